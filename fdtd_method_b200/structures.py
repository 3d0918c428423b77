"""Input structures and enums of the reference, restated for Python callers.

Same names, member order and meaning as reference include/Structures.h:8-38, include/Enums.h:4-7 and
include/Constants.h:4-12 (FP = double, include/FP.h:3).
"""
from __future__ import annotations

import ctypes
import enum
from dataclasses import dataclass, field


class Component(enum.IntEnum):  # include/Enums.h:5
    EX = 0
    EY = 1
    EZ = 2
    BX = 3
    BY = 4
    BZ = 5
    JX = 6
    JY = 7
    JZ = 8


class Axis(enum.IntEnum):  # include/Enums.h:6
    X = 0
    Y = 1
    Z = 2


class FDTD_const:  # include/Constants.h:6-11 (the truncated PI is part of the numerical spec)
    C = 3e10
    R = 1e-12
    EPS0 = 1.0
    MU0 = 1.0
    N = 4.0
    PI = 3.14159265358


class Parameters(ctypes.Structure):
    """FDTD_struct::Parameters (include/Structures.h:24-38); layout-compatible with fdtd_params_t."""
    _fields_ = [("Ni", ctypes.c_int), ("Nj", ctypes.c_int), ("Nk", ctypes.c_int),
                ("ax", ctypes.c_double), ("bx", ctypes.c_double),
                ("ay", ctypes.c_double), ("by", ctypes.c_double),
                ("az", ctypes.c_double), ("bz", ctypes.c_double),
                ("dx", ctypes.c_double), ("dy", ctypes.c_double), ("dz", ctypes.c_double)]

    def __repr__(self):
        return "Parameters(" + ", ".join(f"{n}={getattr(self, n)!r}" for n, _ in self._fields_) + ")"


@dataclass
class SelectedFields:  # include/Structures.h:9-12
    selected_E: Component
    selected_B: Component


@dataclass
class CurrentParameters:  # include/Structures.h:14-22
    period: int
    m: int
    dt: float
    iterations: int = 0
    period_x: float = field(default=None)
    period_y: float = field(default=None)
    period_z: float = field(default=None)

    def __post_init__(self):
        for n in ("period_x", "period_y", "period_z"):
            if getattr(self, n) is None:
                setattr(self, n, float(self.m) * FDTD_const.C)
