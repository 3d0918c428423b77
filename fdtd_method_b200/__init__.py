"""fdtd_method_b200 -- B200 (sm_100a) implementation of the FDTD_Method Yee leapfrog time step.

The implementation is libfdtd_b200.so (hand-written CUDA behind the C ABI in include/fdtd_b200.h);
this package is the thin host-side mirror of the reference's solver API.  No CPU fallback exists.
"""
from .structures import Axis, Component, CurrentParameters, FDTD_const, Parameters, SelectedFields  # noqa: F401
from .solver import FDTD, FDTD_PML, FieldView, nccl_unique_id, pml_profile, slab_range  # noqa: F401
from .multi import FDTDMulti, FDTD_PML_Multi  # noqa: F401

__all__ = ["FDTD", "FDTD_PML", "FDTDMulti", "FDTD_PML_Multi", "FieldView", "Parameters", "CurrentParameters", "SelectedFields", "Component",
           "Axis", "FDTD_const", "nccl_unique_id", "slab_range", "pml_profile"]
