"""Turn `ncu --set full` reports into the markdown / json summaries kept under profiles/ (run here, no GPU needed).

    python tools/ncu_summary.py gpurun_out/t2_r01.ncu-rep [more.ncu-rep ...] > profiles/ncu_t2_r01.md
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
]
STALLS = ["barrier", "mio_throttle", "short_scoreboard", "long_scoreboard", "wait", "math_pipe_throttle", "lg_throttle",
          "not_selected", "selected", "branch_resolving", "dispatch_stall", "no_instructions"]


def rows(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(out)))
    return r[0], r[1], r[2:]


def main():
    for path in sys.argv[1:]:
        h, units, data = rows(path)
        ix = {k: i for i, k in enumerate(h)}
        for row in data:
            name = row[ix["Kernel Name"]]
            print(f"## {name}\n\n(from `{path}`)\n\n| metric | value |\n|---|---|")
            for k in KEYS:
                if k in ix:
                    print(f"| {k} [{units[ix[k]]}] | {row[ix[k]]} |")
            tot = 0
            st = {}
            for s_ in STALLS:
                k = f"smsp__pcsamp_warps_issue_stalled_{s_}"
                if k in ix:
                    st[s_] = float(row[ix[k]])
                    tot += st[s_]
            if tot:
                print("\nWarp-state samples (share of all sampled warps): " +
                      ", ".join(f"{k} {v / tot * 100:.1f} %" for k, v in sorted(st.items(), key=lambda x: -x[1])))
            try:
                rd = float(row[ix["dram__bytes_read.sum"]]); wr = float(row[ix["dram__bytes_write.sum"]])
                sc = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
                rd *= sc[units[ix["dram__bytes_read.sum"]]]; wr *= sc[units[ix["dram__bytes_write.sum"]]]
                t = float(row[ix["gpu__time_duration.sum"]]) * {"ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}[units[ix["gpu__time_duration.sum"]]]
                print(f"\nDRAM traffic {(rd + wr) / 1e9:.3f} GB per launch (read {rd / 1e9:.3f}, write {wr / 1e9:.3f}); "
                      f"{(rd + wr) / t / 1e9:.0f} GB/s under ncu ({t * 1e3:.3f} ms)\n")
            except Exception:
                pass


if __name__ == "__main__":
    main()
