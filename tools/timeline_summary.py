"""Summarise the per-pass timelines bench.py --timeline wrote (gpurun_out/<tag>/timeline_n<N>*/timeline_n<N>_rank<r>.json)
into one JSON line per rank: pass duration quantiles, when the halo copies start / are done relative to the pass
start, gaps between passes.  Usage: python tools/timeline_summary.py <dir> [<dir> ...] > profiles/mgpu_timeline_rNN.jsonl"""
import glob
import json
import os
import sys

import numpy as np

for d in sys.argv[1:]:
    for f in sorted(glob.glob(os.path.join(d, "timeline_n*_rank*.json"))):
        t = json.load(open(f))
        p = np.array(t["passes"])
        if p.size == 0:
            continue
        dur, xs, xe = p[:, 3] - p[:, 0], p[:, 1] - p[:, 0], p[:, 2] - p[:, 0]
        gap = p[1:, 0] - p[:-1, 3]
        q = lambda a: {"p50": round(float(np.percentile(a, 50)), 4), "p90": round(float(np.percentile(a, 90)), 4), "max": round(float(a.max()), 4)}
        print(json.dumps({"run": os.path.basename(os.path.normpath(d)), "world": t["world"], "rank": t["rank"], "passes": int(len(p)),
                          "pass_ms": q(dur), "halo_copies_start_ms_after_pass_start": q(xs), "halo_copies_done_ms_after_pass_start": q(xe),
                          "gap_between_passes_ms": q(gap) if len(gap) else None, "rep_ms_this_rank": t.get("rep_ms_this_rank")}))
