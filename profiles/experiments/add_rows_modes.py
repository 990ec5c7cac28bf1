"""Times dist_b200_add_rows_batch under different assignment distributions.  The committed
add_rows_modes.txt was produced by an experimental build that switched the lane-combining strategy with
DIST_B200_ADD_MODE (0 direct shared-memory atomics, 1 match.any + redux, 2 ballot "peel" of the popular
groups); the shipped kernel uses 0 for the integer statistics and 2 for nich's double sums, and ignores the
variable.  Run on a GPU box:
    python profiles/experiments/add_rows_modes.py > gpurun_out/add_rows_modes.txt
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from distributions_b200 import capi, synth  # noqa: E402


def assignments(kind, G, n, rng):
    if kind == "uniform":
        return rng.integers(0, G, n)
    if kind == "zipf":  # CRP-like: group sizes ~ 1/rank
        p = 1.0 / np.arange(1, G + 1)
        return rng.choice(G, n, p=p / p.sum())
    if kind == "two-heavy":
        return np.where(rng.random(n) < 0.85, np.where(rng.random(n) < 0.7, 3, 5), rng.integers(0, G, n))
    return np.zeros(n, np.int64)  # constant


def main():
    ctx = capi.Context(0)
    rng = np.random.default_rng(0)
    n = 1_000_000
    cases = [("gp", capi.GP, 128, 128), ("bb", capi.BB, 128, 128), ("nich", capi.NICH, 128, 32), ("nich", capi.NICH, 1024, 1)]
    for name, mid, G, F in cases:
        ws = [getattr(synth, name)(100 + i, G, n) for i in range(min(F, 4))]
        feats = [ctx.feature(mid).update_all(ws[i % len(ws)]) for i in range(F)]
        cols = [torch.from_numpy(ws[i % len(ws)]["values"].astype(capi.COLUMN_DTYPE[mid])).cuda() for i in range(F)]
        for dist in ("uniform", "zipf", "two-heavy", "constant"):
            a = torch.from_numpy(assignments(dist, G, n, rng).astype(np.int32)).cuda()
            for mode in (0, 1, 2):
                os.environ["DIST_B200_ADD_MODE"] = str(mode)
                for _ in range(2):
                    ctx.add_rows_batch(feats, cols, a, n)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(5):
                    ctx.add_rows_batch(feats, cols, a, n)
                e1.record()
                torch.cuda.synchronize()
                print("%-5s G=%-5d F=%-4d %-10s mode=%d  %.1f us / batch" % (name, G, F, dist, mode, e0.elapsed_time(e1) / 5 * 1e3), flush=True)


if __name__ == "__main__":
    main()
