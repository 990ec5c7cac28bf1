"""score_data_grid: device (dist_b200_score_data_grid_host) against the unmodified reference's
Mixture::score_data_grid on the host (oracle/_ref), same groups and hyper-parameter grid.
    python profiles/experiments/score_data_grid.py > gpurun_out/score_data_grid.txt
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
from distributions_b200 import capi, synth  # noqa: E402
from oracle.pyoracle import Ref  # noqa: E402


def main():
    ctx = capi.Context(0)
    ref = Ref() if Ref.available() else None
    ids = {"dd": capi.DD, "dpd": capi.DPD, "bb": capi.BB, "gp": capi.GP, "nich": capi.NICH}
    for name, G, kw, n_grid in [("nich", 1024, {}, 1024), ("gp", 1024, {}, 1024), ("dd", 1024, dict(dim=16), 256),
                                ("dpd", 512, dict(V=4096, other_frac=0.02), 64)]:
        w = getattr(synth, name)(1, G, 10, **kw)
        grid = cases.shared_grid(w, n_grid, seed=1)
        f = ctx.feature(ids[name]).update_all(w)
        f.score_data_grid(grid)  # warm
        t0 = time.perf_counter()
        reps = 5
        for _ in range(reps):
            got = f.score_data_grid(grid)
        dev_s = (time.perf_counter() - t0) / reps
        line = "%-5s G=%-5d grid=%-5d device %.3f ms (host buffers in/out)" % (name, G, n_grid, dev_s * 1e3)
        if ref is not None:
            k = ref.kind(G, w["sizes"], synth.PY_ALPHA, synth.PY_D)
            fi = cases.ref_add_feature(k, w)
            t0 = time.perf_counter()
            want = k.score_data_grid(fi, grid, use_grid=True)
            ref_s = time.perf_counter() - t0
            rel = np.max(np.abs(got - want) / (1 + np.abs(want)))
            line += "   reference (1 core) %.1f ms   speed-up %.0fx   max rel diff %.1e" % (ref_s * 1e3, ref_s / dev_s, rel)
        print(line, flush=True)


if __name__ == "__main__":
    main()
