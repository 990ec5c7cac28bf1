#!/usr/bin/env python
"""SM clock and board power while one workload's kernel runs back to back for a few seconds (nvidia-smi sampled
every 100 ms): tells a kernel that is slow from one that runs at a lower clock (power-capped tensor work).

    python profiles/experiments/clocks_under_load.py c5_niw "fused:" "mmaonly:7=4" [--seconds 3]
"""
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    import torch
    from distributions_b200 import capi, synth
    name = sys.argv[1]
    seconds = 3.0
    variants = []
    args = sys.argv[2:]
    while args:
        a = args.pop(0)
        if a == "--seconds":
            seconds = float(args.pop(0))
            continue
        label, _, spec = a.partition(":")
        variants.append((label, {int(k): int(v) for k, v in (kv.split("=") for kv in spec.split(",") if kv)}))
    wl = bench.make_workload(name)
    G, N, F = wl["G"], wl["N"], len(wl["feats"])
    ctx = capi.Context(0)
    feats = [ctx.feature(bench.model_id(capi, w["model"])).update_all(w) for w in wl["feats"]]
    cols = [torch.from_numpy(np.ascontiguousarray(w["values"], dtype=capi.COLUMN_DTYPE[bench.model_id(capi, w["model"])])).cuda() for w in wl["feats"]]
    u = torch.from_numpy(wl["u"]).cuda()
    prior = torch.empty(G, device="cuda")
    ctx.prior_pitman_yor(synth.PY_ALPHA, synth.PY_D, wl["sizes"], prior)
    assign = torch.empty(N, device="cuda", dtype=torch.int32)
    for label, opts in variants:
        for k, v in opts.items():
            ctx.set_option(k, v)
        for _ in range(3):
            ctx.score_sample_batch(feats, cols, N, prior, u, assign)
        torch.cuda.synchronize()
        smi = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_throttle_reasons.active", "--format=csv,noheader,nounits",
                                "-lms", "100", "-i", "0"], stdout=subprocess.PIPE, text=True)
        t0 = time.time()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 0
        a.record()
        while time.time() - t0 < seconds:
            for _ in range(50):
                ctx.score_sample_batch(feats, cols, N, prior, u, assign)
            n += 50
            torch.cuda.synchronize()
        b.record()
        torch.cuda.synchronize()
        smi.terminate()
        rows = [r.split(",") for r in smi.communicate()[0].strip().splitlines()]
        rows = rows[len(rows) // 3:]  # the first third is ramp-up
        mhz = sorted(float(r[0]) for r in rows)
        watts = sorted(float(r[1]) for r in rows)
        print(json.dumps({"workload": name, "variant": label, "ms_per_launch_back_to_back": a.elapsed_time(b) / n, "launches": n,
                          "sm_mhz_median": mhz[len(mhz) // 2], "sm_mhz_min": mhz[0], "power_w_median": watts[len(watts) // 2],
                          "throttle_reasons": sorted({r[2].strip() for r in rows}), "samples": len(rows)}))
        for k in opts:
            ctx.set_option(k, 0)


if __name__ == "__main__":
    main()
