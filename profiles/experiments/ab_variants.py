#!/usr/bin/env python
"""Same-box A/B of kernel variants behind dist_b200_ctx_set_option (boxes of the pool differ by a few %, so
variants are compared inside ONE gpurun call).  Device-resident inputs, CUDA events, L2 flushed between steps.

    python profiles/experiments/ab_variants.py c2_nich "packed:" "scalar:6=1"  [--steps 20]
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    import torch
    from distributions_b200 import capi, synth
    name = sys.argv[1]
    steps = 20
    variants = []
    args = sys.argv[2:]
    while args:
        a = args.pop(0)
        if a == "--steps":
            steps = int(args.pop(0))
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
    flush = torch.empty(64 * 1024 * 1024, device="cuda")
    out = {}
    ref = None
    for rnd in range(2):  # two interleaved rounds: drift shows up as disagreement between them
        for label, opts in variants:
            for k, v in opts.items():
                ctx.set_option(k, v)
            for _ in range(3):
                ctx.score_sample_batch(feats, cols, N, prior, u, assign)
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
            torch.cuda.synchronize()
            for a, b in ev:
                flush.zero_()
                a.record()
                ctx.score_sample_batch(feats, cols, N, prior, u, assign)
                b.record()
            torch.cuda.synchronize()
            ms = sorted(a.elapsed_time(b) for a, b in ev)
            got = assign.cpu().numpy()
            if ref is None:
                ref = got
            out.setdefault(label, []).append({"ms_median": ms[len(ms) // 2], "ms_min": ms[0], "agree_with_first": float(np.mean(got == ref))})
            for k in opts:
                ctx.set_option(k, 0)
    cells = float(N) * F * G
    for label, runs in out.items():
        best = min(r["ms_median"] for r in runs)
        print(json.dumps({"workload": name, "variant": label, "runs": runs, "scores_per_s": cells / (best * 1e-3)}))


if __name__ == "__main__":
    main()
