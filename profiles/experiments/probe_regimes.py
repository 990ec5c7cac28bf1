#!/usr/bin/env python
"""Throughput probes of shapes BASELINE.json does not name (robustness check of the dispatch: no shape should fall
off a cliff).  Device-resident inputs, CUDA events, median of 5 steps.

    python profiles/experiments/probe_regimes.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    import torch
    from distributions_b200 import capi, synth
    ctx = capi.Context(0)
    MID = {"dd": capi.DD, "dpd": capi.DPD, "bb": capi.BB, "gp": capi.GP, "nich": capi.NICH, "niw": capi.NIW, "bnb": capi.BNB}
    probes = [
        ("nich x16 features, G=512", lambda: [synth.nich(100 + k, 512, 200_000) for k in range(16)]),
        ("nich x4 features, G=100", lambda: [synth.nich(120 + k, 100, 500_000) for k in range(4)]),
        ("gp single, G=1024", lambda: [synth.gp(130, 1024, 500_000)]),
        ("gp single, G=100", lambda: [synth.gp(131, 100, 2_000_000)]),
        ("bnb single, G=256", lambda: [synth.bnb(132, 256, 500_000, r=3)]),
        ("niw d=8, G=256", lambda: [synth.niw(133, 256, 500_000, d=8)]),
        ("niw d=3, G=64", lambda: [synth.niw(134, 64, 1_000_000, d=3)]),
        ("nich single, G=128", lambda: [synth.nich(135, 128, 2_000_000)]),
        ("nich single, G=4096", lambda: [synth.nich(136, 4096, 250_000)]),
        ("dd dim=64 single (no shortcut), G=300", lambda: [synth.dd(137, 300, 1_000_000, dim=64)]),
    ]
    for label, make in probes:
        ws = make()
        G, N = ws[0]["sizes"].size, ws[0]["values"].shape[0]
        sizes = ws[0]["sizes"]
        for w in ws[1:]:
            if "count" in w and w["model"] == "nich":
                w["count"] = sizes.astype(w["count"].dtype)
                w["mean"][sizes == 0] = 0
                w["ctv"][sizes == 0] = 0
        feats = [ctx.feature(MID[w["model"]]).update_all(w) for w in ws]
        cols = [torch.from_numpy(np.ascontiguousarray(w["values"], dtype=capi.COLUMN_DTYPE[MID[w["model"]]])).cuda() for w in ws]
        u = torch.from_numpy(ws[0]["u"]).cuda()
        prior = torch.empty(G, device="cuda")
        ctx.prior_pitman_yor(synth.PY_ALPHA, synth.PY_D, sizes, prior)
        assign = torch.empty(N, device="cuda", dtype=torch.int32)
        if "no shortcut" in label:
            ctx.set_option(capi.OPT_VALUE_CDF, 1)
        for _ in range(2):
            ctx.score_sample_batch(feats, cols, N, prior, u, assign)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(5)]
        torch.cuda.synchronize()
        for a, b in ev:
            a.record()
            ctx.score_sample_batch(feats, cols, N, prior, u, assign)
            b.record()
        torch.cuda.synchronize()
        ctx.set_option(capi.OPT_VALUE_CDF, 0)
        ms = sorted(a.elapsed_time(b) for a, b in ev)[2]
        cells = float(N) * len(ws) * G
        print(json.dumps({"probe": label, "rows": N, "ms": round(ms, 4), "scores_per_s": "%.3e" % (cells / (ms * 1e-3))}), flush=True)
        del feats, cols


if __name__ == "__main__":
    main()
