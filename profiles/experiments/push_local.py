#!/usr/bin/env python
"""Where the feature-sharded c3 step spends its time, measured on ONE GPU: rank 0's share of the work at
feature_shards = 2 / 4 / 8 with every owner's slot placed in LOCAL memory.  score+push = the fused kernel writing the
[rows][G] partials (same instruction stream as the multi-GPU run, stores land in local HBM instead of a peer);
sample = the slot-sum sampler over the owned row block.  The difference between the multi-GPU step
(profiles/r02_bench/scale) and this sum is what NVLink adds.

    python profiles/experiments/push_local.py [--steps 10]
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
    from distributions_b200 import capi, sharding, synth
    steps = int(sys.argv[sys.argv.index("--steps") + 1]) if "--steps" in sys.argv else 10
    wl = bench.make_workload("c3_crosscat")
    G, N, F = wl["G"], wl["N"], len(wl["feats"])
    ctx = capi.Context(0)
    u = torch.from_numpy(wl["u"]).cuda()
    prior = torch.empty(G, device="cuda")
    ctx.prior_pitman_yor(synth.PY_ALPHA, synth.PY_D, wl["sizes"], prior)
    for fs, row_shards in ((2, 1), (4, 1), (8, 1), (2, 4)):
        r0, r1 = sharding.row_shard(N, 0, row_shards)
        n = r1 - r0
        mine = sharding.feature_shard(F, 0, fs)
        feats = [ctx.feature(bench.model_id(capi, wl["feats"][f]["model"])).update_all(wl["feats"][f]) for f in mine]
        cols = [torch.from_numpy(np.ascontiguousarray(wl["feats"][f]["values"][r0:r1], dtype=capi.COLUMN_DTYPE[bench.model_id(capi, wl["feats"][f]["model"])])).cuda()
                for f in mine]
        block = sharding.block_rows(n, fs)
        slots = torch.empty(fs * fs * block * G, device="cuda")  # owner j's buffer = fs slots; this rank writes slot 0 of each
        ptrs = [slots.data_ptr() + 4 * j * fs * block * G for j in range(fs)]
        assign = torch.empty(block, device="cuda", dtype=torch.int32)
        out = {}
        for what in ("score_push", "sample"):
            def run():
                if what == "score_push":
                    ctx.score_push_batch(feats, cols, n, 0, prior, ptrs, block)
                else:
                    ctx.sample_from_slots(slots.data_ptr(), fs, block * G, block, G, u[:block], assign)
            for _ in range(3):
                run()
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
            for a, b in ev:
                a.record()
                run()
                b.record()
            torch.cuda.synchronize()
            ms = sorted(a.elapsed_time(b) for a, b in ev)
            out[what + "_ms"] = ms[len(ms) // 2]
        out.update({"feature_shards": fs, "row_shards": row_shards, "features_on_rank": len(mine), "rows": n, "owned_rows": block,
                    "partials_bytes": 4 * n * G, "would_cross_nvlink_bytes": 4 * n * G * (fs - 1) // fs})
        print(json.dumps(out))


if __name__ == "__main__":
    main()
