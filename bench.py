#!/usr/bin/env python
"""bench.py -- value-group scores/sec of the batched score_value + sample_from_scores hot path.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload c2_nich]

A "step" is one pass of the hot path (prior + Mixture::score_value for every row x group +
sample_from_scores) over one batch of synthetic rows with frozen group statistics.  At N=1 the
workload is BASELINE.json configs[1]: NormalInverseChiSq, 1M rows x 1024 groups, fp32.  For N>1 rows
are sharded across ranks with no data-path collective (weak scaling: every rank scores its own
1M-row shard).  Prints ONE JSON line on rank 0 (contract in the task statement / DESIGN.md).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "value-group scores/sec (score_value+sample)"
UNIT = "value-group scores/s"

WORKLOADS = {
    # name: (builder, kwargs, description)
    "c1_dd": dict(model="dd", G=100, N=100_000, seed=20241, dim=16),
    "c1_dd_steady": dict(model="dd", G=100, N=50_000_000, seed=20241, dim=16),
    "c2_nich": dict(model="nich", G=1024, N=1_000_000, seed=20242),
    "c3_crosscat": dict(model="crosscat", G=128, N=1_000_000, seed=20243, n_gp=128, n_bb=128),
    "c4_dpd": dict(model="dpd", G=512, N=10_000_000, seed=20244, V=4096),
    "c5_niw": dict(model="niw", G=256, N=1_000_000, seed=20245, d=32),
}


def ncu_traffic(summary):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch, from a committed ncu --set full summary"""
    path = os.path.join(ROOT, "profiles", summary)
    if not os.path.exists(path):
        return None, None
    total = 0.0
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for line in open(path):
        parts = line.split()
        if len(parts) >= 3 and parts[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            total += float(parts[1]) * unit.get(parts[2], 1.0)
    return total, "profiles/" + summary


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs"""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_workload(name, rank=0):
    from distributions_b200 import synth
    cfg = dict(WORKLOADS[name])
    model = cfg.pop("model")
    seed = cfg.pop("seed") + 1000 * rank
    G, N = cfg.pop("G"), cfg.pop("N")
    w = getattr(synth, model)(seed, G, N, **cfg)
    feats = w["features"] if model == "crosscat" else [w]
    return dict(name=name, G=G, N=N, feats=feats, sizes=w["sizes"], u=w["u"])


def model_id(capi, name):
    return {"dd": capi.DD, "dpd": capi.DPD, "bb": capi.BB, "gp": capi.GP, "nich": capi.NICH, "niw": capi.NIW}[name]


# ------------------------------------------------------------------------------------------------
def cpu_reference_arm(wl, seconds_target, threads=None):
    """The reference's own CPU path (oracle/_ref, unmodified reference code) on a bounded sample of
    the workload, all host threads; falls back to the C port when _ref is absent."""
    from distributions_b200 import synth
    from oracle.pyoracle import Oracle, Ref
    threads = threads or os.cpu_count() or 1
    G = wl["G"]
    cols = [w["values"] for w in wl["feats"]]
    F = len(cols)
    if Ref.available():
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import cases
        r = Ref()
        k = r.kind(G, wl["sizes"], synth.PY_ALPHA, synth.PY_D)
        for w in wl["feats"]:
            cases.ref_add_feature(k, w)
        probe = min(wl["N"], max(threads * 64, 2000 // max(F, 1)))
        secs, _ = k.bench(cols, probe, threads)
        rate = probe / max(secs, 1e-6)
        rows = int(min(wl["N"], max(probe, rate * seconds_target)))
        run = lambda: k.bench(cols, rows, threads)[0]  # noqa: E731
        kind = "reference"
    else:
        if F != 1 or wl["feats"][0]["model"] != "nich":
            raise RuntimeError("oracle/_ref is required for the CPU arm of this workload")
        o = Oracle()
        w = wl["feats"][0]
        cache = o.nich_caches(w["shared"], w["count"], w["mean"], w["ctv"])
        prior = o.py_prior(synth.PY_ALPHA, synth.PY_D, wl["sizes"])
        probe = min(wl["N"], 2000)
        secs, _ = o.bench_nich(cache, prior, w["values"][:probe], wl["u"][:probe], threads)
        rate = probe / max(secs, 1e-6)
        rows = int(min(wl["N"], max(probe, rate * seconds_target)))
        run = lambda: o.bench_nich(cache, prior, w["values"][:rows], wl["u"][:rows], threads)[0]  # noqa: E731
        kind = "port"
    return run, rows, kind, threads


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    wl = make_workload(args.workload)
    F = len(wl["feats"])
    run, rows, kind, threads = cpu_reference_arm(wl, seconds_target=3.0)
    for _ in range(args.warmup):
        run()
    t = [run() for _ in range(args.steps)]
    per_step = float(np.mean(t))
    value = rows * F * wl["G"] / per_step
    sample = "first %d of %d rows per step, %d threads (row shards)" % (rows, wl["N"], threads)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": per_step * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "rows_per_step": rows, "groups": wl["G"], "features": F},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    from distributions_b200 import capi, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)

    wl = make_workload(args.workload, rank)  # every rank: its own shard (weak scaling)
    G, N = wl["G"], wl["N"]
    F = len(wl["feats"])
    ctx = capi.Context(local_rank)
    feats = [ctx.feature(model_id(capi, w["model"])).update_all(w) for w in wl["feats"]]
    cols_host = [np.ascontiguousarray(w["values"], dtype=capi.COLUMN_DTYPE[model_id(capi, w["model"])]) for w in wl["feats"]]
    cols = [torch.from_numpy(c).to(dev) for c in cols_host]
    u = torch.from_numpy(wl["u"]).to(dev)
    prior = torch.empty(G, device=dev, dtype=torch.float32)
    ctx.prior_pitman_yor(synth.PY_ALPHA, synth.PY_D, wl["sizes"], prior)
    assign = torch.empty(N, device=dev, dtype=torch.int32)
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev, dtype=torch.float32)  # > 126 MB L2
    stream = torch.cuda.current_stream().cuda_stream

    scores_buf = torch.empty((N, G), device=dev, dtype=torch.float32) if args.materialise else None

    base_sizes = torch.from_numpy(wl["sizes"].astype(np.int32)).to(dev)
    sizes_dev = base_sizes.clone()
    in_groups = [False]

    def step():
        # --sweep: one blocked Gibbs pass over the rows, all on the device -- the rows leave their groups
        # (remove_value), are scored against the rest and resampled, join their new groups (add_value), and
        # the clustering prior is refreshed from the new group sizes
        if args.sweep and in_groups[0]:
            ctx.remove_rows_batch(feats, cols, assign, N, stream=stream)
        ctx.score_sample_batch(feats, cols, N, prior, u, assign, scores_buf, stream=stream)
        if args.sweep:
            ctx.add_rows_batch(feats, cols, assign, N, stream=stream)
            sizes_dev.copy_(base_sizes)
            ctx.count_assignments(assign, N, G, sizes_dev, accumulate=True, stream=stream)
            ctx.prior_pitman_yor_dev(synth.PY_ALPHA, synth.PY_D, G, sizes_dev, prior, stream=stream)
            in_groups[0] = True

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        flush.zero_()
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    t_wall = time.perf_counter()
    for s0, s1 in ev:
        flush.zero_()  # L2 flush between timed iterations (outside the event pair)
        s0.record()
        step()
        s1.record()
    barrier()
    t_wall = time.perf_counter() - t_wall
    clocks = sampler.stop() if rank == 0 else None
    ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = torch.tensor([float(sum(ms))], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(total_ms.item()) / args.steps
    cells_per_step = float(N) * F * G * world
    value = cells_per_step / (ms_per_step * 1e-3)

    # e2e: the host-buffer C-ABI call a reference-side binding makes; H2D + D2H inside the timed region
    # inputs live in pinned host memory (torch pin_memory), the result lands in a pinned host buffer
    e2e_steps = max(2, min(args.steps, 5))
    pin_cols = [torch.from_numpy(c).pin_memory() for c in cols_host]
    pin_u = torch.from_numpy(wl["u"]).pin_memory()
    pin_assign = torch.empty(N, dtype=torch.int32).pin_memory()
    prior_host = prior.cpu().numpy()
    np_cols, np_u, np_assign = [c.numpy() for c in pin_cols], pin_u.numpy(), pin_assign.numpy()
    ctx.score_sample_batch_host(feats, np_cols, prior_host, np_u, assign_out=np_assign)  # warm (staging alloc)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        a_host, _ = ctx.score_sample_batch_host(feats, np_cols, prior_host, np_u, assign_out=np_assign)
    barrier()
    e2e_t = torch.tensor([(time.perf_counter() - t0) / e2e_steps], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    h2d = sum(c.nbytes for c in cols_host) + wl["u"].nbytes + 4 * G
    d2h = 4 * N
    e2e = {"value": cells_per_step / float(e2e_t.item()), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
           "d2h_bytes_per_step": int(d2h)}
    same = None if args.sweep else bool(np.array_equal(a_host, assign.cpu().numpy()))

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    hbm_peak, peak_src = peaks()
    algo_bytes = sum(c.nbytes for c in cols_host) + wl["u"].nbytes + 4 * N  # values + u in, assign out
    if args.materialise:
        algo_bytes += 4 * N * G  # the [N][G] log scores written once
    if wl["feats"][0]["model"] == "dpd":
        algo_bytes += 4 * (4096 + 1) * G  # the cache table once
    achieved = algo_bytes / (ms_per_step * 1e-3) / 1e9
    traffic, traffic_src = (None, None)
    if wl["name"] == "c2_nich" and not args.sweep:
        traffic, traffic_src = ncu_traffic("r01_c2_nich_materialised_v2.txt" if args.materialise else "r01_c2_nich_v4_tile32.txt")
    roofline = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                "note": "the fused kernel moves 12 B per ROW and is bound by the MUFU pipe, not HBM: the fraction that "
                        "measures kernel quality is roofline_binding.frac (measured MUFU peak)",
                "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "kernel": "score_rows_kernel / gather_rows_kernel (fused)",
                "algorithmic_bytes_per_launch": int(algo_bytes)}
    # the binding limit of this kernel is not HBM: report the measured pipe it is bound by as well
    mufu = ctx.pipe_peak(0)
    fma = ctx.pipe_peak(1)
    mufu_per_cell = {"nich": 2.0, "dd": 1.0, "dpd": 1.0, "gp": 1.0, "bb": 1.0, "niw": 2.0}[wl["feats"][0]["model"]]
    if wl["name"] == "c3_crosscat":
        mufu_per_cell = 1.0 / F
    cells_rank = float(N) * F * G
    t_sfu = cells_rank * mufu_per_cell / mufu
    roofline_binding = {"bound": "sfu", "mufu_lane_ops_per_s_measured": mufu, "ffma_lane_ops_per_s_measured": fma,
                        "mufu_per_cell": mufu_per_cell, "t_roof_ms": t_sfu * 1e3, "frac": t_sfu / (ms_per_step * 1e-3)}

    cpu_baseline = None
    if world == 1 and not args.no_cpu:
        try:
            run, rows, kind, threads = cpu_reference_arm(wl, seconds_target=12.0)
            secs = run()
            cpu_baseline = {"value": rows * F * G / secs, "unit": UNIT, "cores": threads, "kind": kind,
                            "sample": "first %d of %d rows, one pass, %d threads (row shards)" % (rows, N, threads)}
        except Exception as exc:  # the baseline is a report, never a reason to lose the GPU number
            cpu_baseline = {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": str(exc)[:200]}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": args.workload, "rows_per_gpu": N, "groups": G, "features": F,
                   "l2": "flushed between timed steps (256 MB memset outside the event pairs)",
                   "mode": ("score+prior+sample with the [N][G] scores also written to HBM" if args.materialise else
                            "fused score+prior+sample, scores not materialised") +
                           (" + batched remove_value / add_value, cache rebuild and prior refresh on the device "
                            "(one blocked Gibbs pass per step)" if args.sweep else ""), "wall_s_timed_region": t_wall,
                   "e2e_matches_device_assign": same},
        "clocks": clocks, "e2e": e2e, "gpu_launches": args.steps * ((1 if F == 1 or wl["name"] == "c3_crosscat" else F) + ((4 * ((F + 127) // 128) + 2) if args.sweep else 0)),
        "roofline": roofline, "roofline_binding": roofline_binding, "cpu_baseline": cpu_baseline,
    }
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


def run_b200_feature_sharded(args):
    """c3 at N > 1: the features of one cross-cat kind are sharded over the ranks; one NCCL
    reduce-scatter(sum) of per-row, per-group partial scores over NVLink, each rank samples its row
    block (distributions_b200.sharding).  Strong scaling: the 1M x 256 x 128 table is fixed."""
    import torch
    import torch.distributed as dist
    from distributions_b200 import capi, sharding, synth

    world = int(os.environ["WORLD_SIZE"])
    rank = int(os.environ["RANK"])
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist.init_process_group("nccl", device_id=dev)
    wl = make_workload(args.workload, 0)  # the same table on every rank; each rank keeps its features
    G, N = wl["G"], wl["N"]
    F = len(wl["feats"])
    mine = sharding.feature_shard(F, rank, world)
    ctx = capi.Context(local_rank)
    feats = [ctx.feature(model_id(capi, wl["feats"][f]["model"])).update_all(wl["feats"][f]) for f in mine]
    cols = [torch.from_numpy(np.ascontiguousarray(wl["feats"][f]["values"],
                                                  dtype=capi.COLUMN_DTYPE[model_id(capi, wl["feats"][f]["model"])])).to(dev)
            for f in mine]
    u = torch.from_numpy(wl["u"]).to(dev)
    prior = torch.empty(G, device=dev, dtype=torch.float32)
    ctx.prior_pitman_yor(synth.PY_ALPHA, synth.PY_D, wl["sizes"], prior)
    stream = torch.cuda.current_stream().cuda_stream
    comm = torch.cuda.Stream(device=dev)
    launches = [0]

    def score_partial(lo, hi, out):
        if feats:
            ctx.score_batch(feats, [c[lo:hi] for c in cols], hi - lo, prior if rank == 0 else None, out, stream=stream)
            launches[0] += 1
        else:
            out.zero_()

    def sample_block(scores, ub, out):
        ctx.sample_from_scores(scores, scores.shape[0], G, ub, out, stream=stream)
        launches[0] += 1

    peer = sharding.PeerFeatureShards(ctx, N, G) if args.shard_mode == "push" else None
    lo_own, hi_own = peer.owned() if peer else (0, 0)
    assign_own = torch.empty(max(hi_own - lo_own, 1), device=dev, dtype=torch.int32)

    def step():
        if peer is not None:  # reduction fused into the score kernel over NVLink peer memory
            launches[0] += 2
            return peer.step(feats, cols, prior, u, assign_own, stream=stream)
        return sharding.feature_sharded_score_sample(score_partial, sample_block, N, G, u, dev, tile_rows=args.tile_rows,
                                                     comm_stream=comm)

    def barrier():
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches[0] = 0
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for s0, s1 in ev:
        s0.record()
        step()
        s1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    total_ms = torch.tensor([float(sum(a.elapsed_time(b) for a, b in ev))], device=dev, dtype=torch.float64)
    dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(total_ms.item()) / args.steps
    if rank == 0:
        value = float(N) * F * G / (ms_per_step * 1e-3)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "rows": N, "groups": G, "features": F,
                       "parallelism": ("feature shards (%d per rank), partial rows pushed into the owner's slot over NVLink "
                                       "peer memory from inside the score kernel, owner samples the slot sum" % len(mine))
                       if peer is not None else
                       ("feature shards (%d per rank) + NCCL reduce-scatter(sum) of [rows][G] partials, "
                        "tiles of %d rows overlapped on a second stream" % (len(mine), args.tile_rows)),
                       "l2": "inputs (640 MB of columns + 512 MB of partial scores per step) exceed the 126 MB L2"},
            "clocks": clocks, "gpu_launches": launches[0],
            "nvlink_bytes_per_step_per_rank": int(4 * N * G * (world - 1) / world),
        }
        print(json.dumps(line))
    dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2_nich", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--materialise", action="store_true", help="also write the [N][G] log scores (HBM-write-bound mode)")
    ap.add_argument("--sweep", action="store_true",
                    help="each step is a blocked Gibbs pass on the device: remove_value, score+sample, add_value, cache/prior refresh")
    ap.add_argument("--tile-rows", type=int, default=65536, help="row tile of the feature-sharded reduce-scatter")
    ap.add_argument("--shard-mode", default="push", choices=["push", "rs"],
                    help="c3 at N>1: fused NVLink peer push (default) or NCCL reduce-scatter")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.workload == "c3_crosscat" and int(os.environ.get("WORLD_SIZE", "1")) > 1:
        return run_b200_feature_sharded(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
