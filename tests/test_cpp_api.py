"""The C++ host mirror (include/distributions_b200/mixture.hpp) behind the reference's Mixture
choreography: compile check on CPU; on the GPU, a scripted sequence of init / add_value / remove_value
/ add_group / remove_group / per-value score_value / batched score_values is executed by the C++
program and replayed with the oracle (reference test template: tests/test_models.py:498-594,
doc/overview.rst:185-202)."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "test_mixture_api.cc")


def _compile(out):
    from distributions_b200 import build
    build.build()
    lib_dir = os.path.join(ROOT, "distributions_b200", "lib")
    cmd = ["g++", "-std=c++14", "-O1", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), SRC, "-o", out,
           "-L" + lib_dir, "-ldist_b200", "-Wl,-rpath," + lib_dir]
    subprocess.run(cmd, check=True)
    return out


def test_cpp_header_compiles_and_links(tmp_path):
    _compile(str(tmp_path / "test_mixture_api"))


class Replay:
    """Python mirror of the scripted choreography; group statistics via the oracle's Group restatements."""

    def __init__(self, model, oracle):
        self.model, self.o = model, oracle
        self.groups, self.counts = [], []

    def empty(self):
        return {"nich": (0, 0.0, 0.0), "gp": (0, 0), "bnb": (0, 0), "bb": [0, 0], "dd": [0] * 16}[self.model]

    def upd(self, st, v, op):
        m = self.model
        if m == "nich":
            return self.o.nich_group_update(op, st[0], st[1], st[2], [v])
        if m in ("gp", "bnb"):
            return (st[0] + op, st[1] + op * int(v))
        if m == "bb":
            st = list(st); st[0 if v else 1] += op; return st
        st = list(st); st[int(v)] += op; return st

    def add(self, gid, v):
        added = self.counts[gid] == 0
        self.counts[gid] += 1
        self.groups[gid] = self.upd(self.groups[gid], v, +1)
        if added:
            self.counts.append(0)
            self.groups.append(self.empty())

    def remove(self, gid, v):
        self.counts[gid] -= 1
        self.groups[gid] = self.upd(self.groups[gid], v, -1)
        if self.counts[gid] == 0:
            self.counts[gid] = self.counts[-1]
            self.groups[gid] = self.groups[-1]
            self.counts.pop()
            self.groups.pop()

    def score_data(self):
        """oracle MixtureDataScorer::score_data of the current groups under the script's Shared:
        (fp32 group-order sum, sum |term|, the same terms summed in double)"""
        m, G = self.model, len(self.groups)
        sizes = np.asarray(self.counts, np.int32)
        if m == "nich":
            st = np.array(self.groups, dtype=np.float64)
            w = dict(model=m, sizes=sizes, shared=np.array([0, 1, 1, 1], np.float32), count=st[:, 0].astype(np.int32),
                     mean=st[:, 1].astype(np.float32), ctv=st[:, 2].astype(np.float32))
        elif m == "bnb":
            st = np.array(self.groups, dtype=np.int64)
            w = dict(model=m, sizes=sizes, shared=np.array([1.0, 1.0, 3], np.float32), count=st[:, 0].astype(np.uint32),
                     sum=st[:, 1].astype(np.uint32))
        elif m == "bb":
            st = np.array(self.groups, dtype=np.int32)
            w = dict(model=m, sizes=sizes, shared=np.array([0.5, 2.0], np.float32), heads=st[:, 0].copy(), tails=st[:, 1].copy())
        else:
            w = dict(model=m, sizes=sizes, alphas=np.full(16, 0.5, np.float32),
                     counts=np.array(self.groups, dtype=np.int32).reshape(G, 16))
        return self.o.score_data(w)

    def scores(self, values):
        import cases
        from oracle.pyoracle import BB, BNB, DD, GP, NICH
        o, m, G = self.o, self.model, len(self.groups)
        prior = o.py_prior(1.0, 0.1, self.counts)
        out = np.tile(prior, (len(values), 1)).astype(np.float32)
        if m == "nich":
            st = np.array(self.groups, dtype=np.float64)
            cache = o.nich_caches([0, 1, 1, 1], st[:, 0].astype(np.int32), st[:, 1].astype(np.float32), st[:, 2].astype(np.float32))
            o.score_rows(NICH, cache, np.asarray(values, np.float32), out)
        elif m == "gp":
            st = np.array(self.groups, dtype=np.int64)
            o.score_rows(GP, o.gp_caches([1, 1], st[:, 0], st[:, 1]), np.asarray(values, np.uint32), out)
        elif m == "bnb":
            st = np.array(self.groups, dtype=np.int64)
            o.score_rows(BNB, o.bnb_caches([1.0, 1.0, 3], st[:, 0], st[:, 1]), np.asarray(values, np.uint32), out)
        elif m == "bb":
            st = np.array(self.groups, dtype=np.int32)
            o.score_rows(BB, o.bb_caches([0.5, 2.0], st[:, 0], st[:, 1]), np.asarray(values, np.uint8), out)
        else:
            st = np.array(self.groups, dtype=np.int32).reshape(G, 16)
            o.score_rows(DD, o.dd_caches(np.full(16, 0.5, np.float32), st), np.asarray(values, np.int32), out)
        return prior, out


def _value(rng, model):
    if model == "nich":
        return float(np.float32(rng.normal(0, 3)))
    if model in ("gp", "bnb"):
        return int(rng.poisson(6))
    if model == "bb":
        return int(rng.random() < 0.4)
    return int(rng.integers(0, 16))


@pytest.mark.gpu
def test_cpp_mixture_choreography(tmp_path, oracle):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import cases
    exe = _compile(str(tmp_path / "test_mixture_api"))
    rng = np.random.default_rng(42)
    lines, replays, expected = [], {}, []
    for model in ("nich", "gp", "bb", "dd", "bnb"):
        rp = Replay(model, oracle)
        lines.append("model %s" % model)
        G0 = 6
        lines.append("groups %d" % G0)
        rp.groups = [rp.empty() for _ in range(G0)]
        rp.counts = [0] * G0
        history = []
        moved = set()
        for gid in range(G0 - 1):  # last group stays empty
            for _ in range(int(rng.integers(2, 9))):
                v = _value(rng, model)
                lines.append("fill %d %r" % (gid, v))
                rp.groups[gid] = rp.upd(rp.groups[gid], v, +1)
                rp.counts[gid] += 1
                history.append((gid, v))
        lines.append("init")

        def score():
            v = _value(rng, model)
            lines.append("score %r" % v)
            expected.append((model, "score", v, rp.scores([v])))

        score()
        for step in range(14):
            if step % 3 == 2 and history:
                gid, v = history.pop(int(rng.integers(0, len(history))))
                # packed ids move on removal: find where that group lives now is not tracked here, so only
                # remove from groups that have never been moved
                if gid < len(rp.counts) and rp.counts[gid] > 0 and gid not in moved:
                    lines.append("remove %d %r" % (gid, v))
                    last = len(rp.counts) - 1
                    will_remove = rp.counts[gid] == 1
                    rp.remove(gid, v)
                    if will_remove:
                        moved.update({gid, last})
            else:
                gid = int(rng.integers(0, len(rp.counts)))
                v = _value(rng, model)
                lines.append("add %d %r" % (gid, v))
                rp.add(gid, v)
                if gid not in moved:
                    history.append((gid, v))
            score()
        n = 40
        vals = [_value(rng, model) for _ in range(n)]
        u = rng.random(n, dtype=np.float32)
        lines.append("batch %d" % n)
        lines.append(" ".join(repr(v) for v in vals))
        lines.append(" ".join(repr(float(x)) for x in u))
        expected.append((model, "batch", (vals, u), rp.scores(vals)))
        # batched add_value into non-empty groups, then the per-value score sees the merged statistics
        n = 60
        vals = [_value(rng, model) for _ in range(n)]
        live = [g for g, c in enumerate(rp.counts) if c > 0]
        gids = [int(rng.choice(live)) for _ in range(n)]
        lines.append("addbatch %d" % n)
        lines.append(" ".join(repr(v) for v in vals))
        lines.append(" ".join(str(g) for g in gids))
        for g, v in zip(gids, vals):
            rp.add(g, v)
        score()
        if model != "gp":  # the mirror's GammaPoisson::Group carries no log_prod
            lines.append("scoredata")
            expected.append((model, "scoredata", None, (None, rp.score_data())))
        lines.append("end")
        replays[model] = rp
    le_sizes = [3, 0, 17, 1, 250, 0, 42]
    lines.append("lowentropy")
    lines.append("5000 %d %s" % (len(le_sizes), " ".join(str(c) for c in le_sizes)))
    lines.append("selfcheck")  # dpd / niw<3> / MixtureIdTracker: incremental group operations == update_all of the same groups
    script = tmp_path / "script.txt"
    script.write_text("\n".join(lines) + "\n")
    out = subprocess.run([exe, str(script)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    rows = [ln.split() for ln in out.stdout.strip().splitlines()]
    it = iter(rows)
    cur = None
    LOG_STEP = 6.2e-5

    def tol(model, want):
        extra = 25 * LOG_STEP if model == "nich" else (6e-5 if model in ("gp", "bnb") else 0.0)
        return 4e-6 * (1 + np.abs(want)) + extra

    for model, kind, payload, (prior, want) in expected:
        row = next(it)
        while row[0] == "model":
            cur = row[1]
            row = next(it)
        assert cur == model
        if kind == "scoredata":
            assert row[0] == "score_data"
            want32, scale, want64 = want
            # nich statistics after the batched add differ from the sequential ones by ~1e-6 relative
            assert abs(float(row[1]) - want64) <= (2e-5 if model == "nich" else 2e-7) * scale + 1e-5, (model, row, want)
            continue
        if kind == "score":
            assert row[0] == "prior"
            np.testing.assert_allclose(np.array(row[1:], np.float32), prior, atol=2e-6)
            row = next(it)
            assert row[0] == "scores"
            got = np.array(row[1:], np.float32)
            assert got.shape == want[0].shape
            assert np.all(np.abs(got - want[0]) <= tol(model, want[0]))
            row = next(it)
            assert row[0] == "group0"
            assert abs(float(row[1]) - (want[0][0] - prior[0])) <= 2 * tol(model, want[0][:1])[0] + 1e-5
        else:
            vals, u = payload
            assert row[0] == "batch_assign"
            assign = np.array(row[1:], np.float64).astype(np.int32)
            row = next(it)
            assert row[0] == "batch_scores"
            got = np.array(row[1:], np.float32).reshape(want.shape)
            assert np.all(np.abs(got - want) <= tol(model, want))
            a_orc = oracle.sample_rows(got.copy(), u)
            assert cases.explained_mismatch(got.astype(np.float64), u, assign, a_orc, 2e-5).all()
            row = next(it)
            assert row == ["batch_matches_per_value", "1"]
    row = next(it)
    assert row[0] == "low_entropy"
    assert np.array_equal(np.array(row[1:], np.float32), oracle.low_entropy_prior(5000, le_sizes))
    assert next(it) == ["selfcheck", "ok"], out.stdout[-2000:]
