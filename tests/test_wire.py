"""The reference's protobuf wire format (distributions/io/schema.proto) decoded into update_all's SoA arrays
(SURVEY.md §8f rank 4).  Fixtures: messages serialized with the reference's own schema descriptor
(tests/golden/make_golden_wire.py); the source arrays are regenerated from synth.  The decode step needs no
device and runs in the CPU suite; the GPU test loads features through the wire entry."""
import os

import numpy as np
import pytest

import cases
from distributions_b200 import capi, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
IDS = {"dd": capi.DD, "dpd": capi.DPD, "bb": capi.BB, "gp": capi.GP, "nich": capi.NICH, "bnb": capi.BNB, "niw": capi.NIW}


@pytest.fixture(scope="module")
def wire():
    return np.load(os.path.join(ROOT, "tests", "golden", "wire_golden.npz"))


def messages(wire, name):
    lens = wire["%s_group_lens" % name]
    blob = wire["%s_groups" % name].tobytes()
    offs = np.concatenate([[0], np.cumsum(lens)])
    return wire["%s_shared" % name].tobytes(), [blob[offs[i]:offs[i + 1]] for i in range(lens.size)]


def expected(name):
    seed, G, kw = cases.WIRE[name]
    w = getattr(synth, name)(seed, G, 8, **kw)
    f32 = lambda a: np.asarray(a, np.float32).view(np.uint32).ravel()  # noqa: E731
    u32 = lambda a: np.asarray(a).astype(np.uint32).ravel()  # noqa: E731
    if name == "nich":
        return w, np.asarray(w["shared"], np.float32), np.zeros(0, np.uint32), np.concatenate([u32(w["count"]), f32(w["mean"]), f32(w["ctv"])])
    if name == "gp":
        return w, np.asarray(w["shared"], np.float32), np.zeros(0, np.uint32), np.concatenate([u32(w["count"]), u32(w["sum"]), f32(w["log_prod"])])
    if name == "bnb":
        return w, np.asarray(w["shared"], np.float32), np.array([int(w["shared"][2])], np.uint32), np.concatenate([u32(w["count"]), u32(w["sum"])])
    if name == "bb":
        return w, np.asarray(w["shared"], np.float32), np.zeros(0, np.uint32), np.concatenate([u32(w["heads"]), u32(w["tails"])])
    if name == "dd":
        return w, np.asarray(w["alphas"], np.float32), np.zeros(0, np.uint32), u32(w["counts"])
    if name == "niw":
        sh = np.concatenate([[w["kappa"], w["nu"]], w["mu"], w["psi"].ravel()]).astype(np.float32)
        return w, sh, np.zeros(0, np.uint32), np.concatenate([u32(w["count"]), f32(w["sum_x"]), f32(w["sum_xxT"])])
    beta0 = np.float32(max(0.0, 1.0 - float(np.sum(w["betas"].astype(np.float64)))))
    sh = np.concatenate([[np.float32(w["gamma"]), np.float32(w["alpha"]), beta0], w["betas"]]).astype(np.float32)
    return w, sh, w["keys"].astype(np.uint32), u32(w["counts"])


@pytest.mark.parametrize("name", sorted(cases.WIRE))
def test_decode_matches_source_arrays(wire, name):
    sh_msg, g_msgs = messages(wire, name)
    _, sh, keys, stats = expected(name)
    got_sh, got_keys, got_stats = capi.wire_decode(IDS[name], sh_msg, g_msgs)
    assert np.array_equal(got_sh.view(np.uint32), sh.view(np.uint32))
    assert np.array_equal(got_keys, keys)
    assert np.array_equal(got_stats, stats)


@pytest.mark.parametrize("name", sorted(cases.WIRE))
def test_encode_is_the_reference_writers_bytes(wire, name):
    """decode -> encode reproduces the messages the reference's schema serialized, byte for byte (dpd: its
    sparse groups were written in another key order; equal after decoding again)"""
    sh_msg, g_msgs = messages(wire, name)
    _, keys, stats = capi.wire_decode(IDS[name], sh_msg, g_msgs)
    dim = {"dd": 16, "dpd": 40, "niw": 4}.get(name, 0)
    out = capi.wire_encode_groups(IDS[name], len(g_msgs), dim, keys, stats)
    if name == "dpd":
        sh, _, again = capi.wire_decode(IDS[name], sh_msg, out)
        assert np.array_equal(again, stats)
        # Shared: values, betas and the per-value totals (the column sums of the groups' counts, dpd.hpp:126-138)
        totals = stats.reshape(len(g_msgs), dim).sum(axis=0).astype(np.uint32)
        assert capi.wire_encode_shared(IDS[name], sh, np.concatenate([keys, totals])) == sh_msg
        with pytest.raises(ValueError):  # the totals are part of the message
            capi.wire_encode_shared(IDS[name], sh, keys)
    else:
        assert out == g_msgs
        sh, keys, _ = capi.wire_decode(IDS[name], sh_msg, g_msgs)
        assert capi.wire_encode_shared(IDS[name], sh, keys) == sh_msg


def _varint(v):
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        out.append(b | (0x80 if v else 0))
        if not v:
            return bytes(out)


def test_packed_repeated_and_unknown_fields(wire):
    """proto2 writers emit repeated scalars unpacked, proto3-style writers packed: both decode alike;
    unknown fields are skipped"""
    counts = [3, 0, 70000, 1]
    alphas = np.array([0.5, 1.5, 2.0, 0.25], np.float32)
    payload = b"".join(_varint(c) for c in counts)
    packed_group = b"\x0a" + _varint(len(payload)) + payload + b"\x78\x05"          # + unknown varint field 15
    packed_shared = b"\x0a" + _varint(16) + alphas.tobytes() + b"\x7a\x03abc"     # + unknown bytes field 15
    sh, _, st = capi.wire_decode(capi.DD, packed_shared, [packed_group])
    assert np.array_equal(sh, alphas) and np.array_equal(st, np.array(counts, np.uint32))
    unpacked_group = b"".join(b"\x08" + _varint(c) for c in counts)
    _, _, st2 = capi.wire_decode(capi.DD, packed_shared, [unpacked_group])
    assert np.array_equal(st2, st)


def test_malformed_messages_are_rejected(wire):
    sh_msg, g_msgs = messages(wire, "nich")
    with pytest.raises(ValueError):  # truncated fixed32
        capi.wire_decode(capi.NICH, sh_msg[:-2], g_msgs)
    with pytest.raises(ValueError):  # missing required field
        capi.wire_decode(capi.NICH, sh_msg[:5], g_msgs)
    with pytest.raises(ValueError):  # unterminated varint
        capi.wire_decode(capi.NICH, sh_msg, [b"\x08\xff\xff"])
    with pytest.raises(ValueError):  # count that does not fit the reference's 32-bit field
        capi.wire_decode(capi.GP, messages(wire, "gp")[0], [b"\x08" + _varint(1 << 40) + b"\x10\x01\x1d\x00\x00\x00\x00"])
    with pytest.raises(ValueError):  # dd group with the wrong number of counts
        capi.wire_decode(capi.DD, messages(wire, "dd")[0], [b"\x08\x01\x08\x02"])
    with pytest.raises(ValueError):  # dpd group key that the Shared does not list
        capi.wire_decode(capi.DPD, messages(wire, "dpd")[0], [b"\x08" + _varint(0xFFFFFFF0) + b"\x10\x01"])
    with pytest.raises(ValueError):  # length-delimited field running past the end
        capi.wire_decode(capi.DD, b"\x0a\x40\x00\x00", [])


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(n for n in cases.WIRE if n != "niw"))  # niw: tests/test_gpu_niw_stats.py
def test_update_all_wire_equals_update_all(wire, oracle, name):
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    ctx = capi.Context(0)
    try:
        sh_msg, g_msgs = messages(wire, name)
        w = expected(name)[0]
        a = ctx.feature(IDS[name]).update_all(w)
        b = ctx.feature(IDS[name]).update_all_wire(sh_msg, g_msgs)
        rows = {"nich": 4, "gp": 3, "bnb": 3, "bb": 2, "dd": 16, "dpd": 41}[name]
        assert np.array_equal(a.download_caches(rows).view(np.uint32), b.download_caches(rows).view(np.uint32))
        nbytes = 4 * expected(name)[3].size if name != "gp" else 8 * w["count"].size
        assert np.array_equal(a.download_stats(nbytes), b.download_stats(nbytes))
        import struct
        c = ctx.feature(IDS[name]).update_all_stream(sh_msg, b"".join(struct.pack("<I", len(m)) + m for m in g_msgs))
        assert np.array_equal(a.download_caches(rows).view(np.uint32), c.download_caches(rows).view(np.uint32))
        dumped = b.dump_groups_wire()
        if name == "dpd":
            assert np.array_equal(capi.wire_decode(IDS[name], sh_msg, dumped)[2], expected(name)[3])
        else:
            assert dumped == g_msgs
        if name == "gp":  # log_prod travelled too: score_data works straight after the wire load
            assert np.array_equal(a.score_data_grid(np.array([[1.0, 2.0]], np.float32)), b.score_data_grid(np.array([[1.0, 2.0]], np.float32)))
        sizes = w["sizes"]
        assert np.array_equal(ctx.prior_wire_host(wire["clustering_py"].tobytes(), sizes),
                              ctx.prior_pitman_yor_host(synth.PY_ALPHA, synth.PY_D, sizes))
        assert np.array_equal(ctx.prior_wire_host(wire["clustering_le"].tobytes(), sizes), ctx.prior_low_entropy_host(100000, sizes))
    finally:
        ctx.close()


def test_record_stream_split(wire):
    """the reference's dump framing (io/stream.py:141-153): struct.pack('<I', len) + message"""
    import struct
    _, g_msgs = messages(wire, "gp")
    blob = b"".join(struct.pack("<I", len(m)) + m for m in g_msgs)
    assert capi.wire_split_stream(blob) == g_msgs
    assert capi.wire_split_stream(b"") == []
    with pytest.raises(ValueError):
        capi.wire_split_stream(blob[:-1])         # last record cut short
    with pytest.raises(ValueError):
        capi.wire_split_stream(blob + b"\x01\x00")  # dangling length prefix


def test_decoder_survives_garbage(wire):
    """a parser of external bytes must reject, never crash: random strings, truncations and bit flips of valid
    messages for every model (the process surviving this test is the assertion)"""
    rng = np.random.default_rng(2024)
    outcomes = {"ok": 0, "rejected": 0}
    for name in sorted(cases.WIRE):
        sh_msg, g_msgs = messages(wire, name)
        variants = []
        for _ in range(150):
            variants.append((bytes(rng.integers(0, 256, int(rng.integers(0, 40)), dtype=np.uint8)), g_msgs))          # random Shared
            variants.append((sh_msg, [bytes(rng.integers(0, 256, int(rng.integers(0, 40)), dtype=np.uint8))]))       # random Group
            cut = int(rng.integers(0, len(sh_msg) + 1))
            variants.append((sh_msg[:cut], g_msgs))                                                                  # truncated Shared
            g = bytearray(g_msgs[int(rng.integers(0, len(g_msgs)))])
            if g:
                g[int(rng.integers(0, len(g)))] ^= 1 << int(rng.integers(0, 8))                                      # bit flip
            variants.append((sh_msg, [bytes(g)] + g_msgs[1:]))
            s = bytearray(sh_msg)
            s[int(rng.integers(0, len(s)))] ^= 1 << int(rng.integers(0, 8))
            variants.append((bytes(s), g_msgs))
        variants.append((b"", []))
        variants.append((b"\x0a" + b"\xff" * 9 + b"\x7f", g_msgs))   # absurd length prefix
        variants.append((b"\xff" * 64, g_msgs))                        # endless varint
        for s_msg, gs in variants:
            try:
                sh, keys, st = capi.wire_decode(IDS[name], s_msg, gs)
                assert np.all(np.isfinite(sh) | ~np.isfinite(sh))  # touch the outputs
                outcomes["ok"] += 1
            except ValueError:
                outcomes["rejected"] += 1
    assert outcomes["rejected"] > 1000 and outcomes["ok"] > 0


@pytest.mark.parametrize("name", ["nich", "gp", "bnb", "bb", "dd"])
def test_encode_decode_round_trip_random_statistics(wire, name):
    """size-independent property: decode(encode(stats)) == stats for random statistics, including the extremes
    of the 32-bit fields and float bit patterns (inf, -0.0, denormals)"""
    rng = np.random.default_rng(77)
    sh_msg, _ = messages(wire, name)
    G = 257
    specials = np.array([0, 1, 0x7FFFFFFF, 0x80000000, 0xFFFFFFFF, 0x7F800000, 0x00000001], np.uint32)

    def ints(n, signed):
        x = rng.integers(0, 1 << 31 if signed else 1 << 32, n, dtype=np.uint64).astype(np.uint32)
        x[:3] = [0, 1, 0x7FFFFFFF if signed else 0xFFFFFFFF]
        return x

    def floats(n):
        x = rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32)
        x[:specials.size] = specials
        return x

    if name == "nich":
        stats, dim = np.concatenate([ints(G, True), floats(G), floats(G)]), 0
    elif name == "gp":
        stats, dim = np.concatenate([ints(G, False), ints(G, False), floats(G)]), 0
    elif name == "bnb":
        stats, dim = np.concatenate([ints(G, False), ints(G, False)]), 0
    elif name == "bb":
        stats, dim = np.concatenate([ints(G, True), ints(G, True)]), 0
    else:
        stats, dim = ints(G * 16, True), 16
    msgs = capi.wire_encode_groups(IDS[name], G, dim, None, stats)
    _, _, back = capi.wire_decode(IDS[name], sh_msg, msgs)
    assert np.array_equal(back, stats)
