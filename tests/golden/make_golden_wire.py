#!/usr/bin/env python
"""Generate tests/golden/wire_golden.npz: Shared / Group / Clustering messages of seeded synthetic features
serialized with the REFERENCE's protobuf schema (the FileDescriptorProto embedded in
/root/reference/distributions/io/schema_pb2.py, loaded into a fresh descriptor pool -- the generated module
itself predates the installed protobuf runtime).

Run in the container (needs /root/reference):
    python tests/golden/make_golden_wire.py
The tests regenerate the source arrays from distributions_b200.synth (cases.WIRE) and compare the decoder's
output with them."""
import ast
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cases  # noqa: E402
from distributions_b200 import synth  # noqa: E402


def reference_messages():
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
    src = open("/root/reference/distributions/io/schema_pb2.py").read()
    lit = re.search(r"serialized_pb='((?:[^'\\]|\\.)*)'", src).group(1)
    blob = ast.literal_eval("b'" + lit + "'")
    fdp = descriptor_pb2.FileDescriptorProto.FromString(blob)
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fdp)

    def cls(name):
        return message_factory.GetMessageClass(pool.FindMessageTypeByName("protobuf.distributions." + name))
    return cls


def encode(cls, w):
    m = w["model"]
    G = w["sizes"].size
    if m == "nich":
        sh = cls("NormalInverseChiSq.Shared")(mu=float(w["shared"][0]), kappa=float(w["shared"][1]), sigmasq=float(w["shared"][2]),
                                             nu=float(w["shared"][3]))
        gs = [cls("NormalInverseChiSq.Group")(count=int(w["count"][g]), mean=float(w["mean"][g]),
                                             count_times_variance=float(w["ctv"][g])) for g in range(G)]
    elif m == "gp":
        sh = cls("GammaPoisson.Shared")(alpha=float(w["shared"][0]), inv_beta=float(w["shared"][1]))
        gs = [cls("GammaPoisson.Group")(count=int(w["count"][g]), sum=int(w["sum"][g]), log_prod=float(w["log_prod"][g])) for g in range(G)]
    elif m == "bnb":
        sh = cls("BetaNegativeBinomial.Shared")(alpha=float(w["shared"][0]), beta=float(w["shared"][1]), r=int(w["shared"][2]))
        gs = [cls("BetaNegativeBinomial.Group")(count=int(w["count"][g]), sum=int(w["sum"][g])) for g in range(G)]
    elif m == "bb":
        sh = cls("BetaBernoulli.Shared")(alpha=float(w["shared"][0]), beta=float(w["shared"][1]))
        gs = [cls("BetaBernoulli.Group")(heads=int(w["heads"][g]), tails=int(w["tails"][g])) for g in range(G)]
    elif m == "dd":
        sh = cls("DirichletDiscrete.Shared")(alphas=[float(a) for a in w["alphas"]])
        gs = [cls("DirichletDiscrete.Group")(counts=[int(c) for c in w["counts"][g]]) for g in range(G)]
    elif m == "dpd":
        V = w["keys"].size
        sh = cls("DirichletProcessDiscrete.Shared")(gamma=float(w["gamma"]), alpha=float(w["alpha"]),
                                                   values=[int(k) for k in w["keys"]], betas=[float(b) for b in w["betas"]],
                                                   counts=[int(c) for c in w["counts"].sum(axis=0)])
        gs = []
        for g in range(G):
            nz = np.nonzero(w["counts"][g])[0][::-1]  # sparse, and not in Shared's order
            gs.append(cls("DirichletProcessDiscrete.Group")(keys=[int(w["keys"][v]) for v in nz], values=[int(w["counts"][g][v]) for v in nz]))
        assert V == w["betas"].size
    elif m == "niw":
        sh = cls("NormalInverseWishart.Shared")(mu=[float(x) for x in w["mu"]], kappa=float(w["kappa"]),
                                               psi=[float(x) for x in w["psi"].ravel()], nu=float(w["nu"]))
        gs = [cls("NormalInverseWishart.Group")(count=int(w["count"][g]), sum_x=[float(x) for x in w["sum_x"][g]],
                                               sum_xxT=[float(x) for x in w["sum_xxT"][g].ravel()]) for g in range(G)]
    else:
        raise ValueError(m)
    return sh.SerializeToString(), [g.SerializeToString() for g in gs]


def pack(msgs):
    lens = np.array([len(m) for m in msgs], np.int64)
    return np.frombuffer(b"".join(msgs), np.uint8).copy(), lens


def main():
    cls = reference_messages()
    out = {}
    for name, (seed, G, kw) in cases.WIRE.items():
        w = getattr(synth, name)(seed, G, 8, **kw)
        sh, gs = encode(cls, w)
        out["%s_shared" % name] = np.frombuffer(sh, np.uint8).copy()
        out["%s_groups" % name], out["%s_group_lens" % name] = pack(gs)
    C = cls("Clustering")
    py = C()
    py.pitman_yor.alpha, py.pitman_yor.d = synth.PY_ALPHA, synth.PY_D
    le = C()
    le.low_entropy.dataset_size = 100000
    out["clustering_py"] = np.frombuffer(py.SerializeToString(), np.uint8).copy()
    out["clustering_le"] = np.frombuffer(le.SerializeToString(), np.uint8).copy()
    path = os.path.join(ROOT, "tests", "golden", "wire_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
