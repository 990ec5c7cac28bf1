"""The compiled drop-in: the REFERENCE'S OWN MixtureSlave<Model, MixtureDataScorer, ValueScorer> instantiated with the
B200 ValueScorers (include/distributions_b200/reference_value_scorers.hpp) and run through the reference's
test_mixture_score choreography next to the stock FastMixture in one binary (tests/cpp/dropin_mixture_slave.cc).
The binary is built where /root/reference exists (oracle/build.py:build_dropin) and travels to the GPU box."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_dropin_compiles_against_reference_headers():
    """container only: the ValueScorer header compiles against the reference's mixture.hpp / nich.hpp / dd.hpp and links"""
    from oracle import build as oracle_build
    if not os.path.isdir(os.path.join(oracle_build.REF, "include", "distributions")):
        pytest.skip("/root/reference absent (GPU box): the prebuilt binary is used")
    path = oracle_build.build_dropin()
    assert path and os.path.exists(path)


@pytest.mark.gpu
def test_dropin_mixture_slave_choreography():
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    exe = os.path.join(ROOT, "oracle", "_ref", "dropin_mixture_slave")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/dropin_mixture_slave not built (needs /root/reference at build time)")
    env = dict(os.environ)
    torch_lib = os.path.join(os.path.dirname(torch.__file__), "lib")
    env["LD_LIBRARY_PATH"] = os.pathsep.join([os.path.join(ROOT, "distributions_b200", "lib"), torch_lib, env.get("LD_LIBRARY_PATH", "")])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=600, env=env)
    sys.stdout.write(out.stdout[-4000:])
    assert out.returncode == 0, out.stdout[-4000:] + out.stderr[-2000:]
    assert "DROPIN OK" in out.stdout
    assert out.stdout.count("LATENCY") == 2 and "BATCH nich" in out.stdout
