#!/usr/bin/env python
"""Build recipe for the test/bench checkers under oracle/.  TEST INFRASTRUCTURE, not product.

  oracle/liboracle.so          plain-C restatement of the reference's hot path (oracle/oracle.c),
                               built everywhere (needs only gcc).
  oracle/_ref/libref_shim.so   the UNMODIFIED reference (sources compiled where they lie under
                               /root/reference, never copied) + oracle/ref_shim.cc, built only in
                               the container where /root/reference exists.  The output directory is
                               git-ignored but not gpurun-ignored, so the binary travels to the GPU
                               box, where /root/reference does not exist.

Flags for the reference follow its own Release configuration (CMakeLists.txt:17-23,35):
-std=c++11 -O3 -msse4.1 -mfpmath=sse -ffast-math -funsafe-math-optimizations -DNDEBUG=1.
Eigen is not installed; oracle/stub/eigen3/Eigen/Cholesky forward-declares what random.hpp:38
names (all uses are in uninstantiated templates), so NIW (models/niw.hpp) is NOT part of _ref.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("DIST_REFERENCE_ROOT", "/root/reference")
REF_OUT = os.path.join(HERE, "_ref")

REF_SOURCES = ["common", "special", "random", "vector_math", "clustering", "models/gp", "models/nich"]
REF_FLAGS = [
    "-std=c++11", "-O3", "-msse4.1", "-mfpmath=sse", "-ffast-math", "-funsafe-math-optimizations",
    "-fPIC", "-Wno-strict-aliasing", "-Wno-cpp", "-Wno-register", "-DNDEBUG=1", "-pthread", "-w",
]


def _run(cmd):
    subprocess.run(cmd, check=True)


def _stale(out, deps):
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build_oracle(force=False):
    src = os.path.join(HERE, "oracle.c")
    hdr = os.path.join(HERE, "oracle_tables.h")
    out = os.path.join(HERE, "liboracle.so")
    if force or _stale(out, [src, hdr, os.path.join(HERE, "oracle.h")]):
        # no -ffast-math: the restatement states the arithmetic order explicitly
        _run(["gcc", "-std=c11", "-O2", "-fPIC", "-shared", "-pthread", "-o", out, src, "-lm"])
    return out


def build_ref(force=False):
    """Returns the path of libref_shim.so, or None when the reference tree is absent and no
    prebuilt binary exists."""
    out = os.path.join(REF_OUT, "libref_shim.so")
    shim = os.path.join(HERE, "ref_shim.cc")
    if not os.path.isdir(os.path.join(REF, "include", "distributions")):
        return out if os.path.exists(out) else None
    if not (force or _stale(out, [shim])):
        return out
    os.makedirs(REF_OUT, exist_ok=True)
    inc = ["-I" + os.path.join(REF, "include"), "-I" + os.path.join(HERE, "stub")]
    objs = []
    for s in REF_SOURCES:
        o = os.path.join(REF_OUT, s.replace("/", "_") + ".o")
        _run(["g++"] + REF_FLAGS + inc + ["-c", os.path.join(REF, "src", s + ".cc"), "-o", o])
        objs.append(o)
    o = os.path.join(REF_OUT, "ref_shim.o")
    _run(["g++"] + REF_FLAGS + inc + ["-c", shim, "-o", o])
    objs.append(o)
    _run(["g++", "-shared", "-pthread", "-o", out] + objs + ["-lm"])
    return out


def build_dropin(force=False):
    """tests/cpp/dropin_mixture_slave.cc: the reference's own MixtureSlave instantiated with the B200 ValueScorers
    (include/distributions_b200/reference_value_scorers.hpp), linked against the reference objects of _ref and
    libdist_b200.so.  Container only; the binary travels to the GPU box (tests/test_dropin.py runs it there)."""
    out = os.path.join(REF_OUT, "dropin_mixture_slave")
    root = os.path.dirname(HERE)
    src = os.path.join(root, "tests", "cpp", "dropin_mixture_slave.cc")
    hdr = os.path.join(root, "include", "distributions_b200", "reference_value_scorers.hpp")
    lib_dir = os.path.join(root, "distributions_b200", "lib")
    if not os.path.isdir(os.path.join(REF, "include", "distributions")):
        return out if os.path.exists(out) else None
    if build_ref() is None or not os.path.exists(os.path.join(lib_dir, "libdist_b200.so")):
        return None
    if not (force or _stale(out, [src, hdr, os.path.join(root, "include", "dist_b200.h")])):
        return out
    inc = ["-I" + os.path.join(REF, "include"), "-I" + os.path.join(HERE, "stub"), "-I" + os.path.join(root, "include")]
    objs = [os.path.join(REF_OUT, s.replace("/", "_") + ".o") for s in REF_SOURCES]
    # DIST_THROW_ON_ERROR: a failed DIST_ASSERT raises instead of abort(), like the reference's Python build (CMakeLists.txt:33)
    _run(["g++"] + REF_FLAGS + ["-DDIST_THROW_ON_ERROR"] + inc + [src] + objs +
         ["-L" + lib_dir, "-ldist_b200", "-Wl,-rpath,$ORIGIN/../../distributions_b200/lib", "-o", out, "-lm"])
    return out


if __name__ == "__main__":
    force = "--force" in sys.argv
    print("oracle :", build_oracle(force))
    print("ref    :", build_ref(force))
    print("dropin :", build_dropin(force))
