"""Compile libdfprior_b200.so in-tree for sm_100a (nvcc cross-compiles without a GPU).

    python -m attentive_dfprior_b200.build [--force]
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
OUT = os.path.join(_HERE, "libdfprior_b200.so")
SOURCES = ["dfp_api.cu", "dfp_dec_fwd.cu", "dfp_dec_bwd_m0.cu", "dfp_dec_bwd_m1.cu", "dfp_dec_bwd_m2.cu", "dfp_attn.cu", "dfp_attn_tc.cu", "dfp_dec_fwd_tc.cu", "dfp_dec_dh_tc.cu", "dfp_dec_dw_tc.cu", "dfp_mapper.cu", "dfp_fusion.cu", "dfp_mesh.cu"]
HEADERS = ["dfp_math.h", "dfp_layout.h", "dfp_device.cuh", "dfp_dec_bwd.inl", "dfp_tc.cuh", "dfp_planes.cuh", "dfp_attn_tc.cuh", "dfp_mesh_tables.h",
           os.path.join("..", "..", "include", "dfprior_b200.h"), os.path.join("..", "..", "include", "dfprior_mesh.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = True) -> str:
    if not force and not needs_build():
        return OUT
    objs = []
    procs = []
    for s in SOURCES:
        obj = os.path.join(CSRC, s.replace(".cu", ".o"))
        objs.append(obj)
        cmd = [_nvcc(), *NVCC_FLAGS, "-c", os.path.join(CSRC, s), "-o", obj]
        if verbose:
            print(" ".join(cmd), flush=True)
        procs.append(subprocess.Popen(cmd))
    for p in procs:
        if p.wait() != 0:
            raise RuntimeError("nvcc failed")
    cmd = [_nvcc(), "-shared", "-Wno-deprecated-gpu-targets", "-o", OUT, *objs, "-lcudart"]
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    return OUT


# ---- test infrastructure: the tcgen05 building-block self-test / descriptor probes (tests/selftest) ----
REPO = os.path.dirname(_HERE)
SELFTEST_SRC = os.path.join(REPO, "tests", "selftest", "dfp_tc_selftest.cu")
SELFTEST_OUT = os.path.join(REPO, "tests", "selftest", "libdfp_selftest.so")


def build_selftest(force: bool = False, verbose: bool = True) -> str:
    """Not part of the product library: compiled separately so that no probe ships in libdfprior_b200.so."""
    deps = [SELFTEST_SRC] + [os.path.join(CSRC, h) for h in HEADERS]
    if not force and os.path.exists(SELFTEST_OUT) and all(os.path.getmtime(d) <= os.path.getmtime(SELFTEST_OUT) for d in deps):
        return SELFTEST_OUT
    cmd = [_nvcc(), *NVCC_FLAGS, "-I", CSRC, "-shared", "-Wno-deprecated-gpu-targets", "-o", SELFTEST_OUT, SELFTEST_SRC, "-lcudart"]
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    return SELFTEST_OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv)
    build_selftest(force="--force" in sys.argv)
    print(OUT)
