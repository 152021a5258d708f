"""In-tree nvcc build of libpeneo_b200.so (sm_100a only).

    python -m peneo_b200.build [--force] [--verbose]

The library is written next to this file so that it travels to the GPU box with the repository
snapshot.  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libpeneo_b200.so")
GLUE = os.path.join(HERE, "_hostglue" + (sysconfig.get_config_var("EXT_SUFFIX") or ".so"))
STAMP = os.path.join(HERE, "csrc", ".build_stamp")

SOURCES = ["api.cu", "simt_kernels.cu", "gemm_tc.cu", "pair_heads_tc.cu", "pair_heads_tc2.cu", "pair_heads_generic.cu", "pair_bwd_tc.cu", "pair_bwd_tc2.cu", "pair_bwd_elem.cu", "gemm_ds_fused.cu", "gemm_bwd_tc.cu", "gemm_bwd_tc2.cu", "gemm_tf32.cu", "loss.cu", "ohem.cu", "train.cu", "decode.cu", "selftest.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
    "--expt-relaxed-constexpr",
] + os.environ.get("PENEO_NVCC_EXTRA", "").split()  # A/B experiments: e.g. PENEO_NVCC_EXTRA="-DPENEO_T1_EPI_WARPS=16"


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _digest() -> str:
    h = hashlib.sha256()
    for name in sorted(os.listdir(CSRC)):
        if name.endswith((".cu", ".cuh", ".h", ".c")):
            with open(os.path.join(CSRC, name), "rb") as f:
                h.update(name.encode() + b"\0" + f.read())
    with open(os.path.join(os.path.dirname(HERE), "include", "peneo_b200.h"), "rb") as f:
        h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    digest = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(GLUE) and os.path.exists(STAMP):
        with open(STAMP) as f:
            if f.read().strip() == digest:
                return LIB
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(CSRC, src.replace(".cu", ".o"))
        cmd = [_nvcc(), *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    failed = False
    for src, obj, p in procs:
        out, _ = p.communicate()
        log.append(f"==== {src}\n{out}")
        if p.returncode != 0:
            failed = True
        objs.append(obj)
    with open(os.path.join(CSRC, ".build_log.txt"), "w") as f:
        f.write("\n".join(log))
    if failed or verbose:
        sys.stderr.write("\n".join(log) + "\n")
    if failed:
        raise RuntimeError("nvcc failed; see log above")
    cmd = [_nvcc(), "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a"]
    subprocess.run(cmd, check=True)
    # CPython extension with the decode host glue (record blocks -> Python objects)
    cmd = [os.environ.get("CC", "gcc"), "-O2", "-shared", "-fPIC", "-Wall", "-I", sysconfig.get_paths()["include"],
           os.path.join(CSRC, "hostglue.c"), "-o", GLUE]
    subprocess.run(cmd, check=True)
    with open(STAMP, "w") as f:
        f.write(digest)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
