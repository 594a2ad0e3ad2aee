"""Build the sm_100a extension in-tree:  python -m nanorlhf_b200.csrc.build

Every ``.cu`` is compiled by nvcc for ``-gencode arch=compute_100a,code=sm_100a -lineinfo`` (kernel
translation units do not include torch headers, so they compile in seconds); ``bindings.cpp`` /
``runtime.cpp`` are compiled by g++ against the torch headers; everything is linked into
``nanorlhf_b200/_C.so`` (git-ignored, but it travels to the GPU box with the gpurun snapshot).
nvcc cross-compiles without a GPU, so this is also the CPU-side "does it build" check.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
import sysconfig
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
OUT = os.path.join(PKG, "_C.so")
OBJ = os.path.join(HERE, "build")
CUDA_HOME = os.environ.get("CUDA_HOME", "/usr/local/cuda")

CU_SOURCES = ["gemm_sm100.cu", "gemm_sm100_2cta.cu", "gemm_tc.cu", "elementwise.cu", "rl_kernels.cu", "sampling.cu", "attention_decode.cu", "attention_decode_fp8.cu", "rope_kv.cu", "attention_fwd_tc.cu", "attention_bwd_tc.cu",
              "attention_varlen.cu", "comm.cu", "quant.cu"]
CPP_SOURCES = ["bindings.cpp", "runtime.cpp"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--use_fast_math",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v"]
if os.environ.get("NRL_ATTN_PROFILE") == "1":       # cycle accounting build of attention_fwd_tc.cu (bench/attn_prof.py)
    NVCC_FLAGS.append("-DNRL_ATTN_PROFILE")


def _stamp(path: str, extra: str) -> str:
    h = hashlib.sha1(extra.encode())
    for f in sorted(os.listdir(HERE)):
        if f.endswith((".h", ".cuh")) or f == os.path.basename(path):
            with open(os.path.join(HERE, f), "rb") as fh:
                h.update(fh.read())
    return h.hexdigest()


def _run(cmd, log):
    r = subprocess.run(cmd, capture_output=True, text=True)
    with open(log, "w") as f:
        f.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError(f"build step failed: {' '.join(cmd)}\n{r.stdout}\n{r.stderr}")


def build(verbose: bool = True, force: bool = False) -> str:
    import torch
    from torch.utils import cpp_extension as ce

    os.makedirs(OBJ, exist_ok=True)
    inc = ce.include_paths(device_type="cuda") if "device_type" in ce.include_paths.__code__.co_varnames else ce.include_paths(cuda=True)
    inc_flags = [f"-I{p}" for p in inc] + [f"-I{sysconfig.get_paths()['include']}", f"-I{CUDA_HOME}/include"]
    abi = int(torch._C._GLIBCXX_USE_CXX11_ABI)
    cxx_flags = ["-O2", "-std=c++17", "-fPIC", f"-D_GLIBCXX_USE_CXX11_ABI={abi}", "-DTORCH_EXTENSION_NAME=_C",
                 "-DTORCH_API_INCLUDE_EXTENSION_H", "-Wno-deprecated-declarations"]
    jobs = []
    objs = []
    for src in CU_SOURCES:
        p = os.path.join(HERE, src)
        if not os.path.exists(p):
            continue
        o = os.path.join(OBJ, src + ".o")
        objs.append(o)
        st = _stamp(p, " ".join(NVCC_FLAGS))
        if force or not os.path.exists(o) or open(o + ".stamp").read() != st if os.path.exists(o + ".stamp") else True:
            jobs.append(([os.path.join(CUDA_HOME, "bin", "nvcc")] + NVCC_FLAGS + ["-c", p, "-o", o], o, st))
    for src in CPP_SOURCES:
        p = os.path.join(HERE, src)
        o = os.path.join(OBJ, src + ".o")
        objs.append(o)
        st = _stamp(p, " ".join(cxx_flags))
        if force or not os.path.exists(o) or open(o + ".stamp").read() != st if os.path.exists(o + ".stamp") else True:
            jobs.append((["g++"] + cxx_flags + inc_flags + ["-c", p, "-o", o], o, st))

    def work(job):
        cmd, o, st = job
        _run(cmd, o + ".log")
        with open(o + ".stamp", "w") as f:
            f.write(st)
        return o

    if jobs and verbose:
        print(f"[build] compiling {len(jobs)} translation unit(s) for sm_100a ...", flush=True)
    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        list(ex.map(work, jobs))
    if jobs or not os.path.exists(OUT):
        libdirs = ce.library_paths(device_type="cuda") if "device_type" in ce.library_paths.__code__.co_varnames else ce.library_paths(cuda=True)
        link = ["g++", "-shared", "-o", OUT] + objs
        for d in libdirs:
            link += [f"-L{d}", f"-Wl,-rpath,{d}"]
        link += ["-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch", "-ltorch_python", "-lcudart"]
        _run(link, os.path.join(OBJ, "link.log"))
        if verbose:
            print(f"[build] linked {OUT}", flush=True)
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv)
