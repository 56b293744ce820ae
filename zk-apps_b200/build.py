"""In-tree build of libb200zk.so (CUDA, sm_100a only) and of the test-side libraries.

nvcc cross-compiles without a GPU.  Objects are rebuilt when their source or any header is
newer; translation units compile in parallel.  The .so files stay in-tree (git-ignored) so
they travel to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_obj")
LIB = os.path.join(HERE, "libb200zk.so")

NVCC_FLAGS = ["-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3",
              "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC"]


def _newest(paths):
    return max(os.path.getmtime(p) for p in paths)


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("command failed: %s\n%s\n%s" % (" ".join(cmd), r.stdout, r.stderr))
    return r.stdout + r.stderr


def build_cuda(verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h", ".hpp"))]
    headers.append(os.path.join(ROOT, "include", "b200zk.h"))
    hdr_time = _newest(headers)
    sources = sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))
    jobs, objs = [], []
    for s in sources:
        src = os.path.join(CSRC, s)
        obj = os.path.join(OBJ, s[:-3] + ".o")
        objs.append(obj)
        if not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), hdr_time):
            jobs.append(["nvcc", *NVCC_FLAGS, "-c", src, "-o", obj] + (["-Xptxas", "-v"] if verbose else []))
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        for out in ex.map(_run, jobs):
            if verbose and out:
                print(out)
    if jobs or not os.path.exists(LIB) or os.path.getmtime(LIB) < _newest(objs):
        _run(["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB, *objs, "-lcudart_static", "-ldl"])
    return LIB


def build_hostcheck() -> str:
    src = os.path.join(ROOT, "tests", "host", "hostcheck.cpp")
    out = os.path.join(ROOT, "tests", "host", "libhostcheck.so")
    deps = [src] + [os.path.join(CSRC, f) for f in ("field.cuh", "ec.cuh", "pairing.cuh", "pairing_consts.cuh", "verify.cuh", "glv.cuh", "field_dfma.cuh", "ec_batch_affine.cuh")]
    if not os.path.exists(out) or os.path.getmtime(out) < _newest(deps):
        _run(["g++", "-O2", "-std=c++17", "-frounding-math", "-shared", "-fPIC", "-x", "c++", src, "-o", out])  # field_dfma.cuh sets the rounding mode
    return out


def build_oracle() -> str | None:
    mk = os.path.join(ROOT, "oracle", "c", "Makefile")
    if not os.path.exists(mk):
        return None
    _run(["make", "-s", "-C", os.path.join(ROOT, "oracle", "c")])
    return os.path.join(ROOT, "oracle", "c", "liboracle.so")


if __name__ == "__main__":
    print(build_cuda(verbose="-v" in sys.argv))
    print(build_hostcheck())
    print(build_oracle())
