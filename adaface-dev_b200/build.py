"""Builds csrc/*.cu into csrc/libadaface_b200.so with nvcc for sm_100a (in-tree; the .so is git-ignored but
travels to the GPU box with the snapshot).  nvcc cross-compiles without a GPU.  Every source is compiled to its own
object (in parallel, only when stale) and the objects are linked into one shared library."""
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
LIB = os.path.join(CSRC, "libadaface_b200.so")
OBJ = os.path.join(CSRC, "_obj")
SOURCES = ["capi.cu", "gemm_tcgen05.cu", "attn_tcgen05.cu", "attn_tcgen05_tri.cu", "attn_mma.cu", "attn_bwd_mma.cu", "attn_bwd_tcgen05.cu", "attn_cross_bwd.cu", "attn_cross_stream.cu",
           "elementwise.cu", "elementwise_bwd.cu", "resblock_elementwise.cu", "sampler.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--use_fast_math",
              "-Xcompiler", "-fPIC", "-diag-suppress", "128"]


def _headers():
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    deps.append(os.path.join(os.path.dirname(CSRC), "..", "include", "adaface_b200.h"))
    return deps


def _compile(nvcc, src, verbose):
    obj = os.path.join(OBJ, src.replace(".cu", ".o"))
    deps = [os.path.join(CSRC, src)] + _headers()
    if os.path.exists(obj) and all(os.path.getmtime(d) <= os.path.getmtime(obj) for d in deps):
        return obj, ""
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src]
    res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed on {src}:\n" + res.stdout + res.stderr)
    return obj, res.stderr


def build(force=False, verbose=False):
    """Compile every CUDA source for sm_100a.  Returns the path of the shared library."""
    if not force and os.path.exists(LIB):
        # fast path (e.g. on the GPU box, where the library arrives prebuilt and the object directory does not travel):
        # nothing to do when the library is newer than every source and header
        deps = [os.path.join(CSRC, f) for f in SOURCES] + _headers()
        if all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in deps):
            return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(OBJ, exist_ok=True)
    if force:
        for f in os.listdir(OBJ):
            os.remove(os.path.join(OBJ, f))
    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        results = list(ex.map(lambda s: _compile(nvcc, s, verbose), SOURCES))
    objs = [o for o, _ in results]
    if verbose:
        for (_, log), src in zip(results, SOURCES):
            if log:
                print(f"==== {src}\n{log}")
    if force or not os.path.exists(LIB) or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs):
        res = subprocess.run([nvcc, "-shared", "-o", LIB] + objs + ["-lcudart"], cwd=CSRC, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("link failed:\n" + res.stdout + res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
