"""In-tree build of libequss_b200.so (hand-written CUDA for sm_100a, C-ABI in include/equss_b200.h).

    python -m equss_b200.build            # or: python <pkg>/build.py

nvcc cross-compiles without a GPU.  Objects are cached per source under csrc/_obj/ and rebuilt when the
source or any header is newer.  The .so is git-ignored but travels to the GPU box with the snapshot."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
OBJ = os.path.join(CSRC, "_obj")
LIB = os.path.join(PKG_DIR, "libequss_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers_mtime():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hs.append(os.path.join(PKG_DIR, "..", "include", "equss_b200.h"))
    return max(os.path.getmtime(h) for h in hs)


def _compile(src, hdr_mtime, verbose):
    s = os.path.join(CSRC, src)
    o = os.path.join(OBJ, src[:-3] + ".o")
    if os.path.exists(o) and os.path.getmtime(o) > max(os.path.getmtime(s), hdr_mtime):
        return o, ""
    cmd = [NVCC, *FLAGS, "-c", s, "-o", o]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    if verbose:
        log = os.path.join(OBJ, src[:-3] + ".ptxas.log")
        with open(log, "w") as f:
            f.write(r.stderr)
    return o, r.stderr


def build(verbose=True, force=False):
    os.makedirs(OBJ, exist_ok=True)
    if force:
        for f in os.listdir(OBJ):
            os.remove(os.path.join(OBJ, f))
    hm = _headers_mtime()
    srcs = _sources()
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        results = list(ex.map(lambda s: _compile(s, hm, verbose), srcs))
    objs = [o for o, _ in results]
    if (not os.path.exists(LIB)) or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs):
        cmd = [NVCC, "-shared", "-o", LIB, *objs]   # static cudart; driver API resolved at run time
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    lib = build(force="--force" in sys.argv)
    print(lib)
