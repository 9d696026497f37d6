"""In-tree build of liboat.so (sm_100a only) with plain nvcc: one object per .cu, linked into a C-ABI shared library.

`python -m oa_transformer_b200.build` or `build_lib()`; the .so lands next to this file so it travels with the
repo snapshot to the GPU box (the history stays source-only: *.so / *.o are git-ignored).
"""
import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_obj")
LIB = os.path.join(HERE, "liboat.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]
if os.environ.get("OAT_GEMM_STAGES"):       # experiment knob: TMA ring depth of the GEMM (default 4)
    NVCC_FLAGS += ["-DOAT_GEMM_STAGES=%d" % int(os.environ["OAT_GEMM_STAGES"])]
if os.environ.get("OAT_SPACE_DBG"):         # debug knob: clock64 timeline of the pipelined space-attention backward
    NVCC_FLAGS += ["-DOAT_SPACE_DBG=%d" % int(os.environ["OAT_SPACE_DBG"])]


def _nvcc():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def _digest(path, extra):
    h = hashlib.sha1()
    h.update(" ".join(NVCC_FLAGS).encode())
    for p in [path] + extra:
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def build_lib(force=False, verbose=False):
    """Compile every csrc/*.cu for sm_100a and link liboat.so. Incremental (content hash per object)."""
    os.makedirs(OBJ, exist_ok=True)
    sources = sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))
    headers = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h")))
    headers.append(os.path.join(HERE, "..", "include", "oat.h"))
    jobs = []
    for src in sources:
        sp = os.path.join(CSRC, src)
        op = os.path.join(OBJ, src[:-3] + ".o")
        stamp = op + ".sha1"
        dg = _digest(sp, headers)
        if not force and os.path.exists(op) and os.path.exists(stamp) and open(stamp).read() == dg:
            continue
        jobs.append((sp, op, stamp, dg))

    def compile_one(job):
        sp, op, stamp, dg = job
        cmd = [_nvcc()] + NVCC_FLAGS + ["-c", sp, "-o", op]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (sp, r.stdout, r.stderr))
        with open(op + ".ptxas.log", "w") as f:
            f.write(r.stderr)
        with open(stamp, "w") as f:
            f.write(dg)
        if verbose:
            sys.stderr.write(r.stderr)
        return sp

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(compile_one, jobs))
    objs = [os.path.join(OBJ, s[:-3] + ".o") for s in sources]
    if jobs or not os.path.exists(LIB) or force:
        cmd = [_nvcc(), "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart_static",
                                                          "-lpthread", "-ldl", "-lrt"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    return LIB


if __name__ == "__main__":
    print(build_lib(force="--force" in sys.argv, verbose="-v" in sys.argv))
