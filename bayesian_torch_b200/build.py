"""Build recipe of libbtb200.so (sm_100a only, in-tree so the .so travels with the repo snapshot).

    python -m bayesian_torch_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU.  One object per .cu (parallel), then one shared library
with a C ABI (include/btb200.h).  `-lineinfo` keeps ncu's source page usable; `-Xptxas -v`
output (registers / spills / smem per kernel) is kept in build/ptxas.log.
"""
import concurrent.futures
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libbtb200.so")
SOURCES = ["bt_api.cu", "bt_kl.cu", "bt_rng.cu", "bt_mc.cu", "bt_pool.cu", "bt_lstm.cu", "bt_im2col.cu", "bt_tma.cu", "bt_fused.cu"]
HEADERS = ["bt_common.cuh", "bt_philox.cuh", "bt_kernels.cuh", "bt_direct.cuh", "bt_tma.cuh", os.path.join("..", "..", "include", "btb200.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
    "--expt-relaxed-constexpr",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (libbtb200 needs CUDA 12.9 nvcc to build for sm_100a)")


def _digest():
    h = hashlib.sha256()
    for f in SOURCES + HEADERS:
        with open(os.path.join(CSRC, f), "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    os.makedirs(BUILD, exist_ok=True)
    stamp = os.path.join(BUILD, "stamp")
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == dig:
        return LIB
    nvcc = _nvcc()
    logs = []

    def compile_one(src):
        obj = os.path.join(BUILD, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        return obj, f"==== {src}\n{r.stderr}"

    with concurrent.futures.ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        results = list(ex.map(compile_one, SOURCES))
    objs = [o for o, _ in results]
    logs = [l for _, l in results]
    with open(os.path.join(BUILD, "ptxas.log"), "w") as f:
        f.write("\n".join(logs))
    cmd = [nvcc, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as f:
        f.write(dig)
    if verbose:
        print("\n".join(logs))
    return LIB


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(p)
