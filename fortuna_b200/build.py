"""In-tree build of libfortuna_b200.so with nvcc for sm_100a (no torch extension machinery: the
library is a plain C-ABI shared object, see include/fortuna_b200.h)."""

import concurrent.futures
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
REPO_DIR = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_DIR = os.path.join(PKG_DIR, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libfortuna_b200.so")
OBJ_DIR = os.path.join(REPO_DIR, "build", "obj")

NVCC_FLAGS = [
    "-gencode",
    "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-O3",
    "-std=c++17",
    "--expt-relaxed-constexpr",
    "-Xcompiler",
    "-fPIC",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libfortuna_b200.so cannot be built")


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps_mtime():
    m = 0.0
    for root in (CSRC, os.path.join(REPO_DIR, "include")):
        for f in os.listdir(root):
            if f.endswith((".cu", ".cuh", ".h")):
                m = max(m, os.path.getmtime(os.path.join(root, f)))
    return m


def _compile_one(nvcc, src, obj, verbose):
    cmd = [nvcc, *NVCC_FLAGS, "-I", os.path.join(REPO_DIR, "include"), "-c", src, "-o", obj]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    return r.stderr


def build(force=False, verbose=False):
    """Compile every .cu under csrc/ for sm_100a and link lib/libfortuna_b200.so. Returns its path."""
    srcs = sources()
    if not force and os.path.exists(LIB_PATH) and os.path.getmtime(LIB_PATH) >= _deps_mtime():
        return LIB_PATH
    nvcc = _nvcc()
    os.makedirs(OBJ_DIR, exist_ok=True)
    os.makedirs(LIB_DIR, exist_ok=True)
    objs = [os.path.join(OBJ_DIR, os.path.basename(s)[:-3] + ".o") for s in srcs]
    logs = []
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        futs = [ex.submit(_compile_one, nvcc, s, o, verbose) for s, o in zip(srcs, objs)]
        for f in futs:
            logs.append(f.result())
    link = [nvcc, "-shared", "-o", LIB_PATH + ".tmp", *objs, "-lcudart", "-lcuda"]
    r = subprocess.run(link, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    os.replace(LIB_PATH + ".tmp", LIB_PATH)
    if verbose:
        sys.stderr.write("\n".join(logs))
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
