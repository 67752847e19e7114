"""In-tree build of ``libden_b200.so`` (the C-ABI CUDA library) for sm_100a.

``nvcc`` cross-compiles without a GPU; objects are compiled in parallel, one per
``csrc/*.cu``, and linked into ``lib/libden_b200.so`` next to this file so the
built library travels with the repo snapshot to the GPU box.  A content hash of the
sources + flags is stored beside the library; ``build()`` is a no-op when it matches.
"""

import concurrent.futures
import hashlib
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
REPO_ROOT = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
INCLUDE = os.path.join(REPO_ROOT, "include")
LIB_DIR = os.path.join(PKG_DIR, "lib")
OBJ_DIR = os.path.join(LIB_DIR, "obj")
LIB_PATH = os.path.join(LIB_DIR, "libden_b200.so")
STAMP = os.path.join(LIB_DIR, "libden_b200.stamp")
# test-only probe kernels (tests/csrc/*.cu): built into their own library, never into the product's
PROBE_SRC_DIR = os.path.join(REPO_ROOT, "tests", "csrc")
PROBE_LIB_PATH = os.path.join(LIB_DIR, "libden_b200_probe.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "--expt-relaxed-constexpr",
    "-Xcompiler", "-fPIC",
    "-DDEN_BUILDING_LIBRARY",
]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: cannot build libden_b200.so")
    return exe


def _sources():
    return sorted(
        os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest():
    """Content hash of flags + sources; file names enter RELATIVE to the repository root so the
    snapshot copied to another directory (the GPU box) still matches its stamp."""
    h = hashlib.sha256()
    h.update(" ".join(NVCC_FLAGS).encode())
    files = _sources() + sorted(
        os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h")))
    files += [os.path.join(INCLUDE, f) for f in sorted(os.listdir(INCLUDE))]
    if os.path.isdir(PROBE_SRC_DIR):
        files += sorted(os.path.join(PROBE_SRC_DIR, f) for f in os.listdir(PROBE_SRC_DIR)
                        if f.endswith(".cu"))
    for path in files:
        h.update(os.path.relpath(path, REPO_ROOT).encode())
        with open(path, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def is_current():
    if not (os.path.exists(LIB_PATH) and os.path.exists(STAMP)):
        return False
    with open(STAMP) as fh:
        return fh.read().strip() == _digest()


def have_nvcc():
    return bool(shutil.which("nvcc")) or os.path.exists("/usr/local/cuda/bin/nvcc")


def _compile(src, verbose):
    obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".o")
    cmd = [_nvcc(), *NVCC_FLAGS, "-I", INCLUDE, "-I", CSRC, "-c", src, "-o", obj]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{res.stdout}\n{res.stderr}")
    return obj, res.stderr


def build(force=False, verbose=False):
    """Compile every CUDA source for sm_100a and link the shared library."""
    if not force and is_current():
        return LIB_PATH
    os.makedirs(OBJ_DIR, exist_ok=True)
    srcs = _sources()
    probes = sorted(os.path.join(PROBE_SRC_DIR, f) for f in os.listdir(PROBE_SRC_DIR)
                    if f.endswith(".cu")) if os.path.isdir(PROBE_SRC_DIR) else []
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(12, len(srcs) + len(probes))) as pool:
        results = list(pool.map(lambda s: _compile(s, verbose), srcs + probes))
    if verbose:
        for _, log in results:
            sys.stderr.write(log)
    objs = [o for o, _ in results[:len(srcs)]]
    for out, members in ((LIB_PATH, objs),
                         (PROBE_LIB_PATH, [o for o, _ in results[len(srcs):]]
                          + [os.path.join(OBJ_DIR, "den_api.o")])):
        if out == PROBE_LIB_PATH and not probes:
            continue
        res = subprocess.run([_nvcc(), "-shared", "-o", out, *members], capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    with open(STAMP, "w") as fh:
        fh.write(_digest())
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
