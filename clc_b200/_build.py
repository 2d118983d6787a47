"""In-tree build of libclc_b200.so with nvcc for sm_100a (no torch extension machinery: the
library is a plain C-ABI shared object, see include/clc_b200.h)."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libclc_b200.so")
# bring-up build: the same sources with -DCLC_DEBUG_ABI (clc_debug_* entry points, stage masks, in-kernel
# stamps); loaded only by tests/ and scripts/ through _lib.debug_lib(), never by the product path
LIB_DBG = os.path.join(HERE, "libclc_b200_dbg.so")
OBJ_DIR = os.path.join(HERE, "csrc", "_obj")
OBJ_DIR_DBG = os.path.join(HERE, "csrc", "_obj_dbg")

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr",
]


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hs.append(os.path.join(HERE, "..", "include", "clc_b200.h"))
    return hs


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, debug_too=True):
    """Compile every .cu under csrc/ and link libclc_b200.so (and the bring-up variant).  Returns the
    production library path."""
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    hdrs = _headers()
    variants = [(LIB, OBJ_DIR, [])]
    if debug_too:
        variants.append((LIB_DBG, OBJ_DIR_DBG, ["-DCLC_DEBUG_ABI"]))
    jobs, links = [], []
    for lib, obj_dir, defs in variants:
        os.makedirs(obj_dir, exist_ok=True)
        objs, dirty = [], False
        for src in _sources():
            obj = os.path.join(obj_dir, os.path.basename(src)[:-3] + ".o")
            objs.append(obj)
            if force or _stale(obj, [src] + hdrs):
                jobs.append([nvcc] + NVCC_FLAGS + defs + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj])
                dirty = True
        links.append((lib, objs, dirty))

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed: %s\n%s\n%s" % (" ".join(cmd), r.stdout, r.stderr))
        return r.stderr

    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        logs = list(ex.map(run, jobs))
    if verbose:
        for l in logs:
            sys.stderr.write(l)
    for lib, objs, dirty in links:
        if dirty or force or _stale(lib, objs):
            run([nvcc, "-shared", "-o", lib] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
