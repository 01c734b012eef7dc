"""Build libfolax_b200.so (sm_100a) in-tree with nvcc -- no torch dependency, plain C ABI.

    python -m folax_b200.build [--force] [--verbose]

Each .cu is compiled to an object in folax_b200/lib/obj (in parallel) and linked into
folax_b200/lib/libfolax_b200.so.  The .so is git-ignored but travels to the GPU box.
"""
import concurrent.futures as cf
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(LIBDIR, "obj")
LIB = os.path.join(LIBDIR, "libfolax_b200.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC", "-Xptxas", "-v", "--expt-relaxed-constexpr", "-DFOL_HAVE_J2"]


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _stamp():
    h = hashlib.sha256((" ".join(FLAGS) + repr(EXTRA)).encode())
    for f in sorted(os.listdir(CSRC)):
        with open(os.path.join(CSRC, f), "rb") as fh:
            h.update(f.encode() + fh.read())
    with open(os.path.join(HERE, "..", "include", "folax_b200.h"), "rb") as fh:
        h.update(fh.read())
    return h.hexdigest()


# per-file extras (register caps of the tuned kernels)
EXTRA = {"assemble_hex.cu": ["-maxrregcount=144"] + os.environ.get("FOL_HEX_DEFS", "").split(),
         "energy_grid.cu": os.environ.get("FOL_GRID_DEFS", "").split(),
         "krylov_fused.cu": os.environ.get("FOL_FUSED_DEFS", "").split()}


def _headers_hash():
    h = hashlib.sha256()
    for f in sorted(os.listdir(CSRC)):
        if f.endswith((".cuh", ".h")):
            with open(os.path.join(CSRC, f), "rb") as fh:
                h.update(f.encode() + fh.read())
    with open(os.path.join(HERE, "..", "include", "folax_b200.h"), "rb") as fh:
        h.update(fh.read())
    return h.hexdigest()


def _includes(src):
    """Headers a source file includes (transitively, by name) -- recompile only when they change."""
    seen, todo = set(), [src]
    while todo:
        f = todo.pop()
        path = os.path.join(CSRC, f)
        if not os.path.exists(path):
            continue
        for line in open(path):
            line = line.strip()
            if line.startswith('#include "'):
                name = os.path.basename(line.split('"')[1])
                if name not in seen:
                    seen.add(name)
                    todo.append(name)
    return sorted(seen)


def _file_stamp(src):
    h = hashlib.sha256((" ".join(FLAGS) + repr(EXTRA.get(src, []))).encode())
    for f in [src] + _includes(src):
        path = os.path.join(CSRC, f) if f != "folax_b200.h" else os.path.join(HERE, "..", "include", f)
        if os.path.exists(path):
            with open(path, "rb") as fh:
                h.update(f.encode() + fh.read())
    return h.hexdigest()


def _compile(src, verbose, force=False):
    obj = os.path.join(OBJDIR, src[:-3] + ".o")
    stamp_file, stamp = obj + ".stamp", _file_stamp(src)
    if not force and os.path.exists(obj) and os.path.exists(stamp_file) and open(stamp_file).read() == stamp:
        return obj
    cmd = [NVCC] + FLAGS + EXTRA.get(src, []) + ["-c", os.path.join(CSRC, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    with open(obj + ".log", "w") as fh:
        fh.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
    with open(stamp_file, "w") as fh:
        fh.write(stamp)
    if verbose:
        print(r.stderr)
    return obj


def build(force=False, verbose=False):
    os.makedirs(OBJDIR, exist_ok=True)
    stamp_file = os.path.join(LIBDIR, "build.stamp")
    stamp = _stamp()
    if (not force and os.path.exists(LIB) and os.path.exists(stamp_file)
            and open(stamp_file).read().strip() == stamp):
        return LIB
    srcs = _sources()
    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: _compile(s, verbose, force), srcs))
    cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(stamp_file, "w") as fh:
        fh.write(stamp)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
