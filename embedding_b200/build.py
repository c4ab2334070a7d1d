"""Builds embedding_b200/libdge.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m embedding_b200.build [--force]

nvcc cross-compiles without a GPU.  graph.cu and walk.cu carry the bit-exact double-precision table
arithmetic and are compiled with --fmad=false; the other files use the default contraction.
"""
import concurrent.futures
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libdge.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-Xcompiler", "-pthread", "-I", INCLUDE]
SOURCES = {
    "context.cu": [],
    "graph.cu": ["--fmad=false"],
    "walk.cu": ["--fmad=false"],
    "io.cu": [],
    "sgns.cu": [],
    "comm.cu": [],
    "flows.cu": [],
    "eval.cu": ["--fmad=false"],
}


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _deps():
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(INCLUDE, "dge.h"))
    deps.append(os.path.abspath(__file__))
    return deps


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in _deps())


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    nvcc = _nvcc()

    def compile_one(item):
        src, extra = item
        obj = os.path.join(OBJ, src.replace(".cu", ".o"))
        cmd = [nvcc] + ARCH + COMMON + extra + ["-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    items = [(s, e) for s, e in SOURCES.items() if os.path.exists(os.path.join(CSRC, s))]
    with concurrent.futures.ThreadPoolExecutor(max_workers=len(items)) as ex:
        objs = list(ex.map(compile_one, items))
    cmd = [nvcc] + ARCH + ["-shared", "-o", LIB] + objs + ["-lpthread", "-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    return LIB


# Stand-alone CUDA micro-benchmarks (the measured ceilings bench.py and DESIGN.md quote): scripts/*.cu -> scripts/bin/
MICROBENCH = ["red_microbench", "walk_microbench"]


def build_microbenchmarks(force=False):
    root = os.path.dirname(HERE)
    out_dir = os.path.join(root, "scripts", "bin")
    os.makedirs(out_dir, exist_ok=True)
    outs = []
    for name in MICROBENCH:
        src, out = os.path.join(root, "scripts", name + ".cu"), os.path.join(out_dir, name)
        if not os.path.exists(src):
            continue
        if force or not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src):
            r = subprocess.run([_nvcc()] + ARCH + ["-O3", "-lineinfo", "-std=c++17", "-o", out, src], capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
        outs.append(out)
    return outs


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_microbenchmarks(force="--force" in sys.argv))
