"""Build libfdm_b200.so in-tree with nvcc for sm_100a (no GPU needed to compile)."""

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "lib", "libfdm_b200.so")
SOURCES = ["plan.cpp", "kernels_basic.cu", "goals.cu", "spmm.cu", "solver_pcg.cu", "solver_chol.cu", "solver_chol_warp.cu",
           "solver_pcg2.cu", "solver_ldlt.cu", "pipeline.cu", "fixed_point.cu", "capi.cu"]
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "-shared"]


def _headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    hs.append(os.path.join(os.path.dirname(HERE), "include", "fdm_b200.h"))
    return hs


def _stale(out=None):
    out = out or OUT
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    deps = [os.path.join(CSRC, f) for f in SOURCES] + _headers()
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, defines=(), out=None):
    """Compile every source to an object (in parallel, only the stale ones) and link the shared library.
    `defines` / `out`: an alternative build of the same library (e.g. -DFDM_DEBUG bounds asserts)."""
    out = out or OUT
    if not force and not _stale(out):
        return out
    from concurrent.futures import ThreadPoolExecutor
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    objdir = os.path.join(os.path.dirname(out), "obj" + ("_" + "_".join(defines) if defines else ""))
    os.makedirs(objdir, exist_ok=True)
    hdr_t = max(os.path.getmtime(h) for h in _headers())
    flags = [f for f in FLAGS if f != "-shared"] + ["-D" + d for d in defines]

    def compile_one(src):
        path = os.path.join(CSRC, src)
        obj = os.path.join(objdir, os.path.splitext(src)[0] + ".o")
        if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(path), hdr_t):
            return obj, ""
        cmd = [nvcc] + flags + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, path]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError(f"nvcc failed compiling {src}")
        return obj, r.stderr

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 4)) as ex:
        res = list(ex.map(compile_one, SOURCES))
    r = subprocess.run([nvcc, "-shared", "-o", out] + [o for o, _ in res], capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed linking libfdm_b200.so")
    if verbose:
        print("".join(log for _, log in res))
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
