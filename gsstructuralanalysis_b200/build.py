"""In-tree build of libkl_shell.so for sm_100a (explicit nvcc; no JIT cache, the .so travels with the repo)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libkl_shell.so")
SOURCES = ["kl_capi.cu", "kl_pattern.cu", "kl_assemble.cu", "kl_solve.cu", "kl_stress.cu", "kl_multipatch.cu", "kl_stability.cu", "ks_solid.cu"]
HEADERS = ["kl_internal.h", "kl_device.cuh", os.path.join("..", "..", "include", "kl_shell.h"),
           os.path.join("..", "..", "include", "ks_solid.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC"]


def _nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force=False, verbose=False, out=None, extra=()):
    """out / extra: build a variant (e.g. extra=["-DKS_MINB=2"]) next to the product library for A/B measurements."""
    if out is None and os.environ.get("KL_LIB"):
        return os.environ["KL_LIB"]        # a prebuilt variant was selected explicitly
    if out is None and not force and not needs_build():
        return OUT
    out = out or OUT
    # one object per translation unit, compiled concurrently and reused while neither the source nor a header is newer
    from concurrent.futures import ThreadPoolExecutor
    variant = out != OUT or bool(extra)
    objdir = os.path.join(CSRC, "build", "variant_" + str(abs(hash((out, tuple(extra))))) if variant else "product")
    os.makedirs(objdir, exist_ok=True)
    hdr_t = max(os.path.getmtime(os.path.join(CSRC, h)) for h in HEADERS)
    nvcc = _nvcc()

    def compile_one(f):
        src, obj = os.path.join(CSRC, f), os.path.join(objdir, f + ".o")
        if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(src), hdr_t):
            return obj, ""
        cmd = [nvcc] + NVCC_FLAGS + list(extra) + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("nvcc failed on " + f)
        return obj, r.stderr

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 1)) as ex:
        res = list(ex.map(compile_one, SOURCES))
    r = subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", out] + [o for o, _ in res], capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc link failed")
    if verbose:
        print("".join(log for _, log in res))
    return out


def build_examples(force=False):
    """host-side C++ drivers over the C ABI (g++; they link the in-tree libkl_shell.so)"""
    root = os.path.dirname(HERE)
    exe = os.path.join(root, "examples", "apalm_dispatch")
    src = os.path.join(root, "examples", "apalm_dispatch.cpp")
    deps = [src, os.path.join(root, "examples", "problem_file.h"), os.path.join(root, "include", "gsAPALM_b200.h"),
            os.path.join(root, "include", "kl_shell.h")]
    if force or not os.path.exists(exe) or os.path.getmtime(exe) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-pthread", "-o", exe, src, "-L" + HERE, "-l:libkl_shell.so", "-Wl,-rpath," + HERE])
    return exe


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
