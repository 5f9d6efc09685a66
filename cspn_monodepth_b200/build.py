"""Builds ``cspn_monodepth_b200/lib/libcspn_b200.so`` (the C-ABI CUDA library) in-tree with nvcc.

    python -m cspn_monodepth_b200.build [--force]

sm_100a only; ``-lineinfo`` so ncu source pages map back to the .cu files.  The library does not
link against torch - it is a plain CUDA runtime library loaded through ctypes
(``cspn_monodepth_b200/_lib.py``).
"""
from __future__ import annotations

import glob
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIBDIR = os.path.join(PKG, "lib")
SO = os.path.join(LIBDIR, "libcspn_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
    "-Xptxas", "-v",
    "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def needs_build() -> bool:
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + [os.path.join(PKG, "..", "include", "cspn_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return SO
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libcspn_b200.so (and no prebuilt library is present)")
    os.makedirs(LIBDIR, exist_ok=True)
    objs = []
    procs = []
    extra = ["-DCSPN_TRACE"] if os.environ.get("CSPN_TRACE") else []
    if os.environ.get("CSPN_PACKED_SWEEP"):        # A/B: FFMA2 (packed f32x2) sweeps instead of the scalar-FMA default
        extra += ["-DCSPN_PACKED_SWEEP"]
    if os.environ.get("CSPN_FWD_TILE"):            # experiments: "warps x rows", e.g. 12x7
        wv, rw = os.environ["CSPN_FWD_TILE"].split("x")
        extra += [f"-DCSPN_FWD_WARPS={int(wv)}", f"-DCSPN_FWD_ROWS={int(rw)}"]
    for src in sources():
        obj = os.path.join(LIBDIR, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        procs.append((src, subprocess.Popen([nvcc, *NVCC_FLAGS, *extra, "-c", src, "-o", obj], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    for src, p in procs:
        out, _ = p.communicate()
        log.append(out)
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}")
    link = subprocess.run([nvcc, "-shared", "-o", SO, *objs, "-gencode", "arch=compute_100a,code=sm_100a"], capture_output=True, text=True)
    if link.returncode != 0:
        raise RuntimeError("link failed:\n" + link.stdout + link.stderr)
    with open(os.path.join(LIBDIR, "build.log"), "w") as f:
        f.write("\n".join(log))
    if verbose:
        print("\n".join(log))
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
