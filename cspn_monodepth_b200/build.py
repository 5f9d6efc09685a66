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


TORCH_SO = os.path.join(LIBDIR, "libcspn_torch.so")
TORCH_SRC = os.path.join(CSRC, "torch_ext.cpp")


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def torch_ext_needs_build() -> bool:
    if not os.path.exists(TORCH_SO):
        return True
    t = os.path.getmtime(TORCH_SO)
    return any(os.path.getmtime(d) > t for d in (TORCH_SRC, os.path.join(PKG, "..", "include", "cspn_b200.h")))


def build_torch_ext(force: bool = False) -> str:
    """The PyTorch operator layer (TORCH_LIBRARY(cspn, ...), csrc/torch_ext.cpp): plain C++ over the C ABI, linked against
    libcspn_b200.so (rpath $ORIGIN) and the torch of this interpreter.  Built in-tree so that it travels with the snapshot."""
    if not force and not torch_ext_needs_build():
        return TORCH_SO
    with _BuildLock():
        if not force and not torch_ext_needs_build():
            return TORCH_SO
        return _build_torch_ext_locked()


def _build_torch_ext_locked() -> str:
    import torch
    from torch.utils import cpp_extension as ce
    build(False)
    cxx = shutil.which("g++") or "g++"
    cmd = [cxx, "-O2", "-std=c++17", "-shared", "-fPIC", "-fvisibility=hidden", f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}"]
    for inc in ce.include_paths() + [os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")]:
        cmd += ["-isystem", inc]
    cmd += [TORCH_SRC, "-o", TORCH_SO + f".{os.getpid()}.tmp"]
    for lib_dir in ce.library_paths():
        cmd += ["-L" + lib_dir, "-Wl,-rpath," + lib_dir]
    cmd += ["-L" + LIBDIR, "-Wl,-rpath,$ORIGIN", "-lcspn_b200", "-lc10", "-ltorch_cpu", "-ltorch", "-lc10_cuda", "-ltorch_cuda"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("building libcspn_torch.so failed:\n" + r.stdout[-3000:] + r.stderr[-3000:])
    os.replace(TORCH_SO + f".{os.getpid()}.tmp", TORCH_SO)
    return TORCH_SO


def needs_build() -> bool:
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + [os.path.join(PKG, "..", "include", "cspn_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


class _BuildLock:
    """Inter-process lock (fcntl.flock on lib/.lock): under torchrun every rank may find the library stale at once."""

    def __enter__(self):
        import fcntl
        os.makedirs(LIBDIR, exist_ok=True)
        self.f = open(os.path.join(LIBDIR, ".lock"), "w")
        fcntl.flock(self.f, fcntl.LOCK_EX)
        return self

    def __exit__(self, *exc):
        import fcntl
        fcntl.flock(self.f, fcntl.LOCK_UN)
        self.f.close()
        return False


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return SO
    with _BuildLock():
        if not force and not needs_build():       # another process built it while this one waited
            return SO
        return _build_locked(verbose)


def _build_locked(verbose: bool) -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libcspn_b200.so (and no prebuilt library is present)")
    os.makedirs(LIBDIR, exist_ok=True)
    objs = []
    procs = []
    tmp = os.path.join(LIBDIR, f".build-{os.getpid()}")          # objects and the link result go to a private directory ...
    os.makedirs(tmp, exist_ok=True)
    extra = ["-DCSPN_TRACE"] if os.environ.get("CSPN_TRACE") else []
    if os.environ.get("CSPN_PACKED_SWEEP"):        # A/B: FFMA2 (packed f32x2) sweeps instead of the scalar-FMA default
        extra += ["-DCSPN_PACKED_SWEEP"]
    if os.environ.get("CSPN_FWD_TILE"):            # experiments: "warps x rows", e.g. 12x7
        wv, rw = os.environ["CSPN_FWD_TILE"].split("x")
        extra += [f"-DCSPN_FWD_WARPS={int(wv)}", f"-DCSPN_FWD_ROWS={int(rw)}"]
    for src in sources():
        obj = os.path.join(tmp, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        procs.append((src, subprocess.Popen([nvcc, *NVCC_FLAGS, *extra, "-c", src, "-o", obj], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    for src, p in procs:
        out, _ = p.communicate()
        log.append(out)
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}")
    link = subprocess.run([nvcc, "-shared", "-o", os.path.join(tmp, "libcspn_b200.so"), *objs, "-gencode", "arch=compute_100a,code=sm_100a"], capture_output=True, text=True)
    if link.returncode != 0:
        raise RuntimeError("link failed:\n" + link.stdout + link.stderr)
    for obj in objs:                                             # ... and are moved into place atomically: nobody ever dlopens a half-written file
        os.replace(obj, os.path.join(LIBDIR, os.path.basename(obj)))
    os.replace(os.path.join(tmp, "libcspn_b200.so"), SO)
    shutil.rmtree(tmp, ignore_errors=True)
    with open(os.path.join(LIBDIR, "build.log"), "w") as f:
        f.write("\n".join(log))
    if verbose:
        print("\n".join(log))
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_torch_ext(force="--force" in sys.argv))
