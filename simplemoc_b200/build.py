"""Build recipes for the native parts (all in-tree, so the built files travel to the GPU box).

  libmoc_b200.so   simplemoc_b200/csrc/{moc_device.cu, moc_host.c}  nvcc, sm_100a only
  SimpleMOC-b200   simplemoc_b200/csrc/driver_main.c                the C driver (main loop)
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(ROOT, "include")
LIB = os.path.join(HERE, "libmoc_b200.so")   # experiments: build_lib(out=..., extra_nvcc=[...])
DRIVER = os.path.join(HERE, "SimpleMOC-b200")

NVCC = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
GCC = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _run(cmd, verbose):
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd)


def build_lib(force=False, verbose=False, extra_nvcc=(), out=None):
    srcs = [os.path.join(CSRC, f) for f in
            ("moc_device.cu", "moc_sweep.inl", "moc_phases.inl", "moc_comm.inl", "moc_dropin.inl", "moc_kernels.cuh",
             "moc_walk_warp.cuh", "moc_attenuate.cuh", "moc_two_way.cuh", "moc_two_way.inl", "moc_host.c", "moc_internal.h")]
    srcs += [os.path.join(INCLUDE, f) for f in ("moc_b200.h", "moc_rng.h")]
    out = out or LIB
    if not force and not _newer(out, srcs):
        return out
    obj = os.path.join(CSRC, "moc_host.o")
    _run([GCC, "-std=gnu99", "-O2", "-ffp-contract=off", "-fPIC", "-Wall", "-I", INCLUDE, "-I", CSRC,
          "-c", os.path.join(CSRC, "moc_host.c"), "-o", obj], verbose)
    _run([NVCC, *ARCH, "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-shared",
          "-I", INCLUDE, "-I", CSRC, *extra_nvcc,
          os.path.join(CSRC, "moc_device.cu"), obj, "-o", out, "-ldl"], verbose)
    return out


def build_driver(force=False, verbose=False):
    src = os.path.join(CSRC, "driver_main.c")
    if not os.path.exists(src):
        return None
    if not force and not _newer(DRIVER, [src, LIB]):
        return DRIVER
    _run([GCC, "-std=gnu99", "-O2", "-Wall", "-I", INCLUDE, src, "-o", DRIVER,
          "-L", HERE, "-lmoc_b200", "-Wl,-rpath,$ORIGIN", "-lm"], verbose)
    return DRIVER


def build_all(force=False, verbose=False):
    build_lib(force, verbose)
    build_driver(force, verbose)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose=True)
