"""In-tree build of the native libraries (no pip, no JIT cache: the built .so files travel with the repo
snapshot to the GPU box).

  lifcal_b200/liblfba.so        CUDA kernels + C ABI (include/lfba.h), nvcc, sm_100a only
  lifcal_b200/liblfba_scene.so  synthetic scene generator (include/lfba_scene.h), host only
"""
from __future__ import annotations

import glob
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-O3",
    "--expt-relaxed-constexpr", "-Xptxas", "-v", "--fmad=true",
]


def _host_cxx() -> str:
    # the image exports CXX=/opt/gcc/bin/g++, which lacks libgomp.spec; the distro compiler has it
    return "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else (shutil.which("g++") or "g++")


def _stale(target: str, sources) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources if os.path.exists(s))


def _run(cmd, log=None):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if log:
        with open(log, "w") as f:
            f.write(" ".join(cmd) + "\n" + r.stdout)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("build failed: " + " ".join(cmd))
    return r.stdout


def build_scene(force=False) -> str:
    out = os.path.join(HERE, "liblfba_scene.so")
    src = [os.path.join(HERE, "host", "scene_gen.cpp"), os.path.join(ROOT, "include", "lfba_scene.h"),
           os.path.join(ROOT, "include", "lfba.h")]
    if force or _stale(out, src):
        _run([_host_cxx(), "-O2", "-march=x86-64-v3", "-fopenmp", "-fPIC", "-std=c++17", "-Wall", "-shared",
              "-o", out, src[0]])
    return out


def build_cuda(force=False) -> str:
    """nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo: one object per .cu (compiled in parallel), then a
    shared library. -Xptxas -v output is kept in build_cuda.log (registers / spills per kernel)."""
    from concurrent.futures import ThreadPoolExecutor
    out = os.path.join(HERE, "liblfba.so")
    csrc = os.path.join(HERE, "csrc")
    objdir = os.path.join(HERE, "build")
    cu = sorted(glob.glob(os.path.join(csrc, "*.cu")))
    hdr = glob.glob(os.path.join(csrc, "*.cuh")) + glob.glob(os.path.join(csrc, "*.h")) + \
        [os.path.join(ROOT, "include", "lfba.h")]
    if not cu:
        raise RuntimeError("no CUDA sources under lifcal_b200/csrc")
    os.makedirs(objdir, exist_ok=True)
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    common = [nvcc] + NVCC_FLAGS + ["-ccbin", _host_cxx(), "-I", os.path.join(ROOT, "include")]

    def compile_one(src):
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        if force or _stale(obj, [src] + hdr):
            log = _run(common + ["-c", "-o", obj, src])
            with open(obj + ".log", "w") as f:
                f.write(log)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(cu))) as ex:
        objs = list(ex.map(compile_one, cu))
    if force or _stale(out, objs):
        _run([nvcc, "-shared", "-o", out, "-ccbin", _host_cxx()] + objs + ["-ldl", "-lpthread"])
        with open(os.path.join(HERE, "build_cuda.log"), "w") as f:
            for o in objs:
                if os.path.exists(o + ".log"):
                    f.write(open(o + ".log").read())
    return out


def build_all(force=False):
    return build_scene(force), build_cuda(force)


if __name__ == "__main__":
    print(build_all(force="--force" in sys.argv))
