"""In-tree build of the sm_100a extension (``ops/_hfl_C.so``).

nvcc compiles every ``csrc/*.cu`` for ``-gencode arch=compute_100a,code=sm_100a -lineinfo`` (cross-compiles without a
GPU); g++ compiles the pybind/ATen bindings; the shared object stays in the tree (git-ignored, but shipped to the GPU
box by ``gpurun``), so nothing is JIT-compiled on the B200.  Re-run is a no-op while sources and flags are unchanged.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
import sysconfig
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(os.path.dirname(HERE), "csrc")
OUT_SO = os.path.join(HERE, "_hfl_C.so")
OBJ_DIR = os.path.join(HERE, "_build")
STAMP = os.path.join(OBJ_DIR, "stamp.txt")
EXT_NAME = "_hfl_C"

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-lineinfo", "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC",
              "-Xptxas", "-v"] + ARCH_FLAGS + os.environ.get("HFL_NVCC_EXTRA", "").split()


def _sources():
    cu = sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))
    cpp = sorted(f for f in os.listdir(CSRC) if f.endswith(".cpp"))
    hdr = sorted(f for f in os.listdir(CSRC) if f.endswith((".cuh", ".h", ".hpp")))
    return cu, cpp, hdr


def _digest(files, flags) -> str:
    h = hashlib.sha256()
    for f in files:
        with open(os.path.join(CSRC, f), "rb") as fh:
            h.update(f.encode())
            h.update(fh.read())
    h.update(" ".join(flags).encode())
    return h.hexdigest()


def _run(cmd, log):
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    log.append("$ " + " ".join(cmd) + "\n" + p.stdout)
    if p.returncode != 0:
        raise RuntimeError("build failed:\n" + " ".join(cmd) + "\n" + p.stdout)
    return p.stdout


def build(verbose: bool = False, force: bool = False) -> str:
    import torch
    from torch.utils import cpp_extension as ce
    cu, cpp, hdr = _sources()
    inc = ce.include_paths() + [sysconfig.get_paths()["include"], "/usr/local/cuda/include", CSRC]
    abi = "-D_GLIBCXX_USE_CXX11_ABI=%d" % int(torch.compiled_with_cxx11_abi())
    cxx_flags = ["-O2", "-std=c++17", "-fPIC", abi, "-DTORCH_EXTENSION_NAME=" + EXT_NAME,
                 "-DTORCH_API_INCLUDE_EXTENSION_H", "-Wno-deprecated-declarations"]
    os.makedirs(OBJ_DIR, exist_ok=True)
    hdr_digest = _digest(hdr, [])
    stamps = {}
    if os.path.exists(STAMP):
        with open(STAMP) as f:
            for line in f:
                k, _, v = line.strip().partition(" ")
                stamps[k] = v
    log, jobs, objs = [], [], []
    for f in cu:
        obj = os.path.join(OBJ_DIR, f + ".o")
        d = _digest([f], NVCC_FLAGS) + hdr_digest
        objs.append(obj)
        if force or stamps.get(f) != d or not os.path.exists(obj):
            jobs.append((f, d, [NVCC, "-c", os.path.join(CSRC, f), "-o", obj, "-I", CSRC] + NVCC_FLAGS))
    for f in cpp:
        obj = os.path.join(OBJ_DIR, f + ".o")
        d = _digest([f], cxx_flags) + hdr_digest
        objs.append(obj)
        if force or stamps.get(f) != d or not os.path.exists(obj):
            cmd = ["g++", "-c", os.path.join(CSRC, f), "-o", obj] + cxx_flags
            for i in inc:
                cmd += ["-isystem", i]
            jobs.append((f, d, cmd))
    if jobs or not os.path.exists(OUT_SO):
        with ThreadPoolExecutor(max_workers=min(8, max(len(jobs), 1))) as ex:
            outs = list(ex.map(lambda j: _run(j[2], log), jobs))
        for (f, d, _), o in zip(jobs, outs):
            stamps[f] = d
            if verbose:
                print(o)
        tlib = os.path.join(os.path.dirname(torch.__file__), "lib")
        link = ["g++", "-shared", "-o", OUT_SO] + objs + [
            "-L" + tlib, "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch", "-ltorch_python",
            "-L/usr/local/cuda/lib64", "-lcudart", "-Wl,-rpath," + tlib, "-Wl,-rpath,/usr/local/cuda/lib64"]
        _run(link, log)
        with open(STAMP, "w") as f:
            for k, v in stamps.items():
                f.write("%s %s\n" % (k, v))
        with open(os.path.join(OBJ_DIR, "build.log"), "w") as f:
            f.write("\n".join(log))
    return OUT_SO


if __name__ == "__main__":
    path = build(verbose="-v" in sys.argv, force="-f" in sys.argv)
    print(path)
