#!/usr/bin/env python
"""In-tree build of the native library: nvcc (sm_100a) for csrc/**/*.cu, g++ for the
torch binding, linked into ``comfyui_parallelanything_b200/ops/_C.so``.

Kernel translation units do not include torch headers, so they compile in seconds;
objects are cached in ``build/`` by source mtime + flags.  The ``.so`` is git-ignored
but NOT gpurun-ignored: it travels to the GPU box with the snapshot.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "csrc")
BUILD = os.path.join(ROOT, "build")
OUT = os.path.join(ROOT, "comfyui_parallelanything_b200", "ops", "_C.so")
CUDA_HOME = os.environ.get("CUDA_HOME", "/usr/local/cuda")
NVCC = os.path.join(CUDA_HOME, "bin", "nvcc")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O3", "-lineinfo",
              "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def _sources():
    cu, cpp = [], []
    for d, _, files in os.walk(CSRC):
        for f in sorted(files):
            p = os.path.join(d, f)
            if f.endswith(".cu"):
                cu.append(p)
            elif f.endswith(".cpp"):
                cpp.append(p)
    return cu, cpp


def _headers_stamp() -> str:
    h = hashlib.sha1()
    for d, _, files in os.walk(CSRC):
        for f in sorted(files):
            if f.endswith((".cuh", ".h", ".hpp")):
                p = os.path.join(d, f)
                h.update(p.encode())
                h.update(str(os.path.getmtime(p)).encode())
    return h.hexdigest()


def _obj_path(src: str) -> str:
    rel = os.path.relpath(src, CSRC).replace(os.sep, "_")
    return os.path.join(BUILD, rel + ".o")


def _needs(src: str, obj: str, stamp: str, flags) -> bool:
    meta = obj + ".meta"
    key = f"{os.path.getmtime(src)}|{stamp}|{' '.join(flags)}"
    if os.path.exists(obj) and os.path.exists(meta) and open(meta).read() == key:
        return False
    return True


def _mark(src: str, obj: str, stamp: str, flags) -> None:
    with open(obj + ".meta", "w") as f:
        f.write(f"{os.path.getmtime(src)}|{stamp}|{' '.join(flags)}")


def _run(cmd, log_path=None):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if log_path:
        with open(log_path, "w") as f:
            f.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("command failed: " + " ".join(cmd))
    return r.stdout + r.stderr


def build(verbose: bool = False) -> str:
    import torch
    from torch.utils import cpp_extension as ce

    os.makedirs(BUILD, exist_ok=True)
    cu, cpp = _sources()
    stamp = _headers_stamp()
    inc = ["-I" + CSRC, "-I" + os.path.join(CUDA_HOME, "include")]
    torch_inc = ["-I" + p for p in ce.include_paths("cuda")]
    import sysconfig
    py_inc = ["-I" + sysconfig.get_paths()["include"]]
    abi = f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}"
    cxx_flags = ["-std=c++17", "-O2", "-fPIC", abi, "-DTORCH_EXTENSION_NAME=_C", "-DTORCH_API_INCLUDE_EXTENSION_H",
                 "-w"]

    jobs = []
    for s in cu:
        o = _obj_path(s)
        if _needs(s, o, stamp, NVCC_FLAGS):
            jobs.append((s, o, [NVCC] + NVCC_FLAGS + inc + ["-c", s, "-o", o], NVCC_FLAGS))
    for s in cpp:
        o = _obj_path(s)
        if _needs(s, o, stamp, cxx_flags):
            jobs.append((s, o, ["g++"] + cxx_flags + inc + torch_inc + py_inc + ["-c", s, "-o", o], cxx_flags))

    def do(job):
        s, o, cmd, flags = job
        out = _run(cmd, o + ".log")
        _mark(s, o, stamp, flags)
        if verbose:
            print(f"[build] {os.path.relpath(s, ROOT)}")
            for line in out.splitlines():
                if "registers" in line or "spill" in line:
                    print("   ", line.strip())
        return o

    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        list(ex.map(do, jobs))

    objs = [_obj_path(s) for s in cu + cpp]
    newest = max([os.path.getmtime(o) for o in objs]) if objs else 0
    if jobs or not os.path.exists(OUT) or os.path.getmtime(OUT) < newest:
        libdirs = ce.library_paths("cuda")
        link = ["g++", "-shared", "-o", OUT] + objs
        for d in libdirs:
            link += ["-L" + d, "-Wl,-rpath," + d]
        link += ["-lc10", "-ltorch", "-ltorch_cpu", "-ltorch_python", "-lc10_cuda", "-ltorch_cuda", "-lcudart"]
        _run(link)
        if verbose:
            print(f"[build] linked {os.path.relpath(OUT, ROOT)}")
    return OUT


if __name__ == "__main__":
    print(build(verbose=True))
