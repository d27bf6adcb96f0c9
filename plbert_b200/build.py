"""In-tree build of the sm_100a CUDA library (``plbert_b200/libplbert_b200.so``).

nvcc cross-compiles without a GPU, so this runs in the CPU-only build container; the resulting
``.so`` is git-ignored but travels to the GPU box with the repo snapshot.  Objects are rebuilt only
when a source or header is newer, and independent translation units compile in parallel.
"""
from __future__ import annotations

import concurrent.futures
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
REPO_DIR = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
BUILD = os.path.join(PKG_DIR, "csrc", "build")
LIB = os.path.join(PKG_DIR, "libplbert_b200.so")
SELFTEST = os.path.join(REPO_DIR, "tests", "native", "gemm_selftest")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-diag-suppress", "177",
]


def _nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: the CUDA library cannot be built")
    return nvcc


def _sources() -> list[str]:
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers() -> list[str]:
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hs.append(os.path.join(REPO_DIR, "include", "plbert_b200.h"))
    return hs


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(src: str, obj: str, verbose: bool) -> None:
    cmd = [_nvcc(), *NVCC_FLAGS, "-c", src, "-o", obj]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    if verbose:
        sys.stderr.write(r.stderr)


def build(verbose: bool = False, selftest: bool = True) -> str:
    os.makedirs(BUILD, exist_ok=True)
    headers = _headers()
    jobs = []
    objs = []
    for src in _sources():
        obj = os.path.join(BUILD, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        if _stale(obj, [src, *headers]):
            jobs.append((src, obj))
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        for f in [ex.submit(_compile, s, o, verbose) for s, o in jobs]:
            f.result()
    if _stale(LIB, objs):
        cmd = [_nvcc(), "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    if selftest:
        src = SELFTEST + ".cu"
        if os.path.exists(src) and _stale(SELFTEST, [src, LIB]):
            cmd = [_nvcc(), *NVCC_FLAGS, src, "-o", SELFTEST, "-L" + PKG_DIR, "-lplbert_b200",
                   "-Xlinker", "-rpath=$ORIGIN/../../plbert_b200"]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError(f"selftest build failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv))
