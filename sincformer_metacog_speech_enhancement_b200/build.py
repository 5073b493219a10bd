"""In-tree build of libsinc_b200.so (hand-written sm_100a CUDA + the C ABI in include/sinc_b200.h).

nvcc cross-compiles without a GPU; the .so is git-ignored but travels to the GPU box with the snapshot.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
REPO_DIR = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libsinc_b200.so")
STAMP_PATH = os.path.join(PKG_DIR, ".libsinc_b200.stamp")
# hardware probes / micro-benchmarks: a separate library that only tests/ and tools/ load (include/sinc_b200_diag.h)
DIAG_DIR = os.path.join(CSRC, "diag")
DIAG_LIB_PATH = os.path.join(PKG_DIR, "libsinc_b200_diag.so")
DIAG_STAMP_PATH = os.path.join(PKG_DIR, ".libsinc_b200_diag.stamp")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "--expt-relaxed-constexpr", "-extended-lambda",
    "-Xcompiler", "-fPIC", "-shared",
]
# experiment builds only (e.g. SINC_EXTRA_NVCC_FLAGS=-DSINC_ABLATION for tools/ablate.py): part of the digest, so a
# library built with extra flags is never mistaken for the product build
NVCC_FLAGS += os.environ.get("SINC_EXTRA_NVCC_FLAGS", "").split()


def sources() -> list[str]:
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def diag_sources() -> list[str]:
    return sorted(os.path.join(DIAG_DIR, f) for f in os.listdir(DIAG_DIR) if f.endswith(".cu")) + \
        [os.path.join(CSRC, "sinc_support.cu")]


def _digest(diag: bool = False) -> str:
    h = hashlib.sha256()
    files = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if os.path.isfile(os.path.join(CSRC, f)))
    if diag:
        files += sorted(os.path.join(DIAG_DIR, f) for f in os.listdir(DIAG_DIR))
    files.append(os.path.join(REPO_DIR, "include", "sinc_b200.h"))
    files.append(os.path.join(REPO_DIR, "include", "sinc_b200_diag.h"))
    for f in files:
        with open(f, "rb") as fh:
            h.update(f.encode())
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def find_nvcc() -> str | None:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    return None


def is_fresh(diag: bool = False) -> bool:
    lib, stamp = (DIAG_LIB_PATH, DIAG_STAMP_PATH) if diag else (LIB_PATH, STAMP_PATH)
    if not (os.path.exists(lib) and os.path.exists(stamp)):
        return False
    with open(stamp) as fh:
        return fh.read().strip() == _digest(diag)


def _compile(srcs, out, extra, verbose):
    """One nvcc -c per source (in parallel, objects cached by content digest under build/), then one link."""
    from concurrent.futures import ThreadPoolExecutor
    nvcc = find_nvcc()
    if nvcc is None:
        raise RuntimeError(f"nvcc not found: cannot build {os.path.basename(out)}")
    objdir = os.path.join(REPO_DIR, "build", os.path.basename(out).replace(".so", ""))
    os.makedirs(objdir, exist_ok=True)
    hdr = hashlib.sha256()
    hdr_files = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".h", ".cuh"))]
    hdr_files += [os.path.join(REPO_DIR, "include", f) for f in sorted(os.listdir(os.path.join(REPO_DIR, "include")))]
    for f in hdr_files:
        with open(f, "rb") as fh:
            hdr.update(fh.read())
    hdr.update(" ".join(NVCC_FLAGS + list(extra)).encode())
    compile_flags = [f for f in NVCC_FLAGS if f != "-shared"]

    def one(src):
        with open(src, "rb") as fh:
            key = hashlib.sha256(hdr.digest() + fh.read()).hexdigest()[:20]
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + "." + key + ".o")
        if os.path.exists(obj):
            return obj, ""
        for old in os.listdir(objdir):
            if old.startswith(os.path.basename(src)[:-3] + "."):
                os.unlink(os.path.join(objdir, old))
        cmd = [nvcc, *compile_flags, *extra, "-I", os.path.join(REPO_DIR, "include"), "-I", CSRC, "-c", "-o", obj, src]
        if verbose:
            cmd[1:1] = ["-Xptxas", "-v"]
        proc = subprocess.run(cmd, capture_output=True, text=True)
        if proc.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + proc.stdout + proc.stderr)
        return obj, proc.stderr

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as pool:
        results = list(pool.map(one, srcs))
    if verbose:
        for _, err in results:
            sys.stderr.write(err)
    cmd = [nvcc, "-shared", "-Xcompiler", "-fPIC", "-gencode", "arch=compute_100a,code=sm_100a", "-o", out,
           *[o for o, _ in results]]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("link failed:\n" + " ".join(cmd) + "\n" + proc.stdout + proc.stderr)


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu under csrc/ (not csrc/diag/) into the product library.  Returns its path."""
    if not force and is_fresh():
        return LIB_PATH
    _compile(sources(), LIB_PATH, [], verbose)
    with open(STAMP_PATH, "w") as fh:
        fh.write(_digest())
    return LIB_PATH


def build_diag_library(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/diag/*.cu into the diagnostics library (probes time out softly and report a status word)."""
    if not force and is_fresh(diag=True):
        return DIAG_LIB_PATH
    _compile(diag_sources(), DIAG_LIB_PATH, ["-DSINC_DIAG_SOFT_TIMEOUT"], verbose)
    with open(DIAG_STAMP_PATH, "w") as fh:
        fh.write(_digest(diag=True))
    return DIAG_LIB_PATH


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_diag_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
