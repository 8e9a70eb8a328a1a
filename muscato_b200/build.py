"""In-tree build of the CUDA shared library (sm_100a only) with nvcc.

The library is built next to this file (muscato_b200/libmuscato_b200.so) so that it
travels with the source tree; nothing is installed into site-packages.
"""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(HERE, "libmuscato_b200.so")
EXE_PATH = os.path.join(HERE, "bin", "muscato_b200_hotpath")
GENDAT_PATH = os.path.join(HERE, "libmsc_gendat.so")
STAGE_NAMES = ("muscato_screen", "muscato_confirm", "muscato_combine_windows")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC",
]


def _sources():
    return [os.path.join(CSRC, "api.cu")]


def _deps():
    out = []
    for root, _, files in os.walk(CSRC):
        out += [os.path.join(root, f) for f in files if f.endswith((".cu", ".cuh", ".h", ".inc", ".hpp", ".cc"))]
    out.append(os.path.join(HERE, "..", "include", "muscato_b200.h"))
    return out


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH) or not os.path.exists(EXE_PATH) or not os.path.exists(GENDAT_PATH):
        return True
    t = min(os.path.getmtime(LIB_PATH), os.path.getmtime(EXE_PATH), os.path.getmtime(GENDAT_PATH))
    return any(os.path.getmtime(p) > t for p in _deps() if os.path.exists(p))


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile libmuscato_b200.so for sm_100a; returns its path."""
    if not force and not needs_build():
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libmuscato_b200.so")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH] + _sources()
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    build_host_exe()
    build_gendat()
    return LIB_PATH


def build_gendat() -> str:
    """The synthetic-workload generator (host C++, no CUDA): bench / test tooling in the shape of
    muscato_gendat, kept out of libmuscato_b200.so."""
    cxx = shutil.which("g++") or "g++"
    cmd = [cxx, "-O3", "-std=c++17", "-Wall", "-shared", "-fPIC", "-o", GENDAT_PATH,
           os.path.join(CSRC, "host", "gendat.cc")]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("g++ failed:\n" + res.stdout + res.stderr)
    return GENDAT_PATH


def build_host_exe() -> str:
    """The stage-compatible C++ executable (host text I/O above the C ABI), linked against the library."""
    os.makedirs(os.path.dirname(EXE_PATH), exist_ok=True)
    cxx = shutil.which("g++") or "g++"
    cmd = [cxx, "-O2", "-std=c++17", "-Wall", "-pthread", "-o", EXE_PATH,
           os.path.join(CSRC, "host", "muscato_b200_hotpath.cc"),
           "-L" + HERE, "-lmuscato_b200", "-Wl,-rpath,$ORIGIN/.."]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("g++ failed:\n" + res.stdout + res.stderr)
    # the per-stage names of the unmodified reference driver (cmd/muscato/main.go:310, :402, :442-469):
    # the same binary, dispatching on argv[0]
    for name in STAGE_NAMES:
        link = os.path.join(os.path.dirname(EXE_PATH), name)
        if os.path.lexists(link):
            os.remove(link)
        os.symlink(os.path.basename(EXE_PATH), link)
    return EXE_PATH


if __name__ == "__main__":
    print(build(force=True, verbose=True))
