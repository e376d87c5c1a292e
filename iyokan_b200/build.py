"""In-tree build of the CUDA library and the CPU-side test helpers (explicit nvcc / g++ calls)."""
from __future__ import annotations

import os
import shutil
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
CSRC = ROOT / "iyokan_b200" / "csrc"
SIM = ROOT / "tests" / "sim"

NVCC_FLAGS = [
    "-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "--expt-relaxed-constexpr", "-shared", "-Xcompiler", "-fPIC",
]


def _newer(target: Path, sources) -> bool:
    if not target.exists():
        return False
    t = target.stat().st_mtime
    return all(Path(s).stat().st_mtime <= t for s in sources)


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA extension cannot be built (there is no CPU fallback)")


# Two compile-time flavours of every native library, as the reference has them (-DIYOKAN_80BIT_SECURITY, CMakeLists.txt:28-30):
# "" = 128-bit parameters (libb200fhe.so), "80" = 80-bit parameters (libb200fhe80.so, -DB200FHE_80BIT).
FLAVOURS = ("", "80")


def _defs(flavour: str) -> list:
    return ["-DB200FHE_80BIT"] if flavour == "80" else []


def build_cuda(force: bool = False, verbose: bool = False, flavour: str | None = None) -> Path:
    """Compiles the CUDA library; flavour None = both, returns the path of the 128-bit one (or of the one asked for)."""
    outs = {}
    for fl in (FLAVOURS if flavour is None else (flavour,)):
        out = CSRC / f"libb200fhe{fl}.so"
        outs[fl] = out
        srcs = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.h")) + [ROOT / "include" / "b200fhe.h"]
        if not force and _newer(out, srcs):
            continue
        cmd = [_nvcc(), *NVCC_FLAGS, *_defs(fl), *(["-Xptxas", "-v"] if verbose else []), "-o", str(out),
               str(CSRC / "b200fhe.cu"), "-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
        if verbose:
            print(r.stderr)
    return outs.get("" if flavour is None else flavour)


def build_sim(force: bool = False, flavour: str | None = None) -> Path:
    """CPU lock-step simulator of the kernels (test infrastructure, g++), one library per flavour."""
    outs = {}
    for fl in (FLAVOURS if flavour is None else (flavour,)):
        out = SIM / f"libbr_sim{fl}.so"
        outs[fl] = out
        srcs = [SIM / "br_sim.cpp"] + list(CSRC.glob("*.h"))
        if not force and _newer(out, srcs):
            continue
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        cmd = [cxx, "-std=c++17", "-O2", "-fopenmp", "-fPIC", "-shared", *_defs(fl), "-o", str(out), str(SIM / "br_sim.cpp")]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("g++ failed:\n" + r.stdout + r.stderr)
    return outs.get("" if flavour is None else flavour)


def build_host(force: bool = False) -> Path:
    """Host netlist engine (C++, g++), linked against the CUDA library next to it."""
    host = ROOT / "iyokan_b200" / "host"
    first = None
    for fl in FLAVOURS:
        out = host / f"libb200net{fl}.so"
        first = first or out
        srcs = [host / "b200net.cpp", ROOT / "include" / "b200net.h", ROOT / "include" / "b200fhe.h"]
        if not force and _newer(out, srcs) and out.stat().st_mtime >= (CSRC / f"libb200fhe{fl}.so").stat().st_mtime:
            continue
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        cmd = [cxx, "-std=c++17", "-O2", "-fPIC", "-shared", "-Wall", *_defs(fl), "-o", str(out), str(host / "b200net.cpp"),
               f"-L{CSRC}", f"-lb200fhe{fl}", "-Wl,-rpath,$ORIGIN/../csrc"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("g++ failed:\n" + r.stdout + r.stderr)
    return first


def build_microbench(force: bool = False) -> list:
    """Stand-alone micro-benchmarks (scripts/microbench/*.cu -> binaries next to the sources; run by scripts/gpu_round.sh)."""
    outs = []
    mb = ROOT / "scripts" / "microbench"
    for src in sorted(mb.glob("*.cu")):
        out = src.with_suffix("")
        outs.append(out)
        if not force and _newer(out, [src] + list(CSRC.glob("*.h"))):
            continue
        cmd = [_nvcc(), "-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
               "--expt-relaxed-constexpr", "-o", str(out), str(src)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    return outs
