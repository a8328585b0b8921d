"""In-tree build of ``libb2a.so`` (hand-written sm_100a CUDA kernels + the C ABI of include/b2a.h).

``python -m audio_edge_ml_pipeline_b200.build`` or ``build_lib()``; nvcc cross-compiles without a
GPU.  The built library stays inside the package directory (git-ignored, shipped by gpurun).
"""

from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "libb2a.so"
SOURCES = ["api.cu", "frontend.cu", "logmel512.cu", "logmel1024.cu", "cqt.cu", "resample.cu", "augment.cu", "classical.cu", "effects.cu", "tables.cpp", "decode.cpp"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
]


def _nvcc() -> str:
    cand = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(cand).exists():
        raise RuntimeError("nvcc not found: cannot build libb2a.so")
    return cand


def _fingerprint() -> str:
    h = hashlib.sha256()
    for p in sorted([q for q in CSRC.glob("*") if q.is_file()] + [PKG.parent / "include" / "b2a.h", Path(__file__)]):
        if p.is_file():
            h.update(p.name.encode())
            h.update(p.read_bytes())
    h.update(os.environ.get("B2A_NVCC_EXTRA", "").encode())
    return h.hexdigest()


def build_variant(tag: str, defines: list, verbose: bool = False) -> Path:
    """An experimental build with extra -D flags (numerics A/B runs, tools/mfcc_floor.py) as
    build/libb2a_<tag>.so; loaded instead of the product library through the B2A_LIBRARY variable."""
    return build_lib(force=True, verbose=verbose, extra=[f"-D{d}" for d in defines],
                     lib=PKG / "build" / f"libb2a_{tag}.so", objdir=PKG / "build" / f"obj_{tag}")


def build_lib(force: bool = False, verbose: bool = False, extra: list = (), lib: Path = LIB,
              objdir: Path = None) -> Path:
    """Compile (if sources changed) and return the path of libb2a.so."""
    product = lib == LIB
    stamp = PKG / ".libb2a.stamp"
    fp = _fingerprint()
    if product and not force and LIB.exists() and stamp.exists() and stamp.read_text() == fp:
        return LIB
    nvcc = _nvcc()
    gendir = PKG / "build"
    gendir.mkdir(exist_ok=True)
    objdir = objdir or gendir
    objdir.mkdir(exist_ok=True)
    # build-time code generation: straight-line mel code for the headline configuration
    gen = gendir / "gen_mel"
    subprocess.run(["g++", "-O2", "-std=c++17", str(CSRC / "gen_mel.cpp"), str(CSRC / "tables.cpp"), "-o", str(gen)],
                   check=True)
    inc = subprocess.run([str(gen), "16000", "512", "40", "4", "13"], check=True, capture_output=True, text=True).stdout
    (CSRC / "gen").mkdir(exist_ok=True)
    tgt = CSRC / "gen" / "mel_special.inc"
    if not tgt.exists() or tgt.read_text() != inc:
        tgt.write_text(inc)
    inc2 = subprocess.run([str(gen), "decim"], check=True, capture_output=True, text=True).stdout
    tgt2 = CSRC / "gen" / "decim_taps.inc"
    if not tgt2.exists() or tgt2.read_text() != inc2:
        tgt2.write_text(inc2)
    flags = list(NVCC_FLAGS) + os.environ.get("B2A_NVCC_EXTRA", "").split() + list(extra)
    procs = []
    objs = []
    for src in SOURCES:
        obj = objdir / (src.rsplit(".", 1)[0] + ".o")
        objs.append(str(obj))
        cmd = [nvcc, *flags, "-I", str(PKG.parent / "include"), "-c", str(CSRC / src), "-o", str(obj)]
        if verbose:
            cmd[1:1] = ["-Xptxas", "-v"]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            print(out.decode(errors="replace"))
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}")
    subprocess.run([nvcc, "-shared", "-o", str(lib), *objs, "-lcudart_static", "-lpthread", "-ldl", "-lrt"],
                   check=True)
    if product:
        stamp.write_text(fp)
    return lib


if __name__ == "__main__":
    import sys

    print(build_lib(force="--force" in sys.argv, verbose="-v" in sys.argv))
