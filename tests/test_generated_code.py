"""The build-time generated DCT of the headline mfcc shape (csrc/gen_mel.cpp -> gen/mel_special.inc,
B2A_DCT_GROUP0 / B2A_DCT_GROUP1: even / odd coefficients over band sums / differences) is plain
C arithmetic, so it is compiled for the host here and checked against the orthonormal DCT-II the
reference's own `_dct_matrix` defines (export_svm.py:69-79; oracle.dct2_ortho_matrix).  No GPU."""
import shutil
import subprocess

import numpy as np
import pytest

from audio_edge_ml_pipeline_b200.build import CSRC, build_lib
from oracle import librosa_restated as L

SRC = r'''
#include <cmath>
#include <cstdio>
#include "gen/mel_special.inc"
// the arithmetic is the includer's choice (logmel512.cu): float32 FMAs here, as in the product build
#define B2A_DCT_T float
#define B2A_DCT_CVT(x) (x)
#define B2A_DCT_FMA(cf, cd, f, a) std::fmaf((cf), (f), (a))
int main() {
    float v[B2A_MELSPEC_NMELS];
    for (int m = 0; m < B2A_MELSPEC_NMELS; ++m) if (std::scanf("%f", &v[m]) != 1) return 2;
    float a0[B2A_DCT_GROUP0_NK] = {0}, a1[B2A_DCT_GROUP1_NK] = {0};
#define LD(M) v[M]
    B2A_DCT_GROUP0(LD, a0)
    B2A_DCT_GROUP1(LD, a1)
    for (int k = 0; k < B2A_DCTSPEC_NMFCC; ++k) std::printf("%.9g\n", (k & 1) ? a1[k / 2] : a0[k / 2]);
    return 0;
}
'''


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_generated_folded_dct_matches_the_dct_matrix(tmp_path):
    build_lib()                                            # (re)generates gen/mel_special.inc
    src = tmp_path / "dct_host.cpp"
    src.write_text(SRC)
    exe = tmp_path / "dct_host"
    subprocess.run(["g++", "-O1", "-std=c++17", "-I", str(CSRC), str(src), "-o", str(exe)], check=True)
    D = L.dct2_ortho_matrix(13, 40)
    rng = np.random.default_rng(5)
    for scale, offset in [(20.0, -40.0), (0.0, -80.0), (5.0, 0.0)]:      # typical dB, constant row, small
        v = (offset + scale * rng.standard_normal(40)).astype(np.float32)
        out = subprocess.run([str(exe)], input=" ".join(f"{x:.9g}" for x in v), capture_output=True, text=True, check=True)
        got = np.array([float(t) for t in out.stdout.split()])
        ref = D @ v.astype(np.float64)
        assert got.shape == (13,)
        assert np.abs(got - ref).max() <= 2e-5 * max(1.0, float(np.abs(v).max())), (got, ref)


SRC_TAPS = r'''
#include <cstdio>
#define __constant__
struct float2 { float x, y; };
#include "gen/decim_taps.inc"
int main() {
    std::printf("%d\n", kDecimTapsGen);
    for (int i = 0; i < (kDecimTapsGen + 1) / 2; ++i) std::printf("%a %a\n", (double)kDecTapP[i].x, (double)kDecTapP[i].y);
    for (int i = 0; i < kDecimTapsGen; ++i) std::printf("%a\n", (double)kDecTapF[i]);
    return 0;
}
'''


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_generated_decimator_taps_are_the_oracles(tmp_path):
    """gen/decim_taps.inc: the 383 taps in natural order and as (odd, even) pairs for the packed FFMA2
    loop of cqt_decimate_kernel — both must be the float32 rounding of oracle.halfband_taps()."""
    build_lib()
    src = tmp_path / "taps_host.cpp"
    src.write_text(SRC_TAPS)
    exe = tmp_path / "taps_host"
    subprocess.run(["g++", "-O1", "-std=c++17", "-I", str(CSRC), str(src), "-o", str(exe)], check=True)
    tok = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()
    n = int(tok[0])
    h = L.halfband_taps().astype(np.float32)
    assert n == len(h) == 383
    vals = np.array([float.fromhex(t) for t in tok[1:]], dtype=np.float64)
    pairs, flat = vals[:2 * 192].reshape(192, 2), vals[2 * 192:]
    assert np.array_equal(flat.astype(np.float32), h)
    hp = np.append(h, np.float32(0.0))                       # the odd phase is zero-padded to 192 taps
    assert np.array_equal(pairs[:, 0].astype(np.float32), hp[1::2])     # .x = odd tap  h[2i+1]
    assert np.array_equal(pairs[:, 1].astype(np.float32), hp[0::2])     # .y = even tap h[2i]
