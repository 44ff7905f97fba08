"""Host-side checks of the register four-step FFT's index arithmetic (csrc/fft_reg.cuh) in numpy:
the split x = a*RB + b, X = c + RA*d with the step-1 twiddle W_n^(b c), and the four-step twiddle
W_N^(k1 n2) = coarse[e >> s] * fine[e & (2^s - 1)] from two fp32 tables (api.cpp: FftCore::create)."""
import numpy as np


def test_two_level_split_reproduces_the_dft():
    rng = np.random.default_rng(3)
    for RA, RB in ((16, 16), (16, 32), (32, 32)):
        n = RA * RB
        x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
        xa = x.reshape(RA, RB)                                            # [a][b] = x[a*RB + b]
        y = np.fft.fft(xa, axis=0)                                        # step 1: DFT over a -> [c][b]
        y *= np.exp(-2j * np.pi * np.outer(np.arange(RA), np.arange(RB)) / n)      # W_n^(b c)
        z = np.fft.fft(y, axis=1)                                         # step 2: DFT over b -> [c][d]
        X = np.empty(n, complex)
        for c in range(RA):
            X[c + RA * np.arange(RB)] = z[c]                              # X[c + RA*d]
        assert np.max(np.abs(X - np.fft.fft(x))) < 1e-9 * n


def test_table_twiddles_are_within_fp32_rounding():
    for logN, logTW in ((20, 10), (16, 8), (19, 10)):
        N, TW = 1 << logN, 1 << logTW
        s = logN - logTW
        coarse = np.exp(-2j * np.pi * np.arange(TW) / TW).astype(np.complex64)
        fine = np.exp(-2j * np.pi * np.arange(1 << s) / N).astype(np.complex64)
        e = np.random.default_rng(1).integers(0, N, 200000)
        w = (coarse[e >> s].astype(np.complex64) * fine[e & ((1 << s) - 1)]).astype(np.complex64)
        exact = np.exp(-2j * np.pi * e / N)
        assert np.max(np.abs(w.astype(np.complex128) - exact)) < 2.5e-7      # two rounded factors and one fp32 product
