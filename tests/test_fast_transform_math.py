"""The factorizations the product build uses for the IMDCT and the synthesis matrixing, restated in numpy and checked
against the reference's direct definitions (imdct.go:99-107, frame.go:488-497,644-653).

This pins the index maps, signs and scale factors that kernels.cuh (imdct36_emit, lee_dct, matrix_slot) and
mp3gpu.cu (c_winz) rely on; the CUDA code itself is checked against the oracle by the -m gpu parity tests.
"""
import numpy as np


def dct3_9(a):
    j = np.arange(9)[:, None]
    k = np.arange(9)[None, :]
    return (a[:, None] * np.cos(np.pi * j * (2 * k + 1) / 18)).sum(0)


def fast_imdct36(x):
    """out[p] = sum_m x[m] cos(pi/72 (2p + 19)(2m + 1)), p = 0..35, through an 18-point DCT-IV."""
    xp = x.copy()
    xp[1:] += x[:-1]                                   # x'[m] = x[m] + x[m-1]
    a = xp[0::2]                                       # x'[2j]
    b = xp[1::2].copy()
    b[1:] += xp[1:-1:2]                                # x'[2j+1] + x'[2j-1]
    k = np.arange(9)
    E = dct3_9(a)
    O = dct3_9(b) / (2 * np.cos(np.pi * (2 * k + 1) / 36))
    z = np.concatenate([E + O, (E - O)[::-1]])         # z[k], z[17-k]
    y = z / (2 * np.cos(np.pi * (2 * np.arange(18) + 1) / 72))
    out = np.empty(36)
    p = np.arange(36)
    out[:9] = y[p[:9] + 9]
    out[9:27] = -y[26 - p[9:27]]
    out[27:] = -y[p[27:] - 27]
    return out


def test_imdct36_is_a_dct4_read_out_with_signs():
    rng = np.random.default_rng(1)
    m = np.arange(18)[:, None]
    p = np.arange(36)[None, :]
    cos36 = np.cos(np.pi / 72 * (2 * p + 1 + 18) * (2 * m + 1))      # imdct.go:72-79
    for _ in range(20):
        x = rng.standard_normal(18)
        assert np.allclose(fast_imdct36(x), x @ cos36, rtol=0, atol=1e-12)


def test_winz_index_map_matches_the_read_out():
    # mp3gpu.cu: windowed out[p] = z[idx(p)] * (+-sec72[idx(p)] * win[p]) with idx as in imdct36_emit
    p = np.arange(36)
    kk = np.where(p < 9, p + 9, np.where(p < 27, 26 - p, p - 27))
    zidx = np.concatenate([9 + np.arange(9), 9 + (17 - np.arange(9, 18)), 8 - np.arange(9), 8 - (17 - np.arange(9, 18))])
    # first(p) and first(17-p) read z[9+p]; second(q) and second(17-q) read z[8-q]
    assert np.array_equal(kk, zidx)


def lee_dct(x):
    n = len(x)
    if n == 1:
        return x.copy()
    h = n // 2
    i = np.arange(h)
    u = x[:h] + x[::-1][:h]
    v = (x[:h] - x[::-1][:h]) / (2 * np.cos(np.pi * (2 * i + 1) / (2 * n)))
    E, W = lee_dct(u), lee_dct(v)
    X = np.empty(n)
    X[0::2] = E
    X[1::2] = W + np.concatenate([W[1:], [0.0]])
    return X


def test_lee_recursion_is_the_dct2():
    rng = np.random.default_rng(2)
    j = np.arange(32)[None, :]
    n = np.arange(32)[:, None]
    for _ in range(10):
        s = rng.standard_normal(32)
        assert np.allclose(lee_dct(s), (np.cos(n * (2 * j + 1) * np.pi / 64) * s).sum(1), rtol=0, atol=1e-11)


def test_matrixing_rows_are_signed_dct_outputs():
    # frame.go:488-497: V[i] = sum_j cos((16 + i)(2j + 1) pi / 64) s[j]; kernels.cuh stores U[0..15] = c[16..31], U[16] = 0,
    # U[17..32] = c[15..0] and reads V[i] = +-U[ai], V[32+i] = -U[bi]
    rng = np.random.default_rng(3)
    s = rng.standard_normal(32)
    i = np.arange(64)[:, None]
    j = np.arange(32)[None, :]
    V = (np.cos((16 + i) * (2 * j + 1) * np.pi / 64) * s).sum(1)
    c = lee_dct(s)
    U = np.concatenate([c[16:32], [0.0], c[15::-1]])
    for lane in range(32):
        ai = lane if lane <= 16 else 32 - lane
        bi = 0 if lane == 0 else (16 + lane if lane <= 16 else 48 - lane)
        sa = 1.0 if lane <= 16 else -1.0
        assert abs(V[lane] - sa * U[ai]) < 1e-11, lane
        assert abs(V[32 + lane] + U[bi]) < 1e-11, lane
