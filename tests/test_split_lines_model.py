"""CPU model of the split lines of fdes_b200/csrc/fft_core.cuh (split_dif / split_dit_combine) and of the
row masks of atoms.cu (launch_row_masks): threads are emulated with loops, so the index algebra -- which warp
holds which point, which 16 values change hands in the half exchange, which bit of which mask word belongs
to which row -- is checked without a GPU.  The GPU parity tests at 2048^2 exercise the kernels themselves.

A 2048-point line on two warps w = 0, 1 of 32 lanes, 32 points per thread, j = lane + 32 m:
  first step from memory:  warp 0: a[j] = f[j] + f[j+1024];  warp 1: b[j] = (f[j] - f[j+1024]) W^j
                           X[2k] = FFT_1024(a)[k], X[2k+1] = FFT_1024(b)[k]; thread (w, lane) holds
                           x[m] = X[2 (lane + 32 m) + w]
  last step to memory:     A = FFT_1024(g even), B = FFT_1024(g odd); warp 1 multiplies by W^k; warp 0 hands
                           A[k], m >= 16, to warp 1 and takes W^k B[k], m < 16; thread (0, lane) emits k = lane + 32 m,
                           m < 16, thread (1, lane) m >= 16: X[k] = A + W^k B, X[k+1024] = A - W^k B
"""
import numpy as np
import pytest

N, E, H = 2048, 32, 1024


def split_dif_model(f, sign):
    """x[w][lane][m] after the radix-2 step on the loads and the per-warp 1024-point transform."""
    W = np.exp(sign * 2j * np.pi * np.arange(H) / N)
    x = np.zeros((2, 32, E), complex)
    for w in range(2):
        half = np.zeros(H, complex)
        for lane in range(32):
            for m in range(E):
                j = lane + 32 * m
                half[j] = f[j] + f[j + H] if w == 0 else (f[j] - f[j + H]) * W[j]
        Xw = np.fft.fft(half) if sign < 0 else np.fft.ifft(half) * H
        for lane in range(32):
            for m in range(E):
                x[w, lane, m] = Xw[lane + 32 * m]
    return x


def split_dit_model(x, sign):
    """Line spectrum from x[w][lane][m] = g[2 (lane + 32 m) + w] through per-warp transforms and the half exchange."""
    W = np.exp(sign * 2j * np.pi * np.arange(H) / N)
    y = np.zeros((2, 32, E), complex)
    for w in range(2):
        half = np.array([x[w, k % 32, k // 32] for k in range(H)])
        Yw = np.fft.fft(half) if sign < 0 else np.fft.ifft(half) * H
        for lane in range(32):
            for m in range(E):
                k = lane + 32 * m
                y[w, lane, m] = Yw[k] * (W[k] if w == 1 else 1.0)
    # exchange slots: mine[m * 32 + lane], 16 per thread
    slots = np.zeros((2, 16 * 32), complex)
    for lane in range(32):
        for m in range(16):
            slots[0, m * 32 + lane] = y[0, lane, m + 16]       # warp 0 hands A[k], m >= 16
            slots[1, m * 32 + lane] = y[1, lane, m]            # warp 1 hands W^k B[k], m < 16
    out = np.zeros(N, complex)
    written = np.zeros(N, int)
    for lane in range(32):
        for m in range(16):                                    # thread (0, lane)
            k, b = lane + 32 * m, slots[1, m * 32 + lane]
            out[k], out[k + H] = y[0, lane, m] + b, y[0, lane, m] - b
            written[[k, k + H]] += 1
        for m in range(16, 32):                                # thread (1, lane)
            k, a = lane + 32 * m, slots[0, (m - 16) * 32 + lane]
            out[k], out[k + H] = a + y[1, lane, m], a - y[1, lane, m]
            written[[k, k + H]] += 1
    assert (written == 1).all()
    return out


@pytest.mark.parametrize("sign", [-1, 1])
def test_split_dif_holds_the_interleaved_spectrum(sign):
    rng = np.random.default_rng(3)
    f = rng.standard_normal(N) + 1j * rng.standard_normal(N)
    X = np.fft.fft(f) if sign < 0 else np.fft.ifft(f) * N
    x = split_dif_model(f, sign)
    for w in range(2):
        for lane in range(32):
            for m in range(E):
                assert abs(x[w, lane, m] - X[2 * (lane + 32 * m) + w]) < 1e-9 * np.abs(X).max()


@pytest.mark.parametrize("sign", [-1, 1])
def test_split_dit_half_exchange(sign):
    rng = np.random.default_rng(4)
    g = rng.standard_normal(N) + 1j * rng.standard_normal(N)
    x = np.zeros((2, 32, E), complex)
    for w in range(2):
        for lane in range(32):
            for m in range(E):
                x[w, lane, m] = g[2 * (lane + 32 * m) + w]
    X = np.fft.fft(g) if sign < 0 else np.fft.ifft(g) * N
    assert np.abs(split_dit_model(x, sign) - X).max() < 1e-9 * np.abs(X).max()


def test_row_sweep_chain_inverse_pointwise_forward():
    """S5's chain: inverse (split on the loads) of two rows, product at the interleaved positions, forward
    transform (half exchange) == FFT(IFFT(a) * IFFT(b))."""
    rng = np.random.default_rng(5)
    a = rng.standard_normal(N) + 1j * rng.standard_normal(N)
    b = rng.standard_normal(N) + 1j * rng.standard_normal(N)
    prod = split_dif_model(a, +1) * split_dif_model(b, +1)
    ref = np.fft.fft((np.fft.ifft(a) * N) * (np.fft.ifft(b) * N))
    assert np.abs(split_dit_model(prod, -1) - ref).max() < 1e-9 * np.abs(ref).max()


@pytest.mark.parametrize("n,e", [(1024, 32), (2048, 32), (256, 16), (800, 20), (64, 8)])
def test_row_mask_words(n, e):
    """Bit m of word theta <-> row theta + m*T (k_row_masks); the split columns read words lane and lane + 32 and
    find row lane + 32 r at bit r/2 of word r%2 (KeepMask::split)."""
    rng = np.random.default_rng(n)
    T = n // e
    counts = rng.integers(0, 3, n) * (rng.random(n) < 0.3)
    rp = np.concatenate([[0], np.cumsum(counts)])
    words = np.zeros(T, np.uint32)
    for theta in range(T):
        for m in range(e):
            k = theta + m * T
            if rp[k + 1] > rp[k]:
                words[theta] |= np.uint32(1) << np.uint32(m)
    for row in range(n):
        assert bool((words[row % T] >> np.uint32(row // T)) & 1) == bool(counts[row] > 0)
    if n == 2048:
        for lane in range(32):
            w0, w1 = words[lane], words[lane + 32]
            for r in range(64):
                bit = ((w1 if r & 1 else w0) >> np.uint32(r >> 1)) & 1
                assert bool(bit) == bool(counts[lane + 32 * r] > 0)
