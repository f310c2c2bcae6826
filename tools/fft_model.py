#!/usr/bin/env python3
"""Executable model of the register/shared-memory Stockham line FFT used by
fdes_b200/csrc/fft_core.cuh.  Emulates threads with loops so the index algebra
(slot layout theta + m*T, output scatter, twiddle indices, butterfly register permutations)
can be verified on a machine without a GPU.  Run: python tools/fft_model.py"""
import numpy as np

def radix_list(N, E):
    """Greedy: radix E while it divides, then the largest divisor of E that divides the rest."""
    rs, n = [], N
    while n > 1:
        r = max(d for d in range(2, E + 1) if E % d == 0 and n % d == 0)
        rs.append(r); n //= r
    return rs

def dft_small(v, sign):
    R = len(v); n = np.arange(R)
    return np.exp(sign * 2j * np.pi * np.outer(n, n) / R) @ v

def fft_line_model(x, E, radices, sign=-1):
    N = len(x); T = N // E
    tw = np.exp(sign * 2j * np.pi * np.arange(N) / N)
    regs = np.zeros((T, E), complex)
    for th in range(T):
        for m in range(E):
            regs[th, m] = x[th + m * T]
    Ns = 1
    for pi, R in enumerate(radices):
        last = pi == len(radices) - 1
        smem = np.zeros(N, complex)
        for th in range(T):
            for u in range(E // R):
                j = th + u * T
                k = j % Ns
                slots = [u + t * (E // R) for t in range(R)]
                v = regs[th, slots].copy()
                stride = N // (Ns * R)
                for t in range(1, R):
                    v[t] *= tw[(k * t * stride) % N]
                y = dft_small(v, sign)
                if last:
                    assert Ns == N // R
                    regs[th, slots] = y
                else:
                    for t in range(R):
                        smem[(j - k) * R + k + t * Ns] = y[t]
        if not last:
            for th in range(T):
                for m in range(E):
                    regs[th, m] = smem[th + m * T]
        Ns *= R
    out = np.zeros(N, complex)
    for th in range(T):
        for m in range(E):
            out[th + m * T] = regs[th, m]
    return out

if __name__ == "__main__":
    rng = np.random.default_rng(0)
    for N, E in [(16,16),(32,16),(64,16),(128,16),(256,16),(512,16),(1024,16),(2048,16),(4096,16),(320,20),(800,20),(1000,10)]:
        rs = radix_list(N, E)
        x = rng.standard_normal(N) + 1j * rng.standard_normal(N)
        for sign in (-1, 1):
            y = fft_line_model(x, E, rs, sign)
            ref = np.fft.fft(x) if sign < 0 else np.fft.ifft(x) * N
            err = np.abs(y - ref).max() / np.abs(ref).max()
            print(N, E, rs, sign, f"{err:.2e}")
            assert err < 1e-10
