// fdes_b200 -- register/shared-memory Stockham line FFT for sm_100a (complex float32).
//
// One "line" (a grid row or a grid column) of N points is transformed by T = N/E threads that
// each keep E = 16 points in registers at positions theta + m*T (m = 0..E-1).  A transform is a
// sequence of radix-R passes (R = 16 while it divides what is left, then 8/4/2).  In a pass a
// thread performs E/R radix-R butterflies on registers; between passes the line is exchanged
// through padded shared memory (Stockham autosort: natural order in, natural order out).  The
// register layout before the first pass and after the last pass is the same, so
//   * global loads/stores are coalesced (consecutive threads <-> consecutive points), and
//   * transforms can be chained (inverse -> pointwise op -> forward) with the data staying in
//     registers -- this is what lets a whole multislice sweep touch HBM once.
// The index algebra is checked on the CPU by tools/fft_model.py.
//
// Replaces the cuFFT C2C calls of the reference hot path (cufftExecC2C at
// src/crystalMaker.cu:527,531 and src/multisliceSimulation.cu:554,556,608,610).
#pragma once
#include <cuda_runtime.h>

namespace fdes {

typedef float2 cpx;

constexpr int FFT_E = 16;  // points per thread

__device__ __forceinline__ cpx cmul(cpx a, cpx b)
{
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ cpx cadd(cpx a, cpx b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ cpx csub(cpx a, cpx b) { return make_float2(a.x - b.x, a.y - b.y); }
// multiply by DIR * i  (DIR = -1: forward transform, W4 = -i)
template <int DIR>
__device__ __forceinline__ cpx mul_di(cpx a)
{
    return DIR < 0 ? make_float2(a.y, -a.x) : make_float2(-a.y, a.x);
}
// multiply by (c + DIR*i*s)
template <int DIR>
__device__ __forceinline__ cpx mul_w(cpx a, float c, float s)
{
    const float sd = DIR < 0 ? -s : s;
    return make_float2(a.x * c - a.y * sd, a.x * sd + a.y * c);
}

template <int DIR>
__device__ __forceinline__ void bfly2(cpx& a, cpx& b)
{
    const cpx t = a;
    a = cadd(t, b);
    b = csub(t, b);
}

// 4-point DFT, natural order in / out.
template <int DIR>
__device__ __forceinline__ void bfly4(cpx& x0, cpx& x1, cpx& x2, cpx& x3)
{
    const cpx t0 = cadd(x0, x2), t1 = csub(x0, x2);
    const cpx t2 = cadd(x1, x3), t3 = mul_di<DIR>(csub(x1, x3));
    x0 = cadd(t0, t2);
    x2 = csub(t0, t2);
    x1 = cadd(t1, t3);
    x3 = csub(t1, t3);
}

// R-point DFT on v[0..R-1], natural order in / out.
template <int R, int DIR>
struct Butterfly;

template <int DIR>
struct Butterfly<2, DIR> {
    __device__ __forceinline__ static void run(cpx (&v)[2]) { bfly2<DIR>(v[0], v[1]); }
};
template <int DIR>
struct Butterfly<4, DIR> {
    __device__ __forceinline__ static void run(cpx (&v)[4]) { bfly4<DIR>(v[0], v[1], v[2], v[3]); }
};
template <int DIR>
struct Butterfly<8, DIR> {
    // t = t1 + 2*t2 (t1<2, t2<4), s = 4*s1 + s2:  W8^{st} = W2^{s1 t1} W8^{s2 t1} W4^{s2 t2}
    __device__ __forceinline__ static void run(cpx (&v)[8])
    {
        constexpr float h = 0.70710678118654752440f;
        bfly4<DIR>(v[0], v[2], v[4], v[6]);  // Y[0][s2] at v[2*s2]
        bfly4<DIR>(v[1], v[3], v[5], v[7]);  // Y[1][s2] at v[1+2*s2]
        v[3] = mul_w<DIR>(v[3], h, h);       // W8^1
        v[5] = mul_di<DIR>(v[5]);            // W8^2
        v[7] = mul_w<DIR>(v[7], -h, h);      // W8^3
        bfly2<DIR>(v[0], v[1]);              // X[s2] , X[4+s2] at v[2 s2], v[1+2 s2]
        bfly2<DIR>(v[2], v[3]);
        bfly2<DIR>(v[4], v[5]);
        bfly2<DIR>(v[6], v[7]);
        // X[4 s1 + s2] sits at v[s1 + 2 s2] -> natural order
        const cpx a1 = v[1], a2 = v[2], a3 = v[3], a4 = v[4], a5 = v[5], a6 = v[6];
        v[1] = a2; v[2] = a4; v[3] = a6; v[4] = a1; v[5] = a3; v[6] = a5;
    }
};
template <int DIR>
struct Butterfly<16, DIR> {
    // t = t1 + 4*t2, s = 4*s1 + s2:  W16^{st} = W4^{s1 t1} W16^{s2 t1} W4^{s2 t2}
    __device__ __forceinline__ static void run(cpx (&v)[16])
    {
        constexpr float c1 = 0.92387953251128675613f, s1 = 0.38268343236508977173f;
        constexpr float h = 0.70710678118654752440f;
#pragma unroll
        for (int t1 = 0; t1 < 4; t1++) bfly4<DIR>(v[t1], v[t1 + 4], v[t1 + 8], v[t1 + 12]);
        // v[t1 + 4 s2] *= W16^{s2 t1}
        v[5] = mul_w<DIR>(v[5], c1, s1);     // s2=1,t1=1 : n=1
        v[6] = mul_w<DIR>(v[6], h, h);       // n=2
        v[7] = mul_w<DIR>(v[7], s1, c1);     // n=3
        v[9] = mul_w<DIR>(v[9], h, h);       // s2=2,t1=1 : n=2
        v[10] = mul_di<DIR>(v[10]);          // n=4
        v[11] = mul_w<DIR>(v[11], -h, h);    // n=6
        v[13] = mul_w<DIR>(v[13], s1, c1);   // s2=3,t1=1 : n=3
        v[14] = mul_w<DIR>(v[14], -h, h);    // n=6
        v[15] = mul_w<DIR>(v[15], -c1, -s1); // n=9: cos=-c1, sin=-s1
#pragma unroll
        for (int s2 = 0; s2 < 4; s2++) bfly4<DIR>(v[4 * s2], v[4 * s2 + 1], v[4 * s2 + 2], v[4 * s2 + 3]);
        // X[4 s1 + s2] sits at v[s1 + 4 s2] -> transpose 4x4
#pragma unroll
        for (int a = 0; a < 4; a++)
#pragma unroll
            for (int b = a + 1; b < 4; b++) {
                const cpx t = v[a + 4 * b];
                v[a + 4 * b] = v[b + 4 * a];
                v[b + 4 * a] = t;
            }
    }
};

// padded shared-memory index (one extra slot per 16) -- conflict-free 64-bit scatter/gather
__device__ __forceinline__ int smpad(int i) { return i + (i >> 4); }
__host__ __device__ constexpr int line_smem_elems(int N) { return N + (N >> 4); }

// forward twiddle table tw[n] = exp(-2 pi i n / N), n < N (built in double on the host)
template <int DIR>
__device__ __forceinline__ cpx ldtw(const cpx* __restrict__ tw, int idx)
{
    cpx w = __ldg(tw + idx);
    if (DIR > 0) w.y = -w.y;
    return w;
}

// twiddle v[t] *= W^(t*kk) for t = 1..R-1, W = exp(DIR 2 pi i / N), kk = k * N/(NS*R)
template <int R, int DIR>
__device__ __forceinline__ void apply_twiddles(cpx (&v)[R], const cpx* __restrict__ tw, int kk)
{
    if (R == 2) {
        v[1] = cmul(v[1], ldtw<DIR>(tw, kk));
    } else if (R == 4) {
        const cpx w1 = ldtw<DIR>(tw, kk), w2 = ldtw<DIR>(tw, 2 * kk);
        v[1] = cmul(v[1], w1);
        v[2] = cmul(v[2], w2);
        v[3] = cmul(v[3], cmul(w1, w2));
    } else if (R == 8) {
        const cpx w1 = ldtw<DIR>(tw, kk), w2 = ldtw<DIR>(tw, 2 * kk), w4 = ldtw<DIR>(tw, 4 * kk);
        const cpx w3 = cmul(w1, w2);
        v[1] = cmul(v[1], w1);
        v[2] = cmul(v[2], w2);
        v[3] = cmul(v[3], w3);
        v[4] = cmul(v[4], w4);
        v[5] = cmul(v[5], cmul(w4, w1));
        v[6] = cmul(v[6], cmul(w4, w2));
        v[7] = cmul(v[7], cmul(w4, w3));
    } else {  // 16
        const cpx w1 = ldtw<DIR>(tw, kk), w2 = ldtw<DIR>(tw, 2 * kk), w4 = ldtw<DIR>(tw, 4 * kk);
        const cpx w8 = ldtw<DIR>(tw, 8 * kk);
        const cpx w3 = cmul(w1, w2);
        cpx lo[8];
        lo[1] = w1; lo[2] = w2; lo[3] = w3; lo[4] = w4;
        lo[5] = cmul(w4, w1); lo[6] = cmul(w4, w2); lo[7] = cmul(w4, w3);
#pragma unroll
        for (int t = 1; t < 8; t++) v[t] = cmul(v[t], lo[t]);
        v[8] = cmul(v[8], w8);
#pragma unroll
        for (int t = 1; t < 8; t++) v[8 + t] = cmul(v[8 + t], cmul(w8, lo[t]));
    }
}

// One Stockham pass of radix R with NS = product of previous radices.
template <int N, int R, int NS, int DIR, bool LAST>
__device__ __forceinline__ void fft_pass(cpx (&x)[FFT_E], cpx* __restrict__ sm, int theta,
                                         const cpx* __restrict__ tw)
{
    constexpr int E = FFT_E, T = N / E, U = E / R;
#pragma unroll
    for (int u = 0; u < U; u++) {
        const int j = theta + u * T;
        const int k = j & (NS - 1);
        cpx v[R];
#pragma unroll
        for (int t = 0; t < R; t++) v[t] = x[u + t * U];
        if (NS > 1) apply_twiddles<R, DIR>(v, tw, k * (N / (NS * R)));
        Butterfly<R, DIR>::run(v);
        if (LAST) {
#pragma unroll
            for (int t = 0; t < R; t++) x[u + t * U] = v[t];
        } else {
            const int base = (j - k) * R + k;
#pragma unroll
            for (int t = 0; t < R; t++) sm[smpad(base + t * NS)] = v[t];
        }
    }
    if (!LAST) {
        __syncthreads();
#pragma unroll
        for (int m = 0; m < E; m++) x[m] = sm[smpad(theta + m * T)];
        __syncthreads();
    }
}

template <int N, int NS, int DIR>
struct LinePasses {
    __device__ __forceinline__ static void run(cpx (&x)[FFT_E], cpx* sm, int theta, const cpx* tw)
    {
        constexpr int rem = N / NS;
        constexpr int R = rem >= FFT_E ? FFT_E : rem;
        constexpr bool last = (rem == R);
        fft_pass<N, R, NS, DIR, last>(x, sm, theta, tw);
        if constexpr (!last) LinePasses<N, NS * R, DIR>::run(x, sm, theta, tw);
    }
};

// Transform one line held as x[m] = f[theta + m*N/16].  All threads of the CTA must call it
// together (it contains __syncthreads when N > 16).  sm: this line's line_smem_elems(N) slots.
template <int N, int DIR>
__device__ __forceinline__ void fft_line(cpx (&x)[FFT_E], cpx* sm, int theta, const cpx* tw)
{
    static_assert(N >= 16 && (N & (N - 1)) == 0, "power-of-two line length >= 16");
    LinePasses<N, 1, DIR>::run(x, sm, theta, tw);
}

}  // namespace fdes
