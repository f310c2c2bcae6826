// Developer microbenchmark for fdes_b200/csrc/fft_core.cuh: a row sweep (IFFT_row -> scale ->
// FFT_row) and a column sweep (FFT_col -> x table -> IFFT_col) over a batch of N x N complex64
// grids, checked against cuFFT (test tooling only) and timed with CUDA events.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -I ../../fdes_b200/csrc fft_bench.cu -lcufft -o fft_bench
#include "fft_core.cuh"
#include <cufft.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
using namespace fdes;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)

template <int N, int E>
std::vector<cpx> make_twiddles()
{
    std::vector<cpx> tw;
    int NS = 1;
    while (NS < N) {
        const int rem = N / NS, R = rem >= E ? E : rem;
        if (NS > 1)
            for (int t = 0; t < R; t++)
                for (int k = 0; k < NS; k++) {
                    const double a = -2.0 * M_PI * (double)t * (double)k / ((double)NS * R);
                    tw.push_back(make_float2((float)cos(a), (float)sin(a)));
                }
        NS *= R;
    }
    if (tw.empty()) tw.push_back(make_float2(1.f, 0.f));
    return tw;
}

template <int N, int E, int LPB, int MINB>
__global__ void __launch_bounds__(LPB * (N / E), MINB) k_rows(cpx* __restrict__ data, const cpx* __restrict__ tw, float scale)
{
    constexpr int T = N / E;
    constexpr int LS = line_smem_elems<E>(N);
    extern __shared__ cpx smem[];
    const int line = threadIdx.x / T, theta = threadIdx.x % T;
    cpx* row = data + ((size_t)blockIdx.x * LPB + line) * N;
    cpx* sm = smem + line * LS;
    cpx x[E];
#pragma unroll
    for (int m = 0; m < E; m++) x[m] = row[theta + m * T];
    if constexpr (T <= 32) {
        fft_line<N, E, 1>(x, sm, theta, tw, SyncWarp());
#pragma unroll
        for (int m = 0; m < E; m++) x[m] = make_float2(x[m].x * scale, x[m].y * scale);
        fft_line<N, E, -1>(x, sm, theta, tw, SyncWarp());
    } else {
        SyncNamed s{line + 1, T};
        fft_line<N, E, 1>(x, sm, theta, tw, s);
#pragma unroll
        for (int m = 0; m < E; m++) x[m] = make_float2(x[m].x * scale, x[m].y * scale);
        fft_line<N, E, -1>(x, sm, theta, tw, s);
    }
#pragma unroll
    for (int m = 0; m < E; m++) row[theta + m * T] = x[m];
}

// quad layout: 4 adjacent lines interleaved element-wise (32-byte sectors hold one position of 4 lines)
template <int N> __device__ __forceinline__ size_t addr_quad(int line, int pos) { return (size_t)(line >> 2) * (4 * N) + 4 * pos + (line & 3); }
template <int N, int E, int LPB, int MINB>
__global__ void __launch_bounds__(LPB * (N / E), MINB) k_quad(const cpx* __restrict__ in, cpx* __restrict__ out, const cpx* __restrict__ tw, float scale)
{
    constexpr int T = N / E;
    constexpr int LS = line_smem_elems<E>(N);
    extern __shared__ cpx smem[];
    const int line = threadIdx.x / T, theta = threadIdx.x % T;
    const int l = blockIdx.x * LPB + line;
    const size_t boff = (size_t)blockIdx.y * N * N;
    cpx* sm = smem + line * LS;
    cpx x[E];
#pragma unroll
    for (int m = 0; m < E; m++) x[m] = in[boff + addr_quad<N>(l, theta + m * T)];
    fft_line<N, E, 1>(x, sm, theta, tw, SyncWarp());
#pragma unroll
    for (int m = 0; m < E; m++) x[m] = make_float2(x[m].x * scale, x[m].y * scale);
    fft_line<N, E, -1>(x, sm, theta, tw, SyncWarp());
#pragma unroll
    for (int m = 0; m < E; m++) out[boff + addr_quad<N>(theta + m * T, l)] = x[m];
}
// row lines in, transposed out through a shared-memory tile that aliases the exchange buffers
template <int N, int E, int G, int MINB>
__global__ void __launch_bounds__(G * (N / E), MINB) k_rowsT(const cpx* __restrict__ in, cpx* __restrict__ out, const cpx* __restrict__ tw, float scale)
{
    constexpr int T = N / E, THREADS = G * T;
    constexpr int LS = line_smem_elems<E>(N);
    constexpr int TS = G + 1;                       // tile row stride (odd: conflict-free scatter)
    extern __shared__ cpx smem[];
    const int line = threadIdx.x / T, theta = threadIdx.x % T;
    const int l0 = blockIdx.x * G;
    const size_t boff = (size_t)blockIdx.y * N * N;
    cpx* sm = smem + line * LS;
    cpx x[E];
#pragma unroll
    for (int m = 0; m < E; m++) x[m] = in[boff + (size_t)(l0 + line) * N + theta + m * T];
    fft_line<N, E, 1>(x, sm, theta, tw, SyncWarp());
#pragma unroll
    for (int m = 0; m < E; m++) x[m] = make_float2(x[m].x * scale, x[m].y * scale);
    fft_line<N, E, -1>(x, sm, theta, tw, SyncWarp());
    __syncthreads();
#pragma unroll
    for (int m = 0; m < E; m++) smem[(theta + m * T) * TS + line] = x[m];
    __syncthreads();
    const int c = threadIdx.x % G, q0 = threadIdx.x / G;
    constexpr int QPI = THREADS / G;
#pragma unroll 8
    for (int i = 0; i < N / QPI; i++) {
        const int q = q0 + i * QPI;
        out[boff + (size_t)q * N + l0 + c] = smem[q * TS + c];
    }
}
struct KeepAll { __device__ __forceinline__ bool operator()(int) const { return true; } };
template <int N, int E, int CW, bool STAGED, int MINB>
__global__ void __launch_bounds__(CW * (N / E), MINB) k_cols(cpx* __restrict__ data, const cpx* __restrict__ tw,
                                                            const cpx* __restrict__ tab)
{
    using Tile = ColTile<N, E, CW, STAGED>;
    extern __shared__ cpx smem[];
    const Tile ctx(smem);
    const int theta = ctx.theta;
    const int kx0 = blockIdx.x * CW, kx = kx0 + ctx.line;
    cpx* tile = data + (size_t)blockIdx.y * N * N + kx0;
    cpx x[E];
    ctx.load(x, tile, KeepAll());
    fft_line<N, E, -1>(x, ctx.sm, theta, tw, ctx);
#pragma unroll
    for (int m = 0; m < E; m++) x[m] = cmul(x[m], ld_nc(tab + (theta + m * Tile::T)));   // column-independent table
    fft_line<N, E, 1>(x, ctx.sm, theta, tw, ctx);
    ctx.store(x, tile);
}

__global__ void k_scale(cpx* d, size_t n, float s) { for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) d[i] = make_float2(d[i].x * s, d[i].y * s); }
__global__ void k_mul(cpx* d, const cpx* t, size_t n, size_t nt) { for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) { cpx a = d[i], b = t[i % nt]; d[i] = make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); } }

__global__ void k_mul_rows(cpx* d, const cpx* t, size_t n, int N) { for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) { cpx a = d[i], b = t[(i / N) % N]; d[i] = make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); } }
static double rel_err(const std::vector<cpx>& a, const std::vector<cpx>& b)
{
    double num = 0, den = 0;
    for (size_t i = 0; i < a.size(); i++) {
        const double dx = (double)a[i].x - b[i].x, dy = (double)a[i].y - b[i].y;
        num += dx * dx + dy * dy;
        den += (double)b[i].x * b[i].x + (double)b[i].y * b[i].y;
    }
    return sqrt(num / den);
}

template <int N, int E, int LPB, int MINBR, int CW, bool STAGED, int MINBC>
void run(int batch, int reps)
{
    const size_t NN = (size_t)N * N, total = NN * batch;
    std::vector<cpx> h(total), tabh(NN);
    srand(1234);
    for (auto& v : h) v = make_float2(rand() / (float)RAND_MAX - 0.5f, rand() / (float)RAND_MAX - 0.5f);
    for (size_t i = 0; i < NN; i++) { const float a = 0.001f * (float)(i % 977); tabh[i] = make_float2(cosf(a) / N, sinf(a) / N); }
    cpx *d, *ref, *tab, *tw;
    CK(cudaMalloc(&d, total * sizeof(cpx))); CK(cudaMalloc(&ref, total * sizeof(cpx))); CK(cudaMalloc(&tab, NN * sizeof(cpx)));
    auto twh = make_twiddles<N, E>();
    CK(cudaMalloc(&tw, twh.size() * sizeof(cpx)));
    CK(cudaMemcpy(tw, twh.data(), twh.size() * sizeof(cpx), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(tab, tabh.data(), NN * sizeof(cpx), cudaMemcpyHostToDevice));
    constexpr int T = N / E;
    const size_t smem_r = (size_t)LPB * line_smem_elems<E>(N) * sizeof(cpx);
    const size_t smem_c = ColTile<N, E, CW, STAGED>::SMEM;
    CK(cudaFuncSetAttribute(k_rows<N, E, LPB, MINBR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_r));
    CK(cudaFuncSetAttribute(k_cols<N, E, CW, STAGED, MINBC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_c));
    int occ_r = 0, occ_c = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_r, k_rows<N, E, LPB, MINBR>, LPB * T, smem_r);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_c, k_cols<N, E, CW, STAGED, MINBC>, CW * T, smem_c);
    cudaFuncAttributes fr, fc;
    cudaFuncGetAttributes(&fr, k_rows<N, E, LPB, MINBR>); cudaFuncGetAttributes(&fc, k_cols<N, E, CW, STAGED, MINBC>);
    // ---- correctness vs cuFFT
    cufftHandle plan_r, plan_c;
    int n1[1] = {N};
    cufftPlanMany(&plan_r, 1, n1, n1, 1, N, n1, 1, N, CUFFT_C2C, N * batch);          // rows
    int inembed[1] = {N};
    cufftPlanMany(&plan_c, 1, n1, inembed, N, 1, inembed, N, 1, CUFFT_C2C, N);          // columns of one grid
    std::vector<cpx> got(total), want(total);
    const float scale = 1.f / N;
    CK(cudaMemcpy(d, h.data(), total * sizeof(cpx), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ref, h.data(), total * sizeof(cpx), cudaMemcpyHostToDevice));
    k_rows<N, E, LPB, MINBR><<<N * batch / LPB, LPB * T, smem_r>>>(d, tw, scale);
    cufftExecC2C(plan_r, ref, ref, CUFFT_INVERSE); k_scale<<<1184, 256>>>(ref, total, scale); cufftExecC2C(plan_r, ref, ref, CUFFT_FORWARD);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(got.data(), d, total * sizeof(cpx), cudaMemcpyDeviceToHost)); CK(cudaMemcpy(want.data(), ref, total * sizeof(cpx), cudaMemcpyDeviceToHost));
    const double er = rel_err(got, want), er0 = rel_err(got, h);
    CK(cudaMemcpy(d, h.data(), total * sizeof(cpx), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ref, h.data(), total * sizeof(cpx), cudaMemcpyHostToDevice));
    k_cols<N, E, CW, STAGED, MINBC><<<dim3(N / CW, batch), CW * T, smem_c>>>(d, tw, tab);
    for (int b = 0; b < batch; b++) cufftExecC2C(plan_c, ref + b * NN, ref + b * NN, CUFFT_FORWARD);
    k_mul_rows<<<1184, 256>>>(ref, tab, total, N);
    for (int b = 0; b < batch; b++) cufftExecC2C(plan_c, ref + b * NN, ref + b * NN, CUFFT_INVERSE);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(got.data(), d, total * sizeof(cpx), cudaMemcpyDeviceToHost)); CK(cudaMemcpy(want.data(), ref, total * sizeof(cpx), cudaMemcpyDeviceToHost));
    const double ec = rel_err(got, want);
    // ---- timing
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms_r, ms_c;
    for (int w = 0; w < 2; w++) {
        cudaEventRecord(e0);
        for (int r = 0; r < reps; r++) k_rows<N, E, LPB, MINBR><<<N * batch / LPB, LPB * T, smem_r>>>(d, tw, scale);
        cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms_r, e0, e1);
        cudaEventRecord(e0);
        for (int r = 0; r < reps; r++) k_cols<N, E, CW, STAGED, MINBC><<<dim3(N / CW, batch), CW * T, smem_c>>>(d, tw, tab);
        cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms_c, e0, e1);
    }
    CK(cudaDeviceSynchronize());
    const double gb = 16.0 * total / 1e9;
    printf("N=%4d E=%2d LPB=%d/%d CW=%d%s/%d batch=%2d | rows: %7.2f us %6.0f GB/s regs=%3d occ=%d err=%.1e (vs id %.1e) | cols: %7.2f us %6.0f GB/s regs=%3d occ=%d err=%.1e\n",
           N, E, LPB, MINBR, CW, STAGED ? "s" : "u", MINBC, batch, ms_r / reps * 1e3, gb / (ms_r / reps * 1e-3), fr.numRegs, occ_r, er, er0,
           ms_c / reps * 1e3, gb / (ms_c / reps * 1e-3), fc.numRegs, occ_c, ec);
    cufftDestroy(plan_r); cufftDestroy(plan_c);
    cudaFree(d); cudaFree(ref); cudaFree(tab); cudaFree(tw);
}

template <int N, int E, int LPB, int MINB>
void run_quad(int batch, int reps)
{
    const size_t NN = (size_t)N * N, total = NN * batch;
    std::vector<cpx> h(total), got(total);
    srand(99);
    for (auto& v : h) v = make_float2(rand() / (float)RAND_MAX - 0.5f, rand() / (float)RAND_MAX - 0.5f);
    cpx *a, *b, *tw;
    CK(cudaMalloc(&a, total * sizeof(cpx))); CK(cudaMalloc(&b, total * sizeof(cpx)));
    auto twh = make_twiddles<N, E>();
    CK(cudaMalloc(&tw, twh.size() * sizeof(cpx)));
    CK(cudaMemcpy(tw, twh.data(), twh.size() * sizeof(cpx), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(a, h.data(), total * sizeof(cpx), cudaMemcpyHostToDevice));
    constexpr int T = N / E;
    const size_t smem = (size_t)LPB * line_smem_elems<E>(N) * sizeof(cpx);
    CK(cudaFuncSetAttribute(k_quad<N, E, LPB, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_quad<N, E, LPB, MINB>, LPB * T, smem);
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, k_quad<N, E, LPB, MINB>);
    dim3 grid(N / LPB, batch);
    k_quad<N, E, LPB, MINB><<<grid, LPB * T, smem>>>(a, b, tw, 1.f / N);
    k_quad<N, E, LPB, MINB><<<grid, LPB * T, smem>>>(b, a, tw, 1.f / N);   // transposed twice = identity
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(got.data(), a, total * sizeof(cpx), cudaMemcpyDeviceToHost));
    const double err = rel_err(got, h);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms = 0;
    for (int w = 0; w < 2; w++) {
        cudaEventRecord(e0);
        for (int r = 0; r < reps; r++) k_quad<N, E, LPB, MINB><<<grid, LPB * T, smem>>>((r & 1) ? b : a, (r & 1) ? a : b, tw, 1.f / N);
        cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    }
    printf("N=%4d E=%2d LPB=%d/%d batch=%2d | quad-transposing: %7.2f us %6.0f GB/s regs=%3d occ=%d roundtrip err=%.1e\n", N, E, LPB, MINB, batch,
           ms / reps * 1e3, 16.0 * total / 1e9 / (ms / reps * 1e-3), fa.numRegs, occ, err);
    cudaFree(a); cudaFree(b); cudaFree(tw);
}

template <int N, int E, int G, int MINB>
void run_rowsT(int batch, int reps)
{
    const size_t NN = (size_t)N * N, total = NN * batch;
    std::vector<cpx> h(total), got(total);
    srand(99);
    for (auto& v : h) v = make_float2(rand() / (float)RAND_MAX - 0.5f, rand() / (float)RAND_MAX - 0.5f);
    cpx *a, *b, *tw;
    CK(cudaMalloc(&a, total * sizeof(cpx))); CK(cudaMalloc(&b, total * sizeof(cpx)));
    auto twh = make_twiddles<N, E>();
    CK(cudaMalloc(&tw, twh.size() * sizeof(cpx)));
    CK(cudaMemcpy(tw, twh.data(), twh.size() * sizeof(cpx), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(a, h.data(), total * sizeof(cpx), cudaMemcpyHostToDevice));
    constexpr int T = N / E;
    const size_t s1 = (size_t)G * line_smem_elems<E>(N) * sizeof(cpx), s2 = (size_t)N * (G + 1) * sizeof(cpx);
    const size_t smem = s1 > s2 ? s1 : s2;
    CK(cudaFuncSetAttribute(k_rowsT<N, E, G, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_rowsT<N, E, G, MINB>, G * T, smem);
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, k_rowsT<N, E, G, MINB>);
    dim3 grid(N / G, batch);
    k_rowsT<N, E, G, MINB><<<grid, G * T, smem>>>(a, b, tw, 1.f / N);
    k_rowsT<N, E, G, MINB><<<grid, G * T, smem>>>(b, a, tw, 1.f / N);   // transposed twice = identity
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(got.data(), a, total * sizeof(cpx), cudaMemcpyDeviceToHost));
    const double err = rel_err(got, h);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms = 0;
    for (int w = 0; w < 2; w++) {
        cudaEventRecord(e0);
        for (int r = 0; r < reps; r++) k_rowsT<N, E, G, MINB><<<grid, G * T, smem>>>((r & 1) ? b : a, (r & 1) ? a : b, tw, 1.f / N);
        cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    }
    printf("N=%4d E=%2d G=%d/%d batch=%2d | rows-transposing: %7.2f us %6.0f GB/s regs=%3d occ=%d roundtrip err=%.1e\n", N, E, G, MINB, batch,
           ms / reps * 1e3, 16.0 * total / 1e9 / (ms / reps * 1e-3), fa.numRegs, occ, err);
    cudaFree(a); cudaFree(b); cudaFree(tw);
}

int main(int argc, char** argv)
{
    const int reps = 20;
    run<1024, 32, 4, 1, 8, false, 1>(16, reps);
    run<1024, 16, 2, 4, 4, false, 2>(16, reps);
    run<1024, 16, 2, 5, 4, false, 3>(16, reps);
    run<1024, 16, 2, 6, 8, false, 1>(16, reps);
    run<1024, 16, 4, 3, 8, false, 2>(16, reps);
    run<1024, 8, 1, 8, 2, false, 4>(16, reps);
    return 0;
}
