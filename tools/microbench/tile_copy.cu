// Microbenchmark: bandwidth of moving [N rows][CW columns] complex64 tiles of an N x N x batch
// array through registers with the column-fastest thread mapping (the access pattern of a column
// sweep) for different tile widths, against a contiguous row copy.
#include <cuda_runtime.h>
#include <cstdio>
typedef float2 cpx;
template <int N, int CW, int THREADS, int UNROLL>
__global__ void __launch_bounds__(THREADS) k_tile(const cpx* __restrict__ in, cpx* __restrict__ out)
{
    constexpr int RPI = THREADS / CW, ITERS = N / RPI;
    const int c = threadIdx.x % CW, r = threadIdx.x / CW;
    const size_t base = (size_t)blockIdx.y * N * N + blockIdx.x * CW + c;
    for (int i0 = 0; i0 < ITERS; i0 += UNROLL) {
        cpx v[UNROLL];
#pragma unroll
        for (int i = 0; i < UNROLL; i++) v[i] = in[base + (size_t)(r + (i0 + i) * RPI) * N];
#pragma unroll
        for (int i = 0; i < UNROLL; i++) out[base + (size_t)(r + (i0 + i) * RPI) * N] = make_float2(v[i].y, v[i].x);
    }
}
template <int N, int THREADS, int UNROLL>
__global__ void __launch_bounds__(THREADS) k_rowcopy(const cpx* __restrict__ in, cpx* __restrict__ out)
{
    // each warp copies whole rows: lane + 32 m
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const size_t row = ((size_t)blockIdx.x * (THREADS / 32) + warp) * N;
    for (int m0 = 0; m0 < N / 32; m0 += UNROLL) {
        cpx v[UNROLL];
#pragma unroll
        for (int i = 0; i < UNROLL; i++) v[i] = in[row + lane + 32 * (m0 + i)];
#pragma unroll
        for (int i = 0; i < UNROLL; i++) out[row + lane + 32 * (m0 + i)] = make_float2(v[i].y, v[i].x);
    }
}
template <class F> float timeit(F f, int reps)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaEventRecord(e0);
    for (int i = 0; i < reps; i++) f();
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms / reps;
}
template <int CW, int THREADS, int UNROLL>
void run_tile(const cpx* a, cpx* b, int batch)
{
    constexpr int N = 1024;
    float ms = timeit([&] { k_tile<N, CW, THREADS, UNROLL><<<dim3(N / CW, batch), THREADS>>>(a, b); }, 20);
    printf("tile CW=%2d (%3d B segments) threads=%3d unroll=%2d batch=%2d: %7.2f us  %6.0f GB/s\n", CW, CW * 8, THREADS, UNROLL, batch,
           ms * 1e3, 16.0 * N * N * batch / (ms * 1e-3) / 1e9);
}
int main()
{
    constexpr int N = 1024;
    for (int batch : {16, 4}) {
        cpx *a, *b;
        cudaMalloc(&a, sizeof(cpx) * N * N * batch); cudaMalloc(&b, sizeof(cpx) * N * N * batch);
        cudaMemset(a, 0, sizeof(cpx) * N * N * batch);
        float ms = timeit([&] { k_rowcopy<N, 128, 32><<<N * batch / 4, 128>>>(a, b); }, 20);
        printf("row copy threads=128 unroll=32 batch=%2d: %7.2f us  %6.0f GB/s\n", batch, ms * 1e3, 16.0 * N * N * batch / (ms * 1e-3) / 1e9);
        ms = timeit([&] { k_rowcopy<N, 256, 8><<<N * batch / 8, 256>>>(a, b); }, 20);
        printf("row copy threads=256 unroll= 8 batch=%2d: %7.2f us  %6.0f GB/s\n", batch, ms * 1e3, 16.0 * N * N * batch / (ms * 1e-3) / 1e9);
        run_tile<4, 128, 32>(a, b, batch);
        run_tile<8, 256, 32>(a, b, batch);
        run_tile<8, 256, 8>(a, b, batch);
        run_tile<16, 256, 16>(a, b, batch);
        run_tile<16, 512, 32>(a, b, batch);
        run_tile<32, 256, 8>(a, b, batch);
        run_tile<32, 1024, 32>(a, b, batch);
        cudaFree(a); cudaFree(b);
    }
    return 0;
}
