// Microbenchmark: issue rate of scalar FADD/FFMA vs packed FADD2/FFMA2/FMUL2 (f32x2) on sm_100a.
// Prints warp-instructions per cycle per SM for each variant (clock from cudaDevAttrClockRate is
// nominal; the table is for ratios).
#include <cuda_runtime.h>
#include <cstdio>
__device__ __forceinline__ float2 add2(float2 a, float2 b){ float2 c;
  asm("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; add.rn.f32x2 rc, ra, rb; mov.b64 {%0,%1}, rc;}" : "=f"(c.x), "=f"(c.y) : "f"(a.x),"f"(a.y),"f"(b.x),"f"(b.y)); return c; }
__device__ __forceinline__ float2 mul2(float2 a, float2 b){ float2 c;
  asm("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; mul.rn.f32x2 rc, ra, rb; mov.b64 {%0,%1}, rc;}" : "=f"(c.x), "=f"(c.y) : "f"(a.x),"f"(a.y),"f"(b.x),"f"(b.y)); return c; }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c){ float2 d;
  asm("{.reg .b64 ra, rb, rc, rd; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; mov.b64 rc, {%6,%7}; fma.rn.f32x2 rd, ra, rb, rc; mov.b64 {%0,%1}, rd;}" : "=f"(d.x), "=f"(d.y) : "f"(a.x),"f"(a.y),"f"(b.x),"f"(b.y),"f"(c.x),"f"(c.y)); return d; }

template <int MODE>
__global__ void __launch_bounds__(256) k(float2* p, int iters, float2 w)
{
    float2 a[8];
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = p[threadIdx.x + 256 * i];
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (MODE == 0) { a[i].x += w.x; a[i].y += w.y; }                       // 2 FADD
            if (MODE == 1) a[i] = add2(a[i], w);                                   // 1 FADD2
            if (MODE == 2) { a[i].x = fmaf(a[i].x, w.x, w.y); a[i].y = fmaf(a[i].y, w.y, w.x); }  // 2 FFMA
            if (MODE == 3) a[i] = fma2(a[i], w, w);                                // 1 FFMA2
            if (MODE == 4) { float2 t = mul2(make_float2(a[i].y, a[i].y), make_float2(w.y, w.x));   // packed cmul-like: FMUL2 + FFMA2
                             a[i] = fma2(make_float2(a[i].x, a[i].x), w, t); }
            if (MODE == 5) { float ax = a[i].x, ay = a[i].y;                       // scalar cmul: 2 FMUL + 2 FFMA
                             a[i].x = fmaf(ax, w.x, -ay * w.y); a[i].y = fmaf(ax, w.y, ay * w.x); }
            if (MODE == 6) { a[i] = add2(a[i], a[(i + 1) & 7]); }                  // FADD2 with 2 distinct register pairs
            if (MODE == 7) { a[i].x += a[(i + 1) & 7].x; a[i].y += a[(i + 1) & 7].y; }  // 2 FADD, distinct regs
        }
    }
#pragma unroll
    for (int i = 0; i < 8; i++) p[threadIdx.x + 256 * i] = a[i];
}
template <int MODE>
void run(const char* name, float2* d, int nsm, int inst_per_elem, double flop_per_elem)
{
    const int iters = 4096, blocks = nsm * 8;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, 256>>>(d, 16, make_float2(1.0001f, 0.9999f));
    cudaEventRecord(e0);
    k<MODE><<<blocks, 256>>>(d, iters, make_float2(1.0001f, 0.9999f));
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double warp_inst = (double)blocks * 8 /*warps*/ * iters * 8 * inst_per_elem;
    const double cycles = ms * 1e-3 * khz * 1e3;
    printf("%-34s %8.3f ms  %6.2f warp-inst/clk/SM  %7.2f TFLOP/s  (%.2f complex-ops/clk/SM)\n", name, ms,
           warp_inst / cycles / nsm, (double)blocks * 256 * iters * 8 * flop_per_elem / (ms * 1e-3) / 1e12,
           (double)blocks * 8 * iters * 8 / cycles / nsm);
}
int main()
{
    int nsm; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    float2* d; cudaMalloc(&d, 256 * 8 * sizeof(float2)); cudaMemset(d, 0, 256 * 8 * sizeof(float2));
    run<0>("2x FADD  (complex add, scalar)", d, nsm, 2, 2);
    run<1>("1x FADD2 (complex add, packed)", d, nsm, 1, 2);
    run<7>("2x FADD  distinct regs", d, nsm, 2, 2);
    run<6>("1x FADD2 distinct regs", d, nsm, 1, 2);
    run<2>("2x FFMA  scalar", d, nsm, 2, 4);
    run<3>("1x FFMA2 packed", d, nsm, 1, 4);
    run<5>("cmul scalar (2 FMUL + 2 FFMA)", d, nsm, 4, 6);
    run<4>("cmul packed (FMUL2 + FFMA2)", d, nsm, 2, 6);
    return 0;
}
