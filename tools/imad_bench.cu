// imad_bench.cu -- measures the sustained 32-bit integer issue rates that bound the NTT path
// (R_IMAD is not in MEASURED_PEAKS.json and has to be measured on the box).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o imad_bench imad_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 4096
#define ILP 8

template <int MODE>
__global__ void __launch_bounds__(256) k(uint32_t *out, uint32_t a, uint32_t b)
{
    uint32_t v[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) v[i] = threadIdx.x + i * 7 + blockIdx.x;
    uint32_t w = a | 1u, p = b | 3u;
    const double wq = (double)w * 2.3283064365386963e-10, Cq = 4503599627370496.0 - (double)w * 1048576.0;
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (MODE == 0) v[i] = v[i] * w + p;                              // IMAD
            else if (MODE == 1) v[i] = __umulhi(v[i], w) + p;                // IMAD.HI
            else if (MODE == 2) v[i] = (v[i] + w) ^ p;                       // IADD3 + LOP3 (alu pipe)
            else if (MODE == 3) { uint32_t x = v[i] - p; v[i] = min(v[i], x) + w; }   // VIADDMNMX + IADD
            else if (MODE == 4) {                                            // Shoup mulmod: HI + 2 IMAD
                uint32_t q = __umulhi(w, v[i]);
                v[i] = a * v[i] - q * p;
            } else if (MODE == 6) {                                          // IMAD.WIDE.U32 (hi word fed back)
                uint64_t t = (uint64_t)v[i] * (uint64_t)w + (uint64_t)p;
                v[i] = (uint32_t)(t >> 32) ^ (uint32_t)t;
            } else if (MODE == 7) {                                          // Shoup mulmod with ALU filler (3 alu ops)
                uint32_t q = __umulhi(w, v[i]);
                uint32_t t = a * v[i] - q * p;
                uint32_t x = t - p;
                v[i] = (min(t, x) + w) ^ p;
            } else if (MODE == 8 || (MODE == 9 && (i & 1)) || (MODE == 10 && (i % 3) == 0)) {
                // Harvey CT butterfly whose Shoup quotient floor(y * w' / 2^32) is computed on the FP64 pipe:
                // y enters as the exact double 2^52 + y (bit pattern), wq = w' * 2^-32 and C = 2^52 - w' * 2^20 are
                // per-twiddle constants, so fma_rd(D, wq, C) = 2^52 + floor(y * w' / 2^32) exactly: same q as IMAD.HI
                uint32_t x = v[i], y = v[(i + 1) % ILP];
                uint32_t xr = min(x, x - 2 * p);
                double D = __hiloint2double(0x43300000, (int)y);
                uint32_t q = (uint32_t)__double2loint(__fma_rd(D, wq, Cq));
                uint32_t t = a * y - q * p;
                v[i] = xr + t;
                v[(i + 1) % ILP] = xr - t + 2 * p;
            } else if (MODE == 5 || MODE == 9 || MODE == 10) {               // Harvey CT butterfly on a pair
                uint32_t x = v[i], y = v[(i + 1) % ILP];
                uint32_t xr = min(x, x - 2 * p);
                uint32_t q = __umulhi(w, y);
                uint32_t t = a * y - q * p;
                v[i] = xr + t;
                v[(i + 1) % ILP] = xr - t + 2 * p;
            }
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
__global__ void __launch_bounds__(256) kd(double *out, double a, double b, uint32_t ia, uint32_t ib)
{
    double v[ILP];
    uint32_t u[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { v[i] = threadIdx.x + i * 7 + blockIdx.x; u[i] = threadIdx.x + i; }
    const double magic = 6755399441055744.0;
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (MODE == 0) v[i] = fma(v[i], a, b);                           // DFMA
            else if (MODE == 1) {                                            // FP64 mulmod (6 DP ops): x*w mod p, p ~ 2^50
                double x = v[i];
                double h = x * a;
                double l = fma(x, a, -h);
                double q = fma(x, b, magic) - magic;                         // b = a/p
                double r = fma(-q, 1125899906826241.0, h);
                v[i] = r + l;
            } else if (MODE == 2) {                                          // DFMA + independent Shoup mulmod (pipe concurrency)
                v[i] = fma(v[i], a, b);
                uint32_t q = __umulhi(ia, u[i]);
                u[i] = ib * u[i] - q * 1073692673u;
            } else if (MODE == 3) {                                          // 2 DFMA + 1 Shoup mulmod
                v[i] = fma(v[i], a, b);
                v[i] = fma(v[i], a, b);
                uint32_t q = __umulhi(ia, u[i]);
                u[i] = ib * u[i] - q * 1073692673u;
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += v[i] + (double)u[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
double rund(const char *name, double ops_per_iter_lane, double *d, int sms)
{
    const int blocks = sms * 8, threads = 256;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    kd<MODE><<<blocks, threads>>>(d, 1.0000001, 0.9999, 12345u, 777u);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0);
        kd<MODE><<<blocks, threads>>>(d, 1.0000001, 0.9999, 12345u, 777u);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    double ops = (double)blocks * threads * ITERS * ILP * ops_per_iter_lane;
    double rate = ops / (best * 1e-3);
    printf("{\"bench\": \"%s\", \"ms\": %.4f, \"Tops_per_s\": %.3f, \"ops_per_clk_per_sm_at_1965MHz\": %.1f}\n",
           name, best, rate / 1e12, rate / sms / 1.965e9);
    return rate;
}

template <int MODE>
double run(const char *name, double ops_per_iter_lane, uint32_t *d, int sms)
{
    const int blocks = sms * 8, threads = 256;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, threads>>>(d, 12345u, 1073692673u);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0);
        k<MODE><<<blocks, threads>>>(d, 12345u, 1073692673u);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    double ops = (double)blocks * threads * ITERS * ILP * ops_per_iter_lane;
    double rate = ops / (best * 1e-3);
    printf("{\"bench\": \"%s\", \"ms\": %.4f, \"Tops_per_s\": %.3f, \"ops_per_clk_per_sm_at_1965MHz\": %.1f}\n",
           name, best, rate / 1e12, rate / sms / 1.965e9);
    return rate;
}

int main()
{
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    uint32_t *d;
    cudaMalloc(&d, (size_t)prop.multiProcessorCount * 8 * 256 * 4);
    run<0>("imad_lo", 1, d, prop.multiProcessorCount);
    run<1>("imad_hi", 1, d, prop.multiProcessorCount);
    run<2>("iadd3_lop3 (2 alu ops)", 2, d, prop.multiProcessorCount);
    run<3>("viaddmnmx_iadd (2 alu ops)", 2, d, prop.multiProcessorCount);
    run<4>("shoup_mulmod (1 per iter)", 1, d, prop.multiProcessorCount);
    run<5>("harvey_ct_butterfly (1 per iter)", 1, d, prop.multiProcessorCount);
    run<6>("imad_wide_u32", 1, d, prop.multiProcessorCount);
    run<7>("shoup_mulmod + 3 alu (1 per iter)", 1, d, prop.multiProcessorCount);
    run<8>("harvey_ct_butterfly, quotient on the FP64 pipe (1 per iter)", 1, d, prop.multiProcessorCount);
    run<9>("harvey_ct_butterfly, alternating IMAD.HI / FP64 quotient (1 per iter)", 1, d, prop.multiProcessorCount);
    run<10>("harvey_ct_butterfly, one in three on the FP64 pipe (1 per iter)", 1, d, prop.multiProcessorCount);
    double *dd;
    cudaMalloc(&dd, (size_t)prop.multiProcessorCount * 8 * 256 * 8);
    rund<0>("dfma", 1, dd, prop.multiProcessorCount);
    rund<1>("fp64_mulmod_50bit (1 per iter, 6 dp ops)", 1, dd, prop.multiProcessorCount);
    rund<2>("dfma + shoup_mulmod concurrently (pairs per iter)", 1, dd, prop.multiProcessorCount);
    rund<3>("2 dfma + shoup_mulmod concurrently (triples per iter)", 1, dd, prop.multiProcessorCount);
    return 0;
}
