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
            } else if (MODE == 5) {                                          // Harvey CT butterfly on a pair
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
    return 0;
}
