// bfly_occupancy_bench.cu -- Harvey CT butterfly rate as a function of resident warps per SM and of the independent butterflies
// per thread (the engine's kernels run 16 warps per SM -- 128 registers per thread -- with 16 independent butterflies per stage).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bfly_occupancy_bench bfly_occupancy_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITERS 2048
template <int ILP, int TW>
__global__ void k(uint32_t *out, const uint32_t *tw, uint32_t p)
{
    uint32_t v[2 * ILP];
#pragma unroll
    for (int i = 0; i < 2 * ILP; ++i) v[i] = threadIdx.x + i * 7 + blockIdx.x;
    __shared__ uint32_t stw[64];
    if (threadIdx.x < 64) stw[threadIdx.x] = tw[threadIdx.x];
    __syncthreads();
    uint32_t w = tw[0] | 1u, wp = tw[1] | 3u;
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (TW) { w = stw[(2 * i) & 63]; wp = stw[(2 * i + 1) & 63]; }      // twiddles from shared memory (uniform address)
            uint32_t x = v[i], y = v[i + ILP];
            uint32_t xr = min(x, x - 2 * p);
            uint32_t q = __umulhi(wp, y);
            uint32_t t = w * y - q * p;
            v[i] = min(xr + t, 0xfffffffeu);
            v[i + ILP] = xr - t + 2 * p;
        }
        // next "stage": rotate the pairing so that the chain depends on both outputs
        uint32_t t0 = v[0];
#pragma unroll
        for (int i = 0; i < 2 * ILP - 1; ++i) v[i] = v[i + 1];
        v[2 * ILP - 1] = t0;
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 2 * ILP; ++i) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP, int TW>
void run(int warps_per_sm, uint32_t *d, uint32_t *tw, int sms)
{
    const int threads = warps_per_sm >= 8 ? 256 : warps_per_sm * 32, blocks = sms * (warps_per_sm * 32 / threads);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<ILP, TW><<<blocks, threads>>>(d, tw, 1073692673u);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0);
        k<ILP, TW><<<blocks, threads>>>(d, tw, 1073692673u);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    double ops = (double)blocks * threads * ITERS * ILP;
    double rate = ops / (best * 1e-3);
    printf("{\"warps_per_sm\": %d, \"ilp\": %d, \"twiddles_from_smem\": %d, \"ms\": %.4f, \"Tbfly_per_s\": %.3f, \"bfly_per_clk_per_sm_at_1965MHz\": %.2f}\n",
           warps_per_sm, ILP, TW, best, rate / 1e12, rate / sms / 1.965e9);
}
int main()
{
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    uint32_t *d, *tw;
    cudaMalloc(&d, (size_t)prop.multiProcessorCount * 64 * 32 * 4);
    cudaMalloc(&tw, 256);
    cudaMemset(tw, 0x5a, 256);
    const int sms = prop.multiProcessorCount;
    for (int w : {4, 8, 16, 32, 64}) { run<8, 0>(w, d, tw, sms); run<16, 0>(w, d, tw, sms); run<16, 1>(w, d, tw, sms); }
    return 0;
}
