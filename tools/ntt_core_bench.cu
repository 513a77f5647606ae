// ntt_core_bench.cu -- the engine's own transform code (rzk_vm_exec.cuh: fwd_g1 / transpose / fwd_g2 and the inverse) in isolation:
// registers in, registers out, twiddles from shared memory, no global loads, no epilogue.  Tells how much of a kernel's time
// the transforms themselves need at a given number of warps per SM, against tools/bfly_occupancy_bench.cu (bare butterflies).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -I ring-zk_b200/csrc -o ntt_core_bench tools/ntt_core_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "rzk_vm_exec.cuh"
using namespace rzk;

template <int WHAT, int STATICP>     // WHAT 0: forward only, 1: forward + inverse
__global__ void __launch_bounds__(512, 1) k(uint32_t *out, const uint32_t *tab, int reps, uint32_t p_in, uint32_t zero)
{
    extern __shared__ __align__(16) uint32_t smem[];
    const int warps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31, hw = lane >> 4, t = lane & 15;
    uint32_t *g1 = smem, *g2 = smem + 2 * kG1Words, *bufs = g2 + 2 * kLanes * kG2Words;
    for (int i = threadIdx.x; i < 2 * kG1Words + 2 * kLanes * kG2Words; i += blockDim.x) smem[i] = tab[i];
    __syncthreads();
    uint32_t *buf = bufs + (warp * 2 + hw) * (kBufWords + 16);
    const uint32_t p = STATICP ? kStaticPrime0 : p_in, p2 = 2 * p, cap = STATICP ? 4u * kStaticPrime0 - 1u : kAddCap;
    uint32_t cur[kElems];
#pragma unroll
    for (int m = 0; m < kElems; ++m) cur[m] = threadIdx.x * 33 + m + blockIdx.x;
#pragma unroll 1
    for (int r = 0; r < reps; ++r) {
        fwd_g1(cur, g1, p, p2, zero, cap);
#pragma unroll
        for (int m = 0; m < kElems; ++m) { const int i = t + kLanes * m; buf[i + ((i >> 5) << 2)] = cur[m]; }
        __syncwarp();
        {
            const uint4 *row = reinterpret_cast<const uint4 *>(buf + 36 * t);
#pragma unroll
            for (int j = 0; j < 8; ++j) { const uint4 q = row[j]; cur[4 * j] = q.x; cur[4 * j + 1] = q.y; cur[4 * j + 2] = q.z; cur[4 * j + 3] = q.w; }
        }
        fwd_g2(cur, g2 + t * kG2Words, p, p2, zero, cap);
        __syncwarp();
        if (WHAT == 1) {
#pragma unroll
            for (int m = 0; m < kElems; ++m) cur[m] = csub(cur[m], p2);
            inv_g2(cur, g2 + (kLanes + t) * kG2Words, p, p2, zero, cap);
            uint4 *row = reinterpret_cast<uint4 *>(buf + 36 * t);
#pragma unroll
            for (int j = 0; j < 8; ++j) { uint4 w; w.x = cur[4 * j]; w.y = cur[4 * j + 1]; w.z = cur[4 * j + 2]; w.w = cur[4 * j + 3]; row[j] = w; }
            __syncwarp();
#pragma unroll
            for (int m = 0; m < kElems; ++m) { const int i = t + kLanes * m; cur[m] = buf[i + ((i >> 5) << 2)]; }
            inv_g1(cur, g1 + kG1Words, p, p2, zero, cap);
            __syncwarp();
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int m = 0; m < kElems; ++m) s += cur[m];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    (void)warps;
}

template <int WHAT, int STATICP>
void run(int warps, uint32_t *d, uint32_t *tab, int sms)
{
    const int reps = 200;
    const size_t smem = sizeof(uint32_t) * (2 * kG1Words + 2 * kLanes * kG2Words + (size_t)warps * 2 * (kBufWords + 16));
    cudaFuncSetAttribute(k<WHAT, STATICP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<WHAT, STATICP><<<sms, warps * 32, smem>>>(d, tab, reps, 1073692673u, 0u);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0);
        k<WHAT, STATICP><<<sms, warps * 32, smem>>>(d, tab, reps, 1073692673u, 0u);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double transforms = (double)sms * warps * 2 * reps * (WHAT ? 2 : 1);      // half-warp transforms
    const double bfly = transforms * 2304.0;
    printf("{\"what\": \"%s\", \"static_prime\": %d, \"warps_per_sm\": %d, \"ms\": %.4f, \"ns_per_half_warp_transform\": %.3f, \"bfly_per_clk_per_sm_at_1965MHz\": %.2f, \"err\": \"%s\"}\n",
           WHAT ? "fwd+inv" : "fwd", STATICP, warps, best, best * 1e6 / transforms, bfly / (best * 1e-3) / sms / 1.965e9, cudaGetErrorString(cudaGetLastError()));
}

int main()
{
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    uint32_t *d, *tab;
    cudaMalloc(&d, (size_t)prop.multiProcessorCount * 512 * 4);
    cudaMalloc(&tab, 4 * (2 * kG1Words + 2 * kLanes * kG2Words));
    cudaMemset(tab, 0x3b, 4 * (2 * kG1Words + 2 * kLanes * kG2Words));
    for (int w : {4, 8, 12, 16}) { run<0, 0>(w, d, tab, prop.multiProcessorCount); run<0, 1>(w, d, tab, prop.multiProcessorCount); run<1, 1>(w, d, tab, prop.multiProcessorCount); }
    return 0;
}
