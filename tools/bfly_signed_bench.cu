// bfly_signed_bench.cu -- butterfly rate of the signed lazy forms next to the Harvey forms, as a function of resident warps per SM.
//   kind 0: Harvey CT on [0, 4p), p < 2^30            (6 instructions: 3 multiplies + 3 ALU)
//   kind 1: signed lazy CT, p < 2^27, no correction    (4 instructions: 3 multiplies + 1 ALU): q = mulhi(y, w'), t0 = y*w + x, x' = t0 - q*p, y' = 2x - x'
//   kind 2: Harvey GS on [0, 2p)                       (6 instructions)
//   kind 3: signed lazy GS, no correction              (5 instructions): s = x + y, d = x - y, y' = d*w - mulhi(d, w')*p
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bfly_signed_bench tools/bfly_signed_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITERS 2048
template <int ILP, int KIND>
__global__ void k(uint32_t *out, const uint32_t *tw, uint32_t p, uint32_t zero, uint32_t mp)
{
    uint32_t v[2 * ILP];
#pragma unroll
    for (int i = 0; i < 2 * ILP; ++i) v[i] = threadIdx.x + i * 7 + blockIdx.x;
    __shared__ uint32_t stw[64];
    if (threadIdx.x < 64) stw[threadIdx.x] = tw[threadIdx.x];
    __syncthreads();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            const uint32_t w = stw[(2 * i) & 63], wp = stw[(2 * i + 1) & 63];      // twiddles from shared memory (uniform address)
            uint32_t x = v[i], y = v[i + ILP];
            if (KIND == 0) {
                uint32_t xr = min(x, x - 2 * p);
                uint32_t q = __umulhi(wp, y);
                uint32_t t = w * y - q * p;
                v[i] = min(xr + t, 0xfffffffeu);
                v[i + ILP] = xr - t + 2 * p;
            } else if (KIND == 1) {
                const uint32_t q = (uint32_t)__mulhi((int)y, (int)wp);
                const uint32_t t0 = y * w + x;
                const uint32_t xo = q * mp + t0;
                v[i] = xo;
                v[i + ILP] = x + x - xo;
            } else if (KIND == 2) {
                uint32_t s = min(x + y, 0xfffffffeu);
                s = min(s, s - 2 * p);
                const uint32_t d = x - y + 2 * p;
                const uint32_t q = __umulhi(wp, d);
                v[i] = s;
                v[i + ILP] = w * d - q * p;
            } else if (KIND >= 7 && KIND <= 10) {
                // signed lazy CT whose Shoup quotient floor(y * w' / 2^32) comes from the FP64 pipe, bit-identical to mulhi:
                //   Ys = y * 2^-32 exactly (pair lo = y ^ 2^31, hi = 0x41300000 is 2^20 + (y + 2^31) 2^-32; minus 2^20 + 1/2),
                //   Wd = w' exactly (pair lo = w' ^ 2^31, hi = 0x43300000, minus 2^52 + 2^31), q = low word of fma.rd(Ys, Wd, 1.5 * 2^52)
                // KIND 7: every butterfly; 8: every second; 9: two of three; 10: every butterfly, w' converted once per 4 butterflies
                const bool fp = KIND == 7 || KIND == 10 || (KIND == 8 && (i & 1)) || (KIND == 9 && (i % 3) != 0);
                uint32_t q;
                if (fp) {
                    const double ys = __hiloint2double(0x41300000, (int)(y ^ 0x80000000u)) - 1048576.5;
                    const uint32_t wpc = (KIND == 10) ? stw[(2 * (i & ~3) + 1) & 63] : wp;
                    const double wd = __hiloint2double(0x43300000, (int)(wpc ^ 0x80000000u)) - 4503601774854144.0;
                    q = (uint32_t)__double2loint(__fma_rd(ys, wd, 6755399441055744.0));
                } else {
                    q = (uint32_t)__mulhi((int)y, (int)wp);
                }
                const uint32_t t0 = y * w + x;
                const uint32_t xo = q * mp + t0;
                v[i] = xo;
                v[i + ILP] = x + x - xo;
            } else if (KIND == 11) {      // (diagnostic, wrong arithmetic) signed GS whose quotient does not wait for the difference
                const uint32_t sm = x + y, d = x - y;
                const uint32_t q = (uint32_t)__mulhi((int)x, (int)wp);
                v[i] = sm;
                v[i + ILP] = q * mp + d * w;
            } else if (KIND == 12) {      // (diagnostic) signed GS without the sum output: 4 instructions
                const uint32_t d = x - y;
                const uint32_t q = (uint32_t)__mulhi((int)d, (int)wp);
                v[i] = y;
                v[i + ILP] = q * mp + d * w;
            } else if (KIND == 13) {      // signed GS, product from x*w - y*w (no wait on d for the low product): 6 instructions
                const uint32_t sm = x + y, d = x - y;
                const uint32_t q = (uint32_t)__mulhi((int)d, (int)wp);
                const uint32_t t1 = x * w;
                const uint32_t t2 = t1 - y * w;
                v[i] = sm;
                v[i + ILP] = q * mp + t2;
            } else if (KIND == 4) {       // signed GS, adds forced onto the ALU pipe (add-and-max with a bound that never binds)
                const uint32_t s = (uint32_t)max((int)(x + y), -0x7fffffff);
                const uint32_t d = (uint32_t)max((int)(x - y), -0x7fffffff);
                const uint32_t q = (uint32_t)__mulhi((int)d, (int)wp);
                v[i] = s;
                v[i + ILP] = q * mp + d * w;
            } else if (KIND == 5) {       // unsigned lazy GS without correction: d = x - y + K p
                const uint32_t s = min(x + y, 0xfffffffeu);
                const uint32_t d = x - y + zero;
                const uint32_t q = __umulhi(d, wp);
                v[i] = s;
                v[i + ILP] = q * mp + d * w;
            } else if (KIND == 6) {       // signed CT with the last add forced onto the ALU pipe
                const uint32_t q = (uint32_t)__mulhi((int)y, (int)wp);
                const uint32_t t0 = y * w + x;
                const uint32_t xo = q * mp + t0;
                v[i] = xo;
                v[i + ILP] = (uint32_t)max((int)(x + x - xo), -0x7fffffff);
            } else {
                const uint32_t s = x + y;
                const uint32_t d = x - y;
                const uint32_t q = (uint32_t)__mulhi((int)d, (int)wp);
                v[i] = s;
                v[i + ILP] = q * mp + d * w;
            }
        }
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
template <int ILP, int KIND>
void run(int warps_per_sm, uint32_t *d, uint32_t *tw, int sms)
{
    const int threads = warps_per_sm >= 8 ? 256 : warps_per_sm * 32, blocks = sms * (warps_per_sm * 32 / threads);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const uint32_t p = (KIND & 1) || KIND >= 7 ? 68718593u : 1073692673u;   // (the rate does not depend on the value)
    k<ILP, KIND><<<blocks, threads>>>(d, tw, p, 0u, 0u - p);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0);
        k<ILP, KIND><<<blocks, threads>>>(d, tw, p, 0u, 0u - p);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    double ops = (double)blocks * threads * ITERS * ILP;
    double rate = ops / (best * 1e-3);
    const char *names[14] = {"harvey_ct", "signed_ct", "harvey_gs", "signed_gs", "signed_gs_alu", "unsigned_lazy_gs", "signed_ct_alu", "signed_ct_fp64q", "signed_ct_fp64q_1of2", "signed_ct_fp64q_2of3", "signed_ct_fp64q_w4", "diag_gs_q_indep", "diag_gs_no_sum", "signed_gs_split_product"};
    printf("{\"kind\": \"%s\", \"warps_per_sm\": %d, \"ilp\": %d, \"ms\": %.4f, \"Tbfly_per_s\": %.3f, \"bfly_per_clk_per_sm_at_1965MHz\": %.2f}\n",
           names[KIND], warps_per_sm, ILP, best, rate / 1e12, rate / sms / 1.965e9);
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
    for (int w : {8, 16}) { run<16, 11>(w, d, tw, sms); run<16, 12>(w, d, tw, sms); run<16, 13>(w, d, tw, sms); run<8, 3>(w, d, tw, sms); run<8, 1>(w, d, tw, sms); run<16, 7>(w, d, tw, sms); run<16, 8>(w, d, tw, sms); run<16, 9>(w, d, tw, sms); run<16, 10>(w, d, tw, sms); run<16, 0>(w, d, tw, sms); run<16, 1>(w, d, tw, sms); run<16, 2>(w, d, tw, sms); run<16, 3>(w, d, tw, sms); run<16, 4>(w, d, tw, sms); run<16, 5>(w, d, tw, sms); run<16, 6>(w, d, tw, sms); }
    return 0;
}
