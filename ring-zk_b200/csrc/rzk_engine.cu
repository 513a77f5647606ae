// rzk_engine.cu -- sm_100a kernels and the C ABI of include/ringzk_b200.h.
//
// One persistent kernel template (rzk_vm_kernel) evaluates a polynomial-op program
// (rzk_vm.h / rzk_programs.h) for every batch item: one half warp per item, 32
// coefficients per lane in registers, forward / inverse negacyclic NTTs over 1..3
// auxiliary primes with shared-memory transposes, key images and twiddles resident in
// shared memory, Garner CRT and the centred reduction mod q in the epilogue.
// This file contains no CPU fallback: every entry point fails with RZK_ERR_CUDA when the
// device is unavailable.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <thread>
#include <string>
#include <type_traits>
#include <vector>

#include "../../include/ringzk_b200.h"
#include "rzk_vm.h"

#include "rzk_vm_exec.cuh"
#include "rzk_sparse.cuh"
#include "rzk_sample.cuh"
#include "rzk_programs.h"
#include "rzk_tables.h"

using namespace rzk;

// ------------------------------------------------------------------------------ kernels

// warps per CTA: SPLIT fits 16 warps in 128 registers; the one-prime SEQ kernel finishes 32 coefficients per
// lane in the epilogue and gets a larger register budget; the three-prime SEQ kernels finish them four at a time.
template <int NP, int MODE>
#ifndef RZK_SPLIT_WARPS
#define RZK_SPLIT_WARPS 16
#endif
#ifndef RZK_SEQ3_WARPS
#define RZK_SEQ3_WARPS 16      // three-prime programs: chunked epilogue + residue stash in global memory (rzk_vm_exec.cuh ChunkedEpi)
#endif
struct VmCfg { static constexpr int kMaxWarps = !mode_seq(MODE) ? RZK_SPLIT_WARPS : (NP == 1 ? 12 : RZK_SEQ3_WARPS); };

// does a compile-time program multiply by the resident key (OP_MACK)?  Programs that do not (the three-prime product
// sums) leave the key images out of shared memory, which is what lets their two-accumulator form keep 16 warps.
template <class SP>
constexpr bool sp_uses_key()
{
    if constexpr (std::is_void<SP>::value) return true;
    else {
        for (int i = 0; i < kMaxOps && SP::prog.ops[i].code != OP_END; ++i)
            if (SP::prog.ops[i].code == OP_MACK) return true;
        return false;
    }
}

// does a compile-time program keep accumulator 1 in the global stash region (SP::kAcc1Global, rzk_programs.h)?
template <class SP, class = void>
struct SpAcc1Global { static constexpr bool value = false; };
template <class SP>
struct SpAcc1Global<SP, std::enable_if_t<!std::is_void<SP>::value && SP::kAcc1Global>> { static constexpr bool value = true; };

template <int NP, int MODE, bool KEY = true>
struct VmSmem {
    static constexpr int kKP = mode_sk(MODE) ? 2 * kKeyPolys : kKeyPolys;   // key images per prime
    static constexpr int kG1 = NP * 2 * kG1Words;
    static constexpr int kG2 = NP * 2 * kLanes * kG2Words;
    static constexpr int kKey = KEY ? NP * kKP * 2 * kPadWords : 0;
    static constexpr int kTw = (mode_signed(MODE) && RZK_INV_DIT) ? NP * kTwistWords : 0;   // output twists of the signed slots' inverse
    static constexpr int kTables = (kG1 + kG2 + kKey + kTw + 3) / 4 * 4;
    static size_t bytes(int warps, uint32_t hw_words) { return sizeof(uint32_t) * ((size_t)kTables + (size_t)warps * 2 * hw_words); }
};

// ------------------------------------------------------------------------------ TMA staging of the resident tables
// Twiddle and key tiles are brought into shared memory with 1-D bulk copies (cp.async.bulk, SASS UBLKCP) issued by
// one thread and tracked by an mbarrier; every thread then waits on the barrier's phase 0.  Sizes and addresses are
// multiples of 16 bytes by construction (kG1Words, kG2Words, kPadWords; cudaMalloc bases).

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tma_init(uint64_t *bar)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void tma_expect(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void tma_load(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void tma_wait(uint64_t *bar)
{
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "WAIT_%=:\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t"
                 "@!p bra WAIT_%=;\n\t}" ::"r"(smem_u32(bar)) : "memory");
}

// stages the twiddles and the key images of the NP primes of a launch
template <int NP, int MODE, bool KEY = true>
__device__ __forceinline__ void stage_int_tables(const VmLaunch &K, uint32_t *s_g1, uint32_t *s_g2, uint32_t *s_key, uint32_t *s_tw,
                                                 uint64_t *bar, bool issue)
{
    using S = VmSmem<NP, MODE, KEY>;
    constexpr uint32_t g1b = 2 * kG1Words * 4, g2b = 2 * kLanes * kG2Words * 4, keyb = KEY ? S::kKP * 2 * kPadWords * 4 : 0;
    static_assert(g1b % 16 == 0 && g2b % 16 == 0 && keyb % 16 == 0, "bulk copies move multiples of 16 bytes");
    if (issue) {
        constexpr uint32_t twb = S::kTw ? kTwistWords * 4 : 0;
        tma_expect(bar, NP * (g1b + g2b + keyb + twb));
        for (int i = 0; i < NP; ++i) {
            const uint32_t slot = K.pc[i].slot;
            tma_load(s_g1 + i * 2 * kG1Words, K.g1tab + (size_t)slot * (2 * kG1Words), g1b, bar);
            tma_load(s_g2 + i * (2 * kLanes * kG2Words), K.g2tab + (size_t)slot * (2 * kLanes * kG2Words), g2b, bar);
            if constexpr (KEY) tma_load(s_key + i * (S::kKP * 2 * kPadWords), K.keytab + (size_t)i * (S::kKP * 2 * kPadWords), keyb, bar);
            if constexpr (S::kTw != 0) tma_load(s_tw + i * kTwistWords, K.twist[i], kTwistWords * 4, bar);
        }
    }
}

// Persistent kernel: blockDim.x / 32 warps per CTA (as many as the program's shared-memory needs
// allow, up to 16), one CTA per SM, each warp loops over its items.
// SP = void: the generic kernel decodes K.ops at run time.  SP = a compile-time program descriptor
// (rzk_programs.h): the same lane code, unrolled from the constexpr program.
template <int NP, int MODE, class SP = void>
__global__ void __launch_bounds__(VmCfg<NP, MODE>::kMaxWarps * 32, 1) rzk_vm_kernel(const __grid_constant__ VmLaunch K)
{
    constexpr bool SPLIT = !mode_seq(MODE);      // one warp per item
    extern __shared__ __align__(16) uint32_t smem[];
    constexpr bool KEY = sp_uses_key<SP>();
    using S = VmSmem<NP, MODE, KEY>;
    if (K.any_item && *K.any_item == 0) return;     // masked fallback launch with nothing to redo
    uint32_t *s_g1 = smem;
    uint32_t *s_g2 = s_g1 + S::kG1;
    uint32_t *s_key = s_g2 + S::kG2;
    uint32_t *s_tw = s_key + S::kKey;
    uint32_t *s_hw = smem + S::kTables;
    const int nthreads = blockDim.x, warps = nthreads >> 5;

    // stage the twiddles and the key image of every prime of this launch (TMA bulk copies)
    __shared__ __align__(8) uint64_t s_bar;
    if (threadIdx.x == 0) tma_init(&s_bar);
    __syncthreads();
    stage_int_tables<NP, MODE, KEY>(K, s_g1, s_g2, s_key, s_tw, &s_bar, threadIdx.x == 0);
    tma_wait(&s_bar);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, hw = lane >> 4, t = lane & 15;
    uint32_t *mine = s_hw + (warp * 2 + hw) * K.hw_words;
    LaneCtx ctx;
    ctx.buf = mine;
    ctx.slot = mine + K.off_slot;
    ctx.slot_hw[0] = s_hw + (warp * 2 + 0) * K.hw_words + K.off_slot;
    ctx.slot_hw[1] = s_hw + (warp * 2 + 1) * K.hw_words + K.off_slot;
    ctx.stash = K.gstash ? K.gstash + (size_t)((blockIdx.x * warps + warp) * 2 + hw) * K.stash_words : mine + K.off_stash;
    // (decided at compile time for the unrolled programs, so that their accumulator accesses stay LDS / STS or LDG / STG)
    if constexpr (SpAcc1Global<SP>::value) ctx.acc1 = ctx.stash;
    else if constexpr (std::is_void<SP>::value) ctx.acc1 = K.acc1_global ? ctx.stash : mine + K.off_acc1;
    else ctx.acc1 = mine + K.off_acc1;
    ctx.red = SPLIT ? s_hw + (warp * 2) * K.hw_words : mine;
    ctx.ridx = SPLIT ? lane : t;
    ctx.g1 = s_g1;
    ctx.g2 = s_g2;
    ctx.key = s_key;
    ctx.twist = s_tw;
    ctx.t = t;
    ctx.hw = hw;

    const uint32_t units = SPLIT ? 1u : 2u;                          // items per warp
    const uint32_t per_grid = gridDim.x * warps * units;
    const uint32_t first = (blockIdx.x * warps + warp) * units + (SPLIT ? 0u : (uint32_t)hw);
    const uint32_t iters = (K.n_items + per_grid - 1) / per_grid;
    Lane L;
    uint32_t pp_count = 0;
    ctx.pp_count = &pp_count;
    pp_start(K);
#pragma unroll 1
    for (uint32_t it = 0; it < iters; ++it) {
        const uint32_t item = first + it * per_grid;
        ctx.active = item < K.n_items;
        ctx.item = ctx.active ? item : K.n_items - 1;
        if (K.item_mask) {                          // redo only the flagged items (warp-uniform skip)
            ctx.active = ctx.active && K.item_mask[K.mask_div > 1 ? ctx.item / K.mask_div : ctx.item] != 0;
            if (!__any_sync(0xffffffffu, ctx.active)) continue;
        }
        // prefetch the input rows of the item this warp handles next into L2: lane k issues one bulk prefetch
        // (cp.async.bulk.prefetch.L2) for the whole row group of input stream k
        const uint32_t next = item + per_grid;
        if (next < K.n_items && (uint32_t)ctx.ridx < K.n_prefetch) {
            const Stream st = K.st[K.prefetch[ctx.ridx]];
            const uint32_t esz = (st.dtype == DT_I8) ? 1u : 4u;
            const char *base = reinterpret_cast<const char *>(st.base) + (size_t)next * (st.stride * kN * esz);
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(base), "r"(K.prefetch_polys[ctx.ridx] * kN * esz) : "memory");
        }
        if constexpr (std::is_void<SP>::value) vm_run_item<NP, MODE>(K, &L, &ctx);
        else vm_run_static<SP>(K, &L, &ctx);
    }
    pp_finish(K, pp_count);
}

// ---- response z = y + d*r as signed rotations (rzk_sparse.cuh): one warp per item, byte accumulators ----
// (tried: 64 registers, two CTAs per SM, y loaded after the loop: 178 M/s vs 186 M/s -- the shared-memory pipe is the limit)
__global__ void __launch_bounds__(512, 1) rzk_respond_sparse_kernel(const __grid_constant__ SparseLaunch K)
{
    extern __shared__ __align__(16) uint32_t smem[];
    const int warps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    LaneCtxS ctx;
    ctx.sm = smem + (size_t)warp * kSpWarpWords;
    ctx.lane = lane;
    const uint32_t per_grid = gridDim.x * warps;
    const uint32_t first = blockIdx.x * warps + warp;
    const uint32_t iters = (K.n_items + per_grid - 1) / per_grid;
#pragma unroll 1
    for (uint32_t it = 0; it < iters; ++it) {
        const uint32_t item = first + it * per_grid;
        ctx.active = item < K.n_items;
        ctx.item = ctx.active ? item : K.n_items - 1;
        const uint32_t next = item + per_grid;
        if (next < K.n_items) {                     // L2 prefetch of the next item's rows: y 6 KB, r 1.5 KB, d 0.5 KB
            // (one 128-byte line per lane; measured slightly faster here than three bulk prefetches)
            const char *yb = reinterpret_cast<const char *>(K.y) + (size_t)next * 3 * kN * 4;
            const char *rb = reinterpret_cast<const char *>(K.r) + (size_t)next * 3 * kN;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(yb + lane * 128));
            if (lane < 16) asm volatile("prefetch.global.L2 [%0];" ::"l"(yb + 4096 + lane * 128));
            else if (lane < 28) asm volatile("prefetch.global.L2 [%0];" ::"l"(rb + (lane - 16) * 128));
            else asm volatile("prefetch.global.L2 [%0];" ::"l"(K.d + (size_t)(next / K.d_div) * kN + (lane - 28) * 128));
        }
        sparse_respond_item(K, &ctx);
    }
}

// ---- optional on-device samplers (rzk_sample.cuh): one thread per Philox block / per item ----
__global__ void rzk_sample_small_kernel(size_t n_polys, uint32_t b, uint32_t tag, uint32_t k0, uint32_t k1, int8_t *__restrict__ out)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x, total = n_polys * (kN / 4);
    for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += stride) {
        const uint64_t poly = g / (kN / 4);
        const uint32_t i0 = (uint32_t)(g % (kN / 4)) * 4u;
        char4 v;
        v.x = (signed char)sample_small_coeff(poly, i0 + 0, b, tag, k0, k1);
        v.y = (signed char)sample_small_coeff(poly, i0 + 1, b, tag, k0, k1);
        v.z = (signed char)sample_small_coeff(poly, i0 + 2, b, tag, k0, k1);
        v.w = (signed char)sample_small_coeff(poly, i0 + 3, b, tag, k0, k1);
        reinterpret_cast<char4 *>(out)[g] = v;
    }
}

__global__ void rzk_sample_gaussian_kernel(size_t n_polys, double sigma, uint32_t tag, uint32_t k0, uint32_t k1, int32_t *__restrict__ out)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x, total = n_polys * (kN / 2);
    for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += stride) {
        int32_t v[2];
        sample_gaussian_pair(g / (kN / 2), (uint32_t)(g % (kN / 2)), sigma, tag, k0, k1, v);
        reinterpret_cast<int2 *>(out)[g] = make_int2(v[0], v[1]);
    }
}

__global__ void rzk_sample_challenge_kernel(size_t n_items, uint32_t kappa, uint32_t tag, uint32_t k0, uint32_t k1, int8_t *__restrict__ out)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t it = (size_t)blockIdx.x * blockDim.x + threadIdx.x; it < n_items; it += stride) {
        int8_t *d = out + it * kN;
        for (int i = 0; i < kN / 16; ++i) reinterpret_cast<uint4 *>(d)[i] = make_uint4(0, 0, 0, 0);
        sample_challenge_item(it, kN, kappa, tag, k0, k1, d);
    }
}

// bitmap[i>>3] bit (i&7) = (flags[i] & FLAG_FAIL) == 0 ; range_any |= FLAG_RANGE bits
__global__ void rzk_flags_to_bitmap_kernel(size_t n, const uint32_t *__restrict__ flags,
                                           uint8_t *__restrict__ bitmap, uint32_t *range_any)
{
    const size_t byte = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t nbytes = (n + 7) / 8;
    if (byte >= nbytes) return;
    uint32_t bits = 0, rng = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const size_t i = byte * 8 + j;
        if (i < n) {
            const uint32_t f = flags[i];
            bits |= ((f & FLAG_FAIL) ? 0u : 1u) << j;
            rng |= f & FLAG_RANGE;
        }
    }
    bitmap[byte] = (uint8_t)bits;
    if (rng && range_any) atomicOr(range_any, rng);
}

// Last step of a product sum that was cut into `segs` segments per instance (dev_mulsum, small batches): the segment
// results are canonical residues; out = centred(sum of the segments - sub0 - sub1), stored or compared with zero.
__global__ void rzk_partial_reduce_kernel(size_t n_coeffs, uint32_t segs, const int32_t *__restrict__ part,
                                          const int32_t *__restrict__ sub0, const int32_t *__restrict__ sub1,
                                          int32_t *__restrict__ out, uint32_t *__restrict__ flags, uint32_t one_flag_word, int64_t q)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_coeffs) return;
    const size_t inst = i / kN, n = i % kN;
    int64_t acc = 0;
    for (uint32_t sgm = 0; sgm < segs; ++sgm) acc += part[(inst * segs + sgm) * kN + n];
    if (sub0) acc -= sub0[i];
    if (sub1) acc -= sub1[i];
    const int64_t half = (q - 1) / 2;
    int64_t r = acc % q;
    if (r > half) r -= q;
    else if (r < -half) r += q;
    if (out) out[i] = (int32_t)r;
    else if (r != 0) atomicOr(&flags[one_flag_word ? 0 : inst], FLAG_FAIL);
}

// 2-bit packed randomness (rzk_commit_batch_r2): one thread per 32-bit word = 16 coefficients, two's complement fields
// (0, 1, -2, -1), coefficient i of a row in bits 2(i & 3) .. of byte i >> 2; written as 16 int8 (one 128-bit store)
__global__ void rzk_unpack_r2_kernel(size_t nwords, const uint32_t *__restrict__ src, uint4 *__restrict__ dst)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nwords; i += stride) {
        const uint32_t w = src[i];
        uint32_t o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t b = (w >> (8 * k)) & 0xffu;
            // spread the four fields to four bytes, then sign-extend every 2-bit field: (f ^ 2) - 2
            const uint32_t f = (b & 3u) | ((b & 0xcu) << 6) | ((b & 0x30u) << 12) | ((b & 0xc0u) << 18);
            o[k] = __vsub4(f ^ 0x02020202u, 0x02020202u);
        }
        dst[i] = make_uint4(o[0], o[1], o[2], o[3]);
    }
}

// ZqI64::from(i64) for whole arrays: any representative -> canonical centred i32
__global__ void rzk_pack_i64_kernel(size_t n, const int64_t *__restrict__ src, int32_t *__restrict__ dst, int64_t q)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const int64_t half = (q - 1) / 2;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        int64_t r = src[i] % q;
        if (r > half) r -= q;
        else if (r < -half) r += q;
        dst[i] = (int32_t)r;
    }
}

__global__ void rzk_unpack_i64_kernel(size_t n, const int32_t *__restrict__ src, int64_t *__restrict__ dst)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = src[i];
}

// ------------------------------------------------------------------------------ engine

namespace {

constexpr int kPipe = 4;

struct PipeSlot {
    cudaStream_t stream = nullptr;
    char *arena = nullptr;
    size_t cap = 0;
};

thread_local std::string g_create_err;

}  // namespace

struct rzk_engine {
    rzk_params P;
    int device = 0;
    int num_sms = 0;
    std::string err;
    uint32_t *d_g1tab = nullptr;
    uint32_t *d_g2tab = nullptr;
    uint32_t *d_keytab = nullptr;
    uint32_t *d_keytab2 = nullptr;  // split-key images (lo/hi) for prime slot 0, [6][2][576]
    uint32_t *d_keytab3 = nullptr;  // split-key images for the small prime of MODE_SPLITKEY_S (slot kSignedSlot), signed Shoup form
    uint32_t *d_twist = nullptr;    // output twists of the signed slots' decimation-in-time inverse, [kNumPrimeSlots][psi^-i, psi^-i R N^-1][kTwistWords]
    uint32_t *d_gstash[kPipe + 1] = {};   // residue stash of the three-prime programs, [SM][warp][half warp][kStashWordsMax]:
                                    // one per pipeline stream (their kernels may overlap) + one for the `_dev` entry points
    int32_t *d_partial[kPipe + 1] = {};   // segment results of product sums cut into segments (small batches), per stream as above
    uint32_t *d_need = nullptr;     // hand-over words of dev_respond for the `_dev` entry points
    size_t need_cap = 0;
    void *d_wire_toks = nullptr;    // token list of the wire-format calls (rzk_wire.cuh)
    size_t wire_toks_cap = 0;
    uint8_t *d_fs_prefix = nullptr; // transcript prefix of the Fiat-Shamir entry points (rzk_fs.cuh)
    size_t fs_prefix_cap = 0;
    uint32_t *d_misc = nullptr;     // [0] range word, [1] dummy flags word
    uint32_t *h_range = nullptr;    // pinned host copy of the range word (single-chunk calls)
    bool has_key = false;
    bool generic_commit = false;    // b > kSplitKeyLimit: every commitment runs the two-prime program (exact for any int8 r)
    bool small_commit = false;      // b == 1 (Params::default()): the split-key program modulo the small prime, signed lazy arithmetic
    uint64_t sigma = 0, cbound = 0, vbound = 0;
    uint32_t small_lim = 0;
    PipeSlot pipe[kPipe];
    char *scratch = nullptr;        // scratch of the `_dev` entry points
    size_t scratch_cap = 0;
    uint64_t launches = 0;
    uint32_t chunk_items = 8192;    // host pipeline: items per chunk (RZK_CHUNK_ITEMS)
    uint32_t chunk_ramp = 1;        // host pipeline: a batch of several chunks starts with smaller ones (RZK_CHUNK_RAMP=0: off)
    // RZK_TEST_LOWERING: alternative lowerings of the same phases, for the differential tests (tools/soak.py) -- results are
    // identical in every setting.  Comma-separated tokens:
    uint32_t no_static = 0;         //   generic    the runtime-decoded interpreter instead of the compile-time programs
    uint32_t no_sparse = 0;         //   nosparse   responses through the one-prime NTT program only (no rotation kernel)
    uint32_t no_segments = 0;       //   nosegments never cut a small product sum into segments
    uint32_t no_dimg = 0;           //   nodimg     every Sum verify item transforms its challenge itself
    uint32_t no_fuse = 0;           //   nofuse     the Sum prover's two product sums as two launches
    uint32_t no_rot = 0;            //   norot      the verify programs multiply c1*d, c2*d in the NTT domain (no rotation sums)
    uint32_t no_rot_w = 0;          //   norotw     only Open verify uses the rotation sum; Linear / Sum first equations multiply in the NTT domain
    // RZK_TUNE (developer A/B timing, "name=value,..."): the settings below are the measured best
    uint32_t static_respond = 0;
    uint32_t mulsum2_pp = 2;        //   mulsum2_pp phase mixing of the two-accumulator product-sum program (0 / 9 = off)
    uint32_t commit_pp = 2;         //   commit_pp  phase mixing of the split-key commitment program (0 / 9 = off)
    uint32_t mulsum_small = 1;      //   mulsum_small  0: product sums of up to 64 terms use the 30-bit primes too (A/B)
    uint32_t wave_fit = 1;          //   wave_fit   fit the warps per CTA to the wave count of the batch (launch_vm); 0 = always the maximum
    uint32_t commit_small = 1;      //   commit_small  0: engines with b = 1 use the 30-bit split-key program too (A/B)
    uint32_t ld128 = 0;             //   ld128      OP_FWD fetches int32 rows with 128-bit loads + a shared-memory redistribution (A/B)
    uint32_t verify_pp = 22;        //   verify_pp  phase mixing of the Open verify program with the rotation sum (two staggered groups)
    uint32_t verify_w_pp = 21;      //   verify_w_pp  the same for the Linear / Sum first-equation program with two rotation sums (+2 % on Linear verify)
    uint32_t pp_mode = 0;           //   pp         phase mixing between CTA halves for every static program
    uint32_t cta_sync = 8;          //   cta_sync   lock-step barriers (rzk_vm_exec.cuh cta_lockstep): 8 = one per segment, 1 = per transform, 0 = off
};

namespace {

int fail(rzk_engine *e, int code, const std::string &msg)
{
    if (e) e->err = msg;
    else g_create_err = msg;
    return code;
}

#define RZK_CUDA(e, call)                                                                       \
    do {                                                                                        \
        cudaError_t err__ = (call);                                                             \
        if (err__ != cudaSuccess)                                                               \
            return fail((e), RZK_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(err__)); \
    } while (0)

uint64_t isqrt64(uint64_t v)
{
    uint64_t x = 0;
    for (uint64_t bit = 1ull << 31; bit; bit >>= 1)
        if ((x + bit) * (x + bit) <= v) x += bit;
    return x;
}

struct Guard {
    int prev = -1;
    explicit Guard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
    ~Guard() { if (prev >= 0) cudaSetDevice(prev); }
};

void fill_common(const rzk_engine *e, VmLaunch &K, int np, uint32_t n_items, uint32_t flag_div, uint32_t *flags, bool small_primes = false)
{
    static const int slots30[3] = {0, 1, 2}, slots26[3] = {kSignedSlot, kSignedSlot + 1, kSignedSlot + 2};
    const int *slots = small_primes ? slots26 : slots30;
    const uint64_t q = (uint64_t)e->P.q;
    for (int i = 0; i < np; ++i) K.pc[i] = make_prime_consts(slots[i]);
    K.crt = make_crt_consts(slots, np, q);
    K.q = (uint32_t)q;
    K.kqh = (q << 29) + (q - 1) / 2;
    K.qd = (double)q; K.qinvd = 1.0 / (double)q;
    K.p0d = (double)K.pc[0].p; K.p0qinvd = (double)K.pc[0].p / (double)q; K.p0invd = 1.0 / (double)K.pc[0].p;
    K.m30 = (uint32_t)((1ull << 62) / q);
    K.norm_abs_lim[0] = (uint32_t)e->cbound; K.norm_sq_lim[0] = (e->cbound + 1) * (e->cbound + 1) - 1;
    K.norm_abs_lim[1] = (uint32_t)e->vbound; K.norm_sq_lim[1] = (e->vbound + 1) * (e->vbound + 1) - 1;
    K.small_lim = e->small_lim;
    K.ld128 = small_primes ? 0u : e->ld128;
    K.n_items = n_items;
    K.np = (uint32_t)np;
    if (flags) { K.flags = flags; K.flag_div = flag_div; }
    else { K.flags = e->d_misc + 1; K.flag_div = 0xFFFFFFFFu; }
    K.g1tab = e->d_g1tab;
    K.g2tab = e->d_g2tab;
    K.keytab = e->d_keytab;
}

void set_stream(VmLaunch &K, int i, const void *base, uint32_t stride, uint32_t dtype, uint32_t div = 1)
{
    K.st[i].base = base; K.st[i].stride = stride; K.st[i].dtype = dtype; K.st[i].pad_ = 0;
    set_stream_div(K.st[i], div);
}

constexpr uint32_t kStashWordsMax = 2 * (kMaxPrimes - 1) * kSlotWords;    // up to two stashed outputs per three-prime program (prog_mulsum2)

template <int NP, int MODE, class SP = void>
int launch_vm(rzk_engine *e, VmLaunch &K, cudaStream_t s, uint32_t pp_program = 0)
{
    if (K.n_items == 0) return RZK_OK;
    if (K.n_items >= (1u << 28)) return fail(e, RZK_ERR_INVALID, "more than 2^28 items in one launch");
    constexpr bool SPLIT = !mode_seq(MODE);
    auto kern = rzk_vm_kernel<NP, MODE, SP>;
    if (SpAcc1Global<SP>::value != (K.acc1_global != 0) && !std::is_void<SP>::value)
        return fail(e, RZK_ERR_INVALID, "program and kernel disagree on where accumulator 1 lives");
    if ((!SPLIT && NP > 1) || K.acc1_global) {
        // residues of the earlier primes (or the second accumulator of a program with a rotation sum) wait in global memory
        // (one region per resident half warp, reused item after item)
        int si = kPipe;
        for (int i = 0; i < kPipe; ++i) if (e->pipe[i].stream == s) si = i;
        if (!e->d_gstash[si])
            RZK_CUDA(e, cudaMalloc(&e->d_gstash[si], sizeof(uint32_t) * (size_t)e->num_sms * VmCfg<NP, MODE>::kMaxWarps * 2 * kStashWordsMax));
        K.gstash = e->d_gstash[si];
    }
    // (the product-sum programs, MODE_SEQ_S, take the twist that also carries R N^-1)
    for (int i = 0; i < kMaxPrimes; ++i)
        K.twist[i] = e->d_twist + ((size_t)(K.pc[i].slot < (uint32_t)kNumPrimeSlots ? K.pc[i].slot : 0u) * 2 + (MODE == MODE_SEQ_S ? 1 : 0)) * kTwistWords;
    layout_hw(K, SPLIT);
    if (!rot_layout_ok(K.ops, SPLIT, K.acc1_global != 0))
        return fail(e, RZK_ERR_INVALID, "OP_ROT needs a warp-per-item program without operand slot and without a second accumulator in shared memory");
    if (K.stash_words > kStashWordsMax) return fail(e, RZK_ERR_INVALID, "program needs more residue stash than the engine provides");
    list_prefetch(K);
    K.cta_sync = K.item_mask ? 0u : e->cta_sync;      // masked launches skip items per warp: no CTA barriers
    K.pp_mode = 0;
    const size_t max_smem = 227 * 1024 - 64;          // 8 bytes of static shared memory hold the TMA mbarrier
    using S = VmSmem<NP, MODE, sp_uses_key<SP>()>;
    int warps = (int)((max_smem - S::bytes(0, 0)) / (sizeof(uint32_t) * 2 * K.hw_words));
    if (warps > VmCfg<NP, MODE>::kMaxWarps) warps = VmCfg<NP, MODE>::kMaxWarps;
    const uint32_t per_warp = SPLIT ? 1 : 2;
    // do not launch more warps per CTA than the batch can use
    const uint32_t want = (uint32_t)((K.n_items + (uint64_t)e->num_sms * per_warp - 1) / ((uint64_t)e->num_sms * per_warp));
    // (but at least 4: the whole CTA stages the tables, which is what a single call on one item waits for)
    if ((uint32_t)warps > want) warps = (int)std::max<uint32_t>(want, (uint32_t)std::min(4, warps));
    // Items cost the same, so a launch takes ceil(items / resident items) waves and its last wave may be nearly empty
    // (2^14 Linear instances on 16 warps of half-warp items: 3.46 waves, billed as 4).  A CTA with a few warps less can need
    // the same number of waves, each of them shorter: pick the warp count in [warps - 4, warps] with the fewest warp-waves.
    if (e->wave_fit && !K.item_mask && warps > 8) {
        const uint64_t sms = (uint64_t)e->num_sms;
        auto cost = [&](int w) { return ((K.n_items + sms * per_warp * w - 1) / (sms * per_warp * w)) * (uint64_t)w; };
        // (a phase-mixed program needs its groups of equal size)
        const uint32_t ppw = e->pp_mode ? e->pp_mode : pp_program;
        const int div = (ppw && ppw != 9 && !std::is_void<SP>::value) ? (ppw >= 10 ? (int)(ppw / 10) : 2) : 1;
        int best = warps;
        for (int w = warps - 1; w >= warps - 4; --w)
            if (w % div == 0 && cost(w) * 100 < cost(best) * 97) best = w;          // (3 % margin: fewer resident warps hide less latency)
        warps = best;
    }
    // phase mixing: RZK_PP for every static program (experiments), else the program's own setting
    const uint32_t pp = e->pp_mode ? e->pp_mode : pp_program;
    const uint32_t pp_groups = pp >= 10 ? pp / 10 : 2u;
    if (pp && pp != 9 && !std::is_void<SP>::value && pp_groups >= 2 && pp_groups <= 8 && (pp < 10 || pp % 10 >= 1) &&
        (uint32_t)warps >= pp_groups && (uint32_t)warps % pp_groups == 0) {
        K.pp_mode = pp;
        // strict alternation already keeps each group in step; staggered groups (pp >= 10) lock-step per segment inside each
        // group (cta_lockstep); the two-group offset modes 1 / 3 per transform
        K.cta_sync = (pp == 2) ? 0u : (pp >= 10 ? e->cta_sync : 2u);
    }
    const size_t smem = S::bytes(warps, K.hw_words);
    static std::atomic<bool> configured[16];   // per device; engines of a group run on separate threads
    if (!configured[e->device & 15]) {
        RZK_CUDA(e, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_smem));
        configured[e->device & 15] = true;
    }
    const uint32_t per_cta = (uint32_t)warps * per_warp;
    uint32_t grid = (K.n_items + per_cta - 1) / per_cta;
    if (grid > (uint32_t)e->num_sms) grid = (uint32_t)e->num_sms;
    kern<<<grid, warps * 32, smem, s>>>(K);
    RZK_CUDA(e, cudaGetLastError());
    e->launches++;
    return RZK_OK;
}

template <class SP>
int launch_sp(rzk_engine *e, VmLaunch &K, cudaStream_t s, uint32_t pp_program = 0)
{
    if (e->no_static) {      // RZK_NO_STATIC=1: run the same program through the generic interpreter
        if (SP::kMode == MODE_SPLITKEY) return launch_vm<1, MODE_SPLITKEY>(e, K, s);
        if (SP::kMode == MODE_SPLITKEY_S) return launch_vm<1, MODE_SPLITKEY_S>(e, K, s);
        if (SP::kMode == MODE_SEQ_S) return launch_vm<3, MODE_SEQ_S>(e, K, s);
        if (SP::kMode == MODE_SPLIT) return launch_vm<2, MODE_SPLIT>(e, K, s);
        return SP::kNP == 1 ? launch_vm<1, MODE_SEQ>(e, K, s) : launch_vm<3, MODE_SEQ>(e, K, s);
    }
    return launch_vm<SP::kNP, SP::kMode, SP>(e, K, s, pp_program);
}

int launch_np(rzk_engine *e, int np, VmLaunch &K, cudaStream_t s)
{
    if (np == 1) return launch_vm<1, MODE_SEQ>(e, K, s);
    if (np == 2) return launch_vm<2, MODE_SPLIT>(e, K, s);
    return launch_vm<3, MODE_SEQ>(e, K, s);
}

int check_ready(rzk_engine *e, bool need_key = true)
{
    if (!e) return RZK_ERR_INVALID;
    if (need_key && !e->has_key) return fail(e, RZK_ERR_NOKEY, "rzk_set_key has not been called");
    return RZK_OK;
}

int ensure_scratch(rzk_engine *e, size_t bytes)
{
    if (bytes <= e->scratch_cap) return RZK_OK;
    RZK_CUDA(e, cudaDeviceSynchronize());
    if (e->scratch) cudaFree(e->scratch);
    e->scratch = nullptr; e->scratch_cap = 0;
    RZK_CUDA(e, cudaMalloc(&e->scratch, bytes));
    e->scratch_cap = bytes;
    return RZK_OK;
}

int ensure_need(rzk_engine *e, size_t words)
{
    if (words <= e->need_cap) return RZK_OK;
    RZK_CUDA(e, cudaDeviceSynchronize());
    if (e->d_need) cudaFree(e->d_need);
    e->d_need = nullptr; e->need_cap = 0;
    RZK_CUDA(e, cudaMalloc(&e->d_need, words * sizeof(uint32_t)));
    e->need_cap = words;
    return RZK_OK;
}

constexpr size_t kPolyBytes = (size_t)kN * sizeof(int32_t);

#define RZK_TRY(x) do { int rc__ = (x); if (rc__ != RZK_OK) return rc__; } while (0)

// ---- phase lowering on device pointers (scratch supplied by the caller of these helpers) ----

constexpr uint32_t kSplitKeyLimit = 15;    // |r| bound of MODE_SPLITKEY: 2*512*2^15*15 < p/2
constexpr uint32_t kSmallCommitLimit = 1;  // |r| bound of MODE_SPLITKEY_S on the transformed rows: 2*512*2^15 + 127 < kStaticPrimeS/2

// c = [a1;a2].r + [0;x]  (commit.rs:88-128).
// Default: the split-key program, exact for |r| <= 15 on the transformed rows.  An item outside that range is marked
//   * in rmark[item / flag_div] (+ the any-word that follows the marks; rmark holds groups + 1 words) when the caller
//     supplies rmark -- the host entry points do, and
//     then redo exactly the marked item groups with the two-prime program in a second, masked launch on the same stream
//     (it returns at once when nothing is marked), so they are exact for any int8 r without a host round trip;
//   * else by FLAG_RANGE in its flags word (`_dev` entry points: documented in the header).
// Engines created with b > 15 (generic_commit) always run the two-prime program, which is exact for any int8 r.
int dev_commit(rzk_engine *e, size_t B, const int32_t *x, const int8_t *r, int32_t *c, uint32_t *flags, cudaStream_t s,
               uint32_t *rmark = nullptr, uint32_t flag_div = 1)
{
    uint32_t *rmark_any = rmark ? rmark + (B + flag_div - 1) / flag_div : nullptr;       // the any-word follows the marks
    // check_commit_constraint (params.rs:102-108) cannot fail for int8 rows when the bound is at least
    // 127*sqrt(N) (it is 1,359,072 at the default parameters): then the norm pass is skipped.
    const bool norm_vacuous = e->cbound >= 127ull * 23ull;
    auto launch = [&](bool generic, bool masked) -> int {
        VmLaunch K; memset(&K, 0, sizeof(K));
        Prog p;
        if (generic) prog_commit(p, 0, 1, 2, !norm_vacuous);
        else prog_commit_splitkey(p, 0, 1, 2, !norm_vacuous);
        p.end();
        p.install(K);
        fill_common(e, K, generic ? 2 : 1, (uint32_t)B, flag_div, flags);
        set_stream(K, 0, x, 1, DT_I32); set_stream(K, 1, r, 3, DT_I8); set_stream(K, 2, c, 2, DT_I32);
        if (generic) {
            if (masked) { K.item_mask = rmark; K.mask_div = flag_div; K.any_item = rmark_any; }
            return launch_np(e, 2, K, s);
        }
        K.rmark = rmark; K.rmark_any = rmark_any;
        if (e->small_commit) {
            // b = 1: one small prime, signed lazy arithmetic (MODE_SPLITKEY_S); rows with |r| > 1 are redone like the others
            K.pc[0] = make_prime_consts(kSignedSlot);
            K.p0d = (double)K.pc[0].p; K.p0invd = 1.0 / (double)K.pc[0].p;
            K.small_lim = kSmallCommitLimit;
            K.keytab = e->d_keytab3;
            if (norm_vacuous) return launch_sp<SPCommitSplitKeyS>(e, K, s, e->commit_pp);
            return launch_vm<1, MODE_SPLITKEY_S>(e, K, s);
        }
        K.small_lim = kSplitKeyLimit;
        K.keytab = e->d_keytab2;
        // measured best for this program: the two halves of the CTA alternate their multiply-heavy windows (rzk_vm_exec.cuh pp_*)
        if (norm_vacuous) return launch_sp<SPCommitSplitKey>(e, K, s, e->commit_pp);
        return launch_vm<1, MODE_SPLITKEY>(e, K, s);
    };
    if (B == 0) return RZK_OK;
    if (e->generic_commit) return launch(true, false);
    if (rmark) {
        const size_t groups = (B + flag_div - 1) / flag_div;
        RZK_CUDA(e, cudaMemsetAsync(rmark, 0, sizeof(uint32_t) * (groups + 1), s));
    }
    RZK_TRY(launch(false, false));
    return rmark ? launch(true, true) : RZK_OK;
}

// t = A1.y (and optionally w = A2.y) for `items` masking vectors: two-prime program
int dev_keymatvec(rzk_engine *e, size_t items, const int32_t *y, int32_t *t, int32_t *w, uint32_t *flags, uint32_t flag_div,
                  cudaStream_t s)
{
    VmLaunch K; memset(&K, 0, sizeof(K));
    Prog p;
    prog_keymatvec(p, 0, 1, w ? 2 : -1, true);
    p.end();
    p.install(K);
    fill_common(e, K, 2, (uint32_t)items, flag_div, flags);
    set_stream(K, 0, y, 3, DT_I32); set_stream(K, 1, t, 1, DT_I32);
    if (w) set_stream(K, 2, w, 1, DT_I32);
    return w ? launch_sp<SPKeyMatVecTW>(e, K, s) : launch_sp<SPKeyMatVecT>(e, K, s);
}

// open.rs:80-103: the commitment (split-key program) and t = A1.y (two-prime program)
int dev_open_commit(rzk_engine *e, size_t B, const int32_t *x, const int8_t *r, const int32_t *y,
                    int32_t *c, int32_t *t, uint32_t *flags, cudaStream_t s, uint32_t *rmark = nullptr)
{
    RZK_TRY(dev_commit(e, B, x, r, c, flags, s, rmark));
    return dev_keymatvec(e, B, y, t, nullptr, flags, 1, s);
}

// z = y + d*r.  `need` ([items + 1] words of device scratch owned by the caller) carries the hand-over between the
// two launches: the rotation kernel (rzk_sparse.cuh) answers every item whose operands fit its byte accumulators
// -- all honest inputs -- and flags the rest; the one-prime NTT program then redoes exactly the flagged items
// (it returns at once when there are none).  need == nullptr: NTT program for everything.
int dev_respond(rzk_engine *e, size_t items, const int32_t *y, const int8_t *r, const int8_t *d, uint32_t d_div,
                int32_t *z, uint32_t *need, cudaStream_t s)
{
    if (items == 0) return RZK_OK;
    const bool sparse = need && !e->no_sparse;
    if (sparse) {
        SparseLaunch SK;
        memset(&SK, 0, sizeof(SK));
        SK.y = y; SK.r = r; SK.d = d; SK.z = z; SK.need = need; SK.any_need = need + items;
        SK.n_items = (uint32_t)items; SK.d_div = d_div; SK.q = (uint32_t)e->P.q;
        RZK_CUDA(e, cudaMemsetAsync(need + items, 0, sizeof(uint32_t), s));
        int warps = 16;
        const uint32_t want = (uint32_t)((items + (uint64_t)e->num_sms - 1) / (uint64_t)e->num_sms);
        if ((uint32_t)warps > want) warps = (int)(want ? want : 1);
        const size_t smem = sizeof(uint32_t) * (size_t)warps * kSpWarpWords;
        static std::atomic<bool> configured[16];   // per device; engines of a group run on separate threads
        if (!configured[e->device & 15]) {
            RZK_CUDA(e, cudaFuncSetAttribute(rzk_respond_sparse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)(sizeof(uint32_t) * 16 * kSpWarpWords)));
            configured[e->device & 15] = true;
        }
        uint32_t grid = (uint32_t)((items + warps - 1) / warps);
        if (grid > (uint32_t)e->num_sms) grid = (uint32_t)e->num_sms;
        rzk_respond_sparse_kernel<<<grid, warps * 32, smem, s>>>(SK);
        RZK_CUDA(e, cudaGetLastError());
        e->launches++;
    }
    VmLaunch K; memset(&K, 0, sizeof(K));
    Prog p;
    prog_respond(p, 0, 1, 2, 3);
    p.end();
    p.install(K);
    fill_common(e, K, 1, (uint32_t)items, 1, nullptr);
    set_stream(K, 0, y, 3, DT_I32); set_stream(K, 1, r, 3, DT_I8); set_stream(K, 2, d, 1, DT_I8, d_div);
    set_stream(K, 3, z, 3, DT_I32);
    if (sparse) { K.item_mask = need; K.any_item = need + items; }
    // measured: the runtime-decoded kernel is faster than the unrolled one for this program (77 vs 61 M/s)
    return (e->static_respond && !sparse) ? launch_sp<SPRespond>(e, K, s) : launch_np(e, 1, K, s);
}

// norm check + first equation (+ optional w = A2.z - c2*d) for `items` responses
// NTT image (two primes) of `groups` challenges d, for dev_verify_first(..., dimg): 2 polys of words per group
int dev_challenge_image(rzk_engine *e, size_t groups, const int8_t *d, uint32_t *dimg, cudaStream_t s)
{
    VmLaunch K; memset(&K, 0, sizeof(K));
    SPChallengeImage::prog.install(K);
    fill_common(e, K, 2, (uint32_t)groups, 1, nullptr);
    set_stream(K, 0, d, 1, DT_I8); set_stream(K, 1, dimg, 2, DT_I32);
    return launch_sp<SPChallengeImage>(e, K, s);
}

// dimg != nullptr (with w): the image of d comes from dev_challenge_image, one per d_div items.
// Without w (Open verify) the product c1*d is a rotation sum in the epilogue (OP_ROT) instead of two transforms per prime.
int dev_verify_first(rzk_engine *e, size_t items, const int32_t *z, const int32_t *t, const int32_t *c, uint32_t c_stride,
                     const int8_t *d, uint32_t d_div, int32_t *w, uint32_t *flags, uint32_t flag_div, cudaStream_t s,
                     const uint32_t *dimg = nullptr)
{
    VmLaunch K; memset(&K, 0, sizeof(K));
    Prog p;
    const bool rot = !e->no_rot && (!w || !e->no_rot_w);   // c1*d (and c2*d) as signed rotations in the epilogues (OP_ROT)
    if (rot) dimg = nullptr;
    prog_norm_verify(p, 0);
    prog_verify_first(p, 0, 1, 2, 3, w ? 4 : -1, (w && dimg) ? 5 : -1, rot);
    p.end();
    p.install(K);
    fill_common(e, K, 2, (uint32_t)items, flag_div, flags);
    set_stream(K, 0, z, 3, DT_I32); set_stream(K, 1, t, 1, DT_I32); set_stream(K, 2, c, c_stride, DT_I32);
    set_stream(K, 3, d, 1, DT_I8, d_div);
    if (w) set_stream(K, 4, w, 1, DT_I32);
    if (w && dimg) { set_stream(K, 5, dimg, 2, DT_I32, d_div); return launch_sp<SPVerifyFirstWG>(e, K, s); }
    // (tools/ab_time.py, profiles/r2_ab_timings.log: the staggered two-group start decides for the first rotation kernel --
    // 86 -> 98.7 M/s -- and is neutral for the final one)
    if (rot && w) return launch_sp<SPVerifyFirstWRot>(e, K, s, e->verify_w_pp);
    if (rot) return launch_sp<SPVerifyFirstRot>(e, K, s, e->verify_pp);
    return w ? launch_sp<SPVerifyFirstW>(e, K, s) : launch_sp<SPVerifyFirst>(e, K, s);
}

// Commitment::verify (commit.rs:173-210); f == nullptr is the `None` branch
int dev_commitment_verify(rzk_engine *e, size_t B, const int32_t *c, const int32_t *x, const int8_t *r, const int8_t *f,
                          uint32_t *flags, cudaStream_t s)
{
    VmLaunch K; memset(&K, 0, sizeof(K));
    Prog p;
    prog_commitment_verify(p, 0, 1, 2, f ? 3 : -1);
    p.end();
    p.install(K);
    fill_common(e, K, 2, (uint32_t)B, 1, flags);
    set_stream(K, 0, c, 2, DT_I32); set_stream(K, 1, x, 1, DT_I32); set_stream(K, 2, r, 3, DT_I8);
    if (f) set_stream(K, 3, f, 1, DT_I8);
    return launch_np(e, 2, K, s);
}

// A product sum is one item per instance: a half warp walks its T terms.  A batch too small to fill the SMs (single
// calls, the tail shard of a multi-GPU job) is cut into `segs` segments per instance, segs | T, that run as separate
// items and are summed by rzk_partial_reduce_kernel: [B][T] is [B * segs][T / segs] in memory, so no data moves.
uint32_t mulsum_segments(const rzk_engine *e, size_t B, uint32_t T)
{
    const size_t slots = (size_t)e->num_sms * RZK_SEQ3_WARPS * 2;        // half-warp items of one resident wave
    if (e->no_segments || T < 2 || B * 2 > slots) return 1;
    uint32_t best = 1;
    for (uint32_t sg = 2; sg <= T; ++sg)
        if (T % sg == 0 && B * sg <= slots) best = sg;
    return best;
}

int dev_mulsum(rzk_engine *e, size_t B, uint32_t T, const int32_t *a, const int32_t *b, const int32_t *sub0,
               const int32_t *sub1, int32_t *out, uint32_t *flags, cudaStream_t s);

int dev_mulsum_segmented(rzk_engine *e, size_t B, uint32_t T, uint32_t segs, const int32_t *a, const int32_t *b, const int32_t *sub0,
                         const int32_t *sub1, int32_t *out, uint32_t *flags, cudaStream_t s)
{
    int si = kPipe;
    for (int i = 0; i < kPipe; ++i) if (e->pipe[i].stream == s) si = i;
    if (!e->d_partial[si])
        RZK_CUDA(e, cudaMalloc(&e->d_partial[si], kPolyBytes * (size_t)e->num_sms * RZK_SEQ3_WARPS * 2));
    int32_t *part = e->d_partial[si];
    e->no_segments |= 2u;                                                     // the segment launch itself is not cut again
    const int rc = dev_mulsum(e, B * segs, T / segs, a, b, nullptr, nullptr, part, nullptr, s);
    e->no_segments &= ~2u;
    RZK_TRY(rc);
    const size_t n = B * kN;
    rzk_partial_reduce_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(n, segs, part, sub0, sub1, out, flags ? flags : e->d_misc + 1,
                                                                           flags ? 0u : 1u, (int64_t)e->P.q);
    RZK_CUDA(e, cudaGetLastError());
    e->launches++;
    return RZK_OK;
}

// out = sum_{i<T} a_i*b_i - sub0 - sub1 (store) or == 0 (compare)
int dev_mulsum(rzk_engine *e, size_t B, uint32_t T, const int32_t *a, const int32_t *b, const int32_t *sub0,
               const int32_t *sub1, int32_t *out, uint32_t *flags, cudaStream_t s)
{
    if (const uint32_t segs = mulsum_segments(e, B, T); segs > 1)
        return dev_mulsum_segmented(e, B, T, segs, a, b, sub0, sub1, out, flags, s);
    VmLaunch K; memset(&K, 0, sizeof(K));
    Prog p;
    prog_mulsum(p, (int)T, 0, 1, sub0 ? 2 : -1, sub1 ? 3 : -1, out ? 4 : -1, out ? FIN_STORE : FIN_CMPZ);
    p.end();
    p.install(K);
    // up to 64 terms fit the three small primes (any int32 operands: 64 * 512 * 2^62 < p3 p4 p5 / 2): signed lazy arithmetic
    const bool small = e->mulsum_small && T <= (uint32_t)kSignedMaxTerms;
    fill_common(e, K, 3, (uint32_t)B, 1, flags, small);
    set_stream(K, 0, a, T, DT_I32); set_stream(K, 1, b, T, DT_I32);
    if (sub0) set_stream(K, 2, sub0, 1, DT_I32);
    if (sub1) set_stream(K, 3, sub1, 1, DT_I32);
    if (out) set_stream(K, 4, out, 1, DT_I32);
    K.loop_count = T - 1;
    if (small) {
        if (out && !sub0 && !sub1) return launch_sp<SPMulSum0S>(e, K, s);
        if (out && sub0 && !sub1) return launch_sp<SPMulSum1S>(e, K, s);
        if (!out && sub0 && sub1) return launch_sp<SPMulSumCmpS>(e, K, s);
        return launch_vm<3, MODE_SEQ_S>(e, K, s);
    }
    if (out && !sub0 && !sub1) return launch_sp<SPMulSum0>(e, K, s);
    if (out && sub0 && !sub1) return launch_sp<SPMulSum1>(e, K, s);
    if (!out && sub0 && sub1) return launch_sp<SPMulSumCmp>(e, K, s);
    return launch_np(e, 3, K, s);
}

// out0 = sum_{i<T} a_i*b_i,  out1 = sum_{i<T} a_i*c_i - sub   (every a_i transformed once; prog_mulsum2)
int dev_mulsum2(rzk_engine *e, size_t B, uint32_t T, const int32_t *a, const int32_t *b, const int32_t *c, const int32_t *sub,
                int32_t *out0, int32_t *out1, uint32_t *flags, cudaStream_t s)
{
    VmLaunch K; memset(&K, 0, sizeof(K));
    Prog p;
    prog_mulsum2(p, (int)T, 0, 1, 2, 3, 4, 5);
    p.end();
    p.install(K);
    const bool small = e->mulsum_small && T <= (uint32_t)kSignedMaxTerms;
    fill_common(e, K, 3, (uint32_t)B, 1, flags, small);
    set_stream(K, 0, a, T, DT_I32); set_stream(K, 1, b, T, DT_I32); set_stream(K, 2, c, T, DT_I32);
    set_stream(K, 3, sub, 1, DT_I32); set_stream(K, 4, out0, 1, DT_I32); set_stream(K, 5, out1, 1, DT_I32);
    K.loop_count = T - 1;
    if (small) return launch_sp<SPMulSum2S>(e, K, s, e->mulsum2_pp);
    // phase mixing between the CTA halves (as for the commitment program) measured +3.4 % on this kernel at 2^12 x 64 terms;
    // on the single-accumulator product sums, A.y and the verify programs it measured within +-1 % or slower and stays off
    return launch_sp<SPMulSum2>(e, K, s, e->mulsum2_pp);
}

// commit(x; r) -> c  and  t = A1.y, w = A2.y  for `items` (x, r, y) triples
int dev_commit_matvec(rzk_engine *e, size_t items, const int32_t *x, const int8_t *r, const int32_t *y,
                      int32_t *c, int32_t *t, int32_t *w, uint32_t *flags, uint32_t flag_div, cudaStream_t s, uint32_t *rmark = nullptr)
{
    RZK_TRY(dev_commit(e, items, x, r, c, flags, s, rmark, flag_div));
    return dev_keymatvec(e, items, y, t, w, flags, flag_div, s);
}

// scratch: 2*B polys
int dev_linear_commit(rzk_engine *e, size_t B, const int32_t *g, const int32_t *x, const int8_t *rp, const int8_t *r,
                      const int32_t *y, const int32_t *yp, int32_t *gx, int32_t *cp, int32_t *c, int32_t *t,
                      int32_t *tp, int32_t *u, uint32_t *flags, int32_t *scratch, cudaStream_t s, uint32_t *rmark = nullptr)
{
    int32_t *w = scratch, *wp = scratch + B * kN;
    // One launch per product here.  Sharing the transform of g between g*x and u (prog_mulsum2, or a single-term program
    // that inverts one product after the other) was measured 5 % SLOWER for single terms.
    RZK_TRY(dev_mulsum(e, B, 1, g, x, nullptr, nullptr, gx, flags, s));                 // linear.rs:91-95
    RZK_TRY(dev_commit_matvec(e, B, gx, rp, yp, cp, tp, wp, flags, 1, s, rmark));     // linear.rs:96,121,129
    RZK_TRY(dev_commit_matvec(e, B, x, r, y, c, t, w, flags, 1, s, rmark));           // linear.rs:97,118,124-127
    return dev_mulsum(e, B, 1, g, w, wp, nullptr, u, flags, s);                         // linear.rs:124-129
}

int dev_linear_verify(rzk_engine *e, size_t B, const int32_t *z, const int32_t *zp, const int32_t *c, const int32_t *cp,
                      const int32_t *g, const int32_t *t, const int32_t *tp, const int32_t *u, const int8_t *d,
                      uint32_t *flags, int32_t *scratch, cudaStream_t s)
{
    int32_t *w = scratch, *wp = scratch + B * kN;
    // (sharing the transform of d between the two equations was measured slower both ways: as an image from a launch of its
    // own, as in dev_sum_verify, by 1 %; as one program for both equations by 49 % -- twice the straight-line code)
    RZK_TRY(dev_verify_first(e, B, z, t, c, 2, d, 1, w, flags, 1, s));                  // linear.rs:218,225-229
    RZK_TRY(dev_verify_first(e, B, zp, tp, cp, 2, d, 1, wp, flags, 1, s));              // linear.rs:221,231-235
    return dev_mulsum(e, B, 1, g, w, wp, u, nullptr, flags, s);                         // linear.rs:236-249
}

// scratch: (B*T + B) polys
int dev_sum_commit(rzk_engine *e, size_t B, uint32_t T, const int32_t *gs, const int32_t *xs, const int8_t *rp,
                   const int8_t *rs, const int32_t *ys, const int32_t *yp, int32_t *xp, int32_t *cp, int32_t *cs,
                   int32_t *ts, int32_t *tp, int32_t *u, uint32_t *flags, int32_t *scratch, cudaStream_t s, uint32_t *rmark = nullptr)
{
    int32_t *ws = scratch, *wp = scratch + B * T * kN;
    if (e->no_fuse == 1 || T == 1 || mulsum_segments(e, B, T) > 1) {      // (small batches: each product sum is cut into segments)
        RZK_TRY(dev_mulsum(e, B, T, gs, xs, nullptr, nullptr, xp, flags, s));               // sum.rs:107-115
        RZK_TRY(dev_commit_matvec(e, B, xp, rp, yp, cp, tp, wp, flags, 1, s, rmark));     // sum.rs:116,151,160
        RZK_TRY(dev_commit_matvec(e, B * T, xs, rs, ys, cs, ts, ws, flags, T, s, rmark)); // sum.rs:117-120,145-148,157
        return dev_mulsum(e, B, T, gs, ws, wp, nullptr, u, flags, s);                       // sum.rs:154-160
    }
    // same results in an order that lets every g_i be transformed once for both of its products: the masking products
    // first (they do not depend on x'), then x' = sum g_i x_i and u in one launch, then the commitment to x'
    RZK_TRY(dev_keymatvec(e, B, yp, tp, wp, flags, 1, s));                              // sum.rs:151,160
    RZK_TRY(dev_commit_matvec(e, B * T, xs, rs, ys, cs, ts, ws, flags, T, s, rmark)); // sum.rs:117-120,145-148,157
    RZK_TRY(dev_mulsum2(e, B, T, gs, xs, ws, wp, xp, u, flags, s));                     // sum.rs:107-115,154-160
    return dev_commit(e, B, xp, rp, cp, flags, s, rmark);                               // sum.rs:116
}

int dev_sum_verify(rzk_engine *e, size_t B, uint32_t T, const int32_t *zs, const int32_t *zp, const int32_t *cs,
                   const int32_t *cp, const int32_t *gs, const int32_t *ts, const int32_t *tp, const int32_t *u,
                   const int8_t *d, uint32_t *flags, int32_t *scratch, cudaStream_t s)
{
    int32_t *ws = scratch, *wp = scratch + B * T * kN;
    // (only the NTT-domain lowering of c*d needs the challenge's image; the default adds c1*d, c2*d as rotation sums)
    const bool ntt_cd = e->no_rot || e->no_rot_w;
    uint32_t *dimg = (e->no_dimg || !ntt_cd) ? nullptr : reinterpret_cast<uint32_t *>(scratch + (B * T + B) * kN);   // T + 1 equations share d
    if (dimg) RZK_TRY(dev_challenge_image(e, B, d, dimg, s));
    RZK_TRY(dev_verify_first(e, B * T, zs, ts, cs, 2, d, T, ws, flags, T, s, dimg));    // sum.rs:262-268,277-291
    RZK_TRY(dev_verify_first(e, B, zp, tp, cp, 2, d, 1, wp, flags, 1, s, dimg));        // sum.rs:269,293-298
    return dev_mulsum(e, B, T, gs, ws, wp, u, nullptr, flags, s);                       // sum.rs:300-319
}

// ---- chunked host pipeline -----------------------------------------------------------------

struct HArr {
    const void *in;      // host input  (or nullptr)
    void *out;           // host output (or nullptr)
    size_t per_item;     // bytes per item
};

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

constexpr size_t kArenaCap = (size_t)1 << 30;     // one pipeline slot never holds more than 1 GiB

// fn(chunk_items, dptr[], scratch, flags, stream, rmark): rmark = (chunk + 1) words for dev_commit's masked redo
template <class F>
int run_chunked(rzk_engine *e, size_t B, std::vector<HArr> &arrs, size_t scratch_per_item, uint8_t *bitmap, F &&fn)
{
    if (B == 0) return RZK_OK;
    Guard g(e->device);
    e->err.clear();
    size_t per_item = scratch_per_item + 2 * sizeof(uint32_t) + 1;
    for (auto &a : arrs) per_item += a.per_item;
    // (a Sum proof with tens of thousands of terms is hundreds of MB per instance: refuse instead of allocating four
    // multi-GB arenas -- such an instance belongs to the `_dev` entry points with caller-owned buffers)
    if (8 * per_item > kArenaCap)
        return fail(e, RZK_ERR_INVALID, "one item group needs more than the host pipeline's 1 GiB arena: use the _dev entry points");
    size_t chunk = (size_t)(96ull << 20) / per_item;
    chunk = std::max<size_t>(8, std::min<size_t>(chunk, e->chunk_items)) / 8 * 8;
    if (chunk > B) chunk = align_up(B, 8);
    // arena layout for one pipeline slot
    std::vector<size_t> offs(arrs.size());
    size_t off = 0;
    for (size_t i = 0; i < arrs.size(); ++i) { offs[i] = off; off += align_up(arrs[i].per_item * chunk, 256); }
    const size_t off_scratch = off; off += align_up(scratch_per_item * chunk, 256);
    const size_t off_flags = off; off += align_up(sizeof(uint32_t) * chunk, 256);
    const size_t off_rmark = off; off += align_up(sizeof(uint32_t) * (chunk + 1), 256);
    const size_t off_bitmap = off; off += align_up(chunk / 8 + 1, 256);
    const size_t need = off;
    for (int i = 0; i < kPipe; ++i) {
        PipeSlot &ps = e->pipe[i];
        if (ps.cap < need) {
            RZK_CUDA(e, cudaStreamSynchronize(ps.stream));
            if (ps.arena) cudaFree(ps.arena);
            ps.arena = nullptr; ps.cap = 0;
            RZK_CUDA(e, cudaMalloc(&ps.arena, need));
            ps.cap = need;
        }
    }
    // the range word is cleared on the first stream; the other streams are only used (and must wait) when the batch
    // spans several chunks -- a single call on one item stays on one stream with one synchronisation at the end
    // The first results can only start down the bus once the first chunk is up and computed, and the download is the longer
    // direction of a commitment batch: a batch of several chunks therefore starts with small ones (chunk / 8, / 8, / 4, / 2)
    // so that the download engine is busy almost from the start (measured: +4 % end to end at 2^16 commitments)
    auto chunk_len = [&](int ci) -> size_t {
        if (!e->chunk_ramp || B < 2 * chunk || chunk < 512) return chunk;
        const size_t div = ci < 2 ? 8 : ci == 2 ? 4 : ci == 3 ? 2 : 1;
        return chunk / div / 8 * 8;
    };
    size_t nchunks = 0;
    for (size_t c0 = 0; c0 < B; c0 += chunk_len((int)nchunks), ++nchunks) {}
    RZK_CUDA(e, cudaMemsetAsync(e->d_misc, 0, sizeof(uint32_t), e->pipe[0].stream));
    if (nchunks > 1) RZK_CUDA(e, cudaStreamSynchronize(e->pipe[0].stream));
    std::vector<void *> dptr(arrs.size());
    // a failure in any chunk must not return while other streams still copy into the caller's buffers or run kernels on
    // the arenas: every stream is drained first
    auto body = [&]() -> int {
        int ci = 0;
        for (size_t c0 = 0, n = 0; c0 < B; c0 += n, ++ci) {
            n = std::min(chunk_len(ci), B - c0);
            PipeSlot &ps = e->pipe[ci % kPipe];
            cudaStream_t s = ps.stream;
            for (size_t i = 0; i < arrs.size(); ++i) {
                dptr[i] = ps.arena + offs[i];
                if (arrs[i].in)
                    RZK_CUDA(e, cudaMemcpyAsync(dptr[i], (const char *)arrs[i].in + c0 * arrs[i].per_item,
                                                n * arrs[i].per_item, cudaMemcpyHostToDevice, s));
            }
            uint32_t *dflags = reinterpret_cast<uint32_t *>(ps.arena + off_flags);
            RZK_CUDA(e, cudaMemsetAsync(dflags, 0, sizeof(uint32_t) * n, s));
            RZK_TRY(fn(n, dptr.data(), ps.arena + off_scratch, dflags, s, reinterpret_cast<uint32_t *>(ps.arena + off_rmark)));
            if (bitmap) {
                uint8_t *dbm = reinterpret_cast<uint8_t *>(ps.arena + off_bitmap);
                const size_t nbytes = (n + 7) / 8;
                rzk_flags_to_bitmap_kernel<<<(unsigned)((nbytes + 127) / 128), 128, 0, s>>>(n, dflags, dbm, e->d_misc);
                RZK_CUDA(e, cudaGetLastError());
                e->launches++;
                RZK_CUDA(e, cudaMemcpyAsync(bitmap + c0 / 8, dbm, nbytes, cudaMemcpyDeviceToHost, s));
            }
            for (size_t i = 0; i < arrs.size(); ++i)
                if (arrs[i].out)
                    RZK_CUDA(e, cudaMemcpyAsync((char *)arrs[i].out + c0 * arrs[i].per_item, dptr[i],
                                                n * arrs[i].per_item, cudaMemcpyDeviceToHost, s));
        }
        return RZK_OK;
    };
    const int rc = body();
    if (rc != RZK_OK) {
        for (int i = 0; i < kPipe; ++i) cudaStreamSynchronize(e->pipe[i].stream);
        return rc;
    }
    uint32_t range = 0;
    if (nchunks == 1) {
        RZK_CUDA(e, cudaMemcpyAsync(e->h_range, e->d_misc, sizeof(uint32_t), cudaMemcpyDeviceToHost, e->pipe[0].stream));
        RZK_CUDA(e, cudaStreamSynchronize(e->pipe[0].stream));
        range = *e->h_range;
    } else {
        cudaError_t first = cudaSuccess;
        for (int i = 0; i < kPipe; ++i) {
            const cudaError_t ce = cudaStreamSynchronize(e->pipe[i].stream);
            if (ce != cudaSuccess && first == cudaSuccess) first = ce;
        }
        RZK_CUDA(e, first);
        RZK_CUDA(e, cudaMemcpy(&range, e->d_misc, sizeof(uint32_t), cudaMemcpyDeviceToHost));
    }
    if (range & FLAG_RANGE)
        return fail(e, RZK_ERR_RANGE, "a masking vector y exceeds rzk_small_limit(); affected outputs are not exact");
    return RZK_OK;
}

bool any_null(std::initializer_list<const void *> ps)
{
    for (auto p : ps) if (!p) return true;
    return false;
}

}  // namespace

// ------------------------------------------------------------------------------ C ABI

extern "C" {

rzk_params rzk_default_params(int32_t N)
{
    rzk_params P;
    P.q = 3515337053LL; P.b = 1; P.N = N; P.n = 1; P.k = 3; P.l = 1; P.kappa = 36;   // params.rs:121-138
    return P;
}

const char *rzk_last_error(const rzk_engine *e) { return e ? e->err.c_str() : g_create_err.c_str(); }
int rzk_device(const rzk_engine *e) { return e ? e->device : -1; }
uint64_t rzk_sigma(const rzk_engine *e) { return e->sigma; }
uint64_t rzk_commit_bound(const rzk_engine *e) { return e->cbound; }
uint64_t rzk_verify_bound(const rzk_engine *e) { return e->vbound; }
uint32_t rzk_small_limit(const rzk_engine *e) { return e->small_lim; }
uint64_t rzk_kernel_launches(const rzk_engine *e) { return e->launches; }

int rzk_create(const rzk_params *params, int device, rzk_engine **out)
{
    if (!params || !out) return fail(nullptr, RZK_ERR_INVALID, "null argument");
    *out = nullptr;
    const rzk_params &P = *params;
    if (kPrimeList[0] != kStaticPrime0) return fail(nullptr, RZK_ERR_INVALID, "prime slot 0 differs from kStaticPrime0");
    if (kPrimeList[kSignedSlot] != kStaticPrimeS) return fail(nullptr, RZK_ERR_INVALID, "the signed prime slot differs from kStaticPrimeS");
    if (P.N != kN || P.n != 1 || P.k != 3 || P.l != 1)
        return fail(nullptr, RZK_ERR_UNSUPPORTED, "only N=512, (n,k,l)=(1,3,1) is accelerated");
    if (P.q != 3515337053LL || P.b < 1 || P.b > 127 || P.kappa < 1)
        return fail(nullptr, RZK_ERR_UNSUPPORTED, "only q=3515337053 with 1 <= b <= 127 is accelerated");
    // params.rs:94-98, 104, 114
    const uint64_t sigma = (uint64_t)P.b * (uint64_t)(11 * (int64_t)P.kappa) * isqrt64((uint64_t)P.k * (uint64_t)P.N);
    const uint64_t cbound = 4 * sigma * isqrt64((uint64_t)P.N), vbound = 2 * sigma * isqrt64((uint64_t)P.N);
    uint32_t small_lim = 0;
    {
        // Exactness of the two-prime products (small x large), from the actual bounds of this parameter set:
        static const int slots[2] = {0, 1};
        const CrtC c = make_crt_consts(slots, 2, (uint64_t)P.q);
        const uint64_t half = (uint64_t)(P.q - 1) / 2, room = c.P01half - (1ull << 33);
        // (a) prover: |y| <= small_lim keeps y0 + a11*y1 + a12*y2 inside the centred CRT range; the masking vectors are
        //     N(0, sigma) samples, so the limit must sit far in the tail (10 sigma: < 2^-75 per coefficient) or honest
        //     batches would be rejected with RZK_ERR_RANGE
        small_lim = (uint32_t)std::min<uint64_t>(room / ((uint64_t)(P.k - P.n) * (uint64_t)P.N * half), 0x7fffffffu);
        if ((uint64_t)small_lim < 10 * sigma)
            return fail(nullptr, RZK_ERR_UNSUPPORTED, "b*kappa too large: masking vectors N(0, sigma) would leave the exact range of "
                                                        "the two-prime products (needs rzk_small_limit() >= 10 sigma, i.e. b*kappa <= 74)");
        // (b) verifier: a response that passes check_verify_constraint has ||z_i||_1 <= sqrt(N) * vbound, so
        //     |a1j * z_j| <= half * sqrt(N) * vbound per coefficient; c*d with any int8 d and any int32 c adds < 2^47
        const uint64_t sqrtN_up = isqrt64((uint64_t)P.N) + 1;
        const unsigned __int128 worst = (unsigned __int128)(P.k - P.n) * half * sqrtN_up * vbound + ((unsigned __int128)1 << 47) + (1ull << 34);
        if (worst > (unsigned __int128)room)
            return fail(nullptr, RZK_ERR_UNSUPPORTED, "b*kappa too large: A1.z of a response at the norm bound would leave the exact "
                                                        "range of the two-prime products");
    }
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0)
        return fail(nullptr, RZK_ERR_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(ce));
    if (device < 0) RZK_CUDA(nullptr, cudaGetDevice(&device));
    if (device >= ndev) return fail(nullptr, RZK_ERR_INVALID, "device index out of range");
    rzk_engine *e = new rzk_engine();
    e->P = P;
    e->device = device;
    e->sigma = sigma; e->cbound = cbound; e->vbound = vbound; e->small_lim = small_lim;
    e->generic_commit = (uint64_t)P.b > kSplitKeyLimit;
    if (const char *cs = getenv("RZK_CHUNK_ITEMS")) e->chunk_items = (uint32_t)std::max(8, atoi(cs));
    if (const char *cs = getenv("RZK_CHUNK_RAMP")) e->chunk_ramp = (uint32_t)atoi(cs);
    auto has_token = [](const char *list, const char *tok) {
        const size_t n = strlen(tok);
        for (const char *p = list; p && *p;) {
            const char *q = strchr(p, ',');
            const size_t len = q ? (size_t)(q - p) : strlen(p);
            if (len == n && strncmp(p, tok, n) == 0) return true;
            p = q ? q + 1 : nullptr;
        }
        return false;
    };
    if (const char *tl = getenv("RZK_TEST_LOWERING")) {
        e->no_static = has_token(tl, "generic"); e->no_sparse = has_token(tl, "nosparse");
        e->no_segments = has_token(tl, "nosegments"); e->no_dimg = has_token(tl, "nodimg");
        e->no_fuse = has_token(tl, "nofuse"); e->no_rot = has_token(tl, "norot"); e->no_rot_w = has_token(tl, "norotw");
    }
    if (const char *tu = getenv("RZK_TUNE")) {
        auto val = [&](const char *name, uint32_t &dst) {
            const std::string key = std::string(name) + "=";
            const char *p = strstr(tu, key.c_str());
            if (p && (p == tu || p[-1] == ',')) dst = (uint32_t)atoi(p + key.size());
        };
        val("cta_sync", e->cta_sync); val("pp", e->pp_mode); val("commit_pp", e->commit_pp); val("commit_small", e->commit_small); val("wave_fit", e->wave_fit); val("mulsum_small", e->mulsum_small);
        val("mulsum2_pp", e->mulsum2_pp); val("static_respond", e->static_respond); val("verify_pp", e->verify_pp); val("verify_w_pp", e->verify_w_pp); val("ld128", e->ld128);
    }
    e->small_commit = P.b == 1 && e->commit_small != 0;
    Guard g(device);
    cudaDeviceProp prop;
    ce = cudaGetDeviceProperties(&prop, device);
    if (ce != cudaSuccess) { delete e; return fail(nullptr, RZK_ERR_CUDA, cudaGetErrorString(ce)); }
    if (prop.major < 10) { delete e; return fail(nullptr, RZK_ERR_CUDA, "an sm_100a (Blackwell) device is required"); }
    e->num_sms = prop.multiProcessorCount;
    // static tables
    std::vector<uint32_t> g1((size_t)kNumPrimeSlots * 2 * kG1Words, 0), g2((size_t)kNumPrimeSlots * 2 * kLanes * kG2Words);
    for (int s = 0; s < kNumPrimeSlots; ++s) {
        const PrimeTables &T = prime_tables(s);
        for (int d = 0; d < 2; ++d) memcpy(&g1[((size_t)s * 2 + d) * kG1Words], T.g1[d], sizeof(T.g1[d]));
        memcpy(&g2[(size_t)s * 2 * kLanes * kG2Words], T.g2, sizeof(T.g2));
    }
    int rc = RZK_OK;
    auto cu = [&](cudaError_t err, const char *what) {
        if (err != cudaSuccess && rc == RZK_OK) rc = fail(nullptr, RZK_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(err));
    };
    cu(cudaMalloc(&e->d_g1tab, g1.size() * sizeof(uint32_t)), "cudaMalloc(g1)");
    if (rc == RZK_OK) cu(cudaMemcpy(e->d_g1tab, g1.data(), g1.size() * sizeof(uint32_t), cudaMemcpyHostToDevice), "cudaMemcpy(g1)");
    cu(cudaMalloc(&e->d_g2tab, g2.size() * sizeof(uint32_t)), "cudaMalloc(g2)");
    if (rc == RZK_OK) cu(cudaMemcpy(e->d_g2tab, g2.data(), g2.size() * sizeof(uint32_t), cudaMemcpyHostToDevice), "cudaMemcpy(g2)");
    {
        std::vector<uint32_t> tw((size_t)kNumPrimeSlots * 2 * kTwistWords, 0);      // [slot][plain, times R N^-1][kTwistWords]
        for (int s = 0; s < kNumPrimeSlots; ++s)
            if (slot_is_signed(s)) {
                memcpy(&tw[((size_t)s * 2 + 0) * kTwistWords], prime_tables(s).twist, sizeof(uint32_t) * kTwistWords);
                memcpy(&tw[((size_t)s * 2 + 1) * kTwistWords], prime_tables(s).twist_rn, sizeof(uint32_t) * kTwistWords);
            }
        cu(cudaMalloc(&e->d_twist, tw.size() * sizeof(uint32_t)), "cudaMalloc(twist)");
        if (rc == RZK_OK) cu(cudaMemcpy(e->d_twist, tw.data(), tw.size() * sizeof(uint32_t), cudaMemcpyHostToDevice), "cudaMemcpy(twist)");
    }
    cu(cudaMalloc(&e->d_keytab, (size_t)kNumPrimeSlots * kKeyPolys * 2 * kPadWords * sizeof(uint32_t)), "cudaMalloc(key)");
    cu(cudaMalloc(&e->d_keytab2, (size_t)2 * kKeyPolys * 2 * kPadWords * sizeof(uint32_t)), "cudaMalloc(key2)");
    cu(cudaMalloc(&e->d_keytab3, (size_t)2 * kKeyPolys * 2 * kPadWords * sizeof(uint32_t)), "cudaMalloc(key3)");
    cu(cudaMalloc(&e->d_misc, 64), "cudaMalloc(misc)");
    cu(cudaMallocHost(&e->h_range, 64), "cudaMallocHost(range)");
    if (rc == RZK_OK) cu(cudaMemset(e->d_misc, 0, 64), "cudaMemset(misc)");
    for (int i = 0; i < kPipe && rc == RZK_OK; ++i) cu(cudaStreamCreateWithFlags(&e->pipe[i].stream, cudaStreamNonBlocking), "cudaStreamCreate");
    if (rc != RZK_OK) { rzk_destroy(e); return rc; }
    *out = e;
    return RZK_OK;
}

void rzk_destroy(rzk_engine *e)
{
    if (!e) return;
    Guard g(e->device);
    cudaDeviceSynchronize();
    for (int i = 0; i < kPipe; ++i) {
        if (e->pipe[i].arena) cudaFree(e->pipe[i].arena);
        if (e->pipe[i].stream) cudaStreamDestroy(e->pipe[i].stream);
    }
    if (e->scratch) cudaFree(e->scratch);
    if (e->d_g1tab) cudaFree(e->d_g1tab);
    if (e->d_g2tab) cudaFree(e->d_g2tab);
    if (e->d_keytab) cudaFree(e->d_keytab);
    if (e->d_keytab2) cudaFree(e->d_keytab2);
    if (e->d_keytab3) cudaFree(e->d_keytab3);
    if (e->d_twist) cudaFree(e->d_twist);
    if (e->d_need) cudaFree(e->d_need);
    if (e->d_fs_prefix) cudaFree(e->d_fs_prefix);
    if (e->d_wire_toks) cudaFree(e->d_wire_toks);
    for (auto p : e->d_gstash) if (p) cudaFree(p);
    for (auto p : e->d_partial) if (p) cudaFree(p);
    if (e->d_misc) cudaFree(e->d_misc);
    if (e->h_range) cudaFreeHost(e->h_range);
    delete e;
}

int rzk_set_key(rzk_engine *e, const int64_t *a1, const int64_t *a2)
{
    RZK_TRY(check_ready(e, false));
    if (!a1 || !a2) return fail(e, RZK_ERR_INVALID, "null key");
    Guard g(e->device);
    const int k = e->P.k;
    // commit.rs:38-57: a1 = [1 | a11 a12], a2 = [0 | 1 | a22]
    auto is_const = [&](const int64_t *p, int64_t c0) {
        if (p[0] != c0) return false;
        for (int i = 1; i < kN; ++i) if (p[i] != 0) return false;
        return true;
    };
    if (!is_const(a1, 1) || !is_const(a2, 0) || !is_const(a2 + kN, 1))
        return fail(e, RZK_ERR_UNSUPPORTED, "key does not have the [I | a1'], [0 | I | a2'] structure of commit.rs:33-60");
    (void)k;
    const int64_t *polys[kKeyPolys] = {a1 + kN, a1 + 2 * kN, a2 + 2 * kN};
    std::vector<uint32_t> img((size_t)kNumPrimeSlots * kKeyPolys * 2 * kPadWords, 0);
    const int64_t q = e->P.q, half = (q - 1) / 2;
    std::vector<int64_t> cen(kN);
    for (int s = 0; s < 3; ++s) {               // the 30-bit prime slots used by the launches
        const PrimeTables &T = prime_tables(s);
        for (int kk = 0; kk < kKeyPolys; ++kk) {
            for (int i = 0; i < kN; ++i) {
                int64_t r = polys[kk][i] % q;
                if (r > half) r -= q; else if (r < -half) r += q;
                cen[i] = r;
            }
            key_image(T, cen.data(), &img[((size_t)s * kKeyPolys + kk) * 2 * kPadWords]);
        }
    }
    std::vector<uint32_t> img2((size_t)2 * kKeyPolys * 2 * kPadWords, 0), img3(img2.size(), 0);
    for (int kk = 0; kk < kKeyPolys; ++kk) {
        for (int i = 0; i < kN; ++i) {
            int64_t r = polys[kk][i] % q;
            if (r > half) r -= q; else if (r < -half) r += q;
            cen[i] = r;
        }
        key_image_split(prime_tables(0), cen.data(), &img2[(size_t)kk * 4 * kPadWords]);
        key_image_split_signed(prime_tables(kSignedSlot), cen.data(), &img3[(size_t)kk * 4 * kPadWords]);
    }
    RZK_CUDA(e, cudaDeviceSynchronize());
    RZK_CUDA(e, cudaMemcpy(e->d_keytab, img.data(), img.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    RZK_CUDA(e, cudaMemcpy(e->d_keytab2, img2.data(), img2.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    RZK_CUDA(e, cudaMemcpy(e->d_keytab3, img3.data(), img3.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    e->has_key = true;
    return RZK_OK;
}

void *rzk_host_alloc(size_t bytes)
{
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes) != cudaSuccess) return nullptr;
    return p;
}

void rzk_host_free(void *p) { if (p) cudaFreeHost(p); }

int rzk_sync(rzk_engine *e, void *stream)
{
    RZK_TRY(check_ready(e, false));
    Guard g(e->device);
    RZK_CUDA(e, cudaStreamSynchronize((cudaStream_t)stream));
    return RZK_OK;
}

// ---- device-resident entry points ----

int rzk_commit_batch_dev(rzk_engine *e, size_t B, const int32_t *x, const int8_t *r, int32_t *c, uint32_t *flags, void *stream)
{
    RZK_TRY(check_ready(e));
    if (any_null({x, r, c, flags})) return fail(e, RZK_ERR_INVALID, "null argument");
    Guard g(e->device);
    return dev_commit(e, B, x, r, c, flags, (cudaStream_t)stream);
}

int rzk_commitment_verify_batch_dev(rzk_engine *e, size_t B, const int32_t *c, const int32_t *x, const int8_t *r, const int8_t *f,
                                    uint32_t *flags, void *stream)
{
    RZK_TRY(check_ready(e));
    if (any_null({c, x, r, flags})) return fail(e, RZK_ERR_INVALID, "null argument");
    Guard g(e->device);
    return dev_commitment_verify(e, B, c, x, r, f, flags, (cudaStream_t)stream);
}

int rzk_open_commit_batch_dev(rzk_engine *e, size_t B, const int32_t *x, const int8_t *r, const int32_t *y,
                              int32_t *c, int32_t *t, uint32_t *flags, void *stream)
{
    RZK_TRY(check_ready(e));
    if (any_null({x, r, y, c, t, flags})) return fail(e, RZK_ERR_INVALID, "null argument");
    Guard g(e->device);
    return dev_open_commit(e, B, x, r, y, c, t, flags, (cudaStream_t)stream);
}

int rzk_open_respond_batch_dev(rzk_engine *e, size_t B, const int32_t *y, const int8_t *r, const int8_t *d, int32_t *z, void *stream)
{
    RZK_TRY(check_ready(e));
    if (any_null({y, r, d, z})) return fail(e, RZK_ERR_INVALID, "null argument");
    Guard g(e->device);
    RZK_TRY(ensure_need(e, B + 1));
    return dev_respond(e, B, y, r, d, 1, z, e->d_need, (cudaStream_t)stream);
}

int rzk_open_verify_batch_dev(rzk_engine *e, size_t B, const int32_t *z, const int32_t *t, const int32_t *c, uint32_t c_stride,
                              const int8_t *d, uint32_t *flags, void *stream)
{
    RZK_TRY(check_ready(e));
    if (any_null({z, t, c, d, flags}) || c_stride < 1) return fail(e, RZK_ERR_INVALID, "null argument");
    Guard g(e->device);
    return dev_verify_first(e, B, z, t, c, c_stride, d, 1, nullptr, flags, 1, (cudaStream_t)stream);
}

int rzk_linear_commit_batch_dev(rzk_engine *e, size_t B, const int32_t *g, const int32_t *x, const int8_t *rp, const int8_t *r,
                                const int32_t *y, const int32_t *yp, int32_t *gx, int32_t *cp, int32_t *c, int32_t *t,
                                int32_t *tp, int32_t *u, uint32_t *flags, void *stream)
{
    RZK_TRY(check_ready(e));
    if (any_null({g, x, rp, r, y, yp, gx, cp, c, t, tp, u, flags})) return fail(e, RZK_ERR_INVALID, "null argument");
    Guard gd(e->device);
    RZK_TRY(ensure_scratch(e, 2 * B * kPolyBytes));
    return dev_linear_commit(e, B, g, x, rp, r, y, yp, gx, cp, c, t, tp, u, flags, (int32_t *)e->scratch, (cudaStream_t)stream);
}

int rzk_linear_respond_batch_dev(rzk_engine *e, size_t B, const int32_t *y, const int32_t *yp, const int8_t *r, const int8_t *rp,
                                 const int8_t *d, int32_t *z, int32_t *zp, void *stream)
{
    RZK_TRY(check_ready(e));
    if (any_null({y, yp, r, rp, d, z, zp})) return fail(e, RZK_ERR_INVALID, "null argument");
    Guard g(e->device);
    RZK_TRY(ensure_need(e, 2 * (B + 1)));
    RZK_TRY(dev_respond(e, B, y, r, d, 1, z, e->d_need, (cudaStream_t)stream));                  // linear.rs:150-152
    return dev_respond(e, B, yp, rp, d, 1, zp, e->d_need + B + 1, (cudaStream_t)stream);         // linear.rs:154-156
}

int rzk_linear_verify_batch_dev(rzk_engine *e, size_t B, const int32_t *z, const int32_t *zp, const int32_t *c, const int32_t *cp,
                                const int32_t *g, const int32_t *t, const int32_t *tp, const int32_t *u, const int8_t *d,
                                uint32_t *flags, void *stream)
{
    RZK_TRY(check_ready(e));
    if (any_null({z, zp, c, cp, g, t, tp, u, d, flags})) return fail(e, RZK_ERR_INVALID, "null argument");
    Guard gd(e->device);
    RZK_TRY(ensure_scratch(e, 2 * B * kPolyBytes));
    return dev_linear_verify(e, B, z, zp, c, cp, g, t, tp, u, d, flags, (int32_t *)e->scratch, (cudaStream_t)stream);
}

int rzk_sum_commit_batch_dev(rzk_engine *e, size_t B, uint32_t T, const int32_t *gs, const int32_t *xs, const int8_t *rp,
                             const int8_t *rs, const int32_t *ys, const int32_t *yp, int32_t *xp, int32_t *cp, int32_t *cs,
                             int32_t *ts, int32_t *tp, int32_t *u, uint32_t *flags, void *stream)
{
    RZK_TRY(check_ready(e));
    if (T == 0 || T > 65535) return fail(e, RZK_ERR_INVALID, "T must be in 1..65535 (sum.rs:105)");
    if (any_null({gs, xs, rp, rs, ys, yp, xp, cp, cs, ts, tp, u, flags})) return fail(e, RZK_ERR_INVALID, "null argument");
    Guard gd(e->device);
    RZK_TRY(ensure_scratch(e, (B * T + B) * kPolyBytes));
    return dev_sum_commit(e, B, T, gs, xs, rp, rs, ys, yp, xp, cp, cs, ts, tp, u, flags, (int32_t *)e->scratch, (cudaStream_t)stream);
}

int rzk_sum_respond_batch_dev(rzk_engine *e, size_t B, uint32_t T, const int32_t *ys, const int32_t *yp, const int8_t *rs,
                              const int8_t *rp, const int8_t *d, int32_t *zs, int32_t *zp, void *stream)
{
    RZK_TRY(check_ready(e));
    if (T == 0 || T > 65535) return fail(e, RZK_ERR_INVALID, "T must be in 1..65535");
    if (any_null({ys, yp, rs, rp, d, zs, zp})) return fail(e, RZK_ERR_INVALID, "null argument");
    Guard g(e->device);
    RZK_TRY(ensure_need(e, B * T + 1 + B + 1));
    RZK_TRY(dev_respond(e, B * T, ys, rs, d, T, zs, e->d_need, (cudaStream_t)stream));           // sum.rs:188-193
    return dev_respond(e, B, yp, rp, d, 1, zp, e->d_need + B * T + 1, (cudaStream_t)stream);     // sum.rs:195-197
}

int rzk_sum_verify_batch_dev(rzk_engine *e, size_t B, uint32_t T, const int32_t *zs, const int32_t *zp, const int32_t *cs,
                             const int32_t *cp, const int32_t *gs, const int32_t *ts, const int32_t *tp, const int32_t *u,
                             const int8_t *d, uint32_t *flags, void *stream)
{
    RZK_TRY(check_ready(e));
    if (T == 0 || T > 65535) return fail(e, RZK_ERR_INVALID, "T must be in 1..65535");
    if (any_null({zs, zp, cs, cp, gs, ts, tp, u, d, flags})) return fail(e, RZK_ERR_INVALID, "null argument");
    Guard gd(e->device);
    RZK_TRY(ensure_scratch(e, (B * T + 3 * B) * kPolyBytes));
    return dev_sum_verify(e, B, T, zs, zp, cs, cp, gs, ts, tp, u, d, flags, (int32_t *)e->scratch, (cudaStream_t)stream);
}

// ---- optional on-device samplers (not part of the reference's flow: see rzk_sample.cuh) ----

int rzk_sample_small_dev(rzk_engine *e, size_t n_polys, int32_t b, uint64_t seed, uint32_t tag, int8_t *out, void *stream)
{
    RZK_TRY(check_ready(e, false));
    if (!out || b < 1 || b > 127 || tag >= (1u << 24)) return fail(e, RZK_ERR_INVALID, "bad sampler argument");
    if (n_polys == 0) return RZK_OK;
    Guard g(e->device);
    const size_t total = n_polys * (kN / 4);
    const unsigned grid = (unsigned)std::min<size_t>((total + 255) / 256, (size_t)e->num_sms * 8);
    rzk_sample_small_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(n_polys, (uint32_t)b, tag, (uint32_t)seed, (uint32_t)(seed >> 32), out);
    RZK_CUDA(e, cudaGetLastError());
    e->launches++;
    return RZK_OK;
}

int rzk_sample_gaussian_dev(rzk_engine *e, size_t n_polys, double sigma, uint64_t seed, uint32_t tag, int32_t *out, void *stream)
{
    RZK_TRY(check_ready(e, false));
    if (!out || !(sigma > 0.0) || sigma > 1.0e8 || tag >= (1u << 24)) return fail(e, RZK_ERR_INVALID, "bad sampler argument");
    if (n_polys == 0) return RZK_OK;
    Guard g(e->device);
    const size_t total = n_polys * (kN / 2);
    const unsigned grid = (unsigned)std::min<size_t>((total + 255) / 256, (size_t)e->num_sms * 8);
    rzk_sample_gaussian_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(n_polys, sigma, tag, (uint32_t)seed, (uint32_t)(seed >> 32), out);
    RZK_CUDA(e, cudaGetLastError());
    e->launches++;
    return RZK_OK;
}

int rzk_sample_challenge_dev(rzk_engine *e, size_t n_items, int32_t kappa, uint64_t seed, uint32_t tag, int8_t *out, void *stream)
{
    RZK_TRY(check_ready(e, false));
    if (!out || kappa < 1 || tag >= (1u << 24)) return fail(e, RZK_ERR_INVALID, "bad sampler argument");
    if (n_items == 0) return RZK_OK;
    Guard g(e->device);
    const unsigned grid = (unsigned)std::min<size_t>((n_items + 127) / 128, (size_t)e->num_sms * 8);
    rzk_sample_challenge_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(n_items, (uint32_t)kappa, tag, (uint32_t)seed, (uint32_t)(seed >> 32), out);
    RZK_CUDA(e, cudaGetLastError());
    e->launches++;
    return RZK_OK;
}

int rzk_flags_to_bitmap_dev(rzk_engine *e, size_t B, const uint32_t *flags, uint8_t *bitmap, uint32_t *range_any, void *stream)
{
    RZK_TRY(check_ready(e, false));
    if (any_null({flags, bitmap})) return fail(e, RZK_ERR_INVALID, "null argument");
    if (B == 0) return RZK_OK;
    Guard g(e->device);
    const size_t nbytes = (B + 7) / 8;
    rzk_flags_to_bitmap_kernel<<<(unsigned)((nbytes + 127) / 128), 128, 0, (cudaStream_t)stream>>>(B, flags, bitmap, range_any);
    RZK_CUDA(e, cudaGetLastError());
    e->launches++;
    return RZK_OK;
}

// ---- host entry points ----

int rzk_commit_batch(rzk_engine *e, size_t B, const int32_t *x, const int8_t *r, int32_t *c, uint8_t *ok)
{
    RZK_TRY(check_ready(e));
    if (any_null({x, r, c, ok})) return fail(e, RZK_ERR_INVALID, "null argument");
    std::vector<HArr> a = {{x, nullptr, kPolyBytes}, {r, nullptr, 3 * kN}, {nullptr, c, 2 * kPolyBytes}};
    // items with some |r| > 15 are redone by the two-prime program inside dev_commit (masked launch, same stream)
    return run_chunked(e, B, a, 0, ok, [&](size_t n, void **d, char *, uint32_t *fl, cudaStream_t s, uint32_t *rm) {
        return dev_commit(e, n, (const int32_t *)d[0], (const int8_t *)d[1], (int32_t *)d[2], fl, s, rm);
    });
}

int rzk_pack_r2(size_t count, const int8_t *r, uint8_t *r2)
{
    if (!r || !r2 || (count & 3)) return RZK_ERR_INVALID;
    for (size_t i = 0; i < count; i += 4) {
        uint32_t b = 0;
        for (int k = 0; k < 4; ++k) {
            const int v = r[i + k];
            if (v < -2 || v > 1) return RZK_ERR_RANGE;
            b |= ((uint32_t)v & 3u) << (2 * k);
        }
        r2[i >> 2] = (uint8_t)b;
    }
    return RZK_OK;
}

int rzk_unpack_r2_dev(rzk_engine *e, size_t count, const uint8_t *r2, int8_t *r, void *stream)
{
    RZK_TRY(check_ready(e, false));
    if (any_null({r2, r})) return fail(e, RZK_ERR_INVALID, "null argument");
    if (count & 15) return fail(e, RZK_ERR_INVALID, "count must be a multiple of 16 coefficients");
    if (count == 0) return RZK_OK;
    Guard g(e->device);
    const size_t nwords = count / 16;
    const unsigned blocks = (unsigned)std::min<size_t>((nwords + 255) / 256, (size_t)e->num_sms * 8);
    rzk_unpack_r2_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(nwords, reinterpret_cast<const uint32_t *>(r2), reinterpret_cast<uint4 *>(r));
    RZK_CUDA(e, cudaGetLastError());
    e->launches++;
    return RZK_OK;
}

int rzk_commit_batch_r2(rzk_engine *e, size_t B, const int32_t *x, const uint8_t *r2, int32_t *c, uint8_t *ok)
{
    RZK_TRY(check_ready(e));
    if (any_null({x, r2, c, ok})) return fail(e, RZK_ERR_INVALID, "null argument");
    // 384 bytes of randomness per commitment cross the bus instead of 1536; the int8 rows the program reads are unpacked
    // into the chunk's scratch by a small kernel on the same stream
    std::vector<HArr> a = {{x, nullptr, kPolyBytes}, {r2, nullptr, 3 * kN / 4}, {nullptr, c, 2 * kPolyBytes}};
    return run_chunked(e, B, a, 3 * kN, ok, [&](size_t n, void **d, char *sc, uint32_t *fl, cudaStream_t s, uint32_t *rm) {
        RZK_TRY(rzk_unpack_r2_dev(e, n * 3 * kN, (const uint8_t *)d[1], (int8_t *)sc, s));
        return dev_commit(e, n, (const int32_t *)d[0], (const int8_t *)sc, (int32_t *)d[2], fl, s, rm);
    });
}

int rzk_open_commit_batch(rzk_engine *e, size_t B, const int32_t *x, const int8_t *r, const int32_t *y,
                          int32_t *c, int32_t *t, uint8_t *ok)
{
    RZK_TRY(check_ready(e));
    if (any_null({x, r, y, c, t, ok})) return fail(e, RZK_ERR_INVALID, "null argument");
    std::vector<HArr> a = {{x, nullptr, kPolyBytes}, {r, nullptr, 3 * kN}, {y, nullptr, 3 * kPolyBytes},
                           {nullptr, c, 2 * kPolyBytes}, {nullptr, t, kPolyBytes}};
    return run_chunked(e, B, a, 0, ok, [&](size_t n, void **d, char *, uint32_t *fl, cudaStream_t s, uint32_t *rm) {
        return dev_open_commit(e, n, (const int32_t *)d[0], (const int8_t *)d[1], (const int32_t *)d[2],
                               (int32_t *)d[3], (int32_t *)d[4], fl, s, rm);
    });
}

int rzk_open_respond_batch(rzk_engine *e, size_t B, const int32_t *y, const int8_t *r, const int8_t *dch, int32_t *z)
{
    RZK_TRY(check_ready(e));
    if (any_null({y, r, dch, z})) return fail(e, RZK_ERR_INVALID, "null argument");
    std::vector<HArr> a = {{y, nullptr, 3 * kPolyBytes}, {r, nullptr, 3 * kN}, {dch, nullptr, kN}, {nullptr, z, 3 * kPolyBytes}};
    return run_chunked(e, B, a, 2 * sizeof(uint32_t), nullptr, [&](size_t n, void **d, char *sc, uint32_t *, cudaStream_t s, uint32_t *) {
        return dev_respond(e, n, (const int32_t *)d[0], (const int8_t *)d[1], (const int8_t *)d[2], 1, (int32_t *)d[3], (uint32_t *)sc, s);
    });
}

int rzk_commitment_verify_batch(rzk_engine *e, size_t B, const int32_t *c, const int32_t *x, const int8_t *r, const int8_t *f, uint8_t *bm)
{
    RZK_TRY(check_ready(e));
    if (any_null({c, x, r, bm})) return fail(e, RZK_ERR_INVALID, "null argument");
    std::vector<HArr> a = {{c, nullptr, 2 * kPolyBytes}, {x, nullptr, kPolyBytes}, {r, nullptr, 3 * kN}};
    if (f) a.push_back({f, nullptr, kN});
    return run_chunked(e, B, a, 0, bm, [&](size_t n, void **d, char *, uint32_t *fl, cudaStream_t s, uint32_t *rm) {
        return dev_commitment_verify(e, n, (const int32_t *)d[0], (const int32_t *)d[1], (const int8_t *)d[2],
                                     f ? (const int8_t *)d[3] : nullptr, fl, s);
    });
}

int rzk_open_verify_batch(rzk_engine *e, size_t B, const int32_t *z, const int32_t *t, const int32_t *c1, const int8_t *dch, uint8_t *bm)
{
    RZK_TRY(check_ready(e));
    if (any_null({z, t, c1, dch, bm})) return fail(e, RZK_ERR_INVALID, "null argument");
    std::vector<HArr> a = {{z, nullptr, 3 * kPolyBytes}, {t, nullptr, kPolyBytes}, {c1, nullptr, kPolyBytes}, {dch, nullptr, kN}};
    return run_chunked(e, B, a, 0, bm, [&](size_t n, void **d, char *, uint32_t *fl, cudaStream_t s, uint32_t *rm) {
        return dev_verify_first(e, n, (const int32_t *)d[0], (const int32_t *)d[1], (const int32_t *)d[2], 1,
                                (const int8_t *)d[3], 1, nullptr, fl, 1, s);
    });
}

int rzk_linear_commit_batch(rzk_engine *e, size_t B, const int32_t *g, const int32_t *x, const int8_t *rp, const int8_t *r,
                            const int32_t *y, const int32_t *yp, int32_t *gx, int32_t *cp, int32_t *c, int32_t *t,
                            int32_t *tp, int32_t *u, uint8_t *ok)
{
    RZK_TRY(check_ready(e));
    if (any_null({g, x, rp, r, y, yp, gx, cp, c, t, tp, u, ok})) return fail(e, RZK_ERR_INVALID, "null argument");
    std::vector<HArr> a = {{g, nullptr, kPolyBytes}, {x, nullptr, kPolyBytes}, {rp, nullptr, 3 * kN}, {r, nullptr, 3 * kN},
                           {y, nullptr, 3 * kPolyBytes}, {yp, nullptr, 3 * kPolyBytes},
                           {nullptr, gx, kPolyBytes}, {nullptr, cp, 2 * kPolyBytes}, {nullptr, c, 2 * kPolyBytes},
                           {nullptr, t, kPolyBytes}, {nullptr, tp, kPolyBytes}, {nullptr, u, kPolyBytes}};
    return run_chunked(e, B, a, 2 * kPolyBytes, ok, [&](size_t n, void **d, char *sc, uint32_t *fl, cudaStream_t s, uint32_t *rm) {
        return dev_linear_commit(e, n, (const int32_t *)d[0], (const int32_t *)d[1], (const int8_t *)d[2], (const int8_t *)d[3],
                                 (const int32_t *)d[4], (const int32_t *)d[5], (int32_t *)d[6], (int32_t *)d[7], (int32_t *)d[8],
                                 (int32_t *)d[9], (int32_t *)d[10], (int32_t *)d[11], fl, (int32_t *)sc, s, rm);
    });
}

int rzk_linear_respond_batch(rzk_engine *e, size_t B, const int32_t *y, const int32_t *yp, const int8_t *r, const int8_t *rp,
                             const int8_t *dch, int32_t *z, int32_t *zp)
{
    RZK_TRY(check_ready(e));
    if (any_null({y, yp, r, rp, dch, z, zp})) return fail(e, RZK_ERR_INVALID, "null argument");
    std::vector<HArr> a = {{y, nullptr, 3 * kPolyBytes}, {yp, nullptr, 3 * kPolyBytes}, {r, nullptr, 3 * kN}, {rp, nullptr, 3 * kN},
                           {dch, nullptr, kN}, {nullptr, z, 3 * kPolyBytes}, {nullptr, zp, 3 * kPolyBytes}};
    return run_chunked(e, B, a, 4 * sizeof(uint32_t), nullptr, [&](size_t n, void **d, char *sc, uint32_t *, cudaStream_t s, uint32_t *) {
        uint32_t *need = (uint32_t *)sc;
        RZK_TRY(dev_respond(e, n, (const int32_t *)d[0], (const int8_t *)d[2], (const int8_t *)d[4], 1, (int32_t *)d[5], need, s));
        return dev_respond(e, n, (const int32_t *)d[1], (const int8_t *)d[3], (const int8_t *)d[4], 1, (int32_t *)d[6], need + n + 1, s);
    });
}

int rzk_linear_verify_batch(rzk_engine *e, size_t B, const int32_t *z, const int32_t *zp, const int32_t *c, const int32_t *cp,
                            const int32_t *g, const int32_t *t, const int32_t *tp, const int32_t *u, const int8_t *dch, uint8_t *bm)
{
    RZK_TRY(check_ready(e));
    if (any_null({z, zp, c, cp, g, t, tp, u, dch, bm})) return fail(e, RZK_ERR_INVALID, "null argument");
    std::vector<HArr> a = {{z, nullptr, 3 * kPolyBytes}, {zp, nullptr, 3 * kPolyBytes}, {c, nullptr, 2 * kPolyBytes},
                           {cp, nullptr, 2 * kPolyBytes}, {g, nullptr, kPolyBytes}, {t, nullptr, kPolyBytes},
                           {tp, nullptr, kPolyBytes}, {u, nullptr, kPolyBytes}, {dch, nullptr, kN}};
    return run_chunked(e, B, a, 2 * kPolyBytes, bm, [&](size_t n, void **d, char *sc, uint32_t *fl, cudaStream_t s, uint32_t *rm) {
        return dev_linear_verify(e, n, (const int32_t *)d[0], (const int32_t *)d[1], (const int32_t *)d[2], (const int32_t *)d[3],
                                 (const int32_t *)d[4], (const int32_t *)d[5], (const int32_t *)d[6], (const int32_t *)d[7],
                                 (const int8_t *)d[8], fl, (int32_t *)sc, s);
    });
}

int rzk_sum_commit_batch(rzk_engine *e, size_t B, uint32_t T, const int32_t *gs, const int32_t *xs, const int8_t *rp,
                         const int8_t *rs, const int32_t *ys, const int32_t *yp, int32_t *xp, int32_t *cp, int32_t *cs,
                         int32_t *ts, int32_t *tp, int32_t *u, uint8_t *ok)
{
    RZK_TRY(check_ready(e));
    if (T == 0 || T > 65535) return fail(e, RZK_ERR_INVALID, "T must be in 1..65535 (sum.rs:105)");
    if (any_null({gs, xs, rp, rs, ys, yp, xp, cp, cs, ts, tp, u, ok})) return fail(e, RZK_ERR_INVALID, "null argument");
    std::vector<HArr> a = {{gs, nullptr, T * kPolyBytes}, {xs, nullptr, T * kPolyBytes}, {rp, nullptr, 3 * kN},
                           {rs, nullptr, (size_t)T * 3 * kN}, {ys, nullptr, T * 3 * kPolyBytes}, {yp, nullptr, 3 * kPolyBytes},
                           {nullptr, xp, kPolyBytes}, {nullptr, cp, 2 * kPolyBytes}, {nullptr, cs, T * 2 * kPolyBytes},
                           {nullptr, ts, T * kPolyBytes}, {nullptr, tp, kPolyBytes}, {nullptr, u, kPolyBytes}};
    return run_chunked(e, B, a, (T + 1) * kPolyBytes, ok, [&](size_t n, void **d, char *sc, uint32_t *fl, cudaStream_t s, uint32_t *rm) {
        return dev_sum_commit(e, n, T, (const int32_t *)d[0], (const int32_t *)d[1], (const int8_t *)d[2], (const int8_t *)d[3],
                              (const int32_t *)d[4], (const int32_t *)d[5], (int32_t *)d[6], (int32_t *)d[7], (int32_t *)d[8],
                              (int32_t *)d[9], (int32_t *)d[10], (int32_t *)d[11], fl, (int32_t *)sc, s, rm);
    });
}

int rzk_sum_respond_batch(rzk_engine *e, size_t B, uint32_t T, const int32_t *ys, const int32_t *yp, const int8_t *rs,
                          const int8_t *rp, const int8_t *dch, int32_t *zs, int32_t *zp)
{
    RZK_TRY(check_ready(e));
    if (T == 0 || T > 65535) return fail(e, RZK_ERR_INVALID, "T must be in 1..65535");
    if (any_null({ys, yp, rs, rp, dch, zs, zp})) return fail(e, RZK_ERR_INVALID, "null argument");
    std::vector<HArr> a = {{ys, nullptr, T * 3 * kPolyBytes}, {yp, nullptr, 3 * kPolyBytes}, {rs, nullptr, (size_t)T * 3 * kN},
                           {rp, nullptr, 3 * kN}, {dch, nullptr, kN}, {nullptr, zs, T * 3 * kPolyBytes}, {nullptr, zp, 3 * kPolyBytes}};
    return run_chunked(e, B, a, ((size_t)T + 3) * sizeof(uint32_t), nullptr, [&](size_t n, void **d, char *sc, uint32_t *, cudaStream_t s, uint32_t *) {
        uint32_t *need = (uint32_t *)sc;
        RZK_TRY(dev_respond(e, n * T, (const int32_t *)d[0], (const int8_t *)d[2], (const int8_t *)d[4], T, (int32_t *)d[5], need, s));
        return dev_respond(e, n, (const int32_t *)d[1], (const int8_t *)d[3], (const int8_t *)d[4], 1, (int32_t *)d[6], need + n * T + 1, s);
    });
}

int rzk_sum_verify_batch(rzk_engine *e, size_t B, uint32_t T, const int32_t *zs, const int32_t *zp, const int32_t *cs,
                         const int32_t *cp, const int32_t *gs, const int32_t *ts, const int32_t *tp, const int32_t *u,
                         const int8_t *dch, uint8_t *bm)
{
    RZK_TRY(check_ready(e));
    if (T == 0 || T > 65535) return fail(e, RZK_ERR_INVALID, "T must be in 1..65535");
    if (any_null({zs, zp, cs, cp, gs, ts, tp, u, dch, bm})) return fail(e, RZK_ERR_INVALID, "null argument");
    std::vector<HArr> a = {{zs, nullptr, T * 3 * kPolyBytes}, {zp, nullptr, 3 * kPolyBytes}, {cs, nullptr, T * 2 * kPolyBytes},
                           {cp, nullptr, 2 * kPolyBytes}, {gs, nullptr, T * kPolyBytes}, {ts, nullptr, T * kPolyBytes},
                           {tp, nullptr, kPolyBytes}, {u, nullptr, kPolyBytes}, {dch, nullptr, kN}};
    return run_chunked(e, B, a, (T + 3) * kPolyBytes, bm, [&](size_t n, void **d, char *sc, uint32_t *fl, cudaStream_t s, uint32_t *rm) {
        return dev_sum_verify(e, n, T, (const int32_t *)d[0], (const int32_t *)d[1], (const int32_t *)d[2], (const int32_t *)d[3],
                              (const int32_t *)d[4], (const int32_t *)d[5], (const int32_t *)d[6], (const int32_t *)d[7],
                              (const int8_t *)d[8], fl, (int32_t *)sc, s);
    });
}

// ---- i64 staging ----

int rzk_pack_i64(rzk_engine *e, size_t count, const int64_t *src, int32_t *dst)
{
    RZK_TRY(check_ready(e, false));
    if (any_null({src, dst})) return fail(e, RZK_ERR_INVALID, "null argument");
    const int64_t q = e->P.q;
    size_t blocks = (count + kN - 1) / kN;
    std::vector<HArr> a = {{src, nullptr, kN * sizeof(int64_t)}, {nullptr, dst, kN * sizeof(int32_t)}};
    // whole polynomials per "item"; a ragged tail is handled by a final short call
    size_t whole = count / kN;
    int rc = run_chunked(e, whole, a, 0, nullptr, [&](size_t n, void **d, char *, uint32_t *, cudaStream_t s, uint32_t *) {
        rzk_pack_i64_kernel<<<e->num_sms * 4, 256, 0, s>>>(n * kN, (const int64_t *)d[0], (int32_t *)d[1], q);
        e->launches++;
        return cudaGetLastError() == cudaSuccess ? RZK_OK : RZK_ERR_CUDA;
    });
    (void)blocks;
    if (rc != RZK_OK) return rc;
    const int64_t half = (q - 1) / 2;
    for (size_t i = whole * kN; i < count; ++i) {      // < N leftover coefficients: scalar staging
        int64_t r = src[i] % q;
        if (r > half) r -= q; else if (r < -half) r += q;
        dst[i] = (int32_t)r;
    }
    return RZK_OK;
}

int rzk_unpack_i64(rzk_engine *e, size_t count, const int32_t *src, int64_t *dst)
{
    RZK_TRY(check_ready(e, false));
    if (any_null({src, dst})) return fail(e, RZK_ERR_INVALID, "null argument");
    std::vector<HArr> a = {{src, nullptr, kN * sizeof(int32_t)}, {nullptr, dst, kN * sizeof(int64_t)}};
    size_t whole = count / kN;
    int rc = run_chunked(e, whole, a, 0, nullptr, [&](size_t n, void **d, char *, uint32_t *, cudaStream_t s, uint32_t *) {
        rzk_unpack_i64_kernel<<<e->num_sms * 4, 256, 0, s>>>(n * kN, (const int32_t *)d[0], (int64_t *)d[1]);
        e->launches++;
        return cudaGetLastError() == cudaSuccess ? RZK_OK : RZK_ERR_CUDA;
    });
    if (rc != RZK_OK) return rc;
    for (size_t i = whole * kN; i < count; ++i) dst[i] = src[i];
    return RZK_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------ several GPUs behind one handle

struct rzk_group {
    std::vector<rzk_engine *> eng;
    std::string err;
};

namespace {

// poly counts are in units of N coefficients per item; o(ptr, per_item, start) offsets a typed pointer
template <class T>
T *off(T *p, size_t per_item_elems, size_t start) { return p ? p + per_item_elems * start : nullptr; }

// fn(engine, start, count) on one worker thread per engine over contiguous 8-aligned ranges
template <class F>
int group_run(rzk_group *g, size_t B, F &&fn)
{
    if (!g || g->eng.empty()) return RZK_ERR_INVALID;
    const size_t n = g->eng.size();
    size_t per = (B + n - 1) / n;
    per = (per + 7) / 8 * 8;
    std::vector<int> rc(n, RZK_OK);
    std::vector<std::thread> th;
    for (size_t i = 0; i < n; ++i) {
        const size_t lo = std::min(B, i * per), hi = std::min(B, lo + per);
        if (hi == lo) continue;
        th.emplace_back([&, i, lo, hi] { rc[i] = fn(g->eng[i], lo, hi - lo); });
    }
    for (auto &t : th) t.join();
    for (size_t i = 0; i < n; ++i)
        if (rc[i] != RZK_OK) {
            g->err = "device " + std::to_string(g->eng[i]->device) + ": " + g->eng[i]->err;
            return rc[i];
        }
    return RZK_OK;
}

}  // namespace

extern "C" {

int rzk_group_create(const rzk_params *params, const int *device_ids, int n_devices, rzk_group **out)
{
    if (!params || !device_ids || n_devices < 1 || !out) return fail(nullptr, RZK_ERR_INVALID, "null argument");
    *out = nullptr;
    rzk_group *g = new rzk_group();
    for (int i = 0; i < n_devices; ++i) {
        rzk_engine *e = nullptr;
        const int rc = rzk_create(params, device_ids[i], &e);
        if (rc != RZK_OK) { rzk_group_destroy(g); return rc; }
        g->eng.push_back(e);
    }
    *out = g;
    return RZK_OK;
}

void rzk_group_destroy(rzk_group *g)
{
    if (!g) return;
    for (auto *e : g->eng) rzk_destroy(e);
    delete g;
}

int rzk_group_size(const rzk_group *g) { return g ? (int)g->eng.size() : 0; }
const char *rzk_group_last_error(const rzk_group *g) { return g ? g->err.c_str() : g_create_err.c_str(); }

uint64_t rzk_group_kernel_launches(const rzk_group *g)
{
    uint64_t n = 0;
    if (g) for (auto *e : g->eng) n += e->launches;
    return n;
}

int rzk_group_set_key(rzk_group *g, const int64_t *a1, const int64_t *a2)
{
    if (!g) return RZK_ERR_INVALID;
    for (auto *e : g->eng) {
        const int rc = rzk_set_key(e, a1, a2);
        if (rc != RZK_OK) { g->err = e->err; return rc; }
    }
    return RZK_OK;
}

#define POLY(p, polys, s) off(p, (size_t)(polys) * kN, s)
#define BITS(p, s) ((p) ? (p) + (s) / 8 : nullptr)

int rzk_group_commit_batch(rzk_group *g, size_t B, const int32_t *x, const int8_t *r, int32_t *c, uint8_t *ok)
{
    return group_run(g, B, [&](rzk_engine *e, size_t s, size_t n) {
        return rzk_commit_batch(e, n, POLY(x, 1, s), POLY(r, 3, s), POLY(c, 2, s), BITS(ok, s));
    });
}

int rzk_group_commit_batch_r2(rzk_group *g, size_t B, const int32_t *x, const uint8_t *r2, int32_t *c, uint8_t *ok)
{
    return group_run(g, B, [&](rzk_engine *e, size_t s, size_t n) {
        return rzk_commit_batch_r2(e, n, POLY(x, 1, s), r2 + s * (3 * kN / 4), POLY(c, 2, s), BITS(ok, s));
    });
}

int rzk_group_commitment_verify_batch(rzk_group *g, size_t B, const int32_t *c, const int32_t *x, const int8_t *r,
                                      const int8_t *f, uint8_t *bm)
{
    return group_run(g, B, [&](rzk_engine *e, size_t s, size_t n) {
        return rzk_commitment_verify_batch(e, n, POLY(c, 2, s), POLY(x, 1, s), POLY(r, 3, s), POLY(f, 1, s), BITS(bm, s));
    });
}

int rzk_group_open_commit_batch(rzk_group *g, size_t B, const int32_t *x, const int8_t *r, const int32_t *y,
                                int32_t *c, int32_t *t, uint8_t *ok)
{
    return group_run(g, B, [&](rzk_engine *e, size_t s, size_t n) {
        return rzk_open_commit_batch(e, n, POLY(x, 1, s), POLY(r, 3, s), POLY(y, 3, s), POLY(c, 2, s), POLY(t, 1, s), BITS(ok, s));
    });
}

int rzk_group_open_respond_batch(rzk_group *g, size_t B, const int32_t *y, const int8_t *r, const int8_t *d, int32_t *z)
{
    return group_run(g, B, [&](rzk_engine *e, size_t s, size_t n) {
        return rzk_open_respond_batch(e, n, POLY(y, 3, s), POLY(r, 3, s), POLY(d, 1, s), POLY(z, 3, s));
    });
}

int rzk_group_open_verify_batch(rzk_group *g, size_t B, const int32_t *z, const int32_t *t, const int32_t *c1,
                                const int8_t *d, uint8_t *bm)
{
    return group_run(g, B, [&](rzk_engine *e, size_t s, size_t n) {
        return rzk_open_verify_batch(e, n, POLY(z, 3, s), POLY(t, 1, s), POLY(c1, 1, s), POLY(d, 1, s), BITS(bm, s));
    });
}

int rzk_group_linear_commit_batch(rzk_group *g, size_t B, const int32_t *gg, const int32_t *x, const int8_t *rp, const int8_t *r,
                                  const int32_t *y, const int32_t *yp, int32_t *gx, int32_t *cp, int32_t *c, int32_t *t,
                                  int32_t *tp, int32_t *u, uint8_t *ok)
{
    return group_run(g, B, [&](rzk_engine *e, size_t s, size_t n) {
        return rzk_linear_commit_batch(e, n, POLY(gg, 1, s), POLY(x, 1, s), POLY(rp, 3, s), POLY(r, 3, s), POLY(y, 3, s), POLY(yp, 3, s),
                                       POLY(gx, 1, s), POLY(cp, 2, s), POLY(c, 2, s), POLY(t, 1, s), POLY(tp, 1, s), POLY(u, 1, s), BITS(ok, s));
    });
}

int rzk_group_linear_respond_batch(rzk_group *g, size_t B, const int32_t *y, const int32_t *yp, const int8_t *r, const int8_t *rp,
                                   const int8_t *d, int32_t *z, int32_t *zp)
{
    return group_run(g, B, [&](rzk_engine *e, size_t s, size_t n) {
        return rzk_linear_respond_batch(e, n, POLY(y, 3, s), POLY(yp, 3, s), POLY(r, 3, s), POLY(rp, 3, s), POLY(d, 1, s),
                                        POLY(z, 3, s), POLY(zp, 3, s));
    });
}

int rzk_group_linear_verify_batch(rzk_group *g, size_t B, const int32_t *z, const int32_t *zp, const int32_t *c, const int32_t *cp,
                                  const int32_t *gg, const int32_t *t, const int32_t *tp, const int32_t *u, const int8_t *d, uint8_t *bm)
{
    return group_run(g, B, [&](rzk_engine *e, size_t s, size_t n) {
        return rzk_linear_verify_batch(e, n, POLY(z, 3, s), POLY(zp, 3, s), POLY(c, 2, s), POLY(cp, 2, s), POLY(gg, 1, s), POLY(t, 1, s),
                                       POLY(tp, 1, s), POLY(u, 1, s), POLY(d, 1, s), BITS(bm, s));
    });
}

int rzk_group_sum_commit_batch(rzk_group *g, size_t B, uint32_t T, const int32_t *gs, const int32_t *xs, const int8_t *rp,
                               const int8_t *rs, const int32_t *ys, const int32_t *yp, int32_t *xp, int32_t *cp, int32_t *cs,
                               int32_t *ts, int32_t *tp, int32_t *u, uint8_t *ok)
{
    return group_run(g, B, [&](rzk_engine *e, size_t s, size_t n) {
        return rzk_sum_commit_batch(e, n, T, POLY(gs, T, s), POLY(xs, T, s), POLY(rp, 3, s), POLY(rs, 3 * T, s), POLY(ys, 3 * T, s),
                                    POLY(yp, 3, s), POLY(xp, 1, s), POLY(cp, 2, s), POLY(cs, 2 * T, s), POLY(ts, T, s), POLY(tp, 1, s),
                                    POLY(u, 1, s), BITS(ok, s));
    });
}

int rzk_group_sum_respond_batch(rzk_group *g, size_t B, uint32_t T, const int32_t *ys, const int32_t *yp, const int8_t *rs,
                                const int8_t *rp, const int8_t *d, int32_t *zs, int32_t *zp)
{
    return group_run(g, B, [&](rzk_engine *e, size_t s, size_t n) {
        return rzk_sum_respond_batch(e, n, T, POLY(ys, 3 * T, s), POLY(yp, 3, s), POLY(rs, 3 * T, s), POLY(rp, 3, s), POLY(d, 1, s),
                                     POLY(zs, 3 * T, s), POLY(zp, 3, s));
    });
}

int rzk_group_sum_verify_batch(rzk_group *g, size_t B, uint32_t T, const int32_t *zs, const int32_t *zp, const int32_t *cs,
                               const int32_t *cp, const int32_t *gs, const int32_t *ts, const int32_t *tp, const int32_t *u,
                               const int8_t *d, uint8_t *bm)
{
    return group_run(g, B, [&](rzk_engine *e, size_t s, size_t n) {
        return rzk_sum_verify_batch(e, n, T, POLY(zs, 3 * T, s), POLY(zp, 3, s), POLY(cs, 2 * T, s), POLY(cp, 2, s), POLY(gs, T, s),
                                    POLY(ts, T, s), POLY(tp, 1, s), POLY(u, 1, s), POLY(d, 1, s), BITS(bm, s));
    });
}

#undef POLY
#undef BITS

}  // extern "C"

#include "rzk_wire.cuh"
#include "rzk_fs.cuh"
