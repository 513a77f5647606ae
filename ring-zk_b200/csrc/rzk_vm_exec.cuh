// rzk_vm_exec.cuh -- per-lane implementation of the polynomial-op program (rzk_vm.h).
//
// The same source is compiled twice:
//   * by nvcc for sm_100a, where RZK_NL == 1 and "each lane" is the calling thread;
//   * by g++ for the host lane emulator (tests/cpp/emu_check.cpp), where RZK_NL == 32 and the
//     lane loop over one warp is explicit.  The emulator exists so that the exact kernel
//     arithmetic, shared-memory layouts and lane exchanges are checked against the CPU oracle
//     without a GPU; it is test infrastructure, never a product fallback.
//
// Execution modes (template parameter MODE):
//   MODE_SEQ (np == 1 or 3): one half warp owns one item and runs each segment once per prime; the
//       residues of the earlier primes wait in a lane-private shared-memory stash.
//   MODE_SPLIT (np == 2): one warp owns one batch item; half warp h works modulo prime h.  After the
//       inverse transforms the two half warps swap half of their residues with warp shuffles, so
//       every lane recombines 16 coefficients (Garner CRT) -- no residue ever leaves registers.
//   MODE_SPLITKEY (np == 1): one warp owns one item whose small operands (|r| <= 15) meet the key.
//       The key coefficients are split a = a_lo + 2^16 a_hi (|a_lo|, |a_hi| <= 2^15), so each partial
//       product stays below 2^29 and ONE prime is exact.  Half warp 0 accumulates the lo images,
//       half warp 1 the hi images; the forward transforms are shared through shared memory (each
//       half warp transforms one operand) and the result is lo + 2^16 hi.  6 transforms per
//       commitment instead of 8, and no Garner step.
//
// Layouts (N = 512, 16 lanes per polynomial, 32 coefficients per lane in registers):
//   G1 (strided):    lane t holds coefficients i = t + 16*m,  m = 0..31   (stages 0..4: bits 8..4 are lane-local)
//   G2 (contiguous): lane t holds positions   i = 32*t + e,   e = 0..31   (stages 5..8: bits 3..0 are lane-local)
//   transpose buffer word of position i: i + 4*(i>>5)  (uint4 rows of 36 words -> conflict-free)
//   lane-private slot word of element e: ((e>>2)*16 + t)*4 + (e&3)
//
// Replaces Polynomial `*`/`+`/`-`/`==` as composed by Mat::dot, add, sub and
// componentwise_mul (/root/reference/src/mat.rs:95-178).
#pragma once
#include <type_traits>
#include "rzk_arith.cuh"
#include "rzk_vm.h"
#include "rzk_programs.h"

#if defined(__CUDACC__)
#define RZK_VM __device__ __forceinline__
#define RZK_NL 1
#define RZK_SYNC() __syncwarp()
#define RZK_UNROLL _Pragma("unroll")
#define RZK_NOUNROLL _Pragma("unroll 1")
// one lane per thread: no loop, no indexing, so the lane state stays in registers
#define RZK_EACH_LANE if (constexpr int li_ = 0; true)
#else
#define RZK_VM inline
#define RZK_NL 32
#define RZK_SYNC() ((void)0)
#define RZK_UNROLL
#define RZK_NOUNROLL
#define RZK_EACH_LANE for (int li_ = 0; li_ < RZK_NL; ++li_)
struct uint4 { uint32_t x, y, z, w; };
struct uint2 { uint32_t x, y; };
#endif

#ifndef RZK_SMALL_LATE
#define RZK_SMALL_LATE 1      // MODE_SPLITKEY_S: fetch the int8 plain term of an epilogue after the inverse transform instead of before
#endif

#define RZK_LANE Lane &L = lanes[li_]; const LaneCtx &ctx = ctxs[li_]; const int t = ctx.t; (void)t; (void)L; (void)ctx

namespace rzk {

enum { MODE_SEQ = 0, MODE_SPLIT = 1, MODE_SPLITKEY = 2, MODE_SPLITKEY_S = 3, MODE_SEQ_S = 4 };
// MODE_SEQ_S: MODE_SEQ with three SMALL primes (slots 3..5, all within 2^17 above 2^26) and signed lazy arithmetic -- the
// large x large product sums of up to 64 terms (64 * 512 * 2^62 < p3 p4 p5 / 2 = 2^77.0): forward transforms of four
// instructions per butterfly, Montgomery products of five, reductions only where a run of sums needs one.
constexpr bool mode_seq(int mode) { return mode == MODE_SEQ || mode == MODE_SEQ_S; }
constexpr bool mode_signed(int mode) { return mode == MODE_SPLITKEY_S || mode == MODE_SEQ_S; }
constexpr int kSignedMaxTerms = 64;                  // terms of a product sum the three small primes hold for ANY int32 operands
// MODE_SPLITKEY_S: the split-key commitment for |r| <= 1 (Params::default(): b = 1) modulo ONE SMALL prime with signed lazy
// arithmetic (rzk_arith.cuh): |a_lo r1 + a_lo' r2 + r0| <= 2^25 + 127 < p/2 for p = 67153921, and 2^31 / p = 31.98 leaves room
// for nine forward stages without any correction (4-instruction butterflies) and for an inverse in decimation-in-time form
// (inv_g2_dit / inv_g1_dit: the same butterflies, 511 of them without a multiplication, four reductions per lane, then the
// twist psi^-i; the Gentleman-Sande form with its 14 reductions per transform is the RZK_INV_DIT=0 build).  Same program,
// same data flow as MODE_SPLITKEY.
constexpr bool mode_sk(int mode) { return mode == MODE_SPLITKEY || mode == MODE_SPLITKEY_S; }
constexpr uint32_t kStaticPrime0 = 1073692673u;      // kPrimeList[0] (rzk_tables.cpp); 4p - 1 < 2^32
constexpr uint32_t kStaticPrimeS = 67153921u;        // kPrimeList[kSignedSlot]: 2^26 + 45057, == 1 (mod 4096)
constexpr int kSignedSlot = 3;                       // its slot: twiddles and key images in the signed Shoup form
constexpr int kSignedShift = 26;                     // floor(log2 p) (sreduce)
constexpr uint32_t kAddCap = 0xFFFFFFFEu;            // >= 4p - 2 for every p < 2^30: min(a + b, kAddCap) == a + b in the butterflies

struct Lane {
    uint32_t cur[kElems];
    uint32_t acc0[kElems];   // accumulator 0 lives in registers; accumulator 1 in the lane-private smem slot ctx.acc1
    PrimeC pc;               // constants of the prime this lane currently works with
    int pi;                  // its index in the launch's prime list
    uint32_t cap;            // immediate bound of the butterfly adds (rzk_arith.cuh add_alu): 4p - 1 for a compile-time p, else kAddCap
    uint32_t fail;
    uint32_t rerr;
};

struct LaneCtx {
    uint32_t *buf;         // [kBufWords]  transpose buffer of this half warp
    uint32_t *slot;        // [kSlotWords] operand slot of OP_ST / OP_MACV
    uint32_t *slot_hw[2];  // the operand slots of both half warps of this warp (OP_LD)
    uint32_t *acc1;        // [kSlotWords] accumulator 1 (lane-private layout)
    uint32_t *stash;       // SEQ: [nstash][np-1][kSlotWords] residues of earlier primes
    uint32_t *red;         // reduction scratch shared by the lanes that own one item
    const uint32_t *g1;    // staged [np][2][kG1Words]
    const uint32_t *g2;    // staged [np][2][16][60]
    const uint32_t *key;   // staged [np][3][2][576]
    const uint32_t *twist; // staged [np][kTwistWords] (signed modes, RZK_INV_DIT)
    uint32_t item;         // item index (clamped to n_items-1 for idle lanes)
    int t;                 // lane within the half warp, 0..15
    int hw;                // half warp within the warp, 0..1
    int ridx;              // index of this lane among the lanes that own the item
    bool active;
    uint32_t *pp_count;    // device: releases of heavy windows executed so far by this warp (pp_mode 1 / 3)
};

#if defined(__CUDACC__)
// Keeps warps in the same region of the (large, unrolled) program so that they share instruction
// cache lines.  cta_sync = 1: the whole CTA; n >= 2: independent groups of 16/n... warps (named barriers),
// which lets the groups drift apart and mix their pipe usage.
__device__ __forceinline__ void cta_lockstep(const VmLaunch &K)
{
    if (K.pp_mode >= 10) {          // staggered groups (pp_start): lock-step inside each group only, barriers 1 .. G
        const uint32_t G = K.pp_mode / 10, per = (blockDim.x >> 5) / G;
        asm volatile("bar.sync %0, %1;" ::"r"((threadIdx.x >> 5) / per + 1u), "r"(per * 32u));
    } else if (K.cta_sync == 1 || K.cta_sync >= 8) {
        __syncthreads();
    } else if (K.cta_sync >= 2 && K.cta_sync < 8) {
        const uint32_t warps = blockDim.x >> 5;
        const uint32_t per = (warps + K.cta_sync - 1) / K.cta_sync;        // warps per group
        const uint32_t grp = (threadIdx.x >> 5) / per;
        const uint32_t cnt = min(per, warps - grp * per) * 32;
        asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "r"(cnt));
    }
}

// Phase mixing between the two halves of a CTA ("ping-pong", K.pp_mode):
// the multiply-heavy windows (forward transform; inverse transform) saturate the FMA-heavy pipe while
// the windows between them (global loads, int8 conversion, CRT / reduction mod q, stores) leave it idle.
// With all warps in the same window at the same time (lock-step, needed for instruction-cache
// locality: 32 KB L1.5 vs ~80 KB of unrolled program) the pipe idles about a third of the time.
// Two groups of warps, each internally in step, run half an item apart instead:
//   pp_mode 2: strict alternation -- a group enters a heavy window only when the other group has left
//              its own (named barriers 8 + g: 256 threads sync, the other 256 arrive);
//   pp_mode 1 / 3: free running after an initial offset (group 1 starts when group 0 leaves its
//              first / second heavy window).
__device__ __forceinline__ uint32_t pp_group() { return (threadIdx.x >> 5) >= (blockDim.x >> 6) ? 1u : 0u; }

// pp_mode >= 10 encodes G * 10 + n: G groups of consecutive warps (a group of 4 consecutive warps has one warp on every
// scheduler), each in lock-step internally, started in a staggered way -- group g + 1 starts when group g has left its
// n-th heavy window -- and free running afterwards.  Named barriers 8 + g (g = 1 .. G - 1) carry the start signal.
__device__ __forceinline__ uint32_t pp_stagger_groups(const VmLaunch &K) { return K.pp_mode >= 10 ? K.pp_mode / 10 : 0u; }
__device__ __forceinline__ uint32_t pp_stagger_group(const VmLaunch &K) { return (threadIdx.x >> 5) / ((blockDim.x >> 5) / pp_stagger_groups(K)); }

__device__ __forceinline__ void pp_acquire(const VmLaunch &K)
{
    if (K.pp_mode == 2) asm volatile("bar.sync %0, %1;" ::"r"(8u + pp_group()), "r"(blockDim.x) : "memory");
}

__device__ __forceinline__ void pp_release(const VmLaunch &K, uint32_t &pp_count)
{
    if (K.pp_mode == 2) asm volatile("bar.arrive %0, %1;" ::"r"(8u + (pp_group() ^ 1u)), "r"(blockDim.x) : "memory");
    else if (K.pp_mode >= 10) {
        ++pp_count;
        const uint32_t G = pp_stagger_groups(K), g = pp_stagger_group(K);
        if (pp_count == K.pp_mode % 10 && g + 1 < G)
            asm volatile("bar.arrive %0, %1;" ::"r"(8u + g + 1u), "r"(2u * (blockDim.x / G)) : "memory");
    } else if (K.pp_mode != 0) {
        ++pp_count;
        if (pp_count == (K.pp_mode == 1 ? 1u : 2u) && pp_group() == 0)
            asm volatile("bar.arrive 10, %0;" ::"r"(blockDim.x) : "memory");
    }
}

// start of the kernel: every group but the first waits for its predecessor's signal
__device__ __forceinline__ void pp_start(const VmLaunch &K)
{
    if (K.pp_mode == 2) {
        if (pp_group() == 1) asm volatile("bar.arrive 8, %0;" ::"r"(blockDim.x) : "memory");     // group 0 owns the first window
    } else if (K.pp_mode >= 10) {
        const uint32_t G = pp_stagger_groups(K), g = pp_stagger_group(K);
        if (g > 0) asm volatile("bar.sync %0, %1;" ::"r"(8u + g), "r"(2u * (blockDim.x / G)) : "memory");
    } else if (K.pp_mode != 0) {
        if (pp_group() == 1) asm volatile("bar.sync 10, %0;" ::"r"(blockDim.x) : "memory");       // start offset
    }
}

// end of the kernel: a group that never reached the hand-over point (no work) must still release its successor
__device__ __forceinline__ void pp_finish(const VmLaunch &K, uint32_t pp_count)
{
    if (K.pp_mode == 2) {
        if (pp_group() == 0) asm volatile("bar.sync 8, %0;" ::"r"(blockDim.x) : "memory");        // absorbs the last hand-over
    } else if (K.pp_mode >= 10) {
        const uint32_t G = pp_stagger_groups(K), g = pp_stagger_group(K);
        if (g + 1 < G && pp_count < K.pp_mode % 10)
            asm volatile("bar.arrive %0, %1;" ::"r"(8u + g + 1u), "r"(2u * (blockDim.x / G)) : "memory");
    } else if (K.pp_mode != 0) {
        if (pp_group() == 0 && pp_count < (K.pp_mode == 1 ? 1u : 2u)) asm volatile("bar.arrive 10, %0;" ::"r"(blockDim.x) : "memory");
    }
}
#endif

// ---------------------------------------------------------------- transforms

// One butterfly stage with compile-time geometry (all loops have constant trip counts so that
// they unroll fully and the 32 coefficients stay in registers).
//   G1 stages S = 0..4: distance 16>>S in the strided layout, lane-uniform twiddles
//   G2 stages S = 5..8: distance 256>>S in the contiguous layout, lane-specific twiddles
template <int S, int DIR>
RZK_VM void g1_stage(uint32_t (&a)[kElems], const uint2 *g1, uint32_t p, uint32_t p2, uint32_t z, uint32_t cap)
{
    constexpr int half = 16 >> S;
    RZK_UNROLL
    for (int b = 0; b < (1 << S); ++b) {
        const uint2 w = g1[(1 << S) + b];
        RZK_UNROLL
        for (int j = 0; j < half; ++j) {
            const int i0 = b * 2 * half + j;
            if (DIR == 0) ct_bfly(a[i0], a[i0 + half], w.x, w.y, p, p2, z, cap);
            else gs_bfly(a[i0], a[i0 + half], w.x, w.y, p, p2, z, cap);
        }
    }
}

template <int S, int DIR>
RZK_VM void g2_stage(uint32_t (&a)[kElems], const uint4 *tw4, uint32_t p, uint32_t p2, uint32_t z, uint32_t cap)
{
    constexpr int half = 256 >> S;                // 8,4,2,1
    constexpr int nb = 1 << (S - 4);              // 2,4,8,16 blocks
    constexpr int base = (S == 5) ? 0 : (S == 6) ? 1 : (S == 7) ? 3 : 7;
    RZK_UNROLL
    for (int b2 = 0; b2 < nb / 2; ++b2) {
        const uint4 q = tw4[base + b2];
        RZK_UNROLL
        for (int j = 0; j < half; ++j) {
            const int i0 = (2 * b2) * 2 * half + j;
            if (DIR == 0) ct_bfly(a[i0], a[i0 + half], q.x, q.y, p, p2, z, cap);
            else gs_bfly(a[i0], a[i0 + half], q.x, q.y, p, p2, z, cap);
        }
        RZK_UNROLL
        for (int j = 0; j < half; ++j) {
            const int i0 = (2 * b2 + 1) * 2 * half + j;
            if (DIR == 0) ct_bfly(a[i0], a[i0 + half], q.z, q.w, p, p2, z, cap);
            else gs_bfly(a[i0], a[i0 + half], q.z, q.w, p, p2, z, cap);
        }
    }
}

RZK_VM void fwd_g1(uint32_t (&a)[kElems], const uint32_t *g1tab, uint32_t p, uint32_t p2, uint32_t z, uint32_t cap)
{
    const uint2 *g1 = reinterpret_cast<const uint2 *>(g1tab);
    g1_stage<0, 0>(a, g1, p, p2, z, cap);
    g1_stage<1, 0>(a, g1, p, p2, z, cap);
    g1_stage<2, 0>(a, g1, p, p2, z, cap);
    g1_stage<3, 0>(a, g1, p, p2, z, cap);
    g1_stage<4, 0>(a, g1, p, p2, z, cap);
}

RZK_VM void fwd_g2(uint32_t (&a)[kElems], const uint32_t *tw, uint32_t p, uint32_t p2, uint32_t z, uint32_t cap)
{
    const uint4 *tw4 = reinterpret_cast<const uint4 *>(tw);
    g2_stage<5, 0>(a, tw4, p, p2, z, cap);
    g2_stage<6, 0>(a, tw4, p, p2, z, cap);
    g2_stage<7, 0>(a, tw4, p, p2, z, cap);
    g2_stage<8, 0>(a, tw4, p, p2, z, cap);
}

RZK_VM void inv_g2(uint32_t (&a)[kElems], const uint32_t *tw, uint32_t p, uint32_t p2, uint32_t z, uint32_t cap)
{
    const uint4 *tw4 = reinterpret_cast<const uint4 *>(tw);
    g2_stage<8, 1>(a, tw4, p, p2, z, cap);
    g2_stage<7, 1>(a, tw4, p, p2, z, cap);
    g2_stage<6, 1>(a, tw4, p, p2, z, cap);
    g2_stage<5, 1>(a, tw4, p, p2, z, cap);
}

RZK_VM void inv_g1(uint32_t (&a)[kElems], const uint32_t *g1tab, uint32_t p, uint32_t p2, uint32_t z, uint32_t cap)
{
    const uint2 *g1 = reinterpret_cast<const uint2 *>(g1tab);
    g1_stage<4, 1>(a, g1, p, p2, z, cap);
    g1_stage<3, 1>(a, g1, p, p2, z, cap);
    g1_stage<2, 1>(a, g1, p, p2, z, cap);
    g1_stage<1, 1>(a, g1, p, p2, z, cap);
    g1_stage<0, 1>(a, g1, p, p2, z, cap);
}

// ---- signed lazy transforms (MODE_SPLITKEY_S; rzk_arith.cuh): same geometry, same table layout (w centred, w' signed)
// SPECIAL: forward -- the tiny-input first stage (ct_bfly_s_tiny); inverse -- the biased last stage (gs_bfly_s_biased)
template <int S, int DIR, bool BIASED = false>
RZK_VM void g1_stage_s(uint32_t (&a)[kElems], const uint2 *g1, uint32_t mp)
{
    constexpr int half = 16 >> S;
    RZK_UNROLL
    for (int b = 0; b < (1 << S); ++b) {
        const uint2 w = g1[(1 << S) + b];
        RZK_UNROLL
        for (int j = 0; j < half; ++j) {
            const int i0 = b * 2 * half + j;
            if (DIR == 0 && BIASED) ct_bfly_s_tiny(a[i0], a[i0 + half], w.x);      // (forward: the flag marks the tiny-input first stage)
            else if (DIR == 0) ct_bfly_s(a[i0], a[i0 + half], w.x, w.y, mp);
            else if (BIASED) gs_bfly_s_biased(a[i0], a[i0 + half], w.x, w.y, mp);
            else gs_bfly_s(a[i0], a[i0 + half], w.x, w.y, mp);
        }
    }
}

template <int S, int DIR>
RZK_VM void g2_stage_s(uint32_t (&a)[kElems], const uint4 *tw4, uint32_t mp)
{
    constexpr int half = 256 >> S;                // 8,4,2,1
    constexpr int nb = 1 << (S - 4);              // 2,4,8,16 blocks
    constexpr int base = (S == 5) ? 0 : (S == 6) ? 1 : (S == 7) ? 3 : 7;
    RZK_UNROLL
    for (int b2 = 0; b2 < nb / 2; ++b2) {
        const uint4 q = tw4[base + b2];
        RZK_UNROLL
        for (int j = 0; j < half; ++j) {
            const int i0 = (2 * b2) * 2 * half + j;
            if (DIR == 0) ct_bfly_s(a[i0], a[i0 + half], q.x, q.y, mp);
            else gs_bfly_s(a[i0], a[i0 + half], q.x, q.y, mp);
        }
        RZK_UNROLL
        for (int j = 0; j < half; ++j) {
            const int i0 = (2 * b2 + 1) * 2 * half + j;
            if (DIR == 0) ct_bfly_s(a[i0], a[i0 + half], q.z, q.w, mp);
            else gs_bfly_s(a[i0], a[i0 + half], q.z, q.w, mp);
        }
    }
}

// forward: inputs |a| <= 5p/4 (a reduced or Shoup-scaled int32 operand; tiny for int8 rows), every stage adds at most 5p/4:
// below 12.5 p at the end, no correction anywhere
template <bool TINY>      // TINY: every input is in {-1, 0, 1}
RZK_VM void fwd_g1_s(uint32_t (&a)[kElems], const uint32_t *g1tab, uint32_t mp)
{
    const uint2 *g1 = reinterpret_cast<const uint2 *>(g1tab);
    g1_stage_s<0, 0, TINY>(a, g1, mp);
    g1_stage_s<1, 0>(a, g1, mp);
    g1_stage_s<2, 0>(a, g1, mp);
    g1_stage_s<3, 0>(a, g1, mp);
    g1_stage_s<4, 0>(a, g1, mp);
}

RZK_VM void fwd_g2_s(uint32_t (&a)[kElems], const uint32_t *tw, uint32_t mp)
{
    const uint4 *tw4 = reinterpret_cast<const uint4 *>(tw);
    g2_stage_s<5, 0>(a, tw4, mp);
    g2_stage_s<6, 0>(a, tw4, mp);
    g2_stage_s<7, 0>(a, tw4, mp);
    g2_stage_s<8, 0>(a, tw4, mp);
}

// inverse: a sum output doubles the magnitude, a product output resets it to 5p/4; a butterfly needs |x| + |y| < 2^31 =
// 31.98 p.  Inputs: at most two key products, |a| <= 5p/2.  Element e has taken the sum path at a stage exactly when the bit
// of e that the stage pairs is clear, so the magnitudes are known at compile time:
//   contiguous layout (distances 1, 2, 4, 8): the elements with bits 0..2 clear reach 20 p after three stages and are reduced
//   before the fourth; after it the elements with bit 3 clear, bit 2 clear and one of bits 1, 0 set hold 5 p or 10 p and are
//   reduced as well, so that every lane enters the strided layout with at most 5p/2 per element (the history of the first four
//   stages depends on the LANE there, so it has to be uniform);
//   strided layout (distances 1 .. 16): again the elements with bits 0..2 clear before the fourth stage.
// 14 reductions (two instructions each) per transform; the results stay below 20.2 p < 2^31.
RZK_VM void inv_g2_s(uint32_t (&a)[kElems], const uint32_t *tw, uint32_t mp)
{
    const uint4 *tw4 = reinterpret_cast<const uint4 *>(tw);
    g2_stage_s<8, 1>(a, tw4, mp);
    g2_stage_s<7, 1>(a, tw4, mp);
    g2_stage_s<6, 1>(a, tw4, mp);
    RZK_UNROLL
    for (int e = 0; e < kElems; e += 8) a[e] = sreduce<kSignedShift>(a[e], mp);
    g2_stage_s<5, 1>(a, tw4, mp);
    RZK_UNROLL
    for (int e = 0; e < kElems; e += 16) {
        a[e + 1] = sreduce<kSignedShift>(a[e + 1], mp);
        a[e + 2] = sreduce<kSignedShift>(a[e + 2], mp);
        a[e + 3] = sreduce<kSignedShift>(a[e + 3], mp);
    }
}

template <bool BIASED>
RZK_VM void inv_g1_s(uint32_t (&a)[kElems], const uint32_t *g1tab, uint32_t mp)
{
    const uint2 *g1 = reinterpret_cast<const uint2 *>(g1tab);
    g1_stage_s<4, 1>(a, g1, mp);
    g1_stage_s<3, 1>(a, g1, mp);
    g1_stage_s<2, 1>(a, g1, mp);
    RZK_UNROLL
    for (int e = 0; e < kElems; e += 8) a[e] = sreduce<kSignedShift>(a[e], mp);
    g1_stage_s<1, 1>(a, g1, mp);
    g1_stage_s<0, 1, BIASED>(a, g1, mp);    // BIASED: the outputs carry the bias 2^31 (f64_exact_biased)
}

// ---- inverse of a signed slot in decimation-in-time form (RZK_INV_DIT; tables: rzk_tables.cpp) --------------------------
// The forward output is the cyclic DFT of (a_j psi^j) in bit-reversed order, so its inverse is a cyclic decimation-in-time
// transform on that order -- stage m = 2 .. N pairs a[k+j] with a[k+j+m/2] and multiplies the SECOND operand by w_m^j before
// the add / subtract -- followed by the twist a_i *= psi^-i.  That is the forward (Cooley-Tukey) butterfly, four instructions
// and the multiply-first order the pipes like (15.2 against 12.1 butterflies per clock and SM for the Gentleman-Sande form,
// profiles/r2g_bfly_signed.jsonl); the 511 butterflies with j = 0 need no multiplication at all (two instructions), and the
// twist (three instructions per coefficient) hands every output over below 5p/4, with the bias of the conversion for free.
// 620 instructions per lane and transform instead of 748.  Magnitudes in units of p (inputs <= 5/2: two key products, or a
// reduced accumulator): a butterfly without a multiplication doubles, one with a multiplication adds 5/4 to its first operand.
// Contiguous layout: after m = 2 everything is <= 5; after m = 4 the elements e = 0, 2 (mod 4) hold 10, the others 6.25;
// after m = 8 e = 0, 4 (mod 8) hold 20 -- the elements e = 0 (mod 8) are reduced there (4 of 32 per lane) -- so that m = 16
// ends with at most 21.25 (e = 4, 12 mod 16).  The five strided stages multiply every second operand (also lane 0's w = 1)
// and add 5/4 each: below 27.5 < 31.98.
RZK_VM void ct_bfly_s_one(uint32_t &x, uint32_t &y)
{
    const uint32_t s = x + y;
    y = x - y;
    x = s;
}

RZK_VM void inv_g2_dit(uint32_t (&a)[kElems], const uint32_t *g1tab, uint32_t mp)
{
    const uint2 *u = reinterpret_cast<const uint2 *>(g1tab);
    RZK_UNROLL
    for (int e = 0; e < kElems; e += 2) ct_bfly_s_one(a[e], a[e + 1]);
    {
        const uint2 w = u[0];
        RZK_UNROLL
        for (int k = 0; k < kElems; k += 4) {
            ct_bfly_s_one(a[k], a[k + 2]);
            ct_bfly_s(a[k + 1], a[k + 3], w.x, w.y, mp);
        }
    }
    RZK_UNROLL
    for (int j = 0; j < 4; ++j) {
        RZK_UNROLL
        for (int k = 0; k < kElems; k += 8) {
            if (j == 0) ct_bfly_s_one(a[k], a[k + 4]);
            else { const uint2 w = u[j]; ct_bfly_s(a[k + j], a[k + j + 4], w.x, w.y, mp); }
        }
    }
    RZK_UNROLL
    for (int e = 0; e < kElems; e += 8) a[e] = sreduce<kSignedShift>(a[e], mp);
    RZK_UNROLL
    for (int j = 0; j < 8; ++j) {
        RZK_UNROLL
        for (int k = 0; k < kElems; k += 16) {
            if (j == 0) ct_bfly_s_one(a[k], a[k + 8]);
            else { const uint2 w = u[3 + j]; ct_bfly_s(a[k + j], a[k + j + 8], w.x, w.y, mp); }
        }
    }
}

template <bool BIASED>
RZK_VM void inv_g1_dit(uint32_t (&a)[kElems], const uint32_t *g1tab, const uint32_t *g2lane, const uint32_t *twist, int t, uint32_t mp)
{
    const uint2 *u = reinterpret_cast<const uint2 *>(g1tab);
    const uint2 *l = reinterpret_cast<const uint2 *>(g2lane);      // lane t's pairs: stage m = 64 at 0, 128 at 2, 256 at 6, 512 at 14
    {
        const uint2 w = u[16 + t];
        RZK_UNROLL
        for (int k = 0; k < kElems; k += 2) ct_bfly_s(a[k], a[k + 1], w.x, w.y, mp);
    }
    RZK_UNROLL
    for (int c = 0; c < 2; ++c) {
        const uint2 w = l[c];
        RZK_UNROLL
        for (int k = 0; k < kElems; k += 4) ct_bfly_s(a[k + c], a[k + c + 2], w.x, w.y, mp);
    }
    RZK_UNROLL
    for (int c = 0; c < 4; ++c) {
        const uint2 w = l[2 + c];
        RZK_UNROLL
        for (int k = 0; k < kElems; k += 8) ct_bfly_s(a[k + c], a[k + c + 4], w.x, w.y, mp);
    }
    RZK_UNROLL
    for (int c = 0; c < 8; ++c) {
        const uint2 w = l[6 + c];
        RZK_UNROLL
        for (int k = 0; k < kElems; k += 16) ct_bfly_s(a[k + c], a[k + c + 8], w.x, w.y, mp);
    }
    RZK_UNROLL
    for (int c = 0; c < 16; ++c) {
        const uint2 w = l[14 + c];
        ct_bfly_s(a[c], a[c + 16], w.x, w.y, mp);
    }
    const uint4 *tw = reinterpret_cast<const uint4 *>(twist) + t;       // uint4 k of lane t: the pairs of its registers 2k and 2k + 1
    RZK_UNROLL
    for (int k = 0; k < kElems / 2; ++k) {
        const uint4 w = tw[kLanes * k];
        // BIASED: the outputs carry the bias 2^31 (f64_exact_biased)
        a[2 * k] = sshoup_mac(w.x, w.y, a[2 * k], mp, BIASED ? 0x80000000u : 0u);
        a[2 * k + 1] = sshoup_mac(w.z, w.w, a[2 * k + 1], mp, BIASED ? 0x80000000u : 0u);
    }
}

// ---------------------------------------------------------------- global memory

RZK_VM uint64_t stream_poly(const Stream &s, uint32_t item, uint32_t off)
{
    const uint32_t grp = (uint32_t)(((uint64_t)item * s.magic) >> s.shift);      // item / s.div (rzk_vm.h set_stream_div)
    return (uint64_t)grp * s.stride + off;
}

// ---------------------------------------------------------------- ops

RZK_VM uint4 rot_ld128(const void *p)
{
#if defined(__CUDA_ARCH__)
    return __ldg(reinterpret_cast<const uint4 *>(p));
#else
    return *reinterpret_cast<const uint4 *>(p);
#endif
}


template <bool SGN = false, bool TINY = false>
RZK_VM void op_fwd(const VmLaunch &K, Lane *lanes, const LaneCtx *ctxs, const Op &op, int it, uint32_t dtype)
{
    const Stream st = K.st[op.a];
    if (K.ld128 && dtype != DT_I8) {
        // A/B variant (RZK_TUNE=ld128=1): the row arrives as 128-bit loads -- lane t fetches the 32 contiguous coefficients
        // [32 t, 32 t + 32) -- and is redistributed to the strided layout through the transpose buffer (8 LDG.128 +
        // 8 STS.128 + 32 LDS instead of 32 LDG.32).  Measured 9-16 % slower on every kernel: the loads
        // are not what limits these kernels, and the extra shared-memory round trip costs issue slots and LSU wavefronts.
        RZK_EACH_LANE {
            RZK_LANE;
            const uint64_t poly = stream_poly(st, ctx.item, (uint32_t)op.off + (uint32_t)it * op.step +
                                                                ((op.b & FWD_HWPOLY) ? (uint32_t)ctx.hw : 0u));
            const uint4 *src4 = reinterpret_cast<const uint4 *>(reinterpret_cast<const int32_t *>(st.base) + poly * kN) + 8 * t;
            uint4 *row = reinterpret_cast<uint4 *>(ctx.buf + 36 * t);
            RZK_UNROLL
            for (int j = 0; j < 8; ++j) row[j] = rot_ld128(src4 + j);
        }
        RZK_SYNC();
    }
    RZK_EACH_LANE {
        RZK_LANE;
        const PrimeC &pc = L.pc;
        const uint64_t poly = stream_poly(st, ctx.item, (uint32_t)op.off + (uint32_t)it * op.step +
                                                            ((op.b & FWD_HWPOLY) ? (uint32_t)ctx.hw : 0u));
        int32_t v[kElems];
        if (dtype == DT_I8) {
            const int8_t *src = reinterpret_cast<const int8_t *>(st.base) + poly * kN;
            RZK_UNROLL
            for (int m = 0; m < kElems; ++m) v[m] = src[t + kLanes * m];
        } else if (K.ld128) {
            // staged by the 128-bit loads above: every lane reads exactly the words its own transpose will overwrite
            RZK_UNROLL
            for (int m = 0; m < kElems; ++m) {
                const int i = t + kLanes * m;
                v[m] = lift_in((int32_t)ctx.buf[i + ((i >> 5) << 2)], K.q);
            }
        } else {
            const int32_t *src = reinterpret_cast<const int32_t *>(st.base) + poly * kN;
            RZK_UNROLL
            for (int m = 0; m < kElems; ++m) v[m] = SGN ? src[t + kLanes * m] : lift_in(src[t + kLanes * m], K.q);
        }
        if (op.b & FWD_CHECK_SMALL) {
            // |v| <= lim  <=>  (uint32)(v + lim) <= 2 lim: one add-and-max per coefficient (VIADDMNMX)
            uint32_t mx = 0;
            RZK_UNROLL
            for (int m = 0; m < kElems; ++m) mx = umax32(mx, (uint32_t)v[m] + K.small_lim);
            L.rerr |= (mx > 2u * K.small_lim) ? 1u : 0u;
#if defined(__CUDA_ARCH__)
            // (materialise the verdict here: left alone the compiler sinks the whole check to the end of the item and keeps --
            // spills -- the 32 inputs across every transform)
            asm volatile("" : "+r"(L.rerr));
#endif
        }
        if constexpr (SGN) {
            // signed lazy form: an int8 operand is its own input; an int32 operand (ANY representative, |v| <= 2^31) is brought
            // below 5p/4 -- by the Shoup product with R N^-1 that a Montgomery operand needs anyway, or by one shift-reduce
            const uint32_t mp = 0u - pc.p;
            if (dtype == DT_I8) {
                RZK_UNROLL
                for (int m = 0; m < kElems; ++m) L.cur[m] = (uint32_t)v[m];
            } else if ((op.b & FWD_SCALED) && !RZK_INV_DIT) {
                // (with the decimation-in-time inverse the factor R N^-1 of a Montgomery operand rides on the output twist instead:
                // PrimeTables::twist_rn, selected for MODE_SEQ_S at launch -- two multiplies less per coefficient)
                RZK_UNROLL
                for (int m = 0; m < kElems; ++m) L.cur[m] = sshoup_mac(pc.rn, pc.rnp, (uint32_t)v[m], mp, 0u);
            } else {
                RZK_UNROLL
                for (int m = 0; m < kElems; ++m) L.cur[m] = sreduce<kSignedShift>((uint32_t)v[m], mp);
            }
        } else if (op.b & FWD_SCALED) {
            // centred value + 2p lies in (0, 4p): a valid lazy input of the forward butterflies
            RZK_UNROLL
            for (int m = 0; m < kElems; ++m)
                L.cur[m] = shoup_mul(pc.rn, pc.rnp, (uint32_t)v[m] + pc.p2, pc.p);
        } else {
            RZK_UNROLL
            for (int m = 0; m < kElems; ++m) L.cur[m] = (uint32_t)v[m] + pc.p2;
        }
#if defined(__CUDA_ARCH__)
        pp_acquire(K);
#endif
        if constexpr (SGN) fwd_g1_s<TINY>(L.cur, ctx.g1 + (L.pi * 2 + 0) * kG1Words, 0u - pc.p);
        else fwd_g1(L.cur, ctx.g1 + (L.pi * 2 + 0) * kG1Words, pc.p, pc.p2, pc.pad_, L.cap);
        RZK_UNROLL
        for (int m = 0; m < kElems; ++m) {
            const int i = t + kLanes * m;
            ctx.buf[i + ((i >> 5) << 2)] = L.cur[m];
        }
    }
    RZK_SYNC();
    RZK_EACH_LANE {
        RZK_LANE;
        const uint4 *row = reinterpret_cast<const uint4 *>(ctx.buf + 36 * t);
        RZK_UNROLL
        for (int j = 0; j < 8; ++j) {
            const uint4 q = row[j];
            L.cur[4 * j + 0] = q.x; L.cur[4 * j + 1] = q.y; L.cur[4 * j + 2] = q.z; L.cur[4 * j + 3] = q.w;
        }
        if constexpr (SGN) fwd_g2_s(L.cur, ctx.g2 + ((L.pi * 2 + 0) * kLanes + t) * kG2Words, 0u - L.pc.p);
        else fwd_g2(L.cur, ctx.g2 + ((L.pi * 2 + 0) * kLanes + t) * kG2Words, L.pc.p, L.pc.p2, L.pc.pad_, L.cap);
    }
    RZK_SYNC();
#if defined(__CUDA_ARCH__)
    pp_release(K, *ctxs[0].pp_count);
#endif
}

// acc (+)= key (.) cur ; key rows pre-scaled by N^-1, Shoup form
RZK_VM void mac_key(uint32_t (&acc)[kElems], const uint32_t (&cur)[kElems], const uint32_t *krow, int t,
                    uint32_t flags, uint32_t p, uint32_t p2)
{
    const uint4 *w4 = reinterpret_cast<const uint4 *>(krow + 36 * t);
    const uint4 *wp4 = reinterpret_cast<const uint4 *>(krow + kPadWords + 36 * t);
    const bool init = flags & MAC_INIT, neg = flags & MAC_NEG;
    RZK_UNROLL
    for (int j = 0; j < 8; ++j) {
        const uint4 w = w4[j], wp = wp4[j];
        const uint32_t ww[4] = {w.x, w.y, w.z, w.w}, wwp[4] = {wp.x, wp.y, wp.z, wp.w};
        RZK_UNROLL
        for (int c = 0; c < 4; ++c) {
            const int e = 4 * j + c;
            uint32_t tt = shoup_mul(ww[c], wwp[c], cur[e], p);
            if (neg) tt = p2 - tt;
            // (a first, positive term is already in [0, 2p): no correction)
            acc[e] = (init && !neg) ? tt : csub((init ? 0u : acc[e]) + tt, p2);
        }
    }
}

// signed lazy form (MODE_SPLITKEY_S): three instructions per term, the accumulation rides on the first multiply-add;
// |acc| grows by 5p/4 per term.  Key rows hold centred values with signed Shoup companions.
RZK_VM void mac_key_s(uint32_t (&acc)[kElems], const uint32_t (&cur)[kElems], const uint32_t *krow, int t, uint32_t flags, uint32_t mp)
{
    const uint4 *w4 = reinterpret_cast<const uint4 *>(krow + 36 * t);
    const uint4 *wp4 = reinterpret_cast<const uint4 *>(krow + kPadWords + 36 * t);
    const bool init = flags & MAC_INIT;
    RZK_UNROLL
    for (int j = 0; j < 8; ++j) {
        const uint4 w = w4[j], wp = wp4[j];
        const uint32_t ww[4] = {w.x, w.y, w.z, w.w}, wwp[4] = {wp.x, wp.y, wp.z, wp.w};
        RZK_UNROLL
        for (int c = 0; c < 4; ++c) {
            const int e = 4 * j + c;
            acc[e] = sshoup_mac(ww[c], wwp[c], cur[e], mp, init ? 0u : acc[e]);
        }
    }
}

RZK_VM void mac_key_smem_s(uint32_t *acc1, const uint32_t (&cur)[kElems], const uint32_t *krow, int t, uint32_t flags, uint32_t mp)
{
    const uint4 *w4 = reinterpret_cast<const uint4 *>(krow + 36 * t);
    const uint4 *wp4 = reinterpret_cast<const uint4 *>(krow + kPadWords + 36 * t);
    uint4 *a4 = reinterpret_cast<uint4 *>(acc1);
    const bool init = flags & MAC_INIT;
    RZK_UNROLL
    for (int j = 0; j < 8; ++j) {
        const uint4 w = w4[j], wp = wp4[j];
        uint4 a = init ? uint4{0u, 0u, 0u, 0u} : a4[j * kLanes + t];
        a.x = sshoup_mac(w.x, wp.x, cur[4 * j + 0], mp, a.x);
        a.y = sshoup_mac(w.y, wp.y, cur[4 * j + 1], mp, a.y);
        a.z = sshoup_mac(w.z, wp.z, cur[4 * j + 2], mp, a.z);
        a.w = sshoup_mac(w.w, wp.w, cur[4 * j + 3], mp, a.w);
        a4[j * kLanes + t] = a;
    }
}

// acc (+)= slot (.) cur ; slot holds R*N^-1-scaled residues in [0,p), Montgomery product
RZK_VM void mac_var(uint32_t (&acc)[kElems], const uint32_t (&cur)[kElems], const uint32_t *slot, int t,
                    uint32_t flags, uint32_t p, uint32_t p2, uint32_t pinv)
{
    const uint4 *s4 = reinterpret_cast<const uint4 *>(slot);
    const bool init = flags & MAC_INIT, neg = flags & MAC_NEG;
    RZK_UNROLL
    for (int j = 0; j < 8; ++j) {
        const uint4 s = s4[j * kLanes + t];
        const uint32_t ss[4] = {s.x, s.y, s.z, s.w};
        RZK_UNROLL
        for (int c = 0; c < 4; ++c) {
            const int e = 4 * j + c;
            uint32_t tt = mont_mul(csub(cur[e], p2), ss[c], p, pinv);
            if (neg) tt = p2 - tt;
            acc[e] = (init && !neg) ? tt : csub((init ? 0u : acc[e]) + tt, p2);
        }
    }
}

// signed lazy form (MODE_SEQ_S): acc +- slot (.) cur as a signed Montgomery product, five instructions per term.
// |cur| <= 12.5 p, |slot| <= 1.06 p: the product a b 2^-32 (mod p) comes out below 0.71 p in magnitude, so an accumulator
// may take 32 terms between two reductions (sp_exec / vm_run_item reduce it every 32nd iteration of a loop).
RZK_VM void mac_var_s(uint32_t (&acc)[kElems], const uint32_t (&cur)[kElems], const uint32_t *slot, int t,
                      uint32_t flags, uint32_t p, uint32_t pinv)
{
    const uint4 *s4 = reinterpret_cast<const uint4 *>(slot);
    const bool init = flags & MAC_INIT, neg = flags & MAC_NEG;
    RZK_UNROLL
    for (int j = 0; j < 8; ++j) {
        const uint4 s = s4[j * kLanes + t];
        const uint32_t ss[4] = {s.x, s.y, s.z, s.w};
        RZK_UNROLL
        for (int c = 0; c < 4; ++c) {
            const int e = 4 * j + c;
            const uint32_t tt = smont_mul(cur[e], ss[c], p, pinv);
            const uint32_t a0 = init ? 0u : acc[e];
            acc[e] = neg ? a0 - tt : a0 + tt;
        }
    }
}

RZK_VM void mac_var_smem_s(uint32_t *acc1, const uint32_t (&cur)[kElems], const uint32_t *slot, int t,
                           uint32_t flags, uint32_t p, uint32_t pinv)
{
    const uint4 *s4 = reinterpret_cast<const uint4 *>(slot);
    uint4 *a4 = reinterpret_cast<uint4 *>(acc1);
    const bool init = flags & MAC_INIT, neg = flags & MAC_NEG;
    RZK_UNROLL
    for (int j = 0; j < 8; ++j) {
        const uint4 s = s4[j * kLanes + t];
        uint4 a = init ? uint4{0u, 0u, 0u, 0u} : a4[j * kLanes + t];
        const uint32_t t0 = smont_mul(cur[4 * j + 0], s.x, p, pinv), t1 = smont_mul(cur[4 * j + 1], s.y, p, pinv);
        const uint32_t t2 = smont_mul(cur[4 * j + 2], s.z, p, pinv), t3 = smont_mul(cur[4 * j + 3], s.w, p, pinv);
        a.x = neg ? a.x - t0 : a.x + t0; a.y = neg ? a.y - t1 : a.y + t1;
        a.z = neg ? a.z - t2 : a.z + t2; a.w = neg ? a.w - t3 : a.w + t3;
        a4[j * kLanes + t] = a;
    }
}

RZK_VM bool ops_use_acc1(const Op *ops)
{
    bool any = false;
    RZK_NOUNROLL
    for (int i = 0; i < kMaxOps && ops[i].code != OP_END; ++i)
        any = any || ((ops[i].code == OP_MACV || ops[i].code == OP_INV) && ops[i].a == 1);
    return any;
}

// every 32nd term of a looped product sum: the accumulators back below 1.06 p
RZK_VM void reduce_accumulators_s(Lane *lanes, const LaneCtx *ctxs, bool with_acc1)
{
    RZK_EACH_LANE {
        RZK_LANE;
        const uint32_t mp = 0u - L.pc.p;
        RZK_UNROLL
        for (int e = 0; e < kElems; ++e) L.acc0[e] = sreduce<kSignedShift>(L.acc0[e], mp);
        if (with_acc1) {
            uint4 *a4 = reinterpret_cast<uint4 *>(ctx.acc1);
            RZK_UNROLL
            for (int j = 0; j < 8; ++j) {
                uint4 a = a4[j * kLanes + t];
                a.x = sreduce<kSignedShift>(a.x, mp); a.y = sreduce<kSignedShift>(a.y, mp);
                a.z = sreduce<kSignedShift>(a.z, mp); a.w = sreduce<kSignedShift>(a.w, mp);
                a4[j * kLanes + t] = a;
            }
        }
    }
}

// accumulator-1 variants: the accumulator is the lane-private shared-memory slot
RZK_VM void mac_key_smem(uint32_t *acc1, const uint32_t (&cur)[kElems], const uint32_t *krow, int t,
                         uint32_t flags, uint32_t p, uint32_t p2)
{
    const uint4 *w4 = reinterpret_cast<const uint4 *>(krow + 36 * t);
    const uint4 *wp4 = reinterpret_cast<const uint4 *>(krow + kPadWords + 36 * t);
    uint4 *a4 = reinterpret_cast<uint4 *>(acc1);
    const bool init = flags & MAC_INIT, neg = flags & MAC_NEG;
    RZK_UNROLL
    for (int j = 0; j < 8; ++j) {
        const uint4 w = w4[j], wp = wp4[j];
        uint4 a = a4[j * kLanes + t];
        uint32_t tt;
        tt = shoup_mul(w.x, wp.x, cur[4 * j + 0], p); if (neg) tt = p2 - tt; a.x = (init && !neg) ? tt : csub((init ? 0u : a.x) + tt, p2);
        tt = shoup_mul(w.y, wp.y, cur[4 * j + 1], p); if (neg) tt = p2 - tt; a.y = (init && !neg) ? tt : csub((init ? 0u : a.y) + tt, p2);
        tt = shoup_mul(w.z, wp.z, cur[4 * j + 2], p); if (neg) tt = p2 - tt; a.z = (init && !neg) ? tt : csub((init ? 0u : a.z) + tt, p2);
        tt = shoup_mul(w.w, wp.w, cur[4 * j + 3], p); if (neg) tt = p2 - tt; a.w = (init && !neg) ? tt : csub((init ? 0u : a.w) + tt, p2);
        a4[j * kLanes + t] = a;
    }
}

RZK_VM void mac_var_smem(uint32_t *acc1, const uint32_t (&cur)[kElems], const uint32_t *slot, int t,
                         uint32_t flags, uint32_t p, uint32_t p2, uint32_t pinv)
{
    const uint4 *s4 = reinterpret_cast<const uint4 *>(slot);
    uint4 *a4 = reinterpret_cast<uint4 *>(acc1);
    const bool init = flags & MAC_INIT, neg = flags & MAC_NEG;
    RZK_UNROLL
    for (int j = 0; j < 8; ++j) {
        const uint4 s = s4[j * kLanes + t];
        uint4 a = a4[j * kLanes + t];
        uint32_t tt;
        tt = mont_mul(csub(cur[4 * j + 0], p2), s.x, p, pinv); if (neg) tt = p2 - tt; a.x = (init && !neg) ? tt : csub((init ? 0u : a.x) + tt, p2);
        tt = mont_mul(csub(cur[4 * j + 1], p2), s.y, p, pinv); if (neg) tt = p2 - tt; a.y = (init && !neg) ? tt : csub((init ? 0u : a.y) + tt, p2);
        tt = mont_mul(csub(cur[4 * j + 2], p2), s.z, p, pinv); if (neg) tt = p2 - tt; a.z = (init && !neg) ? tt : csub((init ? 0u : a.z) + tt, p2);
        tt = mont_mul(csub(cur[4 * j + 3], p2), s.w, p, pinv); if (neg) tt = p2 - tt; a.w = (init && !neg) ? tt : csub((init ? 0u : a.w) + tt, p2);
        a4[j * kLanes + t] = a;
    }
}

template <bool SGN = false>
RZK_VM void op_st(Lane *lanes, const LaneCtx *ctxs, const Op &op)
{
    RZK_EACH_LANE {
        RZK_LANE;
        const PrimeC &pc = L.pc;
        uint4 *s4 = reinterpret_cast<uint4 *>(ctx.slot);
        if (SGN && !(op.b & ST_RAW)) {      // signed Montgomery operand: a representative below 1.06 p (one shift-reduce)
            const uint32_t mp = 0u - pc.p;
            RZK_UNROLL
            for (int j = 0; j < 8; ++j) {
                uint4 q;
                q.x = sreduce<kSignedShift>(L.cur[4 * j + 0], mp); q.y = sreduce<kSignedShift>(L.cur[4 * j + 1], mp);
                q.z = sreduce<kSignedShift>(L.cur[4 * j + 2], mp); q.w = sreduce<kSignedShift>(L.cur[4 * j + 3], mp);
                s4[j * kLanes + t] = q;
            }
        } else if (op.b & ST_RAW) {            // operand of a Shoup product with the key: any 32-bit value is valid
            RZK_UNROLL
            for (int j = 0; j < 8; ++j) {
                uint4 q;
                q.x = L.cur[4 * j + 0]; q.y = L.cur[4 * j + 1]; q.z = L.cur[4 * j + 2]; q.w = L.cur[4 * j + 3];
                s4[j * kLanes + t] = q;
            }
        } else {                        // operand of a Montgomery product: fully reduced
            RZK_UNROLL
            for (int j = 0; j < 8; ++j) {
                uint4 q;
                q.x = csub(csub(L.cur[4 * j + 0], pc.p2), pc.p);
                q.y = csub(csub(L.cur[4 * j + 1], pc.p2), pc.p);
                q.z = csub(csub(L.cur[4 * j + 2], pc.p2), pc.p);
                q.w = csub(csub(L.cur[4 * j + 3], pc.p2), pc.p);
                s4[j * kLanes + t] = q;
            }
        }
    }
    RZK_SYNC();     // OP_LD may read the other half warp's slot
}

// cur = operand slot of half warp op.a (same lane, same positions)
RZK_VM void op_ld(Lane *lanes, const LaneCtx *ctxs, const Op &op)
{
    RZK_EACH_LANE {
        RZK_LANE;
        const uint4 *s4 = reinterpret_cast<const uint4 *>((op.a & 1) ? ctx.slot_hw[1] : ctx.slot_hw[0]);
        RZK_UNROLL
        for (int j = 0; j < 8; ++j) {
            const uint4 q = s4[j * kLanes + t];
            L.cur[4 * j + 0] = q.x; L.cur[4 * j + 1] = q.y; L.cur[4 * j + 2] = q.z; L.cur[4 * j + 3] = q.w;
        }
    }
}

// NTT images in global memory (OP_STG / OP_MACG): one image = np rows of 512 words in the lane-private order of the
// operand slot; a transform that several items share (the challenge d of the T terms of a Sum proof, of both first
// equations of a Linear proof) is computed once per group by one launch and read by the items of the next.
RZK_VM uint32_t *image_row(const VmLaunch &K, const Stream &st, const LaneCtx &ctx, const Lane &L)
{
    return reinterpret_cast<uint32_t *>(const_cast<void *>(st.base)) + stream_poly(st, ctx.item, (uint32_t)L.pi) * kN;
}

RZK_VM void op_stg(const VmLaunch &K, Lane *lanes, const LaneCtx *ctxs, const Op &op)
{
    const Stream st = K.st[op.a];
    RZK_EACH_LANE {
        RZK_LANE;
        const PrimeC &pc = L.pc;
        if (ctx.active) {
            uint4 *s4 = reinterpret_cast<uint4 *>(image_row(K, st, ctx, L));
            RZK_UNROLL
            for (int j = 0; j < 8; ++j) {
                uint4 q;
                q.x = csub(csub(L.cur[4 * j + 0], pc.p2), pc.p);
                q.y = csub(csub(L.cur[4 * j + 1], pc.p2), pc.p);
                q.z = csub(csub(L.cur[4 * j + 2], pc.p2), pc.p);
                q.w = csub(csub(L.cur[4 * j + 3], pc.p2), pc.p);
                s4[j * kLanes + t] = q;
            }
        }
    }
}

// ---------------------------------------------------------------- epilogue (last prime only)

// Number of coefficients a lane finishes: in the warp-per-item modes two half warps share an item.
template <int MODE>
struct Epi {
    static constexpr int kCount = (!mode_seq(MODE)) ? 16 : 32;
    // Warp-per-item modes: the epilogue value is an exact integer in a double and the reduction mod q runs on the FP64
    // pipe (three FMAs), which relieves the FMA-heavy pipe of the wide multiplies and the mulhi per coefficient.
    //   MODE_SPLITKEY: |lo + 2^16 hi + plain terms| < 2^47.
    //   MODE_SPLIT:    the Garner digit h1 (centred) enters as (p0 * h1) mod q through an exact FMA product
    //                  (crt2_mod_q_f64), so the value is a0 + r + plain terms, < 2^35.
    // MODE_SEQ (1 or 3 primes) stays in int64
    typedef typename std::conditional<!mode_seq(MODE), double, int64_t>::type V_t;
};

// coefficient index m (in the G1 layout, i = t + 16*m) of epilogue element j
template <int MODE>
// Warp-per-item modes: half warp h finishes the rows m = 2j + h, so the two half warps of a warp touch ADJACENT 64-byte
// segments of every row they read or write in the epilogue (coefficients t + 32j and t + 16 + 32j: one 128-byte line per
// warp instruction) and 32 consecutive words of the rotation sum's extended row (OP_ROT: one conflict-free wavefront).
RZK_VM int epi_m(const LaneCtx &ctx, int j) { return (!mode_seq(MODE)) ? (2 * j + ctx.hw) : j; }

// OP_ADDP in two halves: the loads (issued BEFORE the inverse transform by the compile-time programs, so that their latency
// -- an L2 hit, ~300 cycles -- is covered by the butterflies instead of stalling the epilogue) and the accumulation.
// one int8 with a load the compilers keep where it is written (relaxed, gpu scope): a plain load of read-only data is hoisted
// to the top of the item loop and its 16 results are then spilled across the transforms
RZK_VM int32_t ld_i8_pinned(const int8_t *p)
{
#if defined(__CUDA_ARCH__)
    int32_t v;
    asm volatile("ld.relaxed.gpu.global.s8 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
#else
    return *p;
#endif
}

template <int MODE, bool PINNED = false>
RZK_VM void op_addp_load(const VmLaunch &K, const LaneCtx *ctxs, int32_t (&v)[RZK_NL][Epi<MODE>::kCount], const Op &op, int it, uint32_t dtype)
{
    constexpr int CNT = Epi<MODE>::kCount;
    const Stream st = K.st[op.a];
    RZK_EACH_LANE {
        const LaneCtx &ctx = ctxs[li_];
        const int t = ctx.t;
        const uint64_t poly = stream_poly(st, ctx.item, (uint32_t)op.off + (uint32_t)it * op.step);
        if (dtype == DT_I8) {
            const int8_t *src = reinterpret_cast<const int8_t *>(st.base) + poly * kN;
            RZK_UNROLL
            for (int j = 0; j < CNT; ++j) v[li_][j] = PINNED ? ld_i8_pinned(src + t + kLanes * epi_m<MODE>(ctx, j)) : (int32_t)src[t + kLanes * epi_m<MODE>(ctx, j)];
        } else {
            const int32_t *src = reinterpret_cast<const int32_t *>(st.base) + poly * kN;
            RZK_UNROLL
            for (int j = 0; j < CNT; ++j) v[li_][j] = src[t + kLanes * epi_m<MODE>(ctx, j)];   // any representative: reduced in OP_FIN
        }
    }
}

template <int MODE>
RZK_VM void op_addp_apply(typename Epi<MODE>::V_t (&V)[RZK_NL][Epi<MODE>::kCount], const int32_t (&v)[RZK_NL][Epi<MODE>::kCount], const Op &op)
{
    constexpr int CNT = Epi<MODE>::kCount;
    const bool neg = op.c & MAC_NEG;
    RZK_EACH_LANE {
        if constexpr (!mode_seq(MODE)) {
            RZK_UNROLL
            for (int j = 0; j < CNT; ++j) V[li_][j] += neg ? -f64_exact_i32(v[li_][j]) : f64_exact_i32(v[li_][j]);
        } else {
            RZK_UNROLL
            for (int j = 0; j < CNT; ++j) V[li_][j] += neg ? -(int64_t)v[li_][j] : (int64_t)v[li_][j];
        }
    }
}

template <int MODE>
RZK_VM void op_addp(const VmLaunch &K, const LaneCtx *ctxs, typename Epi<MODE>::V_t (&V)[RZK_NL][Epi<MODE>::kCount], const Op &op, int it, uint32_t dtype)
{
    int32_t v[RZK_NL][Epi<MODE>::kCount];
    op_addp_load<MODE>(K, ctxs, v, op, it, dtype);
    op_addp_apply<MODE>(V, v, op);
}

template <int MODE>
RZK_VM void op_fin(const VmLaunch &K, Lane *lanes, const LaneCtx *ctxs, typename Epi<MODE>::V_t (&V)[RZK_NL][Epi<MODE>::kCount], const Op &op, int it)
{
    constexpr int CNT = Epi<MODE>::kCount;
    const Stream st = K.st[op.a];
    RZK_EACH_LANE {
        RZK_LANE;
        const uint64_t poly = stream_poly(st, ctx.item, (uint32_t)op.off + (uint32_t)it * op.step);
        int32_t res[CNT];
        RZK_UNROLL
        for (int j = 0; j < CNT; ++j) {
            if constexpr (!mode_seq(MODE)) res[j] = reduce_q_centered_f64(V[li_][j], K.qd, K.qinvd);
            else res[j] = reduce_q_centered(V[li_][j], K.q, K.m30, K.kqh);
        }
        if (op.b & FIN_CMPZ) {
            uint32_t nz = 0;
            RZK_UNROLL
            for (int j = 0; j < CNT; ++j) nz |= (uint32_t)res[j];
            L.fail |= nz ? 1u : 0u;
        }
        if ((op.b & FIN_STORE) && ctx.active) {
            int32_t *dst = reinterpret_cast<int32_t *>(const_cast<void *>(st.base)) + poly * kN;
            RZK_UNROLL
            for (int j = 0; j < CNT; ++j) dst[t + kLanes * epi_m<MODE>(ctx, j)] = res[j];
        }
    }
}

// OP_ROT -- SURVEY kernel K4 for full-size rows: V += +-(c * d) where d is a challenge (kappa entries +-1,
// challenge_space.rs:12-33) and c an int32 row (a commitment row), as in `c1.componentwise_mul(&d)` of the verification
// equations (open.rs:172, linear.rs:226,232, sum.rs:287,295).  c * d = sum_k d[pos_k] * X^pos_k * c, and X^pos * c is c
// rotated by pos with the wrapped part negated, so the product needs no multiplication at all:
//   * the warp writes the extended row E = [B - c | B + c | B - c] (B = 2^31) to shared memory as 1536 biased uint32 words:
//     the 512 words from offset 512 - pos are B + (X^pos c), the 512 words from offset 1024 - pos are B - (X^pos c)
//     (c is canonicalised first, |c| < 2^31);
//   * it lists one row offset per non-zero entry of d, in pairs: the first term of a pair will be ADDED, so it lists the
//     row holding B + (its contribution); the second will be SUBTRACTED, so it lists the row holding B - (its contribution).
//     Any two terms can be partners whatever their signs; an odd count is completed by a row of biased zeros;
//   * every lane adds its 16 epilogue coefficients (rows m = 2j + hw: the warp reads 32 consecutive words, one conflict-free
//     wavefront) of every term with ONE 32-bit shared-memory load and ONE DADD, on the FP64 pipe the integer transforms
//     leave idle: the loaded word u becomes the low half of the double D = 2^52 + u (high word 0x43300000, no conversion
//     instruction; the high words sit in registers that are written once, before the loop), and V = (V + D1) - D2 cancels
//     2^52 and B at once, so every intermediate is an exact integer below 2^53 (|V| < 2^39 between pairs);
// instead of two forward transforms and a pointwise product per prime.  An item whose d has an entry outside {-1, 0, 1}
// takes a general loop (one scaled rotation per non-zero entry, DFMA), so the op is exact for every int8 d and every
// int32 representative of c.  Warp-per-item modes only.
RZK_VM double rot_biased(uint32_t u)      // the exact double 2^52 + u
{
#if defined(__CUDA_ARCH__)
    return __hiloint2double(0x43300000, (int)u);
#else
    return 4503599627370496.0 + (double)u;
#endif
}

RZK_VM double rot_biased_hi(uint32_t hi, uint32_t u)      // the same with the high word 0x43300000 supplied in a register
{
#if defined(__CUDA_ARCH__)
    return __hiloint2double((int)hi, (int)u);
#else
    (void)hi;
    return 4503599627370496.0 + (double)u;
#endif
}

RZK_VM void rot_ld_pair(const int32_t *p, int32_t &a, int32_t &b)      // 8-byte aligned pair
{
#if defined(__CUDA_ARCH__)
    const int2 v = __ldg(reinterpret_cast<const int2 *>(p));
    a = v.x; b = v.y;
#else
    a = p[0]; b = p[1];
#endif
}

RZK_VM void rot_st_pair(uint32_t *p, uint32_t a, uint32_t b)           // 8-byte aligned pair
{
#if defined(__CUDA_ARCH__)
    *reinterpret_cast<uint2 *>(p) = make_uint2(a, b);
#else
    p[0] = a; p[1] = b;
#endif
}

RZK_VM uint32_t rot_popc(uint32_t v)
{
#if defined(__CUDA_ARCH__)
    return (uint32_t)__popc(v);
#else
    return (uint32_t)__builtin_popcount(v);
#endif
}

RZK_VM uint32_t rot_ctz(uint32_t v)       // v != 0
{
#if defined(__CUDA_ARCH__)
    return (uint32_t)(__ffs((int)v) - 1);
#else
    return (uint32_t)__builtin_ctz(v);
#endif
}

template <int MODE>
RZK_VM void op_rot(const VmLaunch &K, Lane *lanes, const LaneCtx *ctxs, typename Epi<MODE>::V_t (&V)[RZK_NL][Epi<MODE>::kCount], const Op &op, int it)
{
    if constexpr (!mode_seq(MODE)) {
        constexpr int CNT = Epi<MODE>::kCount;
        const Stream sc = K.st[op.a], sd = K.st[op.b];
        const bool neg = op.c & MAC_NEG;          // V -= c*d: every term's sign flips
        RZK_SYNC();          // the transpose buffers this overlays are no longer read
        // ---- E = [B - c | B + c | B - c]: lane l handles the coefficient pairs 2l + 64e, e = 0..7 (8-byte loads and stores)
        RZK_EACH_LANE {
            const LaneCtx &ctx = ctxs[li_];
            uint32_t *E = ctx.red;
            const uint64_t poly = stream_poly(sc, ctx.item, (uint32_t)op.off + (uint32_t)it * op.step);
            const int32_t *src = reinterpret_cast<const int32_t *>(sc.base) + poly * kN + 2 * ctx.ridx;
            RZK_UNROLL
            for (int e = 0; e < 8; ++e) {
                int32_t v0, v1;
                rot_ld_pair(src + 64 * e, v0, v1);
                // any int32 representative works (the sum is reduced mod q at the end); only the lowest values move up by q so
                // that 2^31 - c fits a word (lift_in: one add-and-max)
                const uint32_t c0 = (uint32_t)lift_in(v0, K.q), c1 = (uint32_t)lift_in(v1, K.q);
                rot_st_pair(E + 512 + 2 * ctx.ridx + 64 * e, 0x80000000u + c0, 0x80000000u + c1);
                rot_st_pair(E + 2 * ctx.ridx + 64 * e, 0x80000000u - c0, 0x80000000u - c1);
                rot_st_pair(E + 1024 + 2 * ctx.ridx + 64 * e, 0x80000000u - c0, 0x80000000u - c1);
            }
        }
        // ---- the non-zero entries of d: lane l scans the 16 bytes d[16l .. 16l+16) with word-parallel byte tests
        uint32_t dw[RZK_NL][4], nzw[RZK_NL][4], cnt_l[RZK_NL], pre_l[RZK_NL], bad_l[RZK_NL];
        RZK_EACH_LANE {
            const LaneCtx &ctx = ctxs[li_];
            const uint64_t poly = stream_poly(sd, ctx.item, 0u);
            const uint4 q = rot_ld128(reinterpret_cast<const int8_t *>(sd.base) + poly * kN + 16 * ctx.ridx);
            dw[li_][0] = q.x; dw[li_][1] = q.y; dw[li_][2] = q.z; dw[li_][3] = q.w;
            uint32_t cnt = 0, bad = 0;
            RZK_UNROLL
            for (int c = 0; c < 4; ++c) {
                const uint32_t w = dw[li_][c];
                const uint32_t nz = (((w & 0x7f7f7f7fu) + 0x7f7f7f7fu) | w) & 0x80808080u;            // bit 7: byte != 0
                const uint32_t t1 = ((w & 0x7f7f7f7fu) + 0x01010101u) ^ ((w ^ 0x01010101u) & 0x80808080u);   // byte + 1 (mod 256): 0, 1, 2
                bad |= (((t1 & 0x7f7f7f7fu) + 0x7d7d7d7du) | t1) & 0x80808080u;                         // bit 7: byte + 1 > 2
                nzw[li_][c] = nz;
                cnt += rot_popc(nz);
            }
            cnt_l[li_] = cnt; bad_l[li_] = bad;
        }
        uint32_t n_terms = 0, bad_any = 0;
#if defined(__CUDA_ARCH__)
        {
            uint32_t v = cnt_l[0];          // inclusive scan over the warp
            RZK_UNROLL
            for (int s = 1; s < 32; s <<= 1) { const uint32_t o = __shfl_up_sync(0xffffffffu, v, s); if (ctxs[0].ridx >= s) v += o; }
            pre_l[0] = v - cnt_l[0];
            n_terms = __shfl_sync(0xffffffffu, v, 31);
            bad_any = __any_sync(0xffffffffu, bad_l[0] != 0) ? 1u : 0u;
        }
#else
        for (int li = 0; li < RZK_NL; ++li) { pre_l[li] = n_terms; n_terms += cnt_l[li]; bad_any |= bad_l[li] ? 1u : 0u; }
#endif
        const uint32_t n_pairs = (n_terms + 1u) >> 1;
        // ---- the list: slot k holds the row offset of term k -- an even slot is added (row B + contribution), an odd slot is
        //      subtracted (row B - contribution); the contribution of an entry s at position pos is +-s X^pos c
        RZK_EACH_LANE {
            const LaneCtx &ctx = ctxs[li_];
            uint16_t *list = reinterpret_cast<uint16_t *>(ctx.red + kRotListOff);
            uint32_t slot = pre_l[li_];
            RZK_UNROLL
            for (int c = 0; c < 4; ++c) {
                uint32_t m = nzw[li_][c];
                RZK_NOUNROLL
                while (m) {
                    const uint32_t bit = rot_ctz(m);                        // 7, 15, 23 or 31
                    m &= m - 1u;
                    const uint32_t pos = 16u * (uint32_t)ctx.ridx + 4u * (uint32_t)c + (bit >> 3);
                    const uint32_t minus = ((dw[li_][c] >> bit) & 1u) ^ (neg ? 1u : 0u);    // 1: the contribution is -X^pos c
                    // added slot: contribution row; subtracted slot: the opposite row
                    const uint32_t far = minus ^ (slot & 1u);               // 1: the row from offset 1024 - pos (B - X^pos c)
                    list[slot] = (uint16_t)((far ? 1024u : 512u) - pos);
                    ++slot;
                }
            }
            if (ctx.ridx == 0) list[n_terms] = (uint16_t)kRotZeroOff;      // partner of the last term of an odd count
            uint32_t *zero = ctx.red + kRotZeroOff;
            RZK_UNROLL
            for (int e = 0; e < 16; ++e) zero[ctx.ridx + 32 * e] = 0x80000000u;
        }
        RZK_SYNC();
        if (bad_any) {
            // Not a challenge-space polynomial (an entry outside {-1, 0, 1}): the general form, still exact for any int8 d
            // (|sum| <= 512 * 128 * 2^31 < 2^53) -- every non-zero entry is one rotation scaled by its value, the biased word
            // un-biased first.  Three instructions per coefficient and term instead of two; honest verifiers never get here.
            RZK_EACH_LANE {
                const LaneCtx &ctx = ctxs[li_];
                uint4 q;
                q.x = dw[li_][0]; q.y = dw[li_][1]; q.z = dw[li_][2]; q.w = dw[li_][3];
                reinterpret_cast<uint4 *>(ctx.red + kRotListOff)[ctx.ridx] = q;        // d as 512 bytes over the (unused) list
            }
            RZK_SYNC();
            const double unbias = 4503599627370496.0 + 2147483648.0;
            const double sgn = neg ? -1.0 : 1.0;
            RZK_NOUNROLL
            for (uint32_t k = 0; k < (uint32_t)kN; ++k) {
                const int32_t dv = (int32_t)reinterpret_cast<const int8_t *>(ctxs[0].red + kRotListOff)[k];
                if (dv == 0) continue;
                RZK_EACH_LANE {
                    const LaneCtx &ctx = ctxs[li_];
                    const double sv = sgn * f64_exact_i32(dv);
                    const uint32_t *row = ctx.red + (512u - k) + (uint32_t)(ctx.t + 16 * ctx.hw);
                    RZK_UNROLL
                    for (int j = 0; j < CNT; ++j) V[li_][j] = f64_exact_fma(sv, rot_biased(row[32 * j]) - unbias, V[li_][j]);
                }
            }
        } else {
            // the high words of the doubles 2^52 + u: one register per load slot, written once (the `volatile` keeps the
            // compiler from re-materialising the constant next to every load inside the loop)
            uint32_t hi_a[CNT], hi_s[CNT];
            RZK_UNROLL
            for (int j = 0; j < CNT; ++j) {
#if defined(__CUDA_ARCH__)
                // (item < 2^28, so the shifted term is 0 -- but only at run time: a plain constant would be re-materialised
                // by a move in front of every load, which is what this avoids)
                hi_a[j] = 0x43300000u | ((ctxs[0].item + (uint32_t)j) >> 31);
                hi_s[j] = 0x43300000u | ((ctxs[0].item + (uint32_t)(CNT + j)) >> 31);
#else
                hi_a[j] = hi_s[j] = 0x43300000u;
#endif
            }
            RZK_NOUNROLL
            for (uint32_t k = 0; k < n_pairs; ++k) {
                RZK_EACH_LANE {
                    const LaneCtx &ctx = ctxs[li_];
                    const uint32_t both = (ctx.red + kRotListOff)[k];       // slots 2k (added) and 2k + 1 (subtracted)
                    const uint32_t *ra = ctx.red + (both & 0xffffu) + (uint32_t)(ctx.t + 16 * ctx.hw);
                    const uint32_t *rs = ctx.red + (both >> 16) + (uint32_t)(ctx.t + 16 * ctx.hw);
                    RZK_UNROLL
                    for (int j = 0; j < CNT; ++j)
                        V[li_][j] = (V[li_][j] + rot_biased_hi(hi_a[j], ra[32 * j])) - rot_biased_hi(hi_s[j], rs[32 * j]);
                }
            }
            // (every added word carried the bias 2^31 and every subtracted one too -- the zero row included -- so the biases cancel)
        }
        RZK_SYNC();          // the region is free again (next item's transposes)
    }
}

// Garner recombination of the residues of one coefficient into a signed 64-bit value
// congruent to the exact integer result modulo q (exact integer itself for np <= 2).
template <int NP>
RZK_VM int64_t crt_combine(const VmLaunch &K, const uint32_t (&r)[kMaxPrimes])
{
    if (NP == 1) {
        const uint32_t a0 = r[0];
        return a0 > K.pc[0].half ? (int64_t)a0 - (int64_t)K.pc[0].p : (int64_t)a0;
    }
    const uint32_t p0 = K.pc[0].p, p1 = K.pc[1].p;
    const uint32_t a0 = r[0], a1 = r[1];
    uint32_t h1 = shoup_mul(K.crt.inv01, K.crt.inv01p, a1 - a0 + 2u * p1, p1);
    h1 = csub(h1, p1);
    const uint64_t v01 = (uint64_t)a0 + (uint64_t)p0 * (uint64_t)h1;     // [0, p0*p1)
    if (NP == 2) {
        return v01 > K.crt.P01half ? (int64_t)(v01 - K.crt.P01) : (int64_t)v01;
    }
    const uint32_t p2 = K.pc[2].p;
    const uint32_t a2 = r[2];
    uint32_t s = csub(shoup_mul(K.crt.p0modp2, K.crt.p0modp2p, h1, p2), p2);   // p0*h1 mod p2
    const uint32_t a0r = csub(a0, p2);                                         // a0 < p0 < 2*p2
    uint32_t h2 = shoup_mul(K.crt.inv012, K.crt.inv012p, a2 + 2u * p2 - a0r - s, p2);
    h2 = csub(h2, p2);
    // V = v01 + P01*h2 as a 128-bit integer; negative (centred) iff V > (P-1)/2
    const uint64_t lo_prod = K.crt.P01 * (uint64_t)h2;
    const uint64_t hi_prod = mulhi64(K.crt.P01, (uint64_t)h2);
    const uint64_t lo = lo_prod + v01;
    const uint64_t hi = hi_prod + (lo < lo_prod ? 1u : 0u);
    const bool negv = (hi > K.crt.Phalf_hi) || (hi == K.crt.Phalf_hi && lo > K.crt.Phalf_lo);
    // a value congruent to V mod q with |w| < 2^60 + 2^33 (the range reduce_q_centered accepts):
    // (P01 mod q) * h2 < 2^62 is first brought to [0, 2q) with the same top-bits Barrett step
    uint64_t tq = K.crt.P01modq * (uint64_t)h2;
    tq -= (uint64_t)mulhi32((uint32_t)(tq >> 30), K.m30) * (uint64_t)K.q;
    int64_t w = (int64_t)(v01 + tq);
    if (negv) w -= (int64_t)K.crt.Pmodq;
    return w;
}

// Two-prime Garner recombination straight to a value congruent to the exact integer result modulo q, as an exact
// integer in a double: V = a0 + p0 * h1 with the digit h1 = (a1 - a0) * p0^-1 mod p1 taken centred.  For every result the
// range checks admit (|V| <= P01/2 - 2^33) that IS the integer result (the centred digit cannot be off by p1: the
// difference would exceed the range by P01); (p0 * h1) mod q is an error-free FMA product reduced by rint(. / q) * q.
RZK_VM double crt2_mod_q_f64(const VmLaunch &K, uint32_t a0, uint32_t a1)
{
    const uint32_t p1 = K.pc[1].p;
    uint32_t h1 = shoup_mul(K.crt.inv01, K.crt.inv01p, a1 - a0 + 2u * p1, p1);
    h1 = csub(h1, p1);
    const int32_t h1c = h1 > K.pc[1].half ? (int32_t)(h1 - p1) : (int32_t)h1;
    const double x = f64_exact_i32(h1c);
    const double magic = 6755399441055744.0;
#if defined(__CUDA_ARCH__)
    const double h = __dmul_rn(x, K.p0d);
    const double l = __fma_rn(x, K.p0d, -h);
    const double k = __dadd_rn(__fma_rn(x, K.p0qinvd, magic), -magic);
    const double r = __dadd_rn(__fma_rn(-k, K.qd, h), l);
#else
    const double h = x * K.p0d;
    const double l = __builtin_fma(x, K.p0d, -h);
    const double k = __builtin_fma(x, K.p0qinvd, magic) - magic;
    const double r = __builtin_fma(-k, K.qd, h) + l;
#endif
    return f64_exact_i32((int32_t)a0) + r;
}

// Three-prime MODE_SEQ programs finish an output in chunks of four coefficients per lane (one uint4 of the lane-private
// layout): the residues of the earlier primes wait in a stash in GLOBAL memory (K.gstash, written and read back by the
// same lane, L2-resident), and only four 64-bit values are live at a time.  That keeps the kernel within 128 registers
// (16 warps per SM instead of the 8 that 223 registers allowed) and takes 4 KB per half warp out of shared memory.
template <int NP, int MODE>
struct ChunkedEpi { static constexpr bool value = (mode_seq(MODE) && NP == 3); };
constexpr int kEpiChunk = 4;

template <int NP>
RZK_VM void seq_chunk_values(const VmLaunch &K, Lane *lanes, const LaneCtx *ctxs, const Op &op, int j, int64_t (&V)[RZK_NL][kEpiChunk])
{
    RZK_EACH_LANE {
        RZK_LANE;
        uint4 prev[NP > 1 ? NP - 1 : 1];
        RZK_UNROLL
        for (int k = 0; k < NP - 1; ++k)
            prev[k] = reinterpret_cast<const uint4 *>(ctx.stash + ((int)op.b * (NP - 1) + k) * kSlotWords)[j * kLanes + t];
        RZK_UNROLL
        for (int c = 0; c < kEpiChunk; ++c) {
            uint32_t r[kMaxPrimes] = {0, 0, 0};
            RZK_UNROLL
            for (int k = 0; k < NP - 1; ++k) r[k] = (c == 0) ? prev[k].x : (c == 1) ? prev[k].y : (c == 2) ? prev[k].z : prev[k].w;
            r[NP - 1] = L.cur[kEpiChunk * j + c];
            V[li_][c] = crt_combine<NP>(K, r);
        }
    }
}

// OP_ADDP / OP_FIN on the four coefficients m = 4j .. 4j+3 of every lane (MODE_SEQ; i = t + 16 m)
RZK_VM void op_addp_chunk(const VmLaunch &K, const LaneCtx *ctxs, int64_t (&V)[RZK_NL][kEpiChunk], const Op &op, int it, uint32_t dtype, int j)
{
    const Stream st = K.st[op.a];
    const bool neg = op.c & MAC_NEG;
    RZK_EACH_LANE {
        const LaneCtx &ctx = ctxs[li_];
        const int t = ctx.t;
        const uint64_t poly = stream_poly(st, ctx.item, (uint32_t)op.off + (uint32_t)it * op.step);
        RZK_UNROLL
        for (int c = 0; c < kEpiChunk; ++c) {
            const int i = t + kLanes * (kEpiChunk * j + c);
            const int32_t v = (dtype == DT_I8) ? (int32_t)(reinterpret_cast<const int8_t *>(st.base) + poly * kN)[i]
                                               : (reinterpret_cast<const int32_t *>(st.base) + poly * kN)[i];
            V[li_][c] += neg ? -(int64_t)v : (int64_t)v;
        }
    }
}

RZK_VM void op_fin_chunk(const VmLaunch &K, Lane *lanes, const LaneCtx *ctxs, int64_t (&V)[RZK_NL][kEpiChunk], const Op &op, int it, int j)
{
    const Stream st = K.st[op.a];
    RZK_EACH_LANE {
        RZK_LANE;
        const uint64_t poly = stream_poly(st, ctx.item, (uint32_t)op.off + (uint32_t)it * op.step);
        int32_t res[kEpiChunk];
        RZK_UNROLL
        for (int c = 0; c < kEpiChunk; ++c) res[c] = reduce_q_centered(V[li_][c], K.q, K.m30, K.kqh);
        if (op.b & FIN_CMPZ) {
            uint32_t nz = 0;
            RZK_UNROLL
            for (int c = 0; c < kEpiChunk; ++c) nz |= (uint32_t)res[c];
            L.fail |= nz ? 1u : 0u;
        }
        if ((op.b & FIN_STORE) && ctx.active) {
            int32_t *dst = reinterpret_cast<int32_t *>(const_cast<void *>(st.base)) + poly * kN;
            RZK_UNROLL
            for (int c = 0; c < kEpiChunk; ++c) dst[t + kLanes * (kEpiChunk * j + c)] = res[c];
        }
    }
}

// Inverse transform of acc[a].  On the last prime the residues of all primes are combined and
// the epilogue ops that follow (OP_ADDP*, OP_FIN) are executed here, so that the 64-bit
// values live only inside this function.  Returns the index of the first op after the epilogue.
template <int NP, int MODE>
RZK_VM void inv_core(const VmLaunch &K, Lane *lanes, const LaneCtx *ctxs, const Op &op, int prime_iter,
                     typename Epi<MODE>::V_t (&V)[RZK_NL][Epi<MODE>::kCount],
                     const int32_t (*small)[Epi<MODE>::kCount] = nullptr, const Op *small_late = nullptr, int it = 0)
{
    constexpr int CNT = Epi<MODE>::kCount;
    int32_t small_buf[RZK_NL][CNT];      // `small_late`: the int8 plain term is fetched AFTER the transform (no 16 registers held across it)
    (void)small_buf;
    RZK_SYNC();      // every OP_LD of the partner half warp has finished (the slot may overlay this buffer)
#if defined(__CUDA_ARCH__)
    pp_acquire(K);
#endif
    RZK_EACH_LANE {
        RZK_LANE;
        const PrimeC &pc = L.pc;
        if (op.a == 0) {
            RZK_UNROLL
            for (int e = 0; e < kElems; ++e) L.cur[e] = L.acc0[e];
        } else {
            const uint4 *a4 = reinterpret_cast<const uint4 *>(ctx.acc1);
            RZK_UNROLL
            for (int j = 0; j < 8; ++j) {
                const uint4 a = a4[j * kLanes + t];
                L.cur[4 * j + 0] = a.x; L.cur[4 * j + 1] = a.y; L.cur[4 * j + 2] = a.z; L.cur[4 * j + 3] = a.w;
            }
        }
        if constexpr (MODE == MODE_SEQ_S) {
            // a product sum of up to 33 terms since its last reduction: back below 1.06 p (the inverse admits 5p/2)
            RZK_UNROLL
            for (int e = 0; e < kElems; ++e) L.cur[e] = sreduce<kSignedShift>(L.cur[e], 0u - pc.p);
        }
#if RZK_INV_DIT
        if constexpr (mode_signed(MODE)) inv_g2_dit(L.cur, ctx.g1 + (L.pi * 2 + 1) * kG1Words, 0u - pc.p);
#else
        if constexpr (mode_signed(MODE)) inv_g2_s(L.cur, ctx.g2 + ((L.pi * 2 + 1) * kLanes + t) * kG2Words, 0u - pc.p);
#endif
        else inv_g2(L.cur, ctx.g2 + ((L.pi * 2 + 1) * kLanes + t) * kG2Words, pc.p, pc.p2, pc.pad_, L.cap);
        uint4 *row = reinterpret_cast<uint4 *>(ctx.buf + 36 * t);
        RZK_UNROLL
        for (int j = 0; j < 8; ++j) {
            uint4 w;
            w.x = L.cur[4 * j + 0]; w.y = L.cur[4 * j + 1]; w.z = L.cur[4 * j + 2]; w.w = L.cur[4 * j + 3];
            row[j] = w;
        }
    }
    RZK_SYNC();
    RZK_EACH_LANE {
        RZK_LANE;
        const PrimeC &pc = L.pc;
        RZK_UNROLL
        for (int m = 0; m < kElems; ++m) {
            const int i = t + kLanes * m;
            L.cur[m] = ctx.buf[i + ((i >> 5) << 2)];
        }
        if constexpr (MODE == MODE_SPLITKEY_S) {
#if RZK_INV_DIT
            inv_g1_dit<true>(L.cur, ctx.g1 + (L.pi * 2 + 1) * kG1Words, ctx.g2 + ((L.pi * 2 + 1) * kLanes + t) * kG2Words, ctx.twist + L.pi * kTwistWords, t, 0u - pc.p);
#else
            inv_g1_s<true>(L.cur, ctx.g1 + (L.pi * 2 + 1) * kG1Words, 0u - pc.p);    // any representative below 20.2 p: centred in the epilogue
#endif
        } else if constexpr (MODE == MODE_SEQ_S) {
            // the Garner recombination wants the canonical residue in [0, p): shift-reduce to (-0.06 p, 1.06 p), lift the
            // negative ones by p, one conditional subtraction
            const uint32_t mp = 0u - pc.p;
#if RZK_INV_DIT
            inv_g1_dit<false>(L.cur, ctx.g1 + (L.pi * 2 + 1) * kG1Words, ctx.g2 + ((L.pi * 2 + 1) * kLanes + t) * kG2Words, ctx.twist + L.pi * kTwistWords, t, mp);
#else
            inv_g1_s<false>(L.cur, ctx.g1 + (L.pi * 2 + 1) * kG1Words, mp);
#endif
            RZK_UNROLL
            for (int m = 0; m < kElems; ++m) {
                const uint32_t r = sreduce<kSignedShift>(L.cur[m], mp);
                const uint32_t lifted = r + (((uint32_t)((int32_t)r >> 31)) & pc.p);
                L.cur[m] = csub(lifted, pc.p);
            }
        } else {
            inv_g1(L.cur, ctx.g1 + (L.pi * 2 + 1) * kG1Words, pc.p, pc.p2, pc.pad_, L.cap);
            RZK_UNROLL
            for (int m = 0; m < kElems; ++m) L.cur[m] = csub(L.cur[m], pc.p);     // [0,2p) -> [0,p)
        }
    }
    RZK_SYNC();
#if defined(__CUDA_ARCH__)
    pp_release(K, *ctxs[0].pp_count);
#endif
    if constexpr (!mode_seq(MODE)) {
        if (small_late) { op_addp_load<MODE, true>(K, ctxs, small_buf, *small_late, it, DT_I8); small = small_buf; }
    }
    if (!mode_seq(MODE)) {
        // MODE_SPLIT: half warp 0 holds residues mod p0, half warp 1 mod p1, both for all 512
        // coefficients.  MODE_SPLITKEY: half warp 0 holds the lo part, half warp 1 the hi part.
        // Lane (h, t) finishes the rows m = 2j + h (epi_m): it keeps its own value of those
        // and receives the partner's; it sends its values of the other rows.
        if (MODE == MODE_SPLITKEY) {
            RZK_EACH_LANE {
                RZK_LANE;
                const uint32_t p = L.pc.p, half = L.pc.half;
                RZK_UNROLL
                for (int m = 0; m < kElems; ++m) L.cur[m] = L.cur[m] > half ? L.cur[m] - p : L.cur[m];   // centred lift
            }
        }
        uint32_t recv[RZK_NL][16];
#if defined(__CUDA_ARCH__)
        RZK_UNROLL
        for (int j = 0; j < 16; ++j) {
            const uint32_t send = ctxs[0].hw ? lanes[0].cur[2 * j] : lanes[0].cur[2 * j + 1];
            recv[0][j] = __shfl_xor_sync(0xffffffffu, send, 16);
        }
#else
        for (int li = 0; li < RZK_NL; ++li)
            for (int j = 0; j < 16; ++j) {
                const int partner = li ^ 16;
                recv[li][j] = ctxs[partner].hw ? lanes[partner].cur[2 * j] : lanes[partner].cur[2 * j + 1];
            }
#endif
        RZK_EACH_LANE {
            RZK_LANE;
            RZK_UNROLL
            for (int j = 0; j < 16; ++j) {
                const uint32_t own = ctx.hw ? L.cur[2 * j + 1] : L.cur[2 * j];
                const uint32_t v0 = ctx.hw ? recv[li_][j] : own;      // half warp 0's value
                const uint32_t v1 = ctx.hw ? own : recv[li_][j];      // half warp 1's value
                if (MODE == MODE_SPLITKEY) {
                    // `small`: an int8 plain term of the epilogue (the r0 / r1 row of a commitment) joins the lo part as an
                    // integer (|lo| < 2^29, |small| < 2^7): one add instead of a conversion and a DADD
                    const int32_t lo = small ? (int32_t)v0 + small[li_][j % CNT] : (int32_t)v0;
                    V[li_][j % CNT] = f64_exact_fma(f64_exact_i32((int32_t)v1), 65536.0, f64_exact_i32(lo));
                } else if (MODE == MODE_SPLITKEY_S) {
                    // the residues arrive as arbitrary representatives (|v| < 20.2 p + 2^7 < 2^31): centred on the FP64 pipe,
                    // v - p * rint(v / p), exact because the true parts are below p/2 - 2^14 in magnitude
                    // (the words carry the bias 2^31 of the last inverse stage, the form the exact conversion takes)
                    const uint32_t lo = small ? v0 + (uint32_t)small[li_][j % CNT] : v0;
                    const double lod = center_p_f64(f64_exact_biased(lo), K.p0d, K.p0invd);
                    const double hid = center_p_f64(f64_exact_biased(v1), K.p0d, K.p0invd);
                    V[li_][j % CNT] = f64_exact_fma(hid, 65536.0, lod);
                } else {
                    V[li_][j % CNT] = crt2_mod_q_f64(K, v0, v1);
                }
            }
        }
    } else {
        RZK_EACH_LANE {
            RZK_LANE;
            if (prime_iter < NP - 1) {
                uint4 *s4 = reinterpret_cast<uint4 *>(ctx.stash + ((int)op.b * (NP - 1) + prime_iter) * kSlotWords);
                RZK_UNROLL
                for (int j = 0; j < 8; ++j) {
                    uint4 w;
                    w.x = L.cur[4 * j + 0]; w.y = L.cur[4 * j + 1]; w.z = L.cur[4 * j + 2]; w.w = L.cur[4 * j + 3];
                    s4[j * kLanes + t] = w;
                }
            } else if constexpr (!ChunkedEpi<NP, MODE>::value) {
                uint32_t prev[NP > 1 ? NP - 1 : 1][kElems];
                RZK_UNROLL
                for (int k = 0; k < NP - 1; ++k) {
                    const uint4 *s4 = reinterpret_cast<const uint4 *>(ctx.stash + ((int)op.b * (NP - 1) + k) * kSlotWords);
                    RZK_UNROLL
                    for (int j = 0; j < 8; ++j) {
                        const uint4 w = s4[j * kLanes + t];
                        prev[k][4 * j + 0] = w.x; prev[k][4 * j + 1] = w.y; prev[k][4 * j + 2] = w.z; prev[k][4 * j + 3] = w.w;
                    }
                }
                RZK_UNROLL
                for (int m = 0; m < kElems; ++m) {
                    uint32_t r[kMaxPrimes] = {0, 0, 0};
                    RZK_UNROLL
                    for (int k = 0; k < NP - 1; ++k) r[k] = prev[k][m];
                    r[NP - 1] = L.cur[m];
                    V[li_][m % CNT] = crt_combine<NP>(K, r);
                }
            }
        }
    }
}

// Generic (runtime-decoded) form: inverse transform, then the epilogue ops that follow.
// Returns the index of the first op after the epilogue.
template <int NP, int MODE>
RZK_VM int op_inv(const VmLaunch &K, Lane *lanes, const LaneCtx *ctxs, int q, int it, int prime_iter)
{
    const Op op = K.ops[q];
    const bool last = (!mode_seq(MODE)) || (prime_iter == NP - 1);
    typename Epi<MODE>::V_t V[RZK_NL][Epi<MODE>::kCount];
    inv_core<NP, MODE>(K, lanes, ctxs, op, prime_iter, V);
    ++q;
    if constexpr (ChunkedEpi<NP, MODE>::value) {
        int qe = q;
        while (K.ops[qe].code == OP_ADDP || K.ops[qe].code == OP_FIN) ++qe;
        if (last) {
            RZK_UNROLL
            for (int j = 0; j < kElems / kEpiChunk; ++j) {
                int64_t V4[RZK_NL][kEpiChunk];
                seq_chunk_values<NP>(K, lanes, ctxs, op, j, V4);
                RZK_NOUNROLL
                for (int qq = q; qq < qe; ++qq) {
                    const Op e = K.ops[qq];
                    if (e.code == OP_ADDP) op_addp_chunk(K, ctxs, V4, e, it, K.st[e.a].dtype, j);
                    else op_fin_chunk(K, lanes, ctxs, V4, e, it, j);
                }
            }
        }
        return qe;
    } else {
        RZK_NOUNROLL
        for (;; ++q) {
            const Op e = K.ops[q];
            if (e.code == OP_ADDP) { if (last) op_addp<MODE>(K, ctxs, V, e, it, K.st[e.a].dtype); }
            else if (e.code == OP_ROT) { if (last) op_rot<MODE>(K, lanes, ctxs, V, e, it); }
            else if (e.code == OP_FIN) { if (last) op_fin<MODE>(K, lanes, ctxs, V, e, it); }
            else break;
        }
        return q;
    }
}

// The verifier's norm check in the warp-per-item modes on the device (int32 rows, bound below 2^21: see op_norm): straight-line
// code with the NEXT row in flight while the current one is squared -- the rows are the item's first touch of global memory, and
// each exposed L2 / DRAM latency is overlapped with the arithmetic of the row before.
#if defined(__CUDA_ARCH__)
template <int MODE>
__device__ __forceinline__ void op_norm_fast32(const VmLaunch &K, Lane &L, const LaneCtx &ctx, const Op &op)
{
    constexpr int CNT = Epi<MODE>::kCount;
    static_assert(CNT == 16, "warp-per-item modes");
    const Stream st = K.st[op.a];
    const uint64_t sq_lim = K.norm_sq_lim[op.b];
    const uint64_t lane_lim = sq_lim >> 5;
    const int32_t *src = reinterpret_cast<const int32_t *>(st.base) + stream_poly(st, ctx.item, (uint32_t)op.off) * kN + ctx.ridx * CNT;
    uint4 cur[CNT / 4], nxt[CNT / 4];
    RZK_UNROLL
    for (int j = 0; j < CNT / 4; ++j) cur[j] = rot_ld128(src + 4 * j);
    RZK_NOUNROLL
    for (int c = 0; c < (int)op.c; ++c) {
        if (c + 1 < (int)op.c) {
            RZK_UNROLL
            for (int j = 0; j < CNT / 4; ++j) nxt[j] = rot_ld128(src + (size_t)(c + 1) * kN + 4 * j);
        }
        int64_t acc = 0;
        uint32_t mx = 0;
        RZK_UNROLL
        for (int j = 0; j < CNT / 4; ++j) {
            const int32_t vv[4] = {(int32_t)cur[j].x, (int32_t)cur[j].y, (int32_t)cur[j].z, (int32_t)cur[j].w};
            RZK_UNROLL
            for (int e = 0; e < 4; ++e) {
                mx = umax32(mx, (uint32_t)vv[e] + (1u << 21));
                acc += (int64_t)vv[e] * (int64_t)vv[e];
            }
        }
        const uint32_t bad = (mx >> 22) ? 1u : 0u;
        const uint64_t s = bad ? 0ull : (uint64_t)acc;
        L.fail |= bad;
        if (__any_sync(0xffffffffu, s > lane_lim)) {
            uint32_t lo = (uint32_t)s, hi = (uint32_t)(s >> 32);
            RZK_UNROLL
            for (int d = 16; d >= 1; d >>= 1) {
                const uint32_t olo = __shfl_xor_sync(0xffffffffu, lo, d);
                const uint32_t ohi = __shfl_xor_sync(0xffffffffu, hi, d);
                const uint64_t sum = (((uint64_t)hi << 32) | lo) + (((uint64_t)ohi << 32) | olo);
                lo = (uint32_t)sum; hi = (uint32_t)(sum >> 32);
            }
            L.fail |= ((((uint64_t)hi << 32) | lo) > sq_lim) ? 1u : 0u;
        }
        RZK_UNROLL
        for (int j = 0; j < CNT / 4; ++j) cur[j] = nxt[j];
    }
}
#endif

// params.rs:102-118 via polynomial.rs:60-73: floor(sqrt(sum c^2)) <= bound  <=>  sum c^2 < (bound+1)^2
template <int MODE>
RZK_VM void op_norm(const VmLaunch &K, Lane *lanes, const LaneCtx *ctxs, const Op &op, uint32_t dtype)
{
    constexpr int CNT = Epi<MODE>::kCount;
    constexpr int RED_N = (!mode_seq(MODE)) ? 32 : 16;
    const Stream st = K.st[op.a];
    const uint32_t abs_lim = K.norm_abs_lim[op.b];
    const uint64_t sq_lim = K.norm_sq_lim[op.b];
#if defined(__CUDA_ARCH__)
    if constexpr (!mode_seq(MODE)) {
        if (dtype != DT_I8 && abs_lim < (1u << 21)) { op_norm_fast32<MODE>(K, lanes[0], ctxs[0], op); return; }
    }
#endif
    RZK_NOUNROLL
    for (int c = 0; c < (int)op.c; ++c) {
        RZK_EACH_LANE {
            RZK_LANE;
            const uint64_t poly = stream_poly(st, ctx.item, (uint32_t)op.off + (uint32_t)c);
            uint64_t s = 0;
            uint32_t bad = 0;
            if (dtype == DT_I8) {
                // any partition of the 512 coefficients works for a norm: CNT contiguous bytes per lane
                const int32_t *src = reinterpret_cast<const int32_t *>(reinterpret_cast<const int8_t *>(st.base) + poly * kN) +
                                     ctx.ridx * (CNT / 4);
                uint32_t acc = 0;
                RZK_UNROLL
                for (int j = 0; j < CNT / 4; ++j) acc = dot4_i8(src[j], acc);
                s = acc;
            } else if (abs_lim < (1u << 21)) {
                // Bounds below 2^21 (the default parameters): a coefficient can only pass if its int32 value is
                // already the small canonical residue (a non-canonical int32 representative is > 1.3e9 in magnitude
                // once centred), so |v| < 2^21 is tested on the raw value -- one add-and-max per coefficient -- and the
                // squares are summed in a 64-bit integer (one IMAD.WIDE each; exact whenever the range test passes:
                // CNT * 2^42 < 2^63).  Any partition of the 512 coefficients works for a norm: CNT contiguous words per
                // lane, fetched with 128-bit loads (the warp reads the 2 KB row as one contiguous run).
                const int32_t *src = reinterpret_cast<const int32_t *>(st.base) + poly * kN + ctx.ridx * CNT;
                int64_t acc = 0;
                uint32_t mx = 0;
                RZK_UNROLL
                for (int j = 0; j < CNT / 4; ++j) {
                    const uint4 q4 = rot_ld128(src + 4 * j);
                    const int32_t vv[4] = {(int32_t)q4.x, (int32_t)q4.y, (int32_t)q4.z, (int32_t)q4.w};
                    RZK_UNROLL
                    for (int e = 0; e < 4; ++e) {
                        mx = umax32(mx, (uint32_t)vv[e] + (1u << 21));
                        acc += (int64_t)vv[e] * (int64_t)vv[e];
                    }
                }
                bad = (mx >> 22) ? 1u : 0u;
                s = bad ? 0ull : (uint64_t)acc;
            } else {
                const int32_t *src = reinterpret_cast<const int32_t *>(st.base) + poly * kN;
                RZK_UNROLL
                for (int j = 0; j < CNT; ++j) {
                    const int32_t v = canon_q(src[t + kLanes * epi_m<MODE>(ctx, j)], K.q);
                    const uint32_t av = (uint32_t)(v < 0 ? -v : v);
                    const bool big = av > abs_lim;
                    bad |= big ? 1u : 0u;
                    s += big ? 0ull : (uint64_t)av * (uint64_t)av;
                }
            }
            L.fail |= bad;
#if defined(__CUDA_ARCH__)
            // RED_N lane sums that are each at most sq_lim / RED_N cannot exceed the bound together (an honest response sits at
            // a quarter of it): one vote then replaces the reduction; otherwise
            // sum over the lanes that own the item (half warp in MODE_SEQ, full warp otherwise)
            if (__any_sync(0xffffffffu, s > sq_lim / (uint64_t)RED_N)) {
                uint32_t lo = (uint32_t)s, hi = (uint32_t)(s >> 32);
                RZK_UNROLL
                for (int d = RED_N / 2; d >= 1; d >>= 1) {
                    const uint32_t olo = __shfl_xor_sync(0xffffffffu, lo, d);
                    const uint32_t ohi = __shfl_xor_sync(0xffffffffu, hi, d);
                    const uint64_t sum = (((uint64_t)hi << 32) | lo) + (((uint64_t)ohi << 32) | olo);
                    lo = (uint32_t)sum; hi = (uint32_t)(sum >> 32);
                }
                const uint64_t tot = ((uint64_t)hi << 32) | lo;
                L.fail |= (tot > sq_lim) ? 1u : 0u;
            }
        }
#else
            ctx.red[2 * ctx.ridx] = (uint32_t)s;
            ctx.red[2 * ctx.ridx + 1] = (uint32_t)(s >> 32);
        }
        RZK_EACH_LANE {
            RZK_LANE;
            uint64_t tot = 0;
            for (int j = 0; j < RED_N; ++j) tot += (uint64_t)ctx.red[2 * j] | ((uint64_t)ctx.red[2 * j + 1] << 32);
            L.fail |= (tot > sq_lim) ? 1u : 0u;
        }
#endif
    }
}

// ---------------------------------------------------------------- interpreter

// Folds the status words of the lanes that own an item into the item group's flag word.  A range error goes to the
// launch's mark array instead when there is one (K.rmark: the masked fallback launch of dev_commit redoes the group).
template <int RED_N>
RZK_VM void fold_flags(const VmLaunch &K, Lane *lanes, const LaneCtx *ctxs)
{
#if defined(__CUDA_ARCH__)
    uint32_t f = lanes[0].fail | (lanes[0].rerr << 1);
    RZK_UNROLL
    for (int d = RED_N / 2; d >= 1; d >>= 1) f |= __shfl_xor_sync(0xffffffffu, f, d);
    if (ctxs[0].ridx == 0 && ctxs[0].active && f) {
        const uint32_t grp = ctxs[0].item / K.flag_div;
        if ((f & FLAG_RANGE) && K.rmark) { atomicOr(&K.rmark[grp], 1u); atomicOr(K.rmark_any, 1u); f &= ~(uint32_t)FLAG_RANGE; }
        if (f) atomicOr(&K.flags[grp], f);
    }
#else
    RZK_EACH_LANE { RZK_LANE; ctx.red[ctx.ridx] = L.fail | (L.rerr << 1); }
    RZK_EACH_LANE {
        RZK_LANE;
        if (ctx.ridx == 0 && ctx.active) {
            uint32_t f = 0;
            for (int j = 0; j < RED_N; ++j) f |= ctx.red[j];
            const uint32_t grp = ctx.item / K.flag_div;
            if ((f & FLAG_RANGE) && K.rmark) { K.rmark[grp] |= 1u; *K.rmark_any |= 1u; f &= ~(uint32_t)FLAG_RANGE; }
            if (f) K.flags[grp] |= f;
        }
    }
#endif
}

// Runs the whole program for the item(s) owned by this warp.
template <int NP, int MODE>
RZK_VM void vm_run_item(const VmLaunch &K, Lane *lanes, const LaneCtx *ctxs)
{
    static_assert(MODE != MODE_SPLIT || NP == 2, "MODE_SPLIT maps the two half warps to two primes");
    static_assert(!mode_sk(MODE) || NP == 1, "MODE_SPLITKEY works modulo one prime");
    constexpr int RED_N = (!mode_seq(MODE)) ? 32 : 16;
    constexpr int PRIME_ITERS = (!mode_seq(MODE)) ? 1 : NP;
    constexpr int KP = mode_sk(MODE) ? 2 * kKeyPolys : kKeyPolys;     // key images per prime
    RZK_EACH_LANE { RZK_LANE; L.fail = 0; L.rerr = 0; }
    int pc = 0;
    RZK_NOUNROLL
    while (K.ops[pc].code == OP_NORM) {
        op_norm<MODE>(K, lanes, ctxs, K.ops[pc], K.st[K.ops[pc].a].dtype);
        ++pc;
    }
    RZK_NOUNROLL
    while (K.ops[pc].code == OP_SEG) {
        const int seg_begin = pc + 1;
        int seg_end = seg_begin;
        RZK_NOUNROLL
        for (int prime_iter = 0; prime_iter < PRIME_ITERS; ++prime_iter) {
            RZK_EACH_LANE {
                RZK_LANE;
                L.pi = (MODE == MODE_SPLIT) ? ctx.hw : (mode_sk(MODE) ? 0 : prime_iter);
                L.pc = K.pc[L.pi];
                L.cap = kAddCap;
            }
            int q = seg_begin, loop_start = 0, loop_cnt = 0, it = 0;
            RZK_NOUNROLL
            for (;;) {
                const Op op = K.ops[q];
                if (op.code == OP_SEG || op.code == OP_END) break;
                switch (op.code) {
                case OP_FWD:
#if defined(__CUDA_ARCH__)
                    cta_lockstep(K);
#endif
                    op_fwd<mode_signed(MODE), MODE == MODE_SPLITKEY_S>(K, lanes, ctxs, op, it, K.st[op.a].dtype);
                    break;
                case OP_MACK:
                    RZK_EACH_LANE {
                        RZK_LANE;
                        // MODE_SPLITKEY: half warp h uses the lo (h = 0) / hi (h = 1) image of key poly b
                        const int kidx = mode_sk(MODE) ? (2 * (int)op.b + ctx.hw) : (int)op.b;
                        const uint32_t *krow = ctx.key + ((L.pi * KP + kidx) * 2) * kPadWords;
                        if constexpr (MODE == MODE_SPLITKEY_S) {
                            if (op.a == 0) mac_key_s(L.acc0, L.cur, krow, t, op.c, 0u - L.pc.p);
                            else mac_key_smem_s(ctx.acc1, L.cur, krow, t, op.c, 0u - L.pc.p);
                        } else {
                            if (op.a == 0) mac_key(L.acc0, L.cur, krow, t, op.c, L.pc.p, L.pc.p2);
                            else mac_key_smem(ctx.acc1, L.cur, krow, t, op.c, L.pc.p, L.pc.p2);
                        }
                    }
                    break;
                case OP_MACV:
                    RZK_EACH_LANE {
                        RZK_LANE;
                        if constexpr (MODE == MODE_SEQ_S) {
                            if (op.a == 0) mac_var_s(L.acc0, L.cur, ctx.slot, t, op.c, L.pc.p, L.pc.pinv);
                            else mac_var_smem_s(ctx.acc1, L.cur, ctx.slot, t, op.c, L.pc.p, L.pc.pinv);
                        } else {
                            if (op.a == 0) mac_var(L.acc0, L.cur, ctx.slot, t, op.c, L.pc.p, L.pc.p2, L.pc.pinv);
                            else mac_var_smem(ctx.acc1, L.cur, ctx.slot, t, op.c, L.pc.p, L.pc.p2, L.pc.pinv);
                        }
                    }
                    break;
                case OP_ST:
                    op_st<mode_signed(MODE)>(lanes, ctxs, op);
                    break;
                case OP_STG:
                    op_stg(K, lanes, ctxs, op);
                    break;
                case OP_MACG:
                    RZK_EACH_LANE {
                        RZK_LANE;
                        const uint32_t *img = image_row(K, K.st[op.b], ctx, L);
                        if (op.a == 0) mac_var(L.acc0, L.cur, img, t, op.c, L.pc.p, L.pc.p2, L.pc.pinv);
                        else mac_var_smem(ctx.acc1, L.cur, img, t, op.c, L.pc.p, L.pc.p2, L.pc.pinv);
                    }
                    break;
                case OP_LD:
                    op_ld(lanes, ctxs, op);
                    break;
                case OP_INV:
#if defined(__CUDA_ARCH__)
                    cta_lockstep(K);
#endif
                    q = op_inv<NP, MODE>(K, lanes, ctxs, q, it, prime_iter);
                    continue;
                case OP_LOOP:
                    loop_start = q + 1; loop_cnt = op.off; it = 0;
                    break;
                case OP_ENDLOOP:
                    if constexpr (MODE == MODE_SEQ_S) { if ((it & 31) == 31) reduce_accumulators_s(lanes, ctxs, ops_use_acc1(K.ops)); }
                    if (++it < loop_cnt) { q = loop_start; continue; }
                    it = 0;
                    break;
                default:
                    break;
                }
                ++q;
            }
            seg_end = q;
        }
        pc = seg_end;
    }
    fold_flags<RED_N>(K, lanes, ctxs);
}

// ---------------------------------------------------------------- compile-time programs
//
// The hot programs are also instantiated as templates: SP::prog is a constexpr program (built by the
// same builders), so op decoding, flag tests, dtype branches and the epilogue op lists fold away at
// compile time and the kernel body is straight-line code.  K.loop_count replaces the immediate of
// OP_LOOP (the number of Sum-proof terms is only known at launch).

// how many plain-term rows a compile-time program fetches ahead of each inverse transform (SP::kPreload, default 0)
template <class SP, class = void>
struct SpPreload { static constexpr int value = 0; };
template <class SP>
struct SpPreload<SP, std::void_t<decltype(SP::kPreload)>> { static constexpr int value = SP::kPreload; };

constexpr bool sp_uses_acc1(const Prog &p)
{
    for (int i = 0; i < kMaxOps && p.ops[i].code != OP_END; ++i)
        if ((p.ops[i].code == OP_MACV || p.ops[i].code == OP_MACK || p.ops[i].code == OP_INV) && p.ops[i].a == 1) return true;
    return false;
}

constexpr int sp_find_endloop(const Prog &p, int pc)
{
    while (p.ops[pc].code != OP_ENDLOOP) ++pc;
    return pc;
}

constexpr int sp_next_seg(const Prog &p, int pc)        // index of the next OP_SEG / OP_END at or after pc
{
    while (p.ops[pc].code != OP_SEG && p.ops[pc].code != OP_END) ++pc;
    return pc;
}

template <class SP, int PC>
RZK_VM void sp_epilogue_chunk(const VmLaunch &K, Lane *lanes, const LaneCtx *ctxs, int64_t (&V)[RZK_NL][kEpiChunk], int it, int j)
{
    constexpr Op e = SP::prog.ops[PC];
    if constexpr (e.code == OP_ADDP) {
        constexpr uint32_t dt = SP::dtype[e.a];
        op_addp_chunk(K, ctxs, V, e, it, dt, j);
        sp_epilogue_chunk<SP, PC + 1>(K, lanes, ctxs, V, it, j);
    } else if constexpr (e.code == OP_FIN) {
        op_fin_chunk(K, lanes, ctxs, V, e, it, j);
        sp_epilogue_chunk<SP, PC + 1>(K, lanes, ctxs, V, it, j);
    }
}

template <class SP, int MODE, int PC>
RZK_VM void sp_epilogue(const VmLaunch &K, Lane *lanes, const LaneCtx *ctxs, typename Epi<MODE>::V_t (&V)[RZK_NL][Epi<MODE>::kCount], int it)
{
    constexpr Op e = SP::prog.ops[PC];
    if constexpr (e.code == OP_ADDP) {
        constexpr uint32_t dt = SP::dtype[e.a];
        op_addp<MODE>(K, ctxs, V, e, it, dt);
        sp_epilogue<SP, MODE, PC + 1>(K, lanes, ctxs, V, it);
    } else if constexpr (e.code == OP_ROT) {
        op_rot<MODE>(K, lanes, ctxs, V, e, it);
        sp_epilogue<SP, MODE, PC + 1>(K, lanes, ctxs, V, it);
    } else if constexpr (e.code == OP_FIN) {
        op_fin<MODE>(K, lanes, ctxs, V, e, it);
        sp_epilogue<SP, MODE, PC + 1>(K, lanes, ctxs, V, it);
    }
}

// Plain terms of an epilogue, fetched before the inverse transform that precedes it (warp-per-item modes): at most kMaxPre
// OP_ADDP rows are held in registers across the transform, the others are loaded in place as before.
constexpr int kMaxPre = 2;

constexpr int sp_count_addp(const Prog &p, int pc)
{
    int n = 0;
    while (p.ops[pc].code == OP_ADDP || p.ops[pc].code == OP_FIN || p.ops[pc].code == OP_ROT) { if (p.ops[pc].code == OP_ADDP) ++n; ++pc; }
    return n;
}

template <class SP, int MODE, int PC, int IDX, int NPRE>
RZK_VM void sp_preload(const VmLaunch &K, const LaneCtx *ctxs, int32_t (&pre)[NPRE > 0 ? NPRE : 1][RZK_NL][Epi<MODE>::kCount], int it)
{
    constexpr Op e = SP::prog.ops[PC];
    if constexpr (e.code == OP_ADDP) {
        constexpr uint32_t dt = SP::dtype[e.a];
        if constexpr (IDX < NPRE) op_addp_load<MODE>(K, ctxs, pre[IDX], e, it, dt);
        sp_preload<SP, MODE, PC + 1, IDX + 1, NPRE>(K, ctxs, pre, it);
    } else if constexpr (e.code == OP_ROT || e.code == OP_FIN) {
        sp_preload<SP, MODE, PC + 1, IDX, NPRE>(K, ctxs, pre, it);
    }
}

template <class SP, int MODE, int PC, int IDX, int NPRE>
RZK_VM void sp_epilogue_pre(const VmLaunch &K, Lane *lanes, const LaneCtx *ctxs, typename Epi<MODE>::V_t (&V)[RZK_NL][Epi<MODE>::kCount],
                            const int32_t (&pre)[NPRE > 0 ? NPRE : 1][RZK_NL][Epi<MODE>::kCount], int it)
{
    constexpr Op e = SP::prog.ops[PC];
    if constexpr (e.code == OP_ADDP) {
        constexpr uint32_t dt = SP::dtype[e.a];
        if constexpr (IDX < NPRE) op_addp_apply<MODE>(V, pre[IDX], e);
        else op_addp<MODE>(K, ctxs, V, e, it, dt);
        sp_epilogue_pre<SP, MODE, PC + 1, IDX + 1, NPRE>(K, lanes, ctxs, V, pre, it);
    } else if constexpr (e.code == OP_ROT) {
        op_rot<MODE>(K, lanes, ctxs, V, e, it);
        sp_epilogue_pre<SP, MODE, PC + 1, IDX, NPRE>(K, lanes, ctxs, V, pre, it);
    } else if constexpr (e.code == OP_FIN) {
        op_fin<MODE>(K, lanes, ctxs, V, e, it);
        sp_epilogue_pre<SP, MODE, PC + 1, IDX, NPRE>(K, lanes, ctxs, V, pre, it);
    }
}

constexpr int sp_skip_epilogue(const Prog &p, int pc)
{
    while (p.ops[pc].code == OP_ADDP || p.ops[pc].code == OP_FIN || p.ops[pc].code == OP_ROT) ++pc;
    return pc;
}

// executes ops from PC up to (not including) the next OP_SEG / OP_END / OP_ENDLOOP
template <class SP, int NP, int MODE, int PC>
RZK_VM void sp_exec(const VmLaunch &K, Lane *lanes, const LaneCtx *ctxs, int it, int prime_iter)
{
    constexpr Op op = SP::prog.ops[PC];
    constexpr int KP = mode_sk(MODE) ? 2 * kKeyPolys : kKeyPolys;
    if constexpr (op.code == OP_SEG || op.code == OP_END || op.code == OP_ENDLOOP) {
        return;
    } else if constexpr (op.code == OP_LOOP) {
        constexpr int end = sp_find_endloop(SP::prog, PC);
        RZK_NOUNROLL
        for (int i = 0; i < (int)K.loop_count; ++i) {
            sp_exec<SP, NP, MODE, PC + 1>(K, lanes, ctxs, i, prime_iter);
            if constexpr (MODE == MODE_SEQ_S) { if ((i & 31) == 31) reduce_accumulators_s(lanes, ctxs, sp_uses_acc1(SP::prog)); }
        }
        sp_exec<SP, NP, MODE, end + 1>(K, lanes, ctxs, 0, prime_iter);
    } else if constexpr (op.code == OP_INV) {
        constexpr int next = sp_skip_epilogue(SP::prog, PC + 1);
        {
            typename Epi<MODE>::V_t V[RZK_NL][Epi<MODE>::kCount];
            if constexpr (mode_sk(MODE) && SP::prog.ops[PC + 1].code == OP_ADDP && SP::dtype[SP::prog.ops[PC + 1].a] == DT_I8 &&
                          !(SP::prog.ops[PC + 1].c & MAC_NEG)) {
                constexpr Op e0 = SP::prog.ops[PC + 1];
                if constexpr (MODE == MODE_SPLITKEY_S && RZK_SMALL_LATE) {
                    const Op e0c = e0;
                    inv_core<NP, MODE>(K, lanes, ctxs, op, prime_iter, V, nullptr, &e0c, it);
                } else {
                    int32_t small[RZK_NL][Epi<MODE>::kCount];
                    op_addp_load<MODE>(K, ctxs, small, e0, it, DT_I8);
                    inv_core<NP, MODE>(K, lanes, ctxs, op, prime_iter, V, small);
                }
                sp_epilogue<SP, MODE, PC + 2>(K, lanes, ctxs, V, it);
            } else if constexpr (!mode_seq(MODE) && SpPreload<SP>::value > 0) {
                constexpr int NPRE = sp_count_addp(SP::prog, PC + 1) < SpPreload<SP>::value ? sp_count_addp(SP::prog, PC + 1) : SpPreload<SP>::value;
                int32_t pre[NPRE > 0 ? NPRE : 1][RZK_NL][Epi<MODE>::kCount];
                sp_preload<SP, MODE, PC + 1, 0, NPRE>(K, ctxs, pre, it);
                inv_core<NP, MODE>(K, lanes, ctxs, op, prime_iter, V);
                sp_epilogue_pre<SP, MODE, PC + 1, 0, NPRE>(K, lanes, ctxs, V, pre, it);
            } else {
            inv_core<NP, MODE>(K, lanes, ctxs, op, prime_iter, V);
            if constexpr (ChunkedEpi<NP, MODE>::value) {
                if (prime_iter == NP - 1) {
                    RZK_UNROLL
                    for (int j = 0; j < kElems / kEpiChunk; ++j) {
                        int64_t V4[RZK_NL][kEpiChunk];
                        seq_chunk_values<NP>(K, lanes, ctxs, op, j, V4);
                        sp_epilogue_chunk<SP, PC + 1>(K, lanes, ctxs, V4, it, j);
                    }
                }
            } else if ((!mode_seq(MODE)) || prime_iter == NP - 1) sp_epilogue<SP, MODE, PC + 1>(K, lanes, ctxs, V, it);
            }
        }
        sp_exec<SP, NP, MODE, next>(K, lanes, ctxs, it, prime_iter);
    } else {
        if constexpr (op.code == OP_FWD) {
#if defined(__CUDA_ARCH__)
            // measured on B200: one barrier per forward transform (none before the inverses) is the best
            // trade between instruction-cache locality and pipe mixing; cta_sync >= 8 keeps only the
            // barrier at the start of each segment
            if (K.cta_sync < 8 || SP::prog.ops[PC - 1].code == OP_SEG) cta_lockstep(K);
#endif
            constexpr uint32_t dt = SP::dtype[op.a];
            op_fwd<mode_signed(MODE), MODE == MODE_SPLITKEY_S>(K, lanes, ctxs, op, it, dt);
        } else if constexpr (op.code == OP_MACK) {
            RZK_EACH_LANE {
                RZK_LANE;
                const int kidx = mode_sk(MODE) ? (2 * (int)op.b + ctx.hw) : (int)op.b;
                const uint32_t *krow = ctx.key + ((L.pi * KP + kidx) * 2) * kPadWords;
                if constexpr (MODE == MODE_SPLITKEY_S) {
                    static_assert(MODE != MODE_SPLITKEY_S || !(op.c & MAC_NEG), "the signed lazy key products have no negated form");
                    if constexpr (op.a == 0) mac_key_s(L.acc0, L.cur, krow, t, op.c, 0u - L.pc.p);
                    else mac_key_smem_s(ctx.acc1, L.cur, krow, t, op.c, 0u - L.pc.p);
                } else {
                    if constexpr (op.a == 0) mac_key(L.acc0, L.cur, krow, t, op.c, L.pc.p, L.pc.p2);
                    else mac_key_smem(ctx.acc1, L.cur, krow, t, op.c, L.pc.p, L.pc.p2);
                }
            }
        } else if constexpr (op.code == OP_MACV) {
            RZK_EACH_LANE {
                RZK_LANE;
                if constexpr (MODE == MODE_SEQ_S) {
                    if constexpr (op.a == 0) mac_var_s(L.acc0, L.cur, ctx.slot, t, op.c, L.pc.p, L.pc.pinv);
                    else mac_var_smem_s(ctx.acc1, L.cur, ctx.slot, t, op.c, L.pc.p, L.pc.pinv);
                } else {
                    if constexpr (op.a == 0) mac_var(L.acc0, L.cur, ctx.slot, t, op.c, L.pc.p, L.pc.p2, L.pc.pinv);
                    else mac_var_smem(ctx.acc1, L.cur, ctx.slot, t, op.c, L.pc.p, L.pc.p2, L.pc.pinv);
                }
            }
        } else if constexpr (op.code == OP_ST) {
            op_st<mode_signed(MODE)>(lanes, ctxs, op);
        } else if constexpr (op.code == OP_STG) {
            op_stg(K, lanes, ctxs, op);
        } else if constexpr (op.code == OP_MACG) {
            RZK_EACH_LANE {
                RZK_LANE;
                const uint32_t *img = image_row(K, K.st[op.b], ctx, L);
                if constexpr (op.a == 0) mac_var(L.acc0, L.cur, img, t, op.c, L.pc.p, L.pc.p2, L.pc.pinv);
                else mac_var_smem(ctx.acc1, L.cur, img, t, op.c, L.pc.p, L.pc.p2, L.pc.pinv);
            }
        } else if constexpr (op.code == OP_LD) {
            op_ld(lanes, ctxs, op);
        }
        sp_exec<SP, NP, MODE, PC + 1>(K, lanes, ctxs, it, prime_iter);
    }
}

template <class SP, int NP, int MODE, int PC>
RZK_VM void sp_norms(const VmLaunch &K, Lane *lanes, const LaneCtx *ctxs)
{
    constexpr Op op = SP::prog.ops[PC];
    if constexpr (op.code == OP_NORM) {
        constexpr uint32_t dt = SP::dtype[op.a];
        op_norm<MODE>(K, lanes, ctxs, op, dt);
        sp_norms<SP, NP, MODE, PC + 1>(K, lanes, ctxs);
    }
}

constexpr int sp_first_seg(const Prog &p)
{
    int pc = 0;
    while (p.ops[pc].code == OP_NORM) ++pc;
    return pc;
}

template <class SP, int NP, int MODE, int PC>
RZK_VM void sp_segments(const VmLaunch &K, Lane *lanes, const LaneCtx *ctxs)
{
    constexpr Op op = SP::prog.ops[PC];
    if constexpr (op.code == OP_SEG) {
        constexpr int PRIME_ITERS = (!mode_seq(MODE)) ? 1 : NP;
        RZK_NOUNROLL
        for (int prime_iter = 0; prime_iter < PRIME_ITERS; ++prime_iter) {
            RZK_EACH_LANE {
                RZK_LANE;
                L.pi = (MODE == MODE_SPLIT) ? ctx.hw : (mode_sk(MODE) ? 0 : prime_iter);
                L.pc = K.pc[L.pi];
                L.cap = kAddCap;
                if (MODE == MODE_SPLITKEY_S) {
                    L.pc.p = kStaticPrimeS; L.pc.p2 = 2u * kStaticPrimeS; L.pc.half = (kStaticPrimeS - 1u) / 2u;
                }
                if (MODE == MODE_SPLITKEY) {
                    // one fixed prime (slot 0, checked at key setup): its constants become immediates
                    L.pc.p = kStaticPrime0; L.pc.p2 = 2u * kStaticPrime0; L.pc.half = (kStaticPrime0 - 1u) / 2u;
                    L.cap = 4u * kStaticPrime0 - 1u;
                }
            }
            sp_exec<SP, NP, MODE, PC + 1>(K, lanes, ctxs, 0, prime_iter);
        }
        sp_segments<SP, NP, MODE, sp_next_seg(SP::prog, PC + 1)>(K, lanes, ctxs);
    }
}

// Compile-time counterpart of vm_run_item.
template <class SP>
RZK_VM void vm_run_static(const VmLaunch &K, Lane *lanes, const LaneCtx *ctxs)
{
    constexpr int NP = SP::kNP, MODE = SP::kMode;
    constexpr int RED_N = (!mode_seq(MODE)) ? 32 : 16;
    RZK_EACH_LANE { RZK_LANE; L.fail = 0; L.rerr = 0; }
    sp_norms<SP, NP, MODE, 0>(K, lanes, ctxs);
    sp_segments<SP, NP, MODE, sp_first_seg(SP::prog)>(K, lanes, ctxs);
    fold_flags<RED_N>(K, lanes, ctxs);
}

}  // namespace rzk
