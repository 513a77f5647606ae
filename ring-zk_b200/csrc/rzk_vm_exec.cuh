// rzk_vm_exec.cuh -- per-lane implementation of the polynomial-op program (rzk_vm.h).
//
// The same source is compiled twice:
//   * by nvcc for sm_100a, where RZK_NL == 1 and "each lane" is the calling thread
//     (16 lanes = one half warp = one batch item, 32 coefficients per lane in registers);
//   * by g++ for the host lane emulator (tests/cpp/emu_check.cpp), where RZK_NL == 16
//     and the lane loop is explicit.  The emulator exists so that the exact kernel
//     arithmetic and shared-memory layouts are checked against the CPU oracle without
//     a GPU; it is test infrastructure, never a product fallback.
//
// Layouts (N = 512):
//   G1 (strided):    lane t holds coefficients i = t + 16*m,  m = 0..31   (stages 0..4: bits 8..4 are lane-local)
//   G2 (contiguous): lane t holds positions   i = 32*t + e,   e = 0..31   (stages 5..8: bits 3..0 are lane-local)
//   transpose buffer word of position i: i + 4*(i>>5)  (uint4 rows of 36 words -> conflict-free)
//   lane-private slot word of element e: ((e>>2)*16 + t)*4 + (e&3)
//
// Replaces Polynomial `*`/`+`/`-`/`==` as composed by Mat::dot, add, sub and
// componentwise_mul (/root/reference/src/mat.rs:95-178).
#pragma once
#include "rzk_arith.cuh"
#include "rzk_vm.h"

#if defined(__CUDACC__)
#define RZK_VM __device__ __forceinline__
#define RZK_NL 1
#define RZK_SYNC() __syncwarp()
#define RZK_UNROLL _Pragma("unroll")
#define RZK_NOUNROLL _Pragma("unroll 1")
#else
#define RZK_VM inline
#define RZK_NL 16
#define RZK_SYNC() ((void)0)
#define RZK_UNROLL
#define RZK_NOUNROLL
struct uint4 { uint32_t x, y, z, w; };
#endif

#if defined(__CUDACC__)
// one lane per thread: no loop, no indexing, so the lane state stays in registers
#define RZK_EACH_LANE if (constexpr int li_ = 0; true)
#else
#define RZK_EACH_LANE for (int li_ = 0; li_ < RZK_NL; ++li_)
#endif
#define RZK_LANE Lane &L = lanes[li_]; const int t = t0 + li_; (void)t; (void)L

namespace rzk {

struct Lane {
    uint32_t cur[kElems];
    uint32_t acc0[kElems];   // accumulator 0 lives in registers; accumulator 1 in the lane-private smem slot ctx.acc1
    uint32_t fail;
    uint32_t rerr;
};

struct ItemCtx {
    uint32_t *buf;         // [kBufWords]  transpose buffer of this half warp
    uint32_t *slot;        // [kSlotWords] operand slot of OP_ST / OP_MACV
    uint32_t *acc1;        // [kSlotWords] accumulator 1 (lane-private layout)
    uint32_t *stash;       // [NSTASH][np-1][kSlotWords] residues of earlier primes
    const uint32_t *g2;    // staged [np][2][16][60]
    const uint32_t *key;   // staged [np][3][2][576]
    const uint32_t *g1;    // host emulator only: [slot][2][32][2]
    uint32_t item;         // item index (clamped to n_items-1 for inactive half warps)
    bool active;
};

#if defined(__CUDA_ARCH__)
#define RZK_G1(slot, dir, idx, j) c_g1[slot][dir][idx][j]
#else
#define RZK_G1(slot, dir, idx, j) ctx.g1[((((slot) * 2 + (dir)) * 32 + (idx)) * 2) + (j)]
#endif

RZK_VM int priv_index(int t, int e) { return (((e >> 2) * kLanes + t) << 2) + (e & 3); }

// ---------------------------------------------------------------- transforms

// One butterfly stage with compile-time geometry (all loops have constant trip counts so that
// they unroll fully and the 32 coefficients stay in registers).
//   G1 stages S = 0..4: distance 16>>S in the strided layout, lane-uniform twiddles (constant bank)
//   G2 stages S = 5..8: distance 256>>S in the contiguous layout, lane-specific twiddles (shared memory)
template <int S, int DIR>
RZK_VM void g1_stage(uint32_t (&a)[kElems], const ItemCtx &ctx, uint32_t slot, uint32_t p, uint32_t p2)
{
    (void)ctx;
    constexpr int half = 16 >> S;
    RZK_UNROLL
    for (int b = 0; b < (1 << S); ++b) {
        const uint32_t w = RZK_G1(slot, DIR, (1 << S) + b, 0);
        const uint32_t wp = RZK_G1(slot, DIR, (1 << S) + b, 1);
        RZK_UNROLL
        for (int j = 0; j < half; ++j) {
            const int i0 = b * 2 * half + j;
            if (DIR == 0) ct_bfly(a[i0], a[i0 + half], w, wp, p, p2);
            else gs_bfly(a[i0], a[i0 + half], w, wp, p, p2);
        }
    }
}

template <int S, int DIR>
RZK_VM void g2_stage(uint32_t (&a)[kElems], const uint4 *tw4, uint32_t p, uint32_t p2)
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
            if (DIR == 0) ct_bfly(a[i0], a[i0 + half], q.x, q.y, p, p2);
            else gs_bfly(a[i0], a[i0 + half], q.x, q.y, p, p2);
        }
        RZK_UNROLL
        for (int j = 0; j < half; ++j) {
            const int i0 = (2 * b2 + 1) * 2 * half + j;
            if (DIR == 0) ct_bfly(a[i0], a[i0 + half], q.z, q.w, p, p2);
            else gs_bfly(a[i0], a[i0 + half], q.z, q.w, p, p2);
        }
    }
}

RZK_VM void fwd_g1(uint32_t (&a)[kElems], const ItemCtx &ctx, uint32_t slot, uint32_t p, uint32_t p2)
{
    g1_stage<0, 0>(a, ctx, slot, p, p2);
    g1_stage<1, 0>(a, ctx, slot, p, p2);
    g1_stage<2, 0>(a, ctx, slot, p, p2);
    g1_stage<3, 0>(a, ctx, slot, p, p2);
    g1_stage<4, 0>(a, ctx, slot, p, p2);
}

RZK_VM void fwd_g2(uint32_t (&a)[kElems], const uint32_t *tw, uint32_t p, uint32_t p2)
{
    const uint4 *tw4 = reinterpret_cast<const uint4 *>(tw);
    g2_stage<5, 0>(a, tw4, p, p2);
    g2_stage<6, 0>(a, tw4, p, p2);
    g2_stage<7, 0>(a, tw4, p, p2);
    g2_stage<8, 0>(a, tw4, p, p2);
}

RZK_VM void inv_g2(uint32_t (&a)[kElems], const uint32_t *tw, uint32_t p, uint32_t p2)
{
    const uint4 *tw4 = reinterpret_cast<const uint4 *>(tw);
    g2_stage<8, 1>(a, tw4, p, p2);
    g2_stage<7, 1>(a, tw4, p, p2);
    g2_stage<6, 1>(a, tw4, p, p2);
    g2_stage<5, 1>(a, tw4, p, p2);
}

RZK_VM void inv_g1(uint32_t (&a)[kElems], const ItemCtx &ctx, uint32_t slot, uint32_t p, uint32_t p2)
{
    g1_stage<4, 1>(a, ctx, slot, p, p2);
    g1_stage<3, 1>(a, ctx, slot, p, p2);
    g1_stage<2, 1>(a, ctx, slot, p, p2);
    g1_stage<1, 1>(a, ctx, slot, p, p2);
    g1_stage<0, 1>(a, ctx, slot, p, p2);
}

// ---------------------------------------------------------------- global memory

RZK_VM uint64_t stream_poly(const Stream &s, uint32_t item, uint32_t off)
{
    return (uint64_t)(item / s.div) * s.stride + off;
}

// ---------------------------------------------------------------- ops

RZK_VM void op_fwd(const VmLaunch &K, const ItemCtx &ctx, Lane *lanes, int t0, const Op &op, int it, int pi)
{
    const PrimeC pc = K.pc[pi];
    const Stream st = K.st[op.a];
    const uint64_t poly = stream_poly(st, ctx.item, (uint32_t)op.off + (uint32_t)it * op.step);
    RZK_EACH_LANE {
        RZK_LANE;
        int32_t v[kElems];
        if (st.dtype == DT_I8) {
            const int8_t *src = reinterpret_cast<const int8_t *>(st.base) + poly * kN;
            RZK_UNROLL
            for (int m = 0; m < kElems; ++m) v[m] = src[t + kLanes * m];
        } else {
            const int32_t *src = reinterpret_cast<const int32_t *>(st.base) + poly * kN;
            RZK_UNROLL
            for (int m = 0; m < kElems; ++m) v[m] = canon_q(src[t + kLanes * m], K.q);
        }
        if (op.b & FWD_CHECK_SMALL) {
            uint32_t bad = 0;
            RZK_UNROLL
            for (int m = 0; m < kElems; ++m) {
                const uint32_t av = (uint32_t)(v[m] < 0 ? -v[m] : v[m]);
                bad |= (av > K.small_lim) ? 1u : 0u;
            }
            L.rerr |= bad;
        }
        // centred value + 2p lies in (0, 4p): a valid lazy input of the forward butterflies
        if (op.b & FWD_SCALED) {
            RZK_UNROLL
            for (int m = 0; m < kElems; ++m)
                L.cur[m] = shoup_mul(pc.rn, pc.rnp, (uint32_t)v[m] + pc.p2, pc.p);
        } else {
            RZK_UNROLL
            for (int m = 0; m < kElems; ++m) L.cur[m] = (uint32_t)v[m] + pc.p2;
        }
        fwd_g1(L.cur, ctx, pc.slot, pc.p, pc.p2);
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
        fwd_g2(L.cur, ctx.g2 + ((pi * 2 + 0) * kLanes + t) * kG2Words, pc.p, pc.p2);
    }
    RZK_SYNC();
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
            const uint32_t base = init ? 0u : acc[e];
            acc[e] = csub(base + tt, p2);
        }
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
            const uint32_t base = init ? 0u : acc[e];
            acc[e] = csub(base + tt, p2);
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
        tt = shoup_mul(w.x, wp.x, cur[4 * j + 0], p); if (neg) tt = p2 - tt; a.x = csub((init ? 0u : a.x) + tt, p2);
        tt = shoup_mul(w.y, wp.y, cur[4 * j + 1], p); if (neg) tt = p2 - tt; a.y = csub((init ? 0u : a.y) + tt, p2);
        tt = shoup_mul(w.z, wp.z, cur[4 * j + 2], p); if (neg) tt = p2 - tt; a.z = csub((init ? 0u : a.z) + tt, p2);
        tt = shoup_mul(w.w, wp.w, cur[4 * j + 3], p); if (neg) tt = p2 - tt; a.w = csub((init ? 0u : a.w) + tt, p2);
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
        tt = mont_mul(csub(cur[4 * j + 0], p2), s.x, p, pinv); if (neg) tt = p2 - tt; a.x = csub((init ? 0u : a.x) + tt, p2);
        tt = mont_mul(csub(cur[4 * j + 1], p2), s.y, p, pinv); if (neg) tt = p2 - tt; a.y = csub((init ? 0u : a.y) + tt, p2);
        tt = mont_mul(csub(cur[4 * j + 2], p2), s.z, p, pinv); if (neg) tt = p2 - tt; a.z = csub((init ? 0u : a.z) + tt, p2);
        tt = mont_mul(csub(cur[4 * j + 3], p2), s.w, p, pinv); if (neg) tt = p2 - tt; a.w = csub((init ? 0u : a.w) + tt, p2);
        a4[j * kLanes + t] = a;
    }
}

RZK_VM void op_st(const VmLaunch &K, const ItemCtx &ctx, Lane *lanes, int t0, int pi)
{
    const PrimeC pc = K.pc[pi];
    RZK_EACH_LANE {
        RZK_LANE;
        uint4 *s4 = reinterpret_cast<uint4 *>(ctx.slot);
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
    // lane-private: no cross-lane hazard, no sync needed
}

RZK_VM void op_addp(const VmLaunch &K, const ItemCtx &ctx, int64_t (&V)[RZK_NL][kElems], int t0, const Op &op, int it)
{
    const Stream st = K.st[op.a];
    const uint64_t poly = stream_poly(st, ctx.item, (uint32_t)op.off + (uint32_t)it * op.step);
    const bool neg = op.c & MAC_NEG;
    RZK_EACH_LANE {
        const int t = t0 + li_;
        int32_t v[kElems];
        if (st.dtype == DT_I8) {
            const int8_t *src = reinterpret_cast<const int8_t *>(st.base) + poly * kN;
            RZK_UNROLL
            for (int m = 0; m < kElems; ++m) v[m] = src[t + kLanes * m];
        } else {
            const int32_t *src = reinterpret_cast<const int32_t *>(st.base) + poly * kN;
            RZK_UNROLL
            for (int m = 0; m < kElems; ++m) v[m] = canon_q(src[t + kLanes * m], K.q);
        }
        RZK_UNROLL
        for (int m = 0; m < kElems; ++m) V[li_][m] += neg ? -(int64_t)v[m] : (int64_t)v[m];
    }
}

RZK_VM void op_fin(const VmLaunch &K, const ItemCtx &ctx, Lane *lanes, int64_t (&V)[RZK_NL][kElems], int t0, const Op &op, int it)
{
    const Stream st = K.st[op.a];
    const uint64_t poly = stream_poly(st, ctx.item, (uint32_t)op.off + (uint32_t)it * op.step);
    RZK_EACH_LANE {
        RZK_LANE;
        int32_t res[kElems];
        RZK_UNROLL
        for (int m = 0; m < kElems; ++m) res[m] = reduce_q_centered(V[li_][m], K.q, K.bar, K.kq);
        if (op.b & FIN_CMPZ) {
            uint32_t nz = 0;
            RZK_UNROLL
            for (int m = 0; m < kElems; ++m) nz |= (uint32_t)res[m];
            L.fail |= nz ? 1u : 0u;
        }
        if ((op.b & FIN_STORE) && ctx.active) {
            int32_t *dst = reinterpret_cast<int32_t *>(const_cast<void *>(st.base)) + poly * kN;
            RZK_UNROLL
            for (int m = 0; m < kElems; ++m) dst[t + kLanes * m] = res[m];
        }
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
    // a value congruent to V mod q that stays inside int64: v01 < 2^60, P01modq*h2 < 2^62
    int64_t w = (int64_t)(v01 + K.crt.P01modq * (uint64_t)h2);
    if (negv) w -= (int64_t)K.crt.Pmodq;
    return w;
}

// Inverse transform of acc[a].  On the last prime the residues of all primes are combined and
// the epilogue ops that follow (OP_ADDP*, OP_FIN) are executed here, so that the 64-bit
// values live only inside this function.  Returns the index of the first op after the epilogue.
template <int NP, int NSTASH>
RZK_VM int op_inv(const VmLaunch &K, const ItemCtx &ctx, Lane *lanes, int t0, int q, int it, int pi)
{
    const Op op = K.ops[q];
    int64_t V[RZK_NL][kElems];
    const PrimeC pc = K.pc[pi];
    RZK_EACH_LANE {
        RZK_LANE;
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
        inv_g2(L.cur, ctx.g2 + ((pi * 2 + 1) * kLanes + t) * kG2Words, pc.p, pc.p2);
        uint4 *row = reinterpret_cast<uint4 *>(ctx.buf + 36 * t);
        RZK_UNROLL
        for (int j = 0; j < 8; ++j) {
            uint4 q;
            q.x = L.cur[4 * j + 0]; q.y = L.cur[4 * j + 1]; q.z = L.cur[4 * j + 2]; q.w = L.cur[4 * j + 3];
            row[j] = q;
        }
    }
    RZK_SYNC();
    RZK_EACH_LANE {
        RZK_LANE;
        RZK_UNROLL
        for (int m = 0; m < kElems; ++m) {
            const int i = t + kLanes * m;
            L.cur[m] = ctx.buf[i + ((i >> 5) << 2)];
        }
        inv_g1(L.cur, ctx, pc.slot, pc.p, pc.p2);
        RZK_UNROLL
        for (int m = 0; m < kElems; ++m) L.cur[m] = csub(L.cur[m], pc.p);     // [0,2p) -> [0,p)
        if (pi < NP - 1) {
            uint4 *s4 = reinterpret_cast<uint4 *>(ctx.stash + ((int)op.b * (NP - 1) + pi) * kSlotWords);
            RZK_UNROLL
            for (int j = 0; j < 8; ++j) {
                uint4 q;
                q.x = L.cur[4 * j + 0]; q.y = L.cur[4 * j + 1]; q.z = L.cur[4 * j + 2]; q.w = L.cur[4 * j + 3];
                s4[j * kLanes + t] = q;
            }
        } else {
            uint32_t prev[2][kElems];
            RZK_UNROLL
            for (int k = 0; k < NP - 1; ++k) {
                const uint4 *s4 = reinterpret_cast<const uint4 *>(ctx.stash + ((int)op.b * (NP - 1) + k) * kSlotWords);
                RZK_UNROLL
                for (int j = 0; j < 8; ++j) {
                    const uint4 q = s4[j * kLanes + t];
                    prev[k][4 * j + 0] = q.x; prev[k][4 * j + 1] = q.y; prev[k][4 * j + 2] = q.z; prev[k][4 * j + 3] = q.w;
                }
            }
            RZK_UNROLL
            for (int m = 0; m < kElems; ++m) {
                uint32_t r[kMaxPrimes] = {0, 0, 0};
                RZK_UNROLL
                for (int k = 0; k < NP - 1; ++k) r[k] = prev[k][m];
                r[NP - 1] = L.cur[m];
                V[li_][m] = crt_combine<NP>(K, r);
            }
        }
    }
    RZK_SYNC();
    (void)NSTASH;
    const bool last = (pi == NP - 1);
    ++q;
    RZK_NOUNROLL
    for (;; ++q) {
        const Op e = K.ops[q];
        if (e.code == OP_ADDP) { if (last) op_addp(K, ctx, V, t0, e, it); }
        else if (e.code == OP_FIN) { if (last) op_fin(K, ctx, lanes, V, t0, e, it); }
        else break;
    }
    return q;
}

// params.rs:102-118 via polynomial.rs:60-73: floor(sqrt(sum c^2)) <= bound  <=>  sum c^2 < (bound+1)^2
RZK_VM void op_norm(const VmLaunch &K, const ItemCtx &ctx, Lane *lanes, int t0, const Op &op)
{
    const Stream st = K.st[op.a];
    const uint32_t abs_lim = K.norm_abs_lim[op.b];
    const uint64_t sq_lim = K.norm_sq_lim[op.b];
    RZK_NOUNROLL
    for (int c = 0; c < (int)op.c; ++c) {
        const uint64_t poly = stream_poly(st, ctx.item, (uint32_t)op.off + (uint32_t)c);
        RZK_EACH_LANE {
            RZK_LANE;
            uint64_t s = 0;
            uint32_t bad = 0;
            if (st.dtype == DT_I8) {
                const int8_t *src = reinterpret_cast<const int8_t *>(st.base) + poly * kN;
                RZK_UNROLL
                for (int m = 0; m < kElems; ++m) {
                    const int32_t v = src[t + kLanes * m];
                    s += (uint64_t)(uint32_t)(v * v);
                }
            } else {
                const int32_t *src = reinterpret_cast<const int32_t *>(st.base) + poly * kN;
                RZK_UNROLL
                for (int m = 0; m < kElems; ++m) {
                    const int32_t v = canon_q(src[t + kLanes * m], K.q);
                    const uint32_t av = (uint32_t)(v < 0 ? -v : v);
                    const bool big = av > abs_lim;
                    bad |= big ? 1u : 0u;
                    s += big ? 0ull : (uint64_t)av * (uint64_t)av;
                }
            }
            L.fail |= bad;
            ctx.buf[2 * t] = (uint32_t)s;
            ctx.buf[2 * t + 1] = (uint32_t)(s >> 32);
        }
        RZK_SYNC();
        RZK_EACH_LANE {
            RZK_LANE;
            uint64_t tot = 0;
            RZK_UNROLL
            for (int j = 0; j < kLanes; ++j) tot += (uint64_t)ctx.buf[2 * j] | ((uint64_t)ctx.buf[2 * j + 1] << 32);
            L.fail |= (tot > sq_lim) ? 1u : 0u;
        }
        RZK_SYNC();
    }
}

// ---------------------------------------------------------------- interpreter

template <int NP, int NSTASH>
RZK_VM void vm_run_item(const VmLaunch &K, const ItemCtx &ctx, Lane *lanes, int t0)
{
    RZK_EACH_LANE { RZK_LANE; L.fail = 0; L.rerr = 0; }
    int pc = 0;
    RZK_NOUNROLL
    while (K.ops[pc].code == OP_NORM) {
        op_norm(K, ctx, lanes, t0, K.ops[pc]);
        ++pc;
    }
    RZK_NOUNROLL
    while (K.ops[pc].code == OP_SEG) {
        const int seg_begin = pc + 1;
        int seg_end = seg_begin;
        RZK_NOUNROLL
        for (int pi = 0; pi < NP; ++pi) {
            const PrimeC pcst = K.pc[pi];
            int q = seg_begin, loop_start = 0, loop_cnt = 0, it = 0;
            RZK_NOUNROLL
            for (;;) {
                const Op op = K.ops[q];
                if (op.code == OP_SEG || op.code == OP_END) break;
                switch (op.code) {
                case OP_FWD:
                    op_fwd(K, ctx, lanes, t0, op, it, pi);
                    break;
                case OP_MACK: {
                    const uint32_t *krow = ctx.key + ((pi * kKeyPolys + (int)op.b) * 2) * kPadWords;
                    RZK_EACH_LANE {
                        RZK_LANE;
                        if (op.a == 0) mac_key(L.acc0, L.cur, krow, t, op.c, pcst.p, pcst.p2);
                        else mac_key_smem(ctx.acc1, L.cur, krow, t, op.c, pcst.p, pcst.p2);
                    }
                    break;
                }
                case OP_MACV:
                    RZK_EACH_LANE {
                        RZK_LANE;
                        if (op.a == 0) mac_var(L.acc0, L.cur, ctx.slot, t, op.c, pcst.p, pcst.p2, pcst.pinv);
                        else mac_var_smem(ctx.acc1, L.cur, ctx.slot, t, op.c, pcst.p, pcst.p2, pcst.pinv);
                    }
                    break;
                case OP_ST:
                    op_st(K, ctx, lanes, t0, pi);
                    break;
                case OP_INV:
                    q = op_inv<NP, NSTASH>(K, ctx, lanes, t0, q, it, pi);
                    continue;
                case OP_LOOP:
                    loop_start = q + 1; loop_cnt = op.off; it = 0;
                    break;
                case OP_ENDLOOP:
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
    // fold the 16 lanes' status words into the item-group flag word
    RZK_EACH_LANE { RZK_LANE; ctx.buf[t] = L.fail | (L.rerr << 1); }
    RZK_SYNC();
    RZK_EACH_LANE {
        RZK_LANE;
        if (t == 0 && ctx.active) {
            uint32_t f = 0;
            RZK_UNROLL
            for (int j = 0; j < kLanes; ++j) f |= ctx.buf[j];
            if (f) {
#if defined(__CUDA_ARCH__)
                atomicOr(&K.flags[ctx.item / K.flag_div], f);
#else
                K.flags[ctx.item / K.flag_div] |= f;
#endif
            }
        }
    }
    RZK_SYNC();
}

}  // namespace rzk
