// rzk_f64.cuh -- commitments on the FP64 pipe: the same negacyclic NTT / pointwise / inverse NTT
// structure as rzk_vm_exec.cuh, but modulo ONE 46-bit prime with every residue held exactly in a
// double and every modular product done with fused multiply-adds (error-free product + Barrett
// quotient by a precomputed w/p).  B200 issues DFMA at the full 64 lanes/clk/SM on a pipe of its own
// (measured: tools/imad_bench.cu, profiles/r2_imad_bench.jsonl), idle while the integer kernels run;
// a 46-bit modulus is wide enough for c = A.r + [0; x] with |r| <= 15 (|A.r| < 2^44.7 < p/2), so one
// commitment costs 4 transforms here instead of the 6 of the split-key integer program.
//
// All arithmetic is exact integer arithmetic carried in binary64:
//   mulmod(x, w):  h = x*w (rounded), l = fma(x, w, -h)            => h + l == x*w exactly
//                  k = rint(x * (w/p))   (magic-constant rounding, |x| <= 2^51, |w| <= p/2)
//                  r = fma(-k, p, h) + l                            => r == x*w - k*p exactly,
//                                                                      |r| <= p*(1/2 + |x|*2^-54)
//   butterflies are lazy: no correction in the forward transform (|values| < 5p), one reduction of
//   the running sums in the middle of the inverse transform (|values| < 2^51 throughout).
// The same source is compiled by nvcc (one lane per thread) and by g++ for the host lane emulator
// (tests/cpp/emu_check.cpp), like rzk_vm_exec.cuh.
//
// Replaces CommitmentKey::commit's `a.dot(&r).add(&z)` (/root/reference/src/commit.rs:109-125).
#pragma once
#include <math.h>
#include "rzk_vm_exec.cuh"

namespace rzk {

constexpr int64_t kF64Prime = 70368744137729LL;     // largest prime < 2^46 with p == 1 (mod 1024)
constexpr double kF64P = 70368744137729.0;
constexpr double kF64Magic = 6755399441055744.0;    // 1.5 * 2^52: (x + M) - M == rint(x) for |x| < 2^51
constexpr double kF64Cvt = 4503601774854144.0;      // 2^52 + 2^31: bit pattern (0x43300000, v ^ 0x80000000) minus this == (double)v
constexpr int kF64BufD = 544;                       // transpose buffer per half warp: 512 doubles + 2 per row of 32
constexpr int kF64G2Stride = 31;                   // double2 per (direction, lane): 30 twiddle pairs + 1 pad (31*4 words = 28 mod 32: conflict-free LDS.128)
constexpr int kF64KeyImages = 4;                    // a11, a12, a22, zero
constexpr uint32_t kF64SmallLimit = 15;             // |r| bound: 2*512*((q-1)/2)*15 < p/2

#if !defined(__CUDACC__)
struct alignas(16) double2 { double x, y; };
#endif

struct F64Launch {
    const int32_t *x;          // [B][512]
    const int8_t *r;           // [B][3][512]
    int32_t *c;                // [B][2][512]
    uint32_t *flags;           // [B / flag_div]
    const double *g1;          // device [2][32][2]   (w, w/p), forward then inverse
    const double *g2;          // device [2][16][31][2]
    const double *key;         // device [4][512][2]  lane-private order, pre-scaled by N^-1
    double q, qinv, pinv;
    uint32_t n_items, flag_div, small_lim, pad_;
    uint32_t *work;            // hybrid kernel: item counter
};

struct LaneF {
    double cur[kElems];
    uint32_t rerr;
};

struct LaneCtxF {
    double *buf;               // [kF64BufD] transpose buffer / exchange slot of this half warp
    const double *buf_partner; // the other half warp's
    const double2 *g1;         // [2][32]
    const double2 *g2;         // [2][16][kF64G2Stride]
    const double2 *key;        // [4][512]
    uint32_t item;
    int t, hw, ridx;
    bool active;
};

RZK_HD int32_t f64_ld8(const int8_t *p)
{
#if defined(__CUDA_ARCH__)
    return (int32_t)__ldg(p);
#else
    return (int32_t)*p;
#endif
}
RZK_HD int32_t f64_ld32(const int32_t *p)
{
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}

RZK_HD double f64_mul(double a, double b)
{
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    return a * b;
#endif
}
RZK_HD double f64_add(double a, double b)
{
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
}
RZK_HD double f64_fma(double a, double b, double c)
{
#if defined(__CUDA_ARCH__)
    return __fma_rn(a, b, c);
#else
    return fma(a, b, c);
#endif
}

// exact conversion of an int32 without the conversion pipe
RZK_HD double f64_from_i32(int32_t v)
{
#if defined(__CUDA_ARCH__)
    return __dadd_rn(__hiloint2double(0x43300000, (int)((uint32_t)v ^ 0x80000000u)), -kF64Cvt);
#else
    return (double)v;
#endif
}

// integer-valued double with |v| < 2^31 -> int32
RZK_HD int32_t f64_to_i32(double v)
{
#if defined(__CUDA_ARCH__)
    return __double2loint(__dadd_rn(v, kF64Magic));
#else
    return (int32_t)(int64_t)v;
#endif
}

// x * w - rint(x * w / p) * p, exactly; |result| <= p * (1/2 + |x| * 2^-54)
RZK_HD double f64_mulmod(double x, double w, double winv)
{
    const double h = f64_mul(x, w);
    const double l = f64_fma(x, w, -h);
    const double k = f64_add(f64_fma(x, winv, kF64Magic), -kF64Magic);
    const double r = f64_fma(-k, kF64P, h);
    return f64_add(r, l);
}

// x - rint(x / p) * p
RZK_HD double f64_reduce(double x, double pinv)
{
    const double k = f64_add(f64_fma(x, pinv, kF64Magic), -kF64Magic);
    return f64_fma(-k, kF64P, x);
}

RZK_HD void f64_ct(double &x, double &y, double w, double winv)
{
    const double t = f64_mulmod(y, w, winv);
    y = f64_add(x, -t);
    x = f64_add(x, t);
}

RZK_HD void f64_gs(double &x, double &y, double w, double winv)
{
    const double s = f64_add(x, y);
    const double d = f64_add(x, -y);
    x = s;
    y = f64_mulmod(d, w, winv);
}

template <int S, int DIR>
RZK_VM void f64_g1_stage(double (&a)[kElems], const double2 *g1)
{
    constexpr int half = 16 >> S;
    RZK_UNROLL
    for (int b = 0; b < (1 << S); ++b) {
        // (tried: these lane-uniform twiddles from constant memory -- LDC.64 instead of LDS.128: +8 % for this
        //  program alone, -4 % inside the hybrid launch, where the integer group's parameter reads share the constant cache)
        const double2 w = g1[(1 << S) + b];
        RZK_UNROLL
        for (int j = 0; j < half; ++j) {
            const int i0 = b * 2 * half + j;
            if (DIR == 0) f64_ct(a[i0], a[i0 + half], w.x, w.y);
            else f64_gs(a[i0], a[i0 + half], w.x, w.y);
        }
    }
}

template <int S, int DIR>
RZK_VM void f64_g2_stage(double (&a)[kElems], const double2 *tw)
{
    constexpr int half = 256 >> S;                // 8,4,2,1
    constexpr int nb = 1 << (S - 4);              // 2,4,8,16 blocks
    constexpr int base = (S == 5) ? 0 : (S == 6) ? 2 : (S == 7) ? 6 : 14;
    RZK_UNROLL
    for (int b = 0; b < nb; ++b) {
        const double2 w = tw[base + b];
        RZK_UNROLL
        for (int j = 0; j < half; ++j) {
            const int i0 = b * 2 * half + j;
            if (DIR == 0) f64_ct(a[i0], a[i0 + half], w.x, w.y);
            else f64_gs(a[i0], a[i0 + half], w.x, w.y);
        }
    }
}

// transpose buffer index (in doubles) of position i: rows of 32 padded by 2 -> conflict-free 16-byte reads
RZK_HD int f64_pad(int i) { return i + ((i >> 5) << 1); }

// One commitment per warp: half warp h transforms r[1 + h]; half warp 0 finishes c1 = r0 + a11*r1 + a12*r2,
// half warp 1 finishes c2 = r1 + a22*r2 + x  (commit.rs:123-125 with the key of commit.rs:38-57).
RZK_VM void f64_commit_item(const F64Launch &K, LaneF *lanes, const LaneCtxF *ctxs)
{
    // ---- load r[1 + hw] in the strided layout, forward stages 0..4
    RZK_EACH_LANE {
        LaneF &L = lanes[li_]; const LaneCtxF &ctx = ctxs[li_]; const int t = ctx.t;
        const int8_t *src = K.r + ((size_t)ctx.item * 3 + 1 + ctx.hw) * kN;
        int32_t v[kElems];
        RZK_UNROLL
        for (int m = 0; m < kElems; ++m) v[m] = f64_ld8(src + t + kLanes * m);
        uint32_t bad = 0;
        RZK_UNROLL
        for (int m = 0; m < kElems; ++m) {
            const uint32_t av = (uint32_t)(v[m] < 0 ? -v[m] : v[m]);
            bad |= (av > K.small_lim) ? 1u : 0u;
        }
        L.rerr = bad;
        RZK_UNROLL
        for (int m = 0; m < kElems; ++m) L.cur[m] = f64_from_i32(v[m]);
        f64_g1_stage<0, 0>(L.cur, ctx.g1);
        f64_g1_stage<1, 0>(L.cur, ctx.g1);
        f64_g1_stage<2, 0>(L.cur, ctx.g1);
        f64_g1_stage<3, 0>(L.cur, ctx.g1);
        f64_g1_stage<4, 0>(L.cur, ctx.g1);
        RZK_UNROLL
        for (int m = 0; m < kElems; ++m) ctx.buf[f64_pad(t + kLanes * m)] = L.cur[m];
    }
    RZK_SYNC();
    // ---- contiguous layout, forward stages 5..8
    RZK_EACH_LANE {
        LaneF &L = lanes[li_]; const LaneCtxF &ctx = ctxs[li_]; const int t = ctx.t;
        const double2 *row = reinterpret_cast<const double2 *>(ctx.buf + 34 * t);
        RZK_UNROLL
        for (int j = 0; j < 16; ++j) { const double2 q = row[j]; L.cur[2 * j] = q.x; L.cur[2 * j + 1] = q.y; }
        const double2 *tw = ctx.g2 + (0 * kLanes + t) * kF64G2Stride;
        f64_g2_stage<5, 0>(L.cur, tw);
        f64_g2_stage<6, 0>(L.cur, tw);
        f64_g2_stage<7, 0>(L.cur, tw);
        f64_g2_stage<8, 0>(L.cur, tw);
    }
    RZK_SYNC();          // every lane has read its row: the buffer becomes the exchange slot
    RZK_EACH_LANE {
        LaneF &L = lanes[li_]; const LaneCtxF &ctx = ctxs[li_]; const int t = ctx.t;
        double2 *s2 = reinterpret_cast<double2 *>(ctx.buf);
        RZK_UNROLL
        for (int j = 0; j < 16; ++j) { double2 q; q.x = L.cur[2 * j]; q.y = L.cur[2 * j + 1]; s2[j * kLanes + t] = q; }
    }
    RZK_SYNC();
    // ---- pointwise: half warp 0: a11*R1 + a12*R2, half warp 1: a22*R2 + 0*R1  (images pre-scaled by N^-1)
    RZK_EACH_LANE {
        LaneF &L = lanes[li_]; const LaneCtxF &ctx = ctxs[li_]; const int t = ctx.t;
        const double2 *k0 = ctx.key + (ctx.hw ? 2 : 0) * kN;
        const double2 *k1 = ctx.key + (ctx.hw ? 3 : 1) * kN;
        const double2 *o2 = reinterpret_cast<const double2 *>(ctx.buf_partner);
        RZK_UNROLL
        for (int j = 0; j < 16; ++j) {
            const double2 o = o2[j * kLanes + t];
            const double2 ka = k0[(2 * j) * kLanes + t], kb = k0[(2 * j + 1) * kLanes + t];
            const double2 kc = k1[(2 * j) * kLanes + t], kd = k1[(2 * j + 1) * kLanes + t];
            L.cur[2 * j] = f64_add(f64_mulmod(L.cur[2 * j], ka.x, ka.y), f64_mulmod(o.x, kc.x, kc.y));
            L.cur[2 * j + 1] = f64_add(f64_mulmod(L.cur[2 * j + 1], kb.x, kb.y), f64_mulmod(o.y, kd.x, kd.y));
        }
    }
    RZK_SYNC();          // the partner has read this half warp's slot
    // ---- inverse stages 8..5, one reduction of the running sums, transpose, inverse stages 4..0
    RZK_EACH_LANE {
        LaneF &L = lanes[li_]; const LaneCtxF &ctx = ctxs[li_]; const int t = ctx.t;
        const double2 *tw = ctx.g2 + (1 * kLanes + t) * kF64G2Stride;
        f64_g2_stage<8, 1>(L.cur, tw);
        f64_g2_stage<7, 1>(L.cur, tw);
        f64_g2_stage<6, 1>(L.cur, tw);
        f64_g2_stage<5, 1>(L.cur, tw);
        RZK_UNROLL
        for (int e = 0; e < kElems; ++e)
            if ((e & 8) == 0) L.cur[e] = f64_reduce(L.cur[e], K.pinv);     // sums of the last stage: < 17p -> < 0.51p
        double2 *row = reinterpret_cast<double2 *>(ctx.buf + 34 * t);
        RZK_UNROLL
        for (int j = 0; j < 16; ++j) { double2 q; q.x = L.cur[2 * j]; q.y = L.cur[2 * j + 1]; row[j] = q; }
    }
    RZK_SYNC();
    RZK_EACH_LANE {
        LaneF &L = lanes[li_]; const LaneCtxF &ctx = ctxs[li_]; const int t = ctx.t;
        RZK_UNROLL
        for (int m = 0; m < kElems; ++m) L.cur[m] = ctx.buf[f64_pad(t + kLanes * m)];
        f64_g1_stage<4, 1>(L.cur, ctx.g1 + 32);
        f64_g1_stage<3, 1>(L.cur, ctx.g1 + 32);
        f64_g1_stage<2, 1>(L.cur, ctx.g1 + 32);
        f64_g1_stage<1, 1>(L.cur, ctx.g1 + 32);
        f64_g1_stage<0, 1>(L.cur, ctx.g1 + 32);
        // ---- epilogue: exact integer V = centred residue mod p; add the plain terms; centred residue mod q
        const int8_t *ra = K.r + ((size_t)ctx.item * 3 + ctx.hw) * kN;        // r0 for c1, r1 for c2
        const int32_t *xs = K.x + (size_t)ctx.item * kN;
        int32_t *dst = K.c + ((size_t)ctx.item * 2 + ctx.hw) * kN;
        // addends are fetched eight at a time through the read-only path (they may be hoisted over the stores)
        RZK_UNROLL
        for (int m0 = 0; m0 < kElems; m0 += 8) {
            int32_t ar[8], ax[8];
            RZK_UNROLL
            for (int k = 0; k < 8; ++k) ar[k] = f64_ld8(ra + t + kLanes * (m0 + k));
            RZK_UNROLL
            for (int k = 0; k < 8; ++k) ax[k] = ctx.hw ? f64_ld32(xs + t + kLanes * (m0 + k)) : 0;
            RZK_UNROLL
            for (int k = 0; k < 8; ++k) {
                const int m = m0 + k, i = t + kLanes * m;
                double v = f64_reduce(L.cur[m], K.pinv);
                v = f64_add(v, f64_from_i32(ar[k]));
                v = f64_add(v, f64_from_i32(ax[k]));
                const double kq = f64_add(f64_fma(v, K.qinv, kF64Magic), -kF64Magic);
                const double rem = f64_fma(-kq, K.q, v);                     // |rem| <= (q-1)/2 exactly
                if (ctx.active) dst[i] = f64_to_i32(rem);
            }
        }
    }
    RZK_SYNC();          // the transpose buffer is free for the next item
    // ---- range flag (|r| > small_lim on a transformed row)
#if defined(__CUDA_ARCH__)
    {
        uint32_t f = lanes[0].rerr << 1;
        RZK_UNROLL
        for (int d = 16; d >= 1; d >>= 1) f |= __shfl_xor_sync(0xffffffffu, f, d);
        if (ctxs[0].ridx == 0 && ctxs[0].active && f) atomicOr(&K.flags[ctxs[0].item / K.flag_div], f);
    }
#else
    {
        uint32_t f = 0;
        for (int li = 0; li < RZK_NL; ++li) f |= lanes[li].rerr << 1;
        if (ctxs[0].active && f) K.flags[ctxs[0].item / K.flag_div] |= f;
    }
#endif
}

}  // namespace rzk
