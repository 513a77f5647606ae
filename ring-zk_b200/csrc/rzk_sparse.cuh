// rzk_sparse.cuh -- the response z = y + d*r (componentwise_mul(&d) then add,
// /root/reference/src/prove/open.rs:113-115, linear.rs:150-156, sum.rs:190-198) as signed rotations.
//
// The challenge d has kappa = 36 entries +-1 (challenge_space.rs:12-33) and r is tiny (|r| <= b), so the
// product needs no multiplication at all (SURVEY.md kernel K4): d*r = sum_k s_k * X^{pos_k} * r, and X^{pos} * r
// is r rotated by pos with the wrapped part negated.  One warp per item:
//   * r is written to shared memory as biased bytes u = r + bias in an extended array E+ = [bias - r | bias + r]
//     (1024 bytes: index 512 + i - pos reads the rotated, sign-corrected coefficient) and its mirror
//     E- = [bias + r | bias - r] for the terms with s_k = -1;
//   * every lane accumulates 16 consecutive coefficients of each of the 3 polynomials as packed bytes
//     (4 words): per term 5 word loads, 4 funnel shifts (byte alignment of the rotation), 4 adds.  The biased
//     bytes never carry: nnz * 2 * bias <= 255;
//   * z = y + (byte - bias * nnz), 128-bit global loads and stores.
// The words of E are spread over four planes (word L -> plane L & 3, index L >> 2) so that the lanes' accesses,
// which are 4 words apart, are bank-conflict free.
// Items outside the byte range (|r| > bias, an entry of d outside {-1,0,1}, more than 127 non-zeros) are not
// written; need[item] is set and the NTT program redoes exactly those items (rzk_engine.cu dev_respond).
// Compiled by nvcc (one lane per thread) and by g++ for the host lane emulator, like rzk_vm_exec.cuh.
#pragma once
#include "rzk_vm_exec.cuh"

namespace rzk {

constexpr int kSpPlane = 65;                    // words per plane: 64 + 1 (the word after the end is read at pos = 0)
constexpr int kSpArray = 4 * kSpPlane;          // one extended array (1024 bytes + padding)
constexpr int kSpListWords = 64;                // 128 u16 entries
constexpr int kSpWarpWords = 3 * 2 * kSpArray + kSpListWords;    // 1624 words per warp
constexpr uint32_t kSpMaxNnz = 127;

struct SparseLaunch {
    const int32_t *y;          // [items][3][512]
    const int8_t *r;           // [items][3][512]
    const int8_t *d;           // [items / d_div][512]
    int32_t *z;                // [items][3][512]
    uint32_t *need;            // [items]: 1 = redo with the NTT program, 0 = done here
    uint32_t *any_need;        // one word, set when any item needs the NTT program (zeroed before the launch)
    uint32_t n_items, d_div;
    uint32_t q, pad_;
};

struct LaneCtxS {
    uint32_t *sm;              // [kSpWarpWords] of this warp
    uint32_t item;
    int lane;
    bool active;
};

// byte-wise a + b (mod 256 per byte), no carries across bytes
RZK_HD uint32_t swar_add(uint32_t a, uint32_t b) { return ((a & 0x7f7f7f7fu) + (b & 0x7f7f7f7fu)) ^ ((a ^ b) & 0x80808080u); }
// byte-wise a - b
RZK_HD uint32_t swar_sub(uint32_t a, uint32_t b) { return ((a | 0x80808080u) - (b & 0x7f7f7f7fu)) ^ ((a ^ ~b) & 0x80808080u); }
// bit 7 of a byte set  <=>  that byte of u exceeds lim (lim < 0x7f)
RZK_HD uint32_t swar_gt(uint32_t u, uint32_t lim) { return (((u & 0x7f7f7f7fu) + (0x7fu - lim) * 0x01010101u) | u) & 0x80808080u; }

RZK_HD uint32_t funnel_r(uint32_t lo, uint32_t hi, uint32_t sh)
{
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, sh);
#else
    return sh ? (lo >> sh) | (hi << (32 - sh)) : lo;
#endif
}

RZK_HD uint4 sp_ld128(const void *p)
{
#if defined(__CUDA_ARCH__)
    return __ldg(reinterpret_cast<const uint4 *>(p));
#else
    return *reinterpret_cast<const uint4 *>(p);
#endif
}

// acc[j][c] += bytes [4c, 4c + 4) of the window that starts `sh` bits into word ptr[plane P0]; word P0 + c of the
// window lives in plane (P0 + c) & 3 at index (P0 + c) >> 2 past ptr
template <int P0>
RZK_VM void sp_accumulate(uint32_t (&acc)[3][4], const uint32_t *ptr, uint32_t sh)
{
    RZK_UNROLL
    for (int j = 0; j < 3; ++j) {
        const uint32_t *a = ptr + 2 * j * kSpArray;
        const uint32_t w0 = a[((P0 + 0) & 3) * kSpPlane + ((P0 + 0) >> 2)];
        const uint32_t w1 = a[((P0 + 1) & 3) * kSpPlane + ((P0 + 1) >> 2)];
        const uint32_t w2 = a[((P0 + 2) & 3) * kSpPlane + ((P0 + 2) >> 2)];
        const uint32_t w3 = a[((P0 + 3) & 3) * kSpPlane + ((P0 + 3) >> 2)];
        const uint32_t w4 = a[((P0 + 4) & 3) * kSpPlane + ((P0 + 4) >> 2)];
        acc[j][0] += funnel_r(w0, w1, sh);
        acc[j][1] += funnel_r(w1, w2, sh);
        acc[j][2] += funnel_r(w2, w3, sh);
        acc[j][3] += funnel_r(w3, w4, sh);
    }
}

RZK_VM void sparse_respond_item(const SparseLaunch &K, const LaneCtxS *ctxs)
{
    uint32_t nnz_l[RZK_NL], pre_l[RZK_NL], bad_l[RZK_NL];
    uint32_t dw[RZK_NL][4];
    // ---- scan d: positions and signs of the non-zero entries -> list in shared memory
    RZK_EACH_LANE {
        const LaneCtxS &ctx = ctxs[li_];
        const uint4 q = sp_ld128(K.d + (size_t)(ctx.item / K.d_div) * kN + 16 * ctx.lane);
        dw[li_][0] = q.x; dw[li_][1] = q.y; dw[li_][2] = q.z; dw[li_][3] = q.w;
        uint32_t cnt = 0, bad = 0;
        RZK_UNROLL
        for (int c = 0; c < 4; ++c) {
            const uint32_t t1 = swar_add(dw[li_][c], 0x01010101u);          // d + 1 per byte: valid entries give 0, 1, 2
            bad |= swar_gt(t1, 2u);
            const uint32_t nz = (((dw[li_][c] & 0x7f7f7f7fu) + 0x7f7f7f7fu) | dw[li_][c]) & 0x80808080u;   // bit 7: byte != 0
#if defined(__CUDA_ARCH__)
            cnt += (uint32_t)__popc(nz);
#else
            cnt += (uint32_t)__builtin_popcount(nz);
#endif
        }
        nnz_l[li_] = cnt; bad_l[li_] = bad;
    }
    // exclusive prefix sum of the per-lane counts, total, and OR of the validity words
    uint32_t nnz = 0, bad_any = 0;
#if defined(__CUDA_ARCH__)
    {
        uint32_t v = nnz_l[0];
        RZK_UNROLL
        for (int s = 1; s < 32; s <<= 1) { const uint32_t o = __shfl_up_sync(0xffffffffu, v, s); if (ctxs[0].lane >= s) v += o; }
        pre_l[0] = v - nnz_l[0];
        nnz = __shfl_sync(0xffffffffu, v, 31);
        bad_any = __any_sync(0xffffffffu, bad_l[0] != 0) ? 1u : 0u;
    }
#else
    for (int li = 0; li < RZK_NL; ++li) { pre_l[li] = nnz; nnz += nnz_l[li]; bad_any |= bad_l[li] ? 1u : 0u; }
#endif
    RZK_EACH_LANE {
        const LaneCtxS &ctx = ctxs[li_];
        uint16_t *list = reinterpret_cast<uint16_t *>(ctx.sm + 3 * 2 * kSpArray);
        uint32_t idx = pre_l[li_];
        RZK_UNROLL
        for (int b = 0; b < 16; ++b) {
            const int32_t v = (int32_t)(int8_t)(dw[li_][b >> 2] >> (8 * (b & 3)));
            if (v != 0) {
                if (idx < 128u) list[idx] = (uint16_t)((16 * ctx.lane + b) | (v < 0 ? 0x8000 : 0));
                ++idx;
            }
        }
    }
    const uint32_t bias = (nnz <= 42u) ? 3u : 1u;
    // ---- r -> biased extended arrays (both signs), range check
    uint32_t rbad_l[RZK_NL];
    RZK_EACH_LANE {
        const LaneCtxS &ctx = ctxs[li_];
        const int l = ctx.lane;
        uint32_t rb = 0;
        RZK_UNROLL
        for (int j = 0; j < 3; ++j) {
            const uint4 q = sp_ld128(K.r + ((size_t)ctx.item * 3 + j) * kN + 16 * l);
            const uint32_t rw[4] = {q.x, q.y, q.z, q.w};
            uint32_t *ep = ctx.sm + (2 * j + 0) * kSpArray, *em = ctx.sm + (2 * j + 1) * kSpArray;
            RZK_UNROLL
            for (int c = 0; c < 4; ++c) {
                const uint32_t up = swar_add(rw[c], bias * 0x01010101u);        // bias + r
                const uint32_t un = swar_sub(bias * 0x01010101u, rw[c]);        // bias - r
                rb |= swar_gt(up, 2 * bias);
                ep[c * kSpPlane + 32 + l] = up; ep[c * kSpPlane + l] = un;      // E+ = [bias - r | bias + r]
                em[c * kSpPlane + 32 + l] = un; em[c * kSpPlane + l] = up;      // E- = [bias + r | bias - r]
            }
            if (l == 0) { ep[64] = 0; em[64] = 0; }                              // word 256 (plane 0, index 64): read, never used
        }
        rbad_l[li_] = rb;
    }
    uint32_t rbad_any = 0;
#if defined(__CUDA_ARCH__)
    rbad_any = __any_sync(0xffffffffu, rbad_l[0] != 0) ? 1u : 0u;
#else
    for (int li = 0; li < RZK_NL; ++li) rbad_any |= rbad_l[li] ? 1u : 0u;
#endif
    RZK_SYNC();
    const bool fallback = bad_any || rbad_any || nnz > kSpMaxNnz;
    RZK_EACH_LANE {
        const LaneCtxS &ctx = ctxs[li_];
        if (ctx.lane == 0 && ctx.active) {
            K.need[ctx.item] = fallback ? 1u : 0u;
            if (fallback) *K.any_need = 1u;
        }
    }
    RZK_SYNC();          // reconverge after the one-lane store: without it lane 0 and the other 31 lanes ran the whole
                         // accumulation loop as two separate passes (ncu: every later instruction executed twice per item)
    if (fallback) return;
    // y is fetched now so that its latency hides behind the accumulation loop
    const int32_t corr = -(int32_t)(bias * nnz);
    uint4 yq[RZK_NL][3][4];
    RZK_EACH_LANE {
        const LaneCtxS &ctx = ctxs[li_];
        RZK_UNROLL
        for (int j = 0; j < 3; ++j) {
            const size_t row = ((size_t)ctx.item * 3 + j) * kN + 16 * ctx.lane;
            RZK_UNROLL
            for (int c = 0; c < 4; ++c) yq[li_][j][c] = sp_ld128(K.y + row + 4 * c);      // first use is after the loop
        }
    }
    // ---- accumulate the nnz signed rotations as packed biased bytes
    uint32_t acc[RZK_NL][3][4];
    RZK_EACH_LANE {
        RZK_UNROLL
        for (int j = 0; j < 3; ++j) { acc[li_][j][0] = 0; acc[li_][j][1] = 0; acc[li_][j][2] = 0; acc[li_][j][3] = 0; }
    }
    RZK_NOUNROLL
    for (uint32_t k = 0; k < nnz; ++k) {
        RZK_EACH_LANE {
            const LaneCtxS &ctx = ctxs[li_];
            const uint16_t *list = reinterpret_cast<const uint16_t *>(ctx.sm + 3 * 2 * kSpArray);
            const uint32_t e = list[k];
            const uint32_t o = 512u - (e & 0x1FFu);            // byte offset of coefficient 0 in the extended array
            const uint32_t s = o >> 2, sh = (o & 3u) * 8u;
            // the five words of a rotated 16-byte window start in plane s & 3: four code variants with immediate offsets
            const uint32_t *ptr = ctx.sm + ((e >> 15) ? kSpArray : 0) + ctx.lane + (s >> 2);
            switch (s & 3u) {
            case 0: sp_accumulate<0>(acc[li_], ptr, sh); break;
            case 1: sp_accumulate<1>(acc[li_], ptr, sh); break;
            case 2: sp_accumulate<2>(acc[li_], ptr, sh); break;
            default: sp_accumulate<3>(acc[li_], ptr, sh); break;
            }
        }
    }
    // ---- z = y + (byte - bias * nnz), canonical centred residue mod q (what Polynomial + yields).
    // Honest y is tiny (|y| < 2^30): then y + delta is canonical as it stands; any other int32 representative
    // takes the exact 64-bit path below (warp-uniform choice).
    uint32_t wide_any = 0;
    {
        uint32_t wide_l[RZK_NL];
        RZK_EACH_LANE {
            uint32_t wide = 0;
            RZK_UNROLL
            for (int j = 0; j < 3; ++j) {
                RZK_UNROLL
                for (int c = 0; c < 4; ++c) {
                    const uint4 q = yq[li_][j][c];
                    wide |= (q.x + 0x40000000u) | (q.y + 0x40000000u) | (q.z + 0x40000000u) | (q.w + 0x40000000u);   // bit 31 <=> y outside [-2^30, 2^30)
                }
            }
            wide_l[li_] = wide >> 31;
        }
#if defined(__CUDA_ARCH__)
        wide_any = __any_sync(0xffffffffu, wide_l[0] != 0) ? 1u : 0u;
#else
        for (int li = 0; li < RZK_NL; ++li) wide_any |= wide_l[li];
#endif
    }
    RZK_EACH_LANE {
        const LaneCtxS &ctx = ctxs[li_];
        RZK_UNROLL
        for (int j = 0; j < 3; ++j) {
            const size_t row = ((size_t)ctx.item * 3 + j) * kN + 16 * ctx.lane;
            RZK_UNROLL
            for (int c = 0; c < 4; ++c) {
                const uint4 q = yq[li_][j][c];
                const uint32_t w = acc[li_][j][c];
                const int32_t dl[4] = {(int32_t)(w & 0xffu) + corr, (int32_t)((w >> 8) & 0xffu) + corr,
                                       (int32_t)((w >> 16) & 0xffu) + corr, (int32_t)(w >> 24) + corr};
                const int32_t yy[4] = {(int32_t)q.x, (int32_t)q.y, (int32_t)q.z, (int32_t)q.w};
                int32_t zz[4];
                if (!wide_any) {
                    RZK_UNROLL
                    for (int b = 0; b < 4; ++b) zz[b] = yy[b] + dl[b];
                } else {
                    const int64_t half = (int64_t)((K.q - 1u) >> 1);
                    RZK_UNROLL
                    for (int b = 0; b < 4; ++b) {
                        int64_t v = (int64_t)yy[b] + (int64_t)dl[b];
                        if (v > half) v -= (int64_t)K.q;
                        else if (v < -half) v += (int64_t)K.q;
                        zz[b] = (int32_t)v;
                    }
                }
                uint4 zq;
                zq.x = (uint32_t)zz[0]; zq.y = (uint32_t)zz[1]; zq.z = (uint32_t)zz[2]; zq.w = (uint32_t)zz[3];
                if (ctx.active) *reinterpret_cast<uint4 *>(K.z + row + 4 * c) = zq;
            }
        }
    }
    RZK_SYNC();          // the shared-memory arrays are free for the next item
}

}  // namespace rzk
