// rzk_arith.cuh -- word-level modular arithmetic shared by the sm_100a kernels and
// the host-side lane emulator (tests/cpp/emu_check.cpp compiles this with g++).
//
// Replaces the per-coefficient arithmetic of Polynomial<ZqI64<Q>, N>'s `* + -`
// (crate poly-ring-xnp1; called from /root/reference/src/mat.rs:109-110,135-136,
// 160-161,176).  q = 3515337053 admits no length-512 NTT (q-1 = 2^2*2389*367867),
// so products are computed exactly over the integers through NTT-friendly
// auxiliary primes p < 2^30 (Harvey lazy butterflies, Shoup constants) and
// recombined by CRT before the centred reduction mod q.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define RZK_HD __host__ __device__ __forceinline__
#define RZK_D __device__ __forceinline__
#else
#define RZK_HD inline
#define RZK_D inline
#endif

namespace rzk {

RZK_HD uint32_t mulhi32(uint32_t a, uint32_t b)
{
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}

RZK_HD uint64_t mulhi64(uint64_t a, uint64_t b)
{
#if defined(__CUDA_ARCH__)
    return __umul64hi(a, b);
#else
    return (uint64_t)(((unsigned __int128)a * (unsigned __int128)b) >> 64);
#endif
}

// acc + sum of squares of the four int8 lanes of v
RZK_HD uint32_t dot4_i8(int32_t v, uint32_t acc)
{
#if defined(__CUDA_ARCH__)
    return (uint32_t)__dp4a(v, v, (int32_t)acc);
#else
    for (int k = 0; k < 4; ++k) { const int32_t b = (int8_t)(v >> (8 * k)); acc += (uint32_t)(b * b); }
    return acc;
#endif
}

// exact int32 -> binary64 without the conversion pipe (bit pattern 2^52 + 2^31 + v, minus 2^52 + 2^31) and a fused
// multiply-add that the host emulator evaluates with fma(): used for exact integer sums on the FP64 pipe
RZK_HD double f64_exact_i32(int32_t v)
{
#if defined(__CUDA_ARCH__)
    return __dadd_rn(__hiloint2double(0x43300000, (int)((uint32_t)v ^ 0x80000000u)), -4503601774854144.0);
#else
    return (double)v;
#endif
}
// the same from the biased word u = v + 2^31 (mod 2^32)
RZK_HD double f64_exact_biased(uint32_t u)
{
#if defined(__CUDA_ARCH__)
    return __dadd_rn(__hiloint2double(0x43300000, (int)u), -4503601774854144.0);
#else
    return (double)(int32_t)(u ^ 0x80000000u);
#endif
}
RZK_HD double f64_exact_fma(double a, double b, double c)
{
#if defined(__CUDA_ARCH__)
    return __fma_rn(a, b, c);
#else
    return __builtin_fma(a, b, c);
#endif
}

RZK_HD uint32_t umin32(uint32_t a, uint32_t b) { return a < b ? a : b; }
RZK_HD uint32_t umax32(uint32_t a, uint32_t b) { return a > b ? a : b; }

// x in [0, 2m) -> [0, m)  (unsigned wrap makes x - m huge when x < m)
RZK_HD uint32_t csub(uint32_t x, uint32_t m) { return umin32(x, x - m); }

// Shoup multiplication by a constant w with companion wp = floor(w * 2^32 / p).
// Valid for ANY 32-bit y; result in [0, 2p).
RZK_HD uint32_t shoup_mul(uint32_t w, uint32_t wp, uint32_t y, uint32_t p)
{
    uint32_t q = mulhi32(wp, y);
    return w * y - q * p;
}

// Montgomery product a*b*2^-32 mod p for a*b < p*2^32; pinv = p^-1 mod 2^32.
// Result in (0, 2p).
RZK_HD uint32_t mont_mul(uint32_t a, uint32_t b, uint32_t p, uint32_t pinv)
{
    uint32_t lo = a * b;
    uint32_t hi = mulhi32(a, b);
    uint32_t m = lo * pinv;
    uint32_t mh = mulhi32(m, p);
    return hi - mh + p;
}

// `z` below is an operand that is always 0 but opaque to the compiler (it arrives as a kernel
// parameter).  Writing the two-input adds of the butterflies as a + b + z keeps them three-input
// IADD3 instructions on the ALU pipe; without it ptxas turns them into IMAD.IADD on the FMA-heavy
// pipe, which is the pipe the 32-bit multiplies saturate (ncu: sm__pipe_fmaheavy_cycles_active).

// a + b for a, b in [0, 2p), kept on the ALU pipe.  With `cap` != 0 the add is written as min(a + b, cap) for an
// immediate cap >= 4p - 2 (0xFFFFFFFE covers every p < 2^30) -- an identity the compiler cannot see through, which
// becomes one VIADDMNMX with an immediate: two register operands instead of the three of IADD3 R, R, R with the
// opaque zero z described above (ncu: dispatch stalls on the three-register form).  cap == 0 keeps the z form.
RZK_HD uint32_t add_alu(uint32_t a, uint32_t b, uint32_t z, uint32_t cap) { return cap ? umin32(a + b, cap) : a + b + z; }

// Cooley-Tukey (forward) butterfly, Harvey lazy form: inputs in [0, 4p), outputs in [0, 4p).
RZK_HD void ct_bfly(uint32_t &x, uint32_t &y, uint32_t w, uint32_t wp, uint32_t p, uint32_t p2, uint32_t z, uint32_t cap = 0)
{
    uint32_t xr = csub(x, p2);
    uint32_t t = shoup_mul(w, wp, y, p);
    x = add_alu(xr, t, z, cap);
    y = xr - t + p2;
}

// Gentleman-Sande (inverse) butterfly: inputs in [0, 2p), outputs in [0, 2p).
RZK_HD void gs_bfly(uint32_t &x, uint32_t &y, uint32_t w, uint32_t wp, uint32_t p, uint32_t p2, uint32_t z, uint32_t cap = 0)
{
    uint32_t s = csub(add_alu(x, y, z, cap), p2);
    uint32_t d = x - y + p2;
    x = s;
    y = shoup_mul(w, wp, d, p);
}

// ---------------------------------------------------------------- signed lazy arithmetic for one small prime
// For an auxiliary prime 2^26 < p < 2^31 / 29 every residue is kept as ANY int32 representative and nothing is corrected
// inside a butterfly.  Twiddles and key images are stored centred, w in (-p/2, p/2), with the signed Shoup companion
// w' = round(w * 2^32 / p) (fits int32).  For any |y| < 2^31:
//     t = y*w - floor(y*w' / 2^32) * p   is congruent to y*w and lies in (-p/4, 5p/4)
// (y*w'/2^32 = y*w/p + e with |e| <= |y| / 2^33 < 1/4, and the floor takes off less than one more).
// `mp` is -p as a 32-bit word (a separate operand so that the product term is one multiply-add, no negation).
RZK_HD int32_t mulhi_s32(int32_t a, int32_t b)
{
#if defined(__CUDA_ARCH__)
    return __mulhi(a, b);
#else
    return (int32_t)(((int64_t)a * (int64_t)b) >> 32);
#endif
}

// acc + y*w (mod p), lazy: |result| <= |acc| + 5p/4
RZK_HD uint32_t sshoup_mac(uint32_t w, uint32_t wp, uint32_t y, uint32_t mp, uint32_t acc)
{
    const uint32_t q = (uint32_t)mulhi_s32((int32_t)y, (int32_t)wp);
    return q * mp + (w * y + acc);
}

// Signed Montgomery product a b 2^-32 (mod p) for int32 representatives a, b; pinv = p^-1 mod 2^32.
// m p has the low word of a b, so the difference of the (floor) high words is exact: |result| <= |a b| / 2^32 + p/2 + 1.
RZK_HD uint32_t smont_mul(uint32_t a, uint32_t b, uint32_t p, uint32_t pinv)
{
    const uint32_t lo = a * b;
    const int32_t hi = mulhi_s32((int32_t)a, (int32_t)b);
    const uint32_t m = lo * pinv;
    const int32_t mh = mulhi_s32((int32_t)m, (int32_t)p);
    return (uint32_t)(hi - mh);
}

// Cooley-Tukey (forward) butterfly, four instructions: x' = x + t, y' = x - t = 2x - x'.  Magnitudes grow by 5p/4 per stage.
RZK_HD void ct_bfly_s(uint32_t &x, uint32_t &y, uint32_t w, uint32_t wp, uint32_t mp)
{
    const uint32_t xo = sshoup_mac(w, wp, y, mp, x);
    y = x + x - xo;
    x = xo;
}

// The first forward stage of an operand with |y| <= 1 (the range MODE_SPLITKEY_S admits; larger entries are flagged by the
// program's range check and their item is redone): y*w needs no reduction at all, |y*w| < p/2.  Two instructions.
RZK_HD void ct_bfly_s_tiny(uint32_t &x, uint32_t &y, uint32_t w)
{
    const uint32_t xo = w * y + x;
    y = x + x - xo;
    x = xo;
}

// Gentleman-Sande (inverse) butterfly, five instructions: x' = x + y (magnitude doubles), y' = (x - y) * w (back to 5p/4).
// Needs |x| + |y| < 2^31: the callers reduce the few elements whose run of sums would exceed that (sreduce).
RZK_HD void gs_bfly_s(uint32_t &x, uint32_t &y, uint32_t w, uint32_t wp, uint32_t mp)
{
    const uint32_t s = x + y, d = x - y;
    x = s;
    y = sshoup_mac(w, wp, d, mp, 0u);
}

// The last inverse stage hands its outputs over with the bias 2^31 added (the third operand of the add, the addend of the
// first multiply-add: free), which is the form the exact int32 -> binary64 conversion wants (f64_exact_biased).
RZK_HD void gs_bfly_s_biased(uint32_t &x, uint32_t &y, uint32_t w, uint32_t wp, uint32_t mp)
{
    const uint32_t s = x + y + 0x80000000u, d = x - y;
    x = s;
    y = sshoup_mac(w, wp, d, mp, 0x80000000u);
}

// v -> v - floor(v / 2^SH) * p for 2^SH <= p < 2^SH + 2^SH/640: a representative in (-p/20, 21p/20) from any |v| < 2^31
template <int SH>
RZK_HD uint32_t sreduce(uint32_t v, uint32_t mp)
{
    const uint32_t q = (uint32_t)((int32_t)v >> SH);
    return q * mp + v;
}

// Integer-valued double |v| < 2^31 -> the centred residue mod p as a double (exact; p odd, v/p never within 2^-12 of a half
// integer for the values the split-key program produces: the true result is below p/2 - 2^14 in magnitude).
RZK_HD double center_p_f64(double v, double p, double pinv)
{
    const double magic = 6755399441055744.0;                    // 1.5 * 2^52
#if defined(__CUDA_ARCH__)
    const double k = __dadd_rn(__fma_rn(v, pinv, magic), -magic);
    return __fma_rn(-k, p, v);
#else
    const double k = (__builtin_fma(v, pinv, magic)) - magic;
    return __builtin_fma(-k, p, v);
#endif
}

// Canonical centred representative of an arbitrary i32 representative of a class mod q
// (what ZqI64::from(i64) does).  q < 2^32 < 2q so one step suffices.
RZK_HD int32_t canon_q(int32_t v, uint32_t q)
{
    const int32_t half = (int32_t)((q - 1u) >> 1);
    uint32_t u = (uint32_t)v;
    if (v > half) u -= q;
    else if (v < -half) u += q;
    return (int32_t)u;
}

// Any int32 representative v of a class mod q -> a value congruent to v mod q that is >= -2*p_min for
// every auxiliary prime (so that v + 2p is a valid lazy NTT input in [0, 4p)).  Only the 98,302 lowest
// int32 values need fixing (v + q still fits int32); products only need the class mod q, because
// the integer result is reduced mod q at the end and the CRT ranges are sized for |v| <= 2^31.
// Written as max(v, v + q) on int32 (one add-and-max instruction): v + q wraps to a smaller value unless
// v < q - 2^32 + ... , i.e. exactly when v is so negative that v + q is the representative in [1.36e9, 2^31).
RZK_HD int32_t lift_in(int32_t v, uint32_t q)
{
    const int32_t w = (int32_t)((uint32_t)v + q);
    return w > v ? w : v;
}

// Signed 64-bit value w with |w| < 2^60  ->  centred residue mod q in [-(q-1)/2, (q-1)/2].
//   kqh = q * 2^29 + (q-1)/2   (offset: makes the operand positive and pre-centres it)
//   m30 = floor(2^62 / q)      (32-bit Barrett constant applied to the top 32 bits of the operand)
// The quotient estimate from the top bits is short by at most 1, so one conditional subtraction
// finishes the reduction; subtracting (q-1)/2 then yields the centred representative.
RZK_HD int32_t reduce_q_centered(int64_t w, uint32_t q, uint32_t m30, uint64_t kqh)
{
    const uint64_t wp = (uint64_t)w + kqh;                      // (0, 2^61.7)
    const uint32_t x = (uint32_t)(wp >> 30);
    const uint32_t qh = mulhi32(x, m30);                        // floor(wp/q) or one less
    uint64_t rem = wp - (uint64_t)qh * (uint64_t)q;             // [0, 2q)
    if (rem >= (uint64_t)q) rem -= q;
    return (int32_t)((uint32_t)rem - ((q - 1u) >> 1));
}

// Integer-valued double with |v| < 2^51  ->  centred residue mod q.  k = rint(v / q) (magic-constant rounding), then
// rem = v - k q exactly (one FMA); q is odd and the estimate of v / q is off by < 2^-38, so the rounding always picks the
// nearest integer and |rem| <= (q-1)/2: the canonical representative, no correction step.
RZK_HD int32_t reduce_q_centered_f64(double v, double q, double qinv)
{
    const double magic = 6755399441055744.0;                    // 1.5 * 2^52
#if defined(__CUDA_ARCH__)
    const double k = __dadd_rn(__fma_rn(v, qinv, magic), -magic);
    const double rem = __fma_rn(-k, q, v);
    return __double2loint(__dadd_rn(rem, magic));
#else
    const double k = (__builtin_fma(v, qinv, magic)) - magic;
    const double rem = __builtin_fma(-k, q, v);
    return (int32_t)(int64_t)rem;
#endif
}

}  // namespace rzk
