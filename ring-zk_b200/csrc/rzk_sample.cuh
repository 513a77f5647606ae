// rzk_sample.cuh -- optional on-device samplers for the three random inputs of the
// protocols: r uniform in [-b, b]^N (/root/reference/src/polynomial.rs:14-24), y = trunc(N(0, sigma)) per coefficient
// (polynomial.rs:28-44) and the challenge d with min(kappa, N) entries +-1 at uniformly random distinct positions
// (challenge_space.rs:12-33).  They exist to keep r and y on the device between the prover's two calls; the north-star
// flow (randomness drawn host side and passed in) does not use them, and they do NOT reproduce the stream of Rust's
// `rand` -- the guarantee is distributional, plus bit-reproducibility from (seed, tag) through the counter-based
// generator below, which tests/philox_ref.py restates in numpy.
//
// Generator: Philox4x32-10 (Salmon et al., SC'11), key = the 64-bit seed, counter = (block, index lo, index hi,
// tag << 8 | attempt).
#pragma once
#include <stdint.h>
#include <math.h>
#include "rzk_arith.cuh"

namespace rzk {

struct Philox4 { uint32_t x, y, z, w; };

RZK_HD Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1)
{
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = mulhi32(M0, c0), lo0 = M0 * c0;
        const uint32_t hi1 = mulhi32(M1, c2), lo1 = M1 * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += W0; k1 += W1;
    }
    Philox4 o; o.x = c0; o.y = c1; o.z = c2; o.w = c3;
    return o;
}

// Exactly uniform value in [0, range) from 32-bit words: Lemire's multiply-shift with rejection; word `lane` of
// attempt blocks 0, 1, ... (at most 8 attempts: the rejection probability is range / 2^32 per attempt)
RZK_HD uint32_t sample_below(uint32_t range, uint32_t block, uint32_t idx_lo, uint32_t idx_hi, uint32_t tag, int lane,
                             uint32_t k0, uint32_t k1)
{
    const uint32_t thresh = (0u - range) % range;               // 2^32 mod range
    uint32_t v = 0;
    for (uint32_t attempt = 0; attempt < 8; ++attempt) {
        const Philox4 o = philox4x32_10(block, idx_lo, idx_hi, (tag << 8) | attempt, k0, k1);
        const uint32_t u = lane == 0 ? o.x : lane == 1 ? o.y : lane == 2 ? o.z : o.w;
        const uint64_t m = (uint64_t)u * (uint64_t)range;
        v = (uint32_t)(m >> 32);
        if ((uint32_t)m >= thresh) break;
    }
    return v;
}

// coefficient i of polynomial `poly`: uniform in [-b, b]
RZK_HD int32_t sample_small_coeff(uint64_t poly, uint32_t i, uint32_t b, uint32_t tag, uint32_t k0, uint32_t k1)
{
    return (int32_t)sample_below(2u * b + 1u, i >> 2, (uint32_t)poly, (uint32_t)(poly >> 32), tag, (int)(i & 3u), k0, k1) - (int32_t)b;
}

// coefficients 2g and 2g + 1 of polynomial `poly`: Box-Muller on two 53-bit uniforms, truncated toward zero
RZK_HD void sample_gaussian_pair(uint64_t poly, uint32_t g, double sigma, uint32_t tag, uint32_t k0, uint32_t k1, int32_t *out2)
{
    const Philox4 o = philox4x32_10(g, (uint32_t)poly, (uint32_t)(poly >> 32), tag << 8, k0, k1);
    const double u1 = ((double)((((uint64_t)o.x << 32) | o.y) >> 11) + 0.5) * (1.0 / 9007199254740992.0);
    const double u2 = ((double)((((uint64_t)o.z << 32) | o.w) >> 11) + 0.5) * (1.0 / 9007199254740992.0);
    const double rad = sigma * sqrt(-2.0 * log(u1));
    const double ang = 6.283185307179586476925286766559 * u2;
    out2[0] = (int32_t)trunc(rad * cos(ang));
    out2[1] = (int32_t)trunc(rad * sin(ang));
}

// challenge of one item: min(kappa, N) distinct positions (9 bits of a word: N = 512 is a power of two, so no
// rejection beyond "already taken"), sign from bit 0.  `d` (N int8, zeroed by the caller) is written in place.
RZK_HD void sample_challenge_item(uint64_t item, uint32_t n, uint32_t kappa, uint32_t tag, uint32_t k0, uint32_t k1, int8_t *d)
{
    const uint32_t want = kappa < n ? kappa : n;
    uint32_t have = 0;
    for (uint32_t block = 0; have < want; ++block) {
        const Philox4 o = philox4x32_10(block, (uint32_t)item, (uint32_t)(item >> 32), tag << 8, k0, k1);
        const uint32_t w[4] = {o.x, o.y, o.z, o.w};
        for (int j = 0; j < 4 && have < want; ++j) {
            const uint32_t pos = (w[j] >> 23) & (n - 1u);
            if (d[pos] == 0) { d[pos] = (w[j] & 1u) ? (int8_t)1 : (int8_t)-1; ++have; }
        }
    }
}

}  // namespace rzk
