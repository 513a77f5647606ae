// rzk_tables.cpp -- see rzk_tables.h
#include "rzk_tables.h"

#include <mutex>
#include <stdexcept>
#include <string.h>

namespace rzk {

const uint32_t kPrimeList[kNumPrimeSlots] = {      // slot 0 is also the compile-time prime of the split-key program (kStaticPrime0)
    1073692673u, 1073668097u, 1073655809u,   // < 2^30
    67153921u, 67219457u, 67227649u,         // the three smallest primes == 1 (mod 4096) above 2^26, for the signed lazy arithmetic
                                             // (rzk_arith.cuh): slot 3 = kStaticPrimeS = 2^26 + 45057; product 2^78.006
};

bool slot_is_signed(int slot) { return slot >= 3; }

// w in [0, p) -> centred value and signed Shoup companion round(w * 2^32 / p), both as 32-bit words
void signed_shoup_pair(uint32_t w, uint32_t p, uint32_t &wc, uint32_t &wcp)
{
    int64_t c = (int64_t)w;
    if (c > (int64_t)(p - 1) / 2) c -= (int64_t)p;
    const __int128 num = ((__int128)c << 32) * 2 + (__int128)p;          // floor((2 c 2^32 + p) / (2 p)) = round(c 2^32 / p)
    __int128 qv = num / (2 * (__int128)p);
    if (num % (2 * (__int128)p) < 0) --qv;                               // floor for negative numerators
    wc = (uint32_t)(int32_t)c;
    wcp = (uint32_t)(int32_t)qv;
}

uint32_t mod_pow(uint32_t b, uint64_t e, uint32_t p)
{
    uint64_t r = 1, x = b % p;
    while (e) {
        if (e & 1) r = r * x % p;
        x = x * x % p;
        e >>= 1;
    }
    return (uint32_t)r;
}

uint32_t mod_inv(uint32_t a, uint32_t p) { return mod_pow(a, (uint64_t)p - 2, p); }

uint32_t shoup_companion(uint32_t w, uint32_t p) { return (uint32_t)(((uint64_t)w << 32) / p); }

static int brv9(int x)
{
    int r = 0;
    for (int i = 0; i < 9; ++i) r |= ((x >> i) & 1) << (8 - i);
    return r;
}

static void build(PrimeTables &T, uint32_t p, bool signed_form)
{
    memset(&T, 0, sizeof(T));
    T.p = p;
    // primitive 2N-th root of unity: g = x^((p-1)/2N) with g^N == -1
    uint32_t psi = 0;
    for (uint32_t x = 2; x < 1000; ++x) {
        uint32_t g = mod_pow(x, (uint64_t)(p - 1) / (2 * kN), p);
        if (mod_pow(g, kN, p) == p - 1) { psi = g; break; }
    }
    if (!psi) throw std::runtime_error("no 2N-th root of unity");
    T.psi = psi;
    T.psi_inv = mod_inv(psi, p);
    T.ninv = mod_inv(kN, p);
    T.r = (uint32_t)((1ull << 32) % p);
    T.rn = (uint32_t)((uint64_t)T.r * T.ninv % p);
    T.rnp = shoup_companion(T.rn, p);
    // p^-1 mod 2^32 by Newton iteration
    uint32_t inv = p;                       // correct to 3 bits for odd p
    for (int i = 0; i < 5; ++i) inv *= 2u - p * inv;
    T.pinv = inv;
    for (int idx = 0; idx < kN; ++idx) {
        uint32_t wf = mod_pow(psi, (uint64_t)brv9(idx), p);
        uint32_t wi = mod_pow(T.psi_inv, (uint64_t)brv9(idx), p);
        T.tw[0][idx][0] = wf; T.tw[0][idx][1] = shoup_companion(wf, p);
        T.tw[1][idx][0] = wi; T.tw[1][idx][1] = shoup_companion(wi, p);
    }
    for (int d = 0; d < 2; ++d) {
        for (int idx = 0; idx < 32; ++idx) {
            T.g1[d][idx][0] = T.tw[d][idx][0];
            T.g1[d][idx][1] = T.tw[d][idx][1];
        }
        for (int t = 0; t < kLanes; ++t) {
            int w = 0;
            for (int s = 5; s <= 8; ++s) {                 // stage s: 2^(s-4) twiddles per lane
                int cnt = 1 << (s - 4);
                for (int a = 0; a < cnt; ++a) {
                    int idx = (1 << s) + cnt * t + a;
                    T.g2[d][t][w++] = T.tw[d][idx][0];
                    T.g2[d][t][w++] = T.tw[d][idx][1];
                }
            }
        }
    }
    if (signed_form) {
        // the kernel-side tables of a signed slot hold (centred w, signed companion); T.tw keeps the canonical residues
        for (int d = 0; d < 2; ++d) {
            for (int idx = 0; idx < 32; ++idx) signed_shoup_pair(T.tw[d][idx][0], p, T.g1[d][idx][0], T.g1[d][idx][1]);
            for (int t = 0; t < kLanes; ++t)
                for (int w = 0; w < kG2Words; w += 2) signed_shoup_pair(T.g2[d][t][w], p, T.g2[d][t][w], T.g2[d][t][w + 1]);
        }
#if RZK_INV_DIT
        // Inverse of a signed slot in decimation-in-time form (rzk_vm_exec.cuh inv_g2_dit / inv_g1_dit): the forward output is the
        // cyclic DFT of (a_j psi^j) in bit-reversed order, so the inverse is a cyclic DIT transform on that order -- stage m = 2 .. N:
        // a[k+j], a[k+j+m/2] <- a[k+j] +- w_m^j a[k+j+m/2], w_m = psi^(-2N/m) -- followed by the twist a_i *= psi^-i (N^-1 rides on the
        // key images / the pre-scaled operand as before).  Twiddles depend on the position inside a block, so the roles of the two
        // tables swap: g1[1] holds the lane-uniform twiddles of the contiguous-layout stages m = 4, 8, 16 (j = 1 .. m/2-1 at
        // index m/4 - 1 + j - 1 ... packed 0 .. 10) and, at index 16 + t, lane t's twiddle of stage m = 32; g2[1][t] holds lane t's
        // twiddles of the stages m = 64 .. 512 (j = t + 16 a, a = 0 .. m/32 - 1: 2 + 4 + 8 + 16 pairs).
        auto wpow = [&](int m, int j) { return mod_pow(T.psi_inv, (uint64_t)(2 * kN / m) * (uint64_t)j, p); };
        memset(T.g1[1], 0, sizeof(T.g1[1]));
        int idx = 0;
        for (int m = 4; m <= 16; m <<= 1)
            for (int j = 1; j < m / 2; ++j, ++idx) signed_shoup_pair(wpow(m, j), p, T.g1[1][idx][0], T.g1[1][idx][1]);
        for (int t = 0; t < kLanes; ++t) signed_shoup_pair(wpow(32, t), p, T.g1[1][16 + t][0], T.g1[1][16 + t][1]);
        for (int t = 0; t < kLanes; ++t) {
            int w = 0;
            for (int m = 64; m <= kN; m <<= 1)
                for (int a = 0; a < m / 32; ++a, w += 2) signed_shoup_pair(wpow(m, t + 16 * a), p, T.g2[1][t][w], T.g2[1][t][w + 1]);
        }
        // twist psi^-i in the order the strided layout reads it with 128-bit loads: coefficient i = t + 16 m lives in pair
        // ((m >> 1) * 16 + t) * 2 + (m & 1), so lane t's uint4 number k holds the pairs of its registers m = 2k and 2k + 1
        for (int i = 0; i < kN; ++i) {
            const int t = i & 15, m = i >> 4, slot = ((m >> 1) * 16 + t) * 2 + (m & 1);
            signed_shoup_pair(mod_pow(T.psi_inv, (uint64_t)i, p), p, T.twist[slot][0], T.twist[slot][1]);
            signed_shoup_pair((uint32_t)((uint64_t)mod_pow(T.psi_inv, (uint64_t)i, p) * T.rn % p), p, T.twist_rn[slot][0], T.twist_rn[slot][1]);
        }
#endif
    }
}

const PrimeTables &prime_tables(int slot)
{
    static PrimeTables tabs[kNumPrimeSlots];
    static std::once_flag once[kNumPrimeSlots];
    if (slot < 0 || slot >= kNumPrimeSlots) throw std::out_of_range("prime slot");
    std::call_once(once[slot], [slot] { build(tabs[slot], kPrimeList[slot], slot_is_signed(slot)); });
    return tabs[slot];
}

void ntt_forward_ref(const PrimeTables &T, uint32_t a[kN])
{
    const uint64_t p = T.p;
    int t = kN;
    for (int m = 1; m < kN; m <<= 1) {
        t >>= 1;
        for (int i = 0; i < m; ++i) {
            const uint64_t S = T.tw[0][m + i][0];
            const int j1 = 2 * i * t;
            for (int j = j1; j < j1 + t; ++j) {
                uint64_t U = a[j], V = (uint64_t)a[j + t] * S % p;
                a[j] = (uint32_t)((U + V) % p);
                a[j + t] = (uint32_t)((U + p - V) % p);
            }
        }
    }
}

void ntt_inverse_ref(const PrimeTables &T, uint32_t a[kN])
{
    const uint64_t p = T.p;
    int t = 1;
    for (int m = kN; m > 1; m >>= 1) {
        const int h = m >> 1;
        int j1 = 0;
        for (int i = 0; i < h; ++i) {
            const uint64_t S = T.tw[1][h + i][0];
            for (int j = j1; j < j1 + t; ++j) {
                uint64_t U = a[j], V = a[j + t];
                a[j] = (uint32_t)((U + V) % p);
                a[j + t] = (uint32_t)((U + p - V) % p * S % p);
            }
            j1 += 2 * t;
        }
        t <<= 1;
    }
    for (int j = 0; j < kN; ++j) a[j] = (uint32_t)((uint64_t)a[j] * T.ninv % p);
}

PrimeC make_prime_consts(int slot)
{
    const PrimeTables &T = prime_tables(slot);
    PrimeC c;
    memset(&c, 0, sizeof(c));
    c.p = T.p;
    c.p2 = 2 * T.p;
    c.pinv = T.pinv;
    c.rn = T.rn;
    c.rnp = T.rnp;
    if (slot_is_signed(slot)) signed_shoup_pair(T.rn, T.p, c.rn, c.rnp);       // centred R N^-1 with its signed companion
    c.slot = (uint32_t)slot;
    c.half = (T.p - 1) / 2;
    return c;
}

CrtC make_crt_consts(const int *slots, int np, uint64_t q)
{
    CrtC c;
    memset(&c, 0, sizeof(c));
    typedef unsigned __int128 u128;
    const uint64_t p0 = kPrimeList[slots[0]];
    c.P01 = p0;
    c.P01half = (p0 - 1) / 2;
    if (np >= 2) {
        const uint32_t p1 = kPrimeList[slots[1]];
        c.inv01 = mod_inv((uint32_t)(p0 % p1), p1);
        c.inv01p = shoup_companion(c.inv01, p1);
        c.P01 = p0 * p1;
        c.P01half = (c.P01 - 1) / 2;
    }
    c.P01modq = c.P01 % q;
    if (np >= 3) {
        const uint32_t p2 = kPrimeList[slots[2]];
        c.p0modp2 = (uint32_t)(p0 % p2);
        c.p0modp2p = shoup_companion(c.p0modp2, p2);
        c.inv012 = mod_inv((uint32_t)(c.P01 % p2), p2);
        c.inv012p = shoup_companion(c.inv012, p2);
        u128 P = (u128)c.P01 * p2;
        c.Pmodq = (uint64_t)(P % q);
        u128 half = (P - 1) / 2;
        c.Phalf_lo = (uint64_t)half;
        c.Phalf_hi = (uint64_t)(half >> 64);
    }
    return c;
}

void key_image(const PrimeTables &T, const int64_t *poly, uint32_t *out)
{
    uint32_t a[kN];
    const int64_t p = T.p;
    for (int i = 0; i < kN; ++i) {
        int64_t v = poly[i] % p;
        if (v < 0) v += p;
        a[i] = (uint32_t)v;
    }
    ntt_forward_ref(T, a);
    memset(out, 0, sizeof(uint32_t) * 2 * kPadWords);
    for (int i = 0; i < kN; ++i) {
        uint32_t w = (uint32_t)((uint64_t)a[i] * T.ninv % T.p);
        out[pad_index(i)] = w;
        out[kPadWords + pad_index(i)] = shoup_companion(w, T.p);
    }
}

// the signed-slot form of a key image: centred residues with signed Shoup companions
static void key_image_to_signed(const PrimeTables &T, uint32_t *out)
{
    for (int i = 0; i < kN; ++i)
        signed_shoup_pair(out[pad_index(i)], T.p, out[pad_index(i)], out[kPadWords + pad_index(i)]);
}

void key_image_split_signed(const PrimeTables &T, const int64_t *poly, uint32_t *out)
{
    key_image_split(T, poly, out);
    key_image_to_signed(T, out);
    key_image_to_signed(T, out + 2 * kPadWords);
}

void key_image_split(const PrimeTables &T, const int64_t *poly, uint32_t *out)
{
    int64_t lo[kN], hi[kN];
    for (int i = 0; i < kN; ++i) {
        int64_t l = ((poly[i] + 32768) & 0xFFFF) - 32768;     // centred low 16 bits
        lo[i] = l;
        hi[i] = (poly[i] - l) / 65536;
    }
    key_image(T, lo, out);
    key_image(T, hi, out + 2 * kPadWords);
}

}  // namespace rzk
