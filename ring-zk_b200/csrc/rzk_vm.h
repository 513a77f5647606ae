// rzk_vm.h -- data model of the polynomial-op "program" that every protocol phase of
// the R_q hot path is lowered to.  Shared by the host composer (rzk_engine.cu), the
// sm_100a kernel (rzk_kernels.cu) and the host lane emulator (tests/cpp/emu_check.cpp).
//
// One half-warp (16 lanes, 32 coefficients per lane) evaluates one batch item.  A
// program is a list of NORM ops followed by segments; every segment is run once per
// auxiliary prime and ends in inverse transforms whose residues are CRT-combined,
// reduced mod q and either stored or compared with zero.
//
//   out = reduce_q( sum_i key_i (*) in_i  +  sum_j in_a (*) in_b  +  sum_k +-plain_k )
//
// which covers Mat::dot / componentwise_mul / add / sub of /root/reference/src/mat.rs:95-178
// as composed by commit.rs:88-128 and prove/{open,linear,sum}.rs.
#pragma once
#include <stdint.h>
#include <stddef.h>

namespace rzk {

constexpr int kN = 512;            // ring degree handled by the NTT kernels
constexpr int kLanes = 16;         // lanes per item (half warp)
constexpr int kElems = 32;         // coefficients per lane
constexpr int kMaxPrimes = 3;
constexpr int kNumPrimeSlots = 6;  // global static prime list (rzk_tables.cpp)
constexpr int kPadWords = 576;     // 512 + 4 words of padding per 32 (conflict-free uint4 rows)
constexpr int kBufWords = 592;     // transpose buffer per half-warp (+16: bank shift between halves)
constexpr int kSlotWords = 512;    // lane-private slot (uint4 chunks interleaved over lanes)
constexpr int kG2Words = 60;       // lane-specific twiddle words per (direction, prime, lane)
#ifndef RZK_INV_DIT
#define RZK_INV_DIT 1      // signed slots: inverse transform in decimation-in-time form (Cooley-Tukey butterflies + output twist); 0 = Gentleman-Sande (A/B)
#endif
constexpr int kTwistWords = 1024;  // output twist of a signed slot's inverse: 512 (w, w') pairs
constexpr int kG1Words = 66;       // lane-uniform twiddle words per (prime, direction): 32 (w, w') pairs + 2 pad
                                   // (66 = 2 mod 32: the two half warps of a SPLIT warp hit different banks)
constexpr int kKeyPolys = 3;       // non-trivial key polynomials a1'[0], a1'[1], a2'[0] at (n,k,l)=(1,3,1)
constexpr int kMaxOps = 56;
constexpr int kRotHwWords = 1168;  // half-warp region of a program with OP_ROT: the warp's two regions (2336 words) hold the
                                   // extended row E = [B - c | B + c | B - c] (1536 words, B = 2^31), a row of biased zeros (512
                                   // words) that partners the last term of an odd count, and the list of row offsets, one
                                   // uint16 per term (256 words); 1168 = 16 (mod 32)
constexpr int kRotZeroOff = 1536;  // word offset of the zero row in the warp region
constexpr int kRotListOff = 2048;  // word offset of the list
constexpr int kMaxStreams = 12;

enum OpCode : uint8_t {
    OP_END = 0,
    OP_SEG,      // start of a segment (ops up to the next SEG/END run once per prime)
    OP_FWD,      // cur = NTT_p(stream a, poly off [+ it*step]);  b: FWD_* flags
    OP_MACK,     // acc[a] (+)= key[b] (.) cur            c: MAC_* flags
    OP_MACV,     // acc[a] (+)= slot (.) cur (Montgomery)  c: MAC_* flags
    OP_ST,       // slot = cur  (b: ST_RAW keeps lazy values; default reduces to [0,p) for OP_MACV)
    OP_LD,       // cur = operand slot of half warp a (warp-per-item modes)
    OP_INV,      // inverse NTT of acc[a]; stash b; on the last prime: CRT -> V
    OP_ADDP,     // (last prime only) V += +-stream a poly off      c: MAC_NEG
    OP_FIN,      // (last prime only) res = reduce_q(V);  b: FIN_* flags (store to stream a / compare with 0)
    OP_NORM,     // (once) squared 2-norm check of c consecutive polys of stream a from off; b = bound select
    OP_LOOP,     // repeat the ops up to OP_ENDLOOP `off` times (iteration index scales `step`)
    OP_ENDLOOP,
    OP_STG,      // NTT image out: stream a [group][prime][512 words, lane-private order] = cur reduced to [0,p)
                 // (the Montgomery operand a later launch consumes with OP_MACG; cur comes from a FWD_SCALED transform)
    OP_MACG,     // acc[a] (+)= image stream b (.) cur (Montgomery), the image read from global memory     c: MAC_* flags
    OP_ROT,      // (epilogue, warp-per-item modes) V += +-(stream a poly off) * (sparse int8 polynomial of stream b) as signed
                 // rotations of the int32 row: sum_k d[pos_k] * X^pos_k * c, no multiplication by a transform (SURVEY kernel K4);
                 // c: MAC_NEG.  Exact for any int8 d and any int32 c (d in {-1, 0, 1}^N takes the fast paired form).
};

enum : uint8_t { FWD_SCALED = 1, FWD_CHECK_SMALL = 2, FWD_HWPOLY = 4 };   // HWPOLY: half warp h reads poly off + h
enum : uint8_t { ST_RAW = 1 };
enum : uint8_t { MAC_INIT = 1, MAC_NEG = 2 };
enum : uint8_t { FIN_STORE = 1, FIN_CMPZ = 2 };
enum : uint8_t { DT_I32 = 0, DT_I8 = 1 };
enum : uint32_t { FLAG_FAIL = 1u, FLAG_RANGE = 2u };

struct Op {
    uint8_t code, a, b, c;
    uint16_t off, step;
};

struct Stream {
    const void *base;     // device pointer
    uint32_t stride;      // polynomials per item group
    uint32_t div;         // item group = item / div  (div > 1 only for per-instance streams of Sum proofs)
    uint32_t dtype;       // DT_I32 / DT_I8
    uint32_t magic;       // item / div == (item * magic) >> shift for item < 2^28 (set_stream_div)
    uint32_t shift;
    uint32_t pad_;
};

// Exact division by the stream's (run-time) group size without a divide sequence: round-up multiplier
// magic = ceil(2^shift / div), shift = 28 + ceil(log2 div); exact for item < 2^28, div < 2^16.
inline void set_stream_div(Stream &s, uint32_t div)
{
    uint32_t c = 0;
    while ((1u << c) < div) ++c;
    s.div = div;
    s.shift = 28 + c;
    s.magic = (uint32_t)((((uint64_t)1 << s.shift) + div - 1) / div);
}

struct PrimeC {
    uint32_t p, p2, pinv;   // p, 2p, p^-1 mod 2^32
    uint32_t rn, rnp;       // R * N^-1 mod p (R = 2^32) and its Shoup companion
    uint32_t slot;          // index into the global static prime list (selects G1 twiddles)
    uint32_t half;          // (p-1)/2
    uint32_t pad_;          // always 0; used as the compiler-opaque zero operand `z` of the butterflies
};

struct CrtC {
    // Garner: V = a0 + p0*h1 + p0*p1*h2
    uint32_t inv01, inv01p;       // p0^-1 mod p1 (Shoup pair)
    uint32_t p0modp2, p0modp2p;   // p0 mod p2 (Shoup pair w.r.t. p2)
    uint32_t inv012, inv012p;     // (p0*p1)^-1 mod p2 (Shoup pair)
    uint32_t pad0_, pad1_;
    uint64_t P01;                 // p0*p1
    uint64_t P01half;             // (p0*p1 - 1)/2          (centring for 2 primes)
    uint64_t P01modq;             // (p0*p1) mod q
    uint64_t Pmodq;               // (p0*p1*p2) mod q
    uint64_t Phalf_lo;            // (p0*p1*p2 - 1)/2, low 64 bits
    uint64_t Phalf_hi;            //                    high bits
};

struct VmLaunch {
    Op ops[kMaxOps];
    Stream st[kMaxStreams];
    PrimeC pc[kMaxPrimes];
    CrtC crt;
    uint64_t kqh;              // q * 2^29 + (q-1)/2   (reduce_q_centered offset)
    double qd, qinvd;          // q and 1/q as doubles (reduce_q_centered_f64)
    double p0d, p0qinvd;       // prime 0 and p0 / q as doubles (crt2_mod_q_f64)
    double p0invd;             // 1 / prime 0 (center_p_f64, MODE_SPLITKEY_S)
    uint64_t norm_sq_lim[2];   // (bound+1)^2 - 1 for [0]=commit, [1]=verify constraint
    uint32_t norm_abs_lim[2];  // bound
    uint32_t q;
    uint32_t small_lim;        // |v| limit for FWD_CHECK_SMALL operands
    uint32_t n_items;
    uint32_t flag_div;         // flags index = item / flag_div
    uint32_t np;               // primes in this launch
    uint32_t m30;              // floor(2^62 / q)         (reduce_q_centered Barrett constant)
    uint32_t *flags;           // device, one word per item group
    const uint32_t *g1tab;     // device, [prime slot][dir][kG1Words]
    const uint32_t *g2tab;     // device, [prime slot][dir][16][60]
    const uint32_t *keytab;    // device, [prime slot][3][2][576]
    const uint32_t *twist[kMaxPrimes];   // device; signed slots (RZK_INV_DIT): the output twists psi^-i of the decimation-in-time inverse, staged into shared memory
    // shared-memory layout of one half-warp region (words); sized from what the program uses
    uint32_t hw_words;         // == 16 (mod 32): the two half warps of a warp sit 16 banks apart
    uint32_t off_slot, off_acc1, off_stash;
    // input streams whose next-item rows are prefetched into L2 while the current item is computed
    uint32_t n_prefetch;
    uint8_t prefetch[8];
    uint32_t prefetch_polys[8]; // leading polynomials of the row group that the program reads (a verify launch handed the full
                               // commitment with c_stride 2 reads c1 only: the unused row is not pulled through DRAM)
    uint32_t cta_sync;         // keep the warps of a CTA in step (instruction-cache locality)
    uint32_t loop_count;       // trip count of OP_LOOP in compile-time programs (terms of a Sum proof - 1)
    const uint32_t *item_mask; // optional: only items with item_mask[item / mask_div] != 0 are processed (fallback launches)
    const uint32_t *any_item;  // optional: the whole launch returns at once when *any_item == 0
    uint32_t mask_div;         // item groups per mask word (0 is read as 1)
    uint32_t ld128;            // A/B variant of OP_FWD's int32 loads (rzk_vm_exec.cuh op_fwd); 0 = 32-bit strided loads
    uint32_t *rmark;           // optional: a range error (FWD_CHECK_SMALL) marks rmark[item / flag_div] = 1 and *rmark_any = 1
    uint32_t *rmark_any;       // instead of setting FLAG_RANGE -- the hand-over to a masked fallback launch (dev_commit)
    uint32_t pp_mode;          // phase mixing between the two halves of a CTA (rzk_vm_exec.cuh pp_acquire); 0 = off
    uint32_t alias_slot;       // the operand slot may overlay the transpose buffer (programs whose OP_LDs all
                               // precede the inverse transforms and that never use OP_MACV)
    uint32_t stash_words;      // words of residue stash per half warp: nstash * (np - 1) * kSlotWords (MODE_SEQ, np > 1)
    uint32_t *gstash;          // device: the stash lives in global memory, [CTA][warp][half warp][stash_words]
                               // (nullptr: in the half warp's shared-memory region, host emulator)
    uint32_t acc1_global;      // warp-per-item programs with OP_ROT and a second accumulator: accumulator 1 lives in the half
                               // warp's global stash region (written and read back by the same lane, L2-resident) so that
                               // the rotation sum may overlay the whole shared-memory region of the warp
    uint32_t pad32_;
};

// What a program needs per half warp (decides how many warps fit in shared memory).
struct ProgNeeds {
    bool slot = false, acc1 = false, rot = false;
    int nstash = 0;
};

inline ProgNeeds scan_needs(const Op *ops)
{
    ProgNeeds n;
    for (int i = 0; i < kMaxOps && ops[i].code != OP_END; ++i) {
        const Op &o = ops[i];
        if (o.code == OP_ST || o.code == OP_MACV || o.code == OP_LD) n.slot = true;
        if ((o.code == OP_MACK || o.code == OP_MACV || o.code == OP_MACG || o.code == OP_INV) && o.a == 1) n.acc1 = true;
        if (o.code == OP_INV && (int)o.b + 1 > n.nstash) n.nstash = (int)o.b + 1;
        if (o.code == OP_ROT) n.rot = true;
    }
    return n;
}

// Lists the input streams of a program (read by OP_FWD / OP_ADDP / OP_NORM / OP_ROT) for prefetching, with the number of
// leading polynomials of each row group that are actually read.
inline void list_prefetch(VmLaunch &K)
{
    bool is_out[kMaxStreams] = {}, is_in[kMaxStreams] = {};
    uint32_t used[kMaxStreams] = {};
    bool in_loop = false;
    auto touch = [&](int s, uint32_t off, uint32_t cnt, bool whole) {
        is_in[s] = true;
        const uint32_t hi = whole ? K.st[s].stride : off + cnt;
        if (hi > used[s]) used[s] = hi;
    };
    for (int i = 0; i < kMaxOps && K.ops[i].code != OP_END; ++i) {
        const Op &o = K.ops[i];
        if (o.code == OP_LOOP) in_loop = true;
        if (o.code == OP_ENDLOOP) in_loop = false;
        if (o.code == OP_FIN && (o.b & FIN_STORE)) is_out[o.a] = true;
        if (o.code == OP_FWD) touch(o.a, o.off, (o.b & FWD_HWPOLY) ? 2u : 1u, in_loop && o.step);
        if (o.code == OP_ADDP) touch(o.a, o.off, 1u, in_loop && o.step);
        if (o.code == OP_NORM) touch(o.a, o.off, o.c, false);
        if (o.code == OP_ROT) { touch(o.a, o.off, 1u, in_loop && o.step); touch(o.b, 0u, 1u, false); }
    }
    K.n_prefetch = 0;
    for (int s = 0; s < kMaxStreams && K.n_prefetch < 8; ++s)
        if (is_in[s] && !is_out[s] && K.st[s].div == 1) {
            K.prefetch_polys[K.n_prefetch] = used[s] < K.st[s].stride ? used[s] : K.st[s].stride;
            K.prefetch[K.n_prefetch++] = (uint8_t)s;
        }
}

// Fills the half-warp layout fields; `split` programs keep no stash (residues are swapped by shuffles).
inline void layout_hw(VmLaunch &K, bool split)
{
    const ProgNeeds n = scan_needs(K.ops);
    uint32_t w = kBufWords;
    if (K.alias_slot && n.slot) K.off_slot = 0;
    else { K.off_slot = w; if (n.slot) w += kSlotWords; }
    K.off_acc1 = w;  if (n.acc1 && !K.acc1_global) w += kSlotWords;
    K.stash_words = (!split && K.np > 1) ? (uint32_t)n.nstash * (K.np - 1) * kSlotWords : ((K.acc1_global && n.acc1) ? (uint32_t)kSlotWords : 0u);
    K.off_stash = w; if (!K.gstash) w += K.stash_words;
    // OP_ROT overlays the whole warp region (transpose buffers included: they are dead in the epilogue); programs that keep
    // an operand slot or a second accumulator in the region cannot use it (rot_layout_ok; acc1_global moves the accumulator out)
    if (n.rot && w < (uint32_t)kRotHwWords) w = kRotHwWords;
    K.hw_words = w;
}

// OP_ROT needs the warp region to itself during the epilogue: warp-per-item programs without an operand slot, and with the
// second accumulator (if any) in the global stash region.
inline bool rot_layout_ok(const Op *ops, bool split, bool acc1_global = false)
{
    const ProgNeeds n = scan_needs(ops);
    return !n.rot || (split && !n.slot && (!n.acc1 || acc1_global));
}

// Host-side twiddle tables for one prime (built in rzk_tables.cpp).
struct PrimeTables {
    uint32_t p;
    uint32_t g1[2][32][2];                 // [dir][twiddle index][w, w']  (indices 1..31 used)
    uint32_t g2[2][kLanes][kG2Words];      // [dir][lane][...]
    uint32_t psi, psi_inv, ninv, r, rn, rnp, pinv;
    uint32_t tw[2][kN][2];                 // full tables (reference transform, key setup)
    uint32_t twist[kN][2];                 // signed slots: (centred psi^-i, signed companion) in the lane order of inv_g1_dit (rzk_tables.cpp)
    uint32_t twist_rn[kN][2];              // the same times R N^-1 (R = 2^32): the twist of the product-sum programs (MODE_SEQ_S), whose
                                           // Montgomery products then take both operands as plain transforms
};

}  // namespace rzk
