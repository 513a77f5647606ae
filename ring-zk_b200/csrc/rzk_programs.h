// rzk_programs.h -- lowers each protocol phase of ring-zk to a polynomial-op program
// (rzk_vm.h).  Shapes are the reference's default (n, k, l) = (1, 3, 1)
// (/root/reference/src/params.rs:121-138): the key is a1 = [1, a11, a12],
// a2 = [0, 1, a22] (commit.rs:33-60), so
//     A1 . v = v0 + a11*v1 + a12*v2        A2 . v = v1 + a22*v2
// and the products with the structural 1 / 0 blocks that Mat::dot executes
// (mat.rs:106-113) become plain additions / disappear.
//
// Key polynomial indices: 0 = a11, 1 = a12, 2 = a22.
#pragma once
#include <string.h>
#include "rzk_vm.h"

namespace rzk {

// Program builder.  A literal type so that the same builders produce both the runtime programs the
// generic interpreter decodes and the compile-time programs the specialised kernels are unrolled from.
struct Prog {
    Op ops[kMaxOps] = {};
    int n = 0;
    uint8_t alias_slot = 0;
    uint8_t acc1_global = 0;
    constexpr void add(uint8_t code, int a = 0, int b = 0, int c = 0, int off = 0, int step = 0)
    {
        Op &o = ops[n++];
        o.code = code; o.a = (uint8_t)a; o.b = (uint8_t)b; o.c = (uint8_t)c;
        o.off = (uint16_t)off; o.step = (uint16_t)step;
    }
    constexpr void end() { add(OP_END); }
    void install(VmLaunch &K) const
    {
        for (int i = 0; i < kMaxOps; ++i) K.ops[i] = ops[i];
        K.alias_slot = alias_slot;
        K.acc1_global = acc1_global;
    }
};

// ---- 2-prime programs -------------------------------------------------------------

// [c1; c2] = [a1; a2] . r + [0; x]                       commit.rs:88-128
// streams: 0 = x (1 poly), 1 = r (3 polys), 2 = c out (2 polys)
constexpr inline void prog_commit(Prog &P, int sx = 0, int sr = 1, int sc = 2, bool with_norm = true)
{
    if (with_norm) P.add(OP_NORM, sr, /*commit bound*/ 0, /*count*/ 3, 0);     // commit.rs:102
    P.add(OP_SEG);
    P.add(OP_FWD, sr, 0, 0, 1);
    P.add(OP_MACK, 0, 0, MAC_INIT);
    P.add(OP_FWD, sr, 0, 0, 2);
    P.add(OP_MACK, 0, 1, 0);
    P.add(OP_MACK, 1, 2, MAC_INIT);
    P.add(OP_INV, 0, 0);
    P.add(OP_ADDP, sr, 0, 0, 0);
    P.add(OP_FIN, sc, FIN_STORE, 0, 0);
    P.add(OP_INV, 1, 1);
    P.add(OP_ADDP, sr, 0, 0, 1);
    P.add(OP_ADDP, sx, 0, 0, 0);
    P.add(OP_FIN, sc, FIN_STORE, 0, 1);
}

// Same commitment for small randomness (|r| <= 15) in MODE_SPLITKEY: one prime, the key split
// into 16-bit halves (half warp 0: lo images, half warp 1: hi images), the two forward
// transforms shared between the half warps through their operand slots.
constexpr inline void prog_commit_splitkey(Prog &P, int sx = 0, int sr = 1, int sc = 2, bool with_norm = true)
{
    P.alias_slot = 1;         // ST/LD finish before the first inverse transform touches the buffer
    if (with_norm) P.add(OP_NORM, sr, /*commit bound*/ 0, /*count*/ 3, 0);     // commit.rs:102
    P.add(OP_SEG);
    P.add(OP_FWD, sr, FWD_HWPOLY | FWD_CHECK_SMALL, 0, 1);    // half warp h transforms r[1 + h]
    P.add(OP_ST, 0, ST_RAW);
    P.add(OP_LD, 0);
    P.add(OP_MACK, 0, 0, MAC_INIT);
    P.add(OP_LD, 1);
    P.add(OP_MACK, 0, 1, 0);
    P.add(OP_MACK, 1, 2, MAC_INIT);
    P.add(OP_INV, 0, 0);
    P.add(OP_ADDP, sr, 0, 0, 0);
    P.add(OP_FIN, sc, FIN_STORE, 0, 0);
    P.add(OP_INV, 1, 1);
    P.add(OP_ADDP, sr, 0, 0, 1);
    P.add(OP_ADDP, sx, 0, 0, 0);
    P.add(OP_FIN, sc, FIN_STORE, 0, 1);
}

// t = A1 . y  (open.rs:97, linear.rs:118-121, sum.rs:145-151) and optionally w = A2 . y
// (the inner factor of u, linear.rs:124-129 / sum.rs:154-160).
// sv = y stream (3 polys), st = t out (or -1), sw = w out (or -1)
constexpr inline void prog_keymatvec(Prog &P, int sv, int st, int sw, bool check_small)
{
    const int fl = check_small ? FWD_CHECK_SMALL : 0;
    P.add(OP_SEG);
    if (st >= 0) {
        P.add(OP_FWD, sv, fl, 0, 1);
        P.add(OP_MACK, 0, 0, MAC_INIT);
    }
    P.add(OP_FWD, sv, fl, 0, 2);
    if (st >= 0) P.add(OP_MACK, 0, 1, 0);
    if (sw >= 0) P.add(OP_MACK, 1, 2, MAC_INIT);
    if (st >= 0) {
        P.add(OP_INV, 0, 0);
        P.add(OP_ADDP, sv, 0, 0, 0);
        P.add(OP_FIN, st, FIN_STORE, 0, 0);
    }
    if (sw >= 0) {
        P.add(OP_INV, 1, 1);
        P.add(OP_ADDP, sv, 0, 0, 1);
        P.add(OP_FIN, sw, FIN_STORE, 0, 0);
    }
}

// check_verify_constraint(z) and A1.z == t + c1*d            open.rs:162-174
// optionally also w = A2.z - c2*d  (left/right sides of the third equation folded
// together, linear.rs:236-249 / sum.rs:300-319).
// sz = z (3 polys), st = t, sc = commitment c (2 polys: c1, c2), sd = d, sw = w out or -1
// With sdh >= 0 the NTT image of d comes from stream sdh (written by prog_challenge_image for the group the item
// belongs to) instead of being transformed per item: the T terms of a Sum proof and the two first equations of a
// Linear proof share one challenge.
// rot = true: c1*d (and c2*d) are not multiplied in the NTT domain but added in the epilogues as signed rotations of the rows
// c1, c2 (OP_ROT, SURVEY kernel K4): two forward transforms per item (z1, z2) instead of four / five (d, z1, z2, c1, c2).
// With w the second accumulator waits in the half warp's global stash region (acc1_global) while the first epilogue's
// rotation sum overlays the warp's shared-memory region.
constexpr inline void prog_verify_first(Prog &P, int sz, int st, int sc, int sd, int sw, int sdh = -1, bool rot = false)
{
    P.add(OP_SEG);
    if (rot) {
        if (sw >= 0) P.acc1_global = 1;
        P.add(OP_FWD, sz, 0, 0, 1);
        P.add(OP_MACK, 0, 0, MAC_INIT);
        P.add(OP_FWD, sz, 0, 0, 2);
        P.add(OP_MACK, 0, 1, 0);
        if (sw >= 0) P.add(OP_MACK, 1, 2, MAC_INIT);
        P.add(OP_INV, 0, 0);
        P.add(OP_ADDP, sz, 0, 0, 0);
        P.add(OP_ADDP, st, 0, MAC_NEG, 0);
        P.add(OP_ROT, sc, sd, MAC_NEG, 0);
        P.add(OP_FIN, 0, FIN_CMPZ, 0, 0);
        if (sw >= 0) {
            P.add(OP_INV, 1, 1);
            P.add(OP_ADDP, sz, 0, 0, 1);
            P.add(OP_ROT, sc, sd, MAC_NEG, 1);
            P.add(OP_FIN, sw, FIN_STORE, 0, 0);
        }
        return;
    }
    if (sdh < 0) {
        P.add(OP_FWD, sd, FWD_SCALED, 0, 0);
        P.add(OP_ST);
    }
    P.add(OP_FWD, sz, 0, 0, 1);
    P.add(OP_MACK, 0, 0, MAC_INIT);
    P.add(OP_FWD, sz, 0, 0, 2);
    P.add(OP_MACK, 0, 1, 0);
    if (sw >= 0) P.add(OP_MACK, 1, 2, MAC_INIT);
    P.add(OP_FWD, sc, 0, 0, 0);
    if (sdh < 0) P.add(OP_MACV, 0, 0, MAC_NEG);
    else P.add(OP_MACG, 0, sdh, MAC_NEG);
    if (sw >= 0) {
        P.add(OP_FWD, sc, 0, 0, 1);
        if (sdh < 0) P.add(OP_MACV, 1, 0, MAC_NEG);
        else P.add(OP_MACG, 1, sdh, MAC_NEG);
    }
    P.add(OP_INV, 0, 0);
    P.add(OP_ADDP, sz, 0, 0, 0);
    P.add(OP_ADDP, st, 0, MAC_NEG, 0);
    P.add(OP_FIN, 0, FIN_CMPZ, 0, 0);
    if (sw >= 0) {
        P.add(OP_INV, 1, 1);
        P.add(OP_ADDP, sz, 0, 0, 1);
        P.add(OP_FIN, sw, FIN_STORE, 0, 0);
    }
}

// NTT image of the challenge d (two primes, Montgomery-operand form) for prog_verify_first(..., sdh)
// streams: sd = d (i8, 1 poly per group), sdh = image out
constexpr inline void prog_challenge_image(Prog &P, int sd, int sdh)
{
    P.add(OP_SEG);
    P.add(OP_FWD, sd, FWD_SCALED, 0, 0);
    P.add(OP_STG, sdh);
}

constexpr inline void prog_norm_verify(Prog &P, int sz)
{
    P.add(OP_NORM, sz, /*verify bound*/ 1, 3, 0);      // params.rs:112-118
}

// Commitment::verify (commit.rs:173-210) as a zero test, 2 primes (r is any int8 row, f any int8 polynomial):
//   f = None:     c - ([a1;a2].r + [0;x]) == 0
//   f = Some(f):  f*c - [a1;a2].r - f*[0;x] == 0      (commit.rs:203-207, rearranged by commutativity of R_q)
// preceded by check_commit_constraint(r) (commit.rs:182).
// streams: sc = c (2 polys), sx = x, sr = r (i8, 3 polys), sf = f (i8) or -1
constexpr inline void prog_commitment_verify(Prog &P, int sc, int sx, int sr, int sf)
{
    P.add(OP_NORM, sr, /*commit bound*/ 0, 3, 0);
    P.add(OP_SEG);
    if (sf >= 0) {
        P.add(OP_FWD, sf, FWD_SCALED, 0, 0);
        P.add(OP_ST);
    }
    P.add(OP_FWD, sr, 0, 0, 1);
    P.add(OP_MACK, 0, 0, MAC_INIT | MAC_NEG);
    P.add(OP_FWD, sr, 0, 0, 2);
    P.add(OP_MACK, 0, 1, MAC_NEG);
    P.add(OP_MACK, 1, 2, MAC_INIT | MAC_NEG);
    if (sf >= 0) {
        P.add(OP_FWD, sc, 0, 0, 0);
        P.add(OP_MACV, 0, 0, 0);
        P.add(OP_FWD, sc, 0, 0, 1);
        P.add(OP_MACV, 1, 0, 0);
        P.add(OP_FWD, sx, 0, 0, 0);
        P.add(OP_MACV, 1, 0, MAC_NEG);
    }
    P.add(OP_INV, 0, 0);
    P.add(OP_ADDP, sr, 0, MAC_NEG, 0);
    if (sf < 0) P.add(OP_ADDP, sc, 0, 0, 0);
    P.add(OP_FIN, 0, FIN_CMPZ, 0, 0);
    P.add(OP_INV, 1, 1);
    P.add(OP_ADDP, sr, 0, MAC_NEG, 1);
    if (sf < 0) {
        P.add(OP_ADDP, sc, 0, 0, 1);
        P.add(OP_ADDP, sx, 0, MAC_NEG, 0);
    }
    P.add(OP_FIN, 0, FIN_CMPZ, 0, 0);
}

// ---- 1-prime program ---------------------------------------------------------------

// z = y + r.componentwise_mul(d)                          open.rs:113-115
// streams: sy = y (3), sr = r (3), sd = d (1), sz = z out (3)
constexpr inline void prog_respond(Prog &P, int sy, int sr, int sd, int sz)
{
    P.add(OP_SEG);
    P.add(OP_FWD, sd, FWD_SCALED, 0, 0);
    P.add(OP_ST);
    P.add(OP_FWD, sr, 0, 0, 0);
    P.add(OP_MACV, 0, 0, MAC_INIT);
    P.add(OP_FWD, sr, 0, 0, 1);
    P.add(OP_MACV, 1, 0, MAC_INIT);
    P.add(OP_INV, 0, 0);
    P.add(OP_ADDP, sy, 0, 0, 0);
    P.add(OP_FIN, sz, FIN_STORE, 0, 0);
    P.add(OP_INV, 1, 0);
    P.add(OP_ADDP, sy, 0, 0, 1);
    P.add(OP_FIN, sz, FIN_STORE, 0, 1);
    P.add(OP_SEG);
    P.add(OP_FWD, sd, FWD_SCALED, 0, 0);
    P.add(OP_ST);
    P.add(OP_FWD, sr, 0, 0, 2);
    P.add(OP_MACV, 0, 0, MAC_INIT);
    P.add(OP_INV, 0, 0);
    P.add(OP_ADDP, sy, 0, 0, 2);
    P.add(OP_FIN, sz, FIN_STORE, 0, 2);
}

// ---- 3-prime program ---------------------------------------------------------------

// out = sum_{i<T} a_i * b_i  -  sub0  -  sub1          (large x large products)
//   g*x (linear.rs:91-95), sum g_i*x_i (sum.rs:107-115), u (linear.rs:124-129, sum.rs:154-160),
//   third-equation check (linear.rs:236-249, sum.rs:300-319) with mode = FIN_CMPZ.
// sa, sb: streams holding T polys per item; ssub0/ssub1: single-poly streams or -1.
constexpr inline void prog_mulsum(Prog &P, int T, int sa, int sb, int ssub0, int ssub1, int sout, int mode)
{
    P.add(OP_SEG);
    P.add(OP_FWD, sb, FWD_SCALED, 0, 0);
    P.add(OP_ST);
    P.add(OP_FWD, sa, 0, 0, 0);
    P.add(OP_MACV, 0, 0, MAC_INIT);
    if (T > 1) {
        P.add(OP_LOOP, 0, 0, 0, T - 1);
        P.add(OP_FWD, sb, FWD_SCALED, 0, 1, 1);
        P.add(OP_ST);
        P.add(OP_FWD, sa, 0, 0, 1, 1);
        P.add(OP_MACV, 0, 0, 0);
        P.add(OP_ENDLOOP);
    }
    P.add(OP_INV, 0, 0);
    if (ssub0 >= 0) P.add(OP_ADDP, ssub0, 0, MAC_NEG, 0);
    if (ssub1 >= 0) P.add(OP_ADDP, ssub1, 0, MAC_NEG, 0);
    P.add(OP_FIN, sout >= 0 ? sout : 0, mode, 0, 0);
}

// Two product sums over the same scalars in one pass (the prover's commit phase, linear.rs:91-95 + 124-129,
// sum.rs:107-115 + 154-160):
//     out0 = sum_{i<T} a_i * b_i            (x' = sum g_i x_i)
//     out1 = sum_{i<T} a_i * c_i  -  sub    (u  = sum g_i (A2.y_i) - A2.y')
// Every a_i is transformed once instead of twice: 3 transforms per term and prime instead of 4.
// sa, sb, sc: streams with T polys per item; ssub: single-poly stream.
constexpr inline void prog_mulsum2(Prog &P, int T, int sa, int sb, int sc, int ssub, int sout0, int sout1)
{
    P.add(OP_SEG);
    P.add(OP_FWD, sa, FWD_SCALED, 0, 0);
    P.add(OP_ST);
    P.add(OP_FWD, sb, 0, 0, 0);
    P.add(OP_MACV, 0, 0, MAC_INIT);
    P.add(OP_FWD, sc, 0, 0, 0);
    P.add(OP_MACV, 1, 0, MAC_INIT);
    if (T > 1) {
        P.add(OP_LOOP, 0, 0, 0, T - 1);
        P.add(OP_FWD, sa, FWD_SCALED, 0, 1, 1);
        P.add(OP_ST);
        P.add(OP_FWD, sb, 0, 0, 1, 1);
        P.add(OP_MACV, 0, 0, 0);
        P.add(OP_FWD, sc, 0, 0, 1, 1);
        P.add(OP_MACV, 1, 0, 0);
        P.add(OP_ENDLOOP);
    }
    P.add(OP_INV, 0, 0);
    P.add(OP_FIN, sout0, FIN_STORE, 0, 0);
    P.add(OP_INV, 1, 1);
    P.add(OP_ADDP, ssub, 0, MAC_NEG, 0);
    P.add(OP_FIN, sout1, FIN_STORE, 0, 0);
}

#ifndef RZK_PRELOAD
#define RZK_PRELOAD 0      // plain-term rows fetched ahead of each inverse transform by the warp-per-item programs: measured NO gain
                           // (profiles/r2_ab_timings.log: 150.6 M commitments/s with 0, 1 and 2 rows; the verify kernels spill with 2), so off
#endif

// ---- compile-time program descriptors (vm_run_static) --------------------------------------
// Stream numbering is the one the engine's dev_* helpers use for the same programs.

struct SPCommitSplitKey {        // streams: 0 = x (i32), 1 = r (i8), 2 = c out
    static constexpr int kPreload = RZK_PRELOAD;
    static constexpr int kNP = 1, kMode = 2 /* MODE_SPLITKEY */;
    static constexpr Prog prog = [] { Prog p; prog_commit_splitkey(p, 0, 1, 2, false); p.end(); return p; }();
    static constexpr uint8_t dtype[kMaxStreams] = {DT_I32, DT_I8, DT_I32};
};

struct SPCommitSplitKeyS {       // the same program modulo the small prime with signed lazy arithmetic (|r| <= 1: Params::default())
    static constexpr int kPreload = RZK_PRELOAD;
    static constexpr int kNP = 1, kMode = 3 /* MODE_SPLITKEY_S */;
    static constexpr Prog prog = [] { Prog p; prog_commit_splitkey(p, 0, 1, 2, false); p.end(); return p; }();
    static constexpr uint8_t dtype[kMaxStreams] = {DT_I32, DT_I8, DT_I32};
};

struct SPKeyMatVecT {            // streams: 0 = y (i32), 1 = t out
    static constexpr int kPreload = RZK_PRELOAD;
    static constexpr int kNP = 2, kMode = 1 /* MODE_SPLIT */;
    static constexpr Prog prog = [] { Prog p; prog_keymatvec(p, 0, 1, -1, true); p.end(); return p; }();
    static constexpr uint8_t dtype[kMaxStreams] = {DT_I32, DT_I32};
};

struct SPKeyMatVecTW {           // streams: 0 = y, 1 = t out, 2 = w out
    static constexpr int kPreload = RZK_PRELOAD;
    static constexpr int kNP = 2, kMode = 1;
    static constexpr Prog prog = [] { Prog p; prog_keymatvec(p, 0, 1, 2, true); p.end(); return p; }();
    static constexpr uint8_t dtype[kMaxStreams] = {DT_I32, DT_I32, DT_I32};
};

struct SPVerifyFirst {           // streams: 0 = z, 1 = t, 2 = c, 3 = d (i8)
    static constexpr int kNP = 2, kMode = 1;
    static constexpr Prog prog = [] { Prog p; prog_norm_verify(p, 0); prog_verify_first(p, 0, 1, 2, 3, -1); p.end(); return p; }();
    static constexpr uint8_t dtype[kMaxStreams] = {DT_I32, DT_I32, DT_I32, DT_I8};
};

struct SPVerifyFirstRot {        // streams: 0 = z, 1 = t, 2 = c, 3 = d (i8); c1*d as signed rotations (OP_ROT)
    static constexpr int kPreload = RZK_PRELOAD;
    static constexpr int kNP = 2, kMode = 1;
    static constexpr Prog prog = [] { Prog p; prog_norm_verify(p, 0); prog_verify_first(p, 0, 1, 2, 3, -1, -1, true); p.end(); return p; }();
    static constexpr uint8_t dtype[kMaxStreams] = {DT_I32, DT_I32, DT_I32, DT_I8};
};

struct SPVerifyFirstWRot {       // streams: 0 = z, 1 = t, 2 = c, 3 = d (i8), 4 = w out; c1*d and c2*d as signed rotations
    static constexpr int kPreload = RZK_PRELOAD;
    static constexpr int kNP = 2, kMode = 1;
    static constexpr bool kAcc1Global = true;
    static constexpr Prog prog = [] { Prog p; prog_norm_verify(p, 0); prog_verify_first(p, 0, 1, 2, 3, 4, -1, true); p.end(); return p; }();
    static constexpr uint8_t dtype[kMaxStreams] = {DT_I32, DT_I32, DT_I32, DT_I8, DT_I32};
};

struct SPVerifyFirstW {          // streams: 0 = z, 1 = t, 2 = c, 3 = d (i8), 4 = w out
    static constexpr int kNP = 2, kMode = 1;
    static constexpr Prog prog = [] { Prog p; prog_norm_verify(p, 0); prog_verify_first(p, 0, 1, 2, 3, 4); p.end(); return p; }();
    static constexpr uint8_t dtype[kMaxStreams] = {DT_I32, DT_I32, DT_I32, DT_I8, DT_I32};
};

struct SPVerifyFirstWG {         // streams: 0 = z, 1 = t, 2 = c, 4 = w out, 5 = NTT image of d (per group)
    static constexpr int kNP = 2, kMode = 1;
    static constexpr Prog prog = [] { Prog p; prog_norm_verify(p, 0); prog_verify_first(p, 0, 1, 2, 3, 4, 5); p.end(); return p; }();
    static constexpr uint8_t dtype[kMaxStreams] = {DT_I32, DT_I32, DT_I32, DT_I8, DT_I32, DT_I32};
};

struct SPChallengeImage {        // streams: 0 = d (i8), 1 = image out
    static constexpr int kNP = 2, kMode = 1;
    static constexpr Prog prog = [] { Prog p; prog_challenge_image(p, 0, 1); p.end(); return p; }();
    static constexpr uint8_t dtype[kMaxStreams] = {DT_I8, DT_I32};
};

struct SPRespond {               // streams: 0 = y, 1 = r (i8), 2 = d (i8), 3 = z out
    static constexpr int kNP = 1, kMode = 0 /* MODE_SEQ */;
    static constexpr Prog prog = [] { Prog p; prog_respond(p, 0, 1, 2, 3); p.end(); return p; }();
    static constexpr uint8_t dtype[kMaxStreams] = {DT_I32, DT_I8, DT_I8, DT_I32};
};

// 3-prime product sums; the loop trip count comes from K.loop_count (= T - 1) at launch.
struct SPMulSum0 {               // streams: 0 = a, 1 = b, 4 = out            out = sum a_i*b_i
    static constexpr int kNP = 3, kMode = 0;
    static constexpr Prog prog = [] { Prog p; prog_mulsum(p, 2, 0, 1, -1, -1, 4, FIN_STORE); p.end(); return p; }();
    static constexpr uint8_t dtype[kMaxStreams] = {DT_I32, DT_I32, DT_I32, DT_I32, DT_I32};
};

struct SPMulSum1 {               // streams: 0 = a, 1 = b, 2 = sub0, 4 = out  out = sum a_i*b_i - sub0
    static constexpr int kNP = 3, kMode = 0;
    static constexpr Prog prog = [] { Prog p; prog_mulsum(p, 2, 0, 1, 2, -1, 4, FIN_STORE); p.end(); return p; }();
    static constexpr uint8_t dtype[kMaxStreams] = {DT_I32, DT_I32, DT_I32, DT_I32, DT_I32};
};

struct SPMulSum2 {               // streams: 0 = a, 1 = b, 2 = c, 3 = sub, 4 = out0, 5 = out1   (prog_mulsum2)
    static constexpr int kNP = 3, kMode = 0;
    static constexpr Prog prog = [] { Prog p; prog_mulsum2(p, 2, 0, 1, 2, 3, 4, 5); p.end(); return p; }();
    static constexpr uint8_t dtype[kMaxStreams] = {DT_I32, DT_I32, DT_I32, DT_I32, DT_I32, DT_I32};
};

// the same product sums modulo the three small primes with signed lazy arithmetic (MODE_SEQ_S, up to kSignedMaxTerms terms)
struct SPMulSum0S { static constexpr int kNP = 3, kMode = 4; static constexpr Prog prog = SPMulSum0::prog; static constexpr uint8_t dtype[kMaxStreams] = {DT_I32, DT_I32, DT_I32, DT_I32, DT_I32}; };
struct SPMulSum1S { static constexpr int kNP = 3, kMode = 4; static constexpr Prog prog = SPMulSum1::prog; static constexpr uint8_t dtype[kMaxStreams] = {DT_I32, DT_I32, DT_I32, DT_I32, DT_I32}; };
struct SPMulSum2S { static constexpr int kNP = 3, kMode = 4; static constexpr Prog prog = SPMulSum2::prog; static constexpr uint8_t dtype[kMaxStreams] = {DT_I32, DT_I32, DT_I32, DT_I32, DT_I32, DT_I32}; };

struct SPMulSumCmp {             // streams: 0 = a, 1 = b, 2 = sub0, 3 = sub1  sum a_i*b_i - sub0 - sub1 == 0
    static constexpr int kNP = 3, kMode = 0;
    static constexpr Prog prog = [] { Prog p; prog_mulsum(p, 2, 0, 1, 2, 3, -1, FIN_CMPZ); p.end(); return p; }();
    static constexpr uint8_t dtype[kMaxStreams] = {DT_I32, DT_I32, DT_I32, DT_I32, DT_I32};
};

struct SPMulSumCmpS { static constexpr int kNP = 3, kMode = 4; static constexpr Prog prog = SPMulSumCmp::prog; static constexpr uint8_t dtype[kMaxStreams] = {DT_I32, DT_I32, DT_I32, DT_I32, DT_I32}; };

}  // namespace rzk
