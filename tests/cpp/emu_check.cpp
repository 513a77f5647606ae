// emu_check.cpp -- host lane emulator for the polynomial-op programs.
//
// Compiles ring-zk_b200/csrc/rzk_vm_exec.cuh with g++ (16 explicit lanes per item) and
// checks every program shape used by the engine against the CPU oracle on seeded random
// inputs.  Test infrastructure only (run by tests/test_emulator.py): it lets the exact
// kernel arithmetic and shared-memory layouts be validated without a GPU.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <vector>
#include <type_traits>
#include <limits.h>
#include <stdint.h>

#include "../../ring-zk_b200/csrc/rzk_vm_exec.cuh"
#include "../../ring-zk_b200/csrc/rzk_programs.h"
#include "../../ring-zk_b200/csrc/rzk_sparse.cuh"
#include "../../ring-zk_b200/csrc/rzk_sample.cuh"
#include "../../ring-zk_b200/csrc/rzk_tables.h"
extern "C" {
#include "../../oracle/ringzk_oracle.h"
}

using namespace rzk;

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static uint64_t rnd()
{
    uint64_t z = (rng_state += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static const int64_t Q = 3515337053LL;
static int32_t rnd_q() { return (int32_t)((int64_t)(rnd() % (uint64_t)Q) - (Q - 1) / 2); }
static double rnd_u() { return ((rnd() >> 11) + 0.5) / 9007199254740992.0; }
static int32_t rnd_gauss(double sigma)
{
    double u1 = rnd_u(), u2 = rnd_u();
    return (int32_t)trunc(sigma * sqrt(-2.0 * log(u1)) * cos(2 * M_PI * u2));
}

struct Emu {
    int np;
    int mode;
    int slots[3];
    std::vector<uint32_t> g1, g2, key, twist, flags;
    VmLaunch K;
    std::vector<uint32_t> region;     // two half-warp regions

    Emu(int np_, const int *sl, const int64_t *keypolys /*[3][512]*/, size_t nflags, int mode_ = -1)
    {
        np = np_;
        mode = mode_ >= 0 ? mode_ : (np == 2 ? MODE_SPLIT : MODE_SEQ);
        memset(&K, 0, sizeof(K));
        g1.assign((size_t)np * 2 * kG1Words, 0);
        g2.resize((size_t)np * 2 * kLanes * kG2Words);
        twist.assign((size_t)np * kTwistWords, 0);
        key.resize((size_t)np * kKeyPolys * 2 * kPadWords * (mode_sk(mode) ? 2 : 1));
        for (int i = 0; i < np; ++i) {
            slots[i] = sl[i];
            const PrimeTables &T = prime_tables(sl[i]);
            for (int d = 0; d < 2; ++d) memcpy(&g1[((size_t)i * 2 + d) * kG1Words], T.g1[d], sizeof(T.g1[d]));
            memcpy(&g2[(size_t)i * 2 * kLanes * kG2Words], T.g2, sizeof(T.g2));
            for (int k = 0; k < kKeyPolys; ++k) {
                if (mode == MODE_SPLITKEY_S) key_image_split_signed(T, keypolys + (size_t)k * kN, &key[(size_t)k * 4 * kPadWords]);
                else if (mode == MODE_SPLITKEY) key_image_split(T, keypolys + (size_t)k * kN, &key[(size_t)k * 4 * kPadWords]);
                else key_image(T, keypolys + (size_t)k * kN, &key[((size_t)i * kKeyPolys + k) * 2 * kPadWords]);
            }
            K.pc[i] = make_prime_consts(sl[i]);
            memcpy(&twist[(size_t)i * kTwistWords], mode == MODE_SEQ_S ? T.twist_rn : T.twist, sizeof(T.twist));
        }
        K.crt = make_crt_consts(sl, np, (uint64_t)Q);
        K.q = (uint32_t)Q;
        K.kqh = ((uint64_t)Q << 29) + (uint64_t)(Q - 1) / 2;
        K.qd = (double)Q; K.qinvd = 1.0 / (double)Q;
        K.p0d = (double)K.pc[0].p; K.p0qinvd = (double)K.pc[0].p / (double)Q; K.p0invd = 1.0 / (double)K.pc[0].p;
        K.m30 = (uint32_t)((1ull << 62) / (uint64_t)Q);
        rzko_params P = rzko_default_params(kN);
        uint64_t cb = rzko_commit_bound(&P), vb = rzko_verify_bound(&P);
        K.norm_abs_lim[0] = (uint32_t)cb; K.norm_sq_lim[0] = (cb + 1) * (cb + 1) - 1;
        K.norm_abs_lim[1] = (uint32_t)vb; K.norm_sq_lim[1] = (vb + 1) * (vb + 1) - 1;
        K.small_lim = 1u << 18;
        K.np = np;
        K.flag_div = 1;
        flags.assign(nflags, 0);
        K.flags = flags.data();
    }
    void stream(int i, const void *base, uint32_t stride, uint32_t dtype, uint32_t div = 1)
    {
        K.st[i].base = base; K.st[i].stride = stride; K.st[i].dtype = dtype; set_stream_div(K.st[i], div);
    }
    // Emulates one warp at a time exactly as the kernel maps it: SPLIT (np == 2) gives the warp one
    // item with half warp h on prime h; SEQ gives each half warp its own item.
    template <class SP = void>
    void run(uint32_t n_items)
    {
        K.n_items = n_items;
        const bool split = !mode_seq(mode);
        static std::vector<uint32_t> gst;
        if (K.acc1_global) { gst.assign((size_t)2 * kSlotWords, 0xDEADBEEFu); K.gstash = gst.data(); }   // as the kernel: accumulator 1 outside the region
        layout_hw(K, split);
        if (K.hw_words % 32 != 16) { printf("FAIL: hw_words %u not 16 mod 32\n", K.hw_words); exit(1); }
        region.assign((size_t)2 * K.hw_words, 0xDEADBEEFu);
        static Lane lanes[32];
        static LaneCtx ctxs[32];
        const uint32_t per_warp = split ? 1 : 2;
        for (uint32_t base = 0; base < n_items; base += per_warp) {
            for (int li = 0; li < 32; ++li) {
                LaneCtx &c = ctxs[li];
                const int hw = li >> 4;
                uint32_t *mine = region.data() + (size_t)hw * K.hw_words;
                c.slot_hw[0] = region.data() + K.off_slot; c.slot_hw[1] = region.data() + K.hw_words + K.off_slot;
                c.buf = mine; c.slot = mine + K.off_slot; c.stash = K.gstash ? K.gstash + (size_t)hw * K.stash_words : mine + K.off_stash;
                c.acc1 = K.acc1_global ? c.stash : mine + K.off_acc1;
                c.red = split ? region.data() : mine;
                c.ridx = split ? li : (li & 15);
                c.g1 = g1.data(); c.g2 = g2.data(); c.key = key.data(); c.twist = twist.data();
                c.t = li & 15; c.hw = hw;
                const uint32_t item = base + (split ? 0 : hw);
                c.active = item < n_items;
                c.item = c.active ? item : n_items - 1;
            }
            if constexpr (!std::is_void<SP>::value) vm_run_static<SP>(K, lanes, ctxs);
            else if (mode == MODE_SPLITKEY) vm_run_item<1, MODE_SPLITKEY>(K, lanes, ctxs);
            else if (mode == MODE_SPLITKEY_S) vm_run_item<1, MODE_SPLITKEY_S>(K, lanes, ctxs);
            else if (mode == MODE_SEQ_S) vm_run_item<3, MODE_SEQ_S>(K, lanes, ctxs);
            else if (np == 1) vm_run_item<1, MODE_SEQ>(K, lanes, ctxs);
            else if (np == 2) vm_run_item<2, MODE_SPLIT>(K, lanes, ctxs);
            else vm_run_item<3, MODE_SEQ>(K, lanes, ctxs);
        }
    }
};

static int nfail = 0;
#define CHECK(cond, ...) do { if (!(cond)) { printf("FAIL %s:%d: ", __FILE__, __LINE__); printf(__VA_ARGS__); printf("\n"); ++nfail; } } while (0)

static std::vector<int64_t> widen(const std::vector<int32_t> &v) { return std::vector<int64_t>(v.begin(), v.end()); }
static std::vector<int64_t> widen8(const std::vector<int8_t> &v) { return std::vector<int64_t>(v.begin(), v.end()); }
static bool same(const std::vector<int32_t> &a, const std::vector<int64_t> &b)
{
    if (a.size() != b.size()) return false;
    for (size_t i = 0; i < a.size(); ++i) if ((int64_t)a[i] != b[i]) return false;
    return true;
}

int main(int argc, char **argv)
{
    const int B = argc > 1 ? atoi(argv[1]) : 3;
    const int T = 3;
    rzko_params P = rzko_default_params(kN);
    const size_t N = kN;

    // key
    std::vector<int64_t> keyp(3 * N);
    for (auto &v : keyp) v = rnd_q();
    std::vector<int64_t> a1(3 * N), a2(3 * N);
    rzko_key_expand(&P, keyp.data(), keyp.data() + 2 * N, a1.data(), a2.data());

    // inputs
    std::vector<int32_t> x(B * N), g(B * N), y(B * 3 * N), yp(B * 3 * N);
    std::vector<int8_t> r(B * 3 * N), rp(B * 3 * N), d(B * N, 0);
    for (auto &v : x) v = rnd_q();
    for (auto &v : g) v = rnd_q();
    for (auto &v : r) v = (int8_t)((int)(rnd() % 3) - 1);
    for (auto &v : rp) v = (int8_t)((int)(rnd() % 3) - 1);
    for (auto &v : y) v = rnd_gauss(15444.0);
    for (auto &v : yp) v = rnd_gauss(15444.0);
    for (int b = 0; b < B; ++b)
        for (int k = 0; k < 36;) {
            int pos = (int)(rnd() % N);
            if (d[b * N + pos]) continue;
            d[b * N + pos] = (rnd() & 1) ? 1 : -1;
            ++k;
        }
    // ragged message: zero tail on item 0 (tests/test.rs:95-99)
    for (size_t i = 5; i < N; ++i) x[i] = 0;
    auto x64 = widen(x), g64 = widen(g), y64 = widen(y), yp64 = widen(yp);
    auto r64 = widen8(r), rp64 = widen8(rp), d64 = widen8(d);

    const int L2[3] = {0, 1, 2};

    // ---------------- commit + open commit (2 primes) ----------------
    std::vector<int64_t> c_o(B * 2 * N), t_o(B * N);
    std::vector<uint8_t> ok_o(B);
    rzko_open_commit_batch(&P, a1.data(), a2.data(), B, x64.data(), r64.data(), y64.data(), c_o.data(), t_o.data(), ok_o.data(), 1);
    std::vector<int32_t> c_e(B * 2 * N), t_e(B * N), w_e(B * N);
    for (int ld128 = 1; ld128 >= 0; --ld128) {       // both forms of OP_FWD's int32 loads (the default last: its outputs are reused below)
        Emu E(2, L2, keyp.data(), B);
        E.K.ld128 = (uint32_t)ld128;
        Prog pr;
        prog_commit(pr, 0, 1, 2);
        prog_keymatvec(pr, 3, 4, 5, true);
        pr.end(); pr.install(E.K);
        E.stream(0, x.data(), 1, DT_I32); E.stream(1, r.data(), 3, DT_I8); E.stream(2, c_e.data(), 2, DT_I32);
        E.stream(3, y.data(), 3, DT_I32); E.stream(4, t_e.data(), 1, DT_I32); E.stream(5, w_e.data(), 1, DT_I32);
        E.run(B);
        CHECK(same(c_e, c_o), "commit c mismatch");
        CHECK(same(t_e, t_o), "open commit t mismatch");
        for (int b = 0; b < B; ++b) CHECK(E.flags[b] == 0 && ok_o[b] == 1, "commit flags item %d: %u", b, E.flags[b]);
        // w = A2.y against the oracle's mat_dot
        std::vector<int64_t> w_o(N);
        for (int b = 0; b < B; ++b) {
            rzko_mat_dot(&P, 1, 3, 1, a2.data(), y64.data() + (size_t)b * 3 * N, w_o.data());
            for (size_t i = 0; i < N; ++i) CHECK(w_o[i] == w_e[b * N + i], "w mismatch item %d coef %zu", b, i);
        }
        printf("commit/open_commit: ops=%d\n", pr.n);
    }
    // commit-constraint failure and range flag
    {
        std::vector<int32_t> rbig(B * 3 * N, 0), ybig(y);
        for (int b = 0; b < B; ++b) rbig[(b * 3 + 1) * N + 7] = 1359073;        // > commit bound
        ybig[5] = (1 << 18) + 1;                                                 // item 0, poly 0: not transformed -> no flag
        ybig[N + 5] = (1 << 18) + 1;                                             // item 0, poly 1: flagged
        Emu E(2, L2, keyp.data(), B);
        Prog pr;
        prog_commit(pr, 0, 1, 2);
        prog_keymatvec(pr, 3, 4, -1, true);
        pr.end(); pr.install(E.K);
        E.stream(0, x.data(), 1, DT_I32); E.stream(1, rbig.data(), 3, DT_I32); E.stream(2, c_e.data(), 2, DT_I32);
        E.stream(3, ybig.data(), 3, DT_I32); E.stream(4, t_e.data(), 1, DT_I32);
        E.run(B);
        CHECK(E.flags[0] == (FLAG_FAIL | FLAG_RANGE), "flags[0]=%u", E.flags[0]);
        for (int b = 1; b < B; ++b) CHECK(E.flags[b] == FLAG_FAIL, "flags[%d]=%u", b, E.flags[b]);
        auto rb64 = widen(rbig);
        std::vector<int64_t> c2(B * 2 * N);
        std::vector<uint8_t> ok2(B);
        rzko_commit_batch(&P, a1.data(), a2.data(), B, x64.data(), rb64.data(), c2.data(), ok2.data(), 1);
        CHECK(same(c_e, c2), "commit with large r mismatch");
        for (int b = 0; b < B; ++b) CHECK(ok2[b] == 0, "oracle ok");
    }

    // ---------------- split-key commit (1 prime, key halves) ----------------
    {
        const int L1[1] = {0};
        std::vector<int8_t> r15(r);
        for (size_t i = 0; i < r15.size(); ++i) r15[i] = (int8_t)((int)(rnd() % 31) - 15);      // |r| <= 15: worst case of the bound
        for (size_t i = N; i < 3 * N; ++i) r15[i] = (i & 1) ? 15 : -15;
        std::vector<int32_t> xw(x);
        Emu E(1, L1, keyp.data(), B, MODE_SPLITKEY);
        Prog pr;
        prog_commit_splitkey(pr, 0, 1, 2);
        pr.end(); pr.install(E.K);
        E.K.small_lim = 15;
        E.stream(0, xw.data(), 1, DT_I32); E.stream(1, r15.data(), 3, DT_I8); E.stream(2, c_e.data(), 2, DT_I32);
        E.run(B);
        auto r15_64 = widen8(r15);
        std::vector<int64_t> c2(B * 2 * N);
        std::vector<uint8_t> ok2(B);
        rzko_commit_batch(&P, a1.data(), a2.data(), B, x64.data(), r15_64.data(), c2.data(), ok2.data(), 1);
        CHECK(same(c_e, c2), "split-key commit mismatch");
        for (int b = 0; b < B; ++b) CHECK(E.flags[b] == 0, "split-key flags[%d]=%u", b, E.flags[b]);
        // |r| = 16 on a transformed row must raise the range flag; on row 0 it must not
        r15[2 * N + 3] = 16;
        r15[(size_t)3 * N + 5] = 100;      // item 1, row 0 (never transformed)
        std::vector<uint32_t> f0(B, 0);
        E.flags.assign(B, 0); E.K.flags = E.flags.data();
        E.run(B);
        CHECK(E.flags[0] == FLAG_RANGE, "split-key range flag item 0: %u", E.flags[0]);
        if (B > 1) CHECK(E.flags[1] == 0, "split-key range flag item 1: %u", E.flags[1]);
        // with a mark array (the host entry points' masked redo) the range error goes there and the flags stay clean
        std::vector<uint32_t> rmark(B + 1, 0);
        E.flags.assign(B, 0); E.K.flags = E.flags.data();
        E.K.rmark = rmark.data(); E.K.rmark_any = rmark.data() + B;
        E.run(B);
        CHECK(E.flags[0] == 0 && rmark[0] == 1 && rmark[B] == 1, "range mark: flags %u mark %u any %u", E.flags[0], rmark[0], rmark[B]);
        for (int b = 1; b < B; ++b) CHECK(rmark[b] == 0, "range mark item %d", b);
        printf("split-key commit ok, ops=%d\n", pr.n);
    }

    // ---------------- respond (1 prime) ----------------
    std::vector<int64_t> z_o(B * 3 * N), zp_o(B * 3 * N);
    rzko_linear_respond_batch(&P, B, y64.data(), yp64.data(), r64.data(), rp64.data(), d64.data(), z_o.data(), zp_o.data(), 1);
    std::vector<int32_t> z_e(B * 3 * N), zp_e(B * 3 * N);
    {
        const int L1[1] = {0};
        Emu E(1, L1, keyp.data(), B);
        Prog pr;
        prog_respond(pr, 0, 1, 2, 3);
        prog_respond(pr, 4, 5, 2, 6);
        pr.end(); pr.install(E.K);
        E.stream(0, y.data(), 3, DT_I32); E.stream(1, r.data(), 3, DT_I8); E.stream(2, d.data(), 1, DT_I8);
        E.stream(3, z_e.data(), 3, DT_I32);
        E.stream(4, yp.data(), 3, DT_I32); E.stream(5, rp.data(), 3, DT_I8); E.stream(6, zp_e.data(), 3, DT_I32);
        E.run(B);
        CHECK(same(z_e, z_o), "respond z mismatch");
        CHECK(same(zp_e, zp_o), "respond zp mismatch");
        printf("respond: ops=%d\n", pr.n);
    }

    // ---------------- open verify (2 primes): c1*d in the NTT domain, and as signed rotations (OP_ROT) ----------------
    for (int rot = 0; rot < 2; ++rot) {
        std::vector<int32_t> c32(c_o.begin(), c_o.end()), t32(t_o.begin(), t_o.end());
        for (int variant = 0; variant < 8; ++variant) {
            std::vector<int32_t> zz(z_e), tt(t32), cc(c32);
            std::vector<int8_t> dd(d);
            if (variant == 1) for (int b = 0; b < B; ++b) zz[(b * 3 + 2) * N + 11] += 1;
            if (variant == 2) for (int b = 0; b < B; ++b) tt[b * N + 500] -= 1;
            if (variant == 3) for (int b = 0; b < B; ++b) cc[(b * 2) * N + 1] += 1;
            if (variant == 4) for (int b = 0; b < B; ++b) zz[(b * 3) * N + 3] = 679537;   // norm check
            if (variant == 5) for (int b = 0; b < B; ++b) dd[b * N + 511] = (int8_t)(dd[b * N + 511] ? 0 : 1);   // one more / one fewer rotation
            // a dense challenge with arbitrary int8 entries and non-canonical / extreme representatives of c1: the verdict is
            // `false` (the transcript was made for another d), the point is that both lowerings stay exact and agree with the oracle
            if (variant == 6) for (size_t i = 0; i < dd.size(); ++i) dd[i] = (int8_t)((int)(rnd() % 255) - 127);
            if (variant == 7) {
                for (size_t i = 0; i < dd.size(); ++i) dd[i] = (i & 1) ? 127 : -128;
                for (int b = 0; b < B; ++b) for (size_t i = 0; i < N; ++i) cc[(size_t)b * 2 * N + i] = (i % 3 == 0) ? INT32_MIN : (i % 3 == 1) ? INT32_MAX : (int32_t)((Q - 1) / 2);
            }
            Emu E(2, L2, keyp.data(), B);
            Prog pr;
            prog_norm_verify(pr, 0);
            prog_verify_first(pr, 0, 1, 2, 3, -1, -1, rot != 0);
            pr.end(); pr.install(E.K);
            E.stream(0, zz.data(), 3, DT_I32); E.stream(1, tt.data(), 1, DT_I32);
            E.stream(2, cc.data(), 2, DT_I32); E.stream(3, dd.data(), 1, DT_I8);
            CHECK(rot_layout_ok(E.K.ops, true), "rot layout");
            if (rot) E.run<SPVerifyFirstRot>(B); else E.run(B);      // variants 6, 7: the general (DFMA) form of the rotation sum
            auto z64 = widen(zz), t64 = widen(tt), c64 = widen(cc), dd64 = widen8(dd);
            std::vector<int64_t> c1(B * N);
            for (int b = 0; b < B; ++b) for (size_t i = 0; i < N; ++i) c1[b * N + i] = rzko_center(c64[(size_t)b * 2 * N + i], Q);
            std::vector<uint8_t> okv(B);
            rzko_open_verify_batch(&P, a1.data(), B, z64.data(), t64.data(), c1.data(), dd64.data(), okv.data(), 1);
            for (int b = 0; b < B; ++b) {
                CHECK((E.flags[b] == 0) == (okv[b] == 1), "open verify rot %d variant %d item %d: emu flags %u oracle %u", rot, variant, b, E.flags[b], okv[b]);
                CHECK((okv[b] == 1) == (variant == 0), "oracle verdict variant %d", variant);
            }
        }
        // the rotation sum itself, bit for bit: with z = 0 the program compares -t - c1*d with zero, so feed t = -c1*d
        // (from the oracle's product) and expect "verified" for a dense ternary d, and "failed" after a one-unit change of t
        if (rot) {
            std::vector<int32_t> zz(B * 3 * N, 0), cc(c32);
            std::vector<int8_t> dd(B * N);
            for (auto &v : dd) v = (int8_t)((int)(rnd() % 3) - 1);              // dense challenge: ~340 rotations, unbalanced signs
            for (size_t i = 0; i < N; ++i) dd[i] = (i < 300) ? 1 : (i < 310 ? -1 : 0);    // item 0: 300 additions, 10 subtractions
            for (size_t i = 0; i < N; ++i) cc[i] = (i % 3 == 0) ? INT32_MIN : (i % 3 == 1) ? INT32_MAX : (int32_t)((Q - 1) / 2);   // extreme representatives
            auto c64 = widen(cc), dd64 = widen8(dd);
            for (auto &v : c64) v = rzko_center(v, Q);
            std::vector<int32_t> tt(B * N);
            std::vector<int64_t> prod(N);
            for (int b = 0; b < B; ++b) {
                rzko_poly_mul(&P, c64.data() + (size_t)b * 2 * N, dd64.data() + (size_t)b * N, prod.data());
                for (size_t i = 0; i < N; ++i) tt[b * N + i] = (int32_t)rzko_center(-prod[i], Q);
            }
            for (int bad = 0; bad < 2; ++bad) {
                if (bad) for (int b = 0; b < B; ++b) tt[b * N + (37 * b) % N] += 1;
                Emu E(2, L2, keyp.data(), B);
                Prog pr;
                prog_norm_verify(pr, 0);
                prog_verify_first(pr, 0, 1, 2, 3, -1, -1, true);
                pr.end(); pr.install(E.K);
                E.stream(0, zz.data(), 3, DT_I32); E.stream(1, tt.data(), 1, DT_I32);
                E.stream(2, cc.data(), 2, DT_I32); E.stream(3, dd.data(), 1, DT_I8);
                E.run<SPVerifyFirstRot>(B);
                for (int b = 0; b < B; ++b) CHECK((E.flags[b] == 0) == (bad == 0), "rotation sum vs oracle product: item %d bad %d flags %u", b, bad, E.flags[b]);
            }
        }
        printf("open verify ok (rot=%d)\n", rot);
    }

    // ---------------- linear proof: full lowering vs oracle ----------------
    {
        std::vector<int64_t> gx_o(B * N), cp_o(B * 2 * N), c2_o(B * 2 * N), tl_o(B * N), tp_o(B * N), u_o(B * N);
        std::vector<uint8_t> okl(B);
        rzko_linear_commit_batch(&P, a1.data(), a2.data(), B, g64.data(), x64.data(), rp64.data(), r64.data(), y64.data(), yp64.data(),
                                 gx_o.data(), cp_o.data(), c2_o.data(), tl_o.data(), tp_o.data(), u_o.data(), okl.data(), 1);
        // launch A (3 primes): gx = g*x
        std::vector<int32_t> gx_e(B * N), cp_e(B * 2 * N), cl_e(B * 2 * N), tl_e(B * N), tp_e(B * N), w_l(B * N), wp_l(B * N), u_e(B * N);
        {
            Emu E(3, L2, keyp.data(), B);
            Prog pr;
            prog_mulsum(pr, 1, 0, 1, -1, -1, 2, FIN_STORE);
            pr.end(); pr.install(E.K);
            E.stream(0, g.data(), 1, DT_I32); E.stream(1, x.data(), 1, DT_I32); E.stream(2, gx_e.data(), 1, DT_I32);
            E.run(B);
            CHECK(same(gx_e, gx_o), "linear gx mismatch");
        }
        // launch B (2 primes): both commits, t, tp, w, wp
        {
            Emu E(2, L2, keyp.data(), B);
            Prog pr;
            prog_commit(pr, 0, 1, 2);
            prog_commit(pr, 3, 4, 5);
            prog_keymatvec(pr, 6, 7, 8, true);
            prog_keymatvec(pr, 9, 10, 11, true);
            pr.end(); pr.install(E.K);
            CHECK(pr.n <= kMaxOps, "too many ops %d", pr.n);
            E.stream(0, gx_e.data(), 1, DT_I32); E.stream(1, rp.data(), 3, DT_I8); E.stream(2, cp_e.data(), 2, DT_I32);
            E.stream(3, x.data(), 1, DT_I32); E.stream(4, r.data(), 3, DT_I8); E.stream(5, cl_e.data(), 2, DT_I32);
            E.stream(6, y.data(), 3, DT_I32); E.stream(7, tl_e.data(), 1, DT_I32); E.stream(8, w_l.data(), 1, DT_I32);
            E.stream(9, yp.data(), 3, DT_I32); E.stream(10, tp_e.data(), 1, DT_I32); E.stream(11, wp_l.data(), 1, DT_I32);
            E.run(B);
            CHECK(same(cp_e, cp_o) && same(cl_e, c2_o), "linear commitments mismatch");
            CHECK(same(tl_e, tl_o) && same(tp_e, tp_o), "linear t/tp mismatch");
            printf("linear commit launch B: ops=%d\n", pr.n);
        }
        // launch C (3 primes): u = g*w - wp
        {
            Emu E(3, L2, keyp.data(), B);
            Prog pr;
            prog_mulsum(pr, 1, 0, 1, 2, -1, 3, FIN_STORE);
            pr.end(); pr.install(E.K);
            E.stream(0, g.data(), 1, DT_I32); E.stream(1, w_l.data(), 1, DT_I32); E.stream(2, wp_l.data(), 1, DT_I32);
            E.stream(3, u_e.data(), 1, DT_I32);
            E.run(B);
            CHECK(same(u_e, u_o), "linear u mismatch");
        }
        // verify: launch A (2 primes) + launch B (3 primes), honest and tampered
        for (int variant = 0; variant < 4; ++variant) {
            std::vector<int32_t> zz(z_e), zzp(zp_e), uu(u_e), gg(g);
            if (variant == 1) for (int b = 0; b < B; ++b) uu[b * N + 9] += 1;
            if (variant == 2) for (int b = 0; b < B; ++b) gg[b * N + 100] ^= 1;
            if (variant == 3) for (int b = 0; b < B; ++b) zzp[(b * 3 + 1) * N + 17] -= 1;
            std::vector<uint32_t> flags(B, 0);
            std::vector<int32_t> wv(B * N), wvp(B * N);
            {
                Emu E(2, L2, keyp.data(), B);
                Prog pr;
                prog_norm_verify(pr, 0);
                prog_norm_verify(pr, 4);
                prog_verify_first(pr, 0, 1, 2, 3, 8);
                prog_verify_first(pr, 4, 5, 6, 3, 9);
                pr.end(); pr.install(E.K);
                CHECK(pr.n <= kMaxOps, "too many ops %d", pr.n);
                E.stream(0, zz.data(), 3, DT_I32); E.stream(1, tl_e.data(), 1, DT_I32); E.stream(2, cl_e.data(), 2, DT_I32);
                E.stream(3, d.data(), 1, DT_I8);
                E.stream(4, zzp.data(), 3, DT_I32); E.stream(5, tp_e.data(), 1, DT_I32); E.stream(6, cp_e.data(), 2, DT_I32);
                E.stream(8, wv.data(), 1, DT_I32); E.stream(9, wvp.data(), 1, DT_I32);
                E.run(B);
                flags = E.flags;
                if (variant == 0) printf("linear verify launch A: ops=%d\n", pr.n);
            }
            // the same first equations with c1*d, c2*d as rotation sums (OP_ROT, accumulator 1 in the global stash region):
            // the unrolled program for (z, t, c) and the runtime-decoded one for (z', t', c'); identical flags and w, w'
            {
                std::vector<int32_t> wr(B * N), wrp(B * N);
                Emu E(2, L2, keyp.data(), B);
                SPVerifyFirstWRot::prog.install(E.K);
                E.stream(0, zz.data(), 3, DT_I32); E.stream(1, tl_e.data(), 1, DT_I32); E.stream(2, cl_e.data(), 2, DT_I32);
                E.stream(3, d.data(), 1, DT_I8); E.stream(4, wr.data(), 1, DT_I32);
                CHECK(E.K.acc1_global == 1 && rot_layout_ok(E.K.ops, true, true) && !rot_layout_ok(E.K.ops, true, false), "rot layout (w)");
                E.run<SPVerifyFirstWRot>(B);
                Emu E2(2, L2, keyp.data(), B);
                Prog pr;
                prog_norm_verify(pr, 0);
                prog_verify_first(pr, 0, 1, 2, 3, 4, -1, true);
                pr.end(); pr.install(E2.K);
                E2.stream(0, zzp.data(), 3, DT_I32); E2.stream(1, tp_e.data(), 1, DT_I32); E2.stream(2, cp_e.data(), 2, DT_I32);
                E2.stream(3, d.data(), 1, DT_I8); E2.stream(4, wrp.data(), 1, DT_I32);
                E2.run(B);
                for (int b = 0; b < B; ++b)
                    CHECK((E.flags[b] | E2.flags[b]) == flags[b], "linear verify, rotation sums: variant %d item %d flags %u|%u vs %u", variant, b, E.flags[b], E2.flags[b], flags[b]);
                CHECK(wr == wv && wrp == wvp, "linear verify, rotation sums: w / w' differ from the NTT-domain lowering (variant %d)", variant);
            }
            {
                Emu E(3, L2, keyp.data(), B);
                Prog pr;
                prog_mulsum(pr, 1, 0, 1, 2, 3, -1, FIN_CMPZ);
                pr.end(); pr.install(E.K);
                E.stream(0, gg.data(), 1, DT_I32); E.stream(1, wv.data(), 1, DT_I32); E.stream(2, wvp.data(), 1, DT_I32);
                E.stream(3, uu.data(), 1, DT_I32);
                E.run(B);
                for (int b = 0; b < B; ++b) flags[b] |= E.flags[b];
            }
            auto z64 = widen(zz), zp64 = widen(zzp), u64 = widen(uu), gg64 = widen(gg);
            std::vector<uint8_t> okv(B);
            rzko_linear_verify_batch(&P, a1.data(), a2.data(), B, z64.data(), zp64.data(), c2_o.data(), cp_o.data(), gg64.data(),
                                     tl_o.data(), tp_o.data(), u64.data(), d64.data(), okv.data(), 1);
            for (int b = 0; b < B; ++b) {
                CHECK((flags[b] == 0) == (okv[b] == 1), "linear verify variant %d item %d: emu %u oracle %u", variant, b, flags[b], okv[b]);
                CHECK((okv[b] == 1) == (variant == 0), "oracle linear verdict variant %d", variant);
            }
        }
        printf("linear ok\n");
    }

    // ---------------- sum of T large products (3 primes, looped program) ----------------
    {
        std::vector<int32_t> gs(B * T * N), xs(B * T * N), sub(B * N), out_e(B * N);
        for (auto &v : gs) v = rnd_q();
        for (auto &v : xs) v = rnd_q();
        for (auto &v : sub) v = rnd_q();
        // worst-case magnitudes on item 0 to exercise the 3-prime range
        for (size_t i = 0; i < T * N; ++i) { gs[i] = (i & 1) ? (int32_t)((Q - 1) / 2) : -(int32_t)((Q - 1) / 2); xs[i] = (int32_t)((Q - 1) / 2); }
        // non-canonical int32 representatives (any representative of a class mod q is accepted)
        if (B > 1) for (size_t i = 0; i < N; ++i) {
            gs[(size_t)T * N + i] = (i % 3 == 0) ? INT32_MIN : (i % 3 == 1) ? INT32_MAX : (int32_t)(-2147000001 + (int)(i % 7));
            xs[(size_t)T * N + i] = (i & 1) ? INT32_MAX : INT32_MIN;
            sub[N + i] = (i & 2) ? INT32_MAX : INT32_MIN;
        }
        Emu E(3, L2, keyp.data(), B);
        Prog pr;
        prog_mulsum(pr, T, 0, 1, 2, -1, 3, FIN_STORE);
        pr.end(); pr.install(E.K);
        E.stream(0, gs.data(), T, DT_I32); E.stream(1, xs.data(), T, DT_I32); E.stream(2, sub.data(), 1, DT_I32);
        E.stream(3, out_e.data(), 1, DT_I32);
        E.run(B);
        auto gs64 = widen(gs), xs64 = widen(xs), sub64 = widen(sub);
        for (auto *v : {&gs64, &xs64, &sub64}) for (auto &c : *v) c = rzko_center(c, Q);
        std::vector<int64_t> acc(N), tmp(N);
        for (int b = 0; b < B; ++b) {
            for (int i = 0; i < T; ++i) {
                rzko_poly_mul(&P, xs64.data() + ((size_t)b * T + i) * N, gs64.data() + ((size_t)b * T + i) * N, i ? tmp.data() : acc.data());
                if (i) rzko_poly_add(&P, acc.data(), tmp.data(), acc.data());
            }
            rzko_poly_sub(&P, acc.data(), sub64.data() + (size_t)b * N, acc.data());
            for (size_t i = 0; i < N; ++i) CHECK(acc[i] == out_e[b * N + i], "mulsum item %d coef %zu: %lld vs %d", b, i, (long long)acc[i], out_e[b * N + i]);
        }
        // the compile-time variant of the same program (loop count from K.loop_count)
        {
            std::vector<int32_t> out_s(B * N, 0);
            Emu ES(3, L2, keyp.data(), B);
            SPMulSum1::prog.install(ES.K);
            ES.K.loop_count = T - 1;
            ES.stream(0, gs.data(), T, DT_I32); ES.stream(1, xs.data(), T, DT_I32); ES.stream(2, sub.data(), 1, DT_I32);
            ES.stream(4, out_s.data(), 1, DT_I32);
            ES.run<SPMulSum1>(B);
            CHECK(out_s == out_e, "static mulsum differs from the interpreted one");
            // T = 1 through the same looped program (zero trips)
            std::vector<int32_t> o1(B * N, 0), o2(B * N, 0);
            Emu E1(3, L2, keyp.data(), B);
            SPMulSum0::prog.install(E1.K);
            E1.K.loop_count = 0;
            E1.stream(0, gs.data(), T, DT_I32); E1.stream(1, xs.data(), T, DT_I32); E1.stream(4, o1.data(), 1, DT_I32);
            E1.run<SPMulSum0>(B);
            for (int b = 0; b < B; ++b) {
                rzko_poly_mul(&P, xs64.data() + (size_t)b * T * N, gs64.data() + (size_t)b * T * N, tmp.data());
                for (size_t i = 0; i < N; ++i) CHECK(tmp[i] == o1[b * N + i], "static mulsum T=1 mismatch");
            }
        }
        // two product sums over the same scalars in one pass (prog_mulsum2): out0 = sum g_i*x_i, out1 = sum g_i*c_i - sub,
        // static and interpreted, against the single-output program run twice; T = 3 and T = 1
        for (int TT : {T, 1}) {
            std::vector<int32_t> cs2(xs.size());
            for (size_t i = 0; i < cs2.size(); ++i) cs2[i] = xs[(i * 7 + 3) % xs.size()];
            std::vector<int32_t> r0(B * N, 0), r1(B * N, 0);
            {
                Emu EA(3, L2, keyp.data(), B);
                Prog pa; prog_mulsum(pa, TT, 0, 1, -1, -1, 2, FIN_STORE); pa.end(); pa.install(EA.K);
                EA.stream(0, gs.data(), T, DT_I32); EA.stream(1, xs.data(), T, DT_I32); EA.stream(2, r0.data(), 1, DT_I32);
                EA.run(B);
                Emu EB(3, L2, keyp.data(), B);
                Prog pb; prog_mulsum(pb, TT, 0, 1, 2, -1, 3, FIN_STORE); pb.end(); pb.install(EB.K);
                EB.stream(0, gs.data(), T, DT_I32); EB.stream(1, cs2.data(), T, DT_I32); EB.stream(2, sub.data(), 1, DT_I32);
                EB.stream(3, r1.data(), 1, DT_I32);
                EB.run(B);
            }
            for (int pass = 0; pass < 2; ++pass) {
                std::vector<int32_t> o0(B * N, 0), o1(B * N, 0);
                Emu E2(3, L2, keyp.data(), B);
                if (pass == 0) { Prog p2; prog_mulsum2(p2, TT, 0, 1, 2, 3, 4, 5); p2.end(); p2.install(E2.K); }
                else { SPMulSum2::prog.install(E2.K); E2.K.loop_count = TT - 1; }
                E2.stream(0, gs.data(), T, DT_I32); E2.stream(1, xs.data(), T, DT_I32); E2.stream(2, cs2.data(), T, DT_I32);
                E2.stream(3, sub.data(), 1, DT_I32); E2.stream(4, o0.data(), 1, DT_I32); E2.stream(5, o1.data(), 1, DT_I32);
                if (pass == 0) E2.run(B); else E2.run<SPMulSum2>(B);
                CHECK(o0 == r0, "mulsum2 out0 differs (T=%d pass %d)", TT, pass);
                CHECK(o1 == r1, "mulsum2 out1 differs (T=%d pass %d)", TT, pass);
            }
        }
        // ---- the same product sums modulo the three small primes, signed lazy arithmetic (MODE_SEQ_S) ----
        {
            const int LS3[3] = {kSignedSlot, kSignedSlot + 1, kSignedSlot + 2};
            std::vector<int32_t> o_i(B * N, 0), o_s(B * N, 0);
            Emu EI(3, LS3, keyp.data(), B, MODE_SEQ_S);
            pr.install(EI.K);
            EI.stream(0, gs.data(), T, DT_I32); EI.stream(1, xs.data(), T, DT_I32); EI.stream(2, sub.data(), 1, DT_I32);
            EI.stream(3, o_i.data(), 1, DT_I32);
            EI.run(B);
            CHECK(o_i == out_e, "signed mulsum (interpreted) differs");
            Emu ES(3, LS3, keyp.data(), B, MODE_SEQ_S);
            SPMulSum1S::prog.install(ES.K);
            ES.K.loop_count = T - 1;
            ES.stream(0, gs.data(), T, DT_I32); ES.stream(1, xs.data(), T, DT_I32); ES.stream(2, sub.data(), 1, DT_I32);
            ES.stream(4, o_s.data(), 1, DT_I32);
            ES.run<SPMulSum1S>(B);
            CHECK(o_s == out_e, "signed mulsum (static) differs");
            // two accumulators, and the compare form (a tampered plain term must flag exactly its item)
            std::vector<int32_t> cs2(xs.size());
            for (size_t i = 0; i < cs2.size(); ++i) cs2[i] = xs[(i * 7 + 3) % xs.size()];
            std::vector<int32_t> r0(B * N, 0), r1(B * N, 0), o0(B * N, 0), o1(B * N, 0);
            Emu EA(3, L2, keyp.data(), B);
            SPMulSum2::prog.install(EA.K); EA.K.loop_count = T - 1;
            EA.stream(0, gs.data(), T, DT_I32); EA.stream(1, xs.data(), T, DT_I32); EA.stream(2, cs2.data(), T, DT_I32);
            EA.stream(3, sub.data(), 1, DT_I32); EA.stream(4, r0.data(), 1, DT_I32); EA.stream(5, r1.data(), 1, DT_I32);
            EA.run<SPMulSum2>(B);
            Emu E2(3, LS3, keyp.data(), B, MODE_SEQ_S);
            SPMulSum2S::prog.install(E2.K); E2.K.loop_count = T - 1;
            E2.stream(0, gs.data(), T, DT_I32); E2.stream(1, xs.data(), T, DT_I32); E2.stream(2, cs2.data(), T, DT_I32);
            E2.stream(3, sub.data(), 1, DT_I32); E2.stream(4, o0.data(), 1, DT_I32); E2.stream(5, o1.data(), 1, DT_I32);
            E2.run<SPMulSum2S>(B);
            CHECK(o0 == r0 && o1 == r1, "signed mulsum2 differs");
            std::vector<int32_t> zero(B * N, 0);
            Emu EC(3, LS3, keyp.data(), B, MODE_SEQ_S);
            SPMulSumCmpS::prog.install(EC.K); EC.K.loop_count = T - 1;
            EC.stream(0, gs.data(), T, DT_I32); EC.stream(1, xs.data(), T, DT_I32); EC.stream(2, r0.data(), 1, DT_I32);
            EC.stream(3, zero.data(), 1, DT_I32);
            EC.run<SPMulSumCmpS>(B);
            for (int b = 0; b < B; ++b) CHECK(EC.flags[b] == 0, "signed mulsum compare: honest item %d flagged", b);
            r0[(size_t)(B - 1) * N + 77] ^= 1;
            EC.flags.assign(B, 0); EC.K.flags = EC.flags.data();
            EC.run<SPMulSumCmpS>(B);
            for (int b = 0; b < B; ++b) CHECK(EC.flags[b] == (b == B - 1 ? FLAG_FAIL : 0u), "signed mulsum compare: item %d flag %u", b, EC.flags[b]);
            // 64 terms at the extremes of int32: the edge of the three small primes' range (64 * 512 * 2^62 < P/2), and the
            // accumulator reduction every 32nd term
            const int TB = 64;
            std::vector<int32_t> gb((size_t)2 * TB * N), xb(gb.size()), ob(2 * N, 0), oref(2 * N, 0);
            for (size_t i = 0; i < gb.size(); ++i) {
                gb[i] = i < (size_t)TB * N ? INT32_MIN : rnd_q();
                xb[i] = i < (size_t)TB * N ? ((i % N == 0) ? INT32_MIN : INT32_MAX) : rnd_q();      // X^N = -1: every wrapped term flips sign again
            }
            for (int which = 0; which < 2; ++which) {
                Emu EB(3, which ? LS3 : L2, keyp.data(), 2, which ? MODE_SEQ_S : MODE_SEQ);
                SPMulSum0::prog.install(EB.K); EB.K.loop_count = TB - 1;
                EB.stream(0, gb.data(), TB, DT_I32); EB.stream(1, xb.data(), TB, DT_I32); EB.stream(4, which ? ob.data() : oref.data(), 1, DT_I32);
                if (which) EB.run<SPMulSum0S>(2); else EB.run<SPMulSum0>(2);
            }
            CHECK(ob == oref, "signed mulsum, 64 terms at the int32 extremes, differs from the 30-bit primes");
            auto gb64 = widen(gb), xb64 = widen(xb);
            for (auto *v : {&gb64, &xb64}) for (auto &c : *v) c = rzko_center(c, Q);
            std::vector<int64_t> accb(N), tmpb(N);
            for (int i = 0; i < TB; ++i) {
                rzko_poly_mul(&P, xb64.data() + (size_t)i * N, gb64.data() + (size_t)i * N, i ? tmpb.data() : accb.data());
                if (i) rzko_poly_add(&P, accb.data(), tmpb.data(), accb.data());
            }
            for (size_t i = 0; i < N; ++i) CHECK(accb[i] == ob[i], "signed mulsum 64 terms vs oracle coef %zu", i);
            printf("signed product sums ok\n");
        }
        printf("mulsum T=%d ok, ops=%d\n", T, pr.n);
    }

    // ---------------- split-key commit modulo the small prime, signed lazy arithmetic (|r| <= 1) ----------------
    {
        const int LS[1] = {kSignedSlot};
        for (int pass = 0; pass < 3; ++pass) {
            std::vector<int8_t> r1(r.size());
            for (size_t i = 0; i < r1.size(); ++i) r1[i] = (int8_t)((int)(rnd() % 3) - 1);
            if (pass == 1) for (size_t i = 0; i < r1.size(); ++i) r1[i] = (i % (3 * N) < N) ? (int8_t)((i & 1) ? 127 : -128) : (int8_t)1;   // worst case of the bound: r0 at the int8 limits, r1 = r2 = 1
            if (pass == 2) for (size_t i = 0; i < r1.size(); ++i) r1[i] = (i % (3 * N) < N) ? (int8_t)-128 : (int8_t)((i & 1) ? 1 : -1);
            std::vector<int32_t> c_s(B * 2 * N, 0);
            auto r1_64 = widen8(r1);
            std::vector<int64_t> c2(B * 2 * N);
            std::vector<uint8_t> ok2(B);
            rzko_commit_batch(&P, a1.data(), a2.data(), B, x64.data(), r1_64.data(), c2.data(), ok2.data(), 1);
            {
                Emu E(1, LS, keyp.data(), B, MODE_SPLITKEY_S);
                Prog pr;
                prog_commit_splitkey(pr, 0, 1, 2);
                pr.end(); pr.install(E.K);
                E.K.small_lim = 1;
                E.stream(0, x.data(), 1, DT_I32); E.stream(1, r1.data(), 3, DT_I8); E.stream(2, c_s.data(), 2, DT_I32);
                E.run(B);
                CHECK(same(c_s, c2), "signed split-key commit mismatch (pass %d)", pass);
                for (int b = 0; b < B; ++b) CHECK(E.flags[b] == 0, "signed split-key flags[%d]=%u", b, E.flags[b]);
            }
            {
                std::fill(c_s.begin(), c_s.end(), 0);
                Emu E(1, LS, keyp.data(), B, MODE_SPLITKEY_S);
                SPCommitSplitKeyS::prog.install(E.K);
                E.K.small_lim = 1;
                E.stream(0, x.data(), 1, DT_I32); E.stream(1, r1.data(), 3, DT_I8); E.stream(2, c_s.data(), 2, DT_I32);
                E.run<SPCommitSplitKeyS>(B);
                CHECK(same(c_s, c2), "static signed split-key commit mismatch (pass %d)", pass);
                for (int b = 0; b < B; ++b) CHECK(E.flags[b] == 0, "static signed split-key flags");
                if (pass == 0) {
                    // |r| = 2 on a transformed row must raise the range flag; on row 0 it must not
                    r1[2 * N + 3] = 2;
                    if (B > 1) r1[(size_t)3 * N + 5] = 100;
                    E.flags.assign(B, 0); E.K.flags = E.flags.data();
                    E.run<SPCommitSplitKeyS>(B);
                    CHECK(E.flags[0] == FLAG_RANGE, "signed split-key range flag item 0: %u", E.flags[0]);
                    if (B > 1) CHECK(E.flags[1] == 0, "signed split-key range flag item 1: %u", E.flags[1]);
                }
            }
        }
        printf("signed split-key commit ok\n");
    }

#if RZK_INV_DIT
    // ---------------- decimation-in-time inverse of the signed slots at the extremes of its lazy range ----------------
    // Every input is the representative c + 2p or c - 2p of a residue c (|v| up to 5p/2, the bound inv_g2_dit / inv_g1_dit
    // admit), all of one sign or alternating -- the inputs that drive the running sums furthest -- and the outputs must still
    // be the exact inverse transform (times N: the scaling rides elsewhere) modulo p.
    for (int slot = kSignedSlot; slot < kNumPrimeSlots; ++slot) {
        const PrimeTables &T = prime_tables(slot);
        const uint32_t p = T.p, mp = 0u - p;
        for (int pattern = 0; pattern < 4; ++pattern) {
            std::vector<uint32_t> res(N), ref(N);
            uint64_t st = 0x9e3779b97f4a7c15ull * (uint64_t)(slot * 4 + pattern + 1);
            uint32_t a[kLanes][kElems];
            for (size_t i = 0; i < N; ++i) {
                st = st * 6364136223846793005ull + 1442695040888963407ull;
                const uint32_t r = (uint32_t)((st >> 33) % p);
                ref[i] = r;
                int64_t c = (int64_t)r; if (c > (int64_t)(p - 1) / 2) c -= p;
                const int sign = pattern == 0 ? 1 : pattern == 1 ? -1 : pattern == 2 ? ((i & 1) ? 1 : -1) : (((i >> 3) & 1) ? 1 : -1);
                a[i >> 5][i & 31] = (uint32_t)(int32_t)(c + sign * 2 * (int64_t)p);           // contiguous layout: lane i / 32, element i % 32
            }
            ntt_inverse_ref(T, ref.data());                                                     // includes N^-1
            uint32_t buf[kN];
            for (int t = 0; t < kLanes; ++t) {
                inv_g2_dit(a[t], &T.g1[1][0][0], mp);
                for (int e = 0; e < kElems; ++e) buf[32 * t + e] = a[t][e];
            }
            for (int t = 0; t < kLanes; ++t) {
                uint32_t b[kElems];
                for (int m = 0; m < kElems; ++m) b[m] = buf[t + kLanes * m];                    // strided layout
                inv_g1_dit<false>(b, &T.g1[1][0][0], &T.g2[1][t][0], &T.twist[0][0], t, mp);
                for (int m = 0; m < kElems; ++m) res[t + kLanes * m] = b[m];
            }
            for (size_t i = 0; i < N; ++i) {
                const int64_t v = (int64_t)(int32_t)res[i];
                CHECK(v > -(int64_t)p * 5 / 4 - 2 && v < (int64_t)p * 5 / 4 + 2, "DIT inverse output %zu outside 5p/4 (slot %d pattern %d)", i, slot, pattern);
                const uint32_t got = (uint32_t)(((v % (int64_t)p) + p) % p);
                const uint32_t want = (uint32_t)((uint64_t)ref[i] * kN % p);
                CHECK(got == want, "DIT inverse coefficient %zu (slot %d pattern %d)", i, slot, pattern);
            }
        }
    }
    printf("signed DIT inverse at the range extremes ok\n");
#endif

    // ---------------- compile-time programs (vm_run_static) against the same oracle outputs ----------------
    {
        const int L1[1] = {0};
        // split-key commit without the (vacuous) norm pass
        {
            std::vector<int32_t> c_s(B * 2 * N, 0);
            Emu E(1, L1, keyp.data(), B, MODE_SPLITKEY);
            SPCommitSplitKey::prog.install(E.K);
            E.K.small_lim = 15;
            E.stream(0, x.data(), 1, DT_I32); E.stream(1, r.data(), 3, DT_I8); E.stream(2, c_s.data(), 2, DT_I32);
            E.run<SPCommitSplitKey>(B);
            CHECK(same(c_s, c_o), "static split-key commit mismatch");
            for (int b = 0; b < B; ++b) CHECK(E.flags[b] == 0, "static split-key flags");
        }
        // t = A1.y and (t, w)
        {
            std::vector<int32_t> t_s(B * N, 0), t_s2(B * N, 0), w_s(B * N, 0);
            Emu E(2, L2, keyp.data(), B);
            SPKeyMatVecT::prog.install(E.K);
            E.stream(0, y.data(), 3, DT_I32); E.stream(1, t_s.data(), 1, DT_I32);
            E.run<SPKeyMatVecT>(B);
            CHECK(same(t_s, t_o), "static keymatvec t mismatch");
            Emu E2(2, L2, keyp.data(), B);
            SPKeyMatVecTW::prog.install(E2.K);
            E2.stream(0, y.data(), 3, DT_I32); E2.stream(1, t_s2.data(), 1, DT_I32); E2.stream(2, w_s.data(), 1, DT_I32);
            E2.run<SPKeyMatVecTW>(B);
            CHECK(same(t_s2, t_o), "static keymatvec (t,w) t mismatch");
            std::vector<int64_t> w_o(N);
            for (int b = 0; b < B; ++b) {
                rzko_mat_dot(&P, 1, 3, 1, a2.data(), y64.data() + (size_t)b * 3 * N, w_o.data());
                for (size_t i = 0; i < N; ++i) CHECK(w_o[i] == w_s[b * N + i], "static w mismatch");
            }
        }
        // respond
        {
            std::vector<int32_t> z_s(B * 3 * N, 0);
            Emu E(1, L1, keyp.data(), B);
            SPRespond::prog.install(E.K);
            E.stream(0, y.data(), 3, DT_I32); E.stream(1, r.data(), 3, DT_I8); E.stream(2, d.data(), 1, DT_I8); E.stream(3, z_s.data(), 3, DT_I32);
            E.run<SPRespond>(B);
            CHECK(same(z_s, z_o), "static respond mismatch");
        }
        // verify (honest and tampered) with and without the w output
        {
            std::vector<int32_t> c32(c_o.begin(), c_o.end()), t32(t_o.begin(), t_o.end());
            for (int variant = 0; variant < 3; ++variant) {
                std::vector<int32_t> zz(z_e), w_s(B * N, 0);
                if (variant == 1) for (int b = 0; b < B; ++b) zz[(b * 3 + 1) * N + 77] -= 1;
                if (variant == 2) for (int b = 0; b < B; ++b) zz[(b * 3 + 2) * N] = 679537;
                Emu E(2, L2, keyp.data(), B);
                SPVerifyFirst::prog.install(E.K);
                E.stream(0, zz.data(), 3, DT_I32); E.stream(1, t32.data(), 1, DT_I32); E.stream(2, c32.data(), 2, DT_I32); E.stream(3, d.data(), 1, DT_I8);
                E.run<SPVerifyFirst>(B);
                Emu E2(2, L2, keyp.data(), B);
                SPVerifyFirstW::prog.install(E2.K);
                E2.stream(0, zz.data(), 3, DT_I32); E2.stream(1, t32.data(), 1, DT_I32); E2.stream(2, c32.data(), 2, DT_I32); E2.stream(3, d.data(), 1, DT_I8);
                E2.stream(4, w_s.data(), 1, DT_I32);
                E2.run<SPVerifyFirstW>(B);
                for (int b = 0; b < B; ++b) {
                    CHECK((E.flags[b] == 0) == (variant == 0), "static verify variant %d item %d flags %u", variant, b, E.flags[b]);
                    CHECK(E2.flags[b] == E.flags[b], "static verify W flags differ");
                }
                // the same launch with the challenge's NTT image computed once per group by a launch of its own
                // (OP_STG / OP_MACG), static and interpreted; group size 1 here, shared images are checked on the GPU
                {
                    std::vector<uint32_t> img(B * 2 * N, 0xDEADBEEFu), img_i(B * 2 * N, 0xDEADBEEFu);
                    Emu EI(2, L2, keyp.data(), B);
                    SPChallengeImage::prog.install(EI.K);
                    EI.stream(0, d.data(), 1, DT_I8); EI.stream(1, img.data(), 2, DT_I32);
                    EI.run<SPChallengeImage>(B);
                    Emu EJ(2, L2, keyp.data(), B);
                    SPChallengeImage::prog.install(EJ.K);
                    EJ.stream(0, d.data(), 1, DT_I8); EJ.stream(1, img_i.data(), 2, DT_I32);
                    EJ.run(B);
                    CHECK(img == img_i, "challenge image: static and interpreted differ");
                    for (int pass = 0; pass < 2; ++pass) {
                        std::vector<int32_t> w_g(B * N, 0);
                        Emu E3(2, L2, keyp.data(), B);
                        SPVerifyFirstWG::prog.install(E3.K);
                        E3.stream(0, zz.data(), 3, DT_I32); E3.stream(1, t32.data(), 1, DT_I32); E3.stream(2, c32.data(), 2, DT_I32);
                        E3.stream(3, d.data(), 1, DT_I8); E3.stream(4, w_g.data(), 1, DT_I32); E3.stream(5, img.data(), 2, DT_I32);
                        if (pass == 0) E3.run<SPVerifyFirstWG>(B); else E3.run(B);
                        for (int b = 0; b < B; ++b) CHECK(E3.flags[b] == E.flags[b], "verify with challenge image: flags differ (pass %d)", pass);
                        CHECK(w_g == w_s, "verify with challenge image: w differs (pass %d)", pass);
                    }
                }
                if (variant == 0) {        // w = A2.z - c2*d
                    std::vector<int64_t> az(N), cd(N), wo(N);
                    auto z64 = widen(zz);
                    for (int b = 0; b < B; ++b) {
                        rzko_mat_dot(&P, 1, 3, 1, a2.data(), z64.data() + (size_t)b * 3 * N, az.data());
                        rzko_poly_mul(&P, c_o.data() + ((size_t)b * 2 + 1) * N, d64.data() + (size_t)b * N, cd.data());
                        rzko_poly_sub(&P, az.data(), cd.data(), wo.data());
                        for (size_t i = 0; i < N; ++i) CHECK(wo[i] == w_s[b * N + i], "static verify w mismatch");
                    }
                }
            }
        }
        printf("static programs ok\n");
    }

    // ---------------- optional samplers (rzk_sample.cuh): Philox4x32-10 known answers (Random123) and draw ranges ----------------
    {
        Philox4 o = philox4x32_10(0, 0, 0, 0, 0, 0);
        CHECK(o.x == 0x6627e8d5u && o.y == 0xe169c58du && o.z == 0xbc57ac4cu && o.w == 0x9b00dbd8u, "philox KAT 0");
        o = philox4x32_10(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
        CHECK(o.x == 0x408f276du && o.y == 0x41c83b0eu && o.z == 0xa20bc7c6u && o.w == 0x6d5451fdu, "philox KAT 1");
        o = philox4x32_10(0x243f6a88u, 0x85a308d3u, 0x13198a2eu, 0x03707344u, 0xa4093822u, 0x299f31d0u);
        CHECK(o.x == 0xd16cfe09u && o.y == 0x94fdccebu && o.z == 0x5001e420u && o.w == 0x24126ea1u, "philox KAT 2");
        int hist[3] = {0, 0, 0};
        for (uint32_t i = 0; i < 3000; ++i) { int v = sample_small_coeff(7, i, 1, 1, 123u, 456u); CHECK(v >= -1 && v <= 1, "small range"); hist[v + 1]++; }
        for (int k = 0; k < 3; ++k) CHECK(hist[k] > 900 && hist[k] < 1100, "small histogram %d: %d", k, hist[k]);
        std::vector<int8_t> dd(N, 0);
        sample_challenge_item(3, (uint32_t)N, 36, 2, 123u, 456u, dd.data());
        int nz = 0; for (auto v : dd) { CHECK(v >= -1 && v <= 1, "challenge entry"); nz += v != 0; }
        CHECK(nz == 36, "challenge weight %d", nz);
        printf("samplers ok\n");
    }

    // ---------------- sparse response z = y + d*r as signed rotations (rzk_sparse.cuh) ----------------
    {
        const int BB = B < 6 ? 6 : B;
        std::vector<int32_t> ys(BB * 3 * N), zs(BB * 3 * N, 12345);
        std::vector<int8_t> rsm(BB * 3 * N), dsm(BB * N, 0);
        for (auto &v : ys) v = rnd_gauss(15444.0);
        for (auto &v : rsm) v = (int8_t)((int)(rnd() % 3) - 1);
        auto put_d = [&](int b, int cnt) {
            for (int k = 0; k < cnt;) { int pos = (int)(rnd() % N); if (dsm[b * N + pos]) continue; dsm[b * N + pos] = (rnd() & 1) ? 1 : -1; ++k; }
        };
        put_d(0, 36); put_d(1, 36); put_d(2, 42); put_d(3, 127); put_d(4, 0); put_d(5, 36);
        for (int b = 6; b < BB; ++b) put_d(b, 36);
        dsm[0 * N + 0] = 1; dsm[0 * N + 511] = -1;                             // rotations by 0 and by N-1
        for (size_t i = 0; i < 3 * N; ++i) rsm[1 * 3 * N + i] = (int8_t)((int)(rnd() % 7) - 3);      // |r| <= 3 with 36 terms
        for (size_t i = 0; i < 3 * N; ++i) rsm[2 * 3 * N + i] = (i & 1) ? 3 : -3;                    // extreme bytes: 42 * 6 = 252
        for (size_t i = 0; i < 3 * N; ++i) rsm[3 * 3 * N + i] = (i % 3) ? 1 : -1;                    // 127 terms, bias 1: 254
        // any int32 representative of y, including the extremes and values next to +-(q-1)/2
        for (size_t i = 0; i < N; ++i) {
            ys[5 * 3 * N + i] = (i % 4 == 0) ? INT32_MAX : (i % 4 == 1) ? INT32_MIN : (i % 4 == 2) ? (int32_t)((Q - 1) / 2) : -(int32_t)((Q - 1) / 2);
            ys[5 * 3 * N + N + i] = (int32_t)((Q - 1) / 2) - (int32_t)(i % 40);
        }
        std::vector<uint32_t> need(BB, 7), anyn(1, 0), smem(kSpWarpWords, 0xDEADBEEFu);
        SparseLaunch K;
        memset(&K, 0, sizeof(K));
        K.y = ys.data(); K.r = rsm.data(); K.d = dsm.data(); K.z = zs.data(); K.need = need.data(); K.any_need = anyn.data();
        K.n_items = BB; K.d_div = 1; K.q = (uint32_t)Q;
        static LaneCtxS sctx[32];
        auto run_all = [&]() {
            for (int b = 0; b < BB; ++b) {
                for (int li = 0; li < 32; ++li) { sctx[li].sm = smem.data(); sctx[li].item = b; sctx[li].lane = li; sctx[li].active = true; }
                sparse_respond_item(K, sctx);
            }
        };
        run_all();
        auto ys64 = widen(ys), rs64 = widen8(rsm), ds64 = widen8(dsm);
        for (auto &v : ys64) v = rzko_center(v, Q);
        std::vector<int64_t> zo(BB * 3 * N);
        rzko_open_respond_batch(&P, BB, ys64.data(), rs64.data(), ds64.data(), zo.data(), 1);
        CHECK(same(zs, zo), "sparse respond mismatch");
        for (int b = 0; b < BB; ++b) CHECK(need[b] == 0, "sparse respond need[%d]=%u", b, need[b]);
        CHECK(anyn[0] == 0, "any_need set");
        // items outside the byte range are left to the NTT program: |r| = 4; 43 terms with |r| = 2; d entry 2; 128 terms
        rsm[0 * 3 * N + 700] = 4;
        dsm[2 * N + 0] = dsm[2 * N + 0] ? dsm[2 * N + 0] : 1; { int extra = 0; for (size_t i = 0; i < N && !extra; ++i) if (!dsm[2 * N + i]) { dsm[2 * N + i] = 1; extra = 1; } }
        for (size_t i = 0; i < 3 * N; ++i) rsm[2 * 3 * N + i] = 2;
        dsm[1 * N + 9] = 2;
        { int extra = 0; for (size_t i = 0; i < N && !extra; ++i) if (!dsm[3 * N + i]) { dsm[3 * N + i] = -1; extra = 1; } }
        std::fill(zs.begin(), zs.end(), 777);
        run_all();
        for (int b = 0; b < 4; ++b) {
            CHECK(need[b] == 1, "sparse respond fallback flag item %d: %u", b, need[b]);
            for (size_t i = 0; i < 3 * N; ++i) CHECK(zs[(size_t)b * 3 * N + i] == 777, "fallback item %d was written", b);
        }
        CHECK(need[4] == 0 && need[5] == 0 && anyn[0] == 1, "need flags of in-range items");
        printf("sparse respond ok\n");
    }

    // ---------------- Commitment::verify (commit.rs:173-210), None and Some(f) branches ----------------
    {
        std::vector<int32_t> c32(c_o.begin(), c_o.end());
        // randomised opening: r' = f * r for a challenge-space f (36 entries +-1), so that f*c == A r' + f*[0;x]
        std::vector<int8_t> rf(B * 3 * N);
        std::vector<int64_t> tmp(N);
        for (int b = 0; b < B; ++b)
            for (int j = 0; j < 3; ++j) {
                rzko_poly_mul(&P, r64.data() + ((size_t)b * 3 + j) * N, d64.data() + (size_t)b * N, tmp.data());
                for (size_t i = 0; i < N; ++i) rf[((size_t)b * 3 + j) * N + i] = (int8_t)tmp[i];
            }
        for (int variant = 0; variant < 6; ++variant) {
            const bool withf = variant >= 3;
            std::vector<int32_t> cc(c32), xx(x);
            std::vector<int8_t> rr(withf ? rf : r);
            if (variant == 1 || variant == 4) for (int b = 0; b < B; ++b) cc[((size_t)b * 2 + 1) * N + 5] += 1;
            if (variant == 2 || variant == 5) for (int b = 0; b < B; ++b) rr[((size_t)b * 3) * N + 9] += 1;
            Emu E(2, L2, keyp.data(), B);
            Prog pr;
            prog_commitment_verify(pr, 0, 1, 2, withf ? 3 : -1);
            pr.end(); pr.install(E.K);
            E.stream(0, cc.data(), 2, DT_I32); E.stream(1, xx.data(), 1, DT_I32); E.stream(2, rr.data(), 3, DT_I8);
            if (withf) E.stream(3, d.data(), 1, DT_I8);
            E.run(B);
            auto c64 = widen(cc), rr64 = widen8(rr);
            for (int b = 0; b < B; ++b) {
                const int okv = rzko_commitment_verify(&P, a1.data(), a2.data(), c64.data() + (size_t)b * 2 * N, x64.data() + (size_t)b * N,
                                                       rr64.data() + (size_t)b * 3 * N, withf ? d64.data() + (size_t)b * N : nullptr);
                CHECK((E.flags[b] == 0) == (okv == 1), "commitment verify variant %d item %d: emu %u oracle %d", variant, b, E.flags[b], okv);
                CHECK((okv == 1) == (variant == 0 || variant == 3), "oracle commitment verdict variant %d", variant);
            }
        }
        printf("commitment verify ok\n");
    }

    printf(nfail ? "EMU_CHECK FAILED (%d)\n" : "EMU_CHECK PASSED\n", nfail);
    return nfail ? 1 : 0;
}
