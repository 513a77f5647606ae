// ring_zk.hpp -- C++ host-side mirror of the reference's public Rust API (ring_zk::*,
// /root/reference/src/lib.rs:5-24) on top of the C ABI in include/ringzk_b200.h.
//
// The reference is compiled code (Rust); its toolchain is absent from this image, so the host side
// above the C ABI is written in C++ with the same type and method names, argument meaning and error
// behaviour (assert! -> std::logic_error, verification -> bool), plus the `*_batch` entry points added
// alongside.  Every ring operation runs on the GPU through the C ABI; nothing here computes products.
// The randomness r, y, d is drawn here on the host with the reference's distributions
// (src/polynomial.rs:14-44, src/challenge_space.rs:12-33) from a seeded std::mt19937_64.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <memory>
#include <random>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/ringzk_b200.h"

namespace ring_zk {

using Rng = std::mt19937_64;
constexpr int64_t Q_DEFAULT = 3515337053LL;

inline void rzk_assert(bool c, const char *what) { if (!c) throw std::logic_error(std::string("assertion failed: ") + what); }

// Polynomial<ZqI64<Q>, N>: N canonical centred coefficients
template <int N>
struct Polynomial {
    std::vector<int32_t> c = std::vector<int32_t>(N, 0);
    bool operator==(const Polynomial &o) const { return c == o.c; }
    int deg() const { for (int i = N - 1; i >= 0; --i) if (c[i]) return i; return 0; }
    static Polynomial from_coeffs(const std::vector<int64_t> &v)             // src/params.rs:75,90
    {
        rzk_assert((int)v.size() <= N, "more than N coefficients");
        Polynomial p;
        const int64_t half = (Q_DEFAULT - 1) / 2;
        for (size_t i = 0; i < v.size(); ++i) {
            int64_t r = v[i] % Q_DEFAULT;
            if (r > half) r -= Q_DEFAULT; else if (r < -half) r += Q_DEFAULT;
            p.c[i] = (int32_t)r;
        }
        return p;
    }
};

struct Engine {
    rzk_engine *h = nullptr;
    explicit Engine(int N, int device = -1)
    {
        rzk_params P = rzk_default_params(N);
        if (rzk_create(&P, device, &h) != RZK_OK) throw std::runtime_error(std::string("rzk_create: ") + rzk_last_error(nullptr));
    }
    ~Engine() { rzk_destroy(h); }
    Engine(const Engine &) = delete;
    void check(int rc) const { if (rc != RZK_OK) throw std::runtime_error(std::string("ringzk_b200: ") + rzk_last_error(h)); }
};

inline bool bit(const std::vector<uint8_t> &bm, size_t i) { return (bm[i >> 3] >> (i & 7)) & 1; }

template <int N> struct CommitmentKey;

// Params<ZqI64<Q>>  (src/params.rs:18-36, default 121-138)
struct Params {
    int64_t q = Q_DEFAULT / 2, b = 1;
    int n = 1, k = 3, l = 1, kappa = 36;
    static Params default_() { return Params(); }
    uint64_t standard_deviation(int deg_n) const                               // src/params.rs:94-98
    {
        return (uint64_t)b * (uint64_t)(11 * kappa) * (uint64_t)std::floor(std::sqrt((double)(k * deg_n)));
    }
    template <int N> CommitmentKey<N> generate_commitment_key(Rng &rng) const; // src/params.rs:49-54
    template <int N> std::vector<Polynomial<N>> prepare_value(const std::vector<std::vector<int64_t>> &value) const
    {
        rzk_assert((int)value.size() == l, "value.len() == self.l");           // src/params.rs:71
        std::vector<Polynomial<N>> out;
        for (auto &v : value) out.push_back(Polynomial<N>::from_coeffs(v));
        return out;
    }
    template <int N> Polynomial<N> prepare_scalar(const std::vector<int64_t> &s) const { return Polynomial<N>::from_coeffs(s); }

    // samplers, flat [count][N]
    void sample_small(Rng &rng, size_t polys, int N, std::vector<int8_t> &out) const        // src/polynomial.rs:14-25
    {
        std::uniform_int_distribution<int> d((int)-b, (int)b);
        out.resize(polys * N);
        for (auto &v : out) v = (int8_t)d(rng);
    }
    void sample_gaussian(Rng &rng, size_t polys, int N, std::vector<int32_t> &out) const     // src/polynomial.rs:28-44
    {
        std::normal_distribution<double> d(0.0, (double)standard_deviation(N));
        out.resize(polys * N);
        for (auto &v : out) v = (int32_t)std::trunc(d(rng));
    }
    void sample_challenge(Rng &rng, size_t B, int N, std::vector<int8_t> &out) const         // src/challenge_space.rs:12-33
    {
        out.assign(B * N, 0);
        const int nnz = std::min(kappa, N);
        for (size_t i = 0; i < B; ++i) {
            int8_t *d = out.data() + i * N;
            for (int j = 0; j < nnz; ++j) d[j] = (rng() & 1) ? 1 : -1;
            std::shuffle(d, d + N, rng);
        }
    }
};

template <int N> struct Opening {                                                         // src/commit.rs:223-235
    std::vector<Polynomial<N>> x;
    std::vector<int8_t> r;
    std::vector<int8_t> f;       // empty = None; else a challenge-space polynomial [N] (randomised opening)
};

template <int N>
struct Commitment {                                                                       // src/commit.rs:135-141
    std::vector<int32_t> c;      // [(n+l)][N]
    bool verify(const Opening<N> &opening, const CommitmentKey<N> &ck, const Params &params) const;   // src/commit.rs:173-210
};

template <int N>
struct CommitmentKey {                                                                    // src/commit.rs:19-60
    std::vector<int64_t> a1, a2;     // [n][k][N], [l][k][N]
    std::shared_ptr<Engine> eng;

    static CommitmentKey new_(Rng &rng, const Params &P, int device = -1)
    {
        CommitmentKey ck;
        ck.a1.assign((size_t)P.n * P.k * N, 0);
        ck.a2.assign((size_t)P.l * P.k * N, 0);
        std::uniform_int_distribution<int64_t> d(-P.q, P.q);
        for (int i = 0; i < P.n; ++i) {
            ck.a1[((size_t)i * P.k + i) * N] = 1;                                          // src/commit.rs:39
            for (int j = P.n; j < P.k; ++j) for (int c = 0; c < N; ++c) ck.a1[((size_t)i * P.k + j) * N + c] = d(rng);
        }
        for (int i = 0; i < P.l; ++i) {
            ck.a2[((size_t)i * P.k + P.n + i) * N] = 1;                                    // src/commit.rs:50-51
            for (int j = P.n + P.l; j < P.k; ++j) for (int c = 0; c < N; ++c) ck.a2[((size_t)i * P.k + j) * N + c] = d(rng);
        }
        ck.eng = std::make_shared<Engine>(N, device);
        ck.eng->check(rzk_set_key(ck.eng->h, ck.a1.data(), ck.a2.data()));
        return ck;
    }

    // CommitmentKey::commit for a batch: x [B][l][N] -> r [B][k][N], c [B][n+l][N]   (src/commit.rs:88-128)
    void commit_batch(Rng &rng, size_t B, const std::vector<int32_t> &x, const Params &P,
                      std::vector<int8_t> &r, std::vector<int32_t> &c) const
    {
        rzk_assert(x.size() == B * P.l * N, "l == x.len()");                               // src/commit.rs:95
        P.sample_small(rng, B * P.k, N, r);
        c.resize(B * (P.n + P.l) * N);
        std::vector<uint8_t> ok((B + 7) / 8);
        eng->check(rzk_commit_batch(eng->h, B, x.data(), r.data(), c.data(), ok.data()));
        for (size_t i = 0; i < B; ++i)                                                     // src/commit.rs:98-107 (redraw loop)
            while (!bit(ok, i)) {
                std::vector<int8_t> ri; std::vector<uint8_t> oki(1);
                P.sample_small(rng, P.k, N, ri);
                std::copy(ri.begin(), ri.end(), r.begin() + i * P.k * N);
                eng->check(rzk_commit_batch(eng->h, 1, x.data() + i * P.l * N, ri.data(), c.data() + i * (P.n + P.l) * N, oki.data()));
                if (oki[0] & 1) ok[i >> 3] |= (uint8_t)(1u << (i & 7));
            }
    }

    std::pair<Opening<N>, Commitment<N>> commit(Rng &rng, const std::vector<Polynomial<N>> &x, const Params &P) const
    {
        rzk_assert((int)x.size() == P.l, "l == x.len()");
        std::vector<int32_t> xf; for (auto &p : x) xf.insert(xf.end(), p.c.begin(), p.c.end());
        Opening<N> o; Commitment<N> com;
        o.x = x;
        commit_batch(rng, 1, xf, P, o.r, com.c);
        return {o, com};
    }
};

template <int N> CommitmentKey<N> Params::generate_commitment_key(Rng &rng) const { return CommitmentKey<N>::new_(rng, *this); }

template <int N>
bool Commitment<N>::verify(const Opening<N> &o, const CommitmentKey<N> &ck, const Params &P) const
{
    std::vector<int32_t> xf; for (auto &p : o.x) xf.insert(xf.end(), p.c.begin(), p.c.end());
    rzk_assert(o.f.empty() || (int)o.f.size() == N, "f is one polynomial");
    (void)P;
    std::vector<uint8_t> ok(1);
    // check_commit_constraint(r), then A.r + [0;x] == c (None) or f*c == A.r + f*[0;x] (Some(f))
    ck.eng->check(rzk_commitment_verify_batch(ck.eng->h, 1, c.data(), xf.data(), o.r.data(), o.f.empty() ? nullptr : o.f.data(), ok.data()));
    return ok[0] & 1;
}

// ------------------------------------------------------------------ Open proof (src/prove/open.rs)
template <int N> struct OpenProofResponseContext { Opening<N> opening; std::vector<int32_t> y; };
template <int N> struct OpenProofCommitment { Commitment<N> c; std::vector<int32_t> t; };
template <int N> struct OpenProofVerificationContext { std::vector<int32_t> c1, t; std::vector<int8_t> d; };
template <int N> struct OpenProofChallenge { std::vector<int8_t> d; };
template <int N> struct OpenProofResponse { std::vector<int32_t> z; };

template <int N>
struct OpenProofProver {
    Params params; CommitmentKey<N> ck;
    OpenProofProver(const CommitmentKey<N> &ck_, const Params &p) : params(p), ck(ck_) {}           // open.rs:69

    // open.rs:80-103 for x [B][l][N]
    void commit_batch(Rng &rng, size_t B, const std::vector<int32_t> &x, std::vector<int8_t> &r, std::vector<int32_t> &y,
                      std::vector<int32_t> &c, std::vector<int32_t> &t) const
    {
        rzk_assert(x.size() == B * params.l * N, "l == x.len()");
        params.sample_small(rng, B * params.k, N, r);
        params.sample_gaussian(rng, B * params.k, N, y);
        c.resize(B * 2 * N); t.resize(B * N);
        std::vector<uint8_t> ok((B + 7) / 8);
        ck.eng->check(rzk_open_commit_batch(ck.eng->h, B, x.data(), r.data(), y.data(), c.data(), t.data(), ok.data()));
        for (size_t i = 0; i < B; ++i) rzk_assert(bit(ok, i), "commit constraint (redraw r)");
    }
    // open.rs:107-117
    void create_response_batch(size_t B, const std::vector<int32_t> &y, const std::vector<int8_t> &r, const std::vector<int8_t> &d,
                               std::vector<int32_t> &z) const
    {
        z.resize(B * params.k * N);
        ck.eng->check(rzk_open_respond_batch(ck.eng->h, B, y.data(), r.data(), d.data(), z.data()));
    }
    std::pair<OpenProofResponseContext<N>, OpenProofCommitment<N>> commit(Rng &rng, const std::vector<Polynomial<N>> &x) const
    {
        rzk_assert((int)x.size() == params.l, "l == x.len()");
        std::vector<int32_t> xf; for (auto &p : x) xf.insert(xf.end(), p.c.begin(), p.c.end());
        OpenProofResponseContext<N> ctx; OpenProofCommitment<N> com;
        ctx.opening.x = x;
        commit_batch(rng, 1, xf, ctx.opening.r, ctx.y, com.c.c, com.t);
        return {ctx, com};
    }
    OpenProofResponse<N> create_response(const OpenProofResponseContext<N> &ctx, const OpenProofChallenge<N> &ch) const
    {
        OpenProofResponse<N> r;
        create_response_batch(1, ctx.y, ctx.opening.r, ch.d, r.z);
        return r;
    }
};

template <int N>
struct OpenProofVerifier {
    Params params; CommitmentKey<N> ck;
    OpenProofVerifier(const CommitmentKey<N> &ck_, const Params &p) : params(p), ck(ck_) {}         // open.rs:135

    // open.rs:162-174 -> bitmap
    std::vector<uint8_t> verify_batch(size_t B, const std::vector<int32_t> &z, const std::vector<int32_t> &t,
                                      const std::vector<int32_t> &c1, const std::vector<int8_t> &d) const
    {
        std::vector<uint8_t> bm((B + 7) / 8);
        ck.eng->check(rzk_open_verify_batch(ck.eng->h, B, z.data(), t.data(), c1.data(), d.data(), bm.data()));
        return bm;
    }
    std::pair<OpenProofVerificationContext<N>, OpenProofChallenge<N>> generate_challenge(Rng &rng, const OpenProofCommitment<N> &com) const
    {
        OpenProofVerificationContext<N> v; OpenProofChallenge<N> ch;
        params.sample_challenge(rng, 1, N, ch.d);                                                   // open.rs:148
        v.c1.assign(com.c.c.begin(), com.c.c.begin() + (size_t)params.l * N);                       // commit.rs:213-218
        v.t = com.t; v.d = ch.d;
        return {v, ch};
    }
    bool verify(const OpenProofResponse<N> &resp, const OpenProofVerificationContext<N> &ctx) const
    {
        return bit(verify_batch(1, resp.z, ctx.t, ctx.c1, ctx.d), 0);
    }
};

// ------------------------------------------------------------------ Linear proof (src/prove/linear.rs)
template <int N> struct LinearProofResponseContext { Opening<N> opening, opening_p; std::vector<int32_t> y, yp; };
template <int N> struct LinearProofCommitment { Commitment<N> c, cp; Polynomial<N> g; std::vector<int32_t> t, tp, u; };
template <int N> struct LinearProofVerificationContext { std::vector<int32_t> c, cp, g, t, tp, u; std::vector<int8_t> d; };
template <int N> struct LinearProofChallenge { std::vector<int8_t> d; };
template <int N> struct LinearProofResponse { std::vector<int32_t> z, zp; };

template <int N>
struct LinearProofProver {
    Params params; CommitmentKey<N> ck;
    LinearProofProver(const CommitmentKey<N> &ck_, const Params &p) : params(p), ck(ck_) {}         // linear.rs:71

    std::pair<LinearProofResponseContext<N>, LinearProofCommitment<N>> commit(Rng &rng, const Polynomial<N> &g,
                                                                              const std::vector<Polynomial<N>> &x) const
    {                                                                                               // linear.rs:82-140
        rzk_assert((int)x.size() == params.l, "l == x.len()");
        LinearProofResponseContext<N> ctx; LinearProofCommitment<N> com;
        std::vector<int32_t> xf; for (auto &p : x) xf.insert(xf.end(), p.c.begin(), p.c.end());
        params.sample_small(rng, params.k, N, ctx.opening_p.r);          // draw order r', r, y, y' (linear.rs:96-115)
        params.sample_small(rng, params.k, N, ctx.opening.r);
        params.sample_gaussian(rng, params.k, N, ctx.y);
        params.sample_gaussian(rng, params.k, N, ctx.yp);
        std::vector<int32_t> gx(N);
        com.c.c.resize(2 * N); com.cp.c.resize(2 * N); com.t.resize(N); com.tp.resize(N); com.u.resize(N);
        std::vector<uint8_t> ok(1);
        ck.eng->check(rzk_linear_commit_batch(ck.eng->h, 1, g.c.data(), xf.data(), ctx.opening_p.r.data(), ctx.opening.r.data(),
                                              ctx.y.data(), ctx.yp.data(), gx.data(), com.cp.c.data(), com.c.c.data(),
                                              com.t.data(), com.tp.data(), com.u.data(), ok.data()));
        rzk_assert(ok[0] & 1, "commit constraint (redraw r)");
        ctx.opening.x = x;
        Polynomial<N> gxp; gxp.c = gx; ctx.opening_p.x = {gxp};
        com.g = g;
        return {ctx, com};
    }
    LinearProofResponse<N> create_response(const LinearProofResponseContext<N> &ctx, const LinearProofChallenge<N> &ch) const
    {                                                                                               // linear.rs:144-158
        LinearProofResponse<N> r; r.z.resize(params.k * N); r.zp.resize(params.k * N);
        ck.eng->check(rzk_linear_respond_batch(ck.eng->h, 1, ctx.y.data(), ctx.yp.data(), ctx.opening.r.data(),
                                               ctx.opening_p.r.data(), ch.d.data(), r.z.data(), r.zp.data()));
        return r;
    }
};

template <int N>
struct LinearProofVerifier {
    Params params; CommitmentKey<N> ck;
    LinearProofVerifier(const CommitmentKey<N> &ck_, const Params &p) : params(p), ck(ck_) {}       // linear.rs:176
    std::pair<LinearProofVerificationContext<N>, LinearProofChallenge<N>> generate_challenge(Rng &rng, const LinearProofCommitment<N> &com) const
    {
        LinearProofVerificationContext<N> v; LinearProofChallenge<N> ch;
        params.sample_challenge(rng, 1, N, ch.d);                                                   // linear.rs:192
        v.c = com.c.c; v.cp = com.cp.c; v.g = com.g.c; v.t = com.t; v.tp = com.tp; v.u = com.u; v.d = ch.d;
        return {v, ch};
    }
    bool verify(const LinearProofResponse<N> &r, const LinearProofVerificationContext<N> &v) const  // linear.rs:213-250
    {
        std::vector<uint8_t> bm(1);
        ck.eng->check(rzk_linear_verify_batch(ck.eng->h, 1, r.z.data(), r.zp.data(), v.c.data(), v.cp.data(), v.g.data(),
                                              v.t.data(), v.tp.data(), v.u.data(), v.d.data(), bm.data()));
        return bm[0] & 1;
    }
};

// ------------------------------------------------------------------ Sum proof (src/prove/sum.rs)
template <int N> struct SumProofResponseContext { std::vector<Opening<N>> openings; Opening<N> opening_p; std::vector<int32_t> yp, ys; };
template <int N> struct SumProofCommitment { Commitment<N> cp; std::vector<Commitment<N>> cs; std::vector<int32_t> gs, tp, ts, u; };
template <int N> struct SumProofVerificationContext { std::vector<int32_t> cp, cs, gs, ts, tp, u; std::vector<int8_t> d; uint32_t T; };
template <int N> struct SumProofChallenge { std::vector<int8_t> d; };
template <int N> struct SumProofResponse { std::vector<int32_t> zp, zs; };

template <int N>
struct SumProofProver {
    Params params; CommitmentKey<N> ck;
    SumProofProver(const CommitmentKey<N> &ck_, const Params &p) : params(p), ck(ck_) {}            // sum.rs:84

    std::pair<SumProofResponseContext<N>, SumProofCommitment<N>> commit(Rng &rng, const std::vector<Polynomial<N>> &gs,
                                                                        const std::vector<std::vector<Polynomial<N>>> &xs) const
    {                                                                                               // sum.rs:99-178
        rzk_assert(!gs.empty() && gs.size() == xs.size(), "!gs.is_empty() && gs.len() == xs.len()"); // sum.rs:105
        const uint32_t T = (uint32_t)gs.size();
        SumProofResponseContext<N> ctx; SumProofCommitment<N> com;
        std::vector<int32_t> gf, xf;
        for (auto &g : gs) gf.insert(gf.end(), g.c.begin(), g.c.end());
        for (auto &x : xs) { rzk_assert((int)x.size() == params.l, "l == x.len()"); for (auto &p : x) xf.insert(xf.end(), p.c.begin(), p.c.end()); }
        std::vector<int8_t> rs;
        params.sample_small(rng, params.k, N, ctx.opening_p.r);          // r', r_0.., y_0.., y' (sum.rs:116-142)
        params.sample_small(rng, (size_t)T * params.k, N, rs);
        params.sample_gaussian(rng, (size_t)T * params.k, N, ctx.ys);
        params.sample_gaussian(rng, params.k, N, ctx.yp);
        std::vector<int32_t> xp(N), cs((size_t)T * 2 * N);
        com.cp.c.resize(2 * N); com.ts.resize((size_t)T * N); com.tp.resize(N); com.u.resize(N);
        std::vector<uint8_t> ok(1);
        ck.eng->check(rzk_sum_commit_batch(ck.eng->h, 1, T, gf.data(), xf.data(), ctx.opening_p.r.data(), rs.data(), ctx.ys.data(),
                                           ctx.yp.data(), xp.data(), com.cp.c.data(), cs.data(), com.ts.data(), com.tp.data(),
                                           com.u.data(), ok.data()));
        rzk_assert(ok[0] & 1, "commit constraint (redraw r)");
        for (uint32_t i = 0; i < T; ++i) {
            Opening<N> o; o.x = xs[i]; o.r.assign(rs.begin() + (size_t)i * params.k * N, rs.begin() + (size_t)(i + 1) * params.k * N);
            ctx.openings.push_back(o);
            Commitment<N> c; c.c.assign(cs.begin() + (size_t)i * 2 * N, cs.begin() + (size_t)(i + 1) * 2 * N);
            com.cs.push_back(c);
        }
        Polynomial<N> xpp; xpp.c = xp; ctx.opening_p.x = {xpp};
        com.gs = gf;
        return {ctx, com};
    }
    SumProofResponse<N> create_response(const SumProofResponseContext<N> &ctx, const SumProofChallenge<N> &ch) const
    {                                                                                               // sum.rs:182-200
        const uint32_t T = (uint32_t)ctx.openings.size();
        std::vector<int8_t> rs; for (auto &o : ctx.openings) rs.insert(rs.end(), o.r.begin(), o.r.end());
        SumProofResponse<N> r; r.zs.resize((size_t)T * params.k * N); r.zp.resize(params.k * N);
        ck.eng->check(rzk_sum_respond_batch(ck.eng->h, 1, T, ctx.ys.data(), ctx.yp.data(), rs.data(), ctx.opening_p.r.data(),
                                            ch.d.data(), r.zs.data(), r.zp.data()));
        return r;
    }
};

template <int N>
struct SumProofVerifier {
    Params params; CommitmentKey<N> ck;
    SumProofVerifier(const CommitmentKey<N> &ck_, const Params &p) : params(p), ck(ck_) {}          // sum.rs:219
    std::pair<SumProofVerificationContext<N>, SumProofChallenge<N>> generate_challenge(Rng &rng, const SumProofCommitment<N> &com) const
    {
        SumProofVerificationContext<N> v; SumProofChallenge<N> ch;
        params.sample_challenge(rng, 1, N, ch.d);                                                   // sum.rs:233
        v.T = (uint32_t)com.cs.size();
        for (auto &c : com.cs) v.cs.insert(v.cs.end(), c.c.begin(), c.c.end());
        v.cp = com.cp.c; v.gs = com.gs; v.ts = com.ts; v.tp = com.tp; v.u = com.u; v.d = ch.d;
        return {v, ch};
    }
    bool verify(const SumProofResponse<N> &r, const SumProofVerificationContext<N> &v) const        // sum.rs:257-320
    {
        if (r.zs.size() != (size_t)v.T * params.k * N) return false;                                // Vec inequality, sum.rs:289
        std::vector<uint8_t> bm(1);
        ck.eng->check(rzk_sum_verify_batch(ck.eng->h, 1, v.T, r.zs.data(), r.zp.data(), v.cs.data(), v.cp.data(), v.gs.data(),
                                           v.ts.data(), v.tp.data(), v.u.data(), v.d.data(), bm.data()));
        return bm[0] & 1;
    }
};

}  // namespace ring_zk
