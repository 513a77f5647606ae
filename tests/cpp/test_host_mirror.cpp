// test_host_mirror.cpp -- the reference's integration tests (/root/reference/tests/test.rs:11-93) and
// doctests (README.md:32-55, src/commit.rs:152-171), written against the C++ host mirror
// (ring-zk_b200/host/ring_zk.hpp) so that they read like the reference's own tests.  N = 512.
// Needs a CUDA device (run by tests/test_gpu_host_mirror.py).
#include <cstdio>
#include "../../ring-zk_b200/host/ring_zk.hpp"

using namespace ring_zk;
constexpr int N = 512;

static int failures = 0;
#define ASSERT(c) do { if (!(c)) { printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); ++failures; } } while (0)

static std::vector<int64_t> random_value(Rng &rng, int64_t bound)          // tests/test.rs:95-99
{
    std::uniform_int_distribution<int64_t> d(-bound, bound);
    std::uniform_int_distribution<int> len(1, N);
    std::vector<int64_t> v(len(rng));
    for (auto &c : v) c = d(rng);
    return v;
}

static void test_readme_and_doctests()
{
    Rng rng(1);
    Params params = Params::default_();
    ASSERT(params.standard_deviation(1024) == 21780);                      // params.rs:144-150
    auto ck = params.generate_commitment_key<N>(rng);
    auto x = params.prepare_value<N>({{1, 2, 3, 4}});
    ASSERT(x.size() == 1 && x[0].deg() == 3);                              // params.rs:152-168
    OpenProofProver<N> prover(ck, params);
    OpenProofVerifier<N> verifier(ck, params);
    auto [response_ctx, commitment] = prover.commit(rng, x);               // README.md:44-47
    auto [verification_ctx, challenge] = verifier.generate_challenge(rng, commitment);
    auto response = prover.create_response(response_ctx, challenge);
    ASSERT(verifier.verify(response, verification_ctx));
    // commit.rs:152-171
    auto [open1, com1] = ck.commit(rng, x, params);
    ASSERT(com1.verify(open1, ck, params));
    auto x2 = params.prepare_value<N>({{4, 5, 6, 7}});
    auto [open2, com2] = ck.commit(rng, x2, params);
    ASSERT(com2.verify(open2, ck, params));
    ASSERT(!com2.verify(open1, ck, params));
    ASSERT(!com1.verify(open2, ck, params));
    {   // randomised opening (commit.rs:203-207) with the unit f = -X^3: r' = f*r keeps the norm, f*c == A.r' + f*[0;x]
        auto openf = open1;
        openf.f.assign(N, 0); openf.f[3] = -1;
        for (int j = 0; j < params.k; ++j)
            for (int i = 0; i < N; ++i) {
                const int src = (i - 3 + N) % N;
                const int sign = (i >= 3) ? -1 : 1;            // X^N = -1
                openf.r[(size_t)j * N + i] = (int8_t)(sign * open1.r[(size_t)j * N + src]);
            }
        ASSERT(com1.verify(openf, ck, params));
        ASSERT(!com2.verify(openf, ck, params));
        openf.r[5] = (int8_t)(openf.r[5] + 1);
        ASSERT(!com1.verify(openf, ck, params));
    }
    bool threw = false;
    try { params.prepare_value<N>({{1}, {2}}); } catch (const std::logic_error &) { threw = true; }   // params.rs:71
    ASSERT(threw);
}

static void test_open_proof(int iters)                                      // tests/test.rs:11-31
{
    Rng rng(3);
    Params params = Params::default_();
    const int64_t bound = params.q;
    for (int i = 0; i < iters; ++i) {
        auto ck = params.generate_commitment_key<N>(rng);
        auto x = params.prepare_value<N>({random_value(rng, bound)});
        OpenProofProver<N> prover(ck, params);
        OpenProofVerifier<N> verifier(ck, params);
        auto [response_ctx, commitment] = prover.commit(rng, x);
        ASSERT(commitment.c.verify(response_ctx.opening, ck, params));
        auto [verification_ctx, challenge] = verifier.generate_challenge(rng, commitment);
        auto response = prover.create_response(response_ctx, challenge);
        ASSERT(verifier.verify(response, verification_ctx));
        response.z[7] += 1;
        ASSERT(!verifier.verify(response, verification_ctx));
    }
}

static void test_linear_proof(int iters)                                    // tests/test.rs:33-56
{
    Rng rng(4);
    Params params = Params::default_();
    const int64_t bound = params.q;
    for (int i = 0; i < iters; ++i) {
        auto ck = params.generate_commitment_key<N>(rng);
        auto x = params.prepare_value<N>({random_value(rng, bound)});
        auto g = params.prepare_scalar<N>(random_value(rng, bound));
        LinearProofProver<N> prover(ck, params);
        LinearProofVerifier<N> verifier(ck, params);
        auto [response_ctx, commitment] = prover.commit(rng, g, x);
        ASSERT(commitment.c.verify(response_ctx.opening, ck, params));
        ASSERT(commitment.cp.verify(response_ctx.opening_p, ck, params));
        auto [verification_ctx, challenge] = verifier.generate_challenge(rng, commitment);
        auto response = prover.create_response(response_ctx, challenge);
        ASSERT(verifier.verify(response, verification_ctx));
    }
}

static void test_sum_proof(int iters)                                       // tests/test.rs:58-93
{
    Rng rng(5);
    Params params = Params::default_();
    const int64_t bound = params.q;
    const int VL = 4;
    for (int i = 0; i < iters; ++i) {
        auto ck = params.generate_commitment_key<N>(rng);
        std::vector<std::vector<Polynomial<N>>> xs;
        std::vector<Polynomial<N>> gs;
        for (int j = 0; j < VL; ++j) xs.push_back(params.prepare_value<N>({random_value(rng, bound)}));
        for (int j = 0; j < VL; ++j) gs.push_back(params.prepare_scalar<N>(random_value(rng, bound)));
        SumProofProver<N> prover(ck, params);
        SumProofVerifier<N> verifier(ck, params);
        auto [response_ctx, commitment] = prover.commit(rng, gs, xs);
        ASSERT(commitment.cp.verify(response_ctx.opening_p, ck, params));
        for (size_t j = 0; j < commitment.cs.size(); ++j) ASSERT(commitment.cs[j].verify(response_ctx.openings[j], ck, params));
        auto [verification_ctx, challenge] = verifier.generate_challenge(rng, commitment);
        auto response = prover.create_response(response_ctx, challenge);
        ASSERT(verifier.verify(response, verification_ctx));
    }
    bool threw = false;                                                     // sum.rs:105
    try { auto ck = params.generate_commitment_key<N>(rng); SumProofProver<N>(ck, params).commit(rng, {}, {}); }
    catch (const std::logic_error &) { threw = true; }
    ASSERT(threw);
}

int main(int argc, char **argv)
{
    const int iters = argc > 1 ? atoi(argv[1]) : 3;
    try {
        test_readme_and_doctests();
        test_open_proof(iters);
        test_linear_proof(iters);
        test_sum_proof(iters);
    } catch (const std::exception &e) {
        printf("EXCEPTION: %s\n", e.what());
        return 2;
    }
    printf(failures ? "HOST_MIRROR FAILED (%d)\n" : "HOST_MIRROR PASSED\n", failures);
    return failures ? 1 : 0;
}
