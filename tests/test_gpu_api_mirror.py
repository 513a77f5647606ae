"""The reference's own integration tests (/root/reference/tests/test.rs:11-93) and doctests
(README.md:32-55, commit.rs:152-171), driven through the host-side mirror of its public API
(ring-zk_b200/api.py) so that every ring operation runs on the GPU engine.  N = 512 (the ring
degree the engine accelerates; the reference's tests use N = 16, its doctests N = 512)."""
import importlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

api = importlib.import_module("ring-zk_b200.api")
N = 512
ITERS = 4


def random_value(rng, bound):  # tests/test.rs:95-99
    p = rng.integers(-bound, bound + 1, size=N)
    return list(p[: rng.integers(1, N + 1)])


def test_readme_open_proof():  # README.md:32-55
    rng = np.random.default_rng(1)
    params = api.Params.default()
    ck = params.generate_commitment_key(rng, N)
    x = params.prepare_value([[1, 2, 3, 4]], N)
    prover = api.OpenProofProver(ck, params)
    verifier = api.OpenProofVerifier(ck, params)
    response_ctx, commitment = prover.commit(rng, x)
    verification_ctx, challenge = verifier.generate_challenge(rng, commitment)
    response = prover.create_response(response_ctx, challenge)
    assert verifier.verify(response, verification_ctx)


def test_commitment_doctest():  # commit.rs:152-171
    rng = np.random.default_rng(2)
    params = api.Params.default()
    ck = params.generate_commitment_key(rng, N)
    x = params.prepare_value([[1, 2, 3, 4]], N)
    open1, com1 = ck.commit(rng, x, params)
    assert com1.verify(open1, ck, params)
    x2 = params.prepare_value([[4, 5, 6, 7]], N)
    open2, com2 = ck.commit(rng, x2, params)
    assert com2.verify(open2, ck, params)
    assert not com2.verify(open1, ck, params)
    assert not com1.verify(open2, ck, params)


def test_prepare_value_shape_panics():  # params.rs:71, commit.rs:95
    params = api.Params.default()
    with pytest.raises(AssertionError):
        params.prepare_value([[1], [2]], N)
    assert params.standard_deviation(1024) == 21780      # params.rs:144-150


def test_open_proof():  # tests/test.rs:11-31
    rng = np.random.default_rng(3)
    params = api.Params.default()
    bound = params.q
    for _ in range(ITERS):
        ck = params.generate_commitment_key(rng, N)
        x = params.prepare_value([random_value(rng, bound)], N)
        prover, verifier = api.OpenProofProver(ck, params), api.OpenProofVerifier(ck, params)
        response_ctx, commitment = prover.commit(rng, x)
        assert commitment.c.verify(response_ctx.opening, ck, params)
        verification_ctx, challenge = verifier.generate_challenge(rng, commitment)
        response = prover.create_response(response_ctx, challenge)
        assert verifier.verify(response, verification_ctx)


def test_linear_proof():  # tests/test.rs:33-56
    rng = np.random.default_rng(4)
    params = api.Params.default()
    bound = params.q
    for _ in range(ITERS):
        ck = params.generate_commitment_key(rng, N)
        x = params.prepare_value([random_value(rng, bound)], N)
        g = params.prepare_scalar(random_value(rng, bound), N)
        prover, verifier = api.LinearProofProver(ck, params), api.LinearProofVerifier(ck, params)
        response_ctx, commitment = prover.commit(rng, g, x)
        assert commitment.c.verify(response_ctx.opening, ck, params)
        assert commitment.cp.verify(response_ctx.opening_p, ck, params)
        verification_ctx, challenge = verifier.generate_challenge(rng, commitment)
        response = prover.create_response(response_ctx, challenge)
        assert verifier.verify(response, verification_ctx)


def test_sum_proof():  # tests/test.rs:58-93 (4 terms)
    rng = np.random.default_rng(5)
    params = api.Params.default()
    bound = params.q
    VL = 4
    for _ in range(ITERS):
        ck = params.generate_commitment_key(rng, N)
        xs = [params.prepare_value([random_value(rng, bound)], N) for _ in range(VL)]
        gs = [params.prepare_scalar(random_value(rng, bound), N) for _ in range(VL)]
        prover, verifier = api.SumProofProver(ck, params), api.SumProofVerifier(ck, params)
        response_ctx, commitment = prover.commit(rng, gs, xs)
        assert commitment.cp.verify(response_ctx.opening_p, ck, params)
        for c, o in zip(commitment.cs, response_ctx.openings):
            assert c.verify(o, ck, params)
        verification_ctx, challenge = verifier.generate_challenge(rng, commitment)
        response = prover.create_response(response_ctx, challenge)
        assert verifier.verify(response, verification_ctx)
        # soundness smoke: a tampered response does not verify
        response.zs[VL - 1, 2, 11] += 1
        assert not verifier.verify(response, verification_ctx)


def test_sum_proof_empty_panics():  # sum.rs:105
    rng = np.random.default_rng(6)
    params = api.Params.default()
    ck = params.generate_commitment_key(rng, N)
    with pytest.raises(AssertionError):
        api.SumProofProver(ck, params).commit(rng, [], [])


def test_batched_entry_points():
    """the `*_batch` methods added alongside the reference API"""
    rng = np.random.default_rng(7)
    params = api.Params.default()
    ck = params.generate_commitment_key(rng, N)
    B = 300
    X = rng.integers(-params.q, params.q + 1, size=(B, 1, N)).astype(np.int32)
    prover, verifier = api.OpenProofProver(ck, params), api.OpenProofVerifier(ck, params)
    s = prover.commit_batch(rng, X)
    d = verifier.generate_challenge_batch(rng, B)
    assert (np.abs(d).sum(axis=1) == params.kappa).all() and np.abs(d).max() == 1   # challenge_space.rs:65-71
    z = prover.create_response_batch(s["y"], s["r"], d)
    ok = verifier.verify_batch(z, s["t"], np.ascontiguousarray(s["c"][:, :1]), d)
    assert ok.all()
