"""CPU: the numpy restatement of the optional samplers (tests/philox_ref.py) has the distributions the reference's
samplers have (polynomial.rs:14-44, challenge_space.rs:12-33) and the published Philox4x32-10 known answers."""
import numpy as np

import philox_ref as pr


def test_philox_known_answers():
    # Random123 kat_vectors: philox4x32 10 rounds
    def one(ctr, key):
        seed = key[0] | (key[1] << 32)
        return tuple(int(v) for v in pr.philox(*[np.uint64(c) for c in ctr], seed))
    assert one((0, 0, 0, 0), (0, 0)) == (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)
    assert one((0xffffffff,) * 4, (0xffffffff, 0xffffffff)) == (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)
    assert one((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0)) == (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)


def test_small_is_uniform():
    for b in (1, 3, 127):
        v = pr.sample_small(64, b, seed=12345, tag=1).astype(np.int64).ravel()
        assert v.min() == -b and v.max() == b
        counts = np.bincount(v + b, minlength=2 * b + 1)
        exp = v.size / (2 * b + 1)
        chi2 = ((counts - exp) ** 2 / exp).sum()
        dof = 2 * b
        assert chi2 < dof + 6 * np.sqrt(2 * dof) + 10, (b, chi2)
    assert not (pr.sample_small(2, 1, 1, 1) == pr.sample_small(2, 1, 2, 1)).all()       # seed matters
    assert not (pr.sample_small(2, 1, 1, 1) == pr.sample_small(2, 1, 1, 2)).all()       # tag matters
    assert (pr.sample_small(4, 1, 9, 3)[:2] == pr.sample_small(2, 1, 9, 3)).all()       # counter based: prefix stable


def test_gaussian_moments():
    sigma = 15444.0
    v = pr.sample_gaussian(256, sigma, seed=77, tag=2).astype(np.float64).ravel()
    n = v.size
    assert abs(v.mean()) < 5 * sigma / np.sqrt(n)
    assert abs(v.std() / sigma - 1) < 0.01
    assert abs(((v / sigma) ** 4).mean() - 3) < 0.15                                    # kurtosis of a normal
    assert np.abs(v).max() < 7 * sigma
    assert abs((np.abs(v) < sigma).mean() - 0.6827) < 0.01


def test_challenge_structure():
    d = pr.sample_challenge(200, 36, seed=5, tag=3)
    assert (np.abs(d).sum(axis=1) == 36).all() and np.abs(d).max() == 1                 # challenge_space.rs:65-71
    assert abs(d.sum() / (200 * 36)) < 0.05                                             # signs balanced
    pos = np.nonzero(d)[1]
    assert abs(pos.mean() - 255.5) < 10                                                 # positions spread over [0, N)
    assert (np.abs(pr.sample_challenge(3, 600, 5, 3)).sum(axis=1) == 512).all()         # kappa > N: min(kappa, N)
