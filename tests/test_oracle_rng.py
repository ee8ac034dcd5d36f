"""The oracle's restatements of third-party arithmetic, checked against the real thing (numpy and
CPython are present on every box): MT19937 seeding and 53-bit doubles, numpy's legacy Poisson
draw for draw, Philox4x32-10 known answers (Random123 kat_vectors)."""
import random

import numpy as np
import pytest

from oracle import oracle


@pytest.mark.parametrize('seed', [0, 1, 12345, 2**32 - 1])
def test_mt_numpy_seeding_matches_randomstate(seed):
    rs = np.random.RandomState(seed)
    rng = oracle.Rng('numpy', seed)
    ours = [rng.random() for _ in range(1500)]
    assert ours == list(rs.random_sample(1500))


@pytest.mark.parametrize('seed', [0, 1, 987654321, 2**40 + 17])
def test_mt_python_seeding_matches_random(seed):
    random.seed(seed)
    rng = oracle.Rng('python', seed)
    assert [rng.random() for _ in range(1500)] == [random.random() for _ in range(1500)]


@pytest.mark.parametrize('lam', [0.0, 0.002, 0.5, 2.0, 9.99, 10.0, 37.3, 638.0, 1e5])
def test_poisson_matches_numpy_legacy_draw_for_draw(lam):
    rs = np.random.RandomState(99)
    rng = oracle.Rng('numpy', 99)
    theirs = rs.poisson(lam, 5000)
    ours = [rng.poisson(lam) for _ in range(5000)]
    assert ours == list(theirs)
    # generator states still synchronised afterwards
    assert rng.random() == rs.random_sample()


def test_poisson_scalar_call_pattern_of_the_reference():
    # basepop.py:830 draws `numpy.random.poisson(mean, 1)[0]` from the GLOBAL legacy generator
    np.random.seed(4)
    rng = oracle.Rng('numpy', 4)
    for lam in [0.3, 12.5, 0.0, 2.0, 55.0] * 40:
        assert rng.poisson(lam) == np.random.poisson(lam, 1)[0]


def test_philox4x32_10_known_answers():
    kat = [
        ([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
        ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
        ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
         [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
    ]
    for ctr, key, want in kat:
        assert list(oracle.philox4x32_10(ctr, key)) == want


def test_philox_stream_layout():
    rng = oracle.Rng('philox', seed=0x0123456789abcdef, member=5)
    a, b, c = rng.random(), rng.random(), rng.random()
    o0 = oracle.philox4x32_10([0, 0, 5, 0], [0x89abcdef, 0x01234567])
    o1 = oracle.philox4x32_10([1, 0, 5, 0], [0x89abcdef, 0x01234567])
    pack = lambda x, y: ((int(x) >> 5) * 67108864.0 + (int(y) >> 6)) / 9007199254740992.0
    assert (a, b, c) == (pack(o0[0], o0[1]), pack(o0[2], o0[3]), pack(o1[0], o1[1]))
    assert 0.0 <= a < 1.0


def test_injected_stream_and_exhaustion():
    u = np.array([0.25, 0.5, 0.75])
    rng = oracle.Rng('inject', uniforms=u)
    assert [rng.random() for _ in range(3)] == [0.25, 0.5, 0.75]
    assert not rng.exhausted
    rng.random()
    assert rng.exhausted
