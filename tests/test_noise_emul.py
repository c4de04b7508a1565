"""CPU check of the parallel noise baseline (bronko_b200/csrc/bk_noise.cuh): the device primitives — exact chain
rounds (integer prefix sums over parity maps), speculative table chunks + boundary verification + replay, Thompson
tau — stepped on the CPU by tests/emul and compared bit for bit with the oracle's sequential restatement of
get_baseline_noise (reference src/call.rs:799-967)."""
import numpy as np
import pytest

import emul_lib
from oracle import oracle as O


def _check(fwd, rev):
    mx, _, _ = O.baseline_noise(np.ascontiguousarray(fwd, dtype=np.uint64), np.ascontiguousarray(rev, dtype=np.uint64))
    got, stats = emul_lib.noise(fwd, rev)
    assert np.array_equal(mx, got), "Noise.max differs at %s" % np.nonzero(mx != got)[0][:8]
    return stats


@pytest.mark.parametrize("seed", range(8))
def test_random_pileups(seed):
    rng = np.random.default_rng(seed)
    L = int(rng.integers(100, 4000))
    depth = int(rng.choice([1, 3, 10, 100, 5000, 200000]))
    p = np.array([0.97, 0.01, 0.01, 0.01])
    f = rng.poisson(depth * p, size=(L, 4)).astype(np.uint32)
    r = rng.poisson(depth * p, size=(L, 4)).astype(np.uint32)
    hole = rng.random(L) < rng.choice([0.0, 0.5, 0.95])
    f[hole] = 0
    r[hole] = 0
    for pos in rng.integers(0, L, 5):                 # a few iSNV-like positions: large minor fractions
        f[pos, 1] += depth // 3 + 1
    _check(f, r)


def test_duplicates_and_ties():
    rng = np.random.default_rng(7)
    L = 2500
    f = np.zeros((L, 4), np.uint32)
    r = np.zeros((L, 4), np.uint32)
    f[:, 0] = 1000
    r[:, 0] = 1000
    f[:, 1] = rng.integers(0, 3, L)
    f[:, 2] = rng.integers(0, 2, L)
    _check(f, r)                                      # many identical fractions: evict-by-value hits duplicates
    f[:] = 1
    _check(f, r * 0)                                  # every fraction exactly 1/4: exact ties in the sums
    f = rng.integers(0, 4, (L, 4)).astype(np.uint32)
    stats = _check(f, f * 0)                          # tiny counts: the speculative table converges slowly
    assert stats[2] > 0                               # chain rounds ran
    f = (rng.integers(0, 2, (L, 4)) * rng.integers(1, 1000000, (L, 4))).astype(np.uint32)
    _check(f, f)                                      # operands as large as the sum: binade changes all the time


def test_window_edge_lengths():
    rng = np.random.default_rng(3)
    for L in (100, 101, 127, 128, 129, 178, 206, 255, 256, 257, 334, 462, 1024, 1074):
        f = rng.poisson(50 * np.array([0.9, 0.05, 0.03, 0.02]), size=(L, 4)).astype(np.uint32)
        _check(f, f[::-1].copy())
    out, _ = emul_lib.noise(np.ones((50, 4), np.uint32), np.ones((50, 4), np.uint32))
    assert (out == 0).all()                           # the reference panics below 100 positions: zero noise reported


def test_simulated_sample():
    import bronko_b200
    from bronko_b200 import sim
    from util import oracle_sample
    g = sim.load_genome(sim.HPV16)
    r1, o1, r2, o2, _ = sim.simulate_pairs(g, 400, sim.SEED0 + 5)
    oi = O.Index.build(21, [sim.genome_path(sim.HPV16)])
    _, s = oracle_sample(oi, [(r1, o1), (r2, o2)], bronko_b200.CallArgs())
    p = s.pileup()
    got, _ = emul_lib.noise(np.asarray(p[0]), np.asarray(p[1]))
    assert np.array_equal(got, np.asarray(s.noise_max()))


def test_look_ahead_places_the_zones():
    """nz_hint_*: the lowest zone holds the start, covers the longest run ahead, and a spare zone goes below the run."""
    ef = lambda x: int(np.float64(x).view(np.uint64) >> np.uint64(52))
    assert emul_lib.hint(0.3, [0.3] * 200) == (200, ef(0.3) - 1)                       # one binade: room on both sides
    assert emul_lib.hint(0.3, [0.3, 0.2, 0.1, 0.07]) == (4, ef(0.07))                  # three binades down: the lowest one is zone 0
    assert emul_lib.hint(0.3, [0.3, 0.2, 0.1, 0.07, 0.03, 0.3]) == (4, ef(0.07))       # a fourth does not fit: the run ends before it
    assert emul_lib.hint(0.3, [0.6, 1.1]) == (2, ef(0.3))                              # upwards: the start is the lowest zone
    assert emul_lib.hint(0.3, [0.6, 0.3, 0.6]) == (3, ef(0.3) - 1)                     # two binades: the spare one below
    run, el = emul_lib.hint(0.3, [0.3, 0.0, 0.3])                                      # an empty window ends the run
    assert run == 1 and el == ef(0.3) - 1
    run, el = emul_lib.hint(0.3, [0.25 * (1 + 1e-12)] * 5)                             # next to a power of two: both binades count
    assert run == 5 and el == ef(0.2) - 1                                              # (0.125..0.25 and 0.25..0.5 → spare below)
    assert emul_lib.hint(0.0, [0.3] * 5)[1] == 0 and emul_lib.hint(-1e-18, [0.3] * 5)[1] == 0    # nothing to scale by: serial
    for s0 in (0.3, 0.26, 0.49):                                                       # the zones always hold the start
        for ahead in ([0.9], [0.07], [0.13, 0.6], [1e-9], [3.0]):
            el = emul_lib.hint(s0, ahead)[1]
            assert ef(s0) - 2 <= el <= ef(s0)
