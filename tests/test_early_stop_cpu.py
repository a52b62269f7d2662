"""CPU model of the early stop of the batched RANSAC stage (csrc/two_view.cu: tv_early_kernel).

Reference rule (cpp/src/templering_sfm.cpp:667-677): hypotheses are scored in order and `if (inl.size() > best_inl.size())`
replaces the winner - the FIRST hypothesis with the largest count wins, a hypothesis with 0 inliers never does, and a
count cannot exceed the number of points n.  The stage scores hypotheses [0, 128), and if one of them has count n it
leaves the counts of [128, H) at 0; the winner is then taken by the same arg-max (largest count, lowest index).  This
test brute-forces that the two procedures pick the same winner for every count vector, including ties, full counts at
every position relative to the probe, and all-zero vectors."""
import numpy as np

PROBE = 128  # TV_EARLY_H


def reference_winner(counts):
    best, best_n = -1, 0
    for h, c in enumerate(counts):  # the reference's loop, literally
        if c > best_n:
            best, best_n = h, int(c)
    return best, best_n


def staged_winner(counts, n):
    c = np.array(counts, np.int64)
    if n >= 8 and np.any(c[:PROBE] == n):
        c[PROBE:] = 0  # never solved, never scored
    key_best, best, best_n = 0, -1, 0
    for h, v in enumerate(c):  # ransac_argmax_kernel: max of (count << 32 | ~h), -1 when every count is 0
        key = (int(v) << 32) | (0xFFFFFFFF - h)
        if v > 0 and key > key_best:
            key_best, best, best_n = key, h, int(v)
    return best, best_n


def test_staged_winner_equals_the_reference_loop():
    rng = np.random.default_rng(12)
    cases = 0
    for n in (8, 9, 60, 2000):
        for H in (129, 257, 300, 1000):
            for trial in range(60):
                mode = trial % 6
                if mode == 0:
                    c = rng.integers(0, n + 1, H)  # anything
                elif mode == 1:
                    c = rng.integers(max(0, n - 2), n + 1, H)  # crowded at the top: ties everywhere
                elif mode == 2:
                    c = rng.integers(0, n, H)  # never full
                elif mode == 3:
                    c = rng.integers(0, n, H)
                    c[rng.integers(0, PROBE)] = n  # one full count inside the probe
                elif mode == 4:
                    c = rng.integers(0, n, H)
                    c[rng.integers(PROBE, H)] = n  # one full count beyond the probe: no stop, it still wins
                else:
                    c = np.zeros(H, np.int64)  # no inlier anywhere: no winner
                    if trial % 12 == 5:
                        c[rng.integers(0, H)] = rng.integers(1, n + 1)
                assert staged_winner(c, n) == reference_winner(c), (n, H, mode)
                cases += 1
    # the boundary positions of a full count
    for pos in (0, 1, PROBE - 1, PROBE, PROBE + 1, 299):
        c = np.full(300, 7, np.int64)
        c[pos] = 9
        assert staged_winner(c, 9) == reference_winner(c) == (pos, 9)
    assert cases == 4 * 4 * 60
