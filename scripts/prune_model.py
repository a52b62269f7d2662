#!/usr/bin/env python
"""CPU model of EXACT count-bound pruning for the RANSAC scoring loop (a candidate next step; not built).

The reference keeps the first hypothesis with the largest inlier count (cpp/src/templering_sfm.cpp:667-677).  If the points
are scored in blocks, a hypothesis whose count so far plus the points still to come is below the best FULL count known
cannot be that hypothesis, and scoring it further is wasted - exactly, for any data.  This script measures, on the C4 scene
(10,000 correspondences, 30 % outliers, thr 1e-3, the reference's own hypotheses), what fraction of the hyp x pts
evaluations such a scheme would skip: hypotheses in launch sets of `hset`, every set scored block by block (`block` points),
a hypothesis dropped when count + remaining < the best lower bound known of the winning count (the best full count of the
earlier sets, or the largest count so far in this set).
usage: python scripts/prune_model.py [H=4096] [hset=1024] [block=512]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "structure-from-motion-3d-reconstruction_b200"))
from oracle import oracle  # noqa: E402  (test infrastructure: this is an offline model, not a product path)
from conftest import TEMPLE_K, two_view_scene  # noqa: E402


def sampson_mask(E, x, xp, thr):
    """inlier mask [H, n] in the reference's formulation (FP64)."""
    ex = E[:, 0:1] * x[:, 0] + E[:, 1:2] * x[:, 1] + E[:, 2:3]
    ey = E[:, 3:4] * x[:, 0] + E[:, 4:5] * x[:, 1] + E[:, 5:6]
    ez = E[:, 6:7] * x[:, 0] + E[:, 7:8] * x[:, 1] + E[:, 8:9]
    tx = E[:, 0:1] * xp[:, 0] + E[:, 3:4] * xp[:, 1] + E[:, 6:7]
    ty = E[:, 1:2] * xp[:, 0] + E[:, 4:5] * xp[:, 1] + E[:, 7:8]
    num = xp[:, 0] * ex + xp[:, 1] * ey + ez
    den = ex * ex + ey * ey + tx * tx + ty * ty + 1e-12
    return (num * num) / den < thr


def main():
    H = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    hset = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    block = int(sys.argv[3]) if len(sys.argv) > 3 else 512
    n, thr = 10000, 1e-3
    ref = oracle.best()[0]
    pi, pj = two_view_scene(n)
    x, xp = ref.norm_points(TEMPLE_K, pi), ref.norm_points(TEMPLE_K, pj)
    E, _ = ref.ransac_hypotheses(x, xp, H)
    counts_ref, bh, _ = ref.ransac_score(x, xp, E, thr)
    total, done = H * n, 0
    best_full, winner = 0, -1
    for h0 in range(0, H, hset):
        Es = E[h0:h0 + hset]
        alive = np.ones(len(Es), bool)
        cnt = np.zeros(len(Es), np.int64)
        for p0 in range(0, n, block):
            idx = np.nonzero(alive)[0]
            if not len(idx):
                break
            m = sampson_mask(Es[idx], x[p0:p0 + block], xp[p0:p0 + block], thr)
            cnt[idx] += m.sum(1)
            done += len(idx) * m.shape[1]
            left = n - min(n, p0 + block)
            # strictly below the best full count: can neither beat nor tie-before it (earlier sets have lower indices)
            # a hypothesis's count so far is a lower bound of its full count, so the largest one is a lower bound of the
            # winning count as well
            alive[idx] = cnt[idx] + left >= max(best_full, int(cnt[idx].max()), 1)
        full = np.nonzero(alive)[0]
        for k in full:  # survivors carry exact counts; first strictly larger wins, as in the reference
            if cnt[k] > best_full:
                best_full, winner = int(cnt[k]), h0 + int(k)
    assert (winner, best_full) == (bh, int(counts_ref[bh])), ((winner, best_full), (bh, int(counts_ref[bh])))
    print(f"H {H}, sets of {hset}, blocks of {block} points: winner {winner} with {best_full} inliers (reference: the same); "
          f"{done / total:.3f} of the {total} hyp x pts evaluations performed ({1 - done / total:.1%} skipped)")


if __name__ == "__main__":
    main()
