"""Generates tests/golden/frontend_golden.npz from the COMPILED REFERENCE (oracle/_ref/libsfmref.so, i.e. the
unmodified cpp/src/templering_sfm.cpp built by oracle/Makefile).  Run in the build container, where
/root/reference exists:  python tests/golden/gen_golden.py
Inputs are reproducible from seeds (sfmgpu/synth.py, numpy PCG64), so only outputs are stored (< 100 KB).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "structure-from-motion-3d-reconstruction_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle  # noqa: E402
from conftest import TEMPLE_K, two_view_scene  # noqa: E402
from sfmgpu import synth  # noqa: E402

W, H, SEED = 320, 240, 20261018


def main():
    ref = oracle.ref()
    assert ref is not None, "needs the compiled reference"
    g = {}
    f = [synth.frame(SEED, t, W, H) for t in range(5)]
    g["frame_fnv"] = np.array([synth.fnv1a64(x.tobytes()) for x in f], np.uint64)
    pyr = ref.build_pyr(f[0], 3)
    g["pyr_l1"], g["pyr_l2"] = pyr[1], pyr[2]
    g["corners"] = ref.shi_tomasi(f[0], 400, 0.01, 8).astype(np.int16)
    ties = (f[0] >> 4 << 4)
    g["corners_ties"] = ref.shi_tomasi(ties, 400, 0.01, 8).astype(np.int16)
    kat = synth.kat_image(W, H)
    g["corners_kat"] = ref.shi_tomasi(kat, 600, 0.01, 8).astype(np.int16)
    p0 = g["corners"].astype(np.float64)[:200]
    g["klt_p1"], g["klt_pb"] = ref.klt_track(f[0], f[1], p0, 3, 5, 10)
    trk = ref.tracker(max_tracks=150, min_tracks=120, quality=0.01, min_distance=8, levels=3, radius=5, iters=10, fb=1.0)
    for i, t in enumerate([0, 1, 2, 40, 41]):
        prev, cur, ids = trk.step(synth.frame(SEED, t, W, H))
        xy, tid = trk.tracks()
        g[f"trk{i}_cur"], g[f"trk{i}_ids"], g[f"trk{i}_tracks"], g[f"trk{i}_tids"] = cur, ids, xy, tid
    pi, pj = two_view_scene(300)
    xi, xj = ref.norm_points(TEMPLE_K, pi), ref.norm_points(TEMPLE_K, pj)
    E, idx = ref.ransac_hypotheses(xi, xj, 40)
    counts, bh, inl = ref.ransac_score(xi, xj, E, 1e-3)
    g["rs_xi"], g["rs_xj"], g["rs_E"], g["rs_idx"] = xi, xj, E, idx
    g["rs_counts"], g["rs_best"], g["rs_inl"] = counts, np.array([bh]), inl
    R, t, finl = ref.find_E_ransac(TEMPLE_K, pi, pj, 40, 1e-3, 60)
    g["fe_R"], g["fe_t"], g["fe_inl"] = R, t, finl
    li, lj, nc = ref.pair_frontend(f[2], f[3], 300)
    g["pair_li"], g["pair_lj"], g["pair_nc"] = li, lj, np.array([nc])
    out = os.path.join(ROOT, "tests", "golden", "frontend_golden.npz")
    np.savez_compressed(out, **g)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
