"""Build / iteration round statistics of klt_quad_kernel (library built with -DKLT_ROUND_STATS, SFMGPU_LIB=...)."""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "structure-from-motion-3d-reconstruction_b200"))
import sfmgpu
W, H, F, MC = (int(a) for a in (sys.argv[1:5] + ["1920", "1080", "100", "2000"][len(sys.argv) - 1:]))
ctx = sfmgpu.Context(0)
fr = ctx.frames(W, H, F, 3)
fr.synth(0, F, 20261018, 0)
fr.build_pyramid()
pairs = ctx.pairs(F - 1, MC)
pairs.run(fr, 0, F - 1, sfmgpu.lkcfg(max_tracks=MC))
ctx.sync()
out = (C.c_ulonglong * 66)()
rc = ctx.lib.sfmgpu_debug_klt_round_hist(out)
h = np.array(list(out), dtype=np.float64).reshape(2, 33)
for name, row in zip(("build rounds by lanes building", "iteration rounds by lanes iterating"), h):
    tot = row.sum()
    print(name, "total", int(tot), "per warp-of-32-tracks", tot / ((F - 1) * MC / 32))
    print("  lanes: share of rounds | share of lane-work")
    lanes = np.arange(33)
    work = row * lanes
    for lo, hi in ((1, 1), (2, 2), (3, 4), (5, 8), (9, 16), (17, 24), (25, 31), (32, 32)):
        print(f"  {lo:2d}-{hi:2d}: {row[lo:hi+1].sum()/tot:6.3f} {work[lo:hi+1].sum()/work.sum():6.3f}")
