"""The multi-GPU scheduler entry points of the C ABI (include/sfmgpu.h: sfmgpu_sched_*).

tests/sched_test (C++, built by the package Makefile from tests/cpp/sched_test.cpp) runs one host thread per rank / GPU:
every rank computes its block of pairs of ONE sequence with the RANSAC stage on, the results are gathered to rank 0 over
NCCL (ncclSend / ncclRecv) and compared byte for byte with a single-GPU run of the whole sequence.  On a 1-GPU box the
world degenerates to one rank (no communicator; the gather is a local copy); `gpurun --gpus 2` exercises NCCL."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import sfmgpu

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def test_sched_cpp_program():
    exe = os.path.join(HERE, "sched_test")
    assert os.path.exists(exe), "tests/sched_test is missing: run __graft_entry__.build()"
    r = subprocess.run([exe, "2", "13"], capture_output=True, text=True, timeout=600)
    print(r.stdout, r.stderr)
    assert r.returncode == 0 and "OK" in r.stdout


def test_sched_single_rank_through_ctypes(ctx):
    """world = 1: shard = everything, gather = the pairs object's own results in host arrays."""
    from sfmgpu import synth
    lib = ctx.lib
    a, b = C.c_int(0), C.c_int(0)
    assert lib.sfmgpu_sched_shard(1999, 8, 7, C.byref(a), C.byref(b)) == 0 and (a.value, b.value) == (1750, 1999)
    assert lib.sfmgpu_sched_shard(5, 0, 0, C.byref(a), C.byref(b)) != 0
    s = C.c_void_p()
    ctx._ck(lib.sfmgpu_sched_create(ctx.h, 1, 0, None, None, C.byref(s)))
    p0, p1, f0, f1 = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    ctx._ck(lib.sfmgpu_sched_pair_shard(s, 6, C.byref(p0), C.byref(p1), C.byref(f0), C.byref(f1)))
    assert (p0.value, p1.value, f0.value, f1.value) == (0, 5, 0, 6)
    W, H = 320, 240
    f = ctx.frames(W, H, 6, 3)
    f.upload(0, np.stack([synth.frame(5, t, W, H) for t in range(6)]))
    f.build_pyramid()
    pairs = ctx.pairs(5, 200)
    pairs.run(f, 0, 5, sfmgpu.lkcfg(max_tracks=200))
    li, lj = np.zeros((5, 200, 2)), np.zeros((5, 200, 2))
    nk, nc = np.zeros(5, np.int32), np.zeros(5, np.int32)
    vp = lambda x: x.ctypes.data_as(C.c_void_p)
    ctx._ck(lib.sfmgpu_sched_gather_pairs(ctx.h, s, pairs.h_, 5, 0, 0, vp(li), vp(lj), vp(nk), vp(nc), None, None, None, None, None))
    li2, lj2 = np.zeros((5, 200, 2)), np.zeros((5, 200, 2))
    nk2, nc2 = np.zeros(5, np.int32), np.zeros(5, np.int32)
    pairs.download_all(li2, lj2, nk2, nc2)
    assert np.array_equal(nk, nk2) and np.array_equal(nc, nc2)
    for p in range(5):
        assert np.array_equal(li[p, :nk[p]], li2[p, :nk[p]]) and np.array_equal(lj[p, :nk[p]], lj2[p, :nk[p]])
    # a rank whose block size does not match is refused, loudly
    assert lib.sfmgpu_sched_gather_pairs(ctx.h, s, pairs.h_, 7, 0, 0, None, None, None, None, None, None, None, None, None) != 0
    lib.sfmgpu_sched_destroy(ctx.h, s)
