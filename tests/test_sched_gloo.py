"""The N>1 host path on CPU: world_size-2 gloo processes shard the pairs of a sequence, each computes its block
(with the CPU checker standing in for the GPU), and the results are gathered to rank 0 by sfmgpu.sched — they must
equal the single-process run, in global pair order."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sfmgpu import sched, synth

W, H, NFR, CAP = 128, 96, 6, 60


def test_shard_ranges_cover_everything():
    for n in (0, 1, 5, 8, 999, 1000):
        for world in (1, 2, 3, 8):
            blocks = [sched.shard_range(n, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1
    p0, p1, fr = sched.pair_shard(1000, 8, 3)
    assert fr == (p0, p1 + 1) and (p1 - p0) in (124, 125)
    segs = [sched.shard_range(2000, 8, r) for r in range(8)]
    assert segs[0] == (0, 250) and segs[-1] == (1750, 2000) and all(segs[i][1] == segs[i + 1][0] for i in range(7))


def _worker(rank, world, port, out_path):
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle"))
    import oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    chk = oracle.port()
    p0, p1, (f0, f1) = sched.pair_shard(NFR, world, rank)
    frames = [synth.frame(9, t, W, H) for t in range(f0, f1 + 1)]
    n = p1 - p0
    nk = torch.zeros(n, dtype=torch.int32)
    li = torch.zeros((n, CAP, 2), dtype=torch.float64)
    lj = torch.zeros((n, CAP, 2), dtype=torch.float64)
    for k in range(n):
        a, b, _ = chk.pair_frontend(frames[k], frames[k + 1], CAP)
        nk[k] = len(a)
        li[k, :len(a)] = torch.from_numpy(a)
        lj[k, :len(b)] = torch.from_numpy(b)
    res = sched.gather_pair_results(nk, li, lj, dst=0)
    if rank == 0:
        counts, li_list, lj_list = res
        np.savez(out_path, counts=counts.numpy(), li=np.concatenate([x.numpy() for x in li_list]),
                 lj=np.concatenate([x.numpy() for x in lj_list]))
    else:
        assert res is None
    dist.destroy_process_group()


def test_two_rank_gather_equals_single_process(tmp_path, port):
    out = str(tmp_path / "gathered.npz")
    mp.spawn(_worker, args=(2, 29600 + os.getpid() % 300, out), nprocs=2, join=True)
    g = np.load(out)
    counts, li, lj = [], [], []
    frames = [synth.frame(9, t, W, H) for t in range(NFR)]
    for k in range(NFR - 1):
        a, b, _ = port.pair_frontend(frames[k], frames[k + 1], CAP)
        counts.append(len(a)); li.append(a); lj.append(b)
    assert g["counts"].tolist() == counts
    assert np.array_equal(g["li"], np.concatenate(li)) and np.array_equal(g["lj"], np.concatenate(lj))
