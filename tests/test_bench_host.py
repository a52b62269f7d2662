"""CPU tests of bench.py's host-side logic: the wire format of the N > 1 gather (one packed byte block per rank, li and
inlier indices as int16), the shard partition, and the reference arm's line (runs the compiled reference on a tiny
workload)."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_wire_layout_roundtrip():
    rng = np.random.default_rng(1)
    cap, P, world = 37, 23, 4
    shards = [bench.shard_range(P, world, r) for r in range(world)]
    assert shards[0][0] == 0 and shards[-1][1] == P and all(a[1] == b[0] for a, b in zip(shards, shards[1:]))
    full = {"li": rng.integers(0, 4000, (P, cap, 2)).astype(np.int16), "lj": rng.normal(0, 1000, (P, cap, 2)),
            "nk": rng.integers(0, cap, P).astype(np.int32), "nc": rng.integers(0, cap, P).astype(np.int32),
            "status": rng.integers(0, 3, P).astype(np.int32), "best": rng.integers(-1, 4000, (P, 2)).astype(np.int32),
            "inliers": rng.integers(0, cap, (P, cap)).astype(np.int16), "R": rng.normal(size=(P, 9)), "t": rng.normal(size=(P, 3))}

    class _T:  # stands in for a pinned torch tensor: .numpy() gives the byte block
        def __init__(self, a):
            self.a = a

        def numpy(self):
            return self.a

    host = {}
    for r, (a, b) in enumerate(shards):
        lay, total = bench.wire_layout(b - a, cap)
        assert total % 8 == 0 and [k for k, *_ in lay] == [k for k, *_ in bench.WIRE]
        block = np.zeros(max(total, 8), np.uint8)
        for key, dt, shape, off in lay:
            src = np.ascontiguousarray(full[key][a:b]).view(np.uint8).reshape(-1)
            assert src.dtype == np.uint8 and full[key].dtype == dt and full[key][a:b].shape == shape
            block[off:off + src.size] = src
        host[r] = _T(block)
    got = bench.decode_gathered({"host": host}, shards, P, cap)
    for key in full:
        assert np.array_equal(got[key], full[key]), key


def test_reference_arm_line_tiny():
    """`bench.py --impl reference` prints ONE JSON line with the contract's keys (tiny frames so that it runs in seconds)."""
    env = dict(os.environ)
    code = ("import sys, bench; bench.WORKLOADS['c2'].update(W=160, H=120, corners=60); "
            "sys.argv = ['bench.py', '--impl', 'reference', '--workload', 'c2', '--steps', '1', '--warmup', '0']; bench.main()")
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "feature-tracks/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["e2e"]["h2d_bytes_per_step"] == 0
    assert line["scaling"] == "strong" and line["higher_is_better"] is True
