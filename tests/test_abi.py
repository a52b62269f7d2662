"""C-ABI checks that need no GPU: the library loads, exports every symbol include/sfmgpu.h declares, and refuses
to create a context without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import sfmgpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "sfmgpu.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(sfmgpu_[a-z0-9_]+)\s*\(", txt)))


def test_header_and_binding_agree():
    assert header_symbols() == sorted(sfmgpu.SYMBOLS)


def test_library_exports_every_declared_symbol():
    if not os.path.exists(sfmgpu.LIB_PATH):
        sfmgpu.build_library()
    lib = C.CDLL(sfmgpu.LIB_PATH)
    for s in header_symbols():
        assert hasattr(lib, s), s
    assert lib.sfmgpu_version() == 100


def test_lkcfg_layout_and_defaults():
    lib = sfmgpu.load_library()
    c = sfmgpu.LKCfg()
    lib.sfmgpu_lkcfg_default(C.byref(c))
    assert (c.max_tracks, c.min_tracks, c.quality, c.min_distance, c.pyr_levels, c.win_radius, c.iters, c.fb_thresh) == (
        2200, 900, 0.01, 8, 3, 5, 10, 1.0)  # LKConfig defaults, cpp/src/templering_sfm.cpp:307-316
    assert C.sizeof(sfmgpu.LKCfg) == 40


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(sfmgpu.SfmGpuError):
        sfmgpu.Context(0)


def test_synth_matches_known_hash():
    from sfmgpu import synth
    f = synth.frame(20261018, 3, 64, 48)
    assert f.dtype == np.uint8 and f.shape == (48, 64)
    assert synth.tri(0) == 0 and synth.tri(64) == 64 and synth.tri(100) == 28 and synth.tri(128) == 0
    assert np.array_equal(f, synth.frame(20261018, 3, 64, 48))
    assert not np.array_equal(f, synth.frame(20261018, 4, 64, 48))
