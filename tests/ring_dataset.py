"""Synthetic TempleRing stand-in with real 3-D structure (TEST INFRASTRUCTURE).

The Middlebury TempleRing images are not shipped with the reference, and the value-noise frames of
`sfmgpu.synth` are 2-D translations of one texture: a planar, zero-parallax scene with no ground-truth trajectory to score
against.  This module ray-casts a small textured scene (a sphere in front of a wall, both
carrying a smooth analytic 3-D texture) from cameras on a ring around it, Middlebury temple intrinsics, and writes the
directory layout the reference's CLI reads (`templeRing/templeR_par.txt`, `templeR_ang.txt`, `templeRing_pgm/*.pgm`,
/root/reference/cpp/src/templering_sfm.cpp:120-152, :1678-1712) with the TRUE poses in the par file, so that the
reference's own `ate_keyframes` tool can score a run against ground truth.

The angular step per frame is tiny on purpose: the reference's `lk_step` evaluates the error at the SAME location in both
images (sfm.cpp:439-442), which multiplies small flows by ~31.5 (SURVEY.md §8c KAT 4); a true flow of ~0.05 px per frame
keeps the tracked flow near the 18 px the reference's keyframe rule asks for (sfm.cpp:1576, :1703).  The tests run it with
bundle adjustment off: the reference's map points depend on heap contents (tests/test_gpu_dropin.py explains).
"""
import os

import numpy as np

K_TEMPLE = np.array([[1520.4, 0.0, 302.32], [0.0, 1525.9, 246.87], [0.0, 0.0, 1.0]])


def _texture(P, seed):
    """Smooth analytic 3-D texture in [0, 1] at points P[..., 3]: a sum of sinusoids with random directions."""
    rng = np.random.RandomState(seed)
    acc = np.zeros(P.shape[:-1])
    wsum = 0.0
    for lam, wgt, cnt in ((0.030, 1.0, 5), (0.012, 0.8, 7), (0.005, 0.5, 9)):  # wavelengths in metres
        for _ in range(cnt):
            d = rng.normal(size=3)
            d /= np.linalg.norm(d)
            ph = rng.uniform(0, 2 * np.pi)
            acc += wgt * np.sin((P @ d) * (2 * np.pi / lam) + ph)
            wsum += wgt
    return 0.5 + 0.5 * np.tanh(2.5 * acc / np.sqrt(wsum * 3.0))


def camera(theta, radius=0.60, height=0.05):
    """World->camera (R, t) of a camera on the ring at angle theta (radians), looking at the origin; x right, y down."""
    C = np.array([radius * np.sin(theta), -height, -radius * np.cos(theta)])
    z = -C / np.linalg.norm(C)
    down = np.array([0.0, 1.0, 0.0])  # the world's y axis points down, like the camera's
    x = np.cross(down, z)
    x /= np.linalg.norm(x)
    y = np.cross(z, x)
    R = np.stack([x, y, z])  # rows: camera axes in world coordinates
    return R, -R @ C, C


def render(theta, w=640, h=480, K=K_TEMPLE, seed=7):
    R, t, C = camera(theta)
    u, v = np.meshgrid(np.arange(w, dtype=np.float64), np.arange(h, dtype=np.float64))
    d_cam = np.stack([(u - K[0, 2]) / K[0, 0], (v - K[1, 2]) / K[1, 1], np.ones_like(u)], -1)
    d = d_cam @ R  # = R^T d_cam per pixel: ray directions in the world
    d /= np.linalg.norm(d, axis=-1, keepdims=True)
    # sphere (radius 0.09 at the origin)
    rs = 0.09
    b = d @ C
    disc = b * b - (C @ C - rs * rs)
    hit_s = disc > 0
    ts = -b - np.sqrt(np.where(hit_s, disc, 0.0))
    # wall: plane z_w = 0.25 .. tilted slightly so that it is not fronto-parallel to any camera
    n = np.array([0.15, 0.05, 1.0])
    n /= np.linalg.norm(n)
    d0 = 0.25
    tw = (d0 - C @ n) / (d @ n)
    tt = np.where(hit_s, ts, tw)
    P = C + tt[..., None] * d
    tex = np.where(hit_s, _texture(P, seed), _texture(P * 0.7 + 0.3, seed + 1))
    img = 20.0 + 215.0 * tex
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def write_dataset(root, nframes=24, step_deg=0.02, w=640, h=480):
    """Write the dataset; returns the list of image names."""
    os.makedirs(os.path.join(root, "templeRing"), exist_ok=True)
    os.makedirs(os.path.join(root, "templeRing_pgm"), exist_ok=True)
    par, ang, names = [str(nframes)], [], []
    for i in range(nframes):
        th = np.deg2rad(step_deg * i)
        name = f"templeR{i + 1:04d}"
        img = render(th, w, h)
        with open(os.path.join(root, "templeRing_pgm", name + ".pgm"), "wb") as f:
            f.write(f"P5\n{w} {h}\n255\n".encode())
            f.write(img.tobytes())
        R, t, _ = camera(th)
        vals = list(K_TEMPLE.ravel()) + list(R.ravel()) + list(t)
        par.append(name + ".png " + " ".join(repr(float(x)) for x in vals))
        ang.append(f"{10.0} {step_deg * i} {name}.png")
        names.append(name + ".png")
    open(os.path.join(root, "templeRing", "templeR_par.txt"), "w").write("\n".join(par) + "\n")
    open(os.path.join(root, "templeRing", "templeR_ang.txt"), "w").write("\n".join(ang) + "\n")
    return names
