"""Integer-only synthetic sequences: identical bytes from numpy here and from the CUDA
generator in csrc/synth.cu, so the CPU baseline and the GPU path see the same input without moving frames.

Frame t of sequence `seed` samples a static random field at the fixed-point offset
(16, -8) * tri(t) / 256 px, tri = triangle wave of period 128 frames (a slow camera sway of 1/16 px per
frame; the reference's non-converging LK reports about 32x that).  The field is

    pix = (4*B + 3*V16 + V4) >> 3

with  B   = random two-level blocks on a 32 px grid, AREA-sampled (box filter over the pixel footprint) so
            that sub-pixel motion changes edge pixels smoothly: strong corners at block corners,
      V16 = bilinear value noise on a 16 px lattice, V4 on a 4 px lattice (weak texture everywhere),
all lattice values being bytes of a 32-bit integer hash.  No floating point anywhere.
"""
import numpy as np

_M = np.uint32(0xFFFFFFFF)


def _hash32(a):
    a = a.astype(np.uint32, copy=True)
    a ^= a >> np.uint32(16)
    a *= np.uint32(0x7FEB352D)
    a ^= a >> np.uint32(15)
    a *= np.uint32(0x846CA68B)
    a ^= a >> np.uint32(16)
    return a


def _lattice(seed, octave, ix, iy):
    s = _hash32(np.array([np.uint32((seed * 4 + octave) & 0xFFFFFFFF)], np.uint32))[0]
    h = _hash32(iy.astype(np.uint32) + s)
    h = _hash32(ix.astype(np.uint32) + h)
    return (h >> np.uint32(24)).astype(np.uint32)


def _value_noise(seed, octave, X, Y, log2cell):
    sh = np.uint32(log2cell + 8)
    fs = np.uint32(log2cell)
    ix, iy = X >> sh, Y >> sh
    fx = (X >> fs) & np.uint32(255)
    fy = (Y >> fs) & np.uint32(255)
    l00 = _lattice(seed, octave, ix, iy)
    l10 = _lattice(seed, octave, ix + np.uint32(1), iy)
    l01 = _lattice(seed, octave, ix, iy + np.uint32(1))
    l11 = _lattice(seed, octave, ix + np.uint32(1), iy + np.uint32(1))
    top = l00 * (np.uint32(256) - fx) + l10 * fx
    bot = l01 * (np.uint32(256) - fx) + l11 * fx
    return (top * (np.uint32(256) - fy) + bot * fy) >> np.uint32(16)


def tri(t):
    m = t % 128
    return m if m < 64 else 128 - m


def _blocks(seed, X, Y):
    u = np.uint32
    ix, iy = X >> u(13), Y >> u(13)
    fx, fy = X & u(8191), Y & u(8191)
    w1x = np.where(fx > u(8192 - 256), fx - u(8192 - 256), u(0))
    w1y = np.where(fy > u(8192 - 256), fy - u(8192 - 256), u(0))
    w0x, w0y = u(256) - w1x, u(256) - w1y

    def b(i, j):
        return np.where((_lattice(seed, 0, i, j) & u(3)) == 0, u(255), u(0))

    return (b(ix, iy) * w0x * w0y + b(ix + u(1), iy) * w1x * w0y + b(ix, iy + u(1)) * w0x * w1y
            + b(ix + u(1), iy + u(1)) * w1x * w1y) >> u(16)


def frame(seed, t, w, h):
    """uint8 [h, w] frame t of sequence `seed`."""
    u = np.uint32
    with np.errstate(over="ignore"):
        x = np.arange(w, dtype=np.uint32)[None, :]
        y = np.arange(h, dtype=np.uint32)[:, None]
        X = np.broadcast_to(x * u(256) + u((1 << 20) + 16 * tri(t)), (h, w))
        Y = np.broadcast_to(y * u(256) + u((1 << 20) - 8 * tri(t)), (h, w))
        B = _blocks(seed, X, Y)
        v16 = _value_noise(seed, 1, X, Y, 4)
        v4 = _value_noise(seed, 2, X, Y, 2)
        pix = (u(4) * B + u(3) * v16 + v4) >> u(3)
    return pix.astype(np.uint8)


def sequence(seed, t0, nframes, w, h):
    return np.stack([frame(seed, t0 + k, w, h) for k in range(nframes)])


def kat_image(w=640, h=480):
    """The known-answer image of SURVEY.md §8c(3) (LCG noise over a sheared checker)."""
    s = np.uint64(1)
    n = np.zeros(w * h, np.uint32)
    a, c, m = 1664525, 1013904223, 0xFFFFFFFF
    v = 1
    for i in range(w * h):
        v = (v * a + c) & m
        n[i] = (v >> 24) & 31
    y, x = np.mgrid[0:h, 0:w]
    val = (((x * 7 + y * 13) >> 2) & 63) + (((x // 16 + y // 16) & 1) * 96) + n.reshape(h, w)
    return (val & 255).astype(np.uint8)


def fnv1a64(buf):
    h = 0xCBF29CE484222325
    for b in bytes(buf):
        h = ((h ^ b) * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
    return h
