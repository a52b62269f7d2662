"""ctypes loader for the two CPU checkers — TEST INFRASTRUCTURE, not product code.

`port()`  -> oracle/libsfmoracle.so  (our restatement, prefix orc_; sfm_oracle.cpp)
`ref()`   -> oracle/_ref/libsfmref.so (unmodified reference TU compiled where it lies, prefix ref_;
             ref_harness.cpp).  May be absent (returns None) when neither the prebuilt file nor
             /root/reference is available.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
Every wrapper mirrors one reference entry point (cpp/src/templering_sfm.cpp, see the two .cpp headers).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")


def build(force=False):
    """(Re)build the checkers with oracle/Makefile.  Building the checker is not using it."""
    args = ["make", "-C", _HERE, "-s"] + (["-B"] if force else [])
    subprocess.run(args, check=True)


class CpuFrontEnd:
    """Numpy-facing view of one checker library (prefix 'orc' or 'ref')."""

    def __init__(self, path, prefix):
        self.lib = C.CDLL(path)
        self.prefix = prefix
        self.kind = "reference" if prefix == "ref" else "port"
        self._sig()

    def _f(self, name):
        return getattr(self.lib, f"{self.prefix}_{name}")

    def _sig(self):
        i, d, vp = C.c_int, C.c_double, C.c_void_p
        S = {
            "build_pyr": (i, [_u8p, i, i, i, _u8p]),
            "shi_tomasi": (i, [_u8p, i, i, i, d, i, _f64p, i]),
            "klt_track": (i, [_u8p, _u8p, i, i, i, i, i, _f64p, i, _f64p, _f64p]),
            "tracker_create": (vp, [i, i, d, i, i, i, i, d]),
            "tracker_destroy": (None, [vp]),
            "tracker_reset": (None, [vp, _u8p, i, i]),
            "tracker_step": (i, [vp, _u8p, i, i, _f64p, _f64p, _i32p, i]),
            "tracker_tracks": (i, [vp, _f64p, _i32p, i]),
            "norm_points": (i, [_f64p, _f64p, i, _f64p]),
            "sampson": (d, [_f64p, d, d, d, d]),
            "rng_draws": (i, [i, i, _i32p]),
            "ransac_hypotheses": (i, [_f64p, _f64p, i, i, _f64p, C.c_void_p]),
            "ransac_score": (i, [_f64p, _f64p, i, _f64p, i, d, _i32p, C.POINTER(i), _i32p, C.POINTER(i)]),
            "find_E_ransac": (i, [_f64p, _f64p, _f64p, i, i, d, i, _f64p, _f64p, _i32p, C.POINTER(i)]),
            "pair_frontend": (i, [_u8p, _u8p, i, i, i, d, i, i, i, i, d, _f64p, _f64p, C.POINTER(i)]),
            "pair_frontend_mt": (C.c_long, [_u8p, i, i, i, i, d, i, i, i, i, d, i, C.POINTER(C.c_long)]),
            "ransac_score_mt": (i, [_f64p, _f64p, i, _f64p, i, d, _i32p, i]),
            "two_view_mt": (C.c_long, [_u8p, i, i, i, i, d, i, i, i, i, d, vp, i, d, i, i, i, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
            "global_desc32": (i, [_u8p, i, i, _f32p]),
            "triangulate_dlt": (i, [_f64p, _f64p, i, _i32p, _i32p, _f64p, _f64p, i, _f64p]),
            "desc_search": (i, [_f32p, i, _f32p, _f32p, C.POINTER(i), C.POINTER(C.c_float)]),
        }
        if self.prefix == "orc":
            S.update({
                "score_map": (i, [_u8p, i, i, _f64p]),
                "candidates": (i, [_u8p, i, i, d, i, _i32p, _f64p, i, C.POINTER(d)]),
                "sort_perm_desc": (i, [_f64p, i, _i32p]),
                "klt_track_count": (i, [_u8p, _u8p, i, i, i, i, i, _f64p, i, _f64p, _f64p, _i32p]),
                "synth_frames": (i, [C.c_uint32, i, i, i, i, _u8p, i]),
            })
        for name, (res, args) in S.items():
            fn = self._f(name)
            fn.restype = res
            fn.argtypes = args

    # ---- pyramid (:224-232) ---------------------------------------------------------------
    def build_pyr(self, img, levels):
        h, w = img.shape
        shapes, ww, hh = [], w, h
        for _ in range(1, levels):
            ww, hh = ww // 2, hh // 2
            shapes.append((hh, ww))
        out = np.zeros(max(1, sum(a * b for a, b in shapes)), np.uint8)
        self._f("build_pyr")(np.ascontiguousarray(img), w, h, levels, out)
        lv, off = [np.ascontiguousarray(img)], 0
        for a, b in shapes:
            lv.append(out[off:off + a * b].reshape(a, b).copy())
            off += a * b
        return lv

    # ---- corners (:237-302) -----------------------------------------------------------------
    def shi_tomasi(self, img, max_corners, quality=0.01, min_dist=8):
        h, w = img.shape
        cap = max(1, min(max_corners, w * h))
        xy = np.zeros((cap, 2), np.float64)
        n = self._f("shi_tomasi")(np.ascontiguousarray(img), w, h, max_corners, quality, min_dist, xy, cap)
        return xy[:n].copy()

    def score_map(self, img):
        h, w = img.shape
        s = np.zeros((h, w), np.float64)
        self._f("score_map")(np.ascontiguousarray(img), w, h, s)
        return s

    def candidates(self, img, quality=0.01, sorted_=False):
        h, w = img.shape
        cap = w * h
        xy = np.zeros((cap, 2), np.int32)
        s = np.zeros(cap, np.float64)
        mx = C.c_double(0)
        n = self._f("candidates")(np.ascontiguousarray(img), w, h, quality, int(sorted_), xy, s, cap, C.byref(mx))
        return xy[:n].copy(), s[:n].copy(), mx.value

    def synth_frames(self, seed, t0, nframes, w, h, threads=8):
        out = np.zeros((nframes, h, w), np.uint8)
        self._f("synth_frames")(seed & 0xFFFFFFFF, t0, nframes, w, h, out.reshape(-1), threads)
        return out

    def sort_perm_desc(self, keys):
        keys = np.ascontiguousarray(keys, np.float64)
        perm = np.zeros(max(1, len(keys)), np.int32)
        self._f("sort_perm_desc")(keys, len(keys), perm)
        return perm[:len(keys)]

    # ---- KLT (:396-460) -------------------------------------------------------------------------
    def klt_track(self, im0, im1, p0, levels=3, radius=5, iters=10, count=False):
        h, w = im0.shape
        p0 = np.ascontiguousarray(p0, np.float64).reshape(-1, 2)
        n = len(p0)
        p1 = np.zeros((max(n, 1), 2))
        pb = np.zeros((max(n, 1), 2))
        if count:
            nit = np.zeros(max(n, 1), np.int32)
            self._f("klt_track_count")(np.ascontiguousarray(im0), np.ascontiguousarray(im1), w, h, levels, radius,
                                       iters, p0 if n else np.zeros((1, 2)), n, p1, pb, nit)
            return p1[:n], pb[:n], nit[:n]
        self._f("klt_track")(np.ascontiguousarray(im0), np.ascontiguousarray(im1), w, h, levels, radius, iters,
                             p0 if n else np.zeros((1, 2)), n, p1, pb)
        return p1[:n], pb[:n]

    def tracker(self, max_tracks=2200, min_tracks=900, quality=0.01, min_distance=8, levels=3, radius=5, iters=10,
                fb=1.0):
        return _Tracker(self, max_tracks, min_tracks, quality, min_distance, levels, radius, iters, fb)

    # ---- two-view geometry (:471-501, :609-761) -----------------------------------------------------
    def norm_points(self, K, p):
        p = np.ascontiguousarray(p, np.float64).reshape(-1, 2)
        out = np.zeros_like(p)
        self._f("norm_points")(np.ascontiguousarray(K, np.float64).reshape(9), p, len(p), out)
        return out

    def sampson(self, E, x, xp):
        return self._f("sampson")(np.ascontiguousarray(E, np.float64).reshape(9), x[0], x[1], xp[0], xp[1])

    def rng_draws(self, n, count):
        out = np.zeros(count, np.int32)
        self._f("rng_draws")(n, count, out)
        return out

    def ransac_hypotheses(self, xi, xj, iters):
        xi = np.ascontiguousarray(xi, np.float64).reshape(-1, 2)
        xj = np.ascontiguousarray(xj, np.float64).reshape(-1, 2)
        E = np.zeros((iters, 9))
        idx = np.zeros((iters, 8), np.int32)
        self._f("ransac_hypotheses")(xi, xj, len(xi), iters, E, idx.ctypes.data_as(C.c_void_p))
        return E, idx

    def ransac_score(self, xi, xj, E, thr):
        xi = np.ascontiguousarray(xi, np.float64).reshape(-1, 2)
        xj = np.ascontiguousarray(xj, np.float64).reshape(-1, 2)
        E = np.ascontiguousarray(E, np.float64).reshape(-1, 9)
        H, n = len(E), len(xi)
        counts = np.zeros(max(H, 1), np.int32)
        inl = np.zeros(max(n, 1), np.int32)
        bh, bn = C.c_int(-1), C.c_int(0)
        self._f("ransac_score")(xi, xj, n, E, H, thr, counts, C.byref(bh), inl, C.byref(bn))
        return counts[:H], bh.value, inl[:bn.value].copy()

    def ransac_score_mt(self, xi, xj, E, thr, threads):
        xi = np.ascontiguousarray(xi, np.float64).reshape(-1, 2)
        xj = np.ascontiguousarray(xj, np.float64).reshape(-1, 2)
        E = np.ascontiguousarray(E, np.float64).reshape(-1, 9)
        counts = np.zeros(len(E), np.int32)
        self._f("ransac_score_mt")(xi, xj, len(xi), E, len(E), thr, counts, threads)
        return counts

    # ---- batched triangulate_dlt (:1477-1516) ------------------------------------------------------------------
    def triangulate_dlt(self, K, poses, ia, ib, ui, uj):
        poses = np.ascontiguousarray(poses, np.float64).reshape(-1, 12)
        ia, ib = np.ascontiguousarray(ia, np.int32), np.ascontiguousarray(ib, np.int32)
        ui = np.ascontiguousarray(ui, np.float64).reshape(-1, 2)
        uj = np.ascontiguousarray(uj, np.float64).reshape(-1, 2)
        X = np.zeros((max(len(ui), 1), 3))
        if len(ui):
            self._f("triangulate_dlt")(np.ascontiguousarray(K, np.float64).reshape(9), poses, len(poses), ia, ib, ui, uj, len(ui), X)
        return X[:len(ui)]

    # ---- loop-closure descriptor (:1100-1129) and candidate search (:1823-1831) -----------------------------
    def global_desc32(self, img):
        img = np.ascontiguousarray(img, np.uint8)
        out = np.zeros(1024, np.float32)
        if self._f("global_desc32")(img, img.shape[1], img.shape[0], out) != 1024:
            raise ValueError("degenerate image")
        return out

    def desc_search(self, descs, query, n_search=None):
        descs = np.ascontiguousarray(descs, np.float32).reshape(-1, 1024)
        query = np.ascontiguousarray(query, np.float32).reshape(1024)
        n = len(descs) if n_search is None else n_search
        scores = np.zeros(max(n, 1), np.float32)
        bid, bs = C.c_int(-1), C.c_float(0)
        self._f("desc_search")(descs if len(descs) else np.zeros((1, 1024), np.float32), n, query, scores, C.byref(bid), C.byref(bs))
        return bid.value, np.float32(bs.value), scores[:n]

    def find_E_ransac(self, K, pi, pj, iters, thr, min_inliers):
        pi = np.ascontiguousarray(pi, np.float64).reshape(-1, 2)
        pj = np.ascontiguousarray(pj, np.float64).reshape(-1, 2)
        n = len(pi)
        R, t = np.zeros(9), np.zeros(3)
        inl = np.zeros(max(n, 1), np.int32)
        k = C.c_int(0)
        ok = self._f("find_E_ransac")(np.ascontiguousarray(K, np.float64).reshape(9), pi if n else np.zeros((1, 2)),
                                      pj if n else np.zeros((1, 2)), n, iters, thr, min_inliers, R, t, inl,
                                      C.byref(k))
        if ok != 1:
            return None
        return R.reshape(3, 3), t, inl[:k.value].copy()

    # ---- stateless two-view front end (:1836-1857) ---------------------------------------------------------
    def pair_frontend(self, im0, im1, max_corners, quality=0.01, min_dist=8, levels=3, radius=5, iters=10, fb=1.0):
        h, w = im0.shape
        li = np.zeros((max(1, max_corners), 2))
        lj = np.zeros((max(1, max_corners), 2))
        nc = C.c_int(0)
        k = self._f("pair_frontend")(np.ascontiguousarray(im0), np.ascontiguousarray(im1), w, h, max_corners, quality,
                                     min_dist, levels, radius, iters, fb, li, lj, C.byref(nc))
        return li[:k].copy(), lj[:k].copy(), nc.value

    def two_view_mt(self, frames, max_corners, threads, K=None, rs_iters=4000, rs_thr=2e-3, rs_min_inliers=80, min_points=120,
                    quality=0.01, min_dist=8, levels=3, radius=5, iters=10, fb=1.0, detail=True):
        """The whole two-view unit (:1836-1857) for every consecutive pair of `frames` on `threads` host threads: front end
        plus `if (li.size() >= min_points) find_E_ransac(K, li, lj, rs_iters, rs_thr, rs_min_inliers)` (K None: front end
        only).  Returns (corners processed, dict of per-pair arrays) - the dict is None when detail is False."""
        f, h, w = frames.shape
        P, cap = f - 1, max(1, max_corners)
        out = None
        if detail:
            out = dict(n_corners=np.zeros(P, np.int32), n_kept=np.zeros(P, np.int32), li=np.zeros((P, cap, 2)), lj=np.zeros((P, cap, 2)),
                       status=np.zeros(P, np.int32), n_inl=np.zeros(P, np.int32), inliers=np.zeros((P, cap), np.int32),
                       R=np.zeros((P, 9)), t=np.zeros((P, 3)))
        g = (lambda k: _p(out[k])) if detail else (lambda k: None)
        Kc = None if K is None else np.ascontiguousarray(K, np.float64).reshape(9)
        tracks = self._f("two_view_mt")(np.ascontiguousarray(frames).reshape(-1), f, w, h, max_corners, quality, min_dist, levels, radius,
                                        iters, fb, _p(Kc), rs_iters, rs_thr, rs_min_inliers, min_points, threads, g("n_corners"),
                                        g("n_kept"), g("li"), g("lj"), g("status"), g("n_inl"), g("inliers"), g("R"), g("t"))
        return int(tracks), out

    def pair_frontend_mt(self, frames, max_corners, threads, quality=0.01, min_dist=8, levels=3, radius=5, iters=10,
                         fb=1.0):
        f, h, w = frames.shape
        kept = C.c_long(0)
        tracks = self._f("pair_frontend_mt")(np.ascontiguousarray(frames).reshape(-1), f, w, h, max_corners, quality,
                                             min_dist, levels, radius, iters, fb, threads, C.byref(kept))
        return int(tracks), int(kept.value)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class _Tracker:
    """KLTTracker (:323-466) — reset / step / tracks."""

    def __init__(self, fe, max_tracks, min_tracks, quality, min_distance, levels, radius, iters, fb):
        self.fe = fe
        self.cap = max(1, max_tracks)
        self.h = fe._f("tracker_create")(max_tracks, min_tracks, quality, min_distance, levels, radius, iters, fb)

    def reset(self, img):
        self.fe._f("tracker_reset")(self.h, np.ascontiguousarray(img), img.shape[1], img.shape[0])

    def step(self, img):
        cap = max(self.cap, img.size)
        prev = np.zeros((cap, 2))
        cur = np.zeros((cap, 2))
        ids = np.zeros(cap, np.int32)
        n = self.fe._f("tracker_step")(self.h, np.ascontiguousarray(img), img.shape[1], img.shape[0], prev, cur, ids,
                                       cap)
        return prev[:n].copy(), cur[:n].copy(), ids[:n].copy()

    def tracks(self):
        cap = self.cap * 4 + 16
        xy = np.zeros((cap, 2))
        ids = np.zeros(cap, np.int32)
        n = self.fe._f("tracker_tracks")(self.h, xy, ids, cap)
        return xy[:n].copy(), ids[:n].copy()

    def __del__(self):
        try:
            self.fe._f("tracker_destroy")(self.h)
        except Exception:
            pass


_cache = {}


def port():
    if "port" not in _cache:
        p = os.path.join(_HERE, "libsfmoracle.so")
        if not os.path.exists(p):
            build()
        _cache["port"] = CpuFrontEnd(p, "orc")
    return _cache["port"]


def ref():
    """The compiled reference, or None when it cannot be had on this machine."""
    if "ref" not in _cache:
        p = os.path.join(_HERE, "_ref", "libsfmref.so")
        if not os.path.exists(p) and os.path.exists("/root/reference/cpp/src/templering_sfm.cpp"):
            build()
        _cache["ref"] = CpuFrontEnd(p, "ref") if os.path.exists(p) else None
    return _cache["ref"]


def best():
    """(checker, kind): the compiled reference when present, else the port."""
    r = ref()
    return (r, "reference") if r is not None else (port(), "port")
