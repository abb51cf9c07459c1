"""TEST INFRASTRUCTURE ONLY -- ctypes front end of the CPU oracle (oracle/liboracle.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl reference`` legs may
import this module; the product (diasss_b200/) never does.  See orb_oracle.cpp / match_oracle.cpp
for what is restated (reference file:line) and for the pinning statement.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                     ("octave", "<i4"), ("class_id", "<i4")])  # cv::KeyPoint, 28 B
assert KP_DTYPE.itemsize == 28


class _Frame(C.Structure):
    _fields_ = [("img_id", C.c_int), ("rows", C.c_int), ("cols", C.c_int), ("n", C.c_int),
                ("kps", C.c_void_p), ("desc", C.c_void_p), ("geo_x", C.c_void_p), ("geo_y", C.c_void_p)]


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("orb_oracle.cpp", "match_oracle.cpp", "oracle_capi.h", "orb_pattern.inc")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.orc_fast_atan2.restype = C.c_float
        _LIB.orc_fast_atan2.argtypes = [C.c_float, C.c_float]
        _LIB.orc_extractor_create.restype = C.c_void_p
        _LIB.orc_extractor_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
        _LIB.orc_compute_intersection.restype = C.c_float
        _LIB.orc_mean.restype = C.c_double
    return _LIB


def _p(a):
    return C.c_void_p(a.ctypes.data) if a is not None else C.c_void_p(0)


def _u8img(img):
    img = np.ascontiguousarray(img, dtype=np.uint8)
    assert img.ndim == 2
    return img


# ------------------------------------------------------------------ OpenCV primitives
def resize_linear(src, drows, dcols):
    src = _u8img(src)
    dst = np.empty((drows, dcols), np.uint8)
    lib().orc_resize_linear_u8(_p(src), src.shape[0], src.shape[1], src.strides[0], _p(dst), drows, dcols, dcols)
    return dst


def fast9_16(img, threshold):
    """cv::FAST(img, kps, threshold, true): returns int32 array k x 3 = (x, y, score), row-major order."""
    img = _u8img(img)
    cap = max(16, img.size // 4 + 16)
    out = np.empty((cap, 3), np.int32)
    n = lib().orc_fast9_16(_p(img), img.shape[0], img.shape[1], img.strides[0], int(threshold), _p(out), cap)
    assert n <= cap
    return out[:n].copy()


def gaussian13(img):
    img = _u8img(img)
    dst = np.empty_like(img)
    lib().orc_gaussian13_s2(_p(img), img.shape[0], img.shape[1], img.strides[0], _p(dst), dst.strides[0])
    return dst


def fast_atan2(y, x):
    return float(lib().orc_fast_atan2(C.c_float(y), C.c_float(x)))


def pattern():
    out = np.empty(1024, np.int8)
    lib().orc_pattern(_p(out))
    return out


def rng_draws(n):
    out = np.empty(n, np.uint32)
    lib().orc_rng_draws(_p(out), n)
    return out


def distribute(xys, minX, maxX, minY, maxY, N):
    xys = np.ascontiguousarray(xys, np.int32).reshape(-1, 3)
    out = np.empty((N + 8 + 4 * 64, 3), np.int32)
    n = lib().orc_distribute(_p(xys), len(xys), minX, maxX, minY, maxY, N, _p(out))
    return out[:n].copy()


# ------------------------------------------------------------------ ORBextractor
class Extractor:
    """ORB_SLAM2::ORBextractor in ORB mode (thirdparty/ORBextractor.h:51-61)."""

    def __init__(self, nfeatures=2000, scale_factor=1.2, nlevels=6, ini_th=12, min_th=7):
        self.nfeatures, self.nlevels = nfeatures, nlevels
        self.h = C.c_void_p(lib().orc_extractor_create(nfeatures, C.c_float(scale_factor), nlevels, ini_th, min_th))
        self.scale = np.empty(nlevels, np.float32)
        self.inv_scale = np.empty(nlevels, np.float32)
        self.features_per_level = np.empty(nlevels, np.int32)
        self.umax = np.empty(16, np.int32)
        lib().orc_extractor_tables(self.h, _p(self.scale), _p(self.inv_scale), _p(self.features_per_level), _p(self.umax))
        self._shape = None

    def __del__(self):
        try:
            lib().orc_extractor_destroy(self.h)
        except Exception:
            pass

    def __call__(self, img):
        img = _u8img(img)
        cap = self.nfeatures + 4 * self.nlevels + 64
        kps = np.empty(cap, KP_DTYPE)
        desc = np.empty((cap, 32), np.uint8)
        n = lib().orc_extractor_run(self.h, _p(img), img.shape[0], img.shape[1], img.strides[0], _p(kps), _p(desc), cap)
        assert n <= cap
        self._shape = img.shape
        return kps[:n].copy(), desc[:n].copy()

    def level_size(self, rows, cols, level):
        r, c = C.c_int(), C.c_int()
        lib().orc_extractor_level_size(self.h, rows, cols, level, C.byref(r), C.byref(c))
        return r.value, c.value

    def level_image(self, level):
        r, c = self.level_size(self._shape[0], self._shape[1], level)
        out = np.empty((r, c), np.uint8)
        lib().orc_extractor_level_image(self.h, level, _p(out))
        return out

    def candidates(self, level):
        n = lib().orc_extractor_candidates(self.h, level, C.c_void_p(0), 0)
        out = np.empty((max(n, 1), 3), np.int32)
        lib().orc_extractor_candidates(self.h, level, _p(out), n)
        return out[:n]

    def level_keys(self, level):
        cap = self.nfeatures + 64
        out = np.empty(cap, KP_DTYPE)
        n = lib().orc_extractor_level_keys(self.h, level, _p(out), cap)
        return out[:n].copy()


# ------------------------------------------------------------------ Frame glue
def mask_filter(kps, desc, mask):
    mask = _u8img(mask)
    n = len(kps)
    kps = np.ascontiguousarray(kps, KP_DTYPE)
    desc = np.ascontiguousarray(desc, np.uint8)
    ok, od, oi = np.empty(max(n, 1), KP_DTYPE), np.empty((max(n, 1), 32), np.uint8), np.empty(max(n, 1), np.int32)
    m = lib().orc_mask_filter(_p(kps), _p(desc), n, _p(mask), mask.strides[0], _p(ok), _p(od), _p(oi))
    return ok[:m].copy(), od[:m].copy(), oi[:m].copy()


def geo_img(rows, cols, pose6, g_range):
    pose6 = np.ascontiguousarray(pose6, np.float64).reshape(rows, 6)
    g_range = np.ascontiguousarray(g_range, np.float64)
    assert len(g_range) >= cols - cols // 2 + 1  # SURVEY Appendix B4 (cols/2+1 for even cols)
    gx, gy = np.empty((rows, cols), np.float64), np.empty((rows, cols), np.float64)
    lib().orc_geo_img(rows, cols, _p(pose6), _p(g_range), len(g_range), _p(gx), _p(gy))
    return gx, gy


def mean(raw, order=0):
    """cv::mean of a CV_64F image.  order 0: sequential sum; order 1: the 32-lane order the CUDA library defines."""
    raw = np.ascontiguousarray(raw, np.float64)
    return float(lib().orc_mean(_p(raw), raw.shape[0], raw.shape[1], int(order)))


def normalize_sss(raw, order=0):
    """Frame::GetNormalizeSSS (frame.cpp:57-81)."""
    raw = np.ascontiguousarray(raw, np.float64)
    out = np.empty(raw.shape, np.uint8)
    lib().orc_normalize_sss_m(_p(raw), raw.shape[0], raw.shape[1], C.c_double(mean(raw, order)), _p(out))
    return out


def filtered_mask(raw, order=0):
    """Frame::GetFilteredMask (frame.cpp:83-124, Appendix B5)."""
    raw = np.ascontiguousarray(raw, np.float64)
    out = np.empty(raw.shape, np.uint8)
    lib().orc_filtered_mask_m(_p(raw), raw.shape[0], raw.shape[1], C.c_double(mean(raw, order)), _p(out))
    return out


def compute_intersection(geo_s, geo_t):
    sx, sy = (np.ascontiguousarray(a, np.float64) for a in geo_s)
    tx, ty = (np.ascontiguousarray(a, np.float64) for a in geo_t)
    return float(lib().orc_compute_intersection(_p(sx), _p(sy), sx.size, _p(tx), _p(ty), tx.size))


# ------------------------------------------------------------------ FEAmatcher
class Frame:
    """The Diasss::Frame fields the matcher reads (frame.h:30-46)."""

    def __init__(self, img_id, rows, cols, kps, desc, geo_x, geo_y):
        self.img_id, self.rows, self.cols = int(img_id), int(rows), int(cols)
        self.kps = np.ascontiguousarray(kps, KP_DTYPE)
        self.desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        self.geo_x = np.ascontiguousarray(geo_x, np.float64)
        self.geo_y = np.ascontiguousarray(geo_y, np.float64)
        assert self.geo_x.shape == (rows, cols) and self.geo_y.shape == (rows, cols)

    def c(self):
        return _Frame(self.img_id, self.rows, self.cols, len(self.kps), self.kps.ctypes.data, self.desc.ctypes.data,
                      self.geo_x.ctypes.data, self.geo_y.ctypes.data)


def descriptor_distance(a, b):
    a, b = np.ascontiguousarray(a, np.uint8), np.ascontiguousarray(b, np.uint8)
    return int(lib().orc_descriptor_distance(_p(a), _p(b)))


def geo_nn_search(f, ref):
    """FEAmatcher::GeoNearNeighSearch(f -> ref), ORB branch.  Returns a dict of per-keypoint arrays + scc."""
    n = len(f.kps)
    z = lambda: np.empty(max(n, 1), np.int32)
    corres, pre, best, sec, ncand = z(), z(), z(), z(), z()
    sc, sm, ns = np.empty(1000, np.int32), np.empty(1000, np.float64), C.c_int()
    cf, cr = f.c(), ref.c()
    lib().orc_geo_nn_search(C.byref(cf), C.byref(cr), _p(corres), _p(pre), _p(best), _p(sec), _p(ncand), _p(sc), _p(sm),
                            1000, C.byref(ns))
    return dict(corres=corres[:n], pre=pre[:n], best=best[:n], sec=sec[:n], ncand=ncand[:n],
                scc=list(zip(sc[:ns.value].tolist(), sm[:ns.value].tolist())))


def robust_matching(s, t):
    """FEAmatcher::RobustMatching.  Returns (rows6 [K,6] f64, src_idx, tgt_idx, corres1, corres2)."""
    cap = len(s.kps) + len(t.kps) + 1
    rows6 = np.empty((cap, 6), np.float64)
    si, ti = np.empty(cap, np.int32), np.empty(cap, np.int32)
    c1, c2 = np.empty(max(len(s.kps), 1), np.int32), np.empty(max(len(t.kps), 1), np.int32)
    cs, ct = s.c(), t.c()
    k = lib().orc_robust_matching(C.byref(cs), C.byref(ct), _p(rows6), _p(si), _p(ti), cap, _p(c1), _p(c2))
    return rows6[:k].copy(), si[:k].copy(), ti[:k].copy(), c1[:len(s.kps)].copy(), c2[:len(t.kps)].copy()


def get_kps_pairs(rows6, id_t, alt_s, gra_s, alt_t, gra_t):
    """Optimizer::GetKpsPairs, USE_ANNO = 0 (optimizer.cpp:575-639) on K x 6 corres_kps rows -> [n, 7] float64."""
    rows6 = np.ascontiguousarray(rows6, np.float64).reshape(-1, 6)
    alt_s, gra_s, alt_t, gra_t = (np.ascontiguousarray(a, np.float64) for a in (alt_s, gra_s, alt_t, gra_t))
    out = np.empty((max(len(rows6), 1), 7), np.float64)
    n = lib().orc_get_kps_pairs(_p(rows6), len(rows6), int(id_t), _p(alt_s), _p(gra_s), len(gra_s), _p(alt_t), _p(gra_t), len(gra_t), _p(out))
    return out[:n].copy()
