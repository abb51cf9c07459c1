"""TEST INFRASTRUCTURE ONLY -- ctypes front end of oracle/_ref/: the reference's OWN hot-path code
(thirdparty/ORBextractor.cpp, src/core/FEAmatcher.cpp, src/core/frame.cpp, Util::ComputeIntersection) compiled
unmodified from /root/reference by oracle/build_ref.sh against the OpenCV stand-in of oracle/ref_stub/.

Same surface as oracle/oracle.py so that a test can run both and compare.  Only tests/ and bench.py's reference /
cpu_baseline legs import this module.  The libraries are prebuilt artefacts (git-ignored, shipped to the GPU box);
`build()` rebuilds them only where /root/reference exists.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from .oracle import KP_DTYPE, Frame, _Frame, _p, _u8img  # noqa: F401  (Frame is re-exported)

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS = {}


def build(force=False):
    so = os.path.join(_HERE, "_ref", "libdiasss_ref.so")
    ref_root = os.environ.get("DSX_REFERENCE_ROOT", "/root/reference")
    if os.path.isdir(ref_root):
        srcs = [os.path.join(_HERE, "build_ref.sh")] + [os.path.join(_HERE, "ref_stub", f) for f in
                                                        ("ref_capi.cpp", "ref_cv_impl.cpp", "util.h", "opencv2/opencv.hpp")]
        srcs += [os.path.join(_HERE, f) for f in ("orb_oracle.cpp", "match_oracle.cpp")]
        if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
            subprocess.check_call(["bash", os.path.join(_HERE, "build_ref.sh")])
    return so


def available():
    return os.path.exists(build())


def lib(strict=False):
    """strict=False: reference + S1 S2 B1 B3 (runs everywhere); strict=True: reference + S1 S2 only."""
    if strict not in _LIBS:
        build()
        name = "libdiasss_ref_strict.so" if strict else "libdiasss_ref.so"
        L = C.CDLL(os.path.join(_HERE, "_ref", name))
        L.ref_build_info.restype = C.c_char_p
        L.ref_extractor_create.restype = C.c_void_p
        L.ref_extractor_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
        L.ref_compute_intersection.restype = C.c_float
        L.ref_frame_intersection.restype = C.c_float
        L.ref_frame_create.restype = C.c_void_p
        L.ref_survey_create.restype = C.c_void_p
        L.ref_survey_frame.restype = C.c_void_p
        L.ref_survey_match.argtypes = [C.c_void_p, C.c_float, C.c_int]
        L.ref_consistent_check.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_int,
                                           C.c_double, C.c_void_p, C.c_void_p, C.c_int]
        _LIBS[strict] = L
    return _LIBS[strict]


def set_modes(heap_monotone=True, libm_a5=True, mean_order=1, strict=False):
    """heap_monotone: list nodes get addresses in creation order (the oracle's tie-break B2); False = glibc heap.
    libm_a5: cosf/sinf of the descriptor rotation = double evaluation rounded (oracle definition A5); False = libm.
    mean_order: 0 sequential cv::mean, 1 the 32-lane order the CUDA library documents."""
    L = lib(strict)
    L.ref_set_heap_mode(int(heap_monotone))
    L.ref_set_libm_mode(int(libm_a5))
    L.ref_set_mean_order(int(mean_order))


class Extractor:
    """ORB_SLAM2::ORBextractor (thirdparty/ORBextractor.h:51-61): the reference's class itself."""

    def __init__(self, nfeatures=2000, scale_factor=1.2, nlevels=6, ini_th=12, min_th=7, strict=False):
        self.L = lib(strict)
        self.nfeatures, self.nlevels = nfeatures, nlevels
        self.h = C.c_void_p(self.L.ref_extractor_create(nfeatures, C.c_float(scale_factor), nlevels, ini_th, min_th))
        self.scale = np.empty(nlevels, np.float32)
        self.inv_scale = np.empty(nlevels, np.float32)
        self.features_per_level = np.empty(nlevels, np.int32)
        self.umax = np.empty(16, np.int32)
        self.L.ref_extractor_tables(self.h, _p(self.scale), _p(self.inv_scale), _p(self.features_per_level), _p(self.umax))

    def __del__(self):
        try:
            self.L.ref_extractor_destroy(self.h)
        except Exception:
            pass

    def __call__(self, img):
        img = _u8img(img)
        cap = self.nfeatures + 4 * self.nlevels + 64
        kps = np.empty(cap, KP_DTYPE)
        desc = np.empty((cap, 32), np.uint8)
        n = self.L.ref_extractor_run(self.h, _p(img), img.shape[0], img.shape[1], img.strides[0], _p(kps), _p(desc), cap)
        assert n <= cap
        return kps[:n].copy(), desc[:n].copy()

    def level_image(self, level):
        r, c = C.c_int(), C.c_int()
        self.L.ref_extractor_level_size(self.h, level, C.byref(r), C.byref(c))
        out = np.empty((r.value, c.value), np.uint8)
        self.L.ref_extractor_level_image(self.h, level, _p(out))
        return out

    def candidates(self, level):
        n = self.L.ref_extractor_candidates(self.h, level, C.c_void_p(0), 0)
        out = np.empty((max(n, 1), 3), np.int32)
        self.L.ref_extractor_candidates(self.h, level, _p(out), n)
        return out[:n]

    def distribute(self, xys, minX, maxX, minY, maxY, N):
        xys = np.ascontiguousarray(xys, np.int32).reshape(-1, 3)
        cap = N + 8 + 4 * 64
        out = np.empty((cap, 3), np.int32)
        n = self.L.ref_distribute(self.h, _p(xys), len(xys), minX, maxX, minY, maxY, N, _p(out), cap)
        assert n <= cap
        return out[:n].copy()


def descriptor_distance(a, b, strict=False):
    a, b = np.ascontiguousarray(a, np.uint8), np.ascontiguousarray(b, np.uint8)
    return int(lib(strict).ref_descriptor_distance(_p(a), _p(b)))


def geo_nn_search(f, ref, strict=False):
    """FEAmatcher::GeoNearNeighSearch(f -> ref): final CorresID and the scc pushes."""
    n = len(f.kps)
    corres = np.empty(max(n, 1), np.int32)
    sc, sm, ns = np.empty(1000, np.int32), np.empty(1000, np.float64), C.c_int()
    cf, cr = f.c(), ref.c()
    lib(strict).ref_geo_nn_search(C.byref(cf), C.byref(cr), _p(corres), _p(sc), _p(sm), 1000, C.byref(ns))
    return dict(corres=corres[:n], scc=list(zip(sc[:ns.value].tolist(), sm[:ns.value].tolist())))


def robust_matching(s, t, strict=False):
    """FEAmatcher::RobustMatching: (rows appended to Source.corres_kps, rows appended to Target.corres_kps)."""
    cap = len(s.kps) + len(t.kps) + 1
    rows6, mirror6 = np.empty((cap, 6), np.float64), np.empty((cap, 6), np.float64)
    cs, ct = s.c(), t.c()
    k = lib(strict).ref_robust_matching(C.byref(cs), C.byref(ct), _p(rows6), _p(mirror6), cap)
    assert k <= cap
    return rows6[:k].copy(), mirror6[:k].copy()


def compute_intersection(geo_s, geo_t, strict=False):
    sx, sy = (np.ascontiguousarray(a, np.float64) for a in geo_s)
    tx, ty = (np.ascontiguousarray(a, np.float64) for a in geo_t)
    return float(lib(strict).ref_compute_intersection(_p(sx), _p(sy), sx.shape[0], sx.shape[1], _p(tx), _p(ty),
                                                      tx.shape[0], tx.shape[1]))


def get_kps_pairs(rows6, id_s, id_t, alt_s, gra_s, alt_t, gra_t, strict=False):
    """Optimizer::GetKpsPairs(false, kps, ...) -- the reference's own lines (optimizer.cpp:575-639)."""
    rows6 = np.ascontiguousarray(rows6, np.float64).reshape(-1, 6)
    alt_s, gra_s, alt_t, gra_t = (np.ascontiguousarray(a, np.float64) for a in (alt_s, gra_s, alt_t, gra_t))
    out = np.empty((max(len(rows6), 1), 7), np.float64)
    n = lib(strict).ref_get_kps_pairs(_p(rows6), len(rows6), int(id_s), int(id_t), _p(alt_s), len(alt_s), _p(gra_s), len(gra_s),
                                      _p(alt_t), len(alt_t), _p(gra_t), len(gra_t), _p(out))
    return out[:n].copy()


class RefFrame:
    """Diasss::Frame built by its own constructor (frame.cpp:18-55) from the raw f64 waterfall."""

    def __init__(self, img_id, raw, pose6, altitude, g_range, strict=False, handle=None, owner=None):
        self.L = lib(strict)
        self._owner = owner
        if handle is not None:
            self.h = C.c_void_p(handle)
            return
        raw = np.ascontiguousarray(raw, np.float64)
        pose6 = np.ascontiguousarray(pose6, np.float64).reshape(raw.shape[0], 6)
        altitude = np.ascontiguousarray(altitude, np.float64)
        g_range = np.ascontiguousarray(g_range, np.float64)
        self.shape = raw.shape
        self.h = C.c_void_p(self.L.ref_frame_create(int(img_id), _p(raw), raw.shape[0], raw.shape[1], _p(pose6), _p(altitude),
                                                    len(altitude), _p(g_range), len(g_range)))

    def __del__(self):
        try:
            if self._owner is None:
                self.L.ref_frame_destroy(self.h)
        except Exception:
            pass

    def get(self, shape, planes=True):
        rows, cols = shape
        n = self.L.ref_frame_nkps(self.h)
        out = dict(kps=np.empty(n, KP_DTYPE), desc=np.empty((n, 32), np.uint8))
        if planes:
            out.update(norm_img=np.empty((rows, cols), np.uint8), mask=np.empty((rows, cols), np.uint8),
                       geo_x=np.empty((rows, cols), np.float64), geo_y=np.empty((rows, cols), np.float64))
        g = lambda k: _p(out[k]) if k in out else C.c_void_p(0)
        self.L.ref_frame_get(self.h, g("norm_img"), g("mask"), g("geo_x"), g("geo_y"), _p(out["kps"]) if n else C.c_void_p(0),
                             _p(out["desc"]) if n else C.c_void_p(0))
        return out

    def corres_kps(self):
        k = self.L.ref_frame_corres_rows(self.h)
        rows6 = np.empty((k, 6), np.float64)
        if k:
            self.L.ref_frame_corres(self.h, _p(rows6))
        return rows6


class Survey:
    """test_demo's front end (src/diasss2.cpp:82-97) on the reference's classes, frames and pairs over host threads."""

    def __init__(self, threads=None, strict=False):
        self.L = lib(strict)
        self.strict = strict
        self.h = C.c_void_p(self.L.ref_survey_create(int(threads or os.cpu_count() or 1)))
        self._keep = []
        self.shapes = []

    def __del__(self):
        try:
            self.L.ref_survey_destroy(self.h)
        except Exception:
            pass

    def add(self, img_id, raw, pose6, altitude, g_range):
        raw = np.ascontiguousarray(raw, np.float64)
        pose6 = np.ascontiguousarray(pose6, np.float64).reshape(raw.shape[0], 6)
        altitude = np.ascontiguousarray(altitude, np.float64)
        g_range = np.ascontiguousarray(g_range, np.float64)
        self._keep.append((raw, pose6))
        self.shapes.append(raw.shape)
        self.L.ref_survey_add(self.h, int(img_id), _p(raw), raw.shape[0], raw.shape[1], _p(pose6), _p(altitude), len(altitude),
                              _p(g_range), len(g_range))

    def add_prepared(self, img_id, norm_img, mask, pose6, g_range):
        """A frame given as its u8 norm_img / flt_mask: GetGeoImg + DetectFeature run, GetNormalizeSSS / GetFilteredMask don't."""
        norm_img, mask = _u8img(norm_img), _u8img(mask)
        pose6 = np.ascontiguousarray(pose6, np.float64).reshape(norm_img.shape[0], 6)
        g_range = np.ascontiguousarray(g_range, np.float64)
        self._keep.append((norm_img, mask, pose6))
        self.shapes.append(norm_img.shape)
        self.L.ref_survey_add_prepared(self.h, int(img_id), _p(norm_img), _p(mask), norm_img.shape[0], norm_img.shape[1],
                                       _p(pose6), _p(g_range), len(g_range))

    def sample_step(self, frame_idx, pair_idx):
        """Bounded sample of the survey's work on the built frames (see ref_survey_sample_step); returns the row count."""
        fi = np.ascontiguousarray(frame_idx, np.int32)
        pi = np.ascontiguousarray(pair_idx, np.int32)
        return int(self.L.ref_survey_sample_step(self.h, _p(fi), len(fi), _p(pi), len(pi)))

    @property
    def threads(self):
        return int(self.L.ref_survey_threads(self.h))

    def build(self):
        self.L.ref_survey_build(self.h)
        self._keep = []

    def frame(self, i):
        return RefFrame(None, None, None, None, None, strict=self.strict,
                        handle=self.L.ref_survey_frame(self.h, int(i)), owner=self)

    def match(self, min_overlap=0.4, all_pairs=False):
        """Returns dict(pairs [P,2], overlap [P] f32, matched [P], counts [P], rows6 [K,6]) in (i, j) loop order."""
        total = self.L.ref_survey_match(self.h, C.c_float(min_overlap), int(all_pairs))
        P = self.L.ref_survey_npairs(self.h)
        F = self.L.ref_survey_nframes(self.h)
        overlap, matched, counts = np.empty(P, np.float32), np.empty(P, np.int32), np.empty(P, np.int32)
        self.L.ref_survey_pair_info(self.h, _p(overlap), _p(matched), _p(counts))
        rows6 = np.empty((total, 6), np.float64)
        if total:
            self.L.ref_survey_rows(self.h, _p(rows6))
        pairs = np.array([(i, j) for i in range(F) for j in range(i + 1, F)], np.int32).reshape(-1, 2)
        return dict(pairs=pairs, overlap=overlap, matched=matched, counts=counts, rows6=rows6)
