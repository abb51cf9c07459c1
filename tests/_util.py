"""Shared helpers of the test-suite (seeded inputs that need neither cv2 nor torch)."""
import numpy as np


def textured(rows, cols, seed):
    """numpy-only seeded textured uint8 image; also the input generator of tests/golden/make_golden.py."""
    g = np.random.default_rng(seed)
    a = g.normal(size=(rows + 8, cols + 8))
    k = np.array([1, 4, 6, 4, 1], np.float64); k /= k.sum()
    for _ in range(2):
        a = np.apply_along_axis(lambda v: np.convolve(v, k, mode="same"), 0, a)
        a = np.apply_along_axis(lambda v: np.convolve(v, k, mode="same"), 1, a)
    a = a[4:-4, 4:-4]
    a = (a - a.min()) / (a.max() - a.min()) * 200 + g.normal(size=(rows, cols)) * 5 + 25
    return np.clip(np.rint(a), 0, 255).astype(np.uint8)


def kps_triples(kps):
    return np.stack([kps["x"], kps["y"], kps["response"]], 1).astype(np.int32)


def oracle_frame(O, f, ex=None):
    """Diasss::Frame through the oracle: DetectFeature + GetGeoImg.  f: dict from diasss_b200.synth."""
    ex = ex or O.Extractor()
    k, d = ex(f["norm_img"])
    k, d, _ = O.mask_filter(k, d, f["mask"])
    gx, gy = O.geo_img(f["rows"], f["cols"], f["pose"], f["g_range"])
    return O.Frame(f["img_id"], f["rows"], f["cols"], k, d, gx, gy)
