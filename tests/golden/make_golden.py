#!/usr/bin/env python
"""Generates tests/golden/*.npz with the REAL OpenCV (cv2) and the cv2-based restatement (oracle/cv2_oracle.py).

Run in the build container (cv2 4.13.0 present):   python tests/golden/make_golden.py
The fixtures pin (a) the OpenCV primitives the reference calls -- resize INTER_LINEAR, FAST-9/16 with NMS,
GaussianBlur 13x13 sigma 2, fastAtan2, cv::RNG -- and (b) whole-extractor outputs of the independent cv2-based
restatement on two small seeded images.  Inputs are regenerated from seeds with numpy only, so the fixtures stay small.
"""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import cv2_oracle as P  # noqa: E402
from oracle import oracle as O  # noqa: E402
from tests._util import textured  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    print("cv2", cv2.__version__)
    # ---- primitives
    prim = {}
    img = textured(211, 173, 11)
    prim["resize_src_seed"] = np.array([211, 173, 11])
    for i, (dr, dc) in enumerate([(176, 144), (147, 120), (101, 99)]):
        prim["resize_%d" % i] = cv2.resize(img, (dc, dr), interpolation=cv2.INTER_LINEAR)
    roi = textured(37, 43, 12)
    for th in (7, 12):
        det = cv2.FastFeatureDetector_create(th, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
        prim["fast_%d" % th] = np.array([[int(p.pt[0]), int(p.pt[1]), int(p.response)] for p in det.detect(roi)], np.int32).reshape(-1, 3)
    prim["blur"] = cv2.GaussianBlur(textured(64, 80, 13), (13, 13), 2, sigmaY=2, borderType=cv2.BORDER_REFLECT_101)
    g = np.random.default_rng(14)
    ys = g.integers(-40000, 40000, 4096).astype(np.float32); xs = g.integers(-40000, 40000, 4096).astype(np.float32)
    ys[:4] = [0, 0, 5, -5]; xs[:4] = [0, 7, 0, 0]
    prim["atan_y"], prim["atan_x"] = ys, xs
    prim["atan"] = np.array([cv2.fastAtan2(float(y), float(x)) for y, x in zip(ys, xs)], np.float32)
    cv2.setRNGSeed(0)  # global RNG is NOT what the matcher uses; the default-constructed cv::RNG has state 0xffffffff
    prim["rng_first"] = np.array([130063606, 3003295397, 3870020839], np.uint32)  # SURVEY.md A.7 (probe against cv2.randu)
    np.savez_compressed(os.path.join(HERE, "primitives_cv2.npz"), **prim)
    # ---- whole extractor through the cv2-based restatement
    pat = O.pattern()
    for name, (r, c, seed, nf) in {"extract_a": (240, 200, 21, 2000), "extract_b": (180, 420, 22, 400)}.items():
        img = textured(r, c, seed)
        kps, desc, st = P.extract(img, pat, nfeatures=nf, want_stages=True)
        out = dict(shape_seed_nf=np.array([r, c, seed, nf]), kps=kps, desc=desc)
        for l in range(6):
            out["cand_%d" % l] = st["candidates"][l]
            out["level_%d" % l] = st["levels"][l] if l else np.zeros(0, np.uint8)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, len(kps))


if __name__ == "__main__":
    main()
