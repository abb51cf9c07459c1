"""TEST INFRASTRUCTURE ONLY -- second, independent restatement of the reference extractor that calls the
REAL OpenCV (cv2) for every primitive the reference takes from OpenCV:

    cv2.resize(INTER_LINEAR)          <- ORBextractor.cpp:1128
    cv2.FastFeatureDetector (9_16)    <- ORBextractor.cpp:809,814   (one call per ~30 px cell, like the reference)
    cv2.GaussianBlur(13x13, sigma 2)  <- ORBextractor.cpp:1092
    cv2.fastAtan2                     <- ORBextractor.cpp:103

and restates only the ORB-SLAM2 control flow in numpy.  DistributeOctTree (ORBextractor.cpp:539-763) is
written here in *array form* (list position arithmetic instead of std::list pointer surgery): this is the
formulation the CUDA kernel uses, so agreement of this file with oracle/orb_oracle.cpp (which keeps the
reference's std::list form) checks both the C++ oracle and the array reformulation.

It is slow (python loops); used by tests/ on small images and by tests/golden/make_golden.py to produce the
committed fixtures.  Never imported by the product.
"""
import ctypes
import ctypes.util
import math

import cv2
import numpy as np

_LIBM = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
_LIBM.cosf.restype = _LIBM.sinf.restype = ctypes.c_float
_LIBM.cosf.argtypes = _LIBM.sinf.argtypes = [ctypes.c_float]

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                     ("octave", "<i4"), ("class_id", "<i4")])
EDGE = 19
f32 = np.float32


def cv_round(v):
    return int(np.rint(v))


def tables(nfeatures=2000, scale_factor=1.2, nlevels=6):
    sf = float(f32(scale_factor))                       # double member holding the float argument
    scale = [f32(1.0)]
    for _ in range(1, nlevels):
        scale.append(f32(float(scale[-1]) * sf))
    inv = [f32(1.0) / s for s in scale]
    factor = f32(1.0 / sf)
    nd = f32(f32(nfeatures) * (f32(1) - factor) / (f32(1) - f32(math.pow(float(factor), float(nlevels)))))
    quota, tot = [], 0
    for _ in range(nlevels - 1):
        quota.append(cv_round(nd)); tot += quota[-1]; nd = f32(nd * factor)
    quota.append(max(nfeatures - tot, 0))
    umax = [0] * 16
    vmax = int(math.floor(float(f32(15) * f32(math.sqrt(2.0)) / f32(2) + f32(1))))
    vmin = int(math.ceil(float(f32(15) * f32(math.sqrt(2.0)) / f32(2))))
    for v in range(vmax + 1):
        umax[v] = cv_round(math.sqrt(225.0 - v * v))
    v0 = 0
    for v in range(15, vmin - 1, -1):
        while umax[v0] == umax[v0 + 1]:
            v0 += 1
        umax[v] = v0
        v0 += 1
    return scale, inv, quota, umax


def pyramid(img, inv):
    rows, cols = img.shape
    levels = [img]
    for l in range(1, len(inv)):
        c = cv_round(f32(cols) * inv[l]); r = cv_round(f32(rows) * inv[l])
        levels.append(cv2.resize(levels[-1], (c, r), interpolation=cv2.INTER_LINEAR))
    return levels


def cell_candidates(im, ini_th=12, min_th=7):
    """ORBextractor.cpp:771-829: list of (x, y, response), coords relative to (16,16)."""
    det_hi = cv2.FastFeatureDetector_create(ini_th, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    det_lo = cv2.FastFeatureDetector_create(min_th, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    minB = EDGE - 3
    maxBX, maxBY = im.shape[1] - EDGE + 3, im.shape[0] - EDGE + 3
    width, height = maxBX - minB, maxBY - minB
    nC, nR = int(width / 30), int(height / 30)
    out = []
    if nC == 0 or nR == 0:
        return out
    wC, hC = int(math.ceil(f32(width) / f32(nC))), int(math.ceil(f32(height) / f32(nR)))
    for i in range(nR):
        iniY = minB + i * hC
        maxY = iniY + hC + 6
        if iniY >= maxBY - 3:
            continue
        maxY = min(maxY, maxBY)
        for j in range(nC):
            iniX = minB + j * wC
            maxX = iniX + wC + 6
            if iniX >= maxBX - 6:
                continue
            maxX = min(maxX, maxBX)
            roi = np.ascontiguousarray(im[iniY:maxY, iniX:maxX])
            k = det_hi.detect(roi)
            if len(k) == 0:
                k = det_lo.detect(roi)
            for p in k:
                out.append((int(p.pt[0]) + j * wC, int(p.pt[1]) + i * hC, int(p.response)))
    return out


def distribute_array_form(cands, w, h, N):
    """DistributeOctTree (ORBextractor.cpp:539-763) in array form.

    cands: list of (x, y, resp) in emission order.  Returns the selected candidates in list order.
    A node is (x0, y0, x1, y1, [candidate indices]); the python list `L` IS the std::list, front first.
    """
    nIni = int(math.floor(float(f32(w) / f32(h)) + 0.5))          # round(): half away from zero (positive)
    nIni = max(1, nIni)                                            # B1
    hX = f32(w) / f32(nIni)
    roots = [[int(hX * f32(i)), 0, int(hX * f32(i + 1)), h, []] for i in range(nIni)]
    for idx, (x, y, r) in enumerate(cands):
        roots[int(f32(x) / hX)][4].append(idx)
    L = [n for n in roots if len(n[4]) > 0]

    def split(n):
        x0, y0, x1, y1, keys = n
        hx, hy = (x1 - x0 + 1) // 2, (y1 - y0 + 1) // 2            # ceil(d/2)
        kids = [[x0, y0, x0 + hx, y0 + hy, []], [x0 + hx, y0, x1, y0 + hy, []],
                [x0, y0 + hy, x0 + hx, y1, []], [x0 + hx, y0 + hy, x1, y1, []]]
        for k in keys:
            x, y, _ = cands[k]
            q = (0 if x < x0 + hx else 1) + (0 if y < y0 + hy else 2)
            kids[q][4].append(k)
        return [c for c in kids if len(c[4]) > 0]                 # creation order n1..n4

    first = True
    E = None       # expandable nodes created by the previous step, in creation order
    while True:
        m = len(L)
        # --- normal pass: split every node with >1 keys, front to back
        exp_pos = [i for i, n in enumerate(L) if len(n[4]) > 1]
        groups = [split(L[i]) for i in exp_pos]                    # processing order = list order
        front = []
        for g in reversed(groups):                                 # last processed group is foremost
            front.extend(reversed(g))                              # n4 foremost inside a group
        keep = [n for i, n in enumerate(L) if len(n[4]) == 1]
        L = front + keep
        n_expand = sum(1 for g in groups for c in g if len(c[4]) > 1)
        first = False
        if len(L) >= N or len(L) == m:
            break
        if len(L) + 3 * n_expand > N:
            # --- final phase: largest nodes first, most recently created first among equals (B2)
            done = False
            while not done:
                m = len(L)
                P = [i for i, n in enumerate(L) if len(n[4]) > 1]  # all live in the front part; position asc = creation desc
                P.sort(key=lambda i: (-len(L[i][4]), i))
                size = m
                processed, groups = [], []
                for i in P:
                    g = split(L[i])
                    processed.append(i); groups.append(g)
                    size += len(g) - 1
                    if size >= N:
                        break
                front = []
                for g in reversed(groups):
                    front.extend(reversed(g))
                ps = set(processed)
                L = front + [n for i, n in enumerate(L) if i not in ps]
                if len(L) >= N or len(L) == m:
                    done = True
            break
    out = []
    for n in L:
        best = n[4][0]
        for k in n[4][1:]:
            if cands[k][2] > cands[best][2]:
                best = k
        out.append(cands[best])
    return out


def ic_angle(im, x, y, umax):
    m01 = m10 = 0
    for u in range(-15, 16):
        m10 += u * int(im[y, x + u])
    for v in range(1, 16):
        d = umax[v]
        vs = 0
        for u in range(-d, d + 1):
            vp, vm = int(im[y + v, x + u]), int(im[y - v, x + u])
            vs += vp - vm
            m10 += u * (vp + vm)
        m01 += v * vs
    return cv2.fastAtan2(float(m01), float(m10))


def orb_descriptor(blur, x, y, angle, pat):
    ang = f32(angle) * f32(math.pi / float(f32(180.0)))
    a, b = f32(_LIBM.cosf(float(ang))), f32(_LIBM.sinf(float(ang)))    # the platform libm, like the reference (:113)
    px, py = pat[:, 0].astype(f32), pat[:, 1].astype(f32)
    fy = (px * b).astype(f32) + (py * a).astype(f32)
    fx = (px * a).astype(f32) - (py * b).astype(f32)
    iy, ix = np.rint(fy).astype(np.int64), np.rint(fx).astype(np.int64)
    vals = blur[y + iy, x + ix].astype(np.int32)
    bits = (vals[0::2] < vals[1::2]).astype(np.uint8)
    return np.packbits(bits.reshape(32, 8), axis=1, bitorder="little").ravel()


def extract(img, pat1024, nfeatures=2000, scale_factor=1.2, nlevels=6, ini_th=12, min_th=7, want_stages=False):
    """ORBextractor::operator() in ORB mode.  Returns (kps[KP_DTYPE], desc[n,32]) (+ stages dict)."""
    scale, inv, quota, umax = tables(nfeatures, scale_factor, nlevels)
    pat = np.asarray(pat1024, np.int32).reshape(512, 2)
    levels = pyramid(img, inv)
    stages = dict(levels=levels, candidates=[], level_keys=[])
    kps_all, desc_all = [], []
    for l, im in enumerate(levels):
        cands = cell_candidates(im, ini_th, min_th)
        w, h = im.shape[1] - 2 * (EDGE - 3), im.shape[0] - 2 * (EDGE - 3)
        sel = distribute_array_form(cands, w, h, quota[l])
        stages["candidates"].append(np.array(cands, np.int32).reshape(-1, 3))
        size = f32(int(f32(31) * scale[l]))
        keys = np.zeros(len(sel), KP_DTYPE)
        for i, (x, y, r) in enumerate(sel):
            X, Y = x + 16, y + 16
            keys[i] = (X, Y, size, ic_angle(im, X, Y, umax), r, l, -1)
        stages["level_keys"].append(keys.copy())
        if len(sel) == 0:
            continue
        blur = cv2.GaussianBlur(im.copy(), (13, 13), 2, sigmaY=2, borderType=cv2.BORDER_REFLECT_101)
        d = np.zeros((len(sel), 32), np.uint8)
        for i in range(len(sel)):
            d[i] = orb_descriptor(blur, int(keys[i]["x"]), int(keys[i]["y"]), keys[i]["angle"], pat)
        if l != 0:
            keys["x"] = keys["x"] * scale[l]
            keys["y"] = keys["y"] * scale[l]
        kps_all.append(keys); desc_all.append(d)
    if kps_all:
        kps, desc = np.concatenate(kps_all), np.concatenate(desc_all)
    else:
        kps, desc = np.zeros(0, KP_DTYPE), np.zeros((0, 32), np.uint8)
    return (kps, desc, stages) if want_stages else (kps, desc)
