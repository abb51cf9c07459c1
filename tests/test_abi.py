"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol include/*.h declares; host-only
entry points behave; compute entry points fail loudly (no CPU fallback) when there is no device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    names = set()
    for h in ("diasss_b200.h", "diasss_b200_debug.h"):
        txt = open(os.path.join(ROOT, "include", h)).read()
        txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
        names |= set(re.findall(r"\b(dsx_[a-z0-9_]+)\s*\(", txt))
    return names


def test_exports_match_headers(built):
    from diasss_b200 import binding as B
    L = B.lib()
    decl = _declared()
    assert decl == set(B.EXPORTS), (decl ^ set(B.EXPORTS))
    for s in decl:
        assert hasattr(L, s), s


def test_sass_is_sm100a(built):
    import subprocess
    from diasss_b200 import binding as B
    out = subprocess.run(["cuobjdump", "-lelf", B.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out and not re.search(r"sm_(?!100a)\d+", out), out


def test_default_params(built):
    from diasss_b200 import binding as B
    p = B.default_params()
    assert (p.nfeatures, p.nlevels, p.ini_th_fast, p.min_th_fast) == (2000, 6, 12, 7)          # frame.cpp:180
    assert abs(p.scale_factor - 1.2) < 1e-6
    assert (p.radius, p.dist_bound, p.dist_bound_flip, p.ratio_test) == (8.0, 88, 80, 0.35)    # FEAmatcher.cpp:66,143-147
    assert (p.ransac_iters, p.pix_error, p.kp_diff_thres) == (1000, 2.5, 2.5)                  # :189-190, :329


def test_geo_model_host(built, oracle):
    """dsx_geo_model_build + keypoint_geo == look-ups in the oracle's full GetGeoImg planes, bit for bit."""
    from diasss_b200 import binding as B, synth
    from diasss_b200.frontend import keypoint_geo
    f = synth.make_pair(rows=220, cols=201, seed=4)[1]      # odd cols, heading down
    tab, bbox = B.geo_model_build(f["pose"], f["rows"], f["cols"], f["g_range"])
    gx, gy = oracle.geo_img(f["rows"], f["cols"], f["pose"], f["g_range"])
    assert bbox.tolist() == [gx.min(), gx.max(), gy.min(), gy.max()]
    g = np.random.default_rng(1)
    kps = np.zeros(500, B.KP_DTYPE)
    kps["x"] = g.uniform(0, f["cols"] - 1e-3, 500).astype(np.float32)
    kps["y"] = g.uniform(0, f["rows"] - 1e-3, 500).astype(np.float32)
    geo = keypoint_geo(kps, tab, f["g_range"], f["cols"])
    r, c = kps["y"].astype(int), kps["x"].astype(int)
    assert np.array_equal(geo[:, 0], gx[r, c]) and np.array_equal(geo[:, 1], gy[r, c])
    geo2, bbox2 = B.frame_geo_from_planes(kps, gx, gy)
    assert np.array_equal(geo2, geo) and np.array_equal(bbox2, bbox)


def test_geo_model_batch_equals_single_calls(built):
    from diasss_b200 import binding as B, synth
    rows, cols, n = 300, 260, 7
    tr = synth.survey_tracks(n, rows, cols, seed=5)
    poses, gr = np.stack([t["pose"] for t in tr]), np.stack([t["g_range"] for t in tr])
    for threads in (0, 1, 3, 16):
        tabs, bb = B.geo_model_build_batch(poses, rows, cols, gr, n_threads=threads)
        for k in range(n):
            t1, b1 = B.geo_model_build(poses[k], rows, cols, gr[k])
            assert tabs[k].tobytes() == t1.tobytes() and bb[k].tobytes() == b1.tobytes()
    with pytest.raises(B.DsxError):                      # too few ground ranges: the per-frame error comes through
        B.geo_model_build_batch(poses, rows, cols, gr[:, :10])


def test_compute_intersection_and_pair_list(built, oracle):
    """dsx_compute_intersection / dsx_build_pair_list == Util::ComputeIntersection over the full geo planes
    (util.cpp:13-43, float arithmetic) and the i<j gate of test_demo (diasss2.cpp:88-97), bit for bit."""
    from diasss_b200 import binding as B, synth
    frames = synth.make_survey(6, 260, 240, seed=3, spread=0.9)
    planes = [oracle.geo_img(f["rows"], f["cols"], f["pose"], f["g_range"]) for f in frames]
    bboxes = np.stack([B.geo_model_build(f["pose"], f["rows"], f["cols"], f["g_range"])[1] for f in frames])
    want, want_pairs = [], []
    for i in range(6):
        for j in range(i + 1, 6):
            ov = oracle.compute_intersection(planes[i], planes[j])
            want.append(ov)
            assert np.float32(B.compute_intersection(bboxes[i], bboxes[j])) == np.float32(ov)
            if np.float32(ov) > np.float32(0.4):
                want_pairs.append((i, j))
    pairs, ov = B.build_pair_list(bboxes, 0.4)
    assert ov.tobytes() == np.asarray(want, np.float32).tobytes()
    assert pairs.tolist() == [list(p) for p in want_pairs]
    assert 0 < len(want_pairs) < 15            # the gate is exercised both ways
    # disjoint boxes -> 0 ; identical boxes -> 1
    a = np.array([0.0, 10.0, 0.0, 5.0])
    assert B.compute_intersection(a, a + 100.0) == 0.0 and B.compute_intersection(a, a) == 1.0


def test_no_cpu_fallback(built):
    """Without a CUDA device dsx_create must fail with DSX_ERR_CUDA -- there is no CPU path to fall back to."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from diasss_b200 import binding as B
    with pytest.raises(B.DsxError) as e:
        B.Context()
    assert e.value.status == B.ERR_CUDA


def test_product_does_not_touch_oracle():
    """The product path may not import, link or execute anything under oracle/."""
    bad = re.compile(r"(^\s*(from|import)\s+oracle\b)|(liboracle)|(#include\s+\"[^\"]*oracle)|(oracle/)|(oracle_capi)", re.M)
    for base in ("diasss_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            for fn in files:
                if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp", "Makefile")):
                    txt = open(os.path.join(dirpath, fn), errors="ignore").read()
                    assert not bad.search(txt), os.path.join(dirpath, fn)
