"""GPU parity tests, matching: K7 (gated Hamming top-2 + acceptance), K8 (SCC RANSAC), K9 (ConsistentCheck + rows)
through the C ABI against the CPU oracle: CorresID_1/2, scc[0], emitted index pairs and corres_kps rows -- bit-exact."""
import numpy as np
import pytest

from tests._util import oracle_frame

pytestmark = pytest.mark.gpu


def _gpu_frame(fe, f):
    return fe.make_frame(f)


def _check_pair(O, fe, fa, fb, oa=None, ob=None, ga=None, gb=None):
    ex = O.Extractor()
    oa = oa or oracle_frame(O, fa, ex)
    ob = ob or oracle_frame(O, fb, ex)
    ga = ga or _gpu_frame(fe, fa)
    gb = gb or _gpu_frame(fe, fb)
    assert ga["kps"].tobytes() == oa.kps.tobytes() and gb["kps"].tobytes() == ob.kps.tobytes()
    r = fe.ctx.match_debug(ga, gb)
    rows6, si, ti, c1, c2 = O.robust_matching(oa, ob)
    s1, s2 = O.geo_nn_search(oa, ob), O.geo_nn_search(ob, oa)
    assert np.array_equal(r["corres1"], c1), "CorresID_1"
    assert np.array_equal(r["corres2"], c2), "CorresID_2"
    for d, s in enumerate((s1, s2)):
        if s["scc"]:
            best = sorted(s["scc"], reverse=True)[0]          # std::sort(scc.rbegin(), scc.rend()), FEAmatcher.cpp:331
            assert r["scc_count"][d] == best[0] and r["scc_model"][d] == best[1], "scc[0] dir %d" % d
        else:
            assert r["scc_count"][d] == 0
    assert np.array_equal(r["src_idx"], si) and np.array_equal(r["tgt_idx"], ti), "emitted pairs / order"
    assert r["rows6"].tobytes() == rows6.tobytes(), "corres_kps rows"
    return len(rows6)


@pytest.mark.parametrize("ids,seed,shape", [((0, 1), 7, (420, 360)), ((2, 4), 8, (420, 360)), ((1, 3), 9, (300, 500)),
                                            ((5, 2), 10, (640, 300)), ((0, 1), 12, (1000, 500))])
def test_robust_matching_vs_oracle(oracle, frontend, ids, seed, shape):
    from diasss_b200 import synth
    fa, fb = synth.make_pair(rows=shape[0], cols=shape[1], seed=seed, ids=ids)
    k = _check_pair(oracle, frontend, fa, fb)
    assert k > 10


import os

_SLOW = os.environ.get("DSX_SLOW_TESTS", "0") != "0"


@pytest.mark.parametrize("nf,shape", [(5000, (1500, 1200)), (10000, (1500, 1200)), (20000, (3000, 2400)),
                                      pytest.param(50000, (3000, 2400), marks=pytest.mark.skipif(
                                          not _SLOW, reason="the CPU oracle needs ~1 min for 43k x 43k keypoints; set DSX_SLOW_TESTS=1"))])
def test_high_density_pair_vs_oracle(oracle, nf, shape):
    """BASELINE config 5: 5k-50k keypoints per image through extraction + gated matching + SCC + merge, bit-exact, with
    the default (compacting, gate-culled) matcher.  Beyond ~10k keypoints the per-keypoint working arrays of K7/K8 live
    in global scratch; beyond 16384 keypoints an image's search-axis sort runs in global memory instead of shared."""
    from diasss_b200 import synth
    from diasss_b200.frontend import FrontEnd
    fa, fb = synth.make_pair(rows=shape[0], cols=shape[1], seed=40 + nf // 1000, ids=(0, 1))
    fe = FrontEnd(nfeatures=nf)
    try:
        ex = oracle.Extractor(nf)
        oa, ob = oracle_frame(oracle, fa, ex), oracle_frame(oracle, fb, ex)
        assert len(oa.kps) > 0.6 * nf
        k = _check_pair(oracle, fe, fa, fb, oa, ob)
        assert k > 100
    finally:
        fe.ctx.close()


@pytest.mark.parametrize("nf,cull", [(800, 1), (300, 1), (800, 0)])
def test_small_capacity_pair_vs_oracle(oracle, nf, cull):
    """nfeatures <= ~900 (capacity <= 1024 keypoints): the matcher's one-source-per-lane instantiations, compacting
    (default) and brute force."""
    from diasss_b200 import synth
    from diasss_b200.frontend import FrontEnd
    fa, fb = synth.make_pair(rows=420, cols=360, seed=60 + nf // 100, ids=(0, 1))
    fe = FrontEnd(nfeatures=nf, match_cull=cull)
    try:
        assert fe.ctx.cap <= 1024
        ex = oracle.Extractor(nf)
        oa, ob = oracle_frame(oracle, fa, ex), oracle_frame(oracle, fb, ex)
        k = _check_pair(oracle, fe, fa, fb, oa, ob)
        assert k > 10
    finally:
        fe.ctx.close()


def test_large_drift_partial_overlap(oracle, frontend):
    """DR drift of several metres: some keypoints fall outside the reference bbox / the 8 m gate."""
    from diasss_b200 import synth
    fa, fb = synth.make_survey(2, 420, 360, seed=21, drift_m=6.0, spread=1.2)
    _check_pair(oracle, frontend, fa, fb)


def test_no_overlap_empty_and_noise(oracle, frontend):
    from diasss_b200 import synth
    fa, fb = synth.make_pair(rows=300, cols=280, seed=13)
    ex = oracle.Extractor()
    oa, ob = oracle_frame(oracle, fa, ex), oracle_frame(oracle, fb, ex)
    ga, gb = _gpu_frame(frontend, fa), _gpu_frame(frontend, fb)
    # (1) disjoint geo boxes: B3 -> nothing
    far_o = oracle.Frame(ob.img_id, ob.rows, ob.cols, ob.kps, ob.desc, ob.geo_x + 1e4, ob.geo_y)
    far_g = dict(gb); far_g["geo_xy"] = gb["geo_xy"] + np.array([1e4, 0.0]); far_g["bbox"] = gb["bbox"] + np.array([1e4, 1e4, 0, 0])
    assert _check_pair(oracle, frontend, fa, fb, oa, far_o, ga, far_g) == 0
    # (2) empty target frame
    emp_o = oracle.Frame(3, ob.rows, ob.cols, ob.kps[:0], ob.desc[:0], ob.geo_x, ob.geo_y)
    emp_g = dict(gb); emp_g.update(img_id=3, kps=gb["kps"][:0], desc=gb["desc"][:0], geo_xy=gb["geo_xy"][:0])
    assert _check_pair(oracle, frontend, fa, fb, oa, emp_o, ga, emp_g) == 0
    # (3) descriptors replaced by noise: gate passes, Hamming/ratio reject almost everything
    g = np.random.default_rng(5)
    nd = g.integers(0, 256, ob.desc.shape, dtype=np.uint8)
    noi_o = oracle.Frame(ob.img_id, ob.rows, ob.cols, ob.kps, nd, ob.geo_x, ob.geo_y)
    noi_g = dict(gb); noi_g["desc"] = nd
    _check_pair(oracle, frontend, fa, fb, oa, noi_o, ga, noi_g)
    # (4) identical frames with the same parity: every keypoint matches itself or a duplicate
    same_o = oracle.Frame(2, oa.rows, oa.cols, oa.kps, oa.desc, oa.geo_x, oa.geo_y)
    same_g = dict(ga); same_g["img_id"] = 2
    assert _check_pair(oracle, frontend, fa, fa, oa, same_o, ga, same_g) > 50


@pytest.mark.parametrize("n", [1, 2, 31, 32, 33, 64, 100, 128, 256])
def test_scc_inlier_count_sizes(oracle, frontend, n):
    """SCC_x (FEAmatcher.cpp:186-248) with exactly n tentative matches: a frame against a copy of its first n keypoints
    whose along-track coordinates are spread (so that models differ in their inlier counts).  Covers the sizes at which
    the sorted-offset search of scc_merge_kernel changes shape (powers of two and their neighbours)."""
    from diasss_b200 import synth
    fa, _ = synth.make_pair(rows=300, cols=280, seed=21)
    ex = oracle.Extractor()
    oa = oracle_frame(oracle, fa, ex)
    ga = _gpu_frame(frontend, fa)
    assert len(oa.kps) >= n
    g = np.random.default_rng(n)
    kps = oa.kps[:n].copy()
    kps["y"] += g.integers(-6, 7, n).astype(np.float32)           # keypoint y (along track) moved by a few pings
    sub_o = oracle.Frame(2, oa.rows, oa.cols, kps, oa.desc[:n], oa.geo_x, oa.geo_y)
    kps["y"] = np.clip(kps["y"], 0, oa.rows - 1)
    yi, xi = kps["y"].astype(np.int64), kps["x"].astype(np.int64)        # geo_img.at<double>(int(y), int(x)), FEAmatcher.cpp:81-82
    sub_g = dict(ga); sub_g.update(img_id=2, kps=kps, desc=ga["desc"][:n],
                                   geo_xy=np.ascontiguousarray(np.stack([oa.geo_x[yi, xi], oa.geo_y[yi, xi]], 1)))
    full_o = oracle.Frame(oa.img_id, oa.rows, oa.cols, oa.kps[:n], oa.desc[:n], oa.geo_x, oa.geo_y)
    full_g = dict(ga); full_g.update(kps=ga["kps"][:n], desc=ga["desc"][:n], geo_xy=ga["geo_xy"][:n])
    r = frontend.ctx.match_debug(full_g, sub_g)
    rows6, si, ti, c1, c2 = oracle.robust_matching(full_o, sub_o)
    assert np.array_equal(r["corres1"], c1) and np.array_equal(r["corres2"], c2)
    assert r["rows6"].tobytes() == rows6.tobytes()
    s1 = oracle.geo_nn_search(full_o, sub_o)
    if s1["scc"]:
        best = sorted(s1["scc"], reverse=True)[0]
        assert r["scc_count"][0] == best[0] and r["scc_model"][0] == best[1]


@pytest.mark.parametrize("env", [dict(DSX_MATCH_AUTON="1", DSX_MATCH_COLUMNS="0"), dict(DSX_MATCH_AUTON="1", DSX_MATCH_COLUMNS="2"),
                                 dict(DSX_MATCH_AUTON="0"), dict(DSX_MATCH_COMPACT="0"), dict(DSX_SCC_SORTED="0")])
def test_matcher_forms(oracle, built, env, monkeypatch):
    """The matcher's alternative forms on the same pair: the warp-autonomous form (the default
    only for dense images) in the plain and in the column-major sort order, the CTA-staged form, the non-compacting form
    and the linear SCC count all reproduce RobustMatching (FEAmatcher.cpp:13-50)."""
    from diasss_b200 import synth
    from diasss_b200.frontend import FrontEnd
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    fe = FrontEnd()
    try:
        fa, fb = synth.make_pair(rows=420, cols=360, seed=23, ids=(0, 1))
        assert _check_pair(oracle, fe, fa, fb) > 20
        fa, fb = synth.make_pair(rows=300, cols=500, seed=24, ids=(3, 5))
        _check_pair(oracle, fe, fa, fb)
    finally:
        fe.ctx.close()


def test_geo_near_neigh_search_entry(oracle, frontend):
    """dsx_geo_near_neigh_search == FEAmatcher::GeoNearNeighSearch (one direction)."""
    from diasss_b200 import synth
    fa, fb = synth.make_pair(rows=360, cols=300, seed=15, ids=(4, 7))
    ex = oracle.Extractor()
    oa, ob = oracle_frame(oracle, fa, ex), oracle_frame(oracle, fb, ex)
    ga, gb = _gpu_frame(frontend, fa), _gpu_frame(frontend, fb)
    for (o1, o2, g1, g2) in ((oa, ob, ga, gb), (ob, oa, gb, ga)):
        want = oracle.geo_nn_search(o1, o2)
        corres, cnt, model = frontend.geo_near_neigh_search(g1, g2)
        assert np.array_equal(corres, want["corres"])
        best = sorted(want["scc"], reverse=True)[0]
        assert (cnt, model) == best


def test_descriptor_distance(oracle, frontend):
    g = np.random.default_rng(2)
    a = g.integers(0, 256, (777, 32), dtype=np.uint8)
    b = g.integers(0, 256, (777, 32), dtype=np.uint8)
    b[:5] = a[:5]
    got = frontend.descriptor_distance(a, b)
    want = np.array([oracle.descriptor_distance(x, y) for x, y in zip(a, b)])
    assert np.array_equal(got, want) and np.all(got[:5] == 0)


def test_survey_all_pairs_device_path(oracle):
    """test_demo's two loops on the device for a 5-image survey (10 pairs) == oracle pair by pair, rows in (i,j) order."""
    import torch
    from diasss_b200 import binding as B, synth
    from diasss_b200.frontend import FrontEnd
    n, rows, cols = 5, 400, 360
    frames = synth.make_survey(n, rows, cols, seed=17)
    fe = FrontEnd(max_batch=3)
    try:
        imgs = torch.from_numpy(np.stack([f["norm_img"] for f in frames])).cuda()
        masks = torch.from_numpy(np.stack([f["mask"] for f in frames])).cuda()
        models = [B.geo_model_build(f["pose"], rows, cols, f["g_range"]) for f in frames]
        rowtabs = torch.from_numpy(np.stack([m[0] for m in models])).cuda()
        granges = torch.from_numpy(np.stack([f["g_range"] for f in frames])).cuda()
        bboxes = np.stack([m[1] for m in models])
        pairs = np.array([(i, j) for i in range(n) for j in range(i + 1, n)], np.int32)
        res = fe.process_survey(imgs, masks, rowtabs, granges, [f["img_id"] for f in frames], bboxes, pairs)
        cnt = res["count"].cpu().numpy()[:len(pairs)]
        off = res["offset"].cpu().numpy()
        rows6 = res["rows6"].cpu().numpy()
        ex = oracle.Extractor()
        of = [oracle_frame(oracle, f, ex) for f in frames]
        want = [oracle.robust_matching(of[i], of[j])[0] for i, j in pairs]
        assert cnt.tolist() == [len(w) for w in want]
        assert off[-1] == res["k"] == sum(len(w) for w in want)
        assert rows6.tobytes() == np.concatenate(want).tobytes()
        assert sum(cnt) > 100
    finally:
        fe.ctx.close()


@pytest.mark.parametrize("h2d_chunk,shuffle", [(0, False), (2, True), (3, False)])
def test_survey_call_equals_separate_calls(built, h2d_chunk, shuffle):
    """dsx_survey (extraction, geo look-ups and matching pipelined inside the library, pairs matched as soon as both
    images have arrived) == dsx_detect_feature_batch_dev + dsx_georef_batch_dev + dsx_match_pairs_dev, byte for byte and
    in the caller's pair order -- from page-locked host images, pageable host images and device images."""
    import torch
    from diasss_b200 import binding as B, synth
    from diasss_b200.frontend import FrontEnd
    n, rows, cols = 7, 400, 360
    frames = synth.make_survey(n, rows, cols, seed=23)
    fe = FrontEnd(max_batch=3, h2d_chunk=h2d_chunk)
    try:
        imgs_np = np.stack([f["norm_img"] for f in frames]); masks_np = np.stack([f["mask"] for f in frames])
        imgs, masks = torch.from_numpy(imgs_np).cuda(), torch.from_numpy(masks_np).cuda()
        models = [B.geo_model_build(f["pose"], rows, cols, f["g_range"]) for f in frames]
        rowtabs = torch.from_numpy(np.stack([m[0] for m in models])).cuda()
        granges = torch.from_numpy(np.stack([f["g_range"] for f in frames])).cuda()
        bboxes = np.stack([m[1] for m in models])
        ids = [f["img_id"] for f in frames]
        pairs = np.array([(i, j) for i in range(n) for j in range(i + 1, n)], np.int32)
        if shuffle:
            g = np.random.default_rng(1)
            pairs = pairs[g.permutation(len(pairs))]
            pairs[::3] = pairs[::3, ::-1]                      # some pairs with the later image first
        ref = fe.process_survey(imgs, masks, rowtabs, granges, ids, bboxes, pairs)
        want = (ref["count"].cpu().numpy()[:len(pairs)].copy(), ref["offset"].cpu().numpy().copy(), ref["rows6"].cpu().numpy().copy())
        want_kps = ref["feats"]["kps"].cpu().numpy().copy()
        assert want[1][-1] > 100
        pin_i, pin_m = torch.from_numpy(imgs_np).pin_memory(), torch.from_numpy(masks_np).pin_memory()
        for name, pi, pm in (("pinned", pin_i.data_ptr(), pin_m.data_ptr()), ("pageable", imgs_np.ctypes.data, masks_np.ctypes.data),
                             ("device", imgs.data_ptr(), masks.data_ptr())):
            feats = fe.alloc_features(n)
            out = fe.alloc_match_out(len(pairs), imgs.device)
            k = fe.ctx.survey(pi, pm, n, rows, cols, cols, rows * cols, rowtabs.data_ptr(), granges.data_ptr(), granges.shape[1], ids, bboxes,
                              pairs, feats["c"], out["count"].data_ptr(), out["offset"].data_ptr(), out["rows6"].data_ptr(), out["rows6"].shape[0])
            assert k == want[1][-1], name
            assert np.array_equal(out["count"].cpu().numpy()[:len(pairs)], want[0]), name
            assert np.array_equal(out["offset"].cpu().numpy(), want[1]), name
            assert out["rows6"][:k].cpu().numpy().tobytes() == want[2].tobytes(), name
            assert feats["kps"].cpu().numpy().tobytes() == want_kps.tobytes(), name
    finally:
        fe.ctx.close()


def test_get_kps_pairs_after_the_survey(oracle):
    """The step after the path (first half): Optimizer::GetKpsPairs (optimizer.cpp:575-639, USE_ANNO = 0)
    for every pair of a matched survey on the device == the oracle's restatement (itself pinned against the
    reference's own lines in tests/test_ref_pin.py) applied to each pair's rows."""
    import torch
    from diasss_b200 import synth
    from tests.test_gpu_configs import _device_survey
    n, rows, cols = 6, 900, 700
    frames = synth.make_survey(n, rows, cols, seed=77)
    fe, res, pairs, ids, bboxes = _device_survey(frames, rows, cols, max_batch=4)
    try:
        g = np.random.default_rng(4)
        alts = g.uniform(8.0, 25.0, (n, rows))
        gras = np.stack([f["g_range"] for f in frames])
        n_range = gras.shape[1]
        d_alt, d_gra = torch.from_numpy(alts).cuda(), torch.from_numpy(gras).cuda()
        P = len(pairs)
        cnt, off, rows6 = res["count"], res["offset"], res["rows6"]
        out7 = torch.zeros(max(int(rows6.shape[0]), 1), 7, dtype=torch.float64, device="cuda")
        ocnt = torch.zeros(P, dtype=torch.int32, device="cuda")
        fe.ctx.get_kps_pairs_dev(rows6.data_ptr(), cnt.data_ptr(), off.data_ptr(), pairs, ids, d_alt.data_ptr(), rows, d_gra.data_ptr(),
                                 n_range, n_range, out7.data_ptr(), ocnt.data_ptr())
        torch.cuda.synchronize()
        cnt_h, off_h, rows_h, out_h, ocnt_h = cnt.cpu().numpy(), off.cpu().numpy(), rows6.cpu().numpy(), out7.cpu().numpy(), ocnt.cpu().numpy()
        kept = 0
        for p, (i, j) in enumerate(pairs):
            r = rows_h[off_h[p]:off_h[p] + cnt_h[p]]
            want = oracle.get_kps_pairs(r, ids[j], alts[i], gras[i], alts[j], gras[j])
            assert ocnt_h[p] == len(want), "pair %d" % p
            assert out_h[off_h[p]:off_h[p] + len(want)].tobytes() == want.tobytes(), "pair %d" % p
            kept += len(want)
        assert 0 < kept < int(off_h[P])          # the nadir band dropped some, kept most
    finally:
        fe.ctx.close()
