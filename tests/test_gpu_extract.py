"""GPU parity tests, extraction: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs,
stage by stage -- pyramid levels (K1), FAST candidates in reference order (K2), quadtree selection (K3), orientation +
descriptors + assembly (K4-K6), mask filter -- and against the committed cv2-derived golden fixtures.
Everything is integer / defined-float work: the bar is bit-exact (byte equality)."""
import os

import numpy as np
import pytest

from tests._util import kps_triples, textured

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _ctx(**kw):
    from diasss_b200 import binding as B
    return B.Context(**kw)


def _compare_stages(O, ctx, img, nf, **okw):
    ex = O.Extractor(nf, **okw)
    ok, od = ex(img)
    gk, gd = ctx.extract(img)
    rows, cols = img.shape
    for l in range(ex.nlevels):
        if l:
            assert np.array_equal(ex.level_image(l), ctx.debug_level_image(0, l, rows, cols)), "pyramid level %d" % l
        assert np.array_equal(ex.candidates(l), ctx.debug_candidates(0, l)), "FAST candidates level %d" % l
        assert np.array_equal(kps_triples(ex.level_keys(l)), ctx.debug_level_keys(0, l)), "quadtree level %d" % l
    assert gk.tobytes() == ok.tobytes(), "keypoints"
    assert np.array_equal(gd, od), "descriptors"
    return ok, od


@pytest.mark.parametrize("shape,seed,nf", [((300, 260), 1, 2000), ((420, 640), 2, 500), ((640, 300), 3, 2000),
                                            ((250, 700), 4, 300), ((1000, 500), 5, 2000), ((97, 97), 6, 2000),
                                            ((333, 517), 7, 5000), ((1800, 1000), 8, 2000)])
def test_stages_vs_oracle(oracle, shape, seed, nf):
    from diasss_b200 import synth
    img = synth.make_survey(1, shape[0], shape[1], seed=seed)[0]["norm_img"]
    ctx = _ctx(nfeatures=nf)
    try:
        _compare_stages(oracle, ctx, img, nf)
    finally:
        ctx.close()


@pytest.mark.parametrize("shape,seed,nf", [((1100, 1300), 31, 5000), ((1500, 1400), 32, 10000), ((900, 2600), 33, 12000),
                                            ((3000, 2400), 34, 50000)])
def test_high_density_vs_oracle(oracle, shape, seed, nf):
    """BASELINE config 5 (high-density sweep): nfeatures 5k-50k.  Quotas beyond ~2500 keys per level move the quadtree's
    node arrays to global scratch and deepen the count grid (quadtree.cu); everything stays bit-exact."""
    img = textured(shape[0], shape[1], seed)
    ctx = _ctx(nfeatures=nf)
    try:
        k, _ = _compare_stages(oracle, ctx, img, nf)
        assert len(k) > 0.8 * nf
    finally:
        ctx.close()


@pytest.mark.parametrize("dense", ["0", "1"])
@pytest.mark.parametrize("shape,seed,nf", [((700, 900), 41, 2000), ((1100, 1300), 42, 8000), ((97, 130), 43, 2000)])
def test_both_descriptor_forms_vs_oracle(oracle, monkeypatch, shape, seed, nf, dense):
    """K5 / K6 exist in two forms with identical results: a 37x37 blurred window per keypoint (few keypoints per image) and
    a whole-level 13x13 blur followed by a warp per keypoint (many).  The library picks by bytes moved; here each is forced."""
    monkeypatch.setenv("DSX_DESCRIBE_DENSE", dense)
    img = textured(shape[0], shape[1], seed)
    ctx = _ctx(nfeatures=nf)
    try:
        _compare_stages(oracle, ctx, img, nf)
    finally:
        ctx.close()


@pytest.mark.parametrize("name", ["extract_a", "extract_b"])
def test_vs_golden_fixture(built, name):
    """CUDA path directly against outputs of the cv2-based restatement (real OpenCV primitives)."""
    z = np.load(os.path.join(G, name + ".npz"))
    r, c, seed, nf = (int(v) for v in z["shape_seed_nf"])
    ctx = _ctx(nfeatures=nf)
    try:
        kps, desc = ctx.extract(textured(r, c, seed))
        assert kps.tobytes() == z["kps"].tobytes() and np.array_equal(desc, z["desc"])
        for l in range(6):
            assert np.array_equal(ctx.debug_candidates(0, l), z["cand_%d" % l])
            if l:
                assert np.array_equal(ctx.debug_level_image(0, l, r, c), z["level_%d" % l])
    finally:
        ctx.close()


def test_edge_inputs(oracle):
    from diasss_b200 import binding as B
    ctx = _ctx()
    try:
        g = np.random.default_rng(0)
        for img in (np.zeros((128, 128), np.uint8), np.full((90, 400), 77, np.uint8),
                    g.integers(0, 256, (200, 333), dtype=np.uint8),          # pure noise: every cell saturated
                    (g.integers(0, 2, (160, 160)) * 255).astype(np.uint8),   # binary image: scores up to 254
                    textured(400, 170, 5)):                                   # tall: B1 (nIni clamp)
            _compare_stages(oracle, ctx, img, 2000)
        # empty image: silent return (ORBextractor.cpp:1052)
        k, d = ctx.extract(np.zeros((0, 0), np.uint8))
        assert len(k) == 0
        # a level smaller than 33 px: the reference is undefined -> DSX_ERR_INVALID
        with pytest.raises(B.DsxError) as e:
            ctx.extract(np.zeros((80, 80), np.uint8))
        assert e.value.status == B.ERR_INVALID
        # wrong type (reference: assert(image.type() == CV_8UC1), ORBextractor.cpp:1056)
        with pytest.raises(B.DsxError):
            ctx.extract(np.zeros((100, 100), np.float32))
        # row pitch != cols (ROI view of a larger buffer)
        big = textured(300, 400, 9)
        view = big[10:250, 20:341]
        k1, d1 = ctx.extract(view)
        k2, d2 = ctx.extract(np.ascontiguousarray(view))
        assert k1.tobytes() == k2.tobytes() and np.array_equal(d1, d2)
    finally:
        ctx.close()


@pytest.mark.parametrize("kw", [dict(nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th_fast=20, min_th_fast=7),
                                dict(nfeatures=800, scale_factor=1.5, nlevels=4, ini_th_fast=12, min_th_fast=5),
                                dict(nfeatures=64, scale_factor=1.2, nlevels=6, ini_th_fast=30, min_th_fast=30)])
def test_other_constructor_arguments(oracle, kw):
    img = textured(380, 450, 31)
    ctx = _ctx(**kw)
    try:
        t = ctx.tables()
        ex = oracle.Extractor(kw["nfeatures"], kw["scale_factor"], kw["nlevels"], kw["ini_th_fast"], kw["min_th_fast"])
        assert t["scale"].tobytes() == ex.scale.tobytes() and t["inv_scale"].tobytes() == ex.inv_scale.tobytes()
        assert np.array_equal(t["features_per_level"], ex.features_per_level) and np.array_equal(t["umax"], ex.umax)
        ok, od = ex(img)
        gk, gd = ctx.extract(img)
        assert gk.tobytes() == ok.tobytes() and np.array_equal(gd, od)
    finally:
        ctx.close()


@pytest.mark.parametrize("sf", [1.1, 1.2, 1.3, 1.5])
def test_pyramid_forms(oracle, sf, monkeypatch):
    """K1 has two forms: tiles staged by the TMA unit (16-byte aligned planes, scale factors up to ~1.3) and register-staged
    tiles (everything else; DSX_PYR_TMA=0 forces it).  Both must reproduce ComputePyramid (ORBextractor.cpp:1115-1140)
    bit for bit -- compared through the keypoints and descriptors of all levels -- on an image whose planes are aligned."""
    img = textured(520, 640, 77)
    ex = oracle.Extractor(1500, sf, 6, 20, 7)
    ok, od = ex(img)
    for tma in ("1", "0"):
        monkeypatch.setenv("DSX_PYR_TMA", tma)
        ctx = _ctx(nfeatures=1500, scale_factor=sf, nlevels=6, ini_th_fast=20, min_th_fast=7)
        try:
            gk, gd = ctx.extract(img)
            for l in range(1, ex.nlevels):
                assert np.array_equal(ex.level_image(l), ctx.debug_level_image(0, l, *img.shape)), "level %d, DSX_PYR_TMA=%s, scale %.2f" % (l, tma, sf)
            assert gk.tobytes() == ok.tobytes() and np.array_equal(gd, od), "DSX_PYR_TMA=%s, scale %.2f" % (tma, sf)
        finally:
            ctx.close()


@pytest.mark.parametrize("tma", ["2", "1", "0"])
def test_fast_staging_forms(oracle, tma, monkeypatch):
    """K2 stages a strip as one tensor-map box, as one bulk copy per row, or with 4-byte cp.async copies (DSX_FAST_TMA =
    2 / 1 / 0; unaligned planes always take the last): the same candidates, keypoints and descriptors in every form."""
    img = textured(520, 640, 78)
    monkeypatch.setenv("DSX_FAST_TMA", tma)
    ctx = _ctx()
    try:
        _compare_stages(oracle, ctx, img, 2000)
    finally:
        ctx.close()


def test_detect_feature_mask(oracle):
    """Frame::DetectFeature (frame.cpp:167-203): operator() + mask filter, order preserved."""
    from diasss_b200 import synth
    f = synth.make_pair(rows=420, cols=360, seed=7)[0]
    ctx = _ctx()
    try:
        k, d = oracle.Extractor()(f["norm_img"])
        k, d, _ = oracle.mask_filter(k, d, f["mask"])
        gk, gd = ctx.extract(f["norm_img"], f["mask"])
        assert len(k) < 2007 and gk.tobytes() == k.tobytes() and np.array_equal(gd, d)
        # all-zero mask -> nothing; all-255 mask -> operator() output
        gk0, _ = ctx.extract(f["norm_img"], np.zeros_like(f["mask"]))
        assert len(gk0) == 0
        gk1, gd1 = ctx.extract(f["norm_img"], np.full_like(f["mask"], 255))
        k1, d1 = ctx.extract(f["norm_img"])
        assert gk1.tobytes() == k1.tobytes() and np.array_equal(gd1, d1)
    finally:
        ctx.close()


def test_batched_device_path_equals_host_path(built):
    """dsx_detect_feature_batch_dev over more images than one internal chunk == per-image host calls; idempotent."""
    import torch
    from diasss_b200 import synth
    from diasss_b200.frontend import FrontEnd
    frames = synth.make_survey(5, 400, 360, seed=11)
    fe = FrontEnd(max_batch=2)
    try:
        imgs = torch.from_numpy(np.stack([f["norm_img"] for f in frames])).cuda()
        masks = torch.from_numpy(np.stack([f["mask"] for f in frames])).cuda()
        feats = fe.alloc_features(5)
        for rep in range(2):
            fe.ctx.detect_feature_batch_dev(imgs.data_ptr(), masks.data_ptr(), 5, 400, 360, 360, 400 * 360, feats["c"])
            torch.cuda.synchronize()
            cnt = feats["count"].cpu().numpy()
            kps = feats["kps"].cpu().numpy().view(np.uint8).reshape(5, fe.ctx.cap, 28)
            desc = feats["desc"].cpu().numpy()
            for i, f in enumerate(frames):
                hk, hd = fe.detect_feature(f["norm_img"], f["mask"])
                assert cnt[i] == len(hk)
                assert kps[i, :cnt[i]].tobytes() == hk.tobytes() and np.array_equal(desc[i, :cnt[i]], hd)
    finally:
        fe.ctx.close()


@pytest.mark.parametrize("cols,h2d_chunk", [(360, 0), (363, 2), (364, 3)])
def test_host_batch_pipeline_equals_device_path(built, cols, h2d_chunk):
    """dsx_detect_feature_batch: host images copied chunk by chunk on the library's copy stream, page-locked masks
    read in place by the mask filter (zero-copy), pageable masks staged -- all equal to the per-image host calls.
    cols = 363 exercises a staging pitch different from the caller's step."""
    import torch
    from diasss_b200 import synth
    from diasss_b200.frontend import FrontEnd
    n = 7
    frames = synth.make_survey(n, 400, 360, seed=13)
    g = np.random.default_rng(cols)
    imgs_np = np.stack([np.pad(f["norm_img"], ((0, 0), (0, cols - 360)), mode="reflect") for f in frames])
    masks_np = np.stack([np.pad(f["mask"], ((0, 0), (0, cols - 360)), mode="edge") for f in frames])
    masks_np[g.random(masks_np.shape) < 0.3] = 0        # make the filter drop a good share of the keypoints
    fe = FrontEnd(max_batch=4, h2d_chunk=h2d_chunk)
    try:
        ref = [fe.detect_feature(imgs_np[i], masks_np[i]) for i in range(n)]
        assert 0 < sum(len(r[0]) for r in ref)
        pinned_i, pinned_m = torch.from_numpy(imgs_np).pin_memory(), torch.from_numpy(masks_np).pin_memory()
        variants = [("pinned", pinned_i.data_ptr(), pinned_m.data_ptr()),
                    ("pageable", imgs_np.ctypes.data, masks_np.ctypes.data),
                    ("pinned image, pageable mask", pinned_i.data_ptr(), masks_np.ctypes.data)]
        if cols % 4 == 0:
            dev_i = torch.from_numpy(imgs_np).cuda()
            variants.append(("device image, pinned mask", dev_i.data_ptr(), pinned_m.data_ptr()))
        for name, pi, pm in variants:
            for rep in range(2):                         # second pass reuses the staging buffers
                feats = fe.alloc_features(n)
                fe.ctx.detect_feature_batch(pi, pm, n, 400, cols, cols, 400 * cols, feats["c"])
                torch.cuda.synchronize()
                cnt = feats["count"].cpu().numpy()
                kps = feats["kps"].cpu().numpy().view(np.uint8).reshape(n, fe.ctx.cap, 28)
                desc = feats["desc"].cpu().numpy()
                for i in range(n):
                    hk, hd = ref[i]
                    assert cnt[i] == len(hk), name
                    assert kps[i, :cnt[i]].tobytes() == hk.tobytes() and np.array_equal(desc[i, :cnt[i]], hd), name
        # no mask: operator() only
        feats = fe.alloc_features(n)
        fe.ctx.detect_feature_batch(pinned_i.data_ptr(), 0, n, 400, cols, cols, 400 * cols, feats["c"])
        torch.cuda.synchronize()
        k0, d0 = fe.ctx.extract(imgs_np[0])
        assert int(feats["count"][0]) == len(k0)
    finally:
        fe.ctx.close()


@pytest.mark.parametrize("lanes", ["1", "2"])
def test_host_batch_pipeline_many_chunks(built, monkeypatch, lanes):
    """40 single-image chunks (the chunk-size ramp once overflowed after 31 doublings), one and two extraction lanes,
    more chunks than staging buffers: == the device-resident call."""
    import torch
    from diasss_b200.frontend import FrontEnd
    from tests._util import textured
    monkeypatch.setenv("DSX_H2D_LANES", lanes)
    n, rows, cols = 40, 150, 180
    imgs_np = np.stack([textured(rows, cols, 500 + i % 5) for i in range(n)])
    imgs_np[1::2] = imgs_np[1::2, ::-1]                  # (5 distinct textures, half of them flipped)
    masks_np = np.full_like(imgs_np, 255)
    fe = FrontEnd(max_batch=8, h2d_chunk=1)
    try:
        dev_i, dev_m = torch.from_numpy(imgs_np).cuda(), torch.from_numpy(masks_np).cuda()
        want = fe.alloc_features(n)
        fe.ctx.detect_feature_batch_dev(dev_i.data_ptr(), dev_m.data_ptr(), n, rows, cols, cols, rows * cols, want["c"])
        pin_i, pin_m = torch.from_numpy(imgs_np).pin_memory(), torch.from_numpy(masks_np).pin_memory()
        for rep in range(2):
            got = fe.alloc_features(n)
            fe.ctx.detect_feature_batch(pin_i.data_ptr(), pin_m.data_ptr(), n, rows, cols, cols, rows * cols, got["c"])
            torch.cuda.synchronize()
            fe.ctx.check_error()
            cnt = want["count"].cpu().numpy()
            assert np.array_equal(got["count"].cpu().numpy(), cnt) and cnt.min() > 20
            wk, gk = want["kps"].cpu().numpy(), got["kps"].cpu().numpy()
            wd, gd = want["desc"].cpu().numpy(), got["desc"].cpu().numpy()
            for i in range(n):
                assert gk[i, :cnt[i]].tobytes() == wk[i, :cnt[i]].tobytes() and np.array_equal(gd[i, :cnt[i]], wd[i, :cnt[i]]), i
    finally:
        fe.ctx.close()


def test_frontend_orbextractor_interface(built, oracle):
    """Python mirror of ORB_SLAM2::ORBextractor: ctor arguments, operator(), getters (ORBextractor.h:51-83)."""
    from diasss_b200.frontend import ORBextractor
    orb = ORBextractor(2000, 1.2, 6, 12, 7)
    ex = oracle.Extractor()
    assert orb.GetLevels() == 6 and abs(orb.GetScaleFactor() - 1.2) < 1e-6
    assert orb.GetScaleFactors().tobytes() == ex.scale.tobytes()
    assert orb.GetInverseScaleFactors().tobytes() == ex.inv_scale.tobytes()
    assert np.allclose(orb.GetScaleSigmaSquares(), ex.scale * ex.scale, rtol=0, atol=0)
    img = textured(260, 300, 3)
    k, d = orb(img, mask=None)
    ok, od = ex(img)
    assert k.tobytes() == ok.tobytes() and np.array_equal(d, od)
    k, d = orb(np.zeros((0, 0), np.uint8))
    assert len(k) == 0 and d.shape == (0, 32)
    orb.ctx.close()


def _clustered(rows, cols, seed, n_blobs, rad):
    g = np.random.default_rng(seed)
    img = np.full((rows, cols), 90, np.uint8)
    tex = textured(rows, cols, seed)
    yy, xx = np.mgrid[0:rows, 0:cols]
    m = np.zeros((rows, cols), bool)
    for _ in range(n_blobs):
        cy, cx = g.integers(30, rows - 30), g.integers(30, cols - 30)
        m |= (yy - cy) ** 2 + (xx - cx) ** 2 < rad * rad
    img[m] = tex[m]
    return img


@pytest.mark.parametrize("shape,seed,nf,nb,rad", [((400, 500), 1, 2000, 2, 30), ((300, 900), 3, 1000, 3, 25),
                                                  ((700, 700), 5, 2000, 4, 15), ((1000, 400), 6, 500, 2, 50)])
def test_clustered_keys_general_quadtree_form(oracle, shape, seed, nf, nb, rad):
    """Texture only inside a few discs: the keys are strongly clustered, nodes deeper than the count pyramid must be
    split and the quadtree kernel switches to its general per-key form (and wide images have several root nodes)."""
    img = _clustered(shape[0], shape[1], seed, nb, rad)
    ctx = _ctx(nfeatures=nf)
    try:
        _compare_stages(oracle, ctx, img, nf)
    finally:
        ctx.close()
