"""GPU parity at the sizes BASELINE.json's configs name.

  config 2   two-image pair at the test_data shape (1800 x 1000, the reference-defined shape; 2000 x 1000 needs the
             nIni >= 1 definition of Appendix B1): keypoints, descriptors, CorresID_1/2, scc[0], rows -- bit-exact
  config 3   16-image survey, 2000 x 1000, all 120 pairs on the device path: rows of a sample of pairs == oracle,
             culled matching == brute-force matching on all pairs, batched extraction == single-frame extraction
  config 4   8000 x 2000 (the benchmark shape): one full-size image and one full-size pair against the oracle, plus
             size-independent properties of the batched path (determinism, batch == single, cull == brute force,
             keypoints under the mask and level-major, rows made of existing keypoints of the right frames)
"""
import numpy as np
import pytest

from tests._util import oracle_frame
from tests.test_gpu_match import _check_pair

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape,ids,seed", [((1800, 1000), (0, 1), 101), ((2000, 1000), (3, 4), 102)])
def test_config2_pair_at_test_data_shape(oracle, frontend, shape, ids, seed):
    from diasss_b200 import synth
    fa, fb = synth.make_pair(rows=shape[0], cols=shape[1], seed=seed, ids=ids)
    k = _check_pair(oracle, frontend, fa, fb)
    assert k > 50


def _device_survey(frames, rows, cols, **fe_kw):
    import torch
    from diasss_b200 import binding as B
    from diasss_b200.frontend import FrontEnd
    n = len(frames)
    fe = FrontEnd(**fe_kw)
    imgs = torch.from_numpy(np.stack([f["norm_img"] for f in frames])).cuda()
    masks = torch.from_numpy(np.stack([f["mask"] for f in frames])).cuda()
    models = [B.geo_model_build(f["pose"], rows, cols, f["g_range"]) for f in frames]
    rowtabs = torch.from_numpy(np.stack([m[0] for m in models])).cuda()
    granges = torch.from_numpy(np.stack([f["g_range"] for f in frames])).cuda()
    bboxes = np.stack([m[1] for m in models])
    ids = [f["img_id"] for f in frames]
    pairs = np.array([(i, j) for i in range(n) for j in range(i + 1, n)], np.int32)
    res = fe.process_survey(imgs, masks, rowtabs, granges, ids, bboxes, pairs)
    return fe, res, pairs, ids, bboxes


def test_config3_sixteen_image_survey(oracle):
    from diasss_b200 import synth
    from diasss_b200.frontend import FrontEnd
    n, rows, cols = 16, 2000, 1000
    frames = synth.make_survey(n, rows, cols, seed=303)
    fe, res, pairs, ids, bboxes = _device_survey(frames, rows, cols, max_batch=8)
    fe_bf = FrontEnd(match_cull=0)
    fe1 = FrontEnd()
    try:
        cnt = res["count"].cpu().numpy()[:len(pairs)]
        off = res["offset"].cpu().numpy()
        rows6 = res["rows6"].cpu().numpy()
        assert len(pairs) == 120 and off[-1] == len(rows6) and off[-1] > 1000
        # a sample of pairs against the oracle (neighbours, far apart, first / last)
        ex = oracle.Extractor()
        sample = [0, 1, 14, 15, 57, 100, 119]
        need = sorted({int(v) for p in sample for v in pairs[p]})
        of = {k: oracle_frame(oracle, frames[k], ex) for k in need}
        kps = res["feats"]["kps"].cpu().numpy().view(np.uint8)
        cnts = res["feats"]["count"].cpu().numpy()
        for k in need:                                      # batched extraction == oracle
            assert cnts[k] == len(of[k].kps)
            assert kps[k].reshape(-1)[:28 * cnts[k]].tobytes() == of[k].kps.tobytes()
        for p in sample:
            i, j = pairs[p]
            want = oracle.robust_matching(of[int(i)], of[int(j)])[0]
            assert cnt[p] == len(want), "pair %d" % p
            assert rows6[off[p]:off[p] + cnt[p]].tobytes() == want.tobytes(), "pair %d" % p
        # every distance evaluated (match_cull = 0) == gate-culled search, all 120 pairs
        bf = fe_bf.match_pairs(res["feats"], ids, [rows] * n, bboxes, pairs)
        assert bf["k"] == res["k"] and bf["rows6"].cpu().numpy().tobytes() == rows6.tobytes()
        # batched extraction == the single-frame host entry point (Frame::DetectFeature one frame at a time)
        for k in (0, 7, 15):
            kk, dd = fe1.detect_feature(frames[k]["norm_img"], frames[k]["mask"])
            assert kk.tobytes() == kps[k].reshape(-1)[:28 * cnts[k]].tobytes()
            assert np.array_equal(dd, res["feats"]["desc"][k, :cnts[k]].cpu().numpy())
    finally:
        for f in (fe, fe_bf, fe1):
            f.ctx.close()


def test_config4_full_size_image_and_pair(oracle):
    """8000 x 2000: extraction of one benchmark-size image and RobustMatching of one benchmark-size pair == oracle."""
    from diasss_b200 import synth
    from diasss_b200.frontend import FrontEnd
    rows, cols = 8000, 2000
    fa, fb = synth.make_pair(rows=rows, cols=cols, seed=404, ids=(0, 1))
    fe = FrontEnd()
    try:
        k = _check_pair(oracle, fe, fa, fb)
        assert k > 100
    finally:
        fe.ctx.close()


def test_config4_full_size_properties(built):
    """Size-independent properties at 8000 x 2000 on a 6-image survey (15 pairs)."""
    import torch
    from diasss_b200 import synth
    from diasss_b200.frontend import FrontEnd
    n, rows, cols = 6, 8000, 2000
    frames = synth.make_survey(n, rows, cols, seed=405)
    fe, res, pairs, ids, bboxes = _device_survey(frames, rows, cols, max_batch=4)
    fe_bf = FrontEnd(match_cull=0)
    try:
        rows6 = res["rows6"].cpu().numpy().copy()
        cnt = res["count"].cpu().numpy()[:len(pairs)].copy()
        off = res["offset"].cpu().numpy().copy()
        kps = res["feats"]["kps"].cpu().numpy().copy()
        desc = res["feats"]["desc"].cpu().numpy().copy()
        nk = res["feats"]["count"].cpu().numpy().copy()
        assert off[-1] > 1000 and nk.min() > 1000
        # keypoints lie inside the image, under a non-zero mask byte, level-major with non-decreasing octave
        for k in range(n):
            kk = kps[k, :nk[k]]
            x, y, octave = kk[:, 0], kk[:, 1], kk[:, 5].view(np.int32)
            assert x.min() >= 0 and x.max() < cols and y.min() >= 0 and y.max() < rows
            assert np.all(frames[k]["mask"][y.astype(np.int64), x.astype(np.int64)] != 0)
            assert np.all(np.diff(octave) >= 0) and octave.min() >= 0 and octave.max() <= 5
        # determinism: a second pass over the same inputs gives the same bytes
        imgs = torch.from_numpy(np.stack([f["norm_img"] for f in frames])).cuda()
        masks = torch.from_numpy(np.stack([f["mask"] for f in frames])).cuda()
        feats2 = fe.alloc_features(n)
        fe.ctx.detect_feature_batch_dev(imgs.data_ptr(), masks.data_ptr(), n, rows, cols, cols, rows * cols, feats2["c"])
        assert torch.equal(feats2["count"].cpu(), torch.from_numpy(nk))
        k2, d2_ = feats2["kps"].cpu().numpy(), feats2["desc"].cpu().numpy()
        for k in range(n):
            assert k2[k, :nk[k]].tobytes() == kps[k, :nk[k]].tobytes() and np.array_equal(d2_[k, :nk[k]], desc[k, :nk[k]])
        # one image at a time == the batch
        one = fe.alloc_features(1)
        for k in (0, n - 1):
            fe.ctx.detect_feature_batch_dev(imgs[k].data_ptr(), masks[k].data_ptr(), 1, rows, cols, cols, rows * cols, one["c"])
            assert int(one["count"][0]) == nk[k]
            assert one["kps"][0, :nk[k]].cpu().numpy().tobytes() == kps[k, :nk[k]].tobytes()
        # brute force == culled
        bf = fe_bf.match_pairs(res["feats"], ids, [rows] * n, bboxes, pairs)
        assert bf["rows6"].cpu().numpy().tobytes() == rows6.tobytes()
        # a descriptor's distance to itself is 0 and distances are symmetric (DescriptorDistance)
        d0 = fe.descriptor_distance(desc[0, :256], desc[0, :256])
        d1 = fe.descriptor_distance(desc[0, :256], desc[1, :256])
        d2 = fe.descriptor_distance(desc[1, :256], desc[0, :256])
        assert np.all(d0 == 0) and np.array_equal(d1, d2) and d1.max() <= 256
        # rows name the pair's ids and coordinates of existing keypoints of the two frames
        for p in (0, len(pairs) - 1):
            i, j = pairs[p]
            r = rows6[off[p]:off[p] + cnt[p]]
            assert np.all(r[:, 0] == ids[i]) and np.all(r[:, 1] == ids[j])
            si = {(float(a), float(b)) for a, b in kps[i, :nk[i], :2][:, ::-1]}
            tj = {(float(a), float(b)) for a, b in kps[j, :nk[j], :2][:, ::-1]}
            assert all((a, b) in si for a, b in r[:, 2:4]) and all((a, b) in tj for a, b in r[:, 4:6])
    finally:
        fe.ctx.close(); fe_bf.ctx.close()


def test_survey_host_call_equals_device_path(built):
    """dsx_survey_host (host images, masks, poses in; geo model built inside by host threads under the image copies)
    gives the bytes of the device-resident path, and the frames' geo bounding boxes."""
    import torch
    from diasss_b200 import binding as B, synth
    n, rows, cols = 10, 1500, 1000
    frames = synth.make_survey(n, rows, cols, seed=511)
    fe, res, pairs, ids, bboxes = _device_survey(frames, rows, cols, max_batch=4)
    try:
        want_rows = res["rows6"].cpu().numpy().copy()
        want_cnt = res["count"].cpu().numpy()[:len(pairs)].copy()
        want_kps = res["feats"]["kps"].cpu().numpy().copy()
        h_imgs = torch.from_numpy(np.stack([f["norm_img"] for f in frames])).pin_memory()
        h_masks = torch.from_numpy(np.stack([f["mask"] for f in frames])).pin_memory()
        poses = np.ascontiguousarray(np.stack([f["pose"] for f in frames]), np.float64)
        granges = np.ascontiguousarray(np.stack([f["g_range"] for f in frames]), np.float64)
        feats = fe.alloc_features(n)
        out = fe.alloc_match_out(len(pairs), feats["count"].device)
        bb = np.zeros((n, 4), np.float64)
        for _ in range(2):      # the second call reuses the context's staging
            k = fe.ctx.survey_host(h_imgs.data_ptr(), h_masks.data_ptr(), n, rows, cols, cols, rows * cols, poses, granges, ids, pairs,
                                   feats["c"], out["count"].data_ptr(), out["offset"].data_ptr(), out["rows6"].data_ptr(),
                                   out["rows6"].shape[0], bbox_out=bb)
            assert k == len(want_rows)
            assert out["rows6"][:k].cpu().numpy().tobytes() == want_rows.tobytes()
            assert np.array_equal(out["count"].cpu().numpy()[:len(pairs)], want_cnt)
            assert feats["kps"].cpu().numpy().tobytes() == want_kps.tobytes()
            assert bb.tobytes() == np.ascontiguousarray(bboxes, np.float64).tobytes()
        # no pairs: features only, offset[0] = 0
        k = fe.ctx.survey_host(h_imgs.data_ptr(), h_masks.data_ptr(), n, rows, cols, cols, rows * cols, poses, granges, ids,
                               np.zeros((0, 2), np.int32), feats["c"], out["count"].data_ptr(), out["offset"].data_ptr(),
                               out["rows6"].data_ptr(), out["rows6"].shape[0])
        assert k == 0 and int(out["offset"][0].item()) == 0
        # too few ground ranges: the reference reads past the vector (B4); the call refuses
        with pytest.raises(B.DsxError):
            fe.ctx.survey_host(h_imgs.data_ptr(), h_masks.data_ptr(), n, rows, cols, cols, rows * cols, poses,
                               np.ascontiguousarray(granges[:, :cols // 2]), ids, pairs, feats["c"], out["count"].data_ptr(),
                               out["offset"].data_ptr(), out["rows6"].data_ptr(), out["rows6"].shape[0])
    finally:
        fe.ctx.close()


def test_survey_host_back_to_back_without_synchronisation(built):
    """Two dsx_survey_host calls on DIFFERENT surveys enqueued back to back with no synchronisation in between (what a
    pipelined service does: the second survey's images cross PCIe while the first is still being matched; staging
    buffers, the pinned geo-model block and the matcher scratch are shared) -- both must give the bytes of the
    device-resident path."""
    import torch
    from diasss_b200 import synth
    n, rows, cols = 8, 1200, 1000
    surveys = [synth.make_survey(n, rows, cols, seed=s, drift_m=d) for s, d in ((601, 1.5), (602, 0.7))]
    want = []
    for frames in surveys:
        fe0, res, pairs, ids, bboxes = _device_survey(frames, rows, cols, max_batch=4)
        want.append((res["rows6"].cpu().numpy().copy(), res["count"].cpu().numpy()[:len(pairs)].copy(),
                     res["feats"]["kps"].cpu().numpy().copy(), np.ascontiguousarray(bboxes, np.float64).copy()))
        fe0.ctx.close()
    from diasss_b200.frontend import FrontEnd
    fe = FrontEnd(max_batch=4, h2d_chunk=2)
    try:
        held = []
        for frames in surveys:
            h_imgs = torch.from_numpy(np.stack([f["norm_img"] for f in frames])).pin_memory()
            h_masks = torch.from_numpy(np.stack([f["mask"] for f in frames])).pin_memory()
            poses = np.ascontiguousarray(np.stack([f["pose"] for f in frames]), np.float64)
            granges = np.ascontiguousarray(np.stack([f["g_range"] for f in frames]), np.float64)
            feats = fe.alloc_features(n)
            out = fe.alloc_match_out(len(pairs), feats["count"].device)
            bb = np.zeros((n, 4), np.float64)
            for rep in range(3):        # the same survey several times, then the other one, never synchronising
                fe.ctx.survey_host(h_imgs.data_ptr(), h_masks.data_ptr(), n, rows, cols, cols, rows * cols, poses, granges, ids, pairs,
                                   feats["c"], out["count"].data_ptr(), out["offset"].data_ptr(), out["rows6"].data_ptr(),
                                   out["rows6"].shape[0], sync=False, bbox_out=bb)
            granges_scratch = granges.copy()
            granges[:] = -1.0           # the caller's ground ranges need not outlive the call
            held.append((h_imgs, h_masks, poses, granges_scratch, feats, out, bb))
        torch.cuda.synchronize()
        fe.ctx.check_error()
        for (rows_w, cnt_w, kps_w, bb_w), (_, _, _, _, feats, out, bb) in zip(want, held):
            k = int(out["offset"][len(pairs)].item())
            assert k == len(rows_w) and k > 500
            assert out["rows6"][:k].cpu().numpy().tobytes() == rows_w.tobytes()
            assert np.array_equal(out["count"].cpu().numpy()[:len(pairs)], cnt_w)
            assert feats["kps"].cpu().numpy().tobytes() == kps_w.tobytes()
            assert bb.tobytes() == bb_w.tobytes()
    finally:
        fe.ctx.close()
