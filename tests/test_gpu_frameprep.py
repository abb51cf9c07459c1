"""GPU parity, the step before the path: Frame::GetNormalizeSSS and Frame::GetFilteredMask
(frame.cpp:57-124) on the device against the oracle.  With the mean summed in the library's documented 32-lane order
the planes are bit-exact; against a sequentially summed mean (OpenCV leaves the order undefined) they may differ by
one grey level at isolated pixels."""
import numpy as np
import pytest

from tests._util import textured

pytestmark = pytest.mark.gpu


def raw_sss(rows, cols, seed):
    """Seeded raw side-scan intensities (CV_64F): texture + multiplicative speckle + sparse 'buggy line' samples,
    some of them on the image border where the reference's stamp loop wraps / clips (Appendix B5)."""
    g = np.random.default_rng(seed)
    img = (textured(rows, cols, seed).astype(np.float64) + 3.0) * 1.7e-3 * (1.0 + 0.1 * g.standard_normal((rows, cols)))
    img = np.abs(img)
    hot = g.random((rows, cols)) < 2e-4
    hot[[0, 3, 5, 6, rows - 1, rows - 3], g.integers(0, cols, 6)] = True
    hot[g.integers(0, rows, 6), [0, 2, 5, 6, cols - 1, cols - 4]] = True
    hot[200:203, 150:260] = True                      # a stretch of a sensor line
    img[hot] *= 40.0
    return img


@pytest.mark.parametrize("shape,seed", [((500, 420), 1), ((333, 401), 2), ((700, 1000), 3)])
def test_frame_prepare_vs_oracle(oracle, shape, seed):
    import torch
    from diasss_b200.frontend import FrontEnd
    n = 3
    raws = np.stack([raw_sss(shape[0], shape[1], 10 * seed + k) for k in range(n)])
    fe = FrontEnd()
    try:
        d_raw = torch.from_numpy(raws).cuda()
        step = (shape[1] + 3) & ~3                     # device images of the extractor need a row pitch that is a multiple of 4
        norm_d = torch.zeros(n, shape[0], step, dtype=torch.uint8, device="cuda")
        mask_d = torch.zeros_like(norm_d)
        stats = torch.zeros(n, 3, dtype=torch.float64, device="cuda")
        fe.ctx.frame_prepare_batch_dev(d_raw.data_ptr(), n, shape[0], shape[1], norm_d.data_ptr(), mask_d.data_ptr(), step=step,
                                       stats_ptr=stats.data_ptr())
        torch.cuda.synchronize()
        norm, mask, stats = norm_d.cpu().numpy()[:, :, :shape[1]], mask_d.cpu().numpy()[:, :, :shape[1]], stats.cpu().numpy()
        for k in range(n):
            m1 = oracle.mean(raws[k], 1)
            assert stats[k, 0] == m1 and stats[k, 1] == raws[k].min() and stats[k, 2] == raws[k].max()
            assert np.array_equal(norm[k], oracle.normalize_sss(raws[k], 1)), "norm_img"
            wm = oracle.filtered_mask(raws[k], 1)
            assert np.array_equal(mask[k], wm), "flt_mask"
            assert 0 < np.count_nonzero(wm) < wm.size and np.count_nonzero(wm[160:-160, 100:-100] == 0) > 100   # stamps present
            # sequential-order mean: same to ~1e-15 relative; planes equal up to isolated one-level differences
            m0 = oracle.mean(raws[k], 0)
            assert abs(m0 - m1) <= 1e-13 * abs(m0)
            d = norm[k].astype(int) - oracle.normalize_sss(raws[k], 0).astype(int)
            assert np.abs(d).max() <= 1 and np.count_nonzero(d) <= 1e-4 * d.size
            assert np.count_nonzero(mask[k] != oracle.filtered_mask(raws[k], 0)) <= 1e-4 * d.size
        # the prepared planes drive extraction exactly like host-prepared ones
        feats = fe.alloc_features(n)
        fe.ctx.detect_feature_batch_dev(norm_d.data_ptr(), mask_d.data_ptr(), n, shape[0], shape[1], step, shape[0] * step, feats["c"])
        torch.cuda.synchronize()
        ex = oracle.Extractor()
        k0, d0 = ex(np.ascontiguousarray(norm[0]))
        k0, d0, _ = oracle.mask_filter(k0, d0, np.ascontiguousarray(mask[0]))
        c = int(feats["count"][0])
        assert c == len(k0) and feats["kps"][0, :c].cpu().numpy().view(np.uint8).tobytes() == k0.tobytes()
    finally:
        fe.ctx.close()
