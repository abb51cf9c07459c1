"""CPU: the matcher oracle (oracle/match_oracle.cpp) against a plain-numpy restatement of FEAmatcher.cpp's ORB branch
written independently here, plus the Appendix-B edge cases."""
import numpy as np
import pytest

from diasss_b200 import synth
from tests._util import oracle_frame


def _np_search(f, ref):
    """GeoNearNeighSearch main loop in numpy (FEAmatcher.cpp:79-183, ORB branch); returns pre-SCC CorresID."""
    n = len(f.kps)
    out = np.full(n, -1, np.int64)
    fx = f.geo_x[f.kps["y"].astype(int), f.kps["x"].astype(int)]; fy = f.geo_y[f.kps["y"].astype(int), f.kps["x"].astype(int)]
    rx = ref.geo_x[ref.kps["y"].astype(int), ref.kps["x"].astype(int)]; ry = ref.geo_y[ref.kps["y"].astype(int), ref.kps["x"].astype(int)]
    bound = 80 if (f.img_id % 2 != ref.img_id % 2) else 88
    bits_r = np.unpackbits(ref.desc, axis=1)
    for i in range(n):
        if fx[i] < ref.geo_x.min() or fy[i] < ref.geo_y.min() or fx[i] > ref.geo_x.max() or fy[i] > ref.geo_y.max():
            continue
        cand = np.nonzero(np.sqrt((fx[i] - rx) * (fx[i] - rx) + (fy[i] - ry) * (fy[i] - ry)) < 8)[0]
        if len(cand) == 0:
            continue
        d = (np.unpackbits(f.desc[i])[None, :] != bits_r[cand]).sum(1)
        order = np.argsort(d, kind="stable")
        best, bid = int(d[order[0]]), int(cand[order[0]])
        sec = int(d[order[1]]) if len(cand) > 1 else 1000
        if best <= bound and sec != 1000 and best / sec <= 0.35 if sec else False:
            out[i] = bid
        elif len(cand) == 1 and best <= bound:
            out[i] = bid
    return out


@pytest.mark.parametrize("ids", [(0, 1), (2, 4)])
def test_search_vs_numpy(oracle, ids):
    pair = synth.make_pair(rows=300, cols=280, seed=5, ids=ids)
    ex = oracle.Extractor()
    a, b = (oracle_frame(oracle, f, ex) for f in pair)
    r = oracle.geo_nn_search(a, b)
    assert np.array_equal(r["pre"], _np_search(a, b))
    assert (r["pre"] >= 0).sum() > 20                     # the synthetic pair really matches
    # SCC output is a subset of the tentative matches and reports its own size
    fin = r["corres"]
    assert np.all((fin == -1) | (fin == r["pre"]))
    assert r["scc"] and r["scc"][-1][0] == (fin >= 0).sum()
    counts = [c for c, _ in r["scc"]]
    assert counts == sorted(set(counts))                  # strictly increasing pushes (:237-242)


def test_robust_matching_rows(oracle):
    pair = synth.make_pair(rows=300, cols=280, seed=6, ids=(0, 1))
    ex = oracle.Extractor()
    a, b = (oracle_frame(oracle, f, ex) for f in pair)
    rows6, si, ti, c1, c2 = oracle.robust_matching(a, b)
    assert len(rows6) > 0
    assert np.all(rows6[:, 0] == 0) and np.all(rows6[:, 1] == 1)
    assert np.array_equal(rows6[:, 2], a.kps["y"][si].astype(np.float64)) and np.array_equal(rows6[:, 3], a.kps["x"][si].astype(np.float64))
    assert np.array_equal(rows6[:, 4], b.kps["y"][ti].astype(np.float64)) and np.array_equal(rows6[:, 5], b.kps["x"][ti].astype(np.float64))
    # every emitted pair comes from one of the two directions
    for s, t in zip(si, ti):
        assert c1[s] == t or c2[t] == s


def test_no_overlap_and_empty(oracle):
    """B3: no tentative match in either direction -> nothing emitted (the reference would index an empty vector)."""
    pair = synth.make_pair(rows=260, cols=240, seed=8, ids=(0, 1))
    ex = oracle.Extractor()
    a, b = (oracle_frame(oracle, f, ex) for f in pair)
    far = oracle.Frame(b.img_id, b.rows, b.cols, b.kps, b.desc, b.geo_x + 1e4, b.geo_y)
    rows6, si, ti, c1, c2 = oracle.robust_matching(a, far)
    assert len(rows6) == 0 and np.all(c1 == -1) and np.all(c2 == -1)
    empty = oracle.Frame(3, b.rows, b.cols, b.kps[:0], b.desc[:0], b.geo_x, b.geo_y)
    rows6, *_ = oracle.robust_matching(a, empty)
    assert len(rows6) == 0


def test_frame_glue(oracle):
    """mask filter (frame.cpp:184-195), GetGeoImg (:126-165), ComputeIntersection (util.cpp:13-43)."""
    f = synth.make_pair(rows=200, cols=180, seed=3)[0]
    k, d = oracle.Extractor()(f["norm_img"])
    k2, d2, idx = oracle.mask_filter(k, d, f["mask"])
    keep = f["mask"][k["y"].astype(int), k["x"].astype(int)] != 0
    assert np.array_equal(idx, np.nonzero(keep)[0]) and k2.tobytes() == k[keep].tobytes() and np.array_equal(d2, d[keep])
    gx, gy = oracle.geo_img(f["rows"], f["cols"], f["pose"], f["g_range"])
    half = f["cols"] // 2
    i, j = 17, half + 5
    assert gx[i, j] == f["pose"][i, 3] + f["g_range"][5] * np.cos(f["pose"][i, 2] + 3.14159265359 / 2)
    j = 3
    assert gy[i, j] == f["pose"][i, 4] + f["g_range"][half - 3] * np.sin(f["pose"][i, 2] - 3.14159265359 / 2)
    assert abs(oracle.compute_intersection((gx, gy), (gx, gy)) - 1.0) < 1e-6
    assert oracle.compute_intersection((gx, gy), (gx + 1e5, gy)) == 0.0


def test_frame_prepare_restatement(oracle):
    """Frame::GetNormalizeSSS / GetFilteredMask in the oracle against a direct numpy restatement of frame.cpp:57-124
    (sequential mean), and the library-order mean against the sequential one."""
    g = np.random.default_rng(4)
    rows, cols = 400, 260
    raw = np.abs(g.normal(1.0, 0.3, (rows, cols)))
    raw[g.random((rows, cols)) < 1e-3] *= 30.0
    raw[[2, 7, rows - 2], [3, 9, cols - 1]] = 100.0
    s = 0.0
    for v in raw.ravel().tolist():
        s += v
    mean = s / raw.size
    assert oracle.mean(raw, 0) == mean and abs(oracle.mean(raw, 1) - mean) <= 1e-13 * mean
    mn = raw.min()
    v = np.minimum((raw - mn) / (mean * 2.5 - mn) * 255.0, 255.0)
    assert np.array_equal(oracle.normalize_sss(raw), np.clip(np.rint(v), 0, 255).astype(np.uint8))
    want = np.full((rows, cols), 255, np.uint8)
    for i, j in zip(*np.nonzero(raw > mean * np.float32(2.5))):
        if i >= 6 and j >= 6:
            want[i - 6:i + 6, j - 6:j + 6] = 0
    want[:, cols // 2 - 9:cols // 2 + 10] = 0
    want[:150] = 0; want[rows - 149:] = 0
    want[:, :90] = 0; want[:, cols - 89:] = 0
    assert np.array_equal(oracle.filtered_mask(raw), want)
