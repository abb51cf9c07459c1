"""The reference's on-disk formats and BASELINE config 1's stand-in, on the CPU.

  * dsx_io_read_matrix against FileStorage XML files written by the real OpenCV (tests/golden/io, cv2 4.13) -- the
    writer behind the reference's `ct_img` / `auv_pose` / `anno_kps` files (util.cpp:90-116, :188-191);
  * dsx_io_write_matrix -> dsx_io_read_matrix round trips, and (when cv2 is importable) cv2 reading our files;
  * dsx_io_read_column with the line semantics of util.cpp:127-179;
  * config 1 ("test_demo on bundled test_data, CPU"): the reference's test_data is not published, so a 5-frame synthetic
    survey is stored in the same folder layout and formats, loaded back (bit-exact) and pushed through the CPU oracle's
    Frame construction, overlap gate and RobustMatching -- the reference's front end, restated, on its own input format.
"""
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden", "io")


def test_reads_opencv_written_files(built):
    from diasss_b200 import binding as B
    want = np.load(os.path.join(GOLD, "arrays.npz"))
    for fn, node, key in (("ssh-170_img.xml", "ct_img", "ct"), ("ssh-170_pose.xml", "auv_pose", "pose"), ("ssh-170_anno.xml", "anno_kps", "anno")):
        got = B.io_read_matrix(os.path.join(GOLD, fn), node)
        assert got.dtype == want[key].dtype and got.shape == want[key].shape
        assert got.tobytes() == want[key].tobytes(), node
    p = os.path.join(GOLD, "several_nodes.xml")
    assert B.io_read_matrix(p, "ct_img").tobytes() == want["ct"][3:6].tobytes()            # not the `ct_img_backup` node before it
    assert B.io_read_matrix(p, "ct_img_backup").tobytes() == want["ct"][:3].tobytes()
    assert B.io_read_matrix(p, "f32").tobytes() == want["f32"].tobytes()
    assert B.io_read_matrix(p, "u8").tobytes() == want["u8"].tobytes()
    sp = B.io_read_matrix(p, "special")
    assert np.array_equal(sp, want["special"])           # (.Inf / -.Inf; OpenCV itself writes -0.0 as "0.")
    with pytest.raises(B.DsxError):
        B.io_read_matrix(p, "ct")                       # a prefix of a node name is not a node
    with pytest.raises(B.DsxError):
        B.io_read_matrix(os.path.join(GOLD, "no_such_file.xml"), "ct_img")


def test_write_read_round_trip(built, tmp_path):
    from diasss_b200 import binding as B
    g = np.random.default_rng(3)
    cases = dict(d=g.normal(0, 1e3, (37, 11)) * 10.0 ** g.integers(-12, 12, (37, 11)), f=g.normal(0, 5, (5, 9)).astype(np.float32),
                 i=g.integers(-2 ** 31, 2 ** 31 - 1, (4, 7)).astype(np.int32), s=g.integers(-32768, 32767, (3, 3)).astype(np.int16),
                 w=g.integers(0, 65535, (2, 5)).astype(np.uint16), u=g.integers(0, 255, (8, 8)).astype(np.uint8),
                 c=g.integers(-128, 127, (2, 2)).astype(np.int8))
    cases["d"][0, :4] = [np.inf, -np.inf, 5e-324, 1.7976931348623157e308]
    for k, a in cases.items():
        p = str(tmp_path / ("m_%s.xml" % k))
        B.io_write_matrix(p, "node_" + k, a)
        got = B.io_read_matrix(p, "node_" + k)
        assert got.dtype == a.dtype and got.tobytes() == a.tobytes(), k
    B.io_write_matrix(str(tmp_path / "empty.xml"), "anno_kps", np.zeros((0, 7), np.int32))
    assert B.io_read_matrix(str(tmp_path / "empty.xml"), "anno_kps").shape == (0, 7)
    try:
        import cv2
    except ImportError:
        return
    for k in ("d", "f", "i", "u"):                       # the real OpenCV reads what we write
        fs = cv2.FileStorage(str(tmp_path / ("m_%s.xml" % k)), cv2.FILE_STORAGE_READ)
        m = fs.getNode("node_" + k).mat()
        fs.release()
        assert m.dtype == cases[k].dtype and m.tobytes() == cases[k].tobytes(), k


def test_text_columns(built, tmp_path):
    from diasss_b200 import binding as B
    p = tmp_path / "ssh-170.txt"
    p.write_text("1.5\n\n2e-3\n  7 8\nabc\n-4.25")          # blank lines are skipped, the first number of a line counts,
    assert B.io_read_column(str(p)).tolist() == [1.5, 0.002, 7.0, 0.0, -4.25]      # an unparsable line yields 0 (util.cpp:137-146)
    vals = np.random.default_rng(5).normal(0, 100, 1001)
    p.write_text("".join("%.17g\n" % v for v in vals))
    assert np.array_equal(B.io_read_column(str(p)), vals)
    (tmp_path / "empty.txt").write_text("")
    assert len(B.io_read_column(str(tmp_path / "empty.txt"))) == 0


def make_raw_survey(n, rows, cols, seed, spread=0.7):
    """Raw f64 waterfall images + dead-reckoning data of n overlapping lines (the synthetic stand-in for test_data)."""
    from diasss_b200 import synth
    frames = synth.make_survey(n, rows, cols, seed=seed, spread=spread)
    g = np.random.default_rng(seed + 1)
    raws = []
    for f in frames:
        raw = np.abs((f["norm_img"].astype(np.float64) + 3.0) * 1.7e-3 * (1.0 + 0.02 * g.standard_normal((rows, cols))))
        raw[g.random((rows, cols)) < 1e-4] *= 40.0          # "buggy line" samples that GetFilteredMask stamps out
        raws.append(raw)
    return raws, [f["pose"] for f in frames], [np.full(rows, 10.0) for _ in frames], [f["g_range"] for f in frames]


def oracle_front_end(O, data, order):
    """src/diasss2.cpp:73-97 through the CPU oracle.  order = summation order of cv::mean (0 sequential, 1 library)."""
    n = len(data["imgs"])
    rows, cols = data["imgs"][0].shape
    ex = O.Extractor()
    frames, geos = [], []
    for k in range(n):
        raw = data["imgs"][k]
        kps, desc = ex(O.normalize_sss(raw, order))
        kps, desc, _ = O.mask_filter(kps, desc, O.filtered_mask(raw, order))
        gx, gy = O.geo_img(rows, cols, data["poses"][k], data["granges"][k])
        frames.append(O.Frame(k, rows, cols, kps, desc, gx, gy))
        geos.append((gx, gy))
    pairs, overlap, rows6 = [], [], []
    corres = [[] for _ in range(n)]
    for i in range(n):
        for j in range(i + 1, n):
            ov = O.compute_intersection(geos[i], geos[j])
            overlap.append(ov)
            if ov > np.float32(0.4):
                r = O.robust_matching(frames[i], frames[j])[0]
                pairs.append((i, j)); rows6.append(r)
                corres[i].append(r); corres[j].append(r[:, [1, 0, 4, 5, 2, 3]])       # FEAmatcher.cpp:35-45
    corres = [np.concatenate(c) if c else np.zeros((0, 6)) for c in corres]
    return dict(frames=frames, pairs=np.array(pairs, np.int32).reshape(-1, 2), overlap=np.array(overlap, np.float32), rows6=rows6, corres=corres)


def test_config1_test_demo_stand_in_cpu(oracle, tmp_path):
    from diasss_b200 import demo
    n, rows, cols = 5, 700, 600
    raws, poses, altts, granges = make_raw_survey(n, rows, cols, seed=55)
    paths = demo.write_survey(str(tmp_path), raws, poses, altts, granges)
    assert sorted(os.listdir(paths["image"])) == ["ssh-%d.xml" % (170 + k) for k in range(n)]
    data = demo.load_input_data(**paths)
    for k in range(n):                                       # the formats round-trip bit-exactly
        assert data["imgs"][k].tobytes() == raws[k].tobytes() and data["poses"][k].tobytes() == poses[k].tobytes()
        assert np.array_equal(data["altitudes"][k], altts[k]) and np.array_equal(data["granges"][k], granges[k])
        assert data["annos"][k].shape == (0, 7)
    res = oracle_front_end(oracle, data, order=0)
    assert all(300 < len(f.kps) <= 2012 for f in res["frames"])
    assert 0 < len(res["pairs"]) < n * (n - 1) // 2          # the overlap gate passes some pairs and rejects others
    assert sum(len(r) for r in res["rows6"]) > 200
    for (i, j), r in zip(res["pairs"], res["rows6"]):        # rows name the pair and keypoints of its two frames
        assert np.all(r[:, 0] == i) and np.all(r[:, 1] == j)
        ki = {(float(y), float(x)) for x, y in zip(res["frames"][i].kps["x"], res["frames"][i].kps["y"])}
        assert all((a, b) in ki for a, b in r[:, 2:4])
    assert sum(len(c) for c in res["corres"]) == 2 * sum(len(r) for r in res["rows6"])


def test_cpp_example_loads_folders_and_fails_loudly_without_gpu(built, tmp_path):
    """examples/test_demo_frontend (C ABI only) parses the reference's folders; without a CUDA device it stops at
    dsx_create with the library's message (no CPU fallback)."""
    import subprocess
    import torch
    from diasss_b200 import demo
    exe = os.path.join(os.path.dirname(HERE), "examples", "test_demo_frontend")
    assert os.path.exists(exe)
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by tests/test_gpu_demo.py")
    g = np.random.default_rng(2)
    rows, cols = 40, 32
    paths = demo.write_survey(str(tmp_path), [g.random((rows, cols)) for _ in range(2)], [g.normal(0, 1, (rows, 6)) for _ in range(2)],
                              [np.full(rows, 10.0)] * 2, [np.linspace(1, 20, cols // 2 + 1)] * 2)
    r = subprocess.run([exe, "--image", paths["image"], "--pose", paths["pose"], "--altitude", paths["altitude"], "--groundrange",
                        paths["groundrange"]], capture_output=True, text=True)
    assert r.returncode == 1 and r.stdout.count("image size: 40 32") == 2
    assert "no usable CUDA device" in r.stderr and "no CPU fallback" in r.stderr


def test_malformed_files_are_errors_not_crashes(built, tmp_path):
    from diasss_b200 import binding as B
    good = open(os.path.join(GOLD, "ssh-170_anno.xml")).read()
    cases = {
        "truncated.xml": good[:len(good) // 2],
        "no_rows.xml": good.replace("<rows>9</rows>", ""),
        "bad_dt.xml": good.replace("<dt>i</dt>", "<dt>q</dt>"),
        "three_channels.xml": good.replace("<dt>i</dt>", "<dt>3i</dt>"),
        "short_data.xml": good.replace("<rows>9</rows>", "<rows>900</rows>"),
        "not_a_matrix.xml": good.replace('type_id="opencv-matrix"', 'type_id="opencv-sparse-matrix"'),
        "empty.xml": "",
        "binary.xml": "\x00\x01\x02<anno_kps",
    }
    for name, text in cases.items():
        p = tmp_path / name
        p.write_text(text)
        with pytest.raises(B.DsxError):
            B.io_read_matrix(str(p), "anno_kps")
    with pytest.raises(B.DsxError):
        B.io_read_column(str(tmp_path / "missing.txt"))
    with pytest.raises(B.DsxError):
        B.io_write_matrix(str(tmp_path / "no_such_dir" / "x.xml"), "m", np.zeros((2, 2)))
