"""GPU parity of test_demo's whole front end (diasss_b200/demo.py) from the reference's on-disk formats: frames built on
the device (GetNormalizeSSS, GetFilteredMask, geo model, DetectFeature), the overlap-gated i<j loop and RobustMatching
against the CPU oracle's restatement of src/diasss2.cpp:73-97 -- keypoints, descriptors, pair list, overlap ratios and
every frame's corres_kps byte for byte (cv::mean summed in the library's documented order, see test_gpu_frameprep.py)."""
import numpy as np
import pytest

from tests.test_io_formats import make_raw_survey, oracle_front_end

pytestmark = pytest.mark.gpu


def test_demo_front_end_vs_oracle(oracle, tmp_path):
    from diasss_b200 import demo
    n, rows, cols = 5, 700, 600
    raws, poses, altts, granges = make_raw_survey(n, rows, cols, seed=55)
    paths = demo.write_survey(str(tmp_path), raws, poses, altts, granges)
    data = demo.load_input_data(**paths)
    want = oracle_front_end(oracle, data, order=1)
    got = demo.front_end(data)
    assert np.array_equal(got["pairs"], want["pairs"]) and 0 < len(want["pairs"]) < n * (n - 1) // 2
    assert got["overlap"].tobytes() == want["overlap"].tobytes()
    for k in range(n):
        assert got["frames"][k]["kps"].tobytes() == want["frames"][k].kps.tobytes(), "keypoints of frame %d" % k
        assert np.array_equal(got["frames"][k]["desc"], want["frames"][k].desc)
        assert np.array_equal(got["frames"][k]["norm_img"], oracle.normalize_sss(raws[k], 1))
        assert np.array_equal(got["frames"][k]["flt_mask"], oracle.filtered_mask(raws[k], 1))
        assert got["frames"][k]["corres_kps"].tobytes() == want["corres"][k].tobytes(), "corres_kps of frame %d" % k
    assert got["rows6"].tobytes() == np.concatenate(want["rows6"]).tobytes() and len(got["rows6"]) > 200


def test_demo_cli(built, tmp_path):
    from diasss_b200 import demo
    raws, poses, altts, granges = make_raw_survey(3, 500, 420, seed=77, spread=0.3)
    paths = demo.write_survey(str(tmp_path / "data"), raws, poses, altts, granges)
    out = str(tmp_path / "res.npz")
    demo.main(["--image", paths["image"], "--pose", paths["pose"], "--altitude", paths["altitude"], "--groundrange",
               paths["groundrange"], "--annotation", paths["annotation"], "--out", out])
    r = np.load(out)
    assert len(r["pairs"]) >= 1 and r["rows6"].shape[1] == 6 and r["f0_desc"].shape[1] == 32


def test_cpp_example_equals_python_demo(built, tmp_path):
    """examples/test_demo_frontend.cpp (C ABI only, the reference's command line) == diasss_b200.demo on the same folders."""
    import os
    import subprocess
    from diasss_b200 import demo
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    raws, poses, altts, granges = make_raw_survey(4, 600, 500, seed=91, spread=0.6)
    paths = demo.write_survey(str(tmp_path / "data"), raws, poses, altts, granges)
    out = str(tmp_path / "corres.txt")
    r = subprocess.run([os.path.join(root, "examples", "test_demo_frontend"), "--image", paths["image"], "--pose", paths["pose"],
                        "--altitude", paths["altitude"], "--groundrange", paths["groundrange"], "--annotation", paths["annotation"],
                        "--out", out], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    want = demo.front_end(demo.load_input_data(**paths))
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("The OVERLAPPING RATE")]
    assert len(lines) == 6 and [float(ln.split(":")[1].split()[0]) for ln in lines] == [float("%g" % v) for v in want["overlap"]]
    pairs, rows = [], []
    for ln in open(out):
        t = ln.split()
        if t[0] == "pair":
            pairs.append((int(t[1]), int(t[2])))
        else:
            rows.append([float(v) for v in t])
    assert np.array_equal(np.array(pairs, np.int32).reshape(-1, 2), want["pairs"])
    assert np.array(rows, np.float64).reshape(-1, 6).tobytes() == want["rows6"].tobytes() and len(rows) > 100
