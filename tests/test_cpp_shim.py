"""The C++ host side (include/diasss_b200/shim.hpp): ORB_SLAM2::ORBextractor / Diasss::FEAmatcher used from a C++
program the way the reference uses them (tests/cpp/test_shim.cpp), compared byte for byte with the CPU oracle."""
import os
import struct
import subprocess

import numpy as np
import pytest

from tests._util import oracle_frame

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "test_shim")
KP = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"), ("octave", "<i4"), ("class_id", "<i4")])


def test_shim_builds_and_fails_loudly_without_gpu(built, tmp_path):
    """The shim compiles as C++11 against the C ABI; without a CUDA device the constructor throws (no CPU fallback)."""
    import torch
    assert os.path.exists(EXE)
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by test_shim_vs_oracle")
    inp = tmp_path / "in.bin"
    inp.write_bytes(struct.pack("<i", 0))
    r = subprocess.run([EXE, str(inp), str(tmp_path / "out.bin")], capture_output=True, text=True)
    assert r.returncode == 3
    assert "no usable CUDA device" in r.stderr and "no CPU fallback" in r.stderr


class _Reader:
    def __init__(self, b):
        self.b, self.o = b, 0

    def arr(self, dtype, n):
        dt = np.dtype(dtype)
        a = np.frombuffer(self.b, dt, n, self.o)
        self.o += dt.itemsize * n
        return a

    def i32(self):
        return int(self.arr("<i4", 1)[0])


@pytest.mark.gpu
def test_shim_vs_oracle(built, oracle, tmp_path):
    from diasss_b200 import synth
    O = oracle
    pair = synth.make_pair(rows=420, cols=360, seed=7, ids=(0, 1))
    ex = O.Extractor()
    of = [oracle_frame(O, f, ex) for f in pair]
    with open(tmp_path / "in.bin", "wb") as fh:
        fh.write(struct.pack("<i", 2))
        for f, o in zip(pair, of):
            fh.write(struct.pack("<iii", f["img_id"], f["rows"], f["cols"]))
            fh.write(np.ascontiguousarray(f["norm_img"], np.uint8).tobytes())
            fh.write(np.ascontiguousarray(f["mask"], np.uint8).tobytes())
            fh.write(np.ascontiguousarray(o.geo_x, np.float64).tobytes())
            fh.write(np.ascontiguousarray(o.geo_y, np.float64).tobytes())
    r = subprocess.run([EXE, str(tmp_path / "in.bin"), str(tmp_path / "out.bin")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    R = _Reader((tmp_path / "out.bin").read_bytes())
    for f, o in zip(pair, of):
        k_all, _ = ex(f["norm_img"])
        n_all, n = R.i32(), R.i32()
        assert n_all == len(k_all) and n == len(o.kps)
        assert R.arr(KP, n).tobytes() == o.kps.tobytes(), "Frame::kps"
        assert R.arr(np.uint8, n * 32).tobytes() == o.desc.tobytes(), "Frame::dst"
    rows6, si, ti, c1, c2 = O.robust_matching(of[0], of[1])
    K = R.i32()
    assert K == len(rows6) and K > 10
    src_rows = R.arr("<f8", 6 * K).reshape(K, 6)
    tgt_rows = R.arr("<f8", 6 * K).reshape(K, 6)
    assert src_rows.tobytes() == rows6.tobytes(), "Source.corres_kps"
    assert np.array_equal(tgt_rows, rows6[:, [1, 0, 4, 5, 2, 3]]), "Target.corres_kps (mirrored rows, FEAmatcher.cpp:41-44)"
    s1, s2 = O.geo_nn_search(of[0], of[1]), O.geo_nn_search(of[1], of[0])
    n0 = R.i32(); assert np.array_equal(R.arr("<i4", n0), s1["corres"]), "GeoNearNeighSearch dir 1"
    n1 = R.i32(); assert np.array_equal(R.arr("<i4", n1), s2["corres"]), "GeoNearNeighSearch dir 2"
    K2 = R.i32()
    assert K2 == K
    keys = R.arr(KP, 2 * K2).reshape(K2, 2)
    assert keys[:, 0].tobytes() == of[0].kps[si].tobytes() and keys[:, 1].tobytes() == of[1].kps[ti].tobytes(), "ConsistentCheck keys"
    dd = R.i32()
    assert dd == int(np.unpackbits(of[0].desc[0] ^ of[1].desc[0]).sum()), "DescriptorDistance"
    n2 = R.i32()
    assert n2 == len(of[0].kps)
    assert R.arr(KP, n2).tobytes() == of[0].kps.tobytes() and R.arr(np.uint8, n2 * 32).tobytes() == of[0].desc.tobytes(), "DetectFeatureB200"
    assert R.i32() == 0, "empty image must give no keypoints and an empty descriptor matrix"
    assert R.i32() == 6 and abs(float(R.arr("<f4", 1)[0]) - 1.2) < 1e-6
    sf = R.arr("<f4", 6)
    assert abs(sf[1] - 1.2) < 1e-6 and abs(sf[5] - 2.48832035) < 1e-6
