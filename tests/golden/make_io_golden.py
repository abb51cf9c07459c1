"""Generates tests/golden/io/*: FileStorage XML files written by the REAL OpenCV (cv2.FileStorage, the writer behind the
reference's data files, util.cpp:90-116) and the arrays they hold, for tests/test_io_formats.py.
Run in a container that has cv2:  python tests/golden/make_io_golden.py"""
import os

import cv2
import numpy as np

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "io")
os.makedirs(HERE, exist_ok=True)
g = np.random.default_rng(20240601)
ct = np.abs(g.normal(2e-3, 1e-3, (23, 17)))
ct[0, 0], ct[1, 1], ct[2, 3], ct[5, 5], ct[6, 6] = 1.0, 0.0, 1.2345678e10, 7e-310, 123456789.125
pose = np.concatenate([g.normal(0, 0.01, (31, 3)), g.normal(0, 500, (31, 3))], axis=1)
anno = g.integers(-50, 9000, (9, 7)).astype(np.int32)
f32 = g.normal(0, 3, (4, 5)).astype(np.float32)
u8 = g.integers(0, 256, (6, 10)).astype(np.uint8)
special = np.array([[np.inf, -np.inf, 1.5], [0.0, -0.0, 1e300]])

fs = cv2.FileStorage(os.path.join(HERE, "ssh-170_img.xml"), cv2.FILE_STORAGE_WRITE); fs.write("ct_img", ct); fs.release()
fs = cv2.FileStorage(os.path.join(HERE, "ssh-170_pose.xml"), cv2.FILE_STORAGE_WRITE); fs.write("auv_pose", pose); fs.release()
fs = cv2.FileStorage(os.path.join(HERE, "ssh-170_anno.xml"), cv2.FILE_STORAGE_WRITE); fs.write("anno_kps", anno); fs.release()
fs = cv2.FileStorage(os.path.join(HERE, "several_nodes.xml"), cv2.FILE_STORAGE_WRITE)
fs.write("ct_img_backup", ct[:3]); fs.write("ct_img", ct[3:6]); fs.write("f32", f32); fs.write("u8", u8); fs.write("special", special)
fs.release()
np.savez(os.path.join(HERE, "arrays.npz"), ct=ct, pose=pose, anno=anno, f32=f32, u8=u8, special=special)
print("cv2", cv2.__version__, sorted(os.listdir(HERE)))
