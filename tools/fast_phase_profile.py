#!/usr/bin/env python
"""Where fast_cells_kernel's warps spend their clocks, phase by phase (needs the DSX_FAST_PROFILE variant:
tools/build_variant.sh prof -DDSX_FAST_PROFILE; DSX_LIB=diasss_b200/variants/libdiasss_b200_prof.so python tools/fast_phase_profile.py)."""
import ctypes as C
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch                                          # noqa: E402
from diasss_b200 import binding as B, synth           # noqa: E402
from diasss_b200.frontend import FrontEnd             # noqa: E402

n, rows, cols = 8, 8000, 2000
dev = torch.device("cuda", 0)
field = synth.seabed(2048, 1234, dev)
tracks = synth.survey_tracks(64, rows, cols, seed=1234)[:n]
imgs = torch.empty(n, rows, cols, dtype=torch.uint8, device=dev)
masks = torch.empty_like(imgs)
for k in range(n):
    imgs[k], masks[k] = synth.render(field, tracks[k], device=dev)
fe = FrontEnd()
feats = fe.alloc_features(n)
out = np.zeros(16, np.uint64)
for it in range(3):
    B.lib().dsx_debug_fast_profile(C.c_void_p(out.ctypes.data), 1)
    fe.ctx.detect_feature_batch_dev(imgs.data_ptr(), masks.data_ptr(), n, rows, cols, cols, rows * cols, feats["c"])
    torch.cuda.synchronize()
rc = B.lib().dsx_debug_fast_profile(C.c_void_p(out.ctypes.data), 0)
names = ["staging issue / own wait", "barrier after staging", "shifted copies", "barrier after copies", "scoring", "barrier after scoring",
         "per-cell listing"]
tot = float(sum(out[0:14:2]))
print(json.dumps(dict(tool="fast_phase_profile", rc=rc, images=n, phases=[dict(phase=names[p], warp_clocks=int(out[2 * p]), warps=int(out[2 * p + 1]),
      mean_clocks=float(out[2 * p]) / max(int(out[2 * p + 1]), 1), share=float(out[2 * p]) / tot) for p in range(7)])))
