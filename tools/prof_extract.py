"""Small driver for ncu: extraction of a few 8000x2000 images (device resident), optional matching."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from diasss_b200 import synth, binding as B
from diasss_b200.frontend import FrontEnd
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
do_match = len(sys.argv) > 3 and sys.argv[3] == "match"
R, C = 8000, 2000
dev = torch.device("cuda", 0)
tracks = synth.survey_tracks(64, R, C, seed=1234)[:n]
field = synth.seabed(2048, 1234, dev)
imgs = torch.empty(n, R, C, dtype=torch.uint8, device=dev); masks = torch.empty_like(imgs)
for i, t in enumerate(tracks):
    imgs[i], masks[i] = synth.render(field, t, device=dev)
fe = FrontEnd(max_batch=n)
feats = fe.alloc_features(n)
models = [B.geo_model_build(t["pose"], R, C, t["g_range"]) for t in tracks]
rowtabs = torch.from_numpy(np.stack([m[0] for m in models])).to(dev)
granges = torch.from_numpy(np.stack([t["g_range"] for t in tracks])).to(dev)
pairs = np.array([(i, j) for i in range(n) for j in range(i + 1, n)], np.int32)
fe.ctx.timing_enable(True)
for r in range(reps):
    fe.ctx.detect_feature_batch_dev(imgs.data_ptr(), masks.data_ptr(), n, R, C, C, R * C, feats["c"])
    if do_match:
        fe.ctx.georef_batch_dev(feats["c"], rowtabs.data_ptr(), granges.data_ptr(), R, C, granges.shape[1])
        fe.match_pairs(feats, [t["img_id"] for t in tracks], [R] * n, np.stack([m[1] for m in models]), pairs)
    torch.cuda.synchronize()
    print({k: round(v[0], 3) for k, v in fe.ctx.timing_read().items()})
