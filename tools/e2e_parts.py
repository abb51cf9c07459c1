"""Where the end-to-end step goes: pure H2D copy, extraction from device memory, pipelined extraction from pinned host
memory, matching, row read-back (64-image survey, 8000 x 2000).  Usage (under gpurun): python tools/e2e_parts.py [chunk]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from diasss_b200 import synth, binding as B
from diasss_b200.frontend import FrontEnd
chunk = int(sys.argv[1]) if len(sys.argv) > 1 else 0
n, R, C = 64, 8000, 2000
dev = torch.device("cuda", 0)
tracks = synth.survey_tracks(n, R, C, seed=1234)
field = synth.seabed(2048, 1234, dev)
imgs = torch.empty(n, R, C, dtype=torch.uint8, device=dev); masks = torch.empty_like(imgs)
for i, t in enumerate(tracks):
    imgs[i], masks[i] = synth.render(field, t, device=dev)
h_imgs, h_masks = imgs.cpu().pin_memory(), masks.cpu().pin_memory()
fe = FrontEnd(h2d_chunk=chunk)
feats = fe.alloc_features(n)

def timed(fn, reps=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

stage = torch.empty_like(imgs)
print("H2D 1.024 GB, one copy          : %.2f ms" % timed(lambda: stage.copy_(h_imgs, non_blocking=True)))
def chunks(k):
    for i in range(0, n, k): stage[i:i + k].copy_(h_imgs[i:i + k], non_blocking=True)
for k in (1, 4, 8):
    print("H2D in chunks of %d images        : %.2f ms" % (k, timed(lambda: chunks(k))))
print("extract, device images (1 call) : %.2f ms" % timed(lambda: fe.ctx.detect_feature_batch_dev(imgs.data_ptr(), masks.data_ptr(), n, R, C, C, R * C, feats["c"])))
def dev_chunks(k):
    for i in range(0, n, k):
        fe.ctx.detect_feature_batch_dev(imgs[i:].data_ptr(), masks[i:].data_ptr(), min(k, n - i), R, C, C, R * C, fe.features_view(feats, i, min(k, n - i)))
for k in (1, 2, 4, 8, 16):
    print("extract, device images, chunks of %2d: %.2f ms" % (k, timed(lambda: dev_chunks(k))))
print("extract, pinned host (pipelined): %.2f ms" % timed(lambda: fe.ctx.detect_feature_batch(h_imgs.data_ptr(), h_masks.data_ptr(), n, R, C, C, R * C, feats["c"])))
