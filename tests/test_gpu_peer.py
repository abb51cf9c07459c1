"""GPU tests of the multi-GPU row collection over peer memory (dsx_peer_*, diasss_b200/csrc/peer.cu): the rows rank 0
ends up with must be byte-identical to what one context produces for the whole pair list ("N-GPU output
must be byte-identical to 1-GPU output"), in (i,j) pair order.

  * logical ranks inside one process (dsx_peer_connect_local): several contexts on their own streams share one GPU and
    exchange through the same kernels and flags as real ranks -- runs on a single-GPU box;
  * real ranks (one process per GPU, CUDA IPC handles exchanged with torch.distributed/NCCL): needs >= 2 GPUs.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _survey(n, rows, cols, seed):
    import torch
    from diasss_b200 import binding as B, synth
    frames = synth.make_survey(n, rows, cols, seed=seed)
    imgs = torch.from_numpy(np.stack([f["norm_img"] for f in frames])).cuda()
    masks = torch.from_numpy(np.stack([f["mask"] for f in frames])).cuda()
    models = [B.geo_model_build(f["pose"], rows, cols, f["g_range"]) for f in frames]
    rowtabs = torch.from_numpy(np.stack([m[0] for m in models])).cuda()
    granges = torch.from_numpy(np.stack([f["g_range"] for f in frames])).cuda()
    bboxes = np.stack([m[1] for m in models])
    ids = [f["img_id"] for f in frames]
    pairs = np.array([(i, j) for i in range(n) for j in range(i + 1, n)], np.int32)
    return imgs, masks, rowtabs, granges, bboxes, ids, pairs


@pytest.mark.parametrize("world,order", [(1, (0,)), (3, (2, 1, 0)), (4, (1, 3, 0, 2))])
def test_logical_ranks_equal_single_context(built, world, order):
    import torch
    from diasss_b200 import binding as B, shard
    from diasss_b200.frontend import FrontEnd
    n, rows, cols = 6, 400, 360
    imgs, masks, rowtabs, granges, bboxes, ids, pairs = _survey(n, rows, cols, seed=29)
    P = len(pairs)
    fe0 = FrontEnd(max_batch=3)
    ref = fe0.process_survey(imgs, masks, rowtabs, granges, ids, bboxes, pairs)
    want_cnt = ref["count"].cpu().numpy()[:P].copy()
    want_off = ref["offset"].cpu().numpy().copy()
    want_rows = ref["rows6"].cpu().numpy().copy()
    feats = ref["feats"]
    assert want_off[-1] > 100
    torch.cuda.synchronize()

    streams = [torch.cuda.Stream() for _ in range(world)]
    fes = [FrontEnd(stream=s.cuda_stream) for s in streams]
    plans = [shard.Plan(n, pairs, world, r) for r in range(world)]
    rpp = 2 * fe0.ctx.cap
    peers = [B.Peer(fes[r].ctx, r, world, P, P * rpp) for r in range(world)]
    try:
        for r in range(world):
            for q in range(world):
                if q != r:
                    peers[r].connect_local(q, peers[q])
            # (first use allocates the context's matcher scratch; cudaMalloc may serialise against a spinning kernel)
            fes[r].match_pairs(feats, ids, [rows] * n, bboxes, plans[r].my_pairs if len(plans[r].my_pairs) else pairs[:1])
        torch.cuda.synchronize()
        for seq in (1, 2, 3):                   # both parity halves, and a reuse of the first
            for r in order:                     # launch order != rank order: later ranks wait on the device for earlier ones
                pl = plans[r]
                # (every logical rank sees all features in image order: slot = image index)
                peers[r].match_pairs(feats["c"], ids, [rows] * n, bboxes, pl.my_pairs, pl.pair_begin[r], seq)
            c, o, rws = peers[0].collect(seq)
            cnt = shard._dev_tensor(c, (P,), "<i4", imgs.device, peers[0])
            off = shard._dev_tensor(o, (P + 1,), "<i4", imgs.device, peers[0])
            rows6 = shard._dev_tensor(rws, (P * rpp, 6), "<f8", imgs.device, peers[0])
            with torch.cuda.stream(streams[0]):
                got_cnt, got_off = cnt.cpu().numpy(), off.cpu().numpy()
                got_rows = rows6[:int(got_off[-1])].cpu().numpy()
            for f in fes:
                f.ctx.check_error()
            assert np.array_equal(got_cnt, want_cnt), "seq %d" % seq
            assert np.array_equal(got_off, want_off), "seq %d" % seq
            assert got_rows.tobytes() == want_rows.tobytes(), "seq %d" % seq
    finally:
        torch.cuda.synchronize()
        for p in peers:
            p.close()
        for f in fes:
            f.ctx.close()
        fe0.ctx.close()


_WORKER = r"""
import os, sys
sys.path.insert(0, %(root)r)
import numpy as np, torch, torch.distributed as dist
from diasss_b200 import binding as B, shard, synth
from diasss_b200.frontend import FrontEnd
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
n, rows, cols = 7, 400, 360
frames = synth.make_survey(n, rows, cols, seed=31)
models = [B.geo_model_build(f["pose"], rows, cols, f["g_range"]) for f in frames]
bboxes = np.stack([m[1] for m in models]); ids = [f["img_id"] for f in frames]
pairs = np.array([(i, j) for i in range(n) for j in range(i + 1, n)], np.int32)
plan = shard.Plan(n, pairs, world, rank)
fe = FrontEnd(device=rank)
def dev_of(keys):
    return torch.from_numpy(np.stack(keys)).to(dev)
mine = plan.my_images
imgs = dev_of([frames[k]["norm_img"] for k in mine]); masks = dev_of([frames[k]["mask"] for k in mine])
rowtabs = dev_of([models[k][0] for k in mine]); granges = dev_of([frames[k]["g_range"] for k in mine])
local, allf = fe.alloc_features(plan.n_local), fe.alloc_features(plan.n_slots)
col = shard.PeerCollector(fe, plan, 2 * fe.ctx.cap, dev)
out = None
for step in range(3):
    fe.ctx.detect_feature_batch_dev(imgs.data_ptr(), masks.data_ptr(), len(mine), rows, cols, cols, rows * cols, local["c"])
    fe.ctx.georef_batch_dev(local["c"], rowtabs.data_ptr(), granges.data_ptr(), rows, cols, granges.shape[1])
    shard.all_gather_features(local, allf)
    seq = col.push(allf, plan.slot_ids(ids), [rows] * plan.n_slots, plan.slot_bboxes(bboxes))
    if rank == 0:
        cnt, off, rows6 = col.collect(seq)
        k = int(off[-1].item())
        out = (cnt.cpu().numpy().copy(), off.cpu().numpy().copy(), rows6[:k].cpu().numpy().copy())
    fe.ctx.check_error()
dist.barrier()
if rank == 0:
    # the same survey on this GPU alone
    fe1 = FrontEnd(device=0, max_batch=4)
    f = lambda key: torch.from_numpy(np.stack([fr[key] for fr in frames])).to(dev)
    ref = fe1.process_survey(f("norm_img"), f("mask"), torch.from_numpy(np.stack([m[0] for m in models])).to(dev), f("g_range"), ids, bboxes, pairs)
    assert np.array_equal(out[0], ref["count"].cpu().numpy()[:len(pairs)])
    assert np.array_equal(out[1], ref["offset"].cpu().numpy())
    assert out[2].tobytes() == ref["rows6"].cpu().numpy().tobytes()
    assert out[1][-1] > 100
    print("PEER_OK world=%%d rows=%%d" %% (world, out[1][-1]))
col.close()
dist.destroy_process_group()
"""


def test_real_ranks_over_ipc(built, tmp_path):
    import torch
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    world = min(ngpu, 4)
    script = tmp_path / "peer_worker.py"
    script.write_text(_WORKER % dict(root=ROOT))
    env = dict(os.environ, NCCL_DEBUG="WARN")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr",
                        "127.0.0.1", "--master-port", "29731", str(script)], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0 and "PEER_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
