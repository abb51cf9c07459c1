"""CPU, world_size 2 over gloo: the N > 1 plumbing (diasss_b200/shard.py) -- image/pair ownership, the all-gathered
feature layout, and the gather-v of correspondence rows re-assembled in (i,j) pair order -- must reproduce the
single-process result byte for byte."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from diasss_b200 import shard


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_feats(k, cap):
    """Deterministic fake feature block of image k."""
    g = np.random.default_rng(100 + k)
    n = int(g.integers(0, cap + 1))
    return dict(kps=torch.from_numpy(g.normal(size=(cap, 7)).astype(np.float32)),
                desc=torch.from_numpy(g.integers(0, 256, (cap, 32), dtype=np.uint8)),
                geo_xy=torch.from_numpy(g.normal(size=(cap, 2))), count=n)


def _fake_rows(p):
    """Deterministic fake correspondence rows of global pair p (some pairs have none)."""
    g = np.random.default_rng(5000 + p)
    k = int(g.integers(0, 7)) if p % 5 else 0
    return g.normal(size=(k, 6))


def _worker(rank, world, port, F, cap, q, counts=None):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        pairs = np.array([(i, j) for i in range(F) for j in range(i + 1, F)], np.int32)
        plan = shard.Plan(F, pairs, world, rank, counts=counts)
        dev = torch.device("cpu")
        local = dict(kps=torch.zeros(plan.n_local, cap, 7), desc=torch.zeros(plan.n_local, cap, 32, dtype=torch.uint8),
                     geo_xy=torch.zeros(plan.n_local, cap, 2, dtype=torch.float64), count=torch.zeros(plan.n_local, dtype=torch.int32))
        for s, k in enumerate(plan.my_images):
            f = _fake_feats(k, cap)
            local["kps"][s], local["desc"][s], local["geo_xy"][s], local["count"][s] = f["kps"], f["desc"], f["geo_xy"], f["count"]
        allf = dict(kps=torch.zeros(plan.n_slots, cap, 7), desc=torch.zeros(plan.n_slots, cap, 32, dtype=torch.uint8),
                    geo_xy=torch.zeros(plan.n_slots, cap, 2, dtype=torch.float64), count=torch.zeros(plan.n_slots, dtype=torch.int32))
        shard.all_gather_features(local, allf)
        ok = True
        for k in range(F):
            f, s = _fake_feats(k, cap), int(plan.slot_of(k))
            ok &= bool(torch.equal(allf["kps"][s], f["kps"]) and torch.equal(allf["desc"][s], f["desc"]) and
                       torch.equal(allf["geo_xy"][s], f["geo_xy"]) and int(allf["count"][s]) == f["count"])
        for s in set(range(plan.n_slots)) - {int(plan.slot_of(k)) for k in range(F)}:
            ok &= int(allf["count"][s]) == 0                      # padding slots stay empty
        mine = [_fake_rows(int(p)) for p in plan.my_pair_ids]
        res = dict(count=torch.tensor([len(m) for m in mine] + [0], dtype=torch.int32),
                   rows6=torch.from_numpy(np.concatenate(mine + [np.zeros((0, 6))])))
        cnt, rows = shard.gather_rows(plan, res, dev)
        if rank == 0:
            want = [_fake_rows(p) for p in range(len(pairs))]
            ok &= cnt.tolist() == [len(w) for w in want]
            ok &= rows.numpy().tobytes() == np.concatenate(want).tobytes()
        else:
            ok &= len(rows) == 0
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("F", [6, 7])
def test_two_ranks_equal_single_process(F):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, F, 9, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def test_two_ranks_unequal_shares():
    """Per-rank image shares (bench.py's end-to-end split by measured host-to-device bandwidth): 5 + 2 images."""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, 7, 9, q, [5, 2])) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def test_plan_shares():
    pairs = np.array([(i, j) for i in range(9) for j in range(i + 1, 9)], np.int32)
    counts = [2, 4, 3]
    owned, slots = [], []
    for r in range(3):
        pl = shard.Plan(9, pairs, 3, r, counts=counts)
        assert len(pl.my_images) == counts[r] and pl.n_local == 4 and pl.n_slots == 12
        assert [int(s) for s in pl.slot_of(pl.my_images)] == [r * 4 + i for i in range(counts[r])]      # local order = image order
        owned += pl.my_images
        slots += [int(s) for s in pl.slot_of(pl.my_images)]
    assert sorted(owned) == list(range(9)) and len(set(slots)) == 9
    # the default is the k mod world deal
    a, b = shard.Plan(9, pairs, 3, 1), shard.Plan(9, pairs, 3, 1, counts=[3, 3, 3])
    assert a.my_images == b.my_images and np.array_equal(a.my_pairs_slots, b.my_pairs_slots)
    with pytest.raises(ValueError):
        shard.Plan(9, pairs, 3, 0, counts=[4, 4, 4])


def test_plan_partition():
    pairs = np.array([(i, j) for i in range(8) for j in range(i + 1, 8)], np.int32)
    owned_i, owned_p = [], []
    for r in range(4):
        pl = shard.Plan(8, pairs, 4, r)
        owned_i += pl.my_images
        owned_p += pl.my_pair_ids.tolist()
        assert pl.n_slots == 8 and len(pl.my_pairs_slots) == len(pl.my_pair_ids)
    assert sorted(owned_i) == list(range(8)) and sorted(owned_p) == list(range(28))
    pl = shard.Plan(8, pairs, 1, 0)
    assert np.array_equal(pl.my_pairs_slots, pairs) and pl.slot_ids(np.arange(8) * 3).tolist() == (np.arange(8) * 3).tolist()
