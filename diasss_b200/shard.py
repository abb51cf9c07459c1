"""Multi-GPU partitioning of the front end (one process per GPU, torch.distributed).

The path shards into two independent-unit phases with one exchange step between them:

  extraction  image k runs on rank k mod G                              (no communication)
  exchange    all-gather of the per-image feature blocks                (NCCL all_gather_into_tensor over NVLink)
  matching    pair p (in the reference's i<j loop order) runs on rank p mod G
  collection  gather-v of the per-pair correspondence rows to rank 0, re-assembled in (i,j) pair order so that
              Frame::corres_kps row order equals the single-GPU / reference order (optimizer.cpp:222-231 depends on it)

Everything here is index arithmetic and collectives on torch tensors -- it runs unchanged on CPU tensors with the
gloo backend, which is how tests/test_shard_gloo.py covers the N > 1 path without a GPU.
"""
import numpy as np
import torch
import torch.distributed as dist


class Plan:
    """Who owns which image and which pair; where an image lives in the all-gathered feature block."""

    def __init__(self, n_images, pairs, world, rank):
        self.F, self.world, self.rank = int(n_images), int(world), int(rank)
        self.pairs = np.ascontiguousarray(pairs, np.int32).reshape(-1, 2)
        self.n_local = (self.F + world - 1) // world           # feature slots per rank (equal on every rank)
        self.n_slots = self.n_local * world
        self.my_images = [k for k in range(self.F) if k % world == rank]
        self.pair_owner = np.arange(len(self.pairs)) % world
        self.my_pair_ids = np.nonzero(self.pair_owner == rank)[0]
        self.my_pairs = self.pairs[self.my_pair_ids]
        self.my_pairs_slots = self.slot_of(self.my_pairs)
        self.max_pairs_local = (len(self.pairs) + world - 1) // world

    def slot_of(self, k):
        """Index of image k in the all-gathered block: rank-major, then local slot."""
        k = np.asarray(k)
        return ((k % self.world) * self.n_local + k // self.world).astype(np.int32)

    def _per_slot(self, per_image, fill):
        per_image = np.asarray(per_image)
        out = np.full((self.n_slots,) + per_image.shape[1:], fill, per_image.dtype)
        out[self.slot_of(np.arange(self.F))] = per_image
        return out

    def slot_ids(self, img_ids):
        return self._per_slot(np.asarray(img_ids, np.int32), 0)

    def slot_rows(self, bboxes):
        return self._per_slot(np.asarray(bboxes, np.float64), 0.0)


def all_gather_features(local, gathered):
    """local[key]: [n_local, ...] on every rank  ->  gathered[key]: [world * n_local, ...] (rank-major)."""
    for key in ("kps", "desc", "geo_xy", "count"):
        dist.all_gather_into_tensor(gathered[key], local[key])


def gather_rows(plan, res, dev):
    """res: this rank's match output (count[P_local], rows6[k_local, 6], pair order = plan.my_pair_ids).
    Returns on rank 0: (count per pair in global pair order [P], rows6 [K, 6] in global pair order); elsewhere
    (empty, empty)."""
    W, Pmax = plan.world, plan.max_pairs_local
    cnt_local = torch.zeros(Pmax, dtype=torch.int32, device=dev)
    n_mine = len(plan.my_pair_ids)
    cnt_local[:n_mine] = res["count"][:n_mine]
    cnt_all = torch.empty(W * Pmax, dtype=torch.int32, device=dev)
    dist.all_gather_into_tensor(cnt_all, cnt_local)
    cnt_all = cnt_all.view(W, Pmax).to(torch.int64)
    k_rank = cnt_all.sum(1)
    kmax = int(k_rank.max().item())
    rows = torch.zeros(max(kmax, 1), 6, dtype=torch.float64, device=dev)
    k_mine = int(res["rows6"].shape[0])
    rows[:k_mine] = res["rows6"]
    if plan.rank == 0:
        bufs = [torch.empty_like(rows) for _ in range(W)]
        dist.gather(rows, bufs, dst=0)
    else:
        dist.gather(rows, None, dst=0)
        return torch.empty(0, dtype=torch.int32, device=dev), torch.empty(0, 6, dtype=torch.float64, device=dev)
    P = len(plan.pairs)
    p = torch.arange(P, device=dev)
    r_of, li_of = p % W, p // W
    cg = cnt_all[r_of, li_of]                                   # counts in global pair order
    dst_off = torch.cumsum(cg, 0) - cg
    src_off = torch.cumsum(cnt_all, 1) - cnt_all                # per rank, exclusive over its local pairs
    K = int(cg.sum().item())
    out = torch.empty(K, 6, dtype=torch.float64, device=dev)
    for r in range(W):
        kr = int(k_rank[r].item())
        if kr == 0:
            continue
        li = torch.repeat_interleave(torch.arange(Pmax, device=dev), cnt_all[r])            # local pair of each local row
        within = torch.arange(kr, device=dev) - src_off[r, li]
        gp = li * W + r                                                                   # global pair id
        out[dst_off[gp] + within] = bufs[r][:kr]
    return cg.to(torch.int32), out
