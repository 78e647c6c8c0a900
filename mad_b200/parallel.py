"""Multi-GPU plumbing of the two paths that shard; one process per GPU,
``torch.distributed`` (NCCL on the GPUs; the same code runs over gloo on CPU tensors in the tests).

* batches of maps / conformational snapshots (``MaD.py:143-162,178-189`` loops): independent units,
  map i -> rank i mod G, no data-path collective; ``gather_varlen`` collects per-map result tables;
* all-pairs matching (``MaD.py:420-424``): the lo (reference) axis is cut into G contiguous shards,
  every rank holds all hi rows and computes its local top-k with GLOBAL lo indices; one
  ``all_gather`` of the [M, k] lists and a k-way merge with the (score desc, index asc) rule give
  a result bit-identical to one GPU.
"""
import torch
import torch.distributed as dist


def shard_bounds(n, world):
    """Contiguous near-equal split of range(n): [(start, end)] * world (first n % world shards one longer)."""
    base, extra = divmod(int(n), int(world))
    out, s = [], 0
    for r in range(world):
        e = s + base + (1 if r < extra else 0)
        out.append((s, e))
        s = e
    return out


def assign_units(n_units, rank, world):
    """Round-robin ownership of independent units (maps, snapshots)."""
    return list(range(int(rank), int(n_units), int(world)))


def _world(group):
    return dist.get_world_size(group) if dist.is_initialized() else 1


def gather_topk(idx_local, score_local, group=None):
    """all_gather of per-shard top-k lists: ([M, k] int32, [M, k] float64) -> ([G, M, k], [G, M, k])."""
    g = _world(group)
    if g == 1:
        return idx_local[None], score_local[None]
    idx_all = [torch.empty_like(idx_local) for _ in range(g)]
    sc_all = [torch.empty_like(score_local) for _ in range(g)]
    dist.all_gather(idx_all, idx_local.contiguous(), group=group)
    dist.all_gather(sc_all, score_local.contiguous(), group=group)
    return torch.stack(idx_all), torch.stack(sc_all)


def gather_varlen(t, group=None):
    """all_gather of tensors whose first dimension differs per rank (pair lists, descriptor tables):
    counts first, then one padded all_gather; returns the list of per-rank tensors."""
    g = _world(group)
    if g == 1:
        return [t]
    n = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
    counts = [torch.zeros_like(n) for _ in range(g)]
    dist.all_gather(counts, n, group=group)
    counts = [int(c.item()) for c in counts]
    cap = max(max(counts), 1)
    pad = torch.zeros((cap,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[: t.shape[0]] = t
    bufs = [torch.empty_like(pad) for _ in range(g)]
    dist.all_gather(bufs, pad, group=group)
    return [b[:c] for b, c in zip(bufs, counts)]


def match_topk_sharded(hi_set, lo_shard_set, k, lo_index_base, group=None, impl=None):
    """Per hi row the k best rows of the WHOLE lo set, of which this rank holds the shard starting at
    global row ``lo_index_base``.  Every rank returns the same merged ([M, k] idx, [M, k] score)."""
    from . import pipeline as P
    idx, sc = P.match_topk(hi_set, lo_shard_set, k, lo_index_base=lo_index_base, impl=impl)
    idx_g, sc_g = gather_topk(idx, sc, group)
    if idx_g.shape[0] == 1:
        return idx, sc
    return P.topk_merge(idx_g, sc_g)
