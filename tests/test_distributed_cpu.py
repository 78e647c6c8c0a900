"""N > 1 host logic on CPU: world size 2 over gloo (127.0.0.1).  The collectives and the sharding
arithmetic of mad_b200/parallel.py run for real; the per-shard top-k lists come from the oracle
(NumPy) and the merged result must equal the oracle's unsharded top-k -- the property the GPU
path relies on (the CUDA merge kernel itself is covered by the -m gpu tests)."""
import os
import socket
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _merge_numpy(idx_g, sc_g, k):
    """(score desc, index asc) k-way merge of [G, M, k] lists."""
    g, m, _ = idx_g.shape
    idx = np.full((m, k), -1, dtype=np.int32)
    sc = np.full((m, k), -np.inf)
    for r in range(m):
        cand = [(-(sc_g[s, r, q]), idx_g[s, r, q]) for s in range(g) for q in range(k) if idx_g[s, r, q] >= 0]
        cand.sort()
        for q, (ns, i) in enumerate(cand[:k]):
            idx[r, q], sc[r, q] = i, -ns
    return idx, sc


def _worker(rank, world, port, q):
    sys.path[:0] = [REPO, os.path.join(REPO, "oracle")]
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import mad_oracle as mo
        import synth
        from mad_b200 import parallel as par
        lo = synth.synthetic_descriptors(301, 3)
        hi = synth.synthetic_descriptors(40, 4, noisy_copy_of=lo)
        k = 8
        s, e = par.shard_bounds(len(lo), world)[rank]
        li, ls = mo.match_topk(hi, lo[s:e], k)
        li = np.where(li >= 0, li + s, -1).astype(np.int32)
        idx_g, sc_g = par.gather_topk(torch.from_numpy(li), torch.from_numpy(ls))
        mi, ms = _merge_numpy(idx_g.numpy(), sc_g.numpy(), k)
        oi, osc = mo.match_topk(hi, lo, k)
        ok_topk = bool(np.array_equal(ms, osc) and np.array_equal(mi[ms > osc[:, -1:]], oi[osc > osc[:, -1:]]))
        # variable-length gather: rank r contributes r + 2 rows tagged with its rank
        t = torch.full((rank + 2, 3), rank, dtype=torch.int32)
        parts = par.gather_varlen(t)
        ok_var = [tuple(p.shape) for p in parts] == [(r + 2, 3) for r in range(world)] and \
            all(int(p[0, 0]) == r for r, p in enumerate(parts))
        units = par.assign_units(7, rank, world)
        # threshold mode sharded on the lo axis: per-shard pair lists from the oracle, gathered + merged == unsharded
        lp, lsc = mo.match_threshold(hi, lo[s:e], 0.3)
        got = par.match_threshold_sharded(None, None, 0.3, s, local=(torch.from_numpy(lp[:, 0].copy()),
                                                                     torch.from_numpy(lp[:, 1].copy()), torch.from_numpy(lsc)))
        op, osc2 = mo.match_threshold(hi, lo, 0.3)
        ok_thr = bool(len(op) > 50 and np.array_equal(np.stack([got[0].numpy(), got[1].numpy()], 1), op)
                      and np.abs(got[2].numpy() - osc2).max() < 1e-14)
        # 2-D rank grid (hi blocks x lo shards): both 2-rank grids, local lists from the oracle, merged == unsharded
        ok_grid = True
        for grid in ((1, 2), (2, 1)):
            bh, bl = par.grid_coords(rank, *grid)
            hs, he = par.shard_bounds(len(hi), grid[0])[bh]
            ls_, le_ = par.shard_bounds(len(lo), grid[1])[bl]
            gi, gs = mo.match_topk(hi[hs:he], lo[ls_:le_], k)
            gi = np.where(gi >= 0, gi + ls_, -1).astype(np.int32)
            mi2, ms2 = par.match_topk_grid(None, None, k, len(hi), grid, local=(torch.from_numpy(gi), torch.from_numpy(gs)))
            mi2, ms2 = mi2.numpy(), ms2.numpy()
            ok_grid = ok_grid and bool(np.array_equal(ms2, osc) and np.array_equal(mi2[ms2 > osc[:, -1:]], oi[osc > osc[:, -1:]]))
        # batch path: per-map tables of 5 maps (map u -> rank u mod G) collected on every rank, int16 rows of differing counts
        tabs = [torch.full((u + 1, 4), u, dtype=torch.int16) for u in par.assign_units(5, rank, world)]
        allt = par.collect_units(tabs, 5, like=torch.empty((0, 4), dtype=torch.int16))
        ok_units = [tuple(t.shape) for t in allt] == [(u + 1, 4) for u in range(5)] and all(int(t[0, 0]) == u for u, t in enumerate(allt))
        q.put((rank, ok_topk, ok_var, units, ok_thr, ok_grid, ok_units))
    finally:
        dist.destroy_process_group()


def test_topk_grid_choice():
    sys.path.insert(0, REPO)
    from mad_b200 import parallel as par
    assert par.pick_topk_grid(100000, 100000, 1) == (1, 1)
    assert par.pick_topk_grid(100000, 100000, 2) == (1, 2)          # the reference axis is always cut
    assert par.pick_topk_grid(100000, 100000, 4) == (2, 2)
    assert par.pick_topk_grid(100000, 100000, 8) == (4, 2)
    for w in (2, 4, 8):
        gh, gl = par.pick_topk_grid(1000, 10 ** 6, w)
        assert gh * gl == w and gl >= 2
    assert [par.grid_coords(r, 2, 4) for r in (0, 3, 4, 7)] == [(0, 0), (0, 3), (1, 0), (1, 3)]


def test_shard_bounds_and_unit_assignment():
    sys.path.insert(0, REPO)
    from mad_b200 import parallel as par
    for n in (0, 1, 7, 100000):
        for g in (1, 2, 3, 8):
            b = par.shard_bounds(n, g)
            assert b[0][0] == 0 and b[-1][1] == n and all(x[1] == y[0] for x, y in zip(b, b[1:]))
            assert max(e - s for s, e in b) - min(e - s for s, e in b) <= 1
    owned = sorted(i for r in range(3) for i in par.assign_units(10, r, 3))
    assert owned == list(range(10))


def test_world_size_2_gloo_topk_merge_and_varlen_gather():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[0] for r in res] == [0, 1]
    assert all(r[1] for r in res), "sharded + merged top-k differs from the unsharded oracle"
    assert all(r[2] for r in res), "variable-length gather returned wrong shapes / contents"
    assert res[0][3] == [0, 2, 4, 6] and res[1][3] == [1, 3, 5]
    assert all(r[4] for r in res), "sharded threshold pair lists, gathered and merged, differ from the unsharded oracle"
    assert all(r[5] for r in res), "2-D rank grid top-k differs from the unsharded oracle"
    assert all(r[6] for r in res), "collect_units returned wrong per-map tables"
