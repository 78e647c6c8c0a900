"""Worker of tests/test_multi_gpu.py: run under torchrun, one rank per GPU, NCCL.

Checks, on every rank, that the N-GPU results are BIT-IDENTICAL to the 1-GPU results computed by the same rank:
  * top-k matching with the lo axis sharded + all_gather + merge       (mad/MaD.py:420-424)
  * threshold matching sharded + all-gather-v + (hi, lo) merge         (the parity contract)
  * a batch of maps, map i -> rank i mod G, tables collected           (mad/MaD.py:143-162)
Prints "MGPU-OK <world>" on rank 0 and exits 0; any mismatch raises."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, "oracle"), os.path.join(REPO, "tests")]


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    dist.init_process_group("nccl", device_id=dev)
    import synth
    import helpers as H
    from mad_b200 import pipeline as P
    from mad_b200 import parallel as par

    # ---- matching: identical inputs on every rank, ragged shard sizes
    n, m, k = 6007, 3001, 8
    lo = synth.synthetic_descriptors(n, 3)
    hi = synth.synthetic_descriptors(m, 4, noisy_copy_of=lo)
    hi[5] = 0
    s, e = par.shard_bounds(n, world)[rank]
    hs, ls_full, ls = P.DescriptorSet(hi), P.DescriptorSet(lo), P.DescriptorSet(np.ascontiguousarray(lo[s:e]))
    idx1, sc1 = P.match_topk(hs, ls_full, k)
    idx, sc = par.match_topk_sharded(hs, ls, k, s)
    assert torch.equal(idx, idx1) and torch.equal(sc, sc1), "sharded top-k differs from the 1-GPU result"
    # 2-D rank grid (hi blocks x lo shards), every factorisation of the world
    for gl in range(1, world + 1):
        if world % gl:
            continue
        grid = (world // gl, gl)
        bh, bl = par.grid_coords(rank, *grid)
        h0, h1 = par.shard_bounds(m, grid[0])[bh]
        l0, l1 = par.shard_bounds(n, grid[1])[bl]
        gi, gs = par.match_topk_grid_sets(P.DescriptorSet(np.ascontiguousarray(hi[h0:h1])), P.DescriptorSet(np.ascontiguousarray(lo[l0:l1])),
                                          k, m, l0, grid)
        assert torch.equal(gi, idx1) and torch.equal(gs, sc1), "rank grid %r top-k differs from the 1-GPU result" % (grid,)
    a = P.match_threshold(hs, ls_full, 0.55)
    b = par.match_threshold_sharded(hs, ls, 0.55, s)
    assert a[0].numel() > 1000
    for x, y in zip(a, b):
        assert torch.equal(x, y), "sharded threshold pair list differs from the 1-GPU result"

    # ---- batch of maps: map i -> rank i mod G
    names = ["tiny", "small", "pair_hi", "pair_lo", "tiny", "pair_hi", "small", "tiny", "pair_lo"]
    grids = [synth.dequantise_u16(H.golden(nm)["input_q"]) for nm in names]
    mine = par.assign_units(len(grids), rank, world)
    dsc_l, kp_l, ori_l = [], [], []
    for u in mine:
        sp, kp, ori, dsc = P.describe_struct(grids[u])
        dsc_l.append(dsc)
        kp_l.append(kp.table[:len(kp)])
        ori_l.append(ori.table[:len(ori)])
    like = (torch.empty((0, 1024), dtype=torch.int16, device=dev), torch.empty((0, 12), dtype=torch.int32, device=dev),
            torch.empty((0, 2), dtype=torch.int32, device=dev))          # for a rank that owns no map (world > maps)
    all_dsc = par.collect_units(dsc_l, len(grids), like=like[0])
    all_kp = par.collect_units(kp_l, len(grids), like=like[1])
    all_ori = par.collect_units(ori_l, len(grids), like=like[2])
    for u, g in enumerate(grids):
        sp, kp, ori, dsc = P.describe_struct(g)
        assert torch.equal(all_dsc[u], dsc) and torch.equal(all_kp[u], kp.table[:len(kp)]) and torch.equal(all_ori[u], ori.table[:len(ori)]), \
            "batch path: map %d differs from the 1-GPU tables" % u
        gd = H.golden(names[u])
        assert np.array_equal(H.crc_rows(all_dsc[u].cpu().numpy()), gd["dsc_crc32"])       # and from the reference's
    dist.barrier()
    if rank == 0:
        print("MGPU-OK %d" % world, flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
