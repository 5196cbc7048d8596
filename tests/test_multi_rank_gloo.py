"""N>1 host-side logic on CPU: world_size-2 (and 3) `gloo` process groups exercise exactly what bench.py does across GPUs --
the 16-row band partition (b200pt_partition_* from the product library), padded per-rank band buffers, one gather to rank 0,
reassembly by global row -- with the CPU oracle standing in for the device renderer (it implements the same partition rule).
The assembled image must be bit-identical to the single-rank image (per-pixel RNG is keyed on global coordinates)."""
import os
import sys
import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import util

W, H, FRAMES, BAND = 48, 41, 2, 4


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import vpt_b200 as pt
        S = util.oracle_scene("cornell_box")
        cfg = util.oracle_config("cornell_box", MaxDepth=5)
        rows = pt.partition_rows(H, rank, world, BAND)                       # product's partition rule
        full = np.zeros((H, W, 4), np.float32)
        S.render(cfg, W, H, FRAMES, 1234, image=full, rank=rank, world=world, band_rows=BAND, nthreads=2)
        assert np.all(full[np.setdiff1d(np.arange(H), rows)] == 0)           # a rank touches only its own rows
        max_rows = max(pt.lib().b200pt_partition_local_row_count(H, r, world, BAND) for r in range(world))
        band = torch.zeros((max_rows, W, 4), dtype=torch.float32)
        band[:len(rows)] = torch.from_numpy(full[rows])                      # local-row order, padded (as get_hdr returns it)
        gathered = [torch.zeros_like(band) for _ in range(world)] if rank == 0 else None
        dist.gather(band, gathered, dst=0)                                   # the ONE collective of the data path
        if rank == 0:
            out = np.zeros((H, W, 4), np.float32)
            for r in range(world):
                rr = pt.partition_rows(H, r, world, BAND)
                out[rr] = gathered[r][:len(rr)].numpy()
            q.put(out)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_band_partition_gather_reassembles_bit_identical_image(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + world + (os.getpid() % 200)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs: p.start()
    out = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60); assert p.exitcode == 0
    S = util.oracle_scene("cornell_box")
    ref, _ = S.render(util.oracle_config("cornell_box", MaxDepth=5), W, H, FRAMES, 1234)
    assert np.array_equal(out, ref)
