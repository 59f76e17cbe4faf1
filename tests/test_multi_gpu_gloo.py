"""The N>1 path on CPU: two gloo ranks shard the spp range, render their shard (with the oracle
standing in for the device renderer -- this is a test) and SUM-reduce to rank 0."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import rtb200 as rt
    from raytracer_2022_b200 import multi_gpu as mg
    from oracle_lib import OracleScene

    w = h = 16
    spp = 6
    scene = rt.Scene(seed=2022).build_config("cornell_box")
    orc = OracleScene(scene)
    b, e = mg.shard_range(0, spp, rank, world)
    shard, _ = orc.render(w, h, b, e, seed=3, threads=1)
    accum = torch.from_numpy(shard.astype(np.float32))
    mg.reduce_to_root(accum, world)
    if rank == 0:
        full, _ = orc.render(w, h, 0, spp, seed=3, threads=1)
        q.put((accum.numpy().copy(), full))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_spp_sharding_sums_to_the_full_render():
    from raytracer_2022_b200 import multi_gpu as mg
    assert [mg.shard_range(0, 10, r, 4) for r in range(4)] == [(0, 2), (2, 5), (5, 7), (7, 10)]
    assert mg.step_sample_range(3, 1, 2, 8) == (56, 64)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29000 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got, full = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert np.allclose(got, full, rtol=1e-5, atol=1e-5)
