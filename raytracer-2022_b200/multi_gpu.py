"""spp sharding across the GPUs of one box.

Samples of a pixel are i.i.d. and the reference already sums them (main.rs:144-151), so GPU g of
G renders ALL pixels for a disjoint range of global sample indices; the Philox stream of a path
is keyed by (pixel, global sample index), so the union over ranks is exactly the sample set one
GPU would render.  The only exchange is one SUM-reduce of the float accumulation buffer to
rank 0 (NCCL over NVLink on the GPU box; gloo in the CPU tests).
"""
import torch.distributed as dist


def shard_range(spp_begin, spp_end, rank, world):
    """Contiguous sample-index range of `rank`: [spp_begin + g*n/G, spp_begin + (g+1)*n/G)."""
    n = spp_end - spp_begin
    return spp_begin + (rank * n) // world, spp_begin + ((rank + 1) * n) // world


def step_sample_range(step, rank, world, spp_per_gpu):
    """Weak-scaling bench: step i gives every rank its own `spp_per_gpu` fresh sample indices."""
    b = (step * world + rank) * spp_per_gpu
    return b, b + spp_per_gpu


def reduce_to_root(accum, world, dst=0):
    """One SUM-reduce of the per-rank radiance sums (pre-gamma) to rank `dst`."""
    if world > 1:
        dist.reduce(accum, dst=dst, op=dist.ReduceOp.SUM)
    return accum


def render_sharded(gpu_scene, width, height, spp_begin, spp_end, accum, rank, world, max_depth=50, seed=2022, stream=None):
    """Render this rank's shard into the device tensor `accum` (float32, H x W x 3) and reduce."""
    b, e = shard_range(spp_begin, spp_end, rank, world)
    _, st = gpu_scene.render(width, height, b, e, max_depth=max_depth, seed=seed, accum_ptr=accum.data_ptr(), stream=stream)
    reduce_to_root(accum, world)
    return st
