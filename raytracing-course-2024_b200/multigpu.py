"""One-process-per-GPU orchestration of a sharded frame.

The path shards by SAMPLE: rank r of R renders samples [r*S/R, (r+1)*S/R) of every pixel into a local FP32
accumulator (rgb sums + count) with rt_render_accumulate_device; the counter-based RNG is keyed by the absolute
sample index, so the union over ranks is exactly the single-GPU sample set.  The only communication is ONE
reduce (sum) of the W*H*4-float accumulator to rank 0 (NCCL over NVLink on GPUs, gloo in the CPU tests), after
which rank 0 resolves (color_to_pixel) and copies W*H*3 bytes to the host.  torch.distributed is plumbing only."""
from __future__ import annotations


def shard_range(samples: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, near-equal split of [0, samples) -- empty ranges are possible when samples < world."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, rem = divmod(samples, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def reduce_to_root(accum, root: int = 0):
    """Sum the per-rank accumulators into `root` (in place).  No-op without an initialised process group."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(accum, dst=root, op=dist.ReduceOp.SUM)
    return accum


def render_frame_sharded(scene, accum, rgb_dev, *, seed: int, samples: int, rank: int, world: int, stream_ptr: int, resolve, mark=None, **render_kw):
    """One frame of the sharded path: zero the accumulator -> local sample shard (rt_render_accumulate_device) -> reduce to rank 0 ->
    (rank 0) resolve.  accum / rgb_dev are torch tensors on the rank's device (any object with zero_() / data_ptr()); `scene` needs
    render_accumulate_device(accum_ptr, stream_ptr, **kw); `resolve(accum_ptr, rgb_ptr, stream_ptr)` is color_to_pixel on the device.
    mark(k), if given, is called at the phase boundaries k = 0 (start), 1 (zeroed), 2 (rendered), 3 (reduced), 4 (resolved) -- bench.py
    records CUDA events there.  This is the ONE implementation of the per-rank frame: bench.py's timed loop, its end-to-end leg and the
    gloo test all call it."""
    lo, hi = shard_range(samples, rank, world)
    if mark:
        mark(0)
    accum.zero_()
    if mark:
        mark(1)
    if hi > lo:
        scene.render_accumulate_device(accum.data_ptr(), stream_ptr, seed=seed, sample_begin=lo, sample_end=hi, **render_kw)
    if mark:
        mark(2)
    reduce_to_root(accum, 0)
    if mark:
        mark(3)
    if rank == 0:
        resolve(accum.data_ptr(), rgb_dev.data_ptr(), stream_ptr)
    if mark:
        mark(4)
