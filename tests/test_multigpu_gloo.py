"""World-size-2 gloo test of the N>1 host logic (no GPU): sample shards partition [0, S) exactly, and the reduce
of per-rank accumulators to rank 0 equals the sum -- the only collective of the path."""
import os
import subprocess
import sys

import numpy as np

from conftest import ROOT


def test_shard_ranges_partition_samples(rt):
    from raytracing_course_2024_b200 import multigpu
    for samples in (1, 2, 7, 64, 1000, 1024):
        for world in (1, 2, 3, 4, 8):
            r = [multigpu.shard_range(samples, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == samples
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1])
import numpy as np, torch, torch.distributed as dist
import rtb200
from raytracing_course_2024_b200 import multigpu
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + sys.argv[2], rank=int(sys.argv[3]), world_size=2)
rank = dist.get_rank()
W, H, S = 16, 8, 10
lo, hi = multigpu.shard_range(S, rank, 2)
g = torch.Generator().manual_seed(0)
per_sample = torch.rand(S, H * W, 3, generator=g, dtype=torch.float32)          # same on both ranks
acc = torch.zeros(H * W, 4)
acc[:, :3] = per_sample[lo:hi].sum(0); acc[:, 3] = hi - lo
multigpu.reduce_to_root(acc, 0)
if rank == 0:
    full = per_sample.sum(0)
    assert torch.allclose(acc[:, :3], full, rtol=1e-6, atol=1e-6), "reduce != sum of shards"
    assert torch.all(acc[:, 3] == S)
# the whole per-rank frame (the function bench.py times) with a stand-in scene that "renders" the shard on the CPU
class FakeScene:
    def render_accumulate_device(self, accum_ptr, stream_ptr, seed=0, sample_begin=0, sample_end=0, **kw):
        assert accum_ptr == acc2.data_ptr() and kw == {"kernel_variant": 7}
        acc2[:, :3] += per_sample[sample_begin:sample_end].sum(0); acc2[:, 3] += sample_end - sample_begin
acc2 = torch.full((H * W, 4), 99.0)                                              # stale contents: the frame must zero them
rgb = torch.zeros(H * W, 3)
marks = []
def resolve(a_ptr, r_ptr, s_ptr):
    assert a_ptr == acc2.data_ptr() and r_ptr == rgb.data_ptr()
    rgb[:] = acc2[:, :3] / acc2[:, 3:4]
multigpu.render_frame_sharded(FakeScene(), acc2, rgb, seed=1, samples=S, rank=rank, world=2, stream_ptr=0, resolve=resolve, mark=marks.append, kernel_variant=7)
assert marks == [0, 1, 2, 3, 4]
if rank == 0:
    assert torch.allclose(rgb, per_sample.mean(0), rtol=1e-5, atol=1e-6) and torch.all(acc2[:, 3] == S)
    print("OK")
else:
    assert not rgb.any()                                                         # only the root resolves
dist.barrier()
dist.destroy_process_group()
'''


def test_two_rank_gloo_reduce(rt, tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True) for r in range(2)]
    outs = [p.communicate(timeout=180) for p in procs]
    for p, (o, e) in zip(procs, outs):
        assert p.returncode == 0, e
    assert "OK" in outs[0][0]
