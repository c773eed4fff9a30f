"""world_size-2 gloo tests (CPU) of the multi-GPU host logic (ctn_b200/dist.py): batch shards, global loss mean,
max-over-ranks timing.  The data path itself has no collective."""
import os
import subprocess
import sys

import pytest

from ctn_b200 import dist as D

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys, torch
sys.path.insert(0, os.path.join(%(root)r, "dnn-based_source_separation_b200"))
from ctn_b200 import dist as D
rank, local_rank, world = D.init("gloo")
assert world == 2
G = 7
lo, hi = D.shard_bounds(G, rank, world)
losses = torch.arange(G, dtype=torch.float32) * 1.5 - 2.0        # "per-sample losses" of the global batch
mean = D.global_loss_mean(losses[lo:hi], G)
assert abs(float(mean) - float(losses.mean())) < 1e-6, (float(mean), float(losses.mean()))
t = D.max_over_ranks(10.0 + rank)
assert t == 11.0
s = D.sum_over_ranks(float(hi - lo))
assert s == G
# gradient all-reduce of the batch-sharded training step: unequal shards (4 + 3 of 7), both bucket paths
class M(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.a = torch.nn.Parameter(torch.zeros(5, 3)); self.b = torch.nn.Parameter(torch.zeros(7))
per_sample = torch.arange(G * 22, dtype=torch.float32).reshape(G, 22) / 10.0     # "per-sample gradients"
want = per_sample.mean(dim=0)                                                    # gradient of the GLOBAL batch mean
for aliased in (False, True):
    m = M()
    local = per_sample[lo:hi].mean(dim=0)                                        # gradient of this shard's mean loss
    if aliased:   # the native backward hands out views of one flat buffer
        flat = local.clone(); m.last_flat_grad = flat
        m.a.grad = flat[:15].view(5, 3); m.b.grad = flat[15:].view(7)
    else:
        m.a.grad = local[:15].clone().view(5, 3); m.b.grad = local[15:].clone()
    n = D.allreduce_gradients(m, local_batch=hi - lo, global_batch=G)
    assert n == 22
    got = torch.cat([m.a.grad.reshape(-1), m.b.grad.reshape(-1)])
    assert torch.allclose(got, want, rtol=1e-6, atol=1e-6), (aliased, got, want)
D.barrier()
print("rank", rank, "ok", lo, hi)
"""


def test_shard_bounds_cover_batch():
    for G in (0, 1, 7, 32, 64, 129):
        for world in (1, 2, 3, 8):
            spans = [D.shard_bounds(G, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == G
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c and b >= a
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        D.shard_bounds(8, 2, 2)


def test_gloo_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT})
    port = 29600 + os.getpid() % 300
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=180)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
        assert "ok" in o
