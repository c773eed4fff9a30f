"""2-GPU test of the data-parallel training step (``-m gpu``; skipped on a 1-GPU box): batch shards on two ranks,
native forward/backward per rank, ONE NCCL all-reduce of the flat gradient bucket (ctn_b200.dist.allreduce_gradients)
-- the result must equal the gradients of the same global batch computed on a single GPU."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys, torch
root = %(root)r
sys.path.insert(0, os.path.join(root, "dnn-based_source_separation_b200")); sys.path.insert(0, os.path.join(root, "oracle"))
sys.path.insert(0, os.path.join(root, "tests"))
import convtasnet_oracle as O
from ctn_b200 import dist as D
from ctn_b200.criterion.pit import PIT1d
from ctn_b200.criterion.sdr import NegSISDR
from test_parity_gpu import build_model
rank, local_rank, world = D.init("nccl")
dev = torch.device("cuda", local_rank)
cfg = O.OracleConfig(n_basis=32, kernel_size=16, sep_hidden_channels=64, sep_bottleneck_channels=32, sep_skip_channels=32,
                     sep_num_blocks=2, sep_num_layers=2, causal=False, n_sources=2)
sd = O.synth_state_dict(cfg, seed=7)
G = 6
mixture, sources = O.synth_batch(G, 2, 2000, seed=9)
crit = PIT1d(NegSISDR(), 2)
def grads(lo, hi):
    m = build_model(cfg, sd).to(dev).train()
    loss, _ = crit(m(mixture[lo:hi].to(dev)), sources[lo:hi].to(dev))
    loss.backward()
    return m
lo, hi = D.shard_bounds(G, rank, world)
m = grads(lo, hi)
n = D.allreduce_gradients(m, local_batch=hi - lo, global_batch=G)
full = grads(0, G)                                   # the same global batch on one GPU
worst = 0.0
for (k, p), (_, q) in zip(m.named_parameters(), full.named_parameters()):
    rel = float((p.grad - q.grad).abs().max()) / (float(q.grad.abs().max()) + 1e-30)
    worst = max(worst, rel)
    assert rel < 1e-4, (k, rel)
D.barrier()
print("rank", rank, "ok bucket", n, "worst", worst)
"""


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_ddp_gradients_match_single_gpu(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT})
    port = 29700 + os.getpid() % 200
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=300)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
        assert "ok bucket" in o
