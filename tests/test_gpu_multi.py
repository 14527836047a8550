"""2-GPU parity of the sharded path (fused peer-store epilogue, pipelined push, NCCL all-gather; linear and MLP);
skipped on 1-GPU boxes - bench.py repeats the float64 check of every rank's rows at N > 1 for that reason."""
import os
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent

WORKER = r'''
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["UML_ROOT"])
from oracle import linear as olin
from unionml_b200.engine import Engine
from unionml_b200.sharding import PeerLabelExchange, predict_sharded, shard_bounds, shard_counts
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
z = np.load(os.path.join(os.environ["UML_ROOT"], "tests", "golden", "digits_lr.npz"))
N = 1_000_003
X = np.random.default_rng(7).integers(0, 17, size=(N, 64), dtype=np.uint8).astype(np.float32)   # same on all ranks
want = olin.predict_indices(olin.decision_function(X.astype(np.float64), z["coef"], z["intercept"])).astype(np.int32)
eng = Engine(local); s = torch.cuda.Stream(device=dev); torch.cuda.set_stream(s); eng.set_stream(s.cuda_stream)
m = eng.load_linear(z["coef"], z["intercept"])
lo, hi = shard_bounds(N, rank, world); counts = shard_counts(N, world)
b = eng.stage(X[lo:hi])
got = predict_sharded(eng, m, b, row_offset=lo, counts=counts, exact=True)            # NCCL all-gather
torch.cuda.synchronize(); assert np.array_equal(got.cpu().numpy(), want), "nccl path"
variants = ((torch.int32, False, False, 1), (torch.uint8, False, False, 1), (torch.uint8, True, False, 1),
            (torch.int32, True, False, 1), (torch.uint8, True, True, 1), (torch.uint8, False, True, 1),
            (torch.uint8, True, True, 4), (torch.uint8, False, True, 3))   # last two: pipelined push (sub-batches)
for dtype, mc, push, pipe in variants:
    ex = PeerLabelExchange(N, dev, dtype=dtype, multicast=mc, push=push, pipeline=pipe)   # fused stores / two-step push
    for _ in range(3):
        ex.labels.fill_(99); ex.barrier()
        got = predict_sharded(eng, m, b, row_offset=lo, counts=counts, exact=True, exchange=ex)
        torch.cuda.synchronize(); assert np.array_equal(got.cpu().numpy().astype(np.int32), want), f"fused path {dtype} {mc} {push} {pipe}"
# a default Engine (its own non-blocking stream): predict_sharded must order the exchange after the kernels itself
eng2 = Engine(local); m2 = eng2.load_linear(z["coef"], z["intercept"]); b2 = eng2.stage(X[lo:hi])
ex = PeerLabelExchange(N, dev, dtype=torch.uint8, multicast=True)
for _ in range(5):
    ex.labels.fill_(99); ex.barrier()
    got = predict_sharded(eng2, m2, b2, row_offset=lo, counts=counts, exact=True, exchange=ex)
    torch.cuda.synchronize(); assert np.array_equal(got.cpu().numpy().astype(np.int32), want), "default-engine stream ordering"
# the MLP predictor through the same sharded path (cfg 5): tensor-core kernel with fused uint8 stores, and NCCL
from oracle import mlp as omlp
g = np.load(os.path.join(os.environ["UML_ROOT"], "tests", "golden", "mlp_64_32_10.npz"))
mm = eng.load_mlp(g["w1"], g["b1"], g["w2"], g["b2"])
want_mlp = omlp.predict_indices_f64(X, g["w1"], g["b1"], g["w2"], g["b2"]).astype(np.int32)
got = predict_sharded(eng, mm, b, row_offset=lo, counts=counts, exact=True)
torch.cuda.synchronize(); assert np.array_equal(got.cpu().numpy(), want_mlp), "mlp nccl path"
for dtype, mc, push, pipe in ((torch.uint8, True, False, 1), (torch.int32, False, False, 1), (torch.uint8, True, True, 4)):
    ex = PeerLabelExchange(N, dev, dtype=dtype, multicast=mc, push=push, pipeline=pipe)
    for _ in range(3):
        ex.labels.fill_(99); ex.barrier()
        got = predict_sharded(eng, mm, b, row_offset=lo, counts=counts, exact=True, exchange=ex)
        torch.cuda.synchronize(); assert np.array_equal(got.cpu().numpy().astype(np.int32), want_mlp), f"mlp fused path {dtype} {mc} {push} {pipe}"
dist.barrier(); dist.destroy_process_group()
print(f"rank {rank} ok")
'''


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sharded_predict_two_gpus(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, UML_ROOT=str(ROOT))
    r = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
         "--master-port", "29611", str(script)],
        env=env, capture_output=True, text=True, timeout=600,
    )
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "rank 0 ok" in r.stdout and "rank 1 ok" in r.stdout
