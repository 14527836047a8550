"""The N>1 host logic on CPU: world_size-2 gloo, rows sharded, labels gathered (the kernel is played by the oracle)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from unionml_b200.sharding import shard_bounds, shard_counts


def test_shard_bounds_cover_and_balance():
    for n in (0, 1, 7, 128, 1000, 10_000_001):
        for world in (1, 2, 3, 4, 8):
            blocks = [shard_bounds(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1 and sizes == shard_counts(n, world)
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_rows, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import linear as olin
        from unionml_b200.sharding import gather_labels

        z = np.load(os.path.join(os.path.dirname(__file__), "golden", "digits_lr.npz"))
        X = np.random.default_rng(123).integers(0, 17, size=(n_rows, 64)).astype(np.float64)  # same on both ranks
        full = olin.predict_indices(olin.decision_function(X, z["coef"], z["intercept"])).astype(np.int32)
        lo, hi = shard_bounds(n_rows, rank, world)
        local = torch.from_numpy(full[lo:hi].copy())  # stands in for the kernel's output on this rank's shard
        got = gather_labels(local, shard_counts(n_rows, world))
        assert got.dtype == torch.int32 and got.numel() == n_rows
        np.testing.assert_array_equal(got.numpy(), full)
        with pytest.raises(ValueError):
            gather_labels(local[:-1], shard_counts(n_rows, world))
        open(os.path.join(out_dir, f"ok{rank}"), "w").close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_rows", [4096, 4097])  # equal shards (all_gather_into_tensor) and ragged shards (padded)
def test_gather_labels_world2_gloo(tmp_path, n_rows):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), n_rows, str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))


# ---------------------------------------------------------------------------------------------------------------
# predict_sharded end to end on CPU: the engine is replaced by a stub that writes the oracle's labels through the raw
# pointer it is handed (exactly how the CUDA library is driven), the exchange is a real gloo all-gather
# ---------------------------------------------------------------------------------------------------------------
class _StubEngine:
    device = 0

    def __init__(self, labels_for_shard):
        self._labels = labels_for_shard
        self.calls = []

    def predict(self, model, batch, exact=True, out_device_ptr=None, want_stats=True):
        import ctypes

        src = np.ascontiguousarray(self._labels, dtype=np.int32)
        ctypes.memmove(out_device_ptr, src.ctypes.data, src.nbytes)
        self.calls.append((exact, src.size))
        return None, None

    def predict_mlp(self, model, batch, exact=True, out_device_ptr=None, want_stats=True):
        self.predict(model, batch, exact=exact, out_device_ptr=out_device_ptr, want_stats=want_stats)
        self.calls[-1] = ("mlp",) + self.calls[-1]
        return None, None


class MlpModel:  # predict_sharded dispatches on the class name of the model handle (engine.MlpModel)
    pass


def _sharded_worker(rank, world, port, n_rows, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import linear as olin
        from unionml_b200.sharding import predict_sharded

        z = np.load(os.path.join(os.path.dirname(__file__), "golden", "digits_lr.npz"))
        X = np.random.default_rng(77).integers(0, 17, size=(n_rows, 64)).astype(np.float64)
        full = olin.predict_indices(olin.decision_function(X, z["coef"], z["intercept"])).astype(np.int32)
        lo, hi = shard_bounds(n_rows, rank, world)
        counts = shard_counts(n_rows, world)
        eng = _StubEngine(full[lo:hi])
        labels_all = torch.full((n_rows,), -1, dtype=torch.int32)
        got = predict_sharded(eng, model=None, batch=None, row_offset=lo, counts=counts, exact=True, labels_all=labels_all)
        np.testing.assert_array_equal(got.numpy(), full)
        assert eng.calls == [(True, hi - lo)]
        # the 2-layer MLP goes through the same sharded path (cfg 5): predict_mlp on this rank's rows, same exchange
        labels_all.fill_(-1)
        got = predict_sharded(eng, model=MlpModel(), batch=None, row_offset=lo, counts=counts, exact=False, labels_all=labels_all)
        np.testing.assert_array_equal(got.numpy(), full)
        assert eng.calls[-1] == ("mlp", False, hi - lo)
        open(os.path.join(out_dir, f"sharded_ok{rank}"), "w").close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_rows", [2048, 2051])
def test_predict_sharded_world2_gloo_with_stub_engine(tmp_path, n_rows):
    world = 2
    mp.spawn(_sharded_worker, args=(world, _free_port(), n_rows, str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"sharded_ok{r}").exists() for r in range(world))
