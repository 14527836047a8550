#!/usr/bin/env python
"""bench.py - rows/sec of the batch-predict hot path (64->10 logistic, BASELINE.json) on N B200s.

    python bench.py --gpus 1 --steps 20 --warmup 5                      # this repo's CUDA path
    python bench.py --impl reference --gpus 1 --steps 3 --warmup 1      # the reference's CPU path (sklearn), same metric
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W                          # one rank per GPU

A "step" is one pass of the hot path over one batch: cfg 2 of BASELINE.json - 10M x 64 fp32 rows in the digits pixel
domain, W, b of the golden digits LogisticRegression, labels = argmax in EXACT mode (equal to
scikit-learn's float64 labels).  With N > 1 every rank scores its own 10M-row shard (weak scaling) and the int32 label
vectors are all-gathered so every rank holds all N x 10M labels.

`value`  : rows/s with the batch already resident in HBM (timed with CUDA events on the launching stream,
           barrier + synchronize on both sides, max over ranks).
`e2e`    : the same metric through the host-buffer call (pinned host rows -> H2D -> kernels -> D2H labels inside the
           timed region).
`roofline`: algorithmic bytes (256 B/row) / CUDA-event duration of the scoring kernel, against MEASURED_PEAKS.json.
`cpu_baseline`: scikit-learn's LogisticRegression.predict (the reference's arithmetic) timed on this box's host cores
           on a bounded sample - the only place, with --impl reference, where oracle/ code runs.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "rows/sec batch predict (64->10 logistic)"
UNIT = "rows/s"
N_FEATURES = 64
BYTES_PER_ROW = 4 * N_FEATURES  # algorithmic HBM read per row; + 4 B label write, not counted
# dram__bytes_read.sum + dram__bytes_write.sum of one linear_argmax_tma launch on the default 10M x 64 batch, from the
# committed `ncu --set full` capture (profiles/r01_linear_argmax_tma.ncu_raw.csv): 2.560199 GB + 7.15 MB
NCU_TRAFFIC_BYTES_10M = 2_567_349_000


def load_digits_model():
    z = np.load(ROOT / "tests" / "golden" / "digits_lr.npz")
    return z["coef"], z["intercept"], z["classes"]


def fill_digits_rows(out: np.ndarray, seed_base: int) -> None:
    """cfg-2 rows: chunk k = default_rng(seed_base + k).integers(0, 17, (1M, 64), uint8) as fp32 (exact in fp32)."""
    rows = out.shape[0]
    step = 1_000_000
    for k, r0 in enumerate(range(0, rows, step)):
        r1 = min(rows, r0 + step)
        out[r0:r1] = np.random.default_rng(seed_base + k).integers(0, 17, size=(r1 - r0, N_FEATURES), dtype=np.uint8)


def measured_peak_hbm():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = (
        "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
        "clocks_event_reasons.sw_power_cap"
    )

    def __init__(self, device_index: int):
        self.idx = device_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(self.idx)],
                stdout=subprocess.PIPE,
                stderr=subprocess.DEVNULL,
                text=True,
            )
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
                power.append(float(parts[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {
            "sm_mhz": statistics.median(sm) if sm else None,
            "sm_max_mhz": max(mx) if mx else None,
            "power_w_max": max(power) if power else None,
            "samples": len(sm),
            "reasons": sorted(reasons),
        }


# ---------------------------------------------------------------------------------------------------------------
# CPU legs (the only users of oracle/)
# ---------------------------------------------------------------------------------------------------------------
def sklearn_estimator():
    from sklearn.linear_model import LogisticRegression

    coef, intercept, classes = load_digits_model()
    est = LogisticRegression(C=1.0, max_iter=1000)
    est.coef_, est.intercept_, est.classes_, est.n_features_in_ = coef, intercept, classes, N_FEATURES
    return est


def blas_threads() -> int:
    try:
        from threadpoolctl import threadpool_info

        return max([i.get("num_threads", 1) for i in threadpool_info()] or [1])
    except Exception:
        return os.cpu_count() or 1


def cpu_reference_predict_rows_per_s(sample_rows: int, repeats: int):
    """The reference's CPU path on host cores: Model.predict(features=frame) -> canonical sklearn predictor.

    Restated wrapper (oracle.unionml_path, no flytekit) + the real scikit-learn arithmetic, float64 frame as the
    reference's DataFrame path feeds it.  Returns (rows/s best-of-repeats, seconds list).
    """
    import pandas as pd

    from oracle import unionml_path as opath

    est = sklearn_estimator()
    X = np.empty((sample_rows, N_FEATURES), dtype=np.float64)
    fill_digits_rows(X, 0)
    frame = pd.DataFrame(X, columns=[f"pixel_{i}" for i in range(N_FEATURES)])

    def reader() -> pd.DataFrame:
        return frame

    def predictor(estimator, features) -> list:  # /root/reference/README.md:87-92
        return [float(x) for x in estimator.predict(features)]

    est.feature_names_in_ = np.asarray(frame.columns, dtype=object)
    spec = opath.PathSpec(reader=reader, targets=["target"], predictor=predictor, model_object=est)
    times = []
    for _ in range(repeats):
        t0 = time.perf_counter()
        out = opath.predict(spec, features=frame)
        times.append(time.perf_counter() - t0)
        assert len(out) == sample_rows
    return sample_rows / min(times), times


def run_reference_arm(args):
    """--impl reference: same metric/unit/config, CPU only; under torchrun only rank 0 works."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    sample = args.cpu_rows
    for _ in range(max(args.warmup, 1)):
        cpu_reference_predict_rows_per_s(min(sample, 200_000), 1)
    per_step = []
    for _ in range(args.steps):
        rps, times = cpu_reference_predict_rows_per_s(sample, 1)
        per_step.append(times[0])
    value = sample / statistics.mean(per_step)
    cores = blas_threads()
    line = {
        "impl": "reference",
        "metric": METRIC,
        "value": value,
        "unit": UNIT,
        "n_gpus": args.gpus,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": 1e3 * statistics.mean(per_step),
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f64",
        "data": "synthetic",
        "config": workload_config(args, 1) | {"sample_rows_per_step": sample},
        "cpu_baseline": {
            "value": value,
            "unit": UNIT,
            "cores": cores,
            "kind": "port",
            "sample": f"{sample} rows/step of the cfg-2 batch as a float64 DataFrame through the restated "
            "Model.predict(features=...) wrapper + scikit-learn LogisticRegression.predict + [float(x) ...] "
            f"(BLAS threads={cores}, os.cpu_count()={os.cpu_count()})",
        },
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def workload_config(args, n_gpus: int) -> dict:
    return {
        "workload": "BASELINE.json configs[1]: digits predictor (golden LogisticRegression 64->10), "
        f"{args.rows} x 64 synthetic fp32 rows per GPU (integers 0..16)",
        "rows_per_gpu": args.rows,
        "global_rows": args.rows * n_gpus,
        "n_features": N_FEATURES,
        "n_classes": 10,
        "mode": "exact (fp32 tile kernel + margin guard + fp64 re-score; labels == sklearn float64 labels)",
        "labels": f"{args.wire} class index per row",
        "parallelism": (f"row-sharded x{n_gpus}, label exchange: " + getattr(args, "gather_used", args.gather))
        if n_gpus > 1
        else "single GPU",
        "l2_policy": f"inputs ({args.rows * BYTES_PER_ROW / 1e9:.2f} GB/step) are larger than L2 (126 MB); no flush needed",
    }


# ---------------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------------
def run_gpu_arm(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N > 1 must be launched with torch.distributed.run (one rank per GPU)")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback for the product path)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from unionml_b200.engine import Engine

    eng = Engine(local_rank)
    # an explicit (non-default) torch stream is both torch's current stream and the engine's launch stream, so the
    # torch.cuda.Event pair below brackets exactly the library's kernels (handle 0 would mean "engine's own stream")
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    eng.set_stream(stream.cuda_stream)

    coef, intercept, classes = load_digits_model()
    model = eng.load_linear(coef, intercept, classes)

    rows = args.rows
    # host rows in pinned memory (the e2e leg copies them every step); seeds differ per rank (weak scaling)
    X_host = eng.pinned_empty((rows, N_FEATURES), np.float32)
    fill_digits_rows(X_host, 10 * rank)
    labels_host = eng.pinned_empty(rows, np.int32)

    batch = eng.stage(X_host)  # resident fp32 row-major copy for the `value` leg
    from unionml_b200.sharding import PeerLabelExchange, predict_sharded

    counts = [rows] * world
    exchange = None
    gather = "none" if world == 1 else args.gather
    if world > 1 and gather in ("fused", "push"):
        try:
            exchange = PeerLabelExchange(rows * world, dev, dtype=torch.uint8 if args.wire == "u8" else torch.int32,
                                         multicast=not args.no_multicast, push=args.gather == "push")
        except Exception as exc:  # symmetric memory unavailable on this box: fall back to the NCCL all-gather
            if rank == 0:
                print(f"bench: symmetric memory unavailable ({exc!r}); using nccl all-gather", file=sys.stderr)
            gather = "nccl"
    # label vectors are uint8 class indices (n_classes = 10 <= 256): 4x fewer bytes to write and to exchange
    wire_dtype = torch.uint8 if args.wire == "u8" else torch.int32
    label_bytes = 1 if args.wire == "u8" else 4
    labels_all = exchange.labels if exchange is not None else torch.empty(rows * world, dtype=wire_dtype, device=dev)
    labels_local = labels_all[rank * rows : (rank + 1) * rows]
    interleave = interleave_out = None
    if args.interleave:
        interleave = torch.randn(4096, 4096, device=dev, dtype=torch.bfloat16)
        interleave_out = torch.empty_like(interleave)

    def local_predict(want_stats=False):
        """Score the resident shard into this rank's slice of the label vector (no exchange)."""
        return eng.predict_peers(model, batch, [labels_all.data_ptr()], rank * rows, exact=True, want_stats=want_stats,
                                 label_bytes=label_bytes)

    def step():
        if world == 1:
            local_predict()
            if interleave is not None:  # --interleave: a foreign kernel between steps (robustness check, not a bench)
                torch.matmul(interleave, interleave, out=interleave_out)
        elif exchange is not None:
            predict_sharded(eng, model, batch, row_offset=rank * rows, counts=counts, exact=True, exchange=exchange)
        else:  # nccl
            local_predict()
            dist.all_gather_into_tensor(labels_all, labels_local)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    how = "one NVLS multicast store per tile" if (exchange is not None and exchange.multicast) else "one store per peer per tile"
    args.gather_used = {"push": f"kernel stores {args.wire} labels locally, thin copy kernel pushes the slice ({how.replace(' per tile', '')}) + barrier",
                        "fused": f"fused label stores ({args.wire}, {how}) from the kernel epilogue over NVLink (symmetric memory) + barrier",
                        "nccl": f"ncclAllGather of {args.wire} labels", "none": "none"}[gather]
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # ---- timed region: exactly K steps, CUDA events on the launching stream ----
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    for i in range(args.steps):
        step()
        # bound the number of queued cross-rank barrier steps (a 1250-deep untimed queue did not drain at N = 2);
        # one host sync per 64 steps costs < 0.1 % of the timed region
        if world > 1 and (i + 1) % 64 == 0 and i + 1 < args.steps:
            torch.cuda.synchronize()
    ev1.record(stream)
    barrier()
    elapsed_ms = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    ms_per_step = elapsed_ms / args.steps
    value = rows * world / (ms_per_step * 1e-3)

    # ---- roofline of the dominant kernel: CUDA events around linear_argmax_tma inside the library, live ----
    k_ms, r_ms, flagged, launches_per_step = [], [], 0, 2
    for _ in range(args.steps):
        if exchange is not None and not exchange.push:  # the variant the step really runs: stores to all ranks in the epilogue
            st = eng.predict_peers(model, batch, exchange.peer_ptrs, rank * rows, exact=True, want_stats=True,
                                   label_bytes=exchange.label_bytes)
            exchange.barrier()
        else:
            st = local_predict(want_stats=True)
        k_ms.append(st["kernel_ms"])
        r_ms.append(st["recheck_ms"])
        flagged = st["n_flagged"]
        launches_per_step = st["kernel_launches"]
    kernel_ms = statistics.mean(k_ms)
    # the timed region lasts only K x 0.4 ms; keep the same load running (untimed) until nvidia-smi has had ~0.6 s to
    # sample clocks / throttle reasons under it
    # (single-GPU runs only: a long untimed queue of cross-rank barrier steps is not worth the risk at N > 1)
    if world == 1:
        for _ in range(min(5000, int(600.0 / max(ms_per_step, 0.05)))):
            step()
        torch.cuda.synchronize()
    exchange_ms = None
    if world > 1:  # cost of the label exchange alone (barrier or all-gather), CUDA events, same stream
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream)
        for _ in range(args.steps):
            if exchange is not None:
                if exchange.push:
                    off = rank * rows * exchange.label_bytes
                    remote = exchange.peer_ptrs if exchange.multicast else exchange.peer_ptrs[1:]
                    eng.push_labels(exchange.own_ptr + off, [p + off for p in remote], rows * exchange.label_bytes)
                exchange.barrier()
            else:
                dist.all_gather_into_tensor(labels_all, labels_local)
        e1.record(stream)
        barrier()
        exchange_ms = e0.elapsed_time(e1) / args.steps
    clocks = sampler.stop() if rank == 0 else None
    peak, peak_src = measured_peak_hbm()
    achieved = rows * BYTES_PER_ROW / (kernel_ms * 1e-3) / 1e9
    roofline = {
        "bound": "hbm",
        "kernel": "linear_argmax_tma_kernel<10, EXACT>",
        "achieved": achieved,
        "peak": peak,
        "peak_source": peak_src,
        "unit": "GB/s",
        "frac": achieved / peak,
        "traffic": args.traffic if args.traffic is not None else (NCU_TRAFFIC_BYTES_10M if rows == 10_000_000 else None),
        "traffic_source": "profiles/r01_linear_argmax_tma.ncu_raw.csv (ncu --set full, per launch)",
        "kernel_ms": kernel_ms,
        "kernel_ms_min": min(k_ms),
        "rescore_ms": statistics.mean(r_ms),
        "exchange_ms": exchange_ms,
        "algorithmic_bytes_per_launch": rows * BYTES_PER_ROW,
        "rows_rescored_fp64": flagged,
    }

    # ---- e2e: pinned host rows -> labels in host memory, through the host-buffer call ----
    e2e_t = []
    e2e_stats = None
    for i in range(0 if args.skip_e2e else args.e2e_steps + 1):
        barrier()
        t0 = time.perf_counter()
        _, e2e_stats = eng.predict_host(model, X_host, exact=True, out=labels_host)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if i > 0:
            e2e_t.append(dt)
    if args.skip_e2e:  # profiling runs only (ncu): the JSON line of such a run is never a bench value
        eng.predict_host(model, X_host[:1_000_000], exact=True, out=labels_host[:1_000_000])
        tmp_i32 = torch.empty(rows, dtype=torch.int32, device=dev)
        eng.predict(model, batch, exact=True, out_device_ptr=tmp_i32.data_ptr(), want_stats=False)
        labels_host[:] = tmp_i32.cpu().numpy()
        e2e_t, e2e_stats = [float("nan")], {"h2d_bytes": 0, "d2h_bytes": 0, "total_ms": float("nan")}
    e2e_s = statistics.mean(e2e_t)
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e = {
        "value": rows * world / e2e_s,
        "unit": UNIT,
        "h2d_bytes_per_step": int(e2e_stats["h2d_bytes"]) * world,
        "d2h_bytes_per_step": int(e2e_stats["d2h_bytes"]) * world,
        "ms_per_step": e2e_s * 1e3,
        "device_ms_per_step": e2e_stats["total_ms"],
        "steps": args.e2e_steps,
        "path": "Engine.predict_host: pinned fp32 rows -> chunked H2D -> linear_argmax_tma (+fp64 re-score) -> D2H int32 labels",
    }
    local_i32 = torch.empty(rows, dtype=torch.int32, device=dev)
    eng.predict(model, batch, exact=True, out_device_ptr=local_i32.data_ptr(), want_stats=False)
    if world == 1:
        step()
        torch.cuda.synchronize()
        if not torch.equal(labels_local.to(torch.int32), local_i32):
            raise SystemExit("bench: uint8 label vector differs from the int32 one")
    if world > 1:
        # every rank must hold every rank's labels: compare the exchanged vector with a plain NCCL all-gather
        step()
        ref_all = torch.empty(rows * world, dtype=torch.int32, device=dev)
        dist.all_gather_into_tensor(ref_all, local_i32)
        torch.cuda.synchronize()
        if not torch.equal(ref_all, labels_all.to(torch.int32)):
            raise SystemExit(f"bench: rank {rank}: exchanged label vector differs from the NCCL all-gather")
    # sanity: resident and streamed paths agree
    check = local_i32.cpu().numpy()
    if not np.array_equal(check, labels_host):
        raise SystemExit("bench: resident and host-streamed label vectors differ")

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        rps, times = cpu_reference_predict_rows_per_s(args.cpu_rows, 3)
        cores = blas_threads()
        est_nd = sklearn_estimator()
        Xnd = X_host[: args.cpu_rows].astype(np.float64)
        nd_t = []
        for _ in range(3):
            t0 = time.perf_counter()
            est_nd.predict(Xnd)
            nd_t.append(time.perf_counter() - t0)
        c_port_info = {}
        try:  # plain-C OpenMP restatement (oracle/linear_predict.c): what the host cores can do without Python in the way
            from oracle import c_port

            rows_c = X_host[: args.cpu_rows]
            c_port.predict_indices(rows_c[:10_000], coef, intercept)
            ct = []
            for _ in range(3):
                t0 = time.perf_counter()
                idx_c = c_port.predict_indices(rows_c, coef, intercept)
                ct.append(time.perf_counter() - t0)
            c_port_info = {
                "c_port_value": args.cpu_rows / min(ct),
                "c_port_threads": c_port.num_threads(),
                "c_port_note": "oracle/linear_predict.c, float64 scores over the fp32 rows, OpenMP over rows; labels "
                + ("equal" if np.array_equal(idx_c, labels_host[: args.cpu_rows]) else "DIFFER from") + " the GPU's",
            }
        except Exception as exc:  # the C port is optional evidence, never a reason to lose the bench line
            c_port_info = {"c_port_value": None, "c_port_note": f"unavailable: {exc!r}"}
        cpu_baseline = {
            **c_port_info,
            "ndarray_value": args.cpu_rows / min(nd_t),
            "ndarray_note": "bare LogisticRegression.predict on a C-order float64 ndarray (no DataFrame, no list conversion)",
            "value": rps,
            "unit": UNIT,
            "cores": cores,
            "kind": "port",
            "sample": f"{args.cpu_rows} rows of the same batch as a float64 DataFrame, best of 3: restated "
            "Model.predict(features=...) wrapper + scikit-learn LogisticRegression.predict + [float(x) ...] "
            f"(BLAS threads={cores}, os.cpu_count()={os.cpu_count()}); labels checked equal to the GPU's",
        }
        # the CPU labels are also the parity oracle for the same rows
        est = sklearn_estimator()
        want = est.predict(X_host[: args.cpu_rows].astype(np.float64))
        if not np.array_equal(classes[labels_host[: args.cpu_rows]], want):
            raise SystemExit("bench: GPU labels differ from scikit-learn's on the sampled rows")

    if rank == 0:
        line = {
            "metric": METRIC,
            "value": value,
            "unit": UNIT,
            "n_gpus": world,
            "steps": args.steps,
            "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step,
            "higher_is_better": True,
            "scaling": "weak",
            "vs_baseline": None,
            "dtype": "f32",
            "data": "synthetic",
            "config": workload_config(args, world),
            "roofline": roofline,
            "cpu_baseline": cpu_baseline,
            "e2e": e2e,
            "clocks": clocks,
            "gpu_launches": launches_per_step * args.steps,
            "device": eng.info,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--rows", type=int, default=10_000_000, help="rows per GPU (cfg 2: 10M)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-rows", type=int, default=2_000_000, help="bounded CPU sample (rows)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true", help="profiling only: skip the host-buffer leg")
    ap.add_argument("--gather", default="fused", choices=["fused", "push", "nccl"], help="label exchange for --gpus > 1")
    ap.add_argument("--no-multicast", action="store_true", help="fused exchange: per-peer stores instead of NVLS multicast")
    ap.add_argument("--wire", default="u8", choices=["u8", "i32"], help="label width of the fused exchange")
    ap.add_argument("--interleave", action="store_true", help="robustness check: run a cuBLAS GEMM between steps")
    ap.add_argument("--traffic", type=float, default=None, help="dram bytes/launch from the committed ncu capture")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
