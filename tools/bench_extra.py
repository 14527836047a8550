"""Measurements for the non-headline BASELINE.json configs (one JSON line each; not part of the driver contract).

  cfg 3  MNIST-shaped 784 -> 10 logistic, rows generated on the device (per-GPU shard of the 50M-row batch)
  cfg 5  PyTorch 2-layer MLP 64 -> 32 -> 10, 10M rows
  cfg 4  FastAPI /predict, batch = 32: p50 / p99 latency through the ASGI app (in-process client)

    python tools/bench_extra.py [--rows784 6250000] [--steps 10]
"""
import argparse
import json
import statistics
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from unionml_b200.engine import Engine  # noqa: E402

PEAK = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"] if (ROOT / "MEASURED_PEAKS.json").exists() else 6650.0


def timed(fn, steps, warmup=3):
    for _ in range(warmup):
        fn()
    ks = []
    for _ in range(steps):
        st = fn()
        ks.append(st["kernel_ms"])
    return statistics.mean(ks), min(ks)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows784", type=int, default=6_250_000)
    ap.add_argument("--rows-mlp", type=int, default=10_000_000)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--requests", type=int, default=300)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    eng = Engine(0)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    eng.set_stream(stream.cuda_stream)

    # ---- cfg 3: 784 -> 10 (W ~ N(0, 0.05), b ~ N(0,1), X = uint8/255 generated on the device) ----
    F = 784
    w = (np.random.default_rng(1).standard_normal((10, F)) * 0.05).astype(np.float32)
    b = np.random.default_rng(2).standard_normal(10).astype(np.float32)
    model = eng.load_linear(w, b)
    g = torch.Generator(device=dev).manual_seed(1000)
    X = (torch.randint(0, 256, (args.rows784, F), generator=g, device=dev, dtype=torch.int32).to(torch.float32) / 255.0).contiguous()
    batch = eng.wrap_device(X.data_ptr(), args.rows784, F, F, keepalive=X)
    out = torch.empty(args.rows784, dtype=torch.int32, device=dev)
    mean_ms, min_ms = timed(lambda: eng.predict(model, batch, exact=True, out_device_ptr=out.data_ptr())[1], args.steps)
    # parity on a streamed 1M-row chunk against float64 numpy
    chunk = X[:1_000_000].cpu().numpy().astype(np.float64)
    want = (chunk @ w.astype(np.float64).T + b.astype(np.float64)).argmax(1)
    ok = bool(np.array_equal(out[:1_000_000].cpu().numpy(), want))
    gbs = args.rows784 * 4 * F / (mean_ms * 1e-3) / 1e9
    print(json.dumps({"config": "cfg3 784->10 logistic (per-GPU shard)", "rows": args.rows784, "kernel_ms": mean_ms,
                      "kernel_ms_min": min_ms, "rows_per_s": args.rows784 / (mean_ms * 1e-3), "achieved_GBs": gbs,
                      "roofline_frac_of_measured_hbm": gbs / PEAK, "parity_first_1M_rows_vs_float64": ok}), flush=True)
    del X, batch, out
    torch.cuda.empty_cache()

    # ---- cfg 5: MLP 64 -> 32 -> 10 ----
    z = np.load(ROOT / "tests" / "golden" / "mlp_64_32_10.npz")
    mlp = eng.load_mlp(z["w1"], z["b1"], z["w2"], z["b2"])
    Xh = eng.pinned_empty((args.rows_mlp, 64), np.float32)
    for k, r0 in enumerate(range(0, args.rows_mlp, 1_000_000)):
        Xh[r0:r0 + 1_000_000] = np.random.default_rng(k).integers(0, 17, size=(min(1_000_000, args.rows_mlp - r0), 64), dtype=np.uint8)
    bm = eng.stage(Xh)
    out = torch.empty(args.rows_mlp, dtype=torch.int32, device=dev)
    res = {}
    for mode, exact in (("exact", True), ("fast", False)):
        mean_ms, min_ms = timed(lambda: eng.predict_mlp(mlp, bm, exact=exact, out_device_ptr=out.data_ptr())[1], args.steps)
        res[mode] = {"kernel_ms": mean_ms, "kernel_ms_min": min_ms, "rows_per_s": args.rows_mlp / (mean_ms * 1e-3),
                     "achieved_GBs": args.rows_mlp * 256 / (mean_ms * 1e-3) / 1e9,
                     "fp32_TFLOPs": args.rows_mlp * 4736 / (mean_ms * 1e-3) / 1e12}
    _, st = eng.predict_mlp(mlp, bm, exact=True, out_device_ptr=out.data_ptr())
    # parity spot check in float64 numpy (tools/ may not touch oracle/): argmax of W2 relu(W1 x + b1) + b2
    x64 = Xh[:1_000_000].astype(np.float64)
    hid = np.maximum(x64 @ z["w1"].astype(np.float64).T + z["b1"].astype(np.float64), 0.0)
    want = (hid @ z["w2"].astype(np.float64).T + z["b2"].astype(np.float64)).argmax(1)
    ok = bool(np.array_equal(out[:1_000_000].cpu().numpy(), want))
    print(json.dumps({"config": "cfg5 MLP 64->32->10", "rows": args.rows_mlp, **res, "rows_rescored_fp64": st["n_flagged"],
                      "roofline_frac_of_measured_hbm": res["exact"]["achieved_GBs"] / PEAK,
                      "parity_first_1M_rows_vs_float64": ok}), flush=True)

    # ---- cfg 4: FastAPI /predict, batch = 32 ----
    import pandas as pd
    from fastapi import FastAPI
    from fastapi.testclient import TestClient
    from sklearn.datasets import load_digits
    from sklearn.linear_model import LogisticRegression
    from typing import List

    from unionml_b200 import Dataset, Model, ModelArtifact
    from unionml_b200.predictors import linear_argmax

    dataset = Dataset(name="digits_dataset", test_size=0.2, shuffle=True, targets=["target"])
    m = Model(name="digits_classifier", init=LogisticRegression, dataset=dataset)

    @dataset.reader
    def reader() -> pd.DataFrame:
        return load_digits(as_frame=True).frame

    @m.predictor
    def predictor(estimator: LogisticRegression, features: pd.DataFrame) -> List[float]:
        return linear_argmax(estimator, features)

    zz = np.load(ROOT / "tests" / "golden" / "digits_lr.npz")
    est = LogisticRegression()
    est.coef_, est.intercept_, est.classes_, est.n_features_in_ = zz["coef"], zz["intercept"], zz["classes"], 64
    m.artifact = ModelArtifact(est)
    app = FastAPI()
    m.serve(app)
    frame = load_digits(as_frame=True).frame
    feats = frame[[c for c in frame if c != "target"]]
    lat, lat_cpu = [], []
    with TestClient(app) as client:
        for i in range(args.requests + 20):
            body = {"features": feats.sample(32, random_state=i).to_dict(orient="records")}
            t0 = time.perf_counter()
            r = client.post("/predict", json=body)
            dt = time.perf_counter() - t0
            assert r.status_code == 200 and len(r.json()) == 32
            if i >= 20:
                lat.append(dt * 1e3)
    # the predictor call alone (what the device path adds to a request)
    sample = feats.sample(32, random_state=0)
    pred = []
    for _ in range(200):
        t0 = time.perf_counter()
        linear_argmax(est, sample)
        pred.append((time.perf_counter() - t0) * 1e3)
    ref = []
    est.feature_names_in_ = np.asarray(sample.columns, dtype=object)
    for _ in range(200):
        t0 = time.perf_counter()
        [float(x) for x in est.predict(sample)]
        ref.append((time.perf_counter() - t0) * 1e3)
    q = lambda v, p: float(np.percentile(v, p))  # noqa: E731
    print(json.dumps({"config": "cfg4 FastAPI /predict batch=32 (in-process ASGI client)", "requests": len(lat),
                      "p50_ms": q(lat, 50), "p99_ms": q(lat, 99),
                      "predictor_call_p50_ms": q(pred, 50), "predictor_call_p99_ms": q(pred, 99),
                      "sklearn_cpu_predictor_call_p50_ms": q(ref, 50)}), flush=True)


if __name__ == "__main__":
    main()
