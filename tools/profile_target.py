"""A short, deterministic workload per kernel family for ncu captures (not a bench: numbers under a profiler are never reported).

    ncu --set full --clock-control none --import-source on -k regex:mlp_argmax_tc -s 2 -c 1 -o gpurun_out/x \
        python tools/profile_target.py --what mlp
"""
import argparse
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--what", required=True, choices=["linear", "mlp", "stage", "small"])
    ap.add_argument("--rows", type=int, default=10_000_000)
    ap.add_argument("--reps", type=int, default=4)
    ap.add_argument("--features", type=int, default=64, help="linear: 784 = the cfg-3 shape (random weights, pixel / 255 rows)")
    ap.add_argument("--chunk-rows", type=int, default=0, help="stage: run the staging kernels on chunks of this many rows "
                    "(through predict_host) instead of stage()'s default 64 MiB chunks")
    args = ap.parse_args()
    import torch

    from unionml_b200.engine import Engine

    eng = Engine(0)
    dev = torch.device("cuda", 0)
    rows = args.rows
    if args.what in ("linear", "mlp"):
        X = torch.randint(0, 17, (rows, 64), device=dev, dtype=torch.int32).to(torch.float32)
        b = eng.wrap_device(X.data_ptr(), rows, 64, keepalive=X)
        lab = torch.empty(rows, dtype=torch.uint8, device=dev)
        if args.what == "linear" and args.features != 64:
            F = args.features
            rng = np.random.default_rng(3)
            m = eng.load_linear(rng.standard_normal((10, F)) * 0.05, rng.standard_normal(10))
            X = (torch.randint(0, 256, (rows, F), device=dev, dtype=torch.int32).to(torch.float32) / 255.0).contiguous()
            b = eng.wrap_device(X.data_ptr(), rows, F, keepalive=X)
            for _ in range(args.reps):
                st = eng.predict_peers(m, b, [lab.data_ptr()], 0, exact=True, want_stats=True, label_bytes=1)
            print(st)
        elif args.what == "linear":
            z = np.load(ROOT / "tests" / "golden" / "digits_lr.npz")
            m = eng.load_linear(z["coef"], z["intercept"])
            for _ in range(args.reps):
                eng.predict_peers(m, b, [lab.data_ptr()], 0, exact=True, want_stats=True, label_bytes=1)
        else:
            z = np.load(ROOT / "tests" / "golden" / "mlp_64_32_10.npz")
            m = eng.load_mlp(z["w1"], z["b1"], z["w2"], z["b2"])
            for _ in range(args.reps):
                st = eng.predict_mlp_peers(m, b, [lab.data_ptr()], 0, exact=True, want_stats=True, label_bytes=1)
            print(st)
    elif args.what == "stage":
        rows = min(rows, 4_000_000)
        src = eng.pinned_empty((64, rows), np.float64)  # feature-major float64: a pandas block
        src[:] = np.random.default_rng(0).integers(0, 17, size=(64, rows))
        rm = eng.pinned_empty((rows, 64), np.float64)
        rm[:] = src.T
        if args.chunk_rows:
            z = np.load(ROOT / "tests" / "golden" / "digits_lr.npz")
            m = eng.load_linear(z["coef"], z["intercept"])
            for _ in range(args.reps):
                eng.predict_host(m, src.T, exact=True, chunk_rows=args.chunk_rows)   # stage_featmajor_kernel<double>
                eng.predict_host(m, rm, exact=True, chunk_rows=args.chunk_rows)      # stage_dense_kernel<double>
        else:
            for _ in range(args.reps):
                eng.stage(src.T).free()                      # stage_featmajor_kernel<double>
            for _ in range(args.reps):
                eng.stage(rm).free()                         # stage_dense_kernel<double>
    else:
        z = np.load(ROOT / "tests" / "golden" / "digits_lr.npz")
        m = eng.load_linear(z["coef"], z["intercept"])
        X = np.random.default_rng(0).integers(0, 17, size=(32, 64)).astype(np.float64)
        for _ in range(50):
            eng.predict_host(m, X)
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
