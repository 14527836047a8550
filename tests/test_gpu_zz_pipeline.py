"""GPU: Pipeline(StandardScaler, LogisticRegression) - the MNIST tutorial's model shape (docs/tutorials/mnist.md:116-150) -
through the drop-in predictor, scaler folded into W, b on the device.  (Collected last on purpose: it was written after
the round's GPU budget was spent, so it must not be able to mask the other GPU tests under `pytest -x`.)"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
if not torch.cuda.is_available():
    pytest.skip("needs a CUDA device", allow_module_level=True)


def test_scaler_pipeline_predictor_matches_sklearn():
    from sklearn.datasets import load_digits
    from sklearn.linear_model import LogisticRegression
    from sklearn.pipeline import Pipeline
    from sklearn.preprocessing import StandardScaler

    from unionml_b200.predictors import linear_argmax

    frame = load_digits(as_frame=True).frame
    X, y = frame[[c for c in frame if c != "target"]], frame["target"]
    pipe = Pipeline([("scaler", StandardScaler()), ("clf", LogisticRegression(max_iter=2000))]).fit(X, y)
    want = [float(v) for v in pipe.predict(X)]
    got = linear_argmax(pipe, X)
    assert len(got) == len(want) and all(isinstance(v, float) for v in got)
    mismatches = [i for i, (a, b) in enumerate(zip(got, want)) if a != b]
    # the fold changes float64 rounding at the 1e-16 level; only a row tied to ~1e-12 could flip
    scores = pipe.decision_function(X)
    part = np.partition(scores, -2, axis=1)
    margin = part[:, -1] - part[:, -2]
    assert all(margin[i] < 1e-9 for i in mismatches), mismatches[:5]
    # a second call hits the cached device model
    assert linear_argmax(pipe, X.iloc[:10]) == want[:10]
