"""GPU parity of the 2-layer MLP predictor (BASELINE.json configs[4]) against the oracle and torch itself."""
import numpy as np
import pandas as pd
import pytest
import torch

from oracle import mlp as omlp

pytestmark = pytest.mark.gpu
if not torch.cuda.is_available():
    pytest.skip("needs a CUDA device", allow_module_level=True)

from tests.conftest import GOLDEN  # noqa: E402


@pytest.fixture(scope="module")
def engine():
    from unionml_b200.engine import Engine

    return Engine(0)


@pytest.fixture(scope="module")
def mlp_golden():
    z = np.load(GOLDEN / "mlp_64_32_10.npz")
    return {k: z[k] for k in z.files}


def _weights(g):
    return g["w1"], g["b1"], g["w2"], g["b2"]


def test_golden_fixture_equals_torch_labels(engine, mlp_golden):
    m = engine.load_mlp(*_weights(mlp_golden))
    b = engine.stage(mlp_golden["X"])
    got, st = engine.predict_mlp(m, b, exact=True)
    np.testing.assert_array_equal(got, mlp_golden["labels_torch"])
    assert st["path"] == 3 and st["kernel_launches"] == 2


@pytest.mark.parametrize("rows", [1, 127, 128, 129, 5000, 250_001])
def test_exact_parity_against_float64_network(engine, mlp_golden, rows):
    w = _weights(mlp_golden)
    X = np.random.default_rng(rows).integers(0, 17, size=(rows, 64), dtype=np.uint8).astype(np.float32)
    want = omlp.predict_indices_f64(X, *w).astype(np.int32)
    m = engine.load_mlp(*w)
    got, st = engine.predict_mlp(m, engine.stage(X), exact=True)
    np.testing.assert_array_equal(got, want)
    assert st["n_ambiguous"] == 0


def test_fast_mode_differs_only_inside_the_bound(engine, mlp_golden):
    w = _weights(mlp_golden)
    X = np.random.default_rng(9).standard_normal((1_000_000, 64)).astype(np.float32)
    m = engine.load_mlp(*w)
    b = engine.stage(X)
    exact, st = engine.predict_mlp(m, b, exact=True)
    want = omlp.predict_indices_f64(X, *w).astype(np.int32)
    np.testing.assert_array_equal(exact, want)
    assert 0 <= st["n_flagged"] < 20_000
    fast, stf = engine.predict_mlp(m, b, exact=False)
    assert stf["kernel_launches"] == 1
    diff = np.flatnonzero(fast != want)
    margin = omlp.logit_margin_f64(X, *w)
    assert len(diff) <= st["n_flagged"] and (len(diff) == 0 or margin[diff].max() < 1e-3)


def test_against_torch_cpu_module_and_predictor(mlp_golden):
    """The reference predictor as written vs the drop-in device predictor, on a DataFrame."""
    import torch.nn as nn
    import torch.nn.functional as F

    from unionml_b200.predictors import mlp_argmax

    class PytorchModel(nn.Module):  # tests/integration/pytorch_app/quickstart.py:14-24
        def __init__(self, in_dims, hidden_dims, out_dims):
            super().__init__()
            self.layers = nn.Sequential(nn.Linear(in_dims, hidden_dims), nn.ReLU(), nn.Linear(hidden_dims, out_dims))

        def forward(self, features):
            return F.softmax(self.layers(features), dim=1)

    torch.manual_seed(0)
    module = PytorchModel(64, 32, 10)
    np.testing.assert_array_equal(module.layers[0].weight.detach().numpy(), mlp_golden["w1"])
    frame = pd.DataFrame(np.random.default_rng(4).integers(0, 17, size=(20_000, 64)).astype(np.float64))
    with torch.no_grad():
        want = [float(x) for x in module(torch.from_numpy(frame.values).float()).argmax(1)]
    got = mlp_argmax(module, frame)
    assert all(isinstance(x, float) for x in got)
    mism = [i for i, (a, b) in enumerate(zip(got, want)) if a != b]
    # torch's own fp32 forward is only trusted outside its rounding noise; inside it the float64 network decides
    margin = omlp.logit_margin_f64(frame.values, *_weights(mlp_golden))
    assert all(margin[i] < 1e-4 for i in mism) and len(mism) <= 2


def test_generic_shapes_take_the_fp64_kernel(engine):
    rng = np.random.default_rng(0)
    w1, b1 = rng.standard_normal((20, 13)).astype(np.float32), rng.standard_normal(20).astype(np.float32)
    w2, b2 = rng.standard_normal((7, 20)).astype(np.float32), rng.standard_normal(7).astype(np.float32)
    X = rng.standard_normal((30_000, 13)).astype(np.float32)
    m = engine.load_mlp(w1, b1, w2, b2)
    got, st = engine.predict_mlp(m, engine.stage(X), exact=True)
    np.testing.assert_array_equal(got, omlp.predict_indices_f64(X, w1, b1, w2, b2))
    assert st["path"] == 2


def test_mlp_nonfinite_and_shape_errors(engine, mlp_golden):
    m = engine.load_mlp(*_weights(mlp_golden))
    X = np.ones((1000, 64), dtype=np.float32)
    X[5, 5] = np.nan
    with pytest.raises(ValueError):
        engine.predict_mlp(m, engine.stage(X, check_finite=False), exact=True)
    with pytest.raises(ValueError, match="63 features"):
        engine.predict_mlp(m, engine.stage(np.ones((4, 63), dtype=np.float32)))
