import os, sys, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from unionml_b200.engine import Engine
z = np.load("tests/golden/digits_lr.npz")
eng = Engine(0)
m = eng.load_linear(z["coef"], z["intercept"])
X = np.random.default_rng(0).integers(0, 17, size=(3000, 64), dtype=np.uint8)
cases = {
 "f32_c_stage": lambda: eng.stage(X.astype(np.float32)),
 "f64_c_stage": lambda: eng.stage(X.astype(np.float64)),
 "f64_f_stage": lambda: eng.stage(np.asfortranarray(X.astype(np.float64))),
 "u8_stage": lambda: eng.stage(X),
 "f32_host": lambda: eng.predict_host(m, X.astype(np.float32)),
 "f64_c_host": lambda: eng.predict_host(m, X.astype(np.float64)),
 "f64_f_host": lambda: eng.predict_host(m, np.asfortranarray(X.astype(np.float64))),
 "f64_f_small_host": lambda: eng.predict_host(m, np.asfortranarray(X[:3].astype(np.float64))),
}
for name, fn in cases.items():
    try:
        r = fn()
        print(name, "OK", type(r).__name__, flush=True)
    except Exception as e:
        print(name, "FAIL", repr(e)[:300], flush=True)
