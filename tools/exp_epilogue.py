"""Debug aid: cost of the fused-exchange epilogue variants on ONE GPU (targets in ordinary device memory)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from unionml_b200.engine import Engine
z = np.load("tests/golden/digits_lr.npz")
dev = torch.device("cuda", 0)
eng = Engine(0); s = torch.cuda.Stream(device=dev); torch.cuda.set_stream(s); eng.set_stream(s.cuda_stream)
m = eng.load_linear(z["coef"], z["intercept"])
rows = 10_000_000
X = eng.pinned_empty((rows, 64), np.float32)
for k in range(10):
    X[k*1_000_000:(k+1)*1_000_000] = np.random.default_rng(k).integers(0, 17, size=(1_000_000, 64), dtype=np.uint8)
b = eng.stage(X)
i32 = torch.empty(rows, dtype=torch.int32, device=dev)
u8a = torch.empty(rows, dtype=torch.uint8, device=dev); u8b = torch.empty(rows, dtype=torch.uint8, device=dev)
i32b = torch.empty(rows, dtype=torch.int32, device=dev)
def t(fn, n=15):
    for _ in range(3): fn()
    ks = [fn()["kernel_ms"] for _ in range(n)]
    return round(sum(ks)/len(ks), 4), round(min(ks), 4)
print("plain int32            ", t(lambda: eng.predict(m, b, exact=True, out_device_ptr=i32.data_ptr())[1]))
print("peers int32 x1 (own)   ", t(lambda: eng.predict_peers(m, b, [i32.data_ptr()], 0, exact=True, want_stats=True, label_bytes=4)))
print("peers int32 x2         ", t(lambda: eng.predict_peers(m, b, [i32.data_ptr(), i32b.data_ptr()], 0, exact=True, want_stats=True, label_bytes=4)))
print("peers u8 x1            ", t(lambda: eng.predict_peers(m, b, [u8a.data_ptr()], 0, exact=True, want_stats=True, label_bytes=1)))
print("peers u8 x2            ", t(lambda: eng.predict_peers(m, b, [u8a.data_ptr(), u8b.data_ptr()], 0, exact=True, want_stats=True, label_bytes=1)))
print("fast plain int32       ", t(lambda: eng.predict(m, b, exact=False, out_device_ptr=i32.data_ptr())[1]))
eng.predict(m, b, exact=True, out_device_ptr=i32.data_ptr()); torch.cuda.synchronize(); assert torch.equal(u8a.to(torch.int32), i32); print("u8 == i32 ok")
