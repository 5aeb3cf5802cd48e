"""Where does an end-to-end (host-buffer) MNIST step spend its time?  Run on a GPU box."""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "active-perception-gym_b200"))
import numpy as np, torch
import ap_gym_b200 as apg
from ap_gym_b200 import _capi

n, pw = 65536, 10
ds = apg.SyntheticImageDataset("mnist", seed=0)
env = apg.ImageClassificationVectorEnv(n, apg.ImagePerceptionConfig(dataset=ds), check_errors="never", payload_ring=64)
env.reset(seed=0)
lib = _capi.lib()
h_a = torch.rand((n, 2)).pin_memory(); h_p = torch.randn((n, pw)).pin_memory()
h_r = torch.empty((n,), dtype=torch.float32).pin_memory()
h_te = torch.empty((n,), dtype=torch.uint8).pin_memory(); h_tr = torch.empty((n,), dtype=torch.uint8).pin_memory()
d_a = h_a.cuda(); d_p = h_p.cuda()
def host_step():
    _capi.check(lib.apg_image_step_host(env._h, h_a.data_ptr(), h_p.data_ptr(), h_r.data_ptr(), h_te.data_ptr(), h_tr.data_ptr(), None))
def dev_step():
    _capi.check(lib.apg_image_step(env._h, d_a.data_ptr(), d_p.data_ptr())); _capi.check(lib.apg_image_sync(env._h))
for f, name in ((host_step, "step_host"), (dev_step, "step(device ptrs)+sync")):
    for _ in range(20): f()
    ts = []
    for _ in range(170):
        t0 = time.perf_counter(); f(); ts.append(time.perf_counter() - t0)
    ts = np.array(ts) * 1e6
    print(f"{name:28s} mean {ts.mean():7.1f} us  median {np.median(ts):7.1f}  p90 {np.percentile(ts,90):7.1f}  max {ts.max():8.1f}  (sorted top5 {np.sort(ts)[-5:].round(0)})")
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for name, src, dst in (("H2D action", h_a, d_a), ("H2D pred", h_p, d_p)):
        for _ in range(5): dst.copy_(src, non_blocking=True)
        s.synchronize(); t0 = time.perf_counter()
        for _ in range(50): dst.copy_(src, non_blocking=True); s.synchronize()
        print(f"{name:28s} {(time.perf_counter()-t0)/50*1e6:7.1f} us  ({src.numel()*4/1e6:.2f} MB)")
    d_r = torch.empty((n,), device="cuda")
    s.synchronize(); t0 = time.perf_counter()
    for _ in range(50): h_r.copy_(d_r, non_blocking=True); s.synchronize()
    print(f"{'D2H reward':28s} {(time.perf_counter()-t0)/50*1e6:7.1f} us")
