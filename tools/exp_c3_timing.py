"""Where does the c3 step time go?  Events around the whole step vs around each kernel."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
import ipfa_b200 as ipfa

wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c3"]
dev = torch.device("cuda:0")
inputs = wl.make(0, device=dev)
if len(sys.argv) > 2:  # L2 fetch granularity experiment (bytes)
    import ctypes
    cu = ctypes.CDLL("libcuda.so.1")
    print("cuCtxSetLimit(MAX_L2_FETCH_GRANULARITY,", sys.argv[2], ") ->",
          cu.cuCtxSetLimit(0x05, ctypes.c_size_t(int(sys.argv[2]))))
def step():
    return wl.step(ipfa, inputs)
for _ in range(3): step()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20): step()
b.record(); torch.cuda.synchronize()
print("loop of 20 steps:", a.elapsed_time(b) / 20, "ms/step")
# one step at a time, synchronised
ts = []
for _ in range(10):
    torch.cuda.synchronize(); a.record(); step(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
print("single synchronised steps:", [round(x, 3) for x in ts])
t0 = time.perf_counter()
for _ in range(20): step()
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print("host time to issue 20 steps:", (t1 - t0) / 20 * 1e3, "ms/step; drained after", (t2 - t1) * 1e3, "ms")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(5): step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=8, max_name_column_width=60))
