"""Device time of the NHWC weight-gradient chain of one dense block at cfg4 size (development aid)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import resr_b200
L = resr_b200._lib
n, h, w, cin, cout = int(os.environ.get("N", "16")), 64, 64, 192, 192
x16 = torch.randn(n, h, w, cin, device="cuda").bfloat16()
dy16 = torch.randn(n, h, w, cout, device="cuda").bfloat16()
dw = torch.empty(cout, cin, 3, 3, device="cuda"); db = torch.empty(cout, device="cuda")
need = L.lib().resr_conv3x3_wgrad_nhwc_workspace_bytes()
ws = torch.empty(need + 1024, dtype=torch.uint8, device="cuda")
wp = ws.data_ptr() + (-ws.data_ptr()) % 1024
def run():
    L.check(L.lib().resr_conv3x3_wgrad_nhwc(L.ptr(x16), cin, L.ptr(dy16), cout, n, h, w, cin, cout, L.ptr(dw), L.ptr(db),
                                            ctypes.c_void_p(wp), need, L.stream_ptr()))
for _ in range(5): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50): run()
e1.record(); torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(20): run()
    torch.cuda.synchronize()
per = {}
for e in prof.events():
    if e.device_type is not None and "DeviceType.CUDA" in str(e.device_type):
        per.setdefault(e.name[:40], []).append(e.device_time if hasattr(e, "device_time") else e.cuda_time)
print("  per-kernel us:", {k: round(sum(v) / len(v), 1) for k, v in per.items()})
print(f"wgrad nhwc 192x192 full (4 units) at {n}x{h}x{w}: {e0.elapsed_time(e1)/50*1e3:.1f} us per call (GEMM + reduce)")
