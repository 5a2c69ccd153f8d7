"""Where do the roles of the CTA-pair conv kernel wait? Runs the five convolutions of a dense block at the cfg3 shape through
resr_conv3x3 with a -DRESR_PROFILE_WAITS build (RESR_LIB_PATH=build/variants/libresr_prof.so) and prints, per layer, the
cycles per pipeline step each role spends blocked. Development aid."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import resr_b200

L = resr_b200._lib
n, h, w = (int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (64, 128, 128)))
dev = "cuda"
torch.manual_seed(0)
x16 = (torch.randn(n, h, w, 192, device=dev) * 0.5).half()
out16 = torch.zeros_like(x16)
names = ["total", "wait stage full", "wait slot free", "steps", "prod total", "prod wait empty", "epi total", "epi wait acc", "epi wait tile"]
for cin, cout in ((64, 32), (96, 32), (128, 32), (160, 32), (192, 64)):
    wt = (torch.randn(cout, cin, 3, 3, device=dev) * 0.05).contiguous()
    b = torch.zeros(cout, device=dev)
    d = L.ConvDesc()
    d.in16 = x16.data_ptr()
    d.n, d.h, d.w, d.c_total, d.cin, d.cout = n, h, w, 192, cin, cout
    d.fmt_in, d.mode = 0, -1
    d.weight, d.bias = wt.data_ptr(), b.data_ptr()
    d.ep_mode, d.lrelu = (1 if cout == 64 else 0), (0 if cout == 64 else 1)
    d.out16, d.out16_fmt, d.out16_cstride, d.out16_choff = out16.data_ptr(), 0, 192, (0 if cout == 64 else 64)
    if cout == 64:
        d.res1, d.res_cstride, d.res16, d.res16_fmt = x16.data_ptr(), 192, 1, 0
    for _ in range(2):
        L.check(L.lib().resr_conv3x3(ctypes.byref(d), L.stream_ptr()))
    torch.cuda.synchronize()
    L.check(L.lib().resr_debug_wait_profile(None, 1))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    ms = 0.0
    for _ in range(reps):
        L.check(L.lib().resr_conv3x3(ctypes.byref(d), L.stream_ptr()))
    torch.cuda.synchronize()
    buf = (ctypes.c_ulonglong * 16)()
    L.check(L.lib().resr_debug_wait_profile(buf, 0))
    v = [float(buf[i]) for i in range(13)]
    steps = max(v[3], 1.0)
    leaders = 74.0
    print(f"conv {cin:3d}->{cout:2d}: per step (cycles): MMA-warp total {v[0] / steps:7.0f} | wait full {v[1] / steps:6.0f} | wait slot {v[2] / steps:6.0f}"
          f" || producer total/step {v[4] / 2 / steps:7.0f} wait-empty {v[5] / 2 / steps:6.0f}"
          f" || epilogue(g0) total/step {v[6] / 2 / steps:7.0f} wait-acc {v[7] / 2 / steps:6.0f} wait-tile {v[8] / 2 / steps:6.0f}"
          f" | steps per launch per leader {steps / reps / leaders:.0f}")
    rows = max(v[12], 1.0)
    print(f"      epilogue per emitted pass of group 0 (cycles): drain (ld + zero + arrive) {v[9] / rows:6.0f} | bias/residual math {v[10] / rows:6.0f}"
          f" | tile wait {v[8] / rows:6.0f} | activation + pack + stage + TMA store {v[11] / rows:6.0f} | accumulator wait {v[7] / rows:6.0f}")
