"""Phase timestamps of CTA (0,0) of the conv kernel for RDB-like layer shapes at cfg4 (development aid)."""
import sys, os, ctypes
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import torch
import resr_b200
L = resr_b200._lib
dev = "cuda"
n, h, w = 16, 64, 64
x16 = torch.randn(n, h, w, 192, device=dev).bfloat16()
names = ["entry", "setup", "depwait", "weights", "first_stage", "mma_done", "epi_done", "cta_done"]
for cin, cout, mode, ct, fl in [(64, 32, "o16", 192, 0), (96, 32, "o16", 192, 0), (128, 32, "o16", 192, 0), (160, 32, "o16", 192, 0), (192, 64, "rdb", 192, 0)]:
    x16 = torch.randn(n, h, w, ct, device=dev).bfloat16()
    wt = (torch.randn(cout, cin, 3, 3, device=dev) * 0.05)
    b = torch.randn(cout, device=dev)
    o16 = torch.zeros(n, h, w, 192, device=dev, dtype=torch.bfloat16)
    dbg = torch.zeros(16, dtype=torch.int64, device=dev)
    d = L.ConvDesc()
    d.in16 = x16.data_ptr(); d.n, d.h, d.w, d.c_total, d.cin, d.cout = n, h, w, ct, cin, cout
    d.fmt_in, d.mode = 1, -1
    d.weight, d.bias = wt.data_ptr(), b.data_ptr()
    d.lrelu = 1
    d.out16, d.out16_fmt, d.out16_cstride, d.out16_choff = o16.data_ptr(), 1, 192, 64
    if mode == "o16dense":
        o16 = torch.zeros(n, h, w, 32, device=dev, dtype=torch.bfloat16)
        d.out16, d.out16_fmt, d.out16_cstride, d.out16_choff = o16.data_ptr(), 1, 32, 0
    if mode == "rdb":
        res = torch.randn(n, h, w, 64, device=dev); outf = torch.zeros(n, h, w, 64, device=dev)
        d.ep_mode = 1; d.lrelu = 0; d.res1 = res.data_ptr(); d.res_cstride = 64; d.outf = outf.data_ptr(); d.outf_cstride = 64
        d.out16_choff = 0
    d.dbg = dbg.data_ptr(); d.dbg_flags = fl
    for it in range(3):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        L.check(L.lib().resr_conv3x3(ctypes.byref(d), L.stream_ptr()))
    torch.cuda.synchronize()
    e0.record()
    for it in range(20):
        L.check(L.lib().resr_conv3x3(ctypes.byref(d), L.stream_ptr()))
    e1.record(); torch.cuda.synchronize()
    print(f"   back-to-back: {e0.elapsed_time(e1)/20*1e3:.1f} us per launch")
    t = dbg.cpu().tolist()
    base = t[0]
    print(f"ct={ct} cin={cin} cout={cout} {mode} flags={fl}: " + "  ".join(f"{nm}={(t[i]-base)/1e3:.1f}us" for i, nm in enumerate(names)))
