"""Error map of the NHWC (MN-major) weight-gradient kernel per tap and per 32-channel block (development aid)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import resr_b200
L = resr_b200._lib


def run(n, h, w, cin, xs, cout, ys, seed=0):
    torch.manual_seed(seed)
    dev = "cuda"
    x = torch.randn(n, cin, h, w, device=dev).bfloat16().float()
    dy = torch.randn(n, cout, h, w, device=dev).bfloat16().float()
    x16 = torch.zeros(n, h, w, xs, device=dev).bfloat16()
    x16[..., :cin] = x.permute(0, 2, 3, 1).bfloat16()
    dy16 = torch.zeros(n, h, w, ys, device=dev).bfloat16()
    dy16[..., :cout] = dy.permute(0, 2, 3, 1).bfloat16()
    dw = torch.full((cout, cin, 3, 3), float("nan"), device=dev)
    db = torch.full((cout,), float("nan"), device=dev)
    need = L.lib().resr_conv3x3_wgrad_nhwc_workspace_bytes()
    ws = torch.empty(need + 1024, dtype=torch.uint8, device=dev)
    wp = ws.data_ptr() + (-ws.data_ptr()) % 1024
    L.check(L.lib().resr_conv3x3_wgrad_nhwc(L.ptr(x16), xs, L.ptr(dy16), ys, n, h, w, cin, cout, L.ptr(dw), L.ptr(db),
                                            ctypes.c_void_p(wp), need, L.stream_ptr()))
    torch.cuda.synchronize()
    ref = torch.nn.grad.conv2d_weight(x.double(), (cout, cin, 3, 3), dy.double(), padding=1).float()
    refb = dy.double().sum((0, 2, 3)).float()
    err = (dw - ref).abs()
    print(f"shape n{n} h{h} w{w} cin{cin}/{xs} cout{cout}/{ys}: max err {err.max().item():.3e} scale {ref.abs().max().item():.3e} "
          f"nan {int(torch.isnan(dw).sum())}; bias err {(db - refb).abs().max().item():.3e}")
    if not (err.max().item() <= 1e-3 * ref.abs().max().item()):
        for dyy in range(3):
            print("  tap errors dy", dyy, [f"{err[:, :, dyy, dx].max().item():.2e}" for dx in range(3)])
        for co0 in range(0, cout, 32):
            print("  co", co0, [f"{err[co0:co0 + 32, ci0:ci0 + 32].max().item():.1e}" for ci0 in range(0, cin, 32)])
        # does the result match another tap / shifted reference?
        for dxs in (-1, 1):
            xr = torch.roll(x, dxs, 3)
            r2 = torch.nn.grad.conv2d_weight(xr.double(), (cout, cin, 3, 3), dy.double(), padding=1).float()
            print(f"  vs x rolled {dxs}: center-tap err {(dw - r2)[:, :, 1, 1].abs().max().item():.2e}")


for cfg in [(1, 4, 64, 64, 64, 64, 64), (2, 12, 64, 64, 64, 32, 64), (2, 16, 64, 192, 192, 64, 64), (2, 6, 150, 192, 192, 192, 192),
            (1, 7, 9, 3, 64, 64, 64)]:
    try:
        run(*cfg)
    except Exception as e:
        print("FAILED", cfg, repr(e))
        break
