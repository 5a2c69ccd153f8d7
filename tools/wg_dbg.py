import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import resr_b200
L = resr_b200._lib
n, h, w, cin, c_total, cout = 2, 12, 64, 64, 64, 32
dev = "cuda"
x16 = torch.randn(n, h, w, c_total, device=dev).bfloat16()
dy16 = torch.zeros(n, h, w, 64, device=dev, dtype=torch.bfloat16); dy16[..., :cout] = torch.randn(n, h, w, cout, device=dev).bfloat16()
dw = torch.zeros(cout, cin, 3, 3, device=dev); db = torch.zeros(cout, device=dev)
need = L.lib().resr_conv3x3_wgrad_workspace_bytes(n, h, w, cin, cout)
ws = torch.empty(need + 1024, dtype=torch.uint8, device=dev)
wp = ws.data_ptr() + (-ws.data_ptr()) % 1024
rc = L.lib().resr_conv3x3_wgrad(L.ptr(x16), c_total, 1, L.ptr(dy16), n, h, w, cin, cout, L.ptr(dw), L.ptr(db), ctypes.c_void_p(wp), need, L.stream_ptr())
print("rc", rc)
