"""GPU parity of the tcgen05 3x3 convolution (C ABI resr_conv3x3) against a plain torch fp32 convolution of the same
16-bit-rounded operands. Tolerance: fp32 accumulation-order noise only (operands are exactly representable)."""
import ctypes

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _nhwc16(x, c_total, dtype):
    n, c, h, w = x.shape
    out = torch.randn(n, h, w, c_total, device=x.device).to(dtype)  # finite garbage beyond the used channels
    out[..., :c] = x.permute(0, 2, 3, 1).to(dtype)
    return out.contiguous()


def _run_conv(x16, cin, weight, bias, *, fmt, mode=-1, ep_mode=0, lrelu=0, clamp01=0, out16=None, out16_fmt=1,
              out16_choff=0, out16_up2=0, outf=None, outf_choff=0, res1=None, res2=None, out_nchw=None, res16_fmt=None):
    import resr_b200
    L = resr_b200._lib
    n, h, w, c_total = x16.shape
    d = L.ConvDesc()
    d.in16 = x16.data_ptr()
    d.n, d.h, d.w, d.c_total, d.cin, d.cout = n, h, w, c_total, cin, weight.shape[0]
    d.fmt_in, d.mode = fmt, mode
    d.weight = weight.data_ptr()
    d.bias = bias.data_ptr() if bias is not None else None
    d.ep_mode, d.lrelu, d.clamp01 = ep_mode, lrelu, clamp01
    if out16 is not None:
        d.out16, d.out16_fmt, d.out16_cstride, d.out16_choff, d.out16_up2 = (
            out16.data_ptr(), out16_fmt, out16.shape[-1], out16_choff, out16_up2)
    if outf is not None:
        d.outf, d.outf_cstride, d.outf_choff = outf.data_ptr(), outf.shape[-1], outf_choff
    if res1 is not None:
        d.res1, d.res_cstride, d.res_choff = res1.data_ptr(), res1.shape[-1], 0
    if res2 is not None:
        d.res2 = res2.data_ptr()
    if res16_fmt is not None:
        d.res16, d.res16_fmt = 1, res16_fmt
    if out_nchw is not None:
        d.out_nchw, d.out_nchw_c = out_nchw.data_ptr(), out_nchw.shape[1]
    L.check(L.lib().resr_conv3x3(ctypes.byref(d), L.stream_ptr()))
    torch.cuda.synchronize()


@pytest.fixture(params=["auto", "pairs"], autouse=True)
def conv_kernel_policy(request):
    """Every case runs twice: default kernel choice (small shapes take the single-CTA kernel) and with the CTA-pair
    kernel forced wherever two column groups exist (conv3x3_pair.cu, tcgen05 cta_group::2)."""
    import resr_b200
    lib = resr_b200._lib.lib()
    prev = lib.resr_set_conv_pair_policy(2 if request.param == "pairs" else 1)
    yield request.param
    lib.resr_set_conv_pair_policy(prev)


CASES = [
    # n, h, w, cin, c_total, cout, mode
    (2, 12, 128, 64, 64, 32, 0),
    (2, 12, 128, 64, 64, 32, 1),
    (1, 9, 128, 96, 192, 32, 0),     # zero-padded second K chunk reads finite garbage channels
    (1, 9, 128, 160, 192, 32, 1),
    (2, 7, 128, 192, 192, 64, 0),    # two Cout slices, three K chunks
    (3, 10, 64, 64, 64, 32, -1),     # two images per M tile, odd image count
    (5, 6, 32, 64, 64, 64, -1),      # four images per M tile
    (1, 5, 200, 64, 64, 32, 0),      # ragged second x segment
    (1, 5, 200, 64, 64, 32, 1),
    (2, 1, 128, 64, 64, 32, 0),      # single-row image
    (1, 150, 256, 128, 128, 32, 0),  # many rows: exercises every CTA range split
]


@pytest.mark.parametrize("n,h,w,cin,c_total,cout,mode", CASES)
@pytest.mark.parametrize("fmt", [1, 0])
def test_conv_plain(n, h, w, cin, c_total, cout, mode, fmt):
    torch.manual_seed(n * 1000 + h * 10 + w + cin + cout)
    dev = "cuda"
    dt = torch.bfloat16 if fmt == 1 else torch.float16
    x = torch.randn(n, cin, h, w, device=dev).to(dt).float()
    wt = (torch.randn(cout, cin, 3, 3, device=dev) * 0.05).to(dt).float()
    b = torch.randn(cout, device=dev)
    x16 = _nhwc16(x, c_total, dt)
    outf = torch.full((n, h, w, cout), float("nan"), device=dev)
    _run_conv(x16, cin, wt, b, fmt=fmt, mode=mode, outf=outf)
    ref = F.conv2d(x, wt, b, padding=1).permute(0, 2, 3, 1)
    err = (outf - ref).abs().max().item()
    assert torch.isfinite(outf).all(), "kernel left unwritten / non-finite outputs"
    assert err < 2e-4, f"max abs err {err}"


def test_conv_epilogues():
    torch.manual_seed(7)
    dev = "cuda"
    n, h, w, cin, cout = 2, 10, 128, 192, 64
    x = torch.randn(n, cin, h, w, device=dev).bfloat16().float()
    wt = (torch.randn(cout, cin, 3, 3, device=dev) * 0.03).bfloat16().float()
    b = torch.randn(cout, device=dev) * 0.1
    x16 = _nhwc16(x, cin, torch.bfloat16)
    conv = F.conv2d(x, wt, b, padding=1).permute(0, 2, 3, 1).contiguous()
    r1 = torch.randn(n, h, w, cout, device=dev)
    r2 = torch.randn(n, h, w, cout, device=dev)
    # rdb: 0.2 * v + r1, fp32 + bf16 copy at a channel offset
    outf = torch.zeros(n, h, w, cout, device=dev)
    o16 = torch.zeros(n, h, w, 192, device=dev, dtype=torch.bfloat16)
    _run_conv(x16, cin, wt, b, fmt=1, ep_mode=1, res1=r1, outf=outf, out16=o16, out16_choff=64)
    ref = conv * 0.2 + r1
    assert (outf - ref).abs().max().item() < 2e-4
    assert torch.equal(o16[..., 64:128], outf.bfloat16())
    assert o16[..., :64].abs().max().item() == 0 and o16[..., 128:].abs().max().item() == 0
    # rrdb
    outf.zero_()
    _run_conv(x16, cin, wt, b, fmt=1, ep_mode=2, res1=r1, res2=r2, outf=outf)
    assert (outf - ((conv * 0.2 + r1) * 0.2 + r2)).abs().max().item() < 2e-4
    # skip add + 2x nearest-upsampled fp16 store
    up = torch.zeros(n, 2 * h, 2 * w, cout, device=dev, dtype=torch.float16)
    _run_conv(x16, cin, wt, b, fmt=1, ep_mode=3, res1=r1, out16=up, out16_fmt=0, out16_up2=1)
    ref_up = (r1 + conv).half().repeat_interleave(2, 1).repeat_interleave(2, 2)
    # fp16 ulp of the result plus fp32 accumulation-order noise (which dominates where r1 + conv cancels)
    assert ((up.float() - ref_up.float()).abs() <= ref_up.float().abs() * 2 ** -9 + 2e-4).all()
    # lrelu
    outf.zero_()
    _run_conv(x16, cin, wt, b, fmt=1, lrelu=1, outf=outf)
    assert (outf - F.leaky_relu(conv, 0.2)).abs().max().item() < 2e-4


@pytest.mark.parametrize("fmt", [0, 1])
def test_conv_16bit_residual_stream_in_place(fmt):
    """Inference trunk epilogues: the residuals are 16-bit NHWC tensors (the conv inputs themselves) and the RRDB output
    overwrites the buffer that holds res2 (model.py:94-96, 129-130)."""
    torch.manual_seed(11 + fmt)
    dev = "cuda"
    dt = torch.float16 if fmt == 0 else torch.bfloat16
    n, h, w, cin, cout = 2, 37, 128, 192, 64
    x = torch.randn(n, cin, h, w, device=dev).to(dt).float()
    wt = (torch.randn(cout, cin, 3, 3, device=dev) * 0.03).to(dt).float()
    b = torch.randn(cout, device=dev) * 0.1
    x16 = _nhwc16(x, cin, dt)
    conv = F.conv2d(x, wt, b, padding=1).permute(0, 2, 3, 1).contiguous()
    r1 = torch.randn(n, h, w, 192, device=dev).to(dt)
    r2buf = torch.randn(n, h, w, 192, device=dev).to(dt)
    r2 = r2buf.clone()
    # rdb: 0.2 * v + r1[..., :64]
    o16 = torch.zeros(n, h, w, 192, device=dev, dtype=dt)
    _run_conv(x16, cin, wt, b, fmt=fmt, ep_mode=1, res1=r1, out16=o16, out16_fmt=fmt, res16_fmt=fmt)
    ref = conv * 0.2 + r1[..., :64].float()
    assert torch.equal(o16[..., :64], ref.to(dt)) or (o16[..., :64].float() - ref).abs().max().item() < 2e-2
    assert ((o16[..., :64].float() - ref).abs() <= ref.abs() * 2.0 ** (-8 if fmt else -11) + 2e-4).all()
    # rrdb, output written over the res2 buffer
    _run_conv(x16, cin, wt, b, fmt=fmt, ep_mode=2, res1=r1, res2=r2buf, out16=r2buf, out16_fmt=fmt, res16_fmt=fmt)
    ref = (conv * 0.2 + r1[..., :64].float()) * 0.2 + r2[..., :64].float()
    assert ((r2buf[..., :64].float() - ref).abs() <= ref.abs() * 2.0 ** (-8 if fmt else -11) + 2e-4).all()
    assert torch.equal(r2buf[..., 64:], r2[..., 64:])


def test_conv_rgb_out_clamped():
    torch.manual_seed(9)
    dev = "cuda"
    n, h, w, cin, cout = 2, 16, 256, 64, 3
    x = torch.randn(n, cin, h, w, device=dev).half().float()
    wt = (torch.randn(cout, cin, 3, 3, device=dev) * 0.05).half().float()
    b = torch.rand(cout, device=dev)
    x16 = _nhwc16(x, cin, torch.float16)
    y = torch.full((n, cout, h, w), -5.0, device=dev)
    _run_conv(x16, cin, wt, b, fmt=0, clamp01=1, out_nchw=y)
    ref = F.conv2d(x, wt, b, padding=1).clamp(0, 1)
    assert (y - ref).abs().max().item() < 2e-4


@pytest.mark.parametrize("n,h,w,cin,c_total,cout", [(2, 40, 128, 64, 192, 32), (1, 33, 256, 128, 128, 64), (3, 9, 64, 64, 64, 32)])
def test_conv_out16_only_two_epilogue_groups(n, h, w, cin, c_total, cout):
    """16-bit NHWC output only: the configuration that runs two epilogue groups on alternating rows."""
    torch.manual_seed(h + w)
    dev = "cuda"
    x = torch.randn(n, cin, h, w, device=dev).bfloat16().float()
    wt = (torch.randn(cout, cin, 3, 3, device=dev) * 0.05).bfloat16().float()
    b = torch.randn(cout, device=dev)
    x16 = _nhwc16(x, c_total, torch.bfloat16)
    o16 = torch.zeros(n, h, w, 192, device=dev, dtype=torch.bfloat16)
    _run_conv(x16, cin, wt, b, fmt=1, lrelu=1, out16=o16, out16_choff=96)
    ref = F.leaky_relu(F.conv2d(x, wt, b, padding=1), 0.2).permute(0, 2, 3, 1)
    got = o16[..., 96:96 + cout].float()
    assert ((got - ref).abs() <= ref.abs() * 2 ** -7 + 2e-4).all()
    assert o16[..., :96].abs().max().item() == 0 and o16[..., 96 + cout:].abs().max().item() == 0
