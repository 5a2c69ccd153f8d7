"""GPU parity of the RRDBNet x4 forward (C ABI resr_generator_forward behind resr_b200.model.Generator) against the
fp32 oracle. Contract (BASELINE.json north_star): max-abs <= 2e-2 on [0,1] outputs and PSNR >= 45 dB."""
import math
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

MAX_ABS = 2e-2
MIN_PSNR = 45.0


def _psnr(a, b):
    mse = torch.mean((a.double() - b.double()) ** 2).item()
    return 99.0 if mse == 0 else 10 * math.log10(1.0 / mse)


@pytest.fixture(params=["auto", "pairs"])
def kernel_policy(request):
    """"pairs": the CTA-pair conv kernel wherever two column groups exist, also at small / awkward shapes."""
    import resr_b200
    lib = resr_b200._lib.lib()
    prev = lib.resr_set_conv_pair_policy(2 if request.param == "pairs" else 1)
    yield request.param
    lib.resr_set_conv_pair_policy(prev)


def _make(seed):
    import resr_b200
    from oracle import generator as og
    sd = og.random_state_dict(seed)
    g = resr_b200.model.Generator(3, 3, 4)
    g.load_state_dict(sd)
    return g.cuda().eval(), sd


@pytest.mark.parametrize("seed,shape", [(0, (1, 3, 32, 32)), (1, (2, 3, 24, 40)), (2, (3, 3, 16, 64)), (3, (1, 3, 20, 128))])
def test_generator_vs_oracle(seed, shape, kernel_policy):
    from oracle import generator as og
    g, sd = _make(seed)
    torch.manual_seed(100 + seed)
    x = torch.rand(*shape)
    ref = og.generator_forward(x, sd)
    with torch.no_grad():
        y = g(x.cuda()).cpu()
    assert y.shape == ref.shape
    err = (y - ref).abs().max().item()
    ps = _psnr(y, ref)
    print(f"seed {seed} shape {shape}: max-abs {err:.3e}  psnr {ps:.2f} dB")
    assert err <= MAX_ABS and ps >= MIN_PSNR


def test_generator_golden_reference_vectors(golden_dir):
    """Committed outputs of the UNMODIFIED reference (oracle/make_golden.py)."""
    path = os.path.join(golden_dir, "generator.npz")
    z = np.load(path)
    for tag in ("a", "b"):
        seed = int(z[f"{tag}_seed"])
        g, _ = _make(seed)
        x = torch.from_numpy(z[f"{tag}_x"])
        ref = torch.from_numpy(z[f"{tag}_y"])
        with torch.no_grad():
            y = g(x.cuda()).cpu()
        err = (y - ref).abs().max().item()
        ps = _psnr(y, ref)
        print(f"golden {tag}: max-abs {err:.3e} psnr {ps:.2f} dB")
        assert err <= MAX_ABS and ps >= MIN_PSNR


def test_generator_cfg1_shape_channels_last_and_modes(golden_dir):
    """cfg1 (1x3x128x128): channels_last input accepted; both activation-load modes agree to fp32 noise."""
    from oracle import generator as og
    z = np.load(os.path.join(golden_dir, "generator.npz"))
    g, sd = _make(0)
    torch.manual_seed(0)
    x = torch.rand(1, 3, 128, 128)
    with torch.no_grad():
        y = g(x.cuda().contiguous(memory_format=torch.channels_last)).cpu()
    sub = torch.from_numpy(z["cfg1_y_sub"])  # reference output, every 4th pixel
    err = (y[:, :, ::4, ::4] - sub).abs().max().item()
    ps = _psnr(y[:, :, ::4, ::4], sub)
    print(f"cfg1: max-abs {err:.3e} psnr {ps:.2f} dB")
    assert err <= MAX_ABS and ps >= MIN_PSNR


def test_tiled_inference_matches_whole_image(kernel_policy):
    """configs[4] mechanism at a small size: halo tiles (with ragged edge tiles) reproduce the whole-image forward."""
    import resr_b200
    g, _ = _make(4)
    torch.manual_seed(9)
    x = torch.rand(1, 3, 72, 200, device="cuda")
    with torch.no_grad():
        full = g(x)
    tiled, mine = resr_b200.model.infer_tiled(g, x, tile_h=32, tile_w=128, halo=16)
    assert tiled.shape == full.shape and len(mine) == 6
    err = (tiled - full).abs().max().item()
    print(f"tiled vs whole: max-abs {err:.3e}")
    assert err <= 5e-3
    # two "ranks" cover the image exactly once between them
    out = torch.zeros_like(full)
    for r in range(2):
        resr_b200.model.infer_tiled(g, x, 32, 128, 16, rank=r, world=2, out=out)
    assert torch.equal(out, tiled)


@pytest.mark.gpu
def test_pipelined_host_calls_match_blocking_call():
    """resr_generator_forward_host_async: rotating staging slots, copies overlapping compute; results equal the blocking
    host call bit for bit, in order."""
    import resr_b200
    g, _ = _make(5)
    torch.manual_seed(3)
    xs = [torch.rand(2, 3, 24, 136).pin_memory() for _ in range(5)]
    ref = [g.infer_host(x).clone() for x in xs]
    outs = [torch.empty(2, 3, 96, 544).pin_memory() for _ in range(5)]
    for x, o in zip(xs, outs):
        g.infer_host_async(x, o)
    g.host_sync()
    for r, o in zip(ref, outs):
        assert torch.equal(r, o)
    with pytest.raises(resr_b200._lib.ResrError):
        g.infer_host_async(torch.rand(1, 3, 8, 8), torch.empty(1, 3, 32, 32))  # not pinned


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(1, 3, 1, 1), (1, 3, 2, 3), (1, 3, 5, 7), (2, 3, 9, 130), (1, 3, 3, 257), (5, 3, 8, 8),
                                   (3, 3, 7, 16), (17, 3, 6, 64), (1, 3, 1, 128), (2, 3, 130, 9)])
def test_generator_awkward_shapes(shape, kernel_policy):
    """Single-pixel images, one-row images, ragged widths (W = 130, 257: a 2 / 1 pixel last segment), every lane split
    (W = 8, 16, 64 -> 16, 8, 2 images per M tile) with image counts that do not fill the last tile."""
    from oracle import generator as og
    g, sd = _make(7)
    torch.manual_seed(sum(shape))
    x = torch.rand(*shape)
    with torch.no_grad():
        y = g(x.cuda()).cpu()
    ref = og.generator_forward(x, sd)
    err = (y - ref).abs().max().item()
    psnr = 10 * np.log10(1.0 / max(torch.mean((y - ref) ** 2).item(), 1e-20))
    assert err <= 2e-2 and psnr >= 45.0, (err, psnr)


def test_weight_swap_through_param_data_is_seen():
    """`param.data = t` (EMA.apply_shadow / restore, reference model.py:51-61) does not bump `_version`; the
    packed tensor-core weights must follow anyway."""
    from oracle import generator as og
    g, sd = _make(0)
    sd2 = og.random_state_dict(1)
    torch.manual_seed(5)
    x = torch.rand(1, 3, 16, 24)
    with torch.no_grad():
        y0 = g(x.cuda()).cpu()
        backup = {n: p.data for n, p in g.named_parameters()}
        for n, p in g.named_parameters():      # apply_shadow
            p.data = sd2[n].cuda()
        y1 = g(x.cuda()).cpu()
        for n, p in g.named_parameters():      # restore
            p.data = backup[n]
        y2 = g(x.cuda()).cpu()
    assert (y1 - og.generator_forward(x, sd2)).abs().max().item() <= MAX_ABS
    assert not torch.equal(y0, y1) and torch.equal(y0, y2)
    # frozen-weights serving mode skips the walk; invalidate() forces the repack
    g.assume_static_weights(True)
    with torch.no_grad():
        for n, p in g.named_parameters():
            p.data = sd2[n].cuda()
        assert torch.equal(g(x.cuda()).cpu(), y0)
        g.invalidate()
        assert torch.equal(g(x.cuda()).cpu(), y1)


def test_dense_block_forwards_vs_oracle():
    """ResidualDenseBlock.forward / ResidualResidualDenseBlock.forward (reference model.py:87-98, 123-132) through
    resr_conv3x3, against the fp32 oracle of the same blocks."""
    from oracle import generator as og
    g, sd = _make(2)
    torch.manual_seed(8)
    x = torch.randn(2, 64, 20, 136) * 0.5
    rrdb = g.trunk[3]
    with torch.no_grad():
        y_rdb = rrdb.rdb2(x.cuda()).cpu()
        y_rrdb = rrdb(x.cuda()).cpu()
    ref_rdb = og.rdb_forward(x, sd, "trunk.3.rdb2")
    ref_rrdb = og.rrdb_forward(x, sd, "trunk.3")
    e1, e2 = (y_rdb - ref_rdb).abs().max().item(), (y_rrdb - ref_rrdb).abs().max().item()
    print(f"rdb max-abs {e1:.2e}, rrdb max-abs {e2:.2e} (|x| up to {x.abs().max():.1f})")
    assert e1 <= 4e-3 and e2 <= 4e-3
    with pytest.raises(ValueError):
        rrdb(torch.rand(1, 3, 8, 8, device="cuda"))


def test_input_gradient_request_raises():
    import resr_b200
    g, _ = _make(0)
    x = torch.rand(1, 3, 8, 8, device="cuda", requires_grad=True)
    with pytest.raises(resr_b200._lib.ResrError):
        g(x)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs in one process")
def test_second_device_first_in_one_process():
    """Per-device tables / function attributes: run everything on cuda:1 BEFORE cuda:0 has been
    touched by the library, with cuda:0 the current device, then on cuda:0, and compare."""
    import resr_b200
    from oracle import generator as og
    ip = resr_b200.imgproc
    sd = og.random_state_dict(3)
    rng = np.random.default_rng(0)
    img = torch.from_numpy(rng.random((2, 3, 64, 80), dtype=np.float32))
    q = torch.tensor([35.0, 80.0])
    x = torch.rand(1, 3, 16, 24)
    outs = []
    for dev in ("cuda:1", "cuda:0"):
        g = resr_b200.model.Generator(3, 3, 4)
        g.load_state_dict(sd)
        g = g.to(dev).eval()
        with torch.no_grad():
            y = g(x.to(dev))
        u = ip.USMSharp(50, 0)(img.to(dev), 0.5, 10)
        j = ip.DiffJPEG(False)(img.to(dev), q.to(dev).clone())
        k = torch.zeros(1, 21, 21, device=dev)
        k[0, 8:13, 8:13] = 1 / 25
        f = ip.filter2d_torch(img.to(dev), k)
        assert y.device == torch.device(dev) and u.device == torch.device(dev)
        outs.append([t.cpu() for t in (y, u, j, f)])
    for a, b in zip(*outs):
        assert torch.equal(a, b)
    assert (outs[0][0] - og.generator_forward(x, sd)).abs().max().item() <= MAX_ABS
    assert outs[0][1].abs().sum().item() > 0 and outs[0][2].abs().sum().item() > 0
    # a generator moved between devices re-creates its handle and repacks
    g = g.to("cuda:1")
    with torch.no_grad():
        assert torch.equal(g(x.to("cuda:1")).cpu(), outs[0][0])


@pytest.mark.parametrize("policy", [1, 2])
def test_bf16_recipe_within_contract(policy):
    """north_star's recipe as a switch: bf16 operands + fp32 residual masters (set_precision("bf16")); both recipes meet
    the contract, fp16 is the closer one."""
    import resr_b200
    from oracle import generator as og
    prev = resr_b200._lib.lib().resr_set_conv_pair_policy(policy)
    try:
        g, sd = _make(1)
        torch.manual_seed(21)
        x = torch.rand(2, 3, 40, 136)
        ref = og.generator_forward(x, sd)
        with torch.no_grad():
            y16 = g(x.cuda()).cpu()
            g.set_precision("bf16")
            ybf = g(x.cuda()).cpu()
            g.set_precision("fp16")
            y16b = g(x.cuda()).cpu()
    finally:
        resr_b200._lib.lib().resr_set_conv_pair_policy(prev)
    e16, ebf = (y16 - ref).abs().max().item(), (ybf - ref).abs().max().item()
    print(f"fp16 max-abs {e16:.3e} / {_psnr(y16, ref):.1f} dB; bf16 max-abs {ebf:.3e} / {_psnr(ybf, ref):.1f} dB")
    assert ebf <= MAX_ABS and _psnr(ybf, ref) >= MIN_PSNR
    assert e16 <= MAX_ABS and torch.equal(y16, y16b)
    assert not torch.equal(y16, ybf)
    g.set_precision("bf16")   # the training path runs in either recipe (tests/test_train_gpu.py holds the parity tests)
    loss, _, flat = resr_b200.autograd.l1_loss_backward(g, x[:, :, :8, :8].cuda(), torch.rand(2, 3, 32, 32).cuda())
    assert torch.isfinite(loss) and torch.isfinite(flat).all()


def test_u8_image_io_is_the_reference_conversion_fused():
    """f4: image / 255 -> image_to_tensor fused into the first kernel, tensor_to_image (mul(255).clamp(0, 255), uint8
    truncation; reference imgproc.py:1540-1596, inference.py:40-56) into the last convolution: bit-exact against applying
    the reference's own conversions around the fp32 forward."""
    import resr_b200
    ip = resr_b200.imgproc
    g, _ = _make(6)
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (2, 40, 136, 3), dtype=np.uint8)
    with torch.no_grad():
        y_u8 = g.infer_u8(torch.from_numpy(img).cuda())
        # the reference's path around the same generator: numpy / 255 -> image_to_tensor -> model -> tensor_to_image.
        # (Same batch as the fused call: the forward is deterministic per shape, but fp32 accumulation order -- hence the
        # last fp16 bit of an activation here and there -- depends on how rows map to accumulator slots, i.e. on the tiling.)
        lr = torch.stack([ip.image_to_tensor(img[i].astype(np.float32) / 255.0, False, False) for i in range(2)])
        sr = g(lr.cuda())
        ref = []
        for i in range(2):
            ref.append(sr[i].permute(1, 2, 0).mul(255).clamp(0, 255).cpu().numpy().astype("uint8"))
            assert np.array_equal(ip.tensor_to_image(sr[i:i + 1], False, False), ref[-1])          # device conversion
            assert np.array_equal(ip.tensor_to_image(sr[i:i + 1].cpu(), False, False), ref[-1])    # host path
    assert y_u8.shape == (2, 160, 544, 3) and y_u8.dtype == torch.uint8
    assert np.array_equal(y_u8.cpu().numpy(), np.stack(ref))
    y_host = g.infer_u8_host(torch.from_numpy(img).pin_memory())
    assert np.array_equal(y_host.numpy(), np.stack(ref))
    # range_norm / half variants of tensor_to_image against the reference expression
    t = torch.rand(1, 3, 17, 23, device="cuda") * 2 - 1
    for rn, hf in ((True, False), (False, True), (True, True)):
        tt = t.add(1.0).div(2.0) if rn else t
        tt = tt.half() if hf else tt
        want = tt.squeeze(0).permute(1, 2, 0).mul(255).clamp(0, 255).cpu().numpy().astype("uint8")
        assert np.array_equal(ip.tensor_to_image(t, rn, hf), want), (rn, hf)
