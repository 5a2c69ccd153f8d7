"""GPU parity at BASELINE.json's OWN shapes (VERDICT r01 task 1): the small-fixture tests elsewhere pin the arithmetic,
these pin the launch geometry the benchmark configurations actually use.

  cfg2  16x3x256x256 HR through the second-order degradation: canonical plan S0 + three synth_plan seeds, every draw
        host-fed, each stage fed the ORACLE's stage input (train_realesrnet.py:267-377); per stage <= 1e-5, JPEG
        quantised coefficients compared one by one (1024 + MCUs per launch: the grid-stride loop of jpeg_kernel),
        16 distinct per-sample blur supports; then the whole block end to end on the u8 grid.
  cfg3  64x3x128x128 generator forward, 4 images sampled against the fp32 oracle (model.py:255-272).
  cfg4  TrainStep at 16x3x64x64 -> 16x3x256x256 against fp32 autograd of the oracle (train_realesrnet.py:383-388).
  cfg5  a 512x512 LR image through infer_tiled (halo 16) against the ORACLE on two windows (one straddling a tile
        seam, one in the image corner), not against the repo's own whole-image forward.

Regression guards sit at ~2x the values measured on the round-2 B200 runs (printed by every test), far inside the
contract (generator 2e-2 / 45 dB, degradation 1e-5)."""
import math
import random

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

TOL = 1e-5            # north_star: degradation within 1e-5 in fp32
GEN_GUARD = 6e-3      # 2x the measured 3.0e-3 (contract 2e-2)
GEN_PSNR_GUARD = 58.0  # measured 62.7 .. 69.2 dB (contract 45 dB)


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _psnr(a, b):
    mse = torch.mean((a.double() - b.double()) ** 2).item()
    return 99.0 if mse == 0 else 10 * math.log10(1.0 / mse)


# ---------------------------------------------------------------------------------------------------------- cfg2


def _draw_kernels(batch, seed, force_sinc=False):
    """kernel1 / kernel2 / sinc kernel per sample with the reference's distributions (dataset.py:81-141): host draws
    through the product's own host logic, float64 arithmetic by the numpy oracle, cast to fp32 as the dataset does."""
    import resr_b200
    from oracle import kernels as ok
    P = resr_b200.plan.DEGRADATION_MODEL_PARAMETERS
    random.seed(seed)
    np.random.seed(seed)
    out = [[], [], []]
    for _ in range(batch):
        for which in (1, 2):
            ks = random.choice(P["gaussian_kernel_range"])
            if np.random.uniform() < P[f"sinc_kernel_probability{which}"]:
                om = np.random.uniform(np.pi / 3, np.pi) if ks < 13 else np.random.uniform(np.pi / 5, np.pi)
                p = {"type": "sinc", "kernel_size": ks, "cutoff": om}
            else:
                p, _ = resr_b200.imgproc.draw_mixed_kernel_params(
                    P["gaussian_kernel_type"], P[f"gaussian_kernel_probability{which}"], ks, P[f"gaussian_sigma_range{which}"],
                    P[f"gaussian_sigma_range{which}"], [-math.pi, math.pi], P[f"generalized_kernel_beta_range{which}"],
                    P[f"plateau_kernel_beta_range{which}"])
            out[which - 1].append(ok.from_params(p, 21))
        if force_sinc or np.random.uniform() < P["sinc_kernel_probability3"]:
            ks = random.choice(P["gaussian_kernel_range"])
            out[2].append(ok.from_params({"type": "sinc", "kernel_size": ks, "cutoff": np.random.uniform(np.pi / 3, np.pi)}, 21))
        else:
            out[2].append(ok.from_params({"type": "delta", "kernel_size": 21}, 21))
    return [np.stack(v).astype(np.float32) for v in out]


def _jpeg_flips(q_native, parts):
    flips = unexplained = 0
    for name, qn in zip(("y", "cb", "cr"), q_native):
        qn = qn.cpu().numpy()
        mism = qn != parts[name + "_q"]
        flips += int(mism.sum())
        ratio = parts[name + "_ratio"]
        near_tie = np.abs(np.abs(ratio - np.floor(ratio)) - 0.5) < 2e-3
        unexplained += int((mism & ~near_tie).sum())
        assert np.abs(qn - parts[name + "_q"]).max() <= 1
    return flips, unexplained


def _stagewise(hr, k1, k2, sk, plan, rng):
    """Runs the oracle stage by stage; every CUDA stage is fed the oracle's input of that stage. Returns the report rows
    and the completed plan (Poisson draws attached, so the end-to-end runs replay identical numbers)."""
    import resr_b200
    from oracle import degrade as od
    ip = resr_b200.imgproc
    rows = []
    chain = []   # the oracle's own chain: (stage name, stage output)

    def cmp(name, y, ref, flips_ok=0.0):
        chain.append((name, ref))
        d = np.abs(y.cpu().numpy().astype(np.float64) - ref.astype(np.float64))
        frac = float((d > TOL).mean())
        rows.append((name, tuple(ref.shape), float(d.max()), frac))
        assert frac <= flips_ok, (name, float(d.max()), frac)

    def noise(x, p, name):
        if p["type"] == "poisson" and p.get("samples_color") is None:
            with_gray = bool(p["gray"].sum() > 0)
            rates = od.poisson_rates(x, with_gray)
            p["samples_color"] = rng.poisson(rates["rate"]).astype(np.float32)
            p["samples_gray"] = rng.poisson(rates["rate_g"]).astype(np.float32) if with_gray else None
            rc, rg = ip.poisson_rates(_t(x), with_gray)   # the integer side of the stage: counts -> vals -> rates
            assert np.array_equal(rc.cpu().numpy(), rates["rate"]), "Poisson rates (q * 2^ceil(log2 #unique)) must be bit-exact"
            if with_gray:
                assert np.array_equal(rg.cpu().numpy(), rates["rate_g"])
        if p["type"] == "gaussian":
            ref = od.gaussian_noise_apply(x, p["sigma"], p["gray"], p["noise_color"], p.get("noise_gray"))
        else:
            ref = od.poisson_noise_apply(x, p["scale"], p["gray"], p["samples_color"], p.get("samples_gray"))
        cmp(name, ip._noise(_t(x), p), ref)
        return ref

    def resize(x, r, name, use_scale):
        mode = ("area", "bilinear", "bicubic")[r["mode"]]
        if use_scale and r.get("scale") is not None:
            ref = od.resize(x, r["out_h"], r["out_w"], r["mode"], r["scale"], r["scale"])
            y = ip.interpolate(_t(x), scale_factor=r["scale"], mode=mode)
        else:
            ref = od.resize(x, r["out_h"], r["out_w"], r["mode"])
            y = ip.interpolate(_t(x), size=(r["out_h"], r["out_w"]), mode=mode)
        assert tuple(y.shape) == ref.shape
        cmp(name, y, ref)
        return ref

    def jpeg(x, q, name):
        xin = np.clip(x, 0, 1)
        ref, parts = od.jpeg(xin, q, return_parts=True)
        y, factor, qy, qcb, qcr = ip.DiffJPEG(False)(_t(x), _t(q), return_coefficients=True, clamp_input=True)
        assert np.array_equal(factor.cpu().numpy(), parts["factor"]), "quantisation factor must be bit-exact"
        flips, unexplained = _jpeg_flips((qy, qcb, qcr), parts)
        ncoef = sum(parts[c + "_q"].size for c in ("y", "cb", "cr"))
        rows.append((name + ".coef", (ncoef,), float(flips), float(unexplained)))
        assert unexplained == 0, f"{name}: {unexplained} coefficient flips away from a rounding tie"
        cmp(name, y, ref, flips_ok=0.0 if flips == 0 else 64.0 * flips / y.numel() * 4)
        return ref

    usm_ref, residual, _, _ = od.usm_sharp(hr, 0.5, 10, return_parts=True)
    y = ip.USMSharp(50, 0)(_t(hr), 0.5, 10)
    d = np.abs(y.cpu().numpy() - usm_ref)
    near = np.abs(np.abs(residual) * 255.0 - 10.0) < 1e-3   # mask decisions within fp32 noise of the threshold
    rows.append(("usm", hr.shape, float(d.max()), float((d > TOL).mean())))
    assert float((d > TOL).mean()) <= 1e-3 and np.median(d) <= 1e-6
    rows.append(("usm.near_threshold", (int(near.sum()),), 0.0, 0.0))
    chain.append(("usm", usm_ref))
    out = usm_ref
    if plan["blur1"]:
        ref = od.filter2d(out, k1)
        cmp("blur1", ip.filter2d_torch(_t(out), _t(k1)), ref)
        out = ref
    out = resize(out, plan["resize1"], "resize1", True)
    out = noise(out, plan["noise1"], "noise1")
    out = jpeg(out, plan["jpeg1_quality"], "jpeg1")
    if plan["blur2"]:
        ref = od.filter2d(out, k2)
        cmp("blur2", ip.filter2d_torch(_t(out), _t(k2)), ref)
        out = ref
    out = resize(out, plan["resize2"], "resize2", False)
    out = noise(out, plan["noise2"], "noise2")

    def sinc(x):
        ref = od.filter2d(x, sk)
        cmp("sinc", ip.filter2d_torch(_t(x), _t(sk)), ref)
        return ref

    if plan["final_order"] == 0:
        out = resize(out, plan["resize3"], "resize3", False)
        out = sinc(out)
        out = jpeg(out, plan["jpeg2_quality"], "jpeg2")
    else:
        out = jpeg(out, plan["jpeg2_quality"], "jpeg2")
        out = resize(out, plan["resize3"], "resize3", False)
        out = sinc(out)
    c = plan["crop"]
    lr_ref, hr_ref = od.round_and_crop(out, hr, c["hr_top"], c["hr_left"], c["image_size"], c["upscale"])
    return rows, (lr_ref, hr_ref), chain


@pytest.mark.parametrize("which", ["S0", "seed0", "seed1", "seed2", "seed3"])
def test_cfg2_degradation_per_stage_and_end_to_end(which):
    import resr_b200
    ip = resr_b200.imgproc
    B, H, W = 16, 256, 256
    if which == "S0":
        plan = resr_b200.plan.canonical_plan_s0(B, H, W, seed=0)
        seed = 100
    else:
        seed = int(which[4:])
        plan = resr_b200.plan.synth_plan(B, H, W, seed=seed)
    rng = np.random.default_rng(1000 + seed)
    hr = rng.random((B, 3, H, W), dtype=np.float32)
    hr[:, :, :40, :40] = np.round(hr[:, :, :40, :40] * 255) / 255  # some content on the u8 grid, as decoded images are
    k1, k2, sk = _draw_kernels(B, seed, force_sinc=(which == "S0"))
    supports = {int(np.abs(np.argwhere(k != 0) - 10).max()) * 2 + 1 for k in k1}
    rows, (lr_ref, hr_ref), chain = _stagewise(hr, k1, k2, sk, plan, rng)
    print(f"\ncfg2 {which}: blur1 supports {sorted(supports)}")
    for name, shape, a, b in rows:
        if name.endswith(".coef"):
            print(f"  {name:18s} {shape[0]:>9d} coefficients: {int(a)} flips, {int(b)} not at a tie")
        elif name.endswith("near_threshold"):
            print(f"  {name:18s} {shape[0]} mask decisions within 1e-3 of the threshold")
        else:
            print(f"  {name:18s} {str(shape):>20s} max {a:.2e}  frac>1e-5 {b:.1e}")
    # whole block, three sequencers (Python op sequence, one-call C ABI, CUDA graph) on the completed plan
    args = (_t(hr), _t(k1), _t(k2), _t(sk), plan)
    mine = []
    lr, hrc = ip.degrade_batch(*args, stages=mine)
    lr_n, hrc_n = ip.degrade_batch_native(*args)
    assert torch.equal(lr, lr_n) and torch.equal(hrc, hrc_n)
    pipe = ip.DegradePipeline(*args)
    lr_g, hrc_g = pipe()
    assert torch.equal(lr, lr_g) and torch.equal(hrc, hrc_g)
    assert np.array_equal(hrc.cpu().numpy(), hr_ref)
    lv = np.rint(lr.cpu().numpy() * 255)
    assert np.abs(lr.cpu().numpy() - lv / 255).max() < 1e-7, "lr must sit on the u8 grid"
    diff = np.abs(lv - np.rint(lr_ref * 255))
    nbad = int((diff > 0).sum())
    print(f"  end to end: {nbad}/{diff.size} u8 values differ (max {int(diff.max())} levels)")
    # Free-running chain vs the oracle's chain: per-stage differences of ~1e-6 are harmless until they tip a DECISION
    # (USM mask threshold, the u8 quantisation inside the Poisson stage, JPEG coefficient rounding, final u8 rounding); one flipped JPEG coefficient moves its 8x8
    # block by a few levels and a later upsampling spreads it. So: the first stage at which the two chains part by more
    # than 1e-5 must be a decision stage, and the damage must stay local.
    assert [n for n, _ in mine] == [n for n, _ in chain]
    first = None
    for (name, a), (_, b) in zip(mine, chain):
        d = np.abs(a.cpu().numpy() - b)
        if (d > TOL).any():
            first = (name, int((d > TOL).sum()), float(d.max()))
            break
    print(f"  chains part at: {first}")
    if first is None:
        assert nbad <= 1e-4 * diff.size and diff.max() <= 1   # only final-rounding ties can differ
    else:
        decision = {"usm", "jpeg1", "jpeg2"} | {n for n in ("noise1", "noise2") if plan[n]["type"] == "poisson"}
        assert first[0] in decision, first
        assert nbad <= 1e-2 * diff.size and diff.max() <= 24


# ---------------------------------------------------------------------------------------------------------- cfg3


def test_cfg3_batch_sampled_against_oracle():
    import resr_b200
    from oracle import generator as og
    sd = og.random_state_dict(0)
    g = resr_b200.model.Generator(3, 3, 4)
    g.load_state_dict(sd)
    g = g.cuda().eval()
    gen = torch.Generator().manual_seed(64)
    x = torch.rand(64, 3, 128, 128, generator=gen)
    with torch.no_grad():
        y = g(x.cuda())
    worst, worst_ps = 0.0, 99.0
    for i in (0, 21, 42, 63):
        ref = og.generator_forward(x[i:i + 1], sd)
        yi = y[i:i + 1].cpu()
        err, ps = (yi - ref).abs().max().item(), _psnr(yi, ref)
        print(f"cfg3 image {i}: max-abs {err:.3e}  psnr {ps:.2f} dB")
        worst, worst_ps = max(worst, err), min(worst_ps, ps)
    assert worst <= GEN_GUARD and worst_ps >= GEN_PSNR_GUARD


# ---------------------------------------------------------------------------------------------------------- cfg4


def test_cfg4_train_step_at_size():
    """Both training recipes (fp16 activations + channels-first weight-gradient copies; bf16 activations + the NHWC MN-major
    weight-gradient kernel) against ONE run of the oracle's fp32 autograd on the full batch."""
    import resr_b200
    from oracle import generator as og
    sd = og.random_state_dict(1)
    gen = torch.Generator().manual_seed(4)
    n = 16
    lr = torch.rand(n, 3, 64, 64, generator=gen)
    hr = torch.rand(n, 3, 256, 256, generator=gen)
    # the oracle's fp32 autograd on a quarter of the batch at a time would change the mean: run the full batch
    ref_loss, ref_grads, ref_sr = og.l1_loss_and_grads(lr, hr, sd)
    ref_flat = torch.cat([ref_grads[k].reshape(-1) for k in sd])
    for precision in ("fp16", "bf16"):
        g = resr_b200.model.Generator(3, 3, 4)
        g.load_state_dict(sd)
        g = g.cuda().train()
        g.set_precision(precision)
        ts = resr_b200.autograd.TrainStep(g, n, 64, 64)
        loss, sr, flat = ts.step(lr.cuda(), hr.cuda(), scatter=False)
        torch.cuda.synchronize()
        assert ts.is_graph
        got = flat.cpu()
        rel = abs(loss.item() - ref_loss.item()) / ref_loss.item()
        cos = float(torch.dot(got.double(), ref_flat.double()) / (got.double().norm() * ref_flat.double().norm()))
        rl2 = float((got - ref_flat).double().norm() / ref_flat.double().norm())
        sr_err = (sr.cpu() - ref_sr).abs().max().item()
        print(f"cfg4 16x3x64x64 [{precision}]: loss rel err {rel:.2e}; grad cosine {cos:.6f}; rel-L2 {rl2:.3%}; sr max-abs {sr_err:.2e}")
        assert torch.isfinite(got).all()
        assert rel <= 1e-3 and cos >= 0.9995 and rl2 <= 0.03 and sr_err <= 2e-2
        del ts, g


# ---------------------------------------------------------------------------------------------------------- cfg5


def test_cfg5_tiled_against_oracle_windows():
    import resr_b200
    from oracle import generator as og
    sd = og.random_state_dict(5)
    g = resr_b200.model.Generator(3, 3, 4)
    g.load_state_dict(sd)
    g = g.cuda().eval()
    gen = torch.Generator().manual_seed(512)
    x = torch.rand(1, 3, 512, 512, generator=gen)
    out, mine = resr_b200.model.infer_tiled(g, x.cuda(), tile_h=256, tile_w=256, halo=16)
    assert len(mine) == 4 and out.shape == (1, 3, 2048, 2048)
    out = out.cpu()
    m = 24  # the oracle's own window margin: its truncation error is ~1e-6, below the comparison noise
    # (a) 96x96 window straddling both tile seams (y = x = 256); (b) the top-left image corner (zero padding applies)
    for name, (y0, x0) in (("seam", (208, 208)), ("corner", (0, 0))):
        wy0, wx0 = max(y0 - m, 0), max(x0 - m, 0)
        wy1, wx1 = min(y0 + 96 + m, 512), min(x0 + 96 + m, 512)
        ref = og.generator_forward(x[:, :, wy0:wy1, wx0:wx1], sd)
        ref = ref[:, :, 4 * (y0 - wy0):4 * (y0 - wy0 + 96), 4 * (x0 - wx0):4 * (x0 - wx0 + 96)]
        got = out[:, :, 4 * y0:4 * (y0 + 96), 4 * x0:4 * (x0 + 96)]
        err, ps = (got - ref).abs().max().item(), _psnr(got, ref)
        print(f"cfg5 {name} window: max-abs {err:.3e}  psnr {ps:.2f} dB")
        assert err <= 2e-2 and ps >= 45.0            # the contract
        assert err <= 1.2e-2 and ps >= GEN_PSNR_GUARD - 6  # regression guard (halo-16 truncation adds to the fp16 noise)
