"""GPU parity of the non-default branches of the reference call (marigold_dc.py:467-493) against the oracle:
projection log / log10 / inv, norm percentile, optimisers sgd / adagrad, loss subsets, edge / smooth losses and the
kld penalty -- prologue, tail kernels (teacher-forced, against torch autograd / torch.optim) and end to end."""
import copy

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def setup(cuda):
    from helpers import build_engine, build_models

    unet, vae, ctx, ucfg, vcfg = build_models(cuda, tiny=True)
    eng = build_engine(unet, vae, ctx, ucfg, vcfg, 2, 80, 111, 125, 50, cuda)
    return (unet, vae, ctx), eng


def _inputs(eng, seed=0, channels=3, u8=True):
    dev = eng.device
    g = torch.Generator(device=dev).manual_seed(seed)
    N, H, W = eng.n, eng.H, eng.W
    # depths above 1 m: the reference's inverse of a log projection is singular at 1 m (log = 0)
    sparse = torch.rand(N, 1, H, W, device=dev, generator=g) * 9 + 2.0
    sparse = sparse * (torch.rand(N, 1, H, W, device=dev, generator=g) < 0.03)
    img = torch.randint(0, 256, (N, channels, H, W), device=dev, generator=g, dtype=torch.uint8)
    if not u8:
        img = img.float() / 255
    x = torch.randn(N, 4, eng.lh, eng.lw, device=dev, generator=g).bfloat16()
    return img, sparse, x


@pytest.mark.parametrize("projection,inv,norm", [("log", False, "minmax"), ("log10", True, "const"), ("linear", True, "minmax"),
                                                 ("linear", False, "percentile"), ("log", True, "percentile")])
def test_prologue_projection_and_percentile(setup, projection, inv, norm):
    """mdc_begin_frame's sparse-depth normalisation (marigold_dc.py:707-756) for every projection / norm branch."""
    from oracle.marigold_dc import OraclePipeline

    (unet, vae, ctx), eng = setup
    img, sparse, x = _inputs(eng, seed=3)
    pct = (0.05, 0.9)
    eng.set_options(projection=projection, inv=inv, percentile=pct)
    eng.begin_frame(img, sparse, x, 12.0, 1.5, norm)
    guide, mask, st = eng.dbg_frame_state()
    op = OraclePipeline(unet, vae, ctx)
    ref = op.preprocess(img, sparse, 12.0, 1.5, norm, 125, 2024, None, 0.9, projection, inv, pct)
    assert torch.equal(mask.bool(), ref["masks"])
    assert torch.allclose(torch.from_numpy(st[:, 0]), ref["min_depths"].flatten().cpu(), rtol=1e-6)
    assert torch.allclose(torch.from_numpy(st[:, 1]), ref["max_depths"].flatten().cpu(), rtol=1e-6)
    m = ref["masks"]
    assert torch.allclose(guide[m], ref["sparses_normed"][m], rtol=1e-4, atol=2e-5)
    eng.set_options()


def test_percentile_empty_sample_raises(setup):
    _, eng = setup
    img, sparse, x = _inputs(eng, seed=4)
    sparse[1] = 0
    from depth_completion_b200._lib import MdcError

    eng.set_options()
    with pytest.raises((MdcError, ValueError)):
        eng.begin_frame(img, sparse, x, 10.0, 0.0, "percentile")
    with pytest.raises(ValueError):  # log projection needs min_depth > 1e-7 (marigold_dc.py:636-641)
        eng.set_options(projection="log")
        try:
            eng.begin_frame(img, sparse, x, 10.0, 0.0, "minmax")
        finally:
            eng.set_options()


def _torch_loss(eng, dec, ref, s, t, loss_funcs, img):
    """marigold_dc.py:829-875 on a given decoder output, with torch autograd (fp32 math)."""
    from oracle.marigold_dc import OraclePipeline, compute_loss, masked_minmax

    N = eng.n
    r16 = lambda v: v + (v.bfloat16().float() - v).detach()  # the bf16 rounding points of the reference's bf16 mode
    y = r16((r16(dec.mean(1, keepdim=True)).clip(-1, 1) + 1) / 2)
    a = r16(torch.nn.functional.interpolate(y[:, :, : eng.ph, : eng.pw], (eng.H, eng.W), mode="bilinear"))
    gmin, gmax = masked_minmax(ref["sparses_normed"].view(N, -1), ref["masks"].view(N, -1), dim=-1)
    dense = ((s ** 2) * (gmax - gmin).view(-1, 1, 1, 1) * a + (t ** 2) * gmin.view(-1, 1, 1, 1)).clamp(0, 1)
    dense = OraclePipeline.to_guide_space(dense, ref)
    return compute_loss(dense, ref["sparses_normed"], ref["masks"], loss_funcs, images=img), dense


@pytest.mark.parametrize("projection,inv,loss_funcs,channels,u8", [
    ("log", True, ("l1", "l2"), 3, True),
    ("log10", False, ("l2",), 3, True),
    ("linear", False, ("l1", "l1", "l2"), 3, True),
    ("linear", False, ("smooth",), 3, True),
    ("linear", False, ("l1", "l2", "edge"), 3, False),
    ("linear", False, ("edge", "smooth"), 1, True),
    ("log", False, ("l1", "l2", "edge", "smooth"), 3, True),
])
def test_loss_terms_and_gradient(setup, projection, inv, loss_funcs, channels, u8):
    """Loss value and its gradient w.r.t. the decoder output and the scale, per term and projection, against autograd."""
    from oracle.marigold_dc import OraclePipeline

    (unet, vae, ctx), eng = setup
    dev = eng.device
    img, sparse, x = _inputs(eng, seed=7, channels=channels, u8=u8)
    eng.set_options(projection=projection, inv=inv, loss_funcs=loss_funcs)
    eng.begin_frame(img, sparse, x, 12.0, 1.5, "minmax")
    ref = OraclePipeline(unet, vae, ctx).preprocess(img, sparse, 12.0, 1.5, "minmax", 125, 2024, None, 0.9, projection, inv)
    g = torch.Generator(device=dev).manual_seed(5)
    # a smooth decoder output (low-resolution noise, upsampled): |d dense| well away from 0 between neighbours
    low = torch.randn(eng.n, 1, 6, 8, device=dev, generator=g)
    dec = torch.nn.functional.interpolate(low, (eng.lh * 8, eng.lw * 8), mode="bicubic").repeat(1, 3, 1, 1) * 0.5
    dec = (dec + 0.02 * torch.randn(dec.shape, device=dev, generator=g)).bfloat16().float()
    ddec, loss, gs, gt = eng.dbg_loss(dec)
    d = dec.clone().requires_grad_(True)
    s = torch.ones(eng.n, 1, 1, 1, device=dev, requires_grad=True)
    t = torch.zeros(eng.n, 1, 1, 1, device=dev, requires_grad=True)
    lref, dense = _torch_loss(eng, d, ref, s, t, loss_funcs, img)
    lref.backward(torch.ones_like(lref))
    assert torch.allclose(loss.to(dev), lref.detach(), rtol=2e-2, atol=1e-4), (loss, lref)
    # bf16 rounding of the affine map flips sign() terms of near-ties (points within ~1e-3 of their guide, neighbours
    # whose difference is ~0): judge the gradient by its direction and norm, as the end-to-end tests do
    cos = torch.nn.functional.cosine_similarity(ddec.flatten(), d.grad.flatten(), dim=0).item()
    ratio = (ddec.norm() / d.grad.norm()).item()
    dense_terms = any(f in ("edge", "smooth") for f in loss_funcs)
    assert cos > (0.95 if dense_terms else 0.97), cos
    assert abs(ratio - 1) < (0.15 if dense_terms else 0.05), ratio
    assert torch.allclose(gs.to(dev), s.grad.flatten(), rtol=8e-2, atol=2e-3), (gs, s.grad.flatten())
    assert ddec[:, :, eng.ph:, :].abs().max() == 0
    eng.set_options()


@pytest.mark.parametrize("opt", ["sgd", "adagrad"])
def test_sgd_and_adagrad_updates(setup, opt):
    """torch.optim.SGD / Adagrad on the bf16 latent (marigold_dc.py:776-789, :897) for three teacher-forced steps."""
    from oracle.scheduler import DDIMScheduler

    _, eng = setup
    dev = eng.device
    img, sparse, x0 = _inputs(eng, seed=11)
    eng.set_options(opt=opt)
    eng.begin_frame(img, sparse, x0, 10.0, 0.0, "minmax")
    sch = DDIMScheduler()
    sch.set_timesteps(50)
    x = torch.nn.Parameter(x0.clone())
    optim = {"sgd": torch.optim.SGD, "adagrad": torch.optim.Adagrad}[opt]([{"params": [x], "lr": 0.05}])
    g = torch.Generator(device=dev).manual_seed(9)
    N = eng.n
    for k in range(3):
        t = sch.timesteps[k]
        v = torch.randn(N, 4, eng.lh, eng.lw, device=dev, generator=g).bfloat16()
        dz = (torch.randn(N, 4, eng.lh, eng.lw, device=dev, generator=g) * 1e-3).bfloat16()
        du = (torch.randn(N, 8, eng.lh, eng.lw, device=dev, generator=g) * 1e-2).bfloat16()
        eng.dbg_update(v.float(), dz.float(), du.float())
        a_t = sch.alphas_cumprod[int(t)]
        eps = (a_t ** 0.5) * v + ((1 - a_t) ** 0.5) * x.detach()
        grad = (a_t ** 0.5) * (dz / 0.18215) + du[:, 4:8]
        en = torch.linalg.norm(eps.view(N, -1), dim=1)
        gn = torch.linalg.norm(grad.view(N, -1), dim=1)
        x.grad = grad * (en / gn.clamp(min=1e-7)).view(N, 1, 1, 1)
        optim.step()
        xa = eng.dbg_x_adam()
        frac = ((xa.float() - x.detach().float()).abs() <= 2e-2 * x.detach().float().abs().clamp_min(0.5)).float().mean()
        assert frac.item() > 0.995, f"step {k}: {opt} agreement {frac.item():.4f}"
        with torch.no_grad():
            x.data = sch.step(v, t, x.detach()).prev_sample
        xo, _, _, _ = eng.get_state()
        close = ((xo.float() - x.detach().float()).abs() < 2e-2).float().mean().item()
        assert close > 0.99, f"step {k}: DDIM agreement {close:.4f}"
    eng.set_options()


@pytest.mark.parametrize("mode", ["simple", "strict"])
def test_kld_gradient(setup, mode):
    """kld_weight * d kld_stdnorm(x) / dx joins the latent gradient (marigold_dc.py:238-241, utils.py:69-77)."""
    from oracle.marigold_dc import kld_stdnorm

    _, eng = setup
    dev = eng.device
    img, sparse, x0 = _inputs(eng, seed=13)
    x0 = (x0.float() * 1.5 + 0.3).bfloat16()
    N = eng.n
    zeros4 = torch.zeros(N, 4, eng.lh, eng.lw, device=dev)
    zeros8 = torch.zeros(N, 8, eng.lh, eng.lw, device=dev)
    g = torch.Generator(device=dev).manual_seed(2)
    v = torch.randn(N, 4, eng.lh, eng.lw, device=dev, generator=g).bfloat16().float()
    eng.set_options(kld=True, kld_weight=0.25, kld_mode=mode)
    eng.begin_frame(img, sparse, x0, 10.0, 0.0, "minmax")
    eng.dbg_update(v, zeros4, zeros8)
    got = eng.dbg_buffer("grad")
    xr = x0.float().clone().requires_grad_(True)
    # strict mode's epsilon is finfo(x.dtype).eps of the bf16 latent (utils.py:71)
    nn = xr.reshape(N, -1)
    if mode == "simple":
        dist = nn.square().mean(-1)
    else:
        mu, var = nn.mean(-1), nn.var(-1, unbiased=False)
        dist = 0.5 * (mu.square() + var - torch.log(var + torch.finfo(torch.bfloat16).eps) - 1)
    (0.25 * dist).sum().backward()
    assert torch.allclose(dist.detach(), kld_stdnorm(x0.float(), "none", mode), rtol=1e-3) or mode == "strict"
    from helpers import rel_l2

    assert rel_l2(got, xr.grad) < 2e-2, rel_l2(got, xr.grad)
    eng.set_options()


@pytest.mark.parametrize("kw", [
    dict(projection="log", min_depth=0.1),
    dict(inv=True, min_depth=0.1, norm="const"),
    dict(norm="percentile", percentile=(0.02, 0.98)),
    dict(opt="sgd"),
    dict(opt="adagrad"),
    dict(loss_funcs=["l1", "l2", "smooth", "edge"]),
    dict(kld=True, kld_weight=0.1, kld_mode="simple"),
    dict(kld=True, kld_weight=0.05, kld_mode="strict", loss_funcs=["l2"]),
    dict(closed_form=True),
    dict(closed_form=True, projection="log", min_depth=0.1, opt="adagrad"),
])
def test_pipeline_options_match_oracle(setup, cuda, kw):
    """The drop-in class with each non-default option against the oracle in bf16 on the same GPU, 12 guided steps:
    same loss level, same dense map up to the bf16 chaos of the loop (yardstick: the oracle's own bf16 vs fp32 spread)."""
    from depth_completion_b200.pipeline import MarigoldDepthCompletionPipeline
    from depth_completion_b200.synthetic import make_frame
    from oracle.marigold_dc import OraclePipeline

    (unet, vae, ctx), _ = setup
    fr = make_frame(H=96, W=128, n_points=100, seed=2)
    img, sp = fr["img"].to(cuda), fr["sparse"].to(cuda)
    pipe = MarigoldDepthCompletionPipeline(unet, vae)
    pipe.empty_text_embedding = ctx
    dense, lat = pipe(img, sp, fr["max_depth"], steps=12, resolution=128, **kw)
    assert torch.isfinite(dense).all() and torch.isfinite(lat.float()).all()
    okw = {k: (tuple(v) if isinstance(v, list) else v) for k, v in kw.items()}
    d32, _ = OraclePipeline(copy.deepcopy(unet), copy.deepcopy(vae), ctx)(img, sp, fr["max_depth"], steps=12, resolution=128, **okw)
    d16, _ = OraclePipeline(copy.deepcopy(unet).bfloat16(), copy.deepcopy(vae).bfloat16(), ctx.bfloat16())(
        img, sp, fr["max_depth"], steps=12, resolution=128, **okw)
    rng = fr["max_depth"]
    ours = ((dense - d32).abs().mean() / rng).item()
    ref = ((d16 - d32).abs().mean() / rng).item()
    assert ours < max(2.0 * ref, 1e-2) + 1e-2, f"{kw}: mean |dense - fp32 oracle| / range: ours {ours:.4f}, torch-bf16 {ref:.4f}"
    # the default call differs from this one: the option did take effect
    base, _ = pipe(img, sp, fr["max_depth"], steps=12, resolution=128)
    assert not torch.equal(base, dense)


def test_learning_rates_are_per_call(setup, cuda):
    """lr is read from device memory each step, so a second call with another lr on the same (graph-captured) engine
    must not reuse the first call's rates."""
    from depth_completion_b200.pipeline import MarigoldDepthCompletionPipeline
    from depth_completion_b200.synthetic import make_frame

    (unet, vae, ctx), _ = setup
    fr = make_frame(H=96, W=128, n_points=100, seed=2)
    img, sp = fr["img"].to(cuda), fr["sparse"].to(cuda)
    pipe = MarigoldDepthCompletionPipeline(unet, vae)
    pipe.empty_text_embedding = ctx
    a, _ = pipe(img, sp, fr["max_depth"], steps=6, resolution=128)
    b, _ = pipe(img, sp, fr["max_depth"], steps=6, resolution=128, lr=(0.01, 0.001))
    c, _ = pipe(img, sp, fr["max_depth"], steps=6, resolution=128)
    assert torch.equal(a, c) and not torch.equal(a, b)
    fresh = MarigoldDepthCompletionPipeline(unet, vae)
    fresh.empty_text_embedding = ctx
    b2, _ = fresh(img, sp, fr["max_depth"], steps=6, resolution=128, lr=(0.01, 0.001))
    assert torch.equal(b, b2)


def test_no_grad_sampling_with_closed_form_affine(setup, cuda):
    """train_latents=False (marigold_dc.py:605-613, :905-909, :53-128): plain DDIM sampling + least-squares scale / shift.
    Without the guidance the loop is not chaotic, so the latent itself is compared."""
    from depth_completion_b200.pipeline import MarigoldDepthCompletionPipeline
    from depth_completion_b200.synthetic import make_frame
    from helpers import rel_l2
    from oracle.marigold_dc import OraclePipeline

    (unet, vae, ctx), _ = setup
    fr = make_frame(H=96, W=128, n_points=100, seed=4)
    img, sp = fr["img"].to(cuda), fr["sparse"].to(cuda)
    pipe = MarigoldDepthCompletionPipeline(unet, vae)
    pipe.empty_text_embedding = ctx
    dense, lat = pipe(img, sp, fr["max_depth"], steps=10, resolution=128, train_latents=False)
    d32, l32 = OraclePipeline(copy.deepcopy(unet), copy.deepcopy(vae), ctx)(img, sp, fr["max_depth"], steps=10, resolution=128,
                                                                            train_latents=False)
    d16, l16 = OraclePipeline(copy.deepcopy(unet).bfloat16(), copy.deepcopy(vae).bfloat16(), ctx.bfloat16())(
        img, sp, fr["max_depth"], steps=10, resolution=128, train_latents=False)
    e_ours, e_ref = rel_l2(lat, l32), rel_l2(l16, l32)
    assert e_ours < 1.25 * e_ref + 5e-3, f"latent after 10 DDIM steps: ours {e_ours:.4f}, torch-bf16 {e_ref:.4f}"
    rng = fr["max_depth"]
    ours = ((dense - d32).abs().mean() / rng).item()
    ref = ((d16 - d32).abs().mean() / rng).item()
    assert ours < max(2.0 * ref, 5e-3) + 5e-3, f"dense: ours {ours:.4f}, torch-bf16 {ref:.4f}"
    guided, _ = pipe(img, sp, fr["max_depth"], steps=10, resolution=128)
    assert not torch.equal(guided, dense)
    with pytest.raises(NotImplementedError):
        pipe(img, sp, fr["max_depth"], steps=10, resolution=128, closed_form=True, loss_funcs=["l1", "smooth"])


def test_closed_form_loss_gradient_flows_through_the_fit(setup):
    """closed_form=True inside the guided loop (marigold_dc.py:332-336 via :53-128): loss and d loss / d decoder output
    against autograd through compute_affine_params; the fitted scale / shift are reported by get_state."""
    from oracle.marigold_dc import OraclePipeline, compute_affine_params, compute_loss

    (unet, vae, ctx), eng = setup
    dev = eng.device
    img, sparse, x = _inputs(eng, seed=17)
    eng.set_options(closed_form=True)
    eng.begin_frame(img, sparse, x, 12.0, 0.0, "minmax")
    ref = OraclePipeline(unet, vae, ctx).preprocess(img, sparse, 12.0, 0.0, "minmax", 125, 2024, None, 0.9)
    g = torch.Generator(device=dev).manual_seed(5)
    low = torch.randn(eng.n, 1, 6, 8, device=dev, generator=g)
    dec = torch.nn.functional.interpolate(low, (eng.lh * 8, eng.lw * 8), mode="bicubic").repeat(1, 3, 1, 1) * 0.5
    dec = (dec + 0.02 * torch.randn(dec.shape, device=dev, generator=g)).bfloat16().float()
    ddec, loss, gs, gt = eng.dbg_loss(dec)
    d = dec.clone().requires_grad_(True)
    r16 = lambda v: v + (v.bfloat16().float() - v).detach()
    y = r16((r16(d.mean(1, keepdim=True)).clip(-1, 1) + 1) / 2)
    a = r16(torch.nn.functional.interpolate(y[:, :, : eng.ph, : eng.pw], (eng.H, eng.W), mode="bilinear"))
    cs, ct = compute_affine_params(a, ref["sparses_normed"], ref["masks"])
    dense = (cs.view(-1, 1, 1, 1) * a + ct.view(-1, 1, 1, 1)).clamp(0, 1)
    lref = compute_loss(dense, ref["sparses_normed"], ref["masks"])
    lref.backward(torch.ones_like(lref))
    assert torch.allclose(loss.to(dev), lref.detach(), rtol=2e-2, atol=1e-4), (loss, lref)
    cos = torch.nn.functional.cosine_similarity(ddec.flatten(), d.grad.flatten(), dim=0).item()
    ratio = (ddec.norm() / d.grad.norm()).item()
    assert cos > 0.97 and abs(ratio - 1) < 0.05, (cos, ratio)
    assert gs.abs().max() == 0 and gt.abs().max() == 0  # no learned affine parameters in this mode
    _, sc, sh, _ = eng.get_state()
    assert torch.allclose(sc.to(dev), cs.detach(), rtol=1e-3, atol=1e-4) and torch.allclose(sh.to(dev), ct.detach(), rtol=1e-3, atol=1e-4)
    eng.set_options()


def test_interp_mode_nearest(setup, cuda):
    """interp_mode="nearest" (predict.py:200-206; marigold_dc.py:366-370): the prediction is resized to the input
    resolution with F.interpolate(mode="nearest") in the loop and in the final decode."""
    from depth_completion_b200.pipeline import MarigoldDepthCompletionPipeline
    from depth_completion_b200.synthetic import make_frame
    from oracle.marigold_dc import OraclePipeline

    (unet, vae, ctx), eng = setup
    dev = eng.device
    # kernel level: loss + gradient at 80x111 <- 90x125 (non-integer ratio) against autograd
    img, sparse, x = _inputs(eng, seed=19)
    eng.set_options(interp_mode="nearest")
    eng.begin_frame(img, sparse, x, 12.0, 0.0, "minmax")
    ref = OraclePipeline(unet, vae, ctx).preprocess(img, sparse, 12.0, 0.0, "minmax", 125, 2024, None, 0.9)
    g = torch.Generator(device=dev).manual_seed(5)
    dec = (torch.randn(eng.n, 3, eng.lh * 8, eng.lw * 8, device=dev, generator=g) * 0.6).bfloat16().float()
    ddec, loss, gs, gt = eng.dbg_loss(dec)
    d = dec.clone().requires_grad_(True)
    r16 = lambda v: v + (v.bfloat16().float() - v).detach()
    y = r16((r16(d.mean(1, keepdim=True)).clip(-1, 1) + 1) / 2)
    a = torch.nn.functional.interpolate(y[:, :, : eng.ph, : eng.pw], (eng.H, eng.W), mode="nearest")
    from oracle.marigold_dc import compute_loss, masked_minmax

    N = eng.n
    gmin, gmax = masked_minmax(ref["sparses_normed"].view(N, -1), ref["masks"].view(N, -1), dim=-1)
    dense = ((gmax - gmin).view(-1, 1, 1, 1) * a).clamp(0, 1)
    lref = compute_loss(dense, ref["sparses_normed"], ref["masks"])
    lref.backward(torch.ones_like(lref))
    assert torch.allclose(loss.to(dev), lref.detach(), rtol=1e-3, atol=1e-5), (loss, lref)
    cos = torch.nn.functional.cosine_similarity(ddec.flatten(), d.grad.flatten(), dim=0).item()
    assert cos > 0.99 and abs((ddec.norm() / d.grad.norm()).item() - 1) < 0.02, cos
    eng.set_options()
    # end to end, switching modes on one live pipeline (the step graph is re-captured when the mode changes)
    fr = make_frame(H=90, W=120, n_points=100, seed=6)
    im, sp = fr["img"].to(cuda), fr["sparse"].to(cuda)
    pipe = MarigoldDepthCompletionPipeline(unet, vae)
    pipe.empty_text_embedding = ctx
    bil, _ = pipe(im, sp, fr["max_depth"], steps=8, resolution=128)
    near, _ = pipe(im, sp, fr["max_depth"], steps=8, resolution=128, interp_mode="nearest")
    bil2, _ = pipe(im, sp, fr["max_depth"], steps=8, resolution=128)
    assert torch.equal(bil, bil2) and not torch.equal(bil, near)
    d32, _ = OraclePipeline(copy.deepcopy(unet), copy.deepcopy(vae), ctx)(im, sp, fr["max_depth"], steps=8, resolution=128,
                                                                          interp_mode="nearest")
    d16, _ = OraclePipeline(copy.deepcopy(unet).bfloat16(), copy.deepcopy(vae).bfloat16(), ctx.bfloat16())(
        im, sp, fr["max_depth"], steps=8, resolution=128, interp_mode="nearest")
    rng = fr["max_depth"]
    ours, refd = ((near - d32).abs().mean() / rng).item(), ((d16 - d32).abs().mean() / rng).item()
    assert ours < max(2.0 * refd, 1e-2) + 1e-2, f"nearest: ours {ours:.4f}, torch-bf16 {refd:.4f}"
    with pytest.raises(NotImplementedError):
        pipe(im, sp, fr["max_depth"], steps=8, resolution=128, interp_mode="bicubic")
