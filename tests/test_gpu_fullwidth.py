"""Teacher-forced guided steps at FULL WIDTH (the real SD2 UNet: 320/640/1280/1280 channels, 866 M parameters, and the
real VAE decoder: 128/256/512/512) on the BASELINE.json frame geometries, through the drop-in pipeline class and the C ABI.

For every configuration the CUDA engine executes ONE guided step (marigold_dc.py:801-904) from exactly the state the
fp32 oracle is in, and its UNet output v, loss and total latent gradient are compared with the fp32 oracle, using the
oracle's own bf16 run (torch kernels: cuDNN / cuBLAS / SDPA -- the reference's bf16 mode) from the same state as the
yardstick.  Config (b) is additionally teacher-forced at steps 10, 25 and 49 from the oracle's latent and Adam state.
These are the kernels the benchmark runs: two-pass GroupNorm on 113-226 MB tensors, flash attention over 4800-6912
tokens, split-K convolutions on 9x12 maps, the fused upsample convolutions, the head_dim-512 VAE attention.
"""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu

CONFIGS = {
    "b_nyu_res768": dict(N=1, H=480, W=640, res=768, kind="nyu", max_depth=10.0),      # latent 72x96
    "b_nyu_res640": dict(N=1, H=480, W=640, res=640, kind="nyu", max_depth=10.0),      # latent 60x80 (odd down path 15 -> 8)
    "c_kitti_res1216": dict(N=1, H=352, W=1216, res=1216, kind="kitti", max_depth=80.0),  # latent 44x152 (44 -> 22 -> 11 -> 6)
    "b_batch2": dict(N=2, H=480, W=640, res=768, kind="nyu", max_depth=10.0),
}


@pytest.fixture(scope="module")
def full(cuda):
    """fp32 oracle modules of the full architecture with bf16-representable weights, their bf16 copies, the pipeline."""
    from depth_completion_b200.pipeline import MarigoldDepthCompletionPipeline
    from oracle.marigold_dc import OraclePipeline, make_empty_text_embedding
    from oracle.sd2_modules import AutoencoderKL, UNet2DConditionModel, UNetConfig, VAEConfig

    torch.manual_seed(1234)
    with torch.device(cuda):
        unet, vae = UNet2DConditionModel(UNetConfig()), AutoencoderKL(VAEConfig())
    with torch.no_grad():
        for p in list(unet.parameters()) + list(vae.parameters()):
            p.copy_(p.bfloat16().float())
    unet, vae = unet.requires_grad_(False), vae.requires_grad_(False)
    ctx = make_empty_text_embedding(1024, device=cuda).bfloat16().float()
    o32 = OraclePipeline(unet, vae, ctx)
    o16 = OraclePipeline(copy.deepcopy(unet).bfloat16(), copy.deepcopy(vae).bfloat16(), ctx.bfloat16())
    pipe = MarigoldDepthCompletionPipeline(unet, vae)
    pipe.empty_text_embedding = ctx
    yield dict(o32=o32, o16=o16, pipe=pipe)
    pipe._invalidate()


def _frames(cfg, dev):
    from depth_completion_b200.synthetic import make_batch

    fr = make_batch(cfg["N"], H=cfg["H"], W=cfg["W"], kind=cfg["kind"], n_points=500, max_depth=cfg["max_depth"],
                    min_field=1.0 if cfg["kind"] == "kitti" else 0.5, seed=3)
    return fr["img"].to(dev), fr["sparse"].to(dev)


def _oracle_step_from(o, st, t, state, lr=(0.05, 0.005)):
    """One guided step of oracle pipeline `o` from a given (latent, Adam, scale / shift) state; returns its trace record."""
    dt, dev = o.dtype, o.device
    x = torch.nn.Parameter(state["x_in"].to(dt).clone())
    scales = torch.nn.Parameter(state["opt_in"]["scales_in"].float().clone())
    shifts = torch.nn.Parameter(state["opt_in"]["shifts_in"].float().clone())
    opt = torch.optim.Adam([{"params": [x], "lr": lr[0]}, {"params": [scales, shifts], "lr": lr[1]}])
    if state["idx"] > 0:
        for p, name in ((x, "x"), (scales, "scales"), (shifts, "shifts")):
            opt.state[p] = dict(step=torch.tensor(float(state["idx"])), exp_avg=state["opt_in"][name]["exp_avg"].to(p.dtype).clone(),
                                exp_avg_sq=state["opt_in"][name]["exp_avg_sq"].to(p.dtype).clone())
    rec = []
    o.scheduler.set_timesteps(50, device=dev)
    o.guided_step(st, t, x, scales, shifts, opt, rec.append, state["idx"], dict(loss_funcs=("l1", "l2"), kld=False))
    return rec[0]


def _engine_step_from(pipe, imgs, sparses, cfg, state):
    """The CUDA engine, one guided step from the same state; returns (v, grad, loss, engine)."""
    pipe(imgs, sparses, cfg["max_depth"], steps=50, resolution=cfg["res"], _begin_only=True)
    eng = list(pipe._engines.values())[-1]
    oi = state["opt_in"]
    a6 = torch.stack([oi["scales_in"].flatten(), oi["shifts_in"].flatten(), oi["scales"]["exp_avg"].flatten(),
                      oi["scales"]["exp_avg_sq"].flatten(), oi["shifts"]["exp_avg"].flatten(),
                      oi["shifts"]["exp_avg_sq"].flatten()]).float().cpu().numpy()
    eng.dbg_set_state(state["idx"], state["x_in"], oi["x"]["exp_avg"], oi["x"]["exp_avg_sq"], a6)
    eng.run(1)
    x, sc, sh, ls = eng.get_state()
    return eng.dbg_read("unet.out"), eng.dbg_buffer("grad"), ls, eng


def _compare(tag, pipe, o32, o16, imgs, sparses, cfg, state32, t, st32, st16):
    from helpers import rel_l2

    r16 = _oracle_step_from(o16, st16, t, state32)
    v, grad, loss, eng = _engine_step_from(pipe, imgs, sparses, cfg, state32)
    out = {}
    # The latent gradient is the back-projection of sign(dense - guide) + 2 (dense - guide) at the valid points: one L1
    # sign flip (a point whose residual is within bf16 noise of zero) changes it by ~sqrt(2 / points) in relative L2
    # (6 % for 500 points), for torch's own bf16 run just as for the engine.  The bar for `grad` therefore allows two such
    # flips on top of the torch-bf16 yardstick; `v` (no sign function on its path) is held to the yardstick itself.
    npts = int((sparses > 0).sum().item())
    flip = (2.0 / max(npts, 1)) ** 0.5
    for name, ours, t16, ref, extra in (("v", v, r16["v"], state32["v"], 0.0), ("grad", grad, r16["grad"], state32["grad"], 2.0 * flip)):
        e_ours, e_16 = rel_l2(ours, ref), rel_l2(t16, ref)
        out[name] = (e_ours, e_16)
        assert torch.isfinite(ours).all(), f"{tag} {name}: non-finite"
        assert e_ours <= 1.3 * e_16 + 3e-3 + extra, f"{tag} {name}: engine {e_ours:.3e} vs torch-bf16 {e_16:.3e} (rel L2 to the fp32 oracle)"
    l32, l16 = state32["losses"].float().cpu(), r16["losses"].float().cpu()
    for i in range(cfg["N"]):
        # The loss sums |d| + d^2 over 500 points whose residuals sit near the L1 kink.  Measured over repeated runs (the
        # fp32 oracle's own 50-step trajectory differs from run to run on the GPU): engine 0.15 - 1.7 % from the fp32 value,
        # torch-bf16 0.1 - 5.8 %.  A 1 % floor failed one run in three whenever torch-bf16 happened to land within 0.1 %;
        # the floor is 2x the largest engine deviation seen.
        tol = max(1.5 * abs(l16[i].item() - l32[i].item()), 3e-2 * l32[i].item())
        print(f"[full width] {tag} loss[{i}]: engine {loss[i].item():.5f}, fp32 oracle {l32[i].item():.5f}, torch-bf16 {l16[i].item():.5f}")
        assert abs(loss[i].item() - l32[i].item()) <= tol, f"{tag} loss[{i}]: {loss[i].item():.5f} vs fp32 {l32[i].item():.5f} (bf16 {l16[i].item():.5f})"
    # Adam: elements whose fp32 update direction is unambiguous must move the same way
    xa = eng.dbg_x_adam().float()
    agree = ((xa - state32["x_adam"].float()).abs() < 1.5e-2).float().mean().item()
    agree16 = ((r16["x_adam"].float() - state32["x_adam"].float()).abs() < 1.5e-2).float().mean().item()
    assert agree >= agree16 - 0.03, f"{tag}: Adam agreement {agree:.3f} vs torch-bf16 {agree16:.3f}"
    return out


@pytest.mark.parametrize("name", list(CONFIGS))
def test_first_step_full_width(full, cuda, name):
    cfg = CONFIGS[name]
    o32, o16, pipe = full["o32"], full["o16"], full["pipe"]
    imgs, sparses = _frames(cfg, cuda)
    st32 = o32.preprocess(imgs, sparses, cfg["max_depth"], 0.0, "minmax", cfg["res"], 2024, None, 0.9)
    st16 = o16.preprocess(imgs, sparses, cfg["max_depth"], 0.0, "minmax", cfg["res"], 2024, None, 0.9)
    # same initial latent for all three: the bf16 draw of the seeded generator (what the engine and the bf16 oracle use)
    N = cfg["N"]
    x0 = st16["x"].float()
    zero = lambda ref: dict(exp_avg=torch.zeros_like(ref), exp_avg_sq=torch.zeros_like(ref))
    one = torch.ones(N, 1, 1, 1, device=cuda)
    start = dict(idx=0, x_in=x0, opt_in=dict(x=zero(x0), scales=zero(one), shifts=zero(one), scales_in=one, shifts_in=torch.zeros_like(one)))
    o32.scheduler.set_timesteps(50, device=cuda)
    t = o32.scheduler.timesteps[0]
    state = _oracle_step_from(o32, st32, t, start)
    res = _compare(name, pipe, o32, o16, imgs, sparses, cfg, state, t, st32, st16)
    print(f"[full width] {name}: rel-L2 to fp32 oracle (engine / torch-bf16): v {res['v'][0]:.3e} / {res['v'][1]:.3e}, "
          f"grad {res['grad'][0]:.3e} / {res['grad'][1]:.3e}")


def test_later_steps_full_width(full, cuda):
    """Config (b): 50 steps of the fp32 oracle on the GPU, then the engine and the bf16 oracle are teacher-forced from its
    state (latent, Adam moments, scale / shift) at steps 10, 25 and 49."""
    cfg = CONFIGS["b_nyu_res768"]
    o32, o16, pipe = full["o32"], full["o16"], full["pipe"]
    imgs, sparses = _frames(cfg, cuda)
    keep = {}
    o32(imgs, sparses, cfg["max_depth"], steps=50, resolution=cfg["res"],
        trace=lambda r: keep.__setitem__(r["idx"], r) if r["idx"] in (10, 25, 49) else None)
    st32 = o32.preprocess(imgs, sparses, cfg["max_depth"], 0.0, "minmax", cfg["res"], 2024, None, 0.9)
    st16 = o16.preprocess(imgs, sparses, cfg["max_depth"], 0.0, "minmax", cfg["res"], 2024, None, 0.9)
    o32.scheduler.set_timesteps(50, device=cuda)
    for k in (10, 25, 49):
        # the engine keeps latent and moments in bf16 (like the reference's bf16 mode): start all three from the rounded state
        s = keep[k]
        s["x_in"] = s["x_in"].bfloat16().float()
        for f in ("exp_avg", "exp_avg_sq"):
            s["opt_in"]["x"][f] = s["opt_in"]["x"][f].bfloat16().float()
        t = o32.scheduler.timesteps[k]
        state = _oracle_step_from(o32, st32, t, s)
        res = _compare(f"step {k}", pipe, o32, o16, imgs, sparses, cfg, state, t, st32, st16)
        print(f"[full width] step {k}: rel-L2 to fp32 oracle (engine / torch-bf16): v {res['v'][0]:.3e} / {res['v'][1]:.3e}, "
              f"grad {res['grad'][0]:.3e} / {res['grad'][1]:.3e}")


def test_tapes_full_width(full, cuda):
    """The full-width UNet and VAE-decoder tapes at config (b) (latent 72x96, decoder output 576x768), forward and
    input-gradient backward from a RANDOM output gradient (no sign function anywhere, so nothing is chaotic): every
    full-size kernel of the guided step against torch autograd in fp32, with torch-bf16 as the yardstick."""
    from helpers import rel_l2

    cfg = CONFIGS["b_nyu_res768"]
    o32, o16, pipe = full["o32"], full["o16"], full["pipe"]
    imgs, sparses = _frames(cfg, cuda)
    pipe(imgs, sparses, cfg["max_depth"], steps=50, resolution=cfg["res"], _begin_only=True)
    eng = list(pipe._engines.values())[-1]
    g = torch.Generator(device=cuda).manual_seed(77)

    def check(name, ours, t16, ref):
        e_ours, e_16 = rel_l2(ours, ref), rel_l2(t16, ref)
        print(f"[full width] tape {name}: rel-L2 to fp32 (engine / torch-bf16) {e_ours:.3e} / {e_16:.3e}")
        assert e_ours <= 1.25 * e_16 + 2e-3 and e_ours < 6e-2, f"{name}: engine {e_ours:.3e} vs torch-bf16 {e_16:.3e}"

    # --- decoder
    z = torch.randn(1, 4, eng.lh, eng.lw, device=cuda, generator=g).bfloat16().float()
    x = z.clone().requires_grad_(True)
    y = o32.vae.decode(x)
    dout = torch.randn(y.shape, device=cuda, generator=g).bfloat16().float()
    y.backward(dout)
    x16 = z.bfloat16().requires_grad_(True)
    y16 = o16.vae.decode(x16)
    y16.backward(dout.bfloat16())
    got, din = eng.dbg_forward(1, 0, z), eng.dbg_backward(1, dout)
    check("decoder fwd", got, y16, y.detach())
    check("decoder bwd", din, x16.grad, x.grad)
    del y, y16
    # --- UNet at an early and a late step
    from depth_completion_b200 import ddim
    for step in (0, 40):
        t = torch.tensor(int(ddim.trailing_timesteps(50)[step]), device=cuda)
        xin = torch.randn(1, 8, eng.lh, eng.lw, device=cuda, generator=g).bfloat16().float()
        x = xin.clone().requires_grad_(True)
        y = o32.unet(x, t, o32.empty_text_embedding)
        dout = torch.randn(y.shape, device=cuda, generator=g).bfloat16().float()
        y.backward(dout)
        x16 = xin.bfloat16().requires_grad_(True)
        y16 = o16.unet(x16, t, o16.empty_text_embedding)
        y16.backward(dout.bfloat16())
        got, din = eng.dbg_forward(0, step, xin), eng.dbg_backward(0, dout)
        check(f"unet fwd (step {step})", got, y16, y.detach())
        check(f"unet bwd (step {step})", din, x16.grad, x.grad)


def test_full_width_step_is_deterministic(full, cuda):
    """Config (b) at full width: two guided steps from the same state give bit-identical v, gradient and latent.  The
    step graph has two branches (side stream) and every reduction is fixed-order: any race or missing join shows here."""
    cfg = CONFIGS["b_nyu_res768"]
    pipe = full["pipe"]
    imgs, sparses = _frames(cfg, cuda)
    N = cfg["N"]
    g = torch.Generator(device=cuda).manual_seed(11)
    pipe(imgs, sparses, cfg["max_depth"], steps=50, resolution=cfg["res"], _begin_only=True)
    eng = list(pipe._engines.values())[-1]
    x0 = torch.randn(N, 4, eng.lh, eng.lw, device=cuda, generator=g).bfloat16().float()
    zero = torch.zeros_like(x0)
    a6 = torch.tensor([[1.0] * N, [0.0] * N, [0.0] * N, [0.0] * N, [0.0] * N, [0.0] * N]).numpy()
    outs = []
    for _ in range(2):
        pipe(imgs, sparses, cfg["max_depth"], steps=50, resolution=cfg["res"], _begin_only=True)
        eng = list(pipe._engines.values())[-1]
        eng.dbg_set_state(3, x0, zero, zero, a6)
        eng.run(2)
        x, sc, sh, ls = eng.get_state()
        outs.append((eng.dbg_read("unet.out").clone(), eng.dbg_buffer("grad").clone(), x.clone()))
    for a, b, what in zip(outs[0], outs[1], ("v", "gradient", "latent")):
        assert torch.equal(a, b), f"{what} differs between two identical runs"
