"""Per-step teacher-forced parity over all 50 guided steps at full width (config b: 480x640, resolution 768).

The fp32 oracle runs the 50 steps on the GPU; at EVERY step the CUDA engine and the torch-bf16 oracle are put into the fp32
oracle's state (latent, Adam moments, scale / shift) and execute that one step.  Reported: relative L2 error to the fp32
oracle of the UNet output v, the predicted clean latent x0 (through the decoder input) and the total latent gradient,
for the engine and for torch-bf16.  Writes a markdown table (argv[1], default profiles/r02_teacher_forced_steps.md)."""
import copy
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch  # noqa: E402

from test_gpu_fullwidth import CONFIGS, _engine_step_from, _frames, _oracle_step_from  # noqa: E402


def rel_l2(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-20)).item()


def main():
    from depth_completion_b200.pipeline import MarigoldDepthCompletionPipeline
    from oracle.marigold_dc import OraclePipeline, make_empty_text_embedding
    from oracle.sd2_modules import AutoencoderKL, UNet2DConditionModel, UNetConfig, VAEConfig

    out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02_teacher_forced_steps.md")
    dev = torch.device("cuda:0")
    cfg = CONFIGS["b_nyu_res768"]
    torch.manual_seed(1234)
    with torch.device(dev):
        unet, vae = UNet2DConditionModel(UNetConfig()), AutoencoderKL(VAEConfig())
    with torch.no_grad():
        for p in list(unet.parameters()) + list(vae.parameters()):
            p.copy_(p.bfloat16().float())
    unet, vae = unet.requires_grad_(False), vae.requires_grad_(False)
    ctx = make_empty_text_embedding(1024, device=dev).bfloat16().float()
    o32 = OraclePipeline(unet, vae, ctx)
    o16 = OraclePipeline(copy.deepcopy(unet).bfloat16(), copy.deepcopy(vae).bfloat16(), ctx.bfloat16())
    pipe = MarigoldDepthCompletionPipeline(unet, vae)
    pipe.empty_text_embedding = ctx
    imgs, sparses = _frames(cfg, dev)
    keep = {}
    o32(imgs, sparses, cfg["max_depth"], steps=50, resolution=cfg["res"], trace=lambda r: keep.__setitem__(r["idx"], r))
    st32 = o32.preprocess(imgs, sparses, cfg["max_depth"], 0.0, "minmax", cfg["res"], 2024, None, 0.9)
    st16 = o16.preprocess(imgs, sparses, cfg["max_depth"], 0.0, "minmax", cfg["res"], 2024, None, 0.9)
    o32.scheduler.set_timesteps(50, device=dev)
    rows = []
    for k in range(50):
        s = keep[k]
        s["x_in"] = s["x_in"].bfloat16().float()
        for f in ("exp_avg", "exp_avg_sq"):
            s["opt_in"]["x"][f] = s["opt_in"]["x"][f].bfloat16().float()
        t = o32.scheduler.timesteps[k]
        ref = _oracle_step_from(o32, st32, t, s)
        r16 = _oracle_step_from(o16, st16, t, s)
        v, grad, loss, eng = _engine_step_from(pipe, imgs, sparses, cfg, ref)
        z = eng.dbg_read("vae.in")[:, :4] * 0.18215  # decoder input = x0 / scaling
        rows.append((k, int(t), rel_l2(v, ref["v"]), rel_l2(r16["v"], ref["v"]), rel_l2(z, ref["x0"]), rel_l2(r16["x0"], ref["x0"]),
                     rel_l2(grad, ref["grad"]), rel_l2(r16["grad"], ref["grad"]), loss[0].item(), r16["losses"][0].item(),
                     ref["losses"][0].item()))
        print(rows[-1], flush=True)
    with open(out_path, "w") as f:
        f.write("# Teacher-forced parity, all 50 guided steps, full width (config b: 480x640, resolution 768, 500 points)\n\n"
                "Produced by `python tools/teacher_forced_table.py` on one B200.  At every step the engine and the torch-bf16 oracle start\n"
                "from the fp32 oracle's state (latent and Adam moments rounded to bf16, as the bf16 modes hold them) and execute that one\n"
                "step; entries are relative L2 errors to the fp32 oracle (`ours / torch-bf16`).  The gradient column is chaotic for BOTH\n"
                "bf16 implementations (one L1 sign flip among 500 points is 6 % on its own) and grows as the loss shrinks.\n\n"
                "| step | t | v ours | v torch-bf16 | x0 ours | x0 torch-bf16 | grad ours | grad torch-bf16 | loss ours | loss torch-bf16 | loss fp32 |\n"
                "|---|---|---|---|---|---|---|---|---|---|---|\n")
        for r in rows:
            f.write(f"| {r[0]} | {r[1]} | {r[2]:.2e} | {r[3]:.2e} | {r[4]:.2e} | {r[5]:.2e} | {r[6]:.2e} | {r[7]:.2e} | {r[8]:.5f} | {r[9]:.5f} | {r[10]:.5f} |\n")
        import statistics as st
        f.write(f"\nmedian over the 50 steps: v {st.median(r[2] for r in rows):.2e} / {st.median(r[3] for r in rows):.2e}, "
                f"x0 {st.median(r[4] for r in rows):.2e} / {st.median(r[5] for r in rows):.2e}, "
                f"grad {st.median(r[6] for r in rows):.2e} / {st.median(r[7] for r in rows):.2e}; "
                f"steps where the engine's v error exceeds torch-bf16's: {sum(r[2] > r[3] for r in rows)} of 50\n")
    print("wrote", out_path)


if __name__ == "__main__":
    main()
