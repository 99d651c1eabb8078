"""CPU tests of the host side: C-ABI symbols, config/geometry, DDIM tables, argument validation, sharding (gloo)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from depth_completion_b200 import _lib

    _lib.build()
    lib = _lib.lib()
    syms = []
    for hdr in ("mdc.h", "mdc_debug.h"):
        txt = open(os.path.join(ROOT, "include", hdr)).read()
        syms += re.findall(r"\b(mdc_[a-z0-9_]+)\s*\(", txt)
    syms = sorted(set(syms))
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), s
    assert isinstance(lib.mdc_last_error(), bytes)


def test_config_struct_matches_header():
    """The ctypes mirror of struct mdc_config must list the header's fields in order."""
    from depth_completion_b200.engine import MdcConfig

    txt = open(os.path.join(ROOT, "include", "mdc.h")).read()
    body = txt[txt.index("typedef struct mdc_config {"):txt.index("} mdc_config;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for decl in body.split(";"):
        decl = decl.strip()
        m = re.search(r"\b(int|float)\s+([^{]*)$", decl, flags=re.S)
        if m:
            for part in m.group(2).split(","):
                names.append(re.sub(r"\[.*\]", "", part).strip())
    assert names == [f[0] for f in MdcConfig._fields_]


def test_geometry_and_tables():
    from depth_completion_b200 import ddim
    from depth_completion_b200.config import processed_geometry

    assert processed_geometry(480, 640, 768) == (576, 768, 0, 0)
    assert processed_geometry(480, 640, 640) == (480, 640, 0, 0)
    assert processed_geometry(352, 1216, 1216) == (352, 1216, 0, 0)
    assert processed_geometry(352, 1216, 768) == (222, 768, 2, 0)   # 27 != 28 latent rows
    ts = ddim.trailing_timesteps(50)
    assert ts[0] == 999 and ts[-1] == 19 and len(ts) == 50 and ts.dtype == np.int32
    ac = ddim.alphas_cumprod()
    assert abs(ac[999].item() - 0.0046601) < 1e-6 and abs(ac[0].item() - 0.99915) < 1e-6


def test_tables_match_oracle_scheduler():
    from depth_completion_b200 import ddim
    from oracle.scheduler import DDIMScheduler

    s = DDIMScheduler()
    s.set_timesteps(50)
    assert np.array_equal(ddim.trailing_timesteps(50), s.timesteps.numpy().astype(np.int32))
    assert torch.equal(ddim.alphas_cumprod(), s.alphas_cumprod)
    ac, ts = ddim.tables_from_scheduler(s, 50)
    assert torch.equal(ac, s.alphas_cumprod) and ts[0] == 999


def test_flop_counter_matches_survey():
    from depth_completion_b200.config import UNetConfig, VAEConfig
    from depth_completion_b200.flops import step_flops

    s = step_flops(UNetConfig(), VAEConfig(), 480, 640, 768)
    assert s["latent"] == (72, 96)
    assert abs(s["unet_fwd"] / 1e12 - 1.487) < 2e-3 and abs(s["dec_fwd"] / 1e12 - 4.283) < 2e-3
    assert abs(s["step"] / 1e12 - 11.540) < 5e-3 and abs(s["frame"](50) / 1e12 - 583.2) < 0.2
    s = step_flops(UNetConfig(), VAEConfig(), 352, 1216, 1216)
    assert s["latent"] == (44, 152) and abs(s["step"] / 1e12 - 11.150) < 5e-3


def test_prologue_matches_oracle_on_cpu():
    """Image preprocessing, VAE encoder and masked min/max of the host prologue against the oracle."""
    import torch_reference as prologue
    from depth_completion_b200.config import vae_config_from
    from oracle import image_processor as ip
    from oracle.marigold_dc import masked_minmax
    from oracle.sd2_modules import AutoencoderKL, tiny_vae_config

    torch.manual_seed(0)
    vae = AutoencoderKL(tiny_vae_config())
    img = torch.randint(0, 256, (2, 3, 50, 70), dtype=torch.uint8)
    a, pad = prologue.preprocess_image(img, 64, torch.float32)
    b, pad_ref, _ = ip.preprocess(img, 64, "cpu", torch.float32)
    assert pad == pad_ref and torch.equal(a, b)
    za = prologue.vae_encode_mode(vae.state_dict(), vae_config_from(vae), a)
    zb = vae.encode_mode(b)
    assert torch.allclose(za, zb, atol=1e-5)
    x = torch.rand(3, 40)
    m = torch.rand(3, 40) > 0.6
    lo, hi = prologue.masked_minmax(x, m)
    lo2, hi2 = masked_minmax(x, m, dim=-1)
    assert torch.equal(lo, lo2) and torch.equal(hi, hi2)
    with pytest.raises(ValueError):
        prologue.masked_minmax(x, torch.zeros_like(m))
    with pytest.raises(ValueError):
        prologue.preprocess_image(img.float(), 64, torch.float32)


def test_pipeline_argument_validation_without_gpu():
    """Every ValueError of marigold_dc.py:583-656 fires before any device work, so it can be checked on CPU."""
    from depth_completion_b200._lib import MdcError
    from depth_completion_b200.pipeline import MarigoldDepthCompletionPipeline
    from oracle.sd2_modules import AutoencoderKL, UNet2DConditionModel, tiny_unet_config, tiny_vae_config

    pipe = MarigoldDepthCompletionPipeline(UNet2DConditionModel(tiny_unet_config()), AutoencoderKL(tiny_vae_config()))
    img = torch.zeros(1, 3, 96, 128, dtype=torch.uint8)
    sp = torch.rand(1, 1, 96, 128)
    bad = [dict(train_latents=False, closed_form=False), dict(train_method="x"), dict(train_method="per-input", train_steps=0),
           dict(beta=0.0), dict(norm="percentile", percentile=(0.1, 1.5)), dict(projection="cubic"),
           dict(projection="log10", min_depth=0.0), dict(inv=True), dict(loss_funcs=["l3"]), dict(norm="max"), dict(opt="rmsprop"),
           dict(resolution=128, pred_latents_prev=torch.zeros(2, 4, 12, 16))]
    for kw in bad:
        with pytest.raises(ValueError):
            pipe(img, sp, 10.0, **kw)
    with pytest.raises(ValueError):
        pipe(img, sp[:, :, :, :100], 10.0)
    for kw in [dict(loss_funcs=[]), dict(kld=True, kld_mode="exact")]:  # compute_loss :171-172, utils.py:78-79
        with pytest.raises(ValueError):
            pipe(img, sp, 10.0, **kw)
    for kw in [dict(closed_form=True, loss_funcs=["l1", "edge"]), dict(train_method="per-input"), dict(interp_mode="bicubic")]:
        with pytest.raises(NotImplementedError):
            pipe(img, sp, 10.0, **kw)
    # branches that used to be refused now reach the device (and fail only because there is none here)
    for kw in [dict(opt="adagrad"), dict(kld=True), dict(loss_funcs=["l1", "edge"]), dict(projection="log", min_depth=0.5),
               dict(norm="percentile"), dict(train_latents=False), dict(closed_form=True), dict(interp_mode="nearest")]:
        with pytest.raises(MdcError):
            pipe(img, sp, 10.0, resolution=128, steps=2, **kw)
    with pytest.raises(ValueError):  # 352x1216 at resolution 768
        pipe(torch.zeros(1, 3, 352, 1216, dtype=torch.uint8), torch.rand(1, 1, 352, 1216), 80.0, resolution=768)
    with pytest.raises(MdcError):  # no CPU path exists
        pipe(img, sp, 10.0, resolution=128, steps=2)


def test_shard_frames():
    from depth_completion_b200.pipeline import shard_frames

    for n, w in [(64, 8), (64, 4), (10, 4), (3, 8), (0, 2)]:
        shards = [list(shard_frames(n, r, w)) for r in range(w)]
        assert sum(shards, []) == list(range(n))
        assert max(map(len, shards)) - min(map(len, shards)) <= 1


WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from depth_completion_b200.pipeline import shard_frames
dist.init_process_group("gloo")
r, w = dist.get_rank(), dist.get_world_size()
n_frames = 7
mine = list(shard_frames(n_frames, r, w))
# stand-in for the per-frame result: every rank "completes" its own frames, then outputs are gathered (no collective
# inside the per-frame work), exactly the structure bench.py uses with NCCL
local = torch.zeros(4, 1, 6, 8)
for i, f in enumerate(mine):
    local[i] = float(f) + 1.0
counts = [None] * w
dist.all_gather_object(counts, len(mine))
bufs = [torch.zeros_like(local) for _ in range(w)]
dist.all_gather(bufs, local)
frames = torch.cat([b[:c] for b, c in zip(bufs, counts)], 0)
assert frames.shape[0] == n_frames and [int(v) for v in frames[:, 0, 0, 0]] == list(range(1, n_frames + 1)), frames[:, 0, 0, 0]
t = torch.tensor([float(r + 1)])
dist.all_reduce(t, op=dist.ReduceOp.MAX)
assert t.item() == w
dist.destroy_process_group()
print("ok", r)
'''


def test_two_rank_frame_sharding_gloo(tmp_path):
    """world_size-2 CPU run of the N > 1 host logic: frame shards, output all_gather, max-over-ranks timing reduce."""
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", str(script), ROOT]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=240)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count("ok") == 2


def test_sequence_batches_mirror_the_reference_loop():
    """depth_completion_b200.video.sequence_batches: predict.py:599-603 batching, :423-430 forced batch 1 with the
    temporal prior, frames sharded over ranks only when they are independent."""
    from depth_completion_b200.video import sequence_batches

    assert sequence_batches(5, 2, False) == [(0, 2), (2, 4), (4, 5)]
    assert sequence_batches(5, 4, True) == [(0, 1), (1, 2), (2, 3), (3, 4), (4, 5)]
    assert sequence_batches(0, 3, False) == []
    got = [sequence_batches(64, 4, False, r, 8) for r in range(8)]
    assert sum(len(g) for g in got) == 16 and got[0][0] == (0, 4) and got[7][-1] == (60, 64)
    flat = [f for g in got for (s, e) in g for f in range(s, e)]
    assert flat == list(range(64))
    assert [sequence_batches(10, 4, False, r, 3) for r in range(3)] == [[(0, 4)], [(4, 7)], [(7, 10)]]
    with pytest.raises(ValueError):
        sequence_batches(8, 1, True, 0, 2)
    with pytest.raises(ValueError):
        sequence_batches(8, 0, False)


def test_vae_config_detection_kl_and_tiny():
    """config.vae_config_from: AutoencoderKL vs AutoencoderTiny (diffusers config names), and what is rejected."""
    from depth_completion_b200.config import vae_config_from
    from oracle.sd2_modules import AutoencoderKL, tiny_vae_config
    from oracle.taesd import AutoencoderTiny

    kl = vae_config_from(AutoencoderKL(tiny_vae_config()))
    assert kl.kind == "kl" and kl.block_out_channels == (64, 64, 128, 128) and abs(kl.scaling_factor - 0.18215) < 1e-9
    tiny = vae_config_from(AutoencoderTiny())
    assert tiny.kind == "tiny" and tiny.block_out_channels == (64, 64, 64, 64) and tiny.scaling_factor == 1.0
    assert tiny.num_encoder_blocks == (1, 3, 3, 3) and tiny.num_decoder_blocks == (3, 3, 3, 1) and tiny.latent_magnitude == 3.0
    d = dict(encoder_block_out_channels=(64, 64, 64, 64), decoder_block_out_channels=(64, 64, 64, 64), num_encoder_blocks=(1, 3, 3, 3),
             num_decoder_blocks=(3, 3, 3, 1), latent_magnitude=3, scaling_factor=1.0, act_fn="relu")
    assert vae_config_from(d).kind == "tiny"   # a diffusers-style config dict
    with pytest.raises(ValueError):
        vae_config_from({**d, "decoder_block_out_channels": (64, 64, 64, 32)})
    with pytest.raises(ValueError):
        vae_config_from({**d, "act_fn": "gelu"})


def test_dataset_front_end_round_trip(tmp_path):
    """dataset_io: discovery, pairing, sparse-PNG depth decoding, batch loop and saving (predict.py:512-728) around a
    stand-in pipeline that returns the sparse map it was given."""
    from PIL import Image

    from depth_completion_b200 import dataset_io as dio

    src, dst = tmp_path / "src", tmp_path / "dst"
    g = torch.Generator().manual_seed(0)
    depths = {}
    for ds in ("seq_a", "nested/seq_b"):
        (src / ds / "image" / "cam0").mkdir(parents=True)
        (src / ds / "sparse" / "cam0").mkdir(parents=True)
        for k in range(3):
            img = torch.randint(0, 256, (12, 16, 3), generator=g, dtype=torch.uint8).numpy()
            Image.fromarray(img).save(src / ds / "image" / "cam0" / f"{k:03d}.jpg")
            d = torch.rand(12, 16, generator=g) * 100 * (torch.rand(12, 16, generator=g) < 0.3)
            if k < 2:  # frame 002 has no sparse map -> dropped
                Image.fromarray(dio.encode_depth_png(d, 120.0)).save(src / ds / "sparse" / "cam0" / f"{k:03d}.png")
                depths[(ds.split("/")[-1], k)] = d
    (src / "seq_a" / "image" / "notes.txt").write_text("not an image")
    assert [p.name for p in dio.find_dataset_dirs(src)] == ["seq_b", "seq_a"] or \
        sorted(p.name for p in dio.find_dataset_dirs(src)) == ["seq_a", "seq_b"]
    assert dio.find_dataset_dirs(src / "seq_a") == [src / "seq_a"]
    pairs = dio.find_pairs(src / "seq_a")
    assert [i.name for i, _ in pairs] == ["000.jpg", "001.jpg"] and all(s.suffix == ".png" for _, s in pairs)
    sp = dio.to_depth(torch.stack([dio.load_rgb(s) for _, s in pairs]), 120.0)
    assert sp.shape == (2, 1, 12, 16) and sp.dtype == torch.float32
    assert (sp[0, 0] - depths[("seq_a", 0)]).abs().max() <= 120.0 / 255 / 2 + 1e-5  # 256 levels
    assert ((sp[0, 0] == 0) == (torch.round(depths[("seq_a", 0)] / 120 * 255) == 0)).all()
    assert dio.load_rgb(src / "seq_a" / "image" / "notes.txt") is None

    calls = []

    class EchoPipe:
        device = "cpu"

        def __call__(self, imgs, sparses, max_depth, pred_latents_prev=None, beta=0.9, **kw):
            calls.append((imgs.shape[0], pred_latents_prev is not None, kw))
            assert imgs.dtype == torch.uint8 and imgs.shape[1] == 3 and sparses.shape[1] == 1
            return sparses.clone(), torch.zeros(imgs.shape[0], 4, 1, 2)

    saved = dio.complete_dataset(EchoPipe(), src, dst, 120.0, 120.0, batch_size=2, compress="npz", steps=5)
    assert sorted(saved) == ["seq_a", "seq_b"] and all(len(v) == 2 for v in saved.values())
    assert all(c[0] == 2 and not c[1] and c[2] == {"steps": 5} for c in calls)
    out = dst / "nested" / "seq_b" / "dense" / "cam0" / "001.npz"
    assert out in saved["seq_b"] and out.exists()
    assert np.allclose(dio.load_dense(out)[0], dio.to_depth(dio.load_rgb(src / "nested/seq_b/sparse/cam0/001.png")[None], 120.0)[0, 0])
    calls.clear()
    dio.complete_dataset(EchoPipe(), src / "seq_a", dst / "chain", batch_size=4, use_prev_latent=True, compress=None)
    assert [c[:2] for c in calls] == [(1, False), (1, True)]  # serial chain, batch forced to 1 (predict.py:423-430, :697-699)
    assert (dst / "chain" / "dense" / "cam0" / "000.npy").exists()
    # rank sharding of independent frames: every rank writes its own files, together they cover the dataset
    parts = [dio.complete_dataset(EchoPipe(), src / "seq_a", dst / "sharded", compress="npy", rank=r, world=2)["seq_a"] for r in (0, 1)]
    assert [len(p) for p in parts] == [1, 1] and sorted(p.name for p in parts[0] + parts[1]) == ["000.npy", "001.npy"]
    with pytest.raises(ValueError):  # a serial chain cannot be sharded
        dio.complete_dataset(EchoPipe(), src / "seq_a", dst / "bad", use_prev_latent=True, rank=0, world=2)
    # metrics.evaluate_dataset (analyze.py:225-300): the echo pipeline returned the sparse maps, so the error at the
    # measured points is exactly zero; shifting one stored map by a constant gives that constant back
    from depth_completion_b200 import metrics as mt

    res = mt.evaluate_dataset(src / "seq_a", dst / "seq_a", 120.0, 0.0, 120.0, bin_size=40.0, batch_size=2)
    assert res["overall"] == {"mae": 0.0, "rmse": 0.0} and res["num_points"] > 0 and len(res["bins"]) == 3
    assert sum(b["num_points"] for b in res["bins"]) >= res["num_points"]  # bin edges are inclusive on both sides
    f = dst / "seq_a" / "dense" / "cam0" / "000.npz"
    np.savez_compressed(f, dio.load_dense(f) + 2.0)
    res = mt.evaluate_dataset(src / "seq_a", dst / "seq_a", 120.0, 0.0, 200.0, batch_size=1)
    assert abs(res["overall"]["mae"] - 1.0) < 1e-5 and abs(res["overall"]["rmse"] - 1.0) < 1e-5  # mean of per-batch 2.0 and 0.0
    a, b = torch.tensor([1.0, 2.0, 5.0]), torch.tensor([0.0, 4.0, 5.0])
    assert mt.mae(a, b).item() == 1.0 and abs(mt.rmse(a, b).item() - (5 / 3) ** 0.5) < 1e-6
    assert mt.mae(a, b, torch.tensor([True, False, False])).item() == 1.0
    assert mt.calc_bins(0.0, 100.0, 40.0) == [(0.0, 40.0), (40.0, 80.0), (80.0, 100.0)]
    with pytest.raises(ValueError):
        mt.calc_bins(5.0, 5.0, 1.0)
    with pytest.raises(ValueError):
        dio.save_tensor(torch.zeros(2), tmp_path / "x.npy", compress="npz")
    with pytest.raises(RuntimeError):
        dio.save_tensor(torch.zeros(2), tmp_path / "x.bl2", compress="bl2")
    with pytest.raises(FileNotFoundError):
        dio.complete_dataset(EchoPipe(), tmp_path / "dst", tmp_path / "nowhere")
