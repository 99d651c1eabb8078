"""The command lines around the hot path: `depth_completion_b200.predict` / `.analyze` / `.vis`
against what `/root/reference/predict.py`, `analyze.py` and `utils.py` do with the same files and options.  CPU tests drive
`predict.run` with a stand-in pipeline; the GPU test runs the real command on random-init SD2 modules."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class ListLog:
    def __init__(self):
        self.lines = []

    def __getattr__(self, level):
        return lambda msg: self.lines.append((level, msg))


def _make_dataset(root, names=("seq_a", "nested/seq_b"), frames=3, H=24, W=32, seed=0):
    from PIL import Image

    from depth_completion_b200 import dataset_io as dio

    g = torch.Generator().manual_seed(seed)
    for ds in names:
        (root / ds / "image" / "cam0").mkdir(parents=True)
        (root / ds / "sparse" / "cam0").mkdir(parents=True)
        for k in range(frames):
            img = torch.randint(0, 256, (H, W, 3), generator=g, dtype=torch.uint8).numpy()
            Image.fromarray(img).save(root / ds / "image" / "cam0" / f"{k:03d}.png")
            d = (5 + torch.rand(H, W, generator=g) * 100) * (torch.rand(H, W, generator=g) < 0.3)
            Image.fromarray(dio.encode_depth_png(d, 120.0)).save(root / ds / "sparse" / "cam0" / f"{k:03d}.png")


def _defaults(**over):
    """The parsed defaults of the predict command (what click hands to main)."""
    from depth_completion_b200.predict import main

    o = {p.name: p.type_cast_value(None, p.default) if p.default is not None else None
         for p in main.params if p.name not in ("src_root", "dst_root")}
    o.update(over)
    return o


def test_option_table_matches_the_reference_cli():
    """Names and defaults of predict.py:25-343 / analyze.py:16-108 (typed out here from the reference's option table)."""
    from depth_completion_b200.analyze import main as analyze_main
    from depth_completion_b200.predict import main as predict_main

    got = {p.name: p.default for p in predict_main.params}
    ref = dict(model="original", vae="light", steps=50, res=768, norm="const", percentile="0.01,0.99", max_sparse_depth=120.0,
               max_depth=120.0, min_depth=0.0, vis=True, vis_res=(512, -1), vis_order="image,sparse,dense", save_dense=True,
               log=None, log_level="INFO", precision="bf16", compile_graph=False, compile_mode="reduce-overhead",
               interp_mode="bilinear", loss_funcs="l1,l2", opt="adam", lr_latent=0.05, lr_scaling=0.005, kld=False,
               kld_mode="simple", kld_weight=0.1, batch_size=1, use_prev_latent=False, beta=0.9, use_segmask=False,
               closed_form=False, projection="linear", inv=False, train_latents=True, train_method="per-step", train_steps=10)
    for k, v in ref.items():
        assert got[k] == v, (k, got[k], v)
    assert got["compress"] is None  # resolves to the reference's "bl2" where blosc2 is installed
    short = {o for p in predict_main.params for o in p.opts}
    assert {"-n", "-r", "-v", "-vr", "-vo", "-p", "-c", "-bs"} <= short
    got = {p.name: p.default for p in analyze_main.params}
    ref = dict(log=None, log_level="INFO", metrics="mae,rmse", calc_binned_scores=True, bin_size=10.0, max_sparse_depth=120.0,
               max_depth=120.0, min_depth=0.0, batch_size=32, num_threads=8, cuda=True)
    for k, v in ref.items():
        assert got[k] == v, (k, got[k], v)


def test_comma_separated_type():
    import click

    from depth_completion_b200.cli_common import CommaSeparated

    assert CommaSeparated(float).convert("0.01, 0.99", None, None) == [0.01, 0.99]
    assert CommaSeparated(str).convert("image,dense", None, None) == ["image", "dense"]
    with pytest.raises(click.BadParameter):
        CommaSeparated(int, n=2).convert("1,2,3", None, None)
    with pytest.raises(click.BadParameter):
        CommaSeparated(float).convert("a,b", None, None)
    with pytest.raises(ValueError):
        CommaSeparated(int, n=0)


def test_resolve_options_follows_the_reference_fallbacks():
    """predict.py:396-455, plus the refusals of arithmetic the B200 path does not have."""
    from depth_completion_b200.predict import OptionError, resolve_options

    log = ListLog()
    o = resolve_options(_defaults(vis_order=["dense", "bogus", "image"], loss_funcs=["l1", "huber", "edge"],
                                  use_prev_latent=True, batch_size=4, projection="log", compress="npz"), log)
    assert o["vis_order"] == ["dense", "image"] and o["loss_funcs"] == ["l1", "edge"]
    assert o["batch_size"] == 1 and o["norm"] == "minmax"
    levels = [lv for lv, _ in log.lines]
    assert levels.count("error") == 3 and levels.count("warning") == 1
    assert resolve_options(_defaults(inv=True, compress="npy"), ListLog())["norm"] == "minmax"
    with pytest.raises(OptionError):
        resolve_options(_defaults(vis_order=["bogus"]), ListLog())
    assert resolve_options(_defaults(vis=False, vis_order=["bogus"], compress="npz"), ListLog())["vis_order"] == ["bogus"]
    for bad in (dict(precision="fp32"), dict(model="lcm"), dict(train_method="per-input"),
                dict(closed_form=True, loss_funcs=["l1", "smooth"])):
        with pytest.raises(OptionError):
            resolve_options(_defaults(compress="npz", **bad), ListLog())
    # branches the library does run: nearest resampling, plain DDIM sampling with the closed-form fit (predict.py:437-455)
    o = resolve_options(_defaults(compress="npz", interp_mode="nearest", train_latents=False, train_method="per-input"), ListLog())
    assert o["interp_mode"] == "nearest" and o["train_latents"] is False and o["closed_form"] is True
    from depth_completion_b200 import dataset_io as dio

    o = resolve_options(_defaults(), ListLog())
    assert o["compress"] == ("bl2" if dio.have_blosc2() else "npz")
    if not dio.have_blosc2():
        with pytest.raises(OptionError):
            resolve_options(_defaults(compress="bl2"), ListLog())
        assert resolve_options(_defaults(compress="bl2", save_dense=False), ListLog())["compress"] == "bl2"


def test_visualisation_panels_and_grid(tmp_path):
    """vis.py against utils.py:370-432 (clamp, normalise, Spectral), :973-1066 (grid, -1 keeps the aspect) and
    predict.py:734-757 (missing sparse pixels are black)."""
    from PIL import Image

    from depth_completion_b200 import vis

    d = torch.tensor([[[[-5.0, 0.0, 60.0, 120.0, 500.0]]]])
    v = vis.visualize_depth(d, max_depth=120.0, min_depth=0.0)
    assert v.shape == (1, 3, 1, 5) and v.dtype == torch.uint8
    assert v[0, :, 0, 0].tolist() == v[0, :, 0, 1].tolist() == [158, 1, 66]      # clamped low end = first anchor
    assert v[0, :, 0, 3].tolist() == v[0, :, 0, 4].tolist() == [94, 79, 162]     # clamped high end = last anchor
    assert v[0, :, 0, 2].tolist() == [255, 255, 191]                              # the middle anchor
    q = vis.colormap_spectral(torch.tensor([0.05]), as_bytes=False)[0]            # half way between anchors 0 and 1
    assert torch.allclose(q, (torch.tensor([158.0, 1, 66]) + torch.tensor([213.0, 62, 79])) / 2 / 255, atol=1e-6)
    with pytest.raises(ValueError):
        vis.visualize_depth(d, max_depth=1.0, min_depth=1.0)
    with pytest.raises(ValueError):
        vis.visualize_depth(d[0], max_depth=1.0)

    img = torch.randint(0, 256, (3, 20, 30), dtype=torch.uint8)
    sparse = torch.zeros(1, 20, 30)
    sparse[0, 3, 4] = 50.0
    dense = torch.full((1, 20, 30), 30.0)
    panels = vis.frame_panels(img, sparse, dense, ["image", "sparse", "dense"], max_depth=120.0)
    assert len(panels) == 3 and panels[0] is img
    assert int((panels[1] != 0).any(0).sum()) == 1 and (panels[1][:, 3, 4] != 0).any()
    assert (panels[2] == panels[2][:, :1, :1]).all()
    with pytest.raises(ValueError):
        vis.frame_panels(img, sparse, dense, ["depth"], max_depth=120.0)
    grid = vis.make_grid(panels)
    assert grid.shape == (3, 20 + 4, 3 * 30 + 8)                                  # torchvision's 2-pixel padding
    small = vis.make_grid(panels, resize=(48, -1))
    assert small.shape == (3, 48, int(48 * grid.shape[2] / grid.shape[1]))
    assert vis.make_grid(panels, resize=(-1, -1)).shape == grid.shape
    with pytest.raises(ValueError):
        vis.make_grid([])
    with pytest.raises(ValueError):
        vis.make_grid(panels, resize=(10, 10), interpolation="cubic")
    vis.save_img_tensor(grid, tmp_path / "a" / "g.jpg")
    with Image.open(tmp_path / "a" / "g.jpg") as im:
        assert im.size == (grid.shape[2], grid.shape[1])
    with pytest.raises(ValueError):
        vis.save_img_tensor(torch.full((3, 4, 4), 2.0), tmp_path / "b.png")
    with pytest.raises(ValueError):
        vis.save_img_tensor(torch.zeros(3, 4, 4, dtype=torch.int32), tmp_path / "b.png")


def test_load_many_keeps_order_and_failures(tmp_path):
    """dataset_io.load_many (the parallel batch loader of utils.py:817-970): results in input order, None for a file that
    cannot be decoded, same result on one thread and on many."""
    from PIL import Image

    from depth_completion_b200 import dataset_io as dio

    paths = []
    for k in range(7):
        p = tmp_path / f"{k}.png"
        if k == 3:
            p.write_bytes(b"not a png")
        else:
            Image.fromarray(np.full((4, 5, 3), 10 * k, np.uint8)).save(p)
        paths.append(p)
    one = dio.load_many(paths, num_threads=1)
    many = dio.load_many(paths, num_threads=8)
    assert [t is None for t in one] == [t is None for t in many] == [k == 3 for k in range(7)]
    for k, (a, b) in enumerate(zip(one, many)):
        if k != 3:
            assert a.shape == (3, 4, 5) and int(a[0, 0, 0]) == 10 * k and torch.equal(a, b)
    assert dio.load_many([], num_threads=4) == []


class EchoPipe:
    """Returns the sparse map it was given (+ an offset) as the dense map: the files around the call are what is tested."""
    device = "cpu"

    def __init__(self, offset=0.0):
        self.calls, self.offset = [], offset

    def __call__(self, imgs, sparses, max_depth, pred_latents_prev=None, beta=0.9, **kw):
        self.calls.append((imgs.shape[0], pred_latents_prev is not None, kw))
        return sparses.clone() + self.offset * (sparses > 0), torch.zeros(imgs.shape[0], 4, 1, 2)


def test_predict_run_then_analyze_round_trip(tmp_path):
    """predict.run writes dense/ and vis/ where predict.py:717-764 puts them and hands the call the arguments of
    predict.py:669-694; analyze then scores them in the reference's JSON layout (analyze.py:300-361)."""
    from click.testing import CliRunner

    from depth_completion_b200 import analyze, predict

    src, dst = tmp_path / "src", tmp_path / "dst"
    _make_dataset(src)
    pipe = EchoPipe(offset=2.0)
    o = predict.resolve_options(_defaults(compress="npz", batch_size=2, steps=7, res=64, norm="minmax", vis_res=(32, -1),
                                          vis_order=["sparse", "dense"], loss_funcs=["l1"]), ListLog())
    saved = predict.run(pipe, src, dst, o, ListLog(), progress=False)
    assert sorted(saved) == ["seq_a", "seq_b"] and all(len(v) == 3 for v in saved.values())
    assert [c[0] for c in pipe.calls] == [2, 1, 2, 1]
    kw = pipe.calls[0][2]
    assert kw["steps"] == 7 and kw["resolution"] == 64 and kw["norm"] == "minmax" and kw["lr"] == (0.05, 0.005)
    assert kw["loss_funcs"] == ["l1"] and kw["percentile"] == (0.01, 0.99) and kw["min_depth"] == 0.0
    assert set(kw) == {"min_depth", "projection", "inv", "norm", "percentile", "steps", "resolution", "interp_mode", "loss_funcs",
                       "opt", "lr", "kld", "kld_mode", "kld_weight", "closed_form", "train_latents", "train_method", "train_steps"}
    assert (dst / "nested" / "seq_b" / "dense" / "cam0" / "002.npz").exists()
    from PIL import Image

    with Image.open(dst / "seq_a" / "vis" / "cam0" / "001_vis.jpg") as im:
        assert im.size[1] == 32 and im.size[0] > 32                      # two panels side by side, 32 rows
    # --save-dense False / --vis False
    o2 = dict(o, save_dense=False, vis=False)
    assert predict.run(EchoPipe(), src / "seq_a", tmp_path / "none", o2, ListLog(), progress=False) == {"seq_a": []}
    assert not (tmp_path / "none" / "dense").exists() and not (tmp_path / "none" / "vis").exists()
    with pytest.raises(predict.OptionError):                            # logger.critical + exit(1) in the command
        predict.run(EchoPipe(), tmp_path / "none", tmp_path / "x", o, ListLog(), progress=False)

    # analyze: every measured point is off by exactly the offset (the sparse PNG quantisation cancels)
    res = CliRunner().invoke(analyze.main, [str(src), str(dst), "--cuda", "False", "--bin-size", "40", "--metrics", "mae,rmse,psnr",
                                            "-bs", "2"])
    assert res.exit_code == 0, res.output
    allr = json.loads((dst / "results_all.json").read_text())
    assert abs(allr["overall"]["mae"] - 2.0) < 1e-4 and abs(allr["overall"]["rmse"] - 2.0) < 1e-4
    assert [tuple(b["range"]) for b in allr["binned"]] == [(0.0, 40.0), (40.0, 80.0), (80.0, 120.0)]
    assert all(set(b) == {"range", "metrics", "percentage"} for b in allr["binned"])
    assert 99.0 < sum(b["percentage"] for b in allr["binned"]) < 103.0   # bin edges are inclusive on both sides
    one = json.loads((dst / "nested" / "seq_b" / "results.json").read_text())
    assert set(one) == {"overall", "binned"} and abs(one["overall"]["mae"] - 2.0) < 1e-4
    res = CliRunner().invoke(analyze.main, [str(src), str(dst), "--cuda", "False", "--calc-binned-scores", "False"])
    assert res.exit_code == 0 and json.loads((dst / "results_all.json").read_text())["binned"] == []
    assert "binned" not in json.loads((dst / "seq_a" / "results.json").read_text())
    res = CliRunner().invoke(analyze.main, [str(src), str(dst), "--cuda", "False", "--metrics", "psnr"])
    assert res.exit_code == 1
    res = CliRunner().invoke(analyze.main, [str(src), str(tmp_path / "none"), "--cuda", "False"])
    assert res.exit_code == 1                                            # no results for any dataset


def test_predict_cli_needs_cuda_and_prints_help():
    from click.testing import CliRunner

    from depth_completion_b200 import predict

    res = CliRunner().invoke(predict.main, ["--help"])
    assert res.exit_code == 0 and "--use-prev-latent" in res.output and "--vis-order" in res.output
    if not torch.cuda.is_available():
        res = CliRunner().invoke(predict.main, [REPO, os.path.join(REPO, "_never_written")])
        assert res.exit_code == 1 and not os.path.exists(os.path.join(REPO, "_never_written"))


@pytest.mark.skipif(not __import__("depth_completion_b200.dataset_io", fromlist=["x"]).have_blosc2(), reason="blosc2 not installed")
def test_bl2_round_trip(tmp_path):
    from depth_completion_b200 import dataset_io as dio

    x = torch.rand(1, 7, 9)
    dio.save_tensor(x, tmp_path / "d.bl2", compress="bl2")
    assert dio.find_dense(tmp_path / "d") == tmp_path / "d.bl2"
    assert np.array_equal(dio.load_dense(tmp_path / "d.bl2"), x.numpy())


def test_bl2_without_blosc2_raises_instead_of_substituting(tmp_path):
    from depth_completion_b200 import dataset_io as dio

    if dio.have_blosc2():
        pytest.skip("blosc2 installed")
    with pytest.raises(RuntimeError):
        dio.save_tensor(torch.zeros(1, 2, 2), tmp_path / "d.bl2", compress="bl2")
    (tmp_path / "e.bl2").write_bytes(b"")
    assert dio.find_dense(tmp_path / "e") == tmp_path / "e.bl2"
    with pytest.raises(RuntimeError):
        dio.load_dense(tmp_path / "e.bl2")
    with pytest.raises(ValueError):
        dio.save_tensor(torch.zeros(1, 2, 2), tmp_path / "d.npy", compress="npz")


@pytest.mark.gpu
def test_predict_and_analyze_commands_on_the_gpu(tmp_path):
    """The two commands as a user runs them, on random-init SD2 modules (`--weights synthetic`, no checkpoint on the box):
    dense maps and panels for every frame, with and without the overlapped prologue (bit-identical files), then the
    scores file."""
    src, dst, dst2 = tmp_path / "src", tmp_path / "dst", tmp_path / "dst2"
    _make_dataset(src, names=("seq_a",), frames=3, H=96, W=128)
    env = dict(os.environ, PYTHONPATH=REPO)
    base = [sys.executable, "-m", "depth_completion_b200.predict", str(src), "--weights", "synthetic", "--vae", "light", "-n", "3",
            "-r", "128", "-c", "npz", "--norm", "minmax", "--max-depth", "120"]
    for out, extra in ((dst, []), (dst2, ["--overlap-prologue", "True", "--vis", "False"])):
        cmd = base[:4] + [str(out)] + base[4:] + extra
        r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-3000:]
    from depth_completion_b200 import dataset_io as dio

    for k in range(3):
        a = dio.load_dense(dst / "seq_a" / "dense" / "cam0" / f"{k:03d}.npz")
        b = dio.load_dense(dst2 / "seq_a" / "dense" / "cam0" / f"{k:03d}.npz")
        assert a.shape == (1, 96, 128) and np.isfinite(a).all() and a.min() >= 0.0 and a.max() <= 120.0
        assert np.array_equal(a, b)
        assert (dst / "seq_a" / "vis" / "cam0" / f"{k:03d}_vis.jpg").exists() and not (dst2 / "seq_a" / "vis").exists()
    # branches of the option table beside the default arm: plain DDIM sampling with the closed-form fit, nearest resampling
    dst3 = tmp_path / "dst3"
    cmd = base[:4] + [str(dst3)] + base[4:] + ["--train-latents", "False", "--interp-mode", "nearest", "--vis", "False", "-c", "npy"]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    c = dio.load_dense(dst3 / "seq_a" / "dense" / "cam0" / "000.npy")
    assert c.shape == (1, 96, 128) and np.isfinite(c).all()
    r = subprocess.run([sys.executable, "-m", "depth_completion_b200.analyze", str(src), str(dst)], env=env, capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    allr = json.loads((dst / "results_all.json").read_text())
    assert set(allr["overall"]) == {"mae", "rmse"} and np.isfinite(allr["overall"]["mae"]) and len(allr["binned"]) == 12
