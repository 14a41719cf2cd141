"""the reference's own trainer (byte-compiled, unmodified: oracle/_ref) runs on this renderer.

``EndoSurfTrainer.train_step`` (reference src/trainer/trainer_endosurf.py:94-181) is executed as is; only the two
module globals it constructs its collaborators from are replaced: ``EndoSurfRenderer`` -> ``endosurf_b200`` and
``Dataset`` -> the synthetic scene of ``endosurf_b200.harness`` (there is no EndoNeRF data in this environment)."""
import copy
import importlib
import warnings

import numpy as np
import pytest
import torch
import yaml

from conftest import load_cfg

pytestmark = pytest.mark.gpu


def _trainer(tmp_path, ray_batch=256, ns=16, ni=16, reference_renderer=False, precision_terms=3, n_iter=100):
    tmp_path.mkdir(parents=True, exist_ok=True)
    from oracle import ref_shims
    if not ref_shims.available():
        pytest.skip("oracle/_ref (byte-compiled reference) has not been built: python oracle/build_ref.py")
    ref_shims.install_shims()
    tb = importlib.import_module("src.trainer.trainer_basic")
    te = importlib.import_module("src.trainer.trainer_endosurf")
    from endosurf_b200.harness import patch_reference_trainer
    patch_reference_trainer(te, tb, n_frames=8, hw=(96, 96))
    if precision_terms != 3:  # the single-pass fp16 mode of this renderer (constructor keyword the reference lacks)
        ours_cls = te.EndoSurfRenderer
        te.EndoSurfRenderer = lambda *a, **k: ours_cls(*a, precision_terms=precision_terms, **k)
    if reference_renderer:  # the reference's own renderer (stock PyTorch on the GPU); only the dataset is the stand-in
        te.EndoSurfRenderer = importlib.import_module("src.renderer.endosurf").EndoSurfRenderer
    base = load_cfg()
    cfg = {
        "exp": {"project_name": "endosurf", "exp_name": "b200_test", "exp_dir": str(tmp_path / "logs")},
        "data": {"info_dir": "synthetic", "normalize_time": True},
        "render": dict(copy.deepcopy(base["render"]), n_samples=ns, n_importance=ni),
        "train": {"n_iter": n_iter, "ray_batch": ray_batch, "mask_guided_ray_sampling": True, "color_loss_weight": 1.0,
                  "depth_loss_weight": 1.0, "sdf_loss_weight": 1.0, "angle_loss_weight": 0.1,
                  "eikonal_loss_weight": 0.1, "surf_neig_loss_weight": 0.1, "surf_neig_rad": 0.1, "resume": False,
                  "optim": {"lr": 5e-4, "lr_alpha": 0.05, "warm_up_end": 50}, "eval": {"ray_chunk": 2048}},
        "net": copy.deepcopy(base["net"]),
        "log": {"summary_writer": {"type": "tensorboard"}, "i_eval": 0, "i_save": 0},
    }
    path = tmp_path / "cfg.yml"
    with open(path, "w") as f:
        yaml.safe_dump(cfg, f)
    torch.manual_seed(0)
    np.random.seed(0)
    return te.EndoSurfTrainer(str(path)), te


class _Recorder:
    """Stand-in for the trainer's summary writer: keeps the scalars of the last step."""

    def __init__(self):
        self.scalars = {}

    def add_scalar(self, tag, value, global_step):
        self.scalars[tag] = float(value.item() if torch.is_tensor(value) else value)


def test_first_step_losses_match_the_reference_renderer(tmp_path):
    """Drop-in check through the caller: the unmodified trainer computes the same loss terms on its first step whether
    it drives the reference's own renderer (on the GPU, stock PyTorch) or this one, from the same checkpoint, frame,
    rays and jitter (same RNG stream up to the neighbour sampling of surface_neighbour_error)."""
    import endosurf_b200
    ref, _ = _trainer(tmp_path / "ref", reference_renderer=True)
    ours, te = _trainer(tmp_path / "ours")
    ref_cls = importlib.import_module("src.renderer.endosurf").EndoSurfRenderer
    assert isinstance(ours.renderer, endosurf_b200.EndoSurfRenderer) and isinstance(ref.renderer, ref_cls)
    ours.renderer.load_checkpoint(ref.renderer.save_checkpoint())
    logs = {}
    for name, tr in (("ref", ref), ("ours", ours)):
        tr.writer = _Recorder()
        tr.renderer.train()
        torch.manual_seed(7)
        np.random.seed(7)
        tr.train_step(global_step=3000)
        logs[name] = dict(tr.writer.scalars)
    ours.renderer.sync_check()
    print("first-step scalars (ref, ours):", {k: (logs["ref"][k], logs["ours"][k]) for k in logs["ref"]})
    for k in ["train/loss_color", "train/loss_depth", "train/loss_sdf", "train/loss_angle", "train/loss_eikonal",
              "train/psnr_color", "train/s_val", "train/cdf", "train/weight_max"]:
        a, b = logs["ref"][k], logs["ours"][k]
        assert abs(a - b) <= 2e-3 * max(abs(a), 1e-3), (k, a, b)
    # the neighbour points are drawn from differently shaped random tensors: same statistic, not the same numbers
    a, b = logs["ref"]["train/loss_surf_neig"], logs["ours"]["train/loss_surf_neig"]
    assert abs(a - b) <= 0.5 * max(abs(a), 1e-3), ("train/loss_surf_neig", a, b)


def test_reference_trainer_train_steps_and_checkpoint(tmp_path):
    import endosurf_b200
    trainer, te = _trainer(tmp_path)
    assert isinstance(trainer.renderer, endosurf_b200.EndoSurfRenderer)
    trainer.renderer.train()
    trainer.dset.list_train = trainer.dset.list_train[:1]  # one frame: the loss of consecutive steps is comparable
    losses = []
    for it in range(1, 17):
        losses.append(trainer.train_step(global_step=it))
        trainer.update_learning_rate(it)
    trainer.renderer.sync_check()
    assert all(np.isfinite(losses)), losses
    assert np.mean(losses[-4:]) < np.mean(losses[:4]), f"loss did not decrease: {losses}"
    # checkpoint round trip through the trainer's own save / load (reference state-dict keys)
    trainer.save_checkpoint(16)
    before = {k: v.detach().clone() for k, v in trainer.renderer.state_dict().items()}
    with torch.no_grad():
        for p in trainer.renderer.parameters():
            p.add_(1.0)
    trainer.load_checkpoint()
    assert trainer.step_start == 17
    for k, v in trainer.renderer.state_dict().items():
        assert torch.equal(v, before[k]), k
    # and the restored model still trains
    assert np.isfinite(trainer.train_step(global_step=17))


def test_host_syncs_per_train_step(tmp_path):
    """8f-4: the renderer itself is sync-free; what is left are the trainer's own .item() calls (13 add_scalar
    conversions + loss.item(), reference trainer_endosurf.py:104,165-179, utils.py:96-97) - measured, not fixed,
    because the trainer is run unchanged."""
    trainer, te = _trainer(tmp_path)
    trainer.renderer.train()
    trainer.train_step(global_step=1)
    torch.cuda.synchronize()
    torch.cuda.set_sync_debug_mode("warn")
    try:
        with warnings.catch_warnings(record=True) as w:
            warnings.simplefilter("always")
            trainer.train_step(global_step=2)
    finally:
        torch.cuda.set_sync_debug_mode("default")
    syncs = [x for x in w if "synchroniz" in str(x.message).lower()]
    n_renderer = 0
    for x in syncs:
        if "endosurf_b200" in (x.filename or ""):
            n_renderer += 1
    print(f"host synchronisations in one reference train_step: {len(syncs)} (inside endosurf_b200: {n_renderer})")
    assert n_renderer <= 1, [str(x.filename) + ":" + str(x.lineno) for x in syncs if "endosurf_b200" in (x.filename or "")]
    assert len(syncs) <= 20


def test_loss_curve_parity_200_steps(tmp_path):
    """BASELINE configs[2] ("full EndoSurf training loop ... 1 GPU"): 200 optimisation steps of the UNMODIFIED reference
    trainer, driven three times from the same initial state and the same per-step RNG seeds (frame, rays, jitter):
    with the reference's own renderer (stock PyTorch fp32 on the GPU), with this renderer in the fp32-parity mode
    and in the single-pass fp16 mode (precision_terms=1).  Training amplifies rounding differences (ReLU kinks,
    re-sampling), so the curves are compared as 25-step window means; the band is stated at the assertions."""
    n_steps, win = 200, 25
    terms = ["train/loss_color", "train/loss_depth", "train/loss_sdf", "train/loss_eikonal", "train/loss_total"]

    def run(kind):
        tr, _ = _trainer(tmp_path / kind, reference_renderer=(kind == "ref"),
                         precision_terms=1 if kind == "fp16x1" else 3, n_iter=n_steps)
        return tr

    curves = {}
    init = None
    for kind in ("ref", "fp16x3", "fp16x1"):
        tr = run(kind)
        if init is None:
            init = copy.deepcopy(tr.renderer.save_checkpoint())  # (state dicts alias the live parameters)
        else:
            tr.renderer.load_checkpoint(init)
        tr.writer = _Recorder()
        tr.renderer.train()
        tr.dset.list_train = tr.dset.list_train[:2]
        rec = {k: [] for k in terms}
        for it in range(1, n_steps + 1):
            torch.manual_seed(1000 + it)
            np.random.seed(1000 + it)
            tr.train_step(global_step=it)
            tr.update_learning_rate(it)
            for k in terms:
                rec[k].append(tr.writer.scalars[k])
        if kind != "ref":
            tr.renderer.sync_check()
        curves[kind] = {k: np.asarray(v) for k, v in rec.items()}
    ref = curves["ref"]
    report = {}
    for kind in ("fp16x3", "fp16x1"):
        for k in terms:
            a = ref[k].reshape(-1, win).mean(1)
            b = curves[kind][k].reshape(-1, win).mean(1)
            # relative to the reference window, floored at a tenth of the term's initial level (late in the run single
            # terms are ~1e-3 and the reference itself varies by tens of percent from run to run: atomics)
            report[(kind, k)] = float(np.max(np.abs(a - b) / np.maximum(np.abs(a), 0.1 * np.abs(a[0]))))
    print("loss-curve parity, worst relative deviation of a 25-step window mean from the reference's:",
          {f"{kk[0]}:{kk[1].split('/')[-1]}": round(v, 4) for kk, v in report.items()})
    print("total loss, first 12 steps ref / fp16x3 / fp16x1:",
          [np.round(curves[c]["train/loss_total"][:12], 4).tolist() for c in ("ref", "fp16x3", "fp16x1")])
    print("total-loss window means ref / fp16x3 / fp16x1:",
          [np.round(curves[c]["train/loss_total"].reshape(-1, win).mean(1), 4).tolist() for c in ("ref", "fp16x3", "fp16x1")])
    for c in curves:
        t = curves[c]["train/loss_total"]
        assert np.all(np.isfinite(t)), c
        assert t[-win:].mean() < 0.8 * t[:win].mean(), (c, t[:win].mean(), t[-win:].mean())
    # Stated band (measured on B200: total-loss window means 1.197/0.658/0.309/0.134/0.095/0.062/0.047/0.038 against the
    # reference's 1.199/0.647/0.309/0.145/0.094/0.079/0.052/0.040; identical to 0.3 % over the first 12 steps): the total
    # loss's 25-step window means within 5 % of the reference curve over the first 75 steps and within 30 % everywhere
    # (late in the run the loss is small and the trajectories have decorrelated: different neighbour draws in
    # surface_neighbour_error, ReLU kinks, re-sampling); every single loss term's window means within 60 %.
    for kind in ("fp16x3", "fp16x1"):
        a = ref["train/loss_total"].reshape(-1, win).mean(1)
        b = curves[kind]["train/loss_total"].reshape(-1, win).mean(1)
        rel = np.abs(a - b) / a
        assert np.all(rel[:3] <= 0.05) and np.all(rel <= 0.30), (kind, rel.tolist())
        first = np.abs(ref["train/loss_total"][:12] - curves[kind]["train/loss_total"][:12]) / ref["train/loss_total"][:12]
        assert np.all(first <= 0.01), (kind, first.tolist())
    for (kind, k), v in report.items():
        assert v <= 0.60, (kind, k, v)
