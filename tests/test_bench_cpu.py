"""bench.py's accounting and baseline legs, on the CPU (no GPU, no reference source tree needed at run time)."""
import json
import os

import pytest
import torch

import bench


def test_algorithmic_flops_match_the_survey_table():
    """forward 568,811,520 MAC/ray, training convention 1,481,531,392 MAC/ray at 64+64 samples, k=4;
    half of each at the shipped 32+32; 256^3 grid 16.845 TMAC."""
    assert bench.flops_per_ray(64, 64, 4) == 2.0 * 568_811_520
    assert bench.train_flops_per_ray(64, 64, 4) == 2.0 * 1_481_531_392
    assert bench.flops_per_ray(32, 32, 4) == 2.0 * 284_405_760
    assert 256 ** 3 * (bench.D_MAC + bench.S_MAC) == 16_844_861_734_912  # 16.845 TMAC = 33.69 TFLOP
    # the roofline kernel's per-point figure (geometry chain): 2*(4D+2S)
    assert 2.0 * (4 * bench.D_MAC + 2 * bench.S_MAC) * 524288 == pytest.approx(3.069291003904e12)


def test_mac_counts_follow_the_network_config():
    """D, S, C are the dense MACs of the three 9x256 MLPs of base_pull.yml (encoder widths 39+13, 39, 63+27+3+256).
    SDF / colour: build_mlp_nerf (reference utils.py:11-60, the skip layer's input grows to hidden + in); deform:
    build_mlp_idr (utils.py:63-111, the layer BEFORE the skip shrinks to hidden - in so the concatenation stays 256)."""
    def mlp(d_in, d_out, hidden=256, n_layers=9, skip=4, idr=False):
        tot = 0
        for l in range(n_layers):
            k = d_in if l == 0 else (hidden if idr or l != skip else hidden + d_in)
            n = d_out if l == n_layers - 1 else (hidden - d_in if idr and l + 1 == skip else hidden)
            tot += k * n
        return tot
    assert mlp(39 + 13, 3, idr=True) == bench.D_MAC
    assert mlp(39, 257) == bench.S_MAC
    assert mlp(63 + 27 + 3 + 256, 3) == bench.C_MAC


def test_make_rays_layout():
    r = bench.make_rays(64, frame=5)
    assert r.shape == (64, 9) and torch.allclose(r[:, 3:6].norm(dim=-1), torch.ones(64), atol=1e-6)
    assert torch.all(r[:, 6:8] == 0) and torch.allclose(r[:, 8], torch.full((64,), 5 / 59))
    full = bench.make_rays(0, frame=0, hw=8, all_pixels=True)
    assert full.shape == (64, 9)


def test_committed_ncu_capture_is_readable():
    for mode in ("train", "forward"):
        t, p = bench.ncu_traffic(mode), bench.ncu_capture(mode, "tensor_pipe_pct")
        assert t is not None and t > 1e8
        assert p is not None and 0.0 < p <= 100.0
    assert bench.ncu_traffic("no-such-mode") is None


@pytest.mark.parametrize("mode", ["frame", "grid256"])
def test_reference_latency_legs_run_the_unmodified_reference(mode):
    """BASELINE.md run B2 legs (bench.py --mode frame / grid256, `gpu_reference`), shrunk to a few points on the CPU."""
    from oracle import ref_shims
    if not ref_shims.available():
        pytest.skip("oracle/_ref has not been built")
    ms = bench.reference_latency_ms(mode, "cpu", chunk=32, net_chunk=500, hw=6, res=8)
    assert ms > 0.0


def test_profiles_hold_one_json_line_per_bench_file():
    prof = os.path.join(os.path.dirname(os.path.abspath(bench.__file__)), "profiles")
    names = [n for n in os.listdir(prof) if n.startswith("r2_bench_") and n.endswith(".json")]
    assert names
    for n in names:
        with open(os.path.join(prof, n)) as f:
            line = json.loads(f.read().strip().splitlines()[-1])
        for key in ("metric", "value", "unit", "n_gpus", "higher_is_better", "config", "e2e"):
            assert key in line, (n, key)
