"""Fused renderer (geo.cu + render.cu) vs the oracle's render(): features, rgb, depth, weights."""
from importlib import import_module

import pytest
import torch

from golden_util import generator_case, rel_l2

pytestmark = pytest.mark.gpu


def _run(pkg, port, name, passes=3, over=None):
    ren = import_module("3dhumangan_b200.modules.render_ops")
    cfg, params, cond, z, (u, noise), gold = generator_case(name)
    if over:
        cfg.update(over)
    zz = z if cfg.get("neural_field_latent_input", True) else torch.zeros_like(z)
    with torch.no_grad():
        freq, phase = port.mapping_network(params, zz)
        rgb_r, fmap, depth, w, idx = port.render(params, freq, phase, cond, cfg, u, noise)
    gp = {k: v.cuda() for k, v in params.items()}
    cg = {k: v.cuda() for k, v in cond.items()}
    out = ren.render_forward(gp, freq.cuda(), phase.cuda(), cg, cfg, u.cuda(), noise.cuda(), passes=passes,
                             want_weights=True, want_nearest=True)
    torch.cuda.synchronize()
    B, R = freq.shape[0], cfg["render_width"] * cfg["render_height"]
    ro = out["ray_out"].cpu()
    ref_feat = fmap.permute(0, 2, 3, 1).reshape(B, R, -1)
    ref_rgb = ((rgb_r + 1) / 2).permute(0, 2, 3, 1).reshape(B, R, 3)
    return cfg, ro, ref_feat, ref_rgb, depth, w, idx, out


@pytest.mark.parametrize("name", ["g_tiny_mixed", "g_tiny_dense", "g_tiny_portrait"])
def test_render_matches_oracle(pkg, port, name):
    cfg, ro, ref_feat, ref_rgb, depth, w, idx, out = _run(pkg, port, name)
    assert torch.isfinite(ro).all()
    mism = (out["nearest"].cpu().long() != idx).float().mean()
    assert mism < 2e-3, f"nearest-index mismatch rate {mism:.2e}"
    assert rel_l2(out["weights"].cpu().reshape(w.shape), w) < 1e-3
    assert rel_l2(ro[..., :256], ref_feat) < 1e-3, "feature maps"
    assert rel_l2(ro[..., 256:259], ref_rgb) < 1e-3, "rgb"
    assert rel_l2(ro[..., 259:260], depth) < 1e-4, "depth"


def test_render_last_back_and_no_white(pkg, port):
    cfg, ro, ref_feat, ref_rgb, depth, w, idx, out = _run(pkg, port, "g_tiny_dense",
                                                          over=dict(last_back=True, white_back=False))
    assert rel_l2(ro[..., :256], ref_feat) < 1e-3
    assert rel_l2(ro[..., 259:260], depth) < 1e-4


def test_render_generic_samples_per_ray(pkg, port):
    # S = 16 exercises the shared-memory compositing path (two rays per warp)
    cfg, ro, ref_feat, ref_rgb, depth, w, idx, out = _run(pkg, port, "g_tiny_portrait")
    assert cfg["num_steps"] == 16
    assert rel_l2(ro[..., :256], ref_feat) < 1e-3


def test_render_bf16_mode_is_close(pkg, port):
    cfg, ro, ref_feat, ref_rgb, depth, w, idx, out = _run(pkg, port, "g_tiny_dense", passes=1)
    assert rel_l2(ro[..., :256], ref_feat) < 0.15
