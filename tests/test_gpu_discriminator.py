"""U-Net discriminator on the implicit-GEMM convolution kernels vs the reference golden vector and the oracle."""
import importlib

import pytest
import torch

from golden_util import discriminator_case, rel_l2

pytestmark = pytest.mark.gpu


def _build(cfg, params):
    disc = importlib.import_module("3dhumangan_b200.modules.discriminator")
    D = disc.UNetDiscriminator(**cfg).cuda()
    D.load_state_dict(params, strict=True)
    D.train()
    return D


def test_matches_reference_golden():
    cfg, params, img, gold = discriminator_case("d_tiny")
    D = _build(cfg, params)
    with torch.no_grad():
        out = D(img.cuda(), None, alpha=1.0, **cfg)
    torch.cuda.synchronize()
    for k in ("prediction", "latents", "segments"):
        assert out[k].shape == gold[k].shape, k
        assert rel_l2(out[k].cpu(), gold[k]) < 1e-3, (k, rel_l2(out[k].cpu(), gold[k]))


@pytest.mark.parametrize("hw,B,passes,tol", [((128, 128), 2, 3, 1e-3), ((128, 64), 1, 3, 1e-3), ((64, 64), 2, 1, 5e-2)])
def test_matches_oracle(pkg, port, hw, B, passes, tol):
    dops = importlib.import_module("3dhumangan_b200.modules.discriminator_ops")
    cfg = pkg.configs.baseline_config("C2")
    cfg.update(gen_height=hw[0], gen_width=hw[1])
    params = port.init_discriminator_params(cfg, seed=13)
    img = torch.randn(B, 3, *hw, generator=torch.Generator().manual_seed(5)).clamp(-1, 1)
    stats = {}
    with torch.no_grad():
        ref = port.discriminator_forward(params, img, cfg, training=True, stats_out=stats)
    D = _build(cfg, params)
    out = dops.discriminator_forward(D, img.cuda(), passes=passes)
    torch.cuda.synchronize()
    for k in ("prediction", "latents", "segments"):
        e = rel_l2(out[k].cpu(), ref[k])
        assert e < tol, (k, e)
    if passes == 3:   # spectral-norm buffers advanced exactly like the reference's forward pre-hook
        name = "body_down.1.conv1.1"
        assert rel_l2(D.state_dict()[name + ".weight_u"].cpu(), stats[name + ".weight_u"]) < 1e-5


def test_pool_add_and_dense(pkg):
    abi = importlib.import_module("3dhumangan_b200.abi")
    g = torch.Generator().manual_seed(0)
    a = torch.randn(2, 5, 12, 20, generator=g).cuda()
    b = torch.randn(2, 5, 6, 10, generator=g).cuda()
    import torch.nn.functional as F
    assert torch.allclose(abi.pool_add(a, True, b, False), F.avg_pool2d(a, 2) + b, atol=1e-6)
    assert torch.allclose(abi.pool_add(a, True), F.avg_pool2d(a, 2), atol=1e-6)
    c = torch.randn(2, 5, 12, 20, generator=g).cuda()
    assert torch.allclose(abi.pool_add(a, True, c, True), F.avg_pool2d(a, 2) + F.avg_pool2d(c, 2), atol=1e-6)
    x = torch.randn(11, 3000, generator=g).cuda()
    w = torch.randn(37, 3000, generator=g).cuda()
    bias = torch.randn(37, generator=g).cuda()
    assert torch.allclose(abi.dense(x, w, bias), x @ w.t() + bias, rtol=1e-4, atol=1e-3)
