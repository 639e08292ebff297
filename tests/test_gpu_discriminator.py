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


def test_conv_wgrad_and_dgrad_primitives():
    """Conv2dSame (hg_conv2d forward / data gradient, hg_conv2d_wgrad_tap) against torch autograd in fp64."""
    import torch.nn.functional as TF
    dt = importlib.import_module("3dhumangan_b200.modules.discriminator_train")
    g = torch.Generator().manual_seed(31)
    for (B, Cin, Cout, H, W, k) in [(2, 64, 128, 16, 16, 3), (2, 3, 64, 16, 24, 3), (1, 320, 27, 8, 8, 1), (2, 128, 320, 8, 8, 3)]:
        if Cin % 64 != 0 and k * k * Cin > 64:
            continue
        x0 = torch.randn(B, Cin, H, W, generator=g)
        w0 = torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5
        b0 = torch.randn(Cout, generator=g)
        gy = torch.randn(B, Cout, H, W, generator=g)
        xr, wr, br = (t.double().requires_grad_(True) for t in (x0, w0, b0))
        yr = TF.conv2d(xr, wr, br, padding=k // 2)
        yr.backward(gy.double())
        xg, wg, bg = (t.cuda().requires_grad_(True) for t in (x0, w0, b0))
        yg = dt.Conv2dSame.apply(xg, wg, bg, 3)
        yg.backward(gy.cuda())
        torch.cuda.synchronize()
        for name, a, b in (("y", yg.detach(), yr.detach()), ("dx", xg.grad, xr.grad), ("dw", wg.grad, wr.grad), ("db", bg.grad, br.grad)):
            e = (a.cpu().double() - b).norm() / b.norm()
            assert e < 3e-5, (B, Cin, Cout, H, W, k, name, float(e))


def test_training_graph_matches_inference_and_oracle_gradients(port, monkeypatch):
    """UNetDiscriminator under autograd: same outputs as the fused inference path; gradients w.r.t. the image (generator
    step) and every parameter (discriminator step) against fp64 autograd through the restated reference, evaluated with
    the LeakyReLU masks of our forward (the gradient is discontinuous in them, see tests/test_gpu_synthesis_bwd.py)."""
    import torch.nn.functional as TF
    cfg, params, img, gold = discriminator_case("d_tiny")
    D = _build(cfg, params)
    with torch.no_grad():
        ref_out = D(img.cuda(), None, alpha=1.0, **cfg)
    D2 = _build(cfg, params)
    masks = []
    xg = img.cuda().requires_grad_(True)
    out = D2(xg, None, alpha=1.0, hg_record_masks=masks, **cfg)
    for k in ("prediction", "segments", "latents"):
        assert rel_l2(out[k].detach().cpu(), ref_out[k].cpu()) < 1e-4, k
    g = torch.Generator().manual_seed(41)
    ws = {k: torch.randn(out[k].shape, generator=g) for k in ("prediction", "segments", "latents")}
    sum((out[k] * ws[k].cuda()).sum() for k in ws).backward()
    torch.cuda.synchronize()

    pc = {n: (v.clone().double().requires_grad_(True) if v.is_floating_point() else v.clone()) for n, v in params.items()}
    xc = img.clone().double().requires_grad_(True)
    it = iter([torch.where(m.cpu(), 1.0, 0.2).double() for m in masks])
    with monkeypatch.context() as mp:
        mp.setattr(port.F, "leaky_relu", lambda v, slope: v * next(it))
        ro = port.discriminator_forward(pc, xc, cfg, training=True)
    sum((ro[k] * ws[k].double()).sum() for k in ws).backward()
    assert rel_l2(xg.grad.cpu().double(), xc.grad) < 5e-4
    named = dict(D2.named_parameters())
    bad = {}
    for n, p in named.items():
        if pc[n].grad is None:
            continue
        e = rel_l2(p.grad.cpu().double(), pc[n].grad)
        if e > 5e-4:
            bad[n] = e
    assert not bad, sorted(bad.items(), key=lambda t: -t[1])[:8]


@pytest.mark.parametrize("shape", [
    # B, C1, C2, Cout, H, W, up2, pre_lrelu, residual ('', 'full', 'half')
    (2, 64, 0, 64, 4, 128, False, False, ""),
    (1, 128, 0, 128, 6, 256, False, True, "full"),
    (2, 128, 64, 256, 8, 128, True, True, "half"),
    (1, 256, 0, 64, 10, 384, True, True, ""),
    (3, 64, 0, 32, 2, 128, False, False, "full"),
])
def test_haloed_conv3x3_matches_torch(shape):
    """dconv_halo.cu (one haloed operand tile per K chunk, nine taps by descriptor row offset): every fused option against
    torch in fp64 -- image borders (zero padding), several tiles per CTA, two accumulator sets, two N sub-blocks, concat."""
    import torch.nn.functional as TF
    abi = importlib.import_module("3dhumangan_b200.abi")
    dops = importlib.import_module("3dhumangan_b200.modules.discriminator_ops")
    B, C1, C2, Cout, H, W, up2, pre, res = shape
    g = torch.Generator().manual_seed(sum(shape[:6]))
    Hs, Ws = (H // 2, W // 2) if up2 else (H, W)
    x1 = torch.randn(B, C1, Hs, Ws, generator=g)
    x2 = torch.randn(B, C2, Hs, Ws, generator=g) if C2 else None
    w = torch.randn(Cout, C1 + C2, 3, 3, generator=g) / (9 * (C1 + C2)) ** 0.5
    bias = torch.randn(Cout, generator=g)
    r = None
    if res == "full":
        r = torch.randn(B, Cout, H, W, generator=g)
    elif res == "half":
        r = torch.randn(B, Cout, H // 2, W // 2, generator=g)
    xin = (x1 if x2 is None else torch.cat([x1, x2], 1)).double()
    if pre:
        xin = TF.leaky_relu(xin, 0.2)
    if up2:
        xin = TF.interpolate(xin, scale_factor=2, mode="nearest")
    ref = TF.conv2d(xin, w.double(), bias.double(), padding=1)
    if r is not None:
        ref = ref + (TF.interpolate(r.double(), scale_factor=2, mode="nearest") if res == "half" else r.double())
    img, Nb = dops._pack_conv(w.cuda())
    for passes, tol in ((3, 2e-5), (1, 2e-2)):
        out = abi.conv2d(x1.cuda(), img, Cout, Nb, ksize=3, H=H, W=W, x2=None if x2 is None else x2.cuda(), up2=up2, pre_lrelu=pre,
                         bias=bias.cuda(), residual=None if r is None else r.cuda(), res_up2=(res == "half"), passes=passes)
        torch.cuda.synchronize()
        e = rel_l2(out.cpu(), ref)
        assert e < tol, (shape, passes, e)


@pytest.mark.parametrize("shape", [(2, 64, 128, 4, 128), (1, 128, 64, 34, 128), (2, 192, 256, 8, 256), (1, 96, 200, 6, 128),
                                   (3, 64, 64, 64, 128)])
def test_haloed_conv3x3_weight_gradient_matches_torch(shape):
    """dconv_wgrad_halo.cu (input rows converted once, read as an MN-major operand, taps by descriptor row offset, ring of input
    rows, several strips per CTA) against torch autograd in fp64; also through the Conv2dSame autograd node."""
    import torch.nn.functional as TF
    abi = importlib.import_module("3dhumangan_b200.abi")
    dt = importlib.import_module("3dhumangan_b200.modules.discriminator_train")
    B, Cin, Cout, H, W = shape
    g = torch.Generator().manual_seed(sum(shape))
    x = torch.randn(B, Cin, H, W, generator=g)
    dy = torch.randn(B, Cout, H, W, generator=g)
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) / (9 * Cin) ** 0.5).double().requires_grad_(True)
    b = torch.zeros(Cout, dtype=torch.float64, requires_grad=True)
    TF.conv2d(x.double(), w, b, padding=1).backward(dy.double())
    for passes, tol in ((3, 2e-5), (1, 2e-2)):
        dw, db = abi.conv2d_wgrad(dy.cuda(), x.cuda(), 3, passes=passes)
        torch.cuda.synchronize()
        assert dw.shape == w.shape
        e = rel_l2(dw.cpu(), w.grad)
        assert e < tol, (shape, passes, e)
        assert rel_l2(db.cpu(), b.grad) < 1e-5
    if Cin % 64 == 0:          # the forward / data-gradient kernel takes channel counts in multiples of 64
        xg = x.cuda().requires_grad_(True)
        wg = w.detach().float().cuda().requires_grad_(True)
        bg = torch.zeros(Cout, device="cuda", requires_grad=True)
        dt.Conv2dSame.apply(xg, wg, bg, 3).backward(dy.cuda())
        assert rel_l2(wg.grad.cpu(), w.grad) < 2e-5
