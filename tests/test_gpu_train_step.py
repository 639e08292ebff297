"""One discriminator step + one generator step through the kernels (3dhumangan_b200/train_step.py)."""
import importlib

import pytest
import torch

pytestmark = pytest.mark.gpu


def test_train_iteration_updates_both_networks(pkg):
    gen = importlib.import_module("3dhumangan_b200.modules.generator")
    disc = importlib.import_module("3dhumangan_b200.modules.discriminator")
    ts = importlib.import_module("3dhumangan_b200.train_step")
    cfg = pkg.configs.baseline_config("tiny")
    cfg.update(gen_height=64, gen_width=64, render_height=8, render_width=8, num_steps=32, nerf_noise=0.5)
    B = 2
    torch.manual_seed(0)
    G = gen.Map3DGenerator(**cfg).cuda().train()
    G.set_device(torch.device("cuda:0"))
    D = disc.UNetDiscriminator(**cfg).cuda().train()
    og, od = ts.make_optimizers(G, D, cfg)
    cond = {k: v.cuda() for k, v in pkg.synthetic.make_conditions(B, seed=1).items()}
    batch = dict(z_d=torch.randn(B, cfg["latent_dim"], device="cuda"), z_g=torch.randn(B, cfg["latent_dim"], device="cuda"), cond=cond,
                 images=torch.randn(B, 3, 64, 64, device="cuda").clamp_(-1, 1),
                 labels=torch.randint(1, cfg["label_dim"], (B, 64, 64), device="cuda"))
    g0 = {n: p.detach().clone() for n, p in G.named_parameters()}
    d0 = {n: p.detach().clone() for n, p in D.named_parameters()}
    losses = [ts.train_iteration(G, D, og, od, batch, cfg) for _ in range(3)]
    torch.cuda.synchronize()
    for d, g in losses:
        assert torch.isfinite(d) and torch.isfinite(g)
    moved_g = [n for n, p in G.named_parameters() if not torch.equal(p.detach(), g0[n])]
    moved_d = [n for n, p in D.named_parameters() if not torch.equal(p.detach(), d0[n])]
    # every parameter that takes part in the forward must have received a gradient and moved
    assert any(n.startswith("neural_field.") for n in moved_g) and any(n.startswith("synthesis_network.") for n in moved_g)
    assert any(n.startswith("neural_field_mapping_network.") for n in moved_g)
    assert any(n.startswith("synthesis_mapping_network.") for n in moved_g)
    assert len(moved_d) > 0.9 * len(d0), (len(moved_d), len(d0))
    for p in list(G.parameters()) + list(D.parameters()):
        assert torch.isfinite(p).all()
    # the discriminator loss on fixed data goes down under its own updates
    assert losses[-1][0] < losses[0][0]
    # no two gradients share a buffer: in-place passes over the gradients (GradScaler.unscale_, clip_grad_norm_) would hit an
    # aliased buffer once per alias -- the nine ToRGB biases all receive sum(d_rgb), once returned as views of one tensor
    for net in (G, D):
        ptrs = [p.grad.untyped_storage().data_ptr() for p in net.parameters() if p.grad is not None]
        assert len(ptrs) > 0 and len(ptrs) == len(set(ptrs))


def _oracle_d_step(port, ts, pg, pd, batch, cfg, u, noise, do_r1):
    """The discriminator step of phase_trainer.py:344-444 (gan_lambda = 0, segmentation loss, optional R1) composed from
    the oracle's forward functions under plain torch autograd -- the checker for `Trainer.train_discriminator`."""
    import torch
    with torch.no_grad():
        fake = port.generator_forward(pg, batch["z_d"], batch["cond"], cfg, u, noise, training=True)["rgbs"]
    P = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "weight_u" not in k and "weight_v" not in k else v.clone())
         for k, v in pd.items()}
    real = batch["images"].clone().requires_grad_(True)
    st = {}
    out_real = port.discriminator_forward(P, real, cfg, training=True, stats_out=st)
    pen = 0.0
    if do_r1:         # gan_lambda > 0: f = sum(prediction)   (with gan_lambda = 0 the reference differentiates
        # sum(softmax(segments)), which is identically the pixel count: that penalty is rounding noise by construction)
        target = out_real["prediction"].sum() if cfg["gan_lambda"] > 0 else torch.softmax(out_real["segments"], dim=1).sum()
        g = torch.autograd.grad(target, real, create_graph=True)[0]
        # the reference's arithmetic (phase_trainer.py:281-288): entry 0 of the batch, mean over its channels
        # (train_step.r1_penalty is pinned against the reference's own method in tests/test_cpu_trainer_pin.py)
        g0 = g[0]
        pen = 0.5 * cfg["r1_lambda"] * g0.reshape(g0.shape[0], -1).pow(2).sum(1).mean()
    P2 = dict(P)
    P2.update({k: v.detach() for k, v in st.items()})          # second pass: power iteration continues from the first
    out_gen = port.discriminator_forward(P2, fake, cfg, training=True)
    L = cfg["label_dim"]
    seg = ts.segmentation_loss(out_real["segments"], batch["labels"], L) + \
        ts.segmentation_loss(out_gen["segments"], torch.zeros_like(batch["labels"]), L)
    loss = seg * cfg["segmentation_lambda"] + 4 * pen
    if cfg["gan_lambda"] > 0:
        F = torch.nn.functional
        loss = loss + cfg["gan_lambda"] * (F.softplus(out_gen["prediction"]).mean() + F.softplus(-out_real["prediction"]).mean())
    loss.backward()
    return loss.detach(), (pen.detach() if do_r1 else None), {k: v.grad for k, v in P.items() if isinstance(v, torch.Tensor) and v.requires_grad}


@pytest.mark.parametrize("do_r1", [False, True])
def test_discriminator_step_matches_oracle_composition(pkg, port, monkeypatch, do_r1):
    """Loss value, R1 penalty (double backward through the discriminator; gan_lambda = 1 so that f = sum(prediction) as at phase_trainer.py:261-266) and parameter
    gradients of `Trainer.train_discriminator` against the same step composed from the oracle under torch autograd."""
    gen = importlib.import_module("3dhumangan_b200.modules.generator")
    disc = importlib.import_module("3dhumangan_b200.modules.discriminator")
    ts = importlib.import_module("3dhumangan_b200.train_step")
    rng = importlib.import_module("3dhumangan_b200.rng")
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        cfg = pkg.configs.baseline_config("tiny")
        cfg.update(gen_height=64, gen_width=64, render_height=8, render_width=8, num_steps=32, nerf_noise=0.5,
                   r1_lambda=1.0, grad_clip=1e9, gan_lambda=1.0 if do_r1 else 0)
        cfg["phases"] = [dict(cfg["phases"][3 if do_r1 else 0])]
        B = 2
        pg = {k: v.cuda() for k, v in port.init_generator_params(cfg, seed=5, sigma_gain=200.0, sigma_bias=1.0).items()}
        pd = {k: v.cuda() for k, v in port.init_discriminator_params(cfg, seed=6).items()}
        G = gen.Map3DGenerator(**cfg).cuda().train()
        G.load_state_dict(pg, strict=True)
        G.set_device(torch.device("cuda:0"))
        D = disc.UNetDiscriminator(**cfg).cuda().train()
        D.load_state_dict(pd, strict=True)
        g = torch.Generator().manual_seed(7)
        batch = dict(z_d=torch.randn(B, cfg["latent_dim"], generator=g).cuda(),
                     cond={k: v.cuda() for k, v in pkg.synthetic.make_conditions(B, seed=8).items()},
                     images=torch.randn(B, 3, 64, 64, generator=g).clamp_(-1, 1).cuda(),
                     labels=torch.randint(0, cfg["label_dim"], (B, 64, 64), generator=g).cuda())
        u, noise = rng.draw_render_noise(B, 64, 32, "cuda", cfg["sample_dist"])
        monkeypatch.setattr(rng, "draw_render_noise", lambda *a, **k: (u, noise))
        ref_loss, ref_pen, ref_grads = _oracle_d_step(port, ts, pg, pd, batch, cfg, u, noise, do_r1)
        t = ts.Trainer(G, D, cfg, amp=False, ddp=False)
        loss = t.train_discriminator(batch)
        torch.cuda.synchronize()
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    assert abs(float(loss) - float(ref_loss)) < 2e-3 * abs(float(ref_loss)), (float(loss), float(ref_loss), ref_pen)
    if do_r1:
        assert 4 * float(ref_pen) > 1e-5 * float(ref_loss), (float(ref_pen), float(ref_loss))     # not identically zero
    errs = {}
    scale = max(float(v.norm()) for v in ref_grads.values() if v is not None)
    for n, p in D.named_parameters():
        r = ref_grads.get(n)
        if r is None or float(r.norm()) < 1e-6 * scale:
            continue
        assert p.grad is not None, n
        errs[n] = float((p.grad.double() - r.double()).norm() / r.double().norm())
    assert len(errs) > 25, len(errs)
    vals = sorted(errs.values())
    assert vals[len(vals) // 2] < 2e-2, (vals[len(vals) // 2], sorted(errs.items(), key=lambda kv: -kv[1])[:5])
    assert vals[-1] < 0.3, sorted(errs.items(), key=lambda kv: -kv[1])[:5]


def test_trainer_amp_gradscaler_and_param_groups(pkg):
    """fp16 autocast + GradScaler around both steps (phase_trainer.py:355,396,462), five learning-rate groups
    (phase_trainer.py:57-76), EMA in parameters() order (ema.py:29-48)."""
    gen = importlib.import_module("3dhumangan_b200.modules.generator")
    disc = importlib.import_module("3dhumangan_b200.modules.discriminator")
    ts = importlib.import_module("3dhumangan_b200.train_step")
    cfg = pkg.configs.baseline_config("tiny")
    cfg.update(gen_height=64, gen_width=64, render_height=8, render_width=8, num_steps=32, nerf_noise=0.5)
    B = 2
    torch.manual_seed(0)
    G = gen.Map3DGenerator(**cfg).cuda().train()
    G.set_device(torch.device("cuda:0"))
    D = disc.UNetDiscriminator(**cfg).cuda().train()
    t = ts.Trainer(G, D, cfg, amp=True, ddp=False)
    groups = {g["name"]: g for g in t.optimizer_G.param_groups}
    assert set(groups) == {"generator", "appearance_codes", "neural_field_mapping", "synthesis_mapping", "neural_field"}
    assert groups["neural_field"]["lr"] == pytest.approx(cfg["gen_lr"] * 0.05)
    assert groups["neural_field_mapping"]["lr"] == pytest.approx(cfg["gen_lr"] * 0.05)
    assert groups["synthesis_mapping"]["lr"] == pytest.approx(cfg["gen_lr"])
    assert sum(len(g["params"]) for g in groups.values()) == len(list(G.parameters()))
    batch = dict(cond={k: v.cuda() for k, v in pkg.synthetic.make_conditions(B, seed=1).items()},
                 images=torch.randn(B, 3, 64, 64, device="cuda").clamp_(-1, 1),
                 labels=torch.randint(1, cfg["label_dim"], (B, 64, 64), device="cuda"))
    p0 = [p.detach().clone() for p in G.parameters() if p.requires_grad]
    for _ in range(4):                    # phases 0..3: the last one is a do_r1 phase (r1_lambda = 0: the graph is still built)
        d, g_ = t.iteration(batch)
        assert torch.isfinite(d) and torch.isfinite(g_)
    assert D.step == 4 and G.step == 4
    moved = sum(int(not torch.equal(a, b.detach())) for a, b in zip(p0, [p for p in G.parameters() if p.requires_grad]))
    assert moved > 200
    # EMA follows the parameters: shadow = p0 + sum of lerps, strictly between the start and the current value where moved
    for s, a, b in list(zip(t.ema.shadow_params, p0, [p for p in G.parameters() if p.requires_grad]))[:20]:
        if not torch.equal(a, b.detach()):
            assert not torch.equal(s, a)
            assert float((s - a).abs().max()) <= float((b.detach() - a).abs().max()) * 1.0001 + 1e-12
