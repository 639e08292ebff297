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
