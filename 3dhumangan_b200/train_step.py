"""One training iteration of the shipped 512-pixel curricula on the sm_100a kernels: discriminator step, then
generator step (PhaseTrainer.train_discriminator / train_generator, lib/trainers/phase_trainer.py:297-344, with
`_train_discriminator` :344-444 and `_train_generator` :446-560 for gan_lambda = 0, latent_lambda = 0,
segmentation_lambda = 1, r1_lambda = 0 -- configs/map3d.py:98-191; with r1_lambda = 0 the R1 term is identically
zero, so it is not evaluated).

Everything between the inputs and the two losses runs on this library's kernels (generator: fused inference kernels
in the discriminator step, training kernels in the generator step; discriminator: the autograd graph of
modules/discriminator_train.py).  The loss itself (class-balanced cross entropy), gradient clipping and Adam are the
trainer's own torch code -- SURVEY.md §8f row 1, outside the hot path.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def segmentation_loss(segments, gt, label_dim, prior_weights=None):
    """`cross_entropy_balanced` of PhaseTrainer._calculate_segmentation_loss (phase_trainer.py:203-256)."""
    if not bool((gt > 0).any()):
        return F.cross_entropy(segments, gt)
    pw = torch.ones(label_dim, dtype=segments.dtype, device=segments.device) if prior_weights is None else prior_weights
    pw = pw / pw.mean()
    one_hot = F.one_hot(gt, num_classes=label_dim).permute(0, 3, 1, 2)
    occ = one_hot.sum(dim=(0, 2, 3))
    occ[0] = 0
    n_occ = torch.count_nonzero(occ)
    coef = torch.reciprocal(occ.to(segments.dtype)) * one_hot.numel() / (n_occ * one_hot.shape[1])
    coef[0] = 0
    coef[torch.isinf(coef)] = 0
    coef = coef * pw
    return (F.cross_entropy(segments, gt, reduction="none") * coef[gt]).mean()


def make_optimizers(G, D, cfg):
    """Adam with the curriculum's betas / learning rates (phase_trainer.py:57-76; the per-group multipliers of the
    generator are the trainer's business and do not change the cost of a step)."""
    betas = tuple(float(b) for b in cfg.get("betas", (0, 0.9)))
    og = torch.optim.Adam(G.parameters(), lr=cfg.get("gen_lr", 5e-5), betas=betas)
    od = torch.optim.Adam(D.parameters(), lr=cfg.get("disc_lr", 2e-4), betas=betas)
    return og, od


def average_gradients(module, group=None):
    """Data parallelism: the kernels' parameter gradients are written by hand (`.grad` side effects of the autograd
    functions), so DistributedDataParallel's reducer hooks never see them; average them explicitly, one flat
    NCCL all-reduce per call (19.6 MB for G, 105.8 MB for D)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    grads = [p.grad for p in module.parameters() if p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, group=group)
    flat /= dist.get_world_size(group)
    off = 0
    for g in grads:
        g.copy_(flat[off:off + g.numel()].view_as(g))
        off += g.numel()


def discriminator_step(G, D, opt_d, z, cond, real_images, real_labels, cfg):
    opt_d.zero_grad(set_to_none=True)
    with torch.no_grad():
        fake = G(z, cond, **cfg)["rgbs"]
    # The reference runs the discriminator twice (real, generated: phase_trainer.py:390,402).  It has no batch
    # statistics, so one pass over the concatenated batch computes exactly the same outputs and gradients with half
    # the launches (the low-resolution layers are launch-bound).
    B = real_images.shape[0]
    out = D(torch.cat([real_images.detach(), fake], 0), cond, alpha=1.0, **cfg)
    L = cfg["label_dim"]
    loss = (segmentation_loss(out["segments"][:B], real_labels, L)
            + segmentation_loss(out["segments"][B:], torch.zeros_like(real_labels), L)) * cfg["segmentation_lambda"]
    loss.backward()
    average_gradients(D)
    torch.nn.utils.clip_grad_norm_(D.parameters(), cfg["grad_clip"])
    opt_d.step()
    return loss.detach()


def generator_step(G, D, opt_g, z, cond, labels, cfg):
    opt_g.zero_grad(set_to_none=True)
    flags = [p.requires_grad for p in D.parameters()]
    for p in D.parameters():            # the discriminator is not updated here: skip its weight gradients
        p.requires_grad_(False)
    try:
        fake = G(z, cond, **cfg)["rgbs"]
        out = D(fake, cond, alpha=1.0, **cfg)
        loss = segmentation_loss(out["segments"], labels, cfg["label_dim"]) * cfg["segmentation_lambda"]
        loss.backward()
    finally:
        for p, f in zip(D.parameters(), flags):
            p.requires_grad_(f)
    average_gradients(G)
    torch.nn.utils.clip_grad_norm_(G.parameters(), cfg["grad_clip"])
    opt_g.step()
    return loss.detach()


def train_iteration(G, D, opt_g, opt_d, batch, cfg):
    """batch: dict(z_d, z_g, cond, images, labels).  Returns (d_loss, g_loss)."""
    d = discriminator_step(G, D, opt_d, batch["z_d"], batch["cond"], batch["images"], batch["labels"], cfg)
    g = generator_step(G, D, opt_g, batch["z_g"], batch["cond"], batch["labels"], cfg)
    return d, g
