"""`lib.discriminators` surface: UNetDiscriminator (reference: lib/discriminators/unet_discriminators.py:7-160).

Parameter tree, names and initialisation follow the reference (spectral-normed 3x3 / 1x1 convs in
`body_down` / `body_up` ResBlocks, plain heads `layer_up_last`, `output_layer`, `latent_layer`), so a
reference `state_dict` loads strictly.  The forward pass runs the implicit-GEMM convolution kernels
of csrc/dconv.cu when they are present in the library; there is no cuDNN / eager fallback.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn


def _sn(conv, disable):
    return conv if disable else nn.utils.spectral_norm(conv)


class ResBlock(nn.Module):
    """unet_discriminators.py:7-72 (parameter holder; index positions inside the Sequentials match
    the reference so that keys read `conv1.1.*` / `conv1.2.*` / `conv2.1.*`)."""

    def __init__(self, fin, fout, up_or_down, first=False, **kwargs):
        super().__init__()
        self.up_or_down, self.first = up_or_down, first
        self.learned_shortcut = fin != fout
        dis = kwargs.get("disable_spectral_norm", False)
        if first:
            self.conv1 = _sn(nn.Conv2d(fin, fout, 3, 1, 1), dis)
        elif up_or_down > 0:
            self.conv1 = nn.Sequential(nn.LeakyReLU(0.2, False), nn.Upsample(scale_factor=2), _sn(nn.Conv2d(fin, fout, 3, 1, 1), dis))
        else:
            self.conv1 = nn.Sequential(nn.LeakyReLU(0.2, False), _sn(nn.Conv2d(fin, fout, 3, 1, 1), dis))
        self.conv2 = nn.Sequential(nn.LeakyReLU(0.2, False), _sn(nn.Conv2d(fout, fout, 3, 1, 1), dis))
        if self.learned_shortcut:
            self.conv_s = _sn(nn.Conv2d(fin, fout, 1, 1, 0), dis)


class UNetDiscriminator(nn.Module):
    def __init__(self, **kwargs):
        super().__init__()
        self.epoch = 0
        self.step = 0
        self.semantic_dim = kwargs.get("semantic_dim", 0)
        self.label_dim = kwargs.get("label_dim", 0)
        self.latent_dim = kwargs["latent_dim"]
        self.output_dim = self.semantic_dim + self.label_dim
        self.num_blocks = min(kwargs.get("discriminator_blocks", 6),
                              int(math.log2(max(kwargs["gen_height"], kwargs["gen_width"]))) - 1)
        cin = 6 if kwargs.get("dual_discrimination", False) else 3
        self.channels = [cin, 128, 128, 256, 256, 512, 512, 512, 512]
        ch, nb = self.channels, self.num_blocks
        self.body_up = nn.ModuleList([])
        self.body_down = nn.ModuleList([])
        for i in range(nb):
            self.body_down.append(ResBlock(ch[i], ch[i + 1], -1, first=(i == 0), **kwargs))
        self.body_up.append(ResBlock(ch[nb], ch[nb - 1], 1, **kwargs))
        for i in range(1, nb - 1):
            self.body_up.append(ResBlock(2 * ch[nb - i], ch[nb - i - 1], 1, **kwargs))
        self.body_up.append(ResBlock(2 * ch[1], 64, 1, **kwargs))
        self.layer_up_last = nn.Conv2d(64, 1, 1, 1, 0)
        self.output_layer = nn.Conv2d(64, self.output_dim, 1, 1)
        ds = 2 ** nb
        self.latent_layer = nn.Conv2d(ch[nb], self.latent_dim, (kwargs["gen_height"] // ds, kwargs["gen_width"] // ds))
        # `self.apply(kaiming_leaky_init)` (:121, :74-79): for spectral-normed convs the reference re-draws
        # the derived `.weight` attribute, which the next forward overwrites -- `weight_orig` keeps the
        # default Conv2d initialisation; only the three plain heads actually get the kaiming draw.
        for m in (self.layer_up_last, self.output_layer, self.latent_layer):
            nn.init.kaiming_normal_(m.weight, a=0.2, mode="fan_in", nonlinearity="leaky_relu")
        with torch.no_grad():
            self.output_layer.weight *= 0.25
        self._cfg = {k: v for k, v in kwargs.items() if isinstance(k, str)}

    def forward(self, images, conditions, alpha, **kwargs):
        """-> {"prediction": [B,1,H,W], "latents": [B,L], "segments": [B,label_dim,H,W]} (:125-160).
        `conditions`, `alpha` and other kwargs are accepted and ignored, as in the reference."""
        from . import discriminator_ops
        from .generator import _precision_passes
        if torch.is_grad_enabled() and (images.requires_grad or any(p.requires_grad for p in self.parameters())):
            # discriminator / generator step of the trainer: the autograd graph over the sm_100a primitives
            from . import discriminator_train
            return discriminator_train.discriminator_forward_train(self, images, passes=_precision_passes(kwargs),
                                                                   masks=kwargs.get("hg_record_masks"))
        passes = _precision_passes(kwargs)
        return discriminator_ops.discriminator_forward(self, images, passes=passes)
