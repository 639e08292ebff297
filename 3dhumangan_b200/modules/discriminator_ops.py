"""Host-side schedule of the U-Net discriminator on the sm_100a convolution kernels."""
import torch

from .. import abi


def discriminator_forward(module, images):
    abi.require_device()
    if not hasattr(abi.lib(), "hg_conv3x3"):
        raise RuntimeError("hg3d: the discriminator convolution kernels (csrc/dconv.cu) are not built into this library "
                           "yet; there is no cuDNN / eager fallback on this path")
    raise RuntimeError("hg3d: discriminator forward is not wired up in this build")
