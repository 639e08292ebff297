import torch.nn as nn


class SMPL(nn.Module):
    """Placeholder: the SMPL body model (licence-gated asset) is data preparation, not hot path."""

    def __init__(self, *a, **k):
        super().__init__()
        raise RuntimeError("smplx.SMPL stub: SMPL_NEUTRAL.pkl is not available in this environment")
