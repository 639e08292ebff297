import torch
Tensor = torch.Tensor


class SMPLOutput(dict):
    pass
