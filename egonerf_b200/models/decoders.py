"""Parameter containers for the colour decoders (reference: models/tensorBase.py:30-129).  The modules keep the
reference's submodule names (`mlp.0/2/4`) so state_dict keys match; the arithmetic of the decode runs inside
libegn_b200 as part of `EgoNeRF.forward`."""
import torch


def _mlp(in_c, feature_c):
    l1, l2, l3 = torch.nn.Linear(in_c, feature_c), torch.nn.Linear(feature_c, feature_c), torch.nn.Linear(feature_c, 3)
    torch.nn.init.constant_(l3.bias, 0)
    return torch.nn.Sequential(l1, torch.nn.ReLU(inplace=True), l2, torch.nn.ReLU(inplace=True), l3)


class MLPRender_Fea(torch.nn.Module):
    """tensorBase.py:54-78: input [features, viewdirs, PE(features), PE(viewdirs)]."""

    def __init__(self, inChannel, viewpe=6, feape=6, featureC=128):
        super().__init__()
        self.in_mlpC = 2 * viewpe * 3 + 2 * feape * inChannel + 3 + inChannel
        self.viewpe, self.feape = viewpe, feape
        self.mlp = _mlp(self.in_mlpC, featureC)


class MLPRender(torch.nn.Module):
    """tensorBase.py:107-129: input [features, viewdirs, PE(viewdirs)]."""

    def __init__(self, inChannel, viewpe=6, featureC=128):
        super().__init__()
        self.in_mlpC = (3 + 2 * viewpe * 3) + inChannel
        self.viewpe = viewpe
        self.mlp = _mlp(self.in_mlpC, featureC)


def SHRender(*_a, **_k):     # tensorBase.py:30-34 — marker; evaluated inside the compositing kernel
    raise RuntimeError("SHRender is evaluated inside libegn_b200; call EgoNeRF.forward")


def RGBRender(*_a, **_k):    # tensorBase.py:37-39
    raise RuntimeError("RGBRender is evaluated inside libegn_b200; call EgoNeRF.forward")
