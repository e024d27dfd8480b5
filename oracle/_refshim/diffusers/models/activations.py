from torch import nn


def get_activation(name):
    name = name.lower()
    if name in ("silu", "swish"):
        return nn.SiLU()
    if name == "relu":
        return nn.ReLU()
    if name == "gelu":
        return nn.GELU()
    if name == "mish":
        return nn.Mish()
    raise ValueError(name)
