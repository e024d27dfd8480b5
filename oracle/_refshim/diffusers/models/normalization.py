from torch import nn


class AdaGroupNorm(nn.Module):
    def __init__(self, *a, **k):
        raise NotImplementedError("not on the VAE path")


class RMSNorm(nn.Module):
    def __init__(self, *a, **k):
        raise NotImplementedError("not on the VAE path")
