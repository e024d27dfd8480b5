import torch


def randn_tensor(shape, generator=None, device=None, dtype=None, layout=None):
    """diffusers.utils.torch_utils.randn_tensor: draw on the generator's device, then move."""
    device = device or torch.device("cpu")
    gen_device = generator.device if generator is not None else device
    t = torch.randn(shape, generator=generator, device=gen_device, dtype=dtype)
    return t.to(device)
