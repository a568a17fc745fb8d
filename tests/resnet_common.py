"""Shared by the ResNet18 tests and tests/golden/make_golden_resnet.py: seeded weights are REGENERATED (11.7 M parameters
are not stored): construct the network under torch.manual_seed(seed) -- the mirror and the reference draw the same
kaiming_normal_ values in the same order -- then give every BatchNorm non-trivial affine parameters and running statistics."""
import numpy as np
import torch


def randomize_bn(model, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.weight.copy_(torch.rand(m.weight.shape, generator=g) + 0.5)
                m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.1)
                m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=g) * 0.1)
                m.running_var.copy_(torch.rand(m.running_var.shape, generator=g) + 0.5)
    return model


def make_images(seed, b, size):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(b, 3, size, size, generator=g)


def np_state(model):
    return {k: v.detach().cpu().numpy() for k, v in model.state_dict().items()}


CASES = [  # name, weight seed, bn seed, image seed, batch, image size
    ("resnet18_64", 21, 22, 23, 2, 64),
    ("resnet18_224", 31, 32, 33, 1, 224),
    ("resnet18_96_b3", 41, 42, 43, 3, 96),
]
