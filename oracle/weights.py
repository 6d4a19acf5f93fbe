"""Deterministic synthetic weights and inputs shared by the oracle, the golden
generator and the tests (TEST INFRASTRUCTURE ONLY - see oracle/__init__.py).

The reference's checkpoint files are absent, so weights are random:
shapes/keys follow the reference `state_dict` layout (`cnn.py:15-40`,
`train_cnn.py:54-56`), values come from an explicit `torch.Generator` stream so
that the reference classes (in the build container), the oracle port and the
CUDA modules all load *identical* tensors without shipping 40 MB of fixtures.
"""
import math

import torch

# (key prefix, kind, cin, cout, k)  -- order is the reference module order, cnn.py:15-40
TRANSFER_LAYERS = (
    [("ConvBlock.0", "conv", 3, 32, 9), ("ConvBlock.2", "conv", 32, 64, 3),
     ("ConvBlock.4", "conv", 64, 128, 3), ("ConvBlock.6", "conv", 128, 128, 1)]
    + [(f"ResidualBlock.{i}.{c}", "conv", 128, 128, 3) for i in range(5) for c in ("conv1", "conv2")]
    + [("DeconvBlock.0", "deconv", 128, 128, 1), ("DeconvBlock.2", "deconv", 128, 64, 3),
       ("DeconvBlock.4", "deconv", 64, 32, 3), ("DeconvBlock.6", "conv_nonorm", 32, 3, 9)]
)

# torchvision vgg16().features indices of the 13 convs; the path only runs 0..21 (train_cnn.py:70-77)
VGG_CONVS = [(0, 3, 64), (2, 64, 64), (5, 64, 128), (7, 128, 128), (10, 128, 256), (12, 256, 256),
             (14, 256, 256), (17, 256, 512), (19, 512, 512), (21, 512, 512),
             (24, 512, 512), (26, 512, 512), (28, 512, 512)]
VGG_TAPS = {3: "relu1_2", 8: "relu2_2", 15: "relu3_3", 22: "relu4_3"}
VGG_POOLS = (4, 9, 16, 23, 30)

IMAGENET_NEG_MEAN = (-103.939, -116.779, -123.68)  # BGR, train_cnn.py:164


def transfer_state_dict(seed=2, perturb_affine=True, dtype=torch.float32):
    """70 tensors keyed like the reference `StyleTransfer.state_dict()`."""
    g = torch.Generator().manual_seed(1000 + seed)
    sd = {}
    for prefix, kind, cin, cout, k in TRANSFER_LAYERS:
        bound = 1.0 / math.sqrt(cin * k * k)
        if kind == "deconv":
            wname, shape = prefix + ".conv_transpose", (cin, cout, k, k)  # ConvTranspose2d layout
        else:
            wname, shape = prefix + ".conv_layer", (cout, cin, k, k)
        sd[wname + ".weight"] = (torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1) * bound
        sd[wname + ".bias"] = (torch.rand(cout, generator=g, dtype=torch.float64) * 2 - 1) * bound
        if kind != "conv_nonorm":
            gam = torch.ones(cout, dtype=torch.float64)
            bet = torch.zeros(cout, dtype=torch.float64)
            if perturb_affine:
                gam = gam + 0.2 * torch.randn(cout, generator=g, dtype=torch.float64)
                bet = bet + 0.2 * torch.randn(cout, generator=g, dtype=torch.float64)
            sd[prefix + ".norm_layer.weight"] = gam
            sd[prefix + ".norm_layer.bias"] = bet
    return {k: v.to(dtype) for k, v in sd.items()}


def vgg_state_dict(seed=2, with_bias=True, dtype=torch.float32, all_layers=True):
    """`features.N.{weight,bias}` like torchvision vgg16 (kaiming-normal fan_out scale)."""
    g = torch.Generator().manual_seed(2000 + seed)
    sd = {}
    for idx, cin, cout in VGG_CONVS:
        if not all_layers and idx > 21:
            break
        std = math.sqrt(2.0 / (cout * 9))
        sd[f"features.{idx}.weight"] = torch.randn((cout, cin, 3, 3), generator=g, dtype=torch.float64) * std
        b = torch.randn(cout, generator=g, dtype=torch.float64) * (0.5 if with_bias else 0.0)
        sd[f"features.{idx}.bias"] = b
    return {k: v.to(dtype) for k, v in sd.items()}


def content_batch(batch, size, seed=2, step=0, rank=0, dtype=torch.float32, width=None):
    """uint8-valued BGR images as floats, like dataset.py:97-108 feeds train_cnn.py:298."""
    g = torch.Generator().manual_seed(3000 + seed + 1000 * step + rank)
    w = size if width is None else width
    return torch.randint(0, 256, (batch, 3, size, w), generator=g).to(dtype)


def style_image(size, seed=2, index=0, dtype=torch.float32):
    """One (3,H,W) painting tensor (train_cnn.py:184, dataset.py:120-229)."""
    g = torch.Generator().manual_seed(4000 + seed + index)
    return torch.randint(0, 256, (3, size, size), generator=g).to(dtype)
