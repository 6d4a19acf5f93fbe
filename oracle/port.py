"""Plain-PyTorch CPU restatement of the reference hot path (TEST INFRASTRUCTURE ONLY).

Every function cites the reference lines it follows.  It is written functionally
(explicit parameter dicts + torch.nn.functional) so that it shares no structure
with the product modules; parity with the real reference classes is pinned by
`tests/test_oracle_golden.py` against `tests/golden/*.npz` (made by
`oracle/make_golden.py`, which imports `/root/reference`).

Nothing under `artist_style_transfer_b200/` may import this module.
"""
import torch
import torch.nn.functional as F

from .weights import IMAGENET_NEG_MEAN, TRANSFER_LAYERS, VGG_CONVS, VGG_POOLS, VGG_TAPS

CONTENT_WEIGHT = 17.0  # train_cnn.py:40
STYLE_WEIGHT = 25.0    # train_cnn.py:41


# ----------------------------------------------------------------------------- transform net
def _conv_layer(x, sd, prefix, k, stride, norm=True):
    """ConvLayer.forward, cnn.py:72-79: reflect-pad k//2 (k>1) -> Conv2d(no pad) -> InstanceNorm(affine)."""
    if k > 1:
        x = F.pad(x, (k // 2,) * 4, mode="reflect")                      # cnn.py:55-58
    x = F.conv2d(x, sd[prefix + ".conv_layer.weight"], sd[prefix + ".conv_layer.bias"], stride=stride)
    if norm:                                                             # cnn.py:66-70,75-78
        x = F.instance_norm(x, weight=sd[prefix + ".norm_layer.weight"],
                            bias=sd[prefix + ".norm_layer.bias"], eps=1e-5)
    return x


def _deconv_layer(x, sd, prefix, k, stride, output_padding):
    """DeconvLayer.forward, cnn.py:118-124: ConvTranspose2d(pad k//2, output_padding) -> InstanceNorm."""
    x = F.conv_transpose2d(x, sd[prefix + ".conv_transpose.weight"], sd[prefix + ".conv_transpose.bias"],
                           stride=stride, padding=k // 2, output_padding=output_padding)
    return F.instance_norm(x, weight=sd[prefix + ".norm_layer.weight"],
                           bias=sd[prefix + ".norm_layer.bias"], eps=1e-5)


def transfer_forward(x, sd):
    """StyleTransfer.forward, cnn.py:45-49 over the layer list cnn.py:15-40."""
    x = F.relu(_conv_layer(x, sd, "ConvBlock.0", 9, 1))
    x = F.relu(_conv_layer(x, sd, "ConvBlock.2", 3, 2))
    x = F.relu(_conv_layer(x, sd, "ConvBlock.4", 3, 2))
    x = F.relu(_conv_layer(x, sd, "ConvBlock.6", 1, 1))
    for i in range(5):                                                   # ResidualLayer.forward cnn.py:94-99
        y = F.relu(_conv_layer(x, sd, f"ResidualBlock.{i}.conv1", 3, 1))
        x = _conv_layer(y, sd, f"ResidualBlock.{i}.conv2", 3, 1) + x
    x = F.relu(_deconv_layer(x, sd, "DeconvBlock.0", 1, 1, 0))
    x = F.relu(_deconv_layer(x, sd, "DeconvBlock.2", 3, 2, 1))
    x = F.relu(_deconv_layer(x, sd, "DeconvBlock.4", 3, 2, 1))
    return _conv_layer(x, sd, "DeconvBlock.6", 9, 1, norm=False)         # cnn.py:39


# ----------------------------------------------------------------------------- VGG16 taps
def vgg_features(x, sd, upto=22):
    """VGG16.forward, train_cnn.py:63-78: torchvision vgg16.features, taps at 3/8/15/22, stop at 22."""
    convs = {idx: (cin, cout) for idx, cin, cout in VGG_CONVS}
    feats = {}
    for idx in range(upto + 1):
        if idx in convs:
            x = F.conv2d(x, sd[f"features.{idx}.weight"], sd[f"features.{idx}.bias"], padding=1)
        elif idx in VGG_POOLS:
            x = F.max_pool2d(x, 2, 2)
        else:
            x = F.relu(x)
        if idx in VGG_TAPS:
            feats[VGG_TAPS[idx]] = x
    return feats


def vgg_content_only(x, sd):
    """VGG16(just_content=True).forward, train_cnn.py:64-68: the tensor at features idx 8."""
    return vgg_features(x, sd, upto=8)["relu2_2"]


def gram(f):
    """gram(), train_cnn.py:103-107."""
    b, c, h, w = f.shape
    m = f.reshape(b, c, h * w)
    return torch.bmm(m, m.transpose(1, 2)) / (c * h * w)


def neg_mean(dtype=torch.float32, device=None):
    return torch.tensor(IMAGENET_NEG_MEAN, dtype=torch.float32, device=device).reshape(1, 3, 1, 1).to(dtype)


# ----------------------------------------------------------------------------- style-gram setup
def style_grams_single(style_img, vgg_sd, batch):
    """'random'/'average' setup, train_cnn.py:184-190,199-204: one image expanded to the batch."""
    st = style_img + neg_mean(style_img.dtype, style_img.device)                            # (3,H,W)+(1,3,1,1) -> (1,3,H,W)
    feats = vgg_features(st.expand(batch, -1, -1, -1), vgg_sd)
    return {k: gram(v) for k, v in feats.items()}


def style_grams_smartaverage(paintings, vgg_sd, batch, mode="reference"):
    """'smartaverage', train_cnn.py:224-244.

    mode='reference': sum VGG features over paintings, divide by count, ONE Gram of the mean
    feature (what the reference does).  mode='mean_gram': mean of per-painting Grams
    (the north-star wording; not pinned by the reference).
    """
    acc = None
    for p in paintings:
        st = p + neg_mean(p.dtype, p.device)
        feats = vgg_features(st.expand(batch, -1, -1, -1), vgg_sd)
        cur = feats if mode == "reference" else {k: gram(v) for k, v in feats.items()}
        if acc is None:
            acc = {k: v.clone() for k, v in cur.items()}
        else:
            for k in acc:
                acc[k] += cur[k]
    n = len(paintings)
    if mode == "reference":
        return {k: gram(v / n) for k, v in acc.items()}
    return {k: v / n for k, v in acc.items()}


# ----------------------------------------------------------------------------- the training step
def perceptual_losses(generated, content, vgg_sd, style_gram,
                      content_weight=CONTENT_WEIGHT, style_weight=STYLE_WEIGHT):
    """train_cnn.py:300-329 (method != 3): returns content, style, total and the generated Grams."""
    nm = neg_mean(generated.dtype, generated.device)
    cf = vgg_features(content + nm, vgg_sd)                               # :300
    gf = vgg_features(generated + nm, vgg_sd)                             # :301
    content_loss = F.mse_loss(gf["relu2_2"], cf["relu2_2"]) * content_weight  # :307-308
    grams = {k: gram(v) for k, v in gf.items()}
    style_loss = 0
    for k in gf:                                                          # :321-325
        style_loss = style_loss + F.mse_loss(grams[k], style_gram[k])
    style_loss = style_loss * style_weight
    return content_loss, style_loss, content_loss + style_loss, grams


def training_step(tsd, vgg_sd, content, style_gram, content_weight=CONTENT_WEIGHT,
                  style_weight=STYLE_WEIGHT, want_grads=True):
    """One pass of train_cnn.py:295-333 (no optimizer step).  `tsd` values must be leaf tensors
    with requires_grad when `want_grads`.  Returns dict(losses, grams, generated, grads)."""
    generated = transfer_forward(content, tsd)                            # :299
    c, s, t, grams = perceptual_losses(generated, content, vgg_sd, style_gram, content_weight, style_weight)
    out = {"content": c.detach(), "style": s.detach(), "total": t.detach(),
           "grams": {k: v.detach() for k, v in grams.items()}, "generated": generated.detach()}
    if want_grads:
        keys = list(tsd.keys())
        grads = torch.autograd.grad(t, [tsd[k] for k in keys], allow_unused=True)  # :333
        out["grads"] = {k: (g if g is not None else torch.zeros_like(tsd[k])) for k, g in zip(keys, grads)}
    return out


def make_leaf(sd, dtype=torch.float32, requires_grad=True):
    return {k: v.detach().to(dtype).clone().requires_grad_(requires_grad) for k, v in sd.items()}


def adam_l2_step(params, grads, state, lr, step, wd=1e-4, b1=0.9, b2=0.999, eps=1e-8):
    """optim.Adam(lr, weight_decay=1e-4) update, train_cnn.py:247,334 (L2 folded into the grad)."""
    for k in params:
        g = grads[k] + wd * params[k]
        m, v = state.setdefault(k, (torch.zeros_like(g), torch.zeros_like(g)))
        m.mul_(b1).add_(g, alpha=1 - b1)
        v.mul_(b2).addcmul_(g, g, value=1 - b2)
        mhat = m / (1 - b1 ** step)
        vhat = v / (1 - b2 ** step)
        params[k] = params[k] - lr * mhat / (vhat.sqrt() + eps)
    return params


def transfer_param_keys():
    keys = []
    for prefix, kind, *_ in TRANSFER_LAYERS:
        w = prefix + (".conv_transpose" if kind == "deconv" else ".conv_layer")
        keys += [w + ".weight", w + ".bias"]
        if kind != "conv_nonorm":
            keys += [prefix + ".norm_layer.weight", prefix + ".norm_layer.bias"]
    return keys
