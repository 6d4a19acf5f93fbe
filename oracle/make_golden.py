"""Generate tests/golden/*.npz by running the UNMODIFIED reference classes on CPU.

Run in the build container only (`python -m oracle.make_golden`); `/root/reference`
does not exist on the GPU box, so the outputs are committed as small fixtures.
Recipe: stub matplotlib, import `cnn` / `train_cnn` from
/root/reference, patch `torch.load` so `VGG16()` builds without the absent weight
file, load the seeded synthetic weights from `oracle.weights`, then execute the
lines of `train_cnn.py:295-333` / `:184-190` / `:224-244` around those classes.
"""
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def import_reference():
    for name in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.path.insert(0, REF)
    import cnn  # noqa: E402
    import train_cnn  # noqa: E402
    return cnn, train_cnn


def build_reference_nets(cnn, train_cnn, tsd, vsd, dtype):
    cpu = torch.device("cpu")
    transfer = cnn.StyleTransfer(device=cpu)                 # .double() inside, cnn.py:43
    transfer.load_state_dict({k: v.double() for k, v in tsd.items()}, strict=True)
    real_load = torch.load
    torch.load = lambda *a, **k: {}                           # weight file absent
    try:
        vgg = train_cnn.VGG16()
    finally:
        torch.load = real_load
    missing, unexpected = vgg.load_state_dict({k: v.double() for k, v in vsd.items()}, strict=False)
    assert not unexpected and not missing, (missing, unexpected)
    for p in vgg.parameters():
        p.requires_grad = False
    return transfer.to(dtype), vgg.to(dtype)


def summarize_gram(g):
    g = g.detach().double()
    return {"block": g[:, :32, :32].numpy(), "diag": torch.diagonal(g, dim1=1, dim2=2).numpy(),
            "fro": g.flatten(1).norm(dim=1).numpy(), "sum": g.flatten(1).sum(dim=1).numpy()}


def run_forward(cnn, train_cnn, batch, height, width, dtype, seed=2):
    """StyleTransfer.forward (cnn.py:45-49) alone at an arbitrary H x W (BASELINE config 4: 1080 x 1920 stylisation,
    inference.py:115)."""
    from oracle import weights
    tsd = weights.transfer_state_dict(seed)
    transfer, _ = build_reference_nets(cnn, train_cnn, tsd, weights.vgg_state_dict(seed), dtype)
    content = weights.content_batch(batch, height, seed, width=width).to(dtype)
    with torch.no_grad():
        gen = transfer(content).double()
    sy, sx = max(1, height // 32), max(1, width // 32)
    return {"generated_sub": gen[:, :, ::sy, ::sx].numpy(), "generated_norm": np.array(float(gen.norm())),
            "generated_rowsum": gen.sum(dim=3).numpy(), "generated_colsum": gen.sum(dim=2).numpy()}


def run_step(cnn, train_cnn, batch, size, dtype, seed=2, with_grads=True):
    """Body of train_cnn.py:295-333 with method 0 ('random') style setup :184-190."""
    from oracle import weights
    tsd = weights.transfer_state_dict(seed)
    vsd = weights.vgg_state_dict(seed)
    transfer, vgg = build_reference_nets(cnn, train_cnn, tsd, vsd, dtype)
    neg_mean = torch.tensor([-103.939, -116.779, -123.68], dtype=torch.float32).reshape(1, 3, 1, 1)
    content = weights.content_batch(batch, size, seed).to(dtype)
    style_tensor = weights.style_image(size, seed).to(dtype).add(neg_mean)        # :184-185
    b, c, h, w = style_tensor.shape
    style_gram = {k: train_cnn.gram(v) for k, v in vgg(style_tensor.expand([batch, c, h, w])).items()}
    mse = torch.nn.MSELoss()
    transfer.zero_grad()
    generated = transfer(content)                                                # :299
    cf = vgg(content.add(neg_mean))                                              # :300
    gf = vgg(generated.add(neg_mean))                                            # :301
    content_loss = mse(gf["relu2_2"], cf["relu2_2"]) * 17                        # :307-308
    style_loss = 0
    grams = {}
    for key, value in gf.items():                                                # :321-325
        grams[key] = train_cnn.gram(value)
        style_loss = style_loss + mse(grams[key], style_gram[key])
    style_loss = style_loss * 25
    total = content_loss + style_loss                                            # :329
    out = {"content": float(content_loss), "style": float(style_loss), "total": float(total)}
    arrays = {"losses": np.array([out["content"], out["style"], out["total"]], dtype=np.float64)}
    if with_grads:
        total.backward()                                                         # :333
        for name, p in transfer.named_parameters():
            gflat = p.grad.detach().double().flatten()
            arrays["grad_norm/" + name] = np.array(float(gflat.norm()))
            arrays["grad_head/" + name] = gflat[:16].numpy()
    gen = generated.detach().double()
    arrays["generated_sub"] = gen[:, :, ::max(1, size // 32), ::max(1, size // 32)].numpy()
    arrays["generated_norm"] = np.array(float(gen.norm()))
    for k, g in grams.items():
        for n, v in summarize_gram(g).items():
            arrays[f"gram/{k}/{n}"] = v
    for k, g in style_gram.items():
        for n, v in summarize_gram(g).items():
            arrays[f"style_gram/{k}/{n}"] = v
    for k, v in gf.items():
        arrays[f"feat_norm/{k}"] = np.array(float(v.detach().double().norm()))
    return arrays


def run_smartaverage(cnn, train_cnn, batch, size, count, dtype, seed=2):
    """train_cnn.py:224-244 on `count` synthetic paintings."""
    from oracle import weights
    vsd = weights.vgg_state_dict(seed)
    _, vgg = build_reference_nets(cnn, train_cnn, weights.transfer_state_dict(seed), vsd, dtype)
    neg_mean = torch.tensor([-103.939, -116.779, -123.68], dtype=torch.float32).reshape(1, 3, 1, 1)
    paintings = [weights.style_image(size, seed, i).to(dtype) for i in range(count)]
    style_gram = {}
    first = paintings[0].add(neg_mean)
    b, c, h, w = first.shape
    for key, value in vgg(first.expand([batch, c, h, w])).items():               # :230-233
        style_gram[key] = value
    for i in range(1, count):                                                    # :234-239
        st = paintings[i].add(neg_mean)
        for key, value in vgg(st.expand([batch, c, h, w])).items():
            style_gram[key] += value
    arrays = {}
    for key, value in style_gram.items():                                        # :242-243
        g = train_cnn.gram(value / count)
        for n, v in summarize_gram(g).items():
            arrays[f"gram/{key}/{n}"] = v
    return arrays


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    cnn, train_cnn = import_reference()
    jobs = [
        ("step_b2_s32_f64", lambda: run_step(cnn, train_cnn, 2, 32, torch.float64)),
        ("step_b2_s64_f64", lambda: run_step(cnn, train_cnn, 2, 64, torch.float64)),
        ("step_b4_s256_f32", lambda: run_step(cnn, train_cnn, 4, 256, torch.float32)),
        ("step_b4_s256_f64", lambda: run_step(cnn, train_cnn, 4, 256, torch.float64)),
        ("smartavg_b2_s64_n5_f64", lambda: run_smartaverage(cnn, train_cnn, 2, 64, 5, torch.float64)),
        # BASELINE configs[1] at the bench's own batch (B=32/GPU), configs[2] smartaverage at 512^2, configs[3] 1080p
        # stylisation, configs[4] 1024^2 step (relu1_2 Gram with K = 1,048,576): fp32 = north_star's "PyTorch fp32 path"
        ("step_b32_s256_f32", lambda: run_step(cnn, train_cnn, 32, 256, torch.float32, with_grads=False)),
        ("smartavg_b1_s512_n8_f32", lambda: run_smartaverage(cnn, train_cnn, 1, 512, 8, torch.float32)),
        ("fwd_b1_1080x1920_f64", lambda: run_forward(cnn, train_cnn, 1, 1080, 1920, torch.float64)),
        ("step_b1_s1024_f32", lambda: run_step(cnn, train_cnn, 1, 1024, torch.float32, with_grads=False)),
    ]
    only = set(sys.argv[1:])
    for name, fn in jobs:
        if only and name not in only:
            continue
        arrays = fn()
        path = os.path.join(OUT, name + ".npz")
        np.savez_compressed(path, **arrays)
        print(name, {k: float(v) for k, v in zip(("content", "style", "total"), arrays.get("losses", []))},
              os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
