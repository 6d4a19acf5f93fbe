"""Same-box PyTorch-eager comparator (BASELINE.md section 4 / VERDICT r01 item 9): the reference's training step
(train_cnn.py:295-334) written with stock torch.nn.functional ops - the oracle port, which is pinned to the reference
classes - run ON THE B200 with cuDNN / cuBLAS kernels, B=32 at 256^2, in three precisions:
  fp32-strict (TF32 off), TF32 (allow_tf32), bf16 autocast + channels_last.
This is context for the headline number (what the library kernels reach on the same silicon); it is NOT bench.py's
reference arm and nothing in the product imports it.  usage: python scratch/eager_comparator.py [B] [S]"""
import json, sys, time, torch
sys.path.insert(0, '.')
from oracle import port, weights

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
S = int(sys.argv[2]) if len(sys.argv) > 2 else 256
dev = torch.device('cuda')
out = {"batch": B, "size": S, "gpu": torch.cuda.get_device_name(0)}
vsd = {k: v.to(dev) for k, v in weights.vgg_state_dict(2).items()}
style_img = weights.style_image(S, 2).to(dev)


def run(tag, tf32, autocast):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    torch.backends.cudnn.allow_tf32 = tf32
    torch.backends.cudnn.benchmark = True
    params = {k: v.to(dev).requires_grad_(True) for k, v in weights.transfer_state_dict(2).items()}
    opt = torch.optim.Adam(list(params.values()), lr=0.0024, weight_decay=1e-4, fused=True)
    with torch.no_grad():
        style = port.style_grams_single(style_img, vsd, B)
    batches = [weights.content_batch(B, S, 2, step=i).to(dev) for i in range(4)]
    if autocast:
        batches = [b.contiguous(memory_format=torch.channels_last) for b in batches]

    def step(x):
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            gen = port.transfer_forward(x, params)
            c, s, t, _ = port.perceptual_losses(gen.float() if autocast else gen, x, vsd, style)
        t.backward()
        opt.step()
        return t
    for i in range(5):
        step(batches[i % 4])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    K = 10
    e0.record()
    for i in range(K):
        t = step(batches[i % 4])
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    out[tag] = {"ms_per_step": ms, "images_per_s": B / ms * 1e3, "total_loss_last": float(t),
                "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9}
    print(tag, out[tag], flush=True)


run("eager_fp32_strict", False, False)
run("eager_tf32", True, False)
run("eager_bf16_autocast_channels_last", True, True)
print(json.dumps(out))
