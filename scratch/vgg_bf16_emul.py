"""CPU emulation: would a bf16-activation VGG (fp32 accumulate) meet the 2e-3 Gram tolerance?  (scratch, not product)"""
import sys, torch, torch.nn.functional as F
sys.path.insert(0, ".")
from oracle import port, weights
from oracle.port import VGG_CONVS, VGG_POOLS, VGG_TAPS
torch.manual_seed(0)
def rb(x): return x.to(torch.bfloat16).to(torch.float32)
def rt(x):  # tf32 rna
    i = x.view(torch.int32); i = (i + 0x1000) & ~0x1FFF; return i.view(torch.float32)
def vgg_q(x, sd, q, first_tf32=True):
    convs = {idx: (cin, cout) for idx, cin, cout in VGG_CONVS}; feats = {}
    for idx in range(23):
        if idx in convs:
            qq = rt if (idx == 0 and first_tf32) else q
            x = F.conv2d(qq(x), qq(sd[f"features.{idx}.weight"]), sd[f"features.{idx}.bias"], padding=1)
        elif idx in VGG_POOLS: x = F.max_pool2d(x, 2, 2)
        else: x = q(F.relu(x))
        if idx in VGG_TAPS: feats[VGG_TAPS[idx]] = x
    return feats
def rel(a, b): return float((a.double() - b.double()).norm() / b.double().norm())
vsd = weights.vgg_state_dict(); tsd = weights.transfer_state_dict()
for size in (64, 128):
    c = weights.content_batch(2, size)
    with torch.no_grad():
        y = port.transfer_forward(c, tsd)
        for name, inp in (("content", c), ("generated", y)):
            x = inp + port.neg_mean()
            ref = port.vgg_features(x, vsd)
            for qn, q in (("tf32", rt), ("bf16", rb)):
                f = vgg_q(x, vsd, q)
                print(size, name, qn, {k: f"{rel(port.gram(f[k]), port.gram(ref[k])):.2e}" for k in ref},
                      "feat2_2", f"{rel(f['relu2_2'], ref['relu2_2']):.2e}")
