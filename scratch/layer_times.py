"""Per-launch device times of one fast-mode training step (CUDA events), grouped by op shape."""
import sys, torch
sys.path.insert(0, '.')
import artist_style_transfer_b200 as ast
from artist_style_transfer_b200 import ops
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
S = int(sys.argv[2]) if len(sys.argv) > 2 else 256
dev = torch.device('cuda')
torch.manual_seed(2)
net = ast.StyleTransfer(device=dev, precision='fast'); vgg = ast.VGG16(vgg_path=None, precision='fast').to(dev)
style = ast.style_grams_single(vgg, torch.randint(0, 256, (3, S, S), device=dev).float(), B)
tr = ast.PerceptualTrainer(net, vgg, style)
x = torch.randint(0, 256, (B, 3, S, S), device=dev).float()
for _ in range(3): tr.step(x)
ops.PROFILE_DETAIL = True
ops.profile_begin()
for _ in range(3): tr.step(x)
prof = ops.profile_end()
tot = sum(v[0] for v in prof.values())
print(f"total {tot/3:.2f} ms/step over instrumented launches")
for k, (ms, cnt) in sorted(prof.items(), key=lambda kv: -kv[1][0]):
    print(f"{ms/3:8.3f} ms {cnt/3:5.1f}x  {k}")
