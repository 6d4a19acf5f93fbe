"""Which non-library kernels does one training step launch? (torch.profiler, eager step, B=32 256^2)"""
import sys, collections, torch
sys.path.insert(0, '.')
import artist_style_transfer_b200 as ast
from torch.profiler import profile, ProfilerActivity
dev = torch.device('cuda')
net = ast.StyleTransfer(device=dev, precision='fast'); vgg = ast.VGG16(vgg_path=None, precision='fast').to(dev)
style = ast.style_grams_single(vgg, torch.randint(0, 256, (3, 256, 256), device=dev).float(), 32)
tr = ast.PerceptualTrainer(net, vgg, style)
x = torch.randint(0, 256, (32, 3, 256, 256), device=dev).float()
for _ in range(3): tr.step(x)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    tr.step(x); torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        agg[ev.name][0] += 1; agg[ev.name][1] += ev.device_time if hasattr(ev, 'device_time') else ev.cuda_time
tot_lib = sum(v[1] for k, v in agg.items() if k.startswith('void ast::') or 'ast::' in k or 'fold_rows' in k)
tot_other = sum(v[1] for k, v in agg.items() if not ('ast::' in k or 'fold_rows' in k))
print(f"library kernels {tot_lib/1e3:.2f} ms, other {tot_other/1e3:.2f} ms")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    if 'ast::' in k or 'fold_rows' in k: continue
    print(f"{v[1]/1e3:8.3f} ms {v[0]:4d}x  {k[:150]}")
