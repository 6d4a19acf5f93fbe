"""Per-kernel device time of the CUDA-graph replay of one training step (CUPTI through torch.profiler).

Unlike the CUDA-event timing around eager calls (scratch/layer_times.py) these durations contain no host launch gaps:
the sum over kernels is <= the replayed step time.  usage: graph_profile.py [B] [S] [replays] [out.json]"""
import collections
import json
import sys

import torch

sys.path.insert(0, '.')
import artist_style_transfer_b200 as ast  # noqa: E402
from torch.profiler import profile, ProfilerActivity  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
S = int(sys.argv[2]) if len(sys.argv) > 2 else 256
R = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dev = torch.device('cuda')
torch.manual_seed(2)
net = ast.StyleTransfer(device=dev, precision='fast')
vgg = ast.VGG16(vgg_path=None, precision='fast').to(dev)
style = ast.style_grams_single(vgg, torch.randint(0, 256, (3, S, S), device=dev).float(), B)
tr = ast.PerceptualTrainer(net, vgg, style, cuda_graph=True)
x = torch.randint(0, 256, (B, 3, S, S), device=dev).float()
for _ in range(8):
    tr.step(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    tr.step(x)
e1.record()
torch.cuda.synchronize()
step_ms = e0.elapsed_time(e1) / 10
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(R):
        tr.step(x)
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        t = ev.device_time if hasattr(ev, 'device_time') else ev.cuda_time
        agg[ev.name][0] += 1
        agg[ev.name][1] += t
tot = sum(v[1] for v in agg.values()) / 1e3 / R
print(f"replayed step {step_ms:.3f} ms; sum of kernel durations {tot:.3f} ms/step ({R} replays)")
rows = []
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    rows.append({"kernel": k, "ms_per_step": v[1] / 1e3 / R, "launches_per_step": v[0] / R})
    print(f"{v[1]/1e3/R:8.3f} ms {v[0]/R:6.1f}x  {k[:140]}")
if len(sys.argv) > 4:
    json.dump({"step_ms": step_ms, "sum_kernel_ms": tot, "kernels": rows}, open(sys.argv[4], "w"), indent=1)
