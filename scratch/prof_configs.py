"""Per-kernel device time (CUPTI) of BASELINE configs[3] (1080p stylisation, B=8) and configs[4] (1024^2 step, B=4)."""
import collections, sys, torch
sys.path.insert(0, '.')
import artist_style_transfer_b200 as ast
from torch.profiler import profile, ProfilerActivity
dev = torch.device('cuda')
torch.manual_seed(2)
net = ast.StyleTransfer(device=dev, precision='fast'); vgg = ast.VGG16(vgg_path=None, precision='fast').to(dev)
def prof(fn, tag, reps=2):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    wall = e0.elapsed_time(e1) / reps
    with profile(activities=[ProfilerActivity.CUDA]) as p:
        for _ in range(reps): fn()
        torch.cuda.synchronize()
    agg = collections.defaultdict(lambda: [0, 0.0])
    for ev in p.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA:
            agg[ev.name][0] += 1; agg[ev.name][1] += ev.device_time
    tot = sum(v[1] for v in agg.values()) / 1e3 / reps
    print(f"== {tag}: {wall:.2f} ms per call (events), kernels sum {tot:.2f} ms")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
        print(f"{v[1]/1e3/reps:8.3f} ms {v[0]/reps:5.1f}x  {k[:110]}")
which = sys.argv[1] if len(sys.argv) > 1 else "45"
if "4" in which:
    x = torch.randint(0, 256, (8, 1080, 1920, 3), device=dev, dtype=torch.uint8)
    prof(lambda: net.stylize(x), "config4 stylize uint8 B=8 1080p")
    xf = x.permute(0, 3, 1, 2).float()
    with torch.no_grad():
        prof(lambda: net(xf), "config4 forward fp32 in/out B=8 1080p")
    del x, xf
if "5" in which:
    style = ast.style_grams_single(vgg, torch.randint(0, 256, (3, 1024, 1024), device=dev).float(), 4)
    tr = ast.PerceptualTrainer(net, vgg, style)
    xb = torch.randint(0, 256, (4, 3, 1024, 1024), device=dev, dtype=torch.uint8)
    prof(lambda: tr.step(xb), "config5 step B=4 1024^2")
