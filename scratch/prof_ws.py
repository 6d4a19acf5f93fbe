"""conv_ws representative launches of the B=32 256^2 step (small-channel layers)."""
import sys, torch
sys.path.insert(0, '.')
from artist_style_transfer_b200 import _lib, ops, conv_geometry as cg
torch.manual_seed(0)
n = 32
def run(name, dtype, cin, cout, hin, win, launches, hout, wout, stats=False, mask=False, reps=10, sets=3):
    bufs = []
    nt = sum(len(l.taps) for l in launches)
    for _ in range(sets):
        x = torch.randn(n, hin, win, cin, device='cuda').to(dtype)
        y = torch.empty(n, hout, wout, cout, device='cuda', dtype=dtype)
        m = torch.randn(n, hout, wout, cout, device='cuda') if mask else None
        bufs.append((x, y, m))
    wp = (torch.randn(nt, cout, cin, device='cuda') / (cin * nt) ** 0.5).to(dtype)
    sums = torch.zeros(2 * n * cout, dtype=torch.float64, device='cuda') if stats else None
    def go(i):
        x, y, m = bufs[i % sets]
        ops.conv_gather(x, wp, launches, y, mask=m, tensor=True, stats=sums)
    before = _lib.family_stats()
    for i in range(3): go(i)
    fam = [k for k, v in _lib.family_delta(before).items() if v[0]]
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps): go(i)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    gf = 2.0 * n * sum(l.mi * l.mj * len(l.taps) for l in launches) * cin * cout / 1e9
    mb = (x.numel() * x.element_size() + y.numel() * y.element_size()) / 1e6
    print(f"{name:44s} {fam} {us:8.1f} us  {gf/us*1e-3:6.1f} TF/s  {mb/us*1e-3:6.2f} TB/s(in+out)")
bf, f32 = torch.bfloat16, torch.float32
vt9 = [cg.Launch(256, 256, 1, 1, 0, 0, [(d, 0) for d in range(9)], [(d, 0) for d in range(9)], 0)]
run("T first layer vt9 32->32 256^2 +stats", bf, 32, 32, 264, 256, vt9, 256, 256, stats=True)
run("T D4 convT 3x3 s2 64->32 128^2->256^2 +stats", bf, 64, 32, 128, 128, cg.convT_fwd(3, 2, 1, 1, 128, 128), 256, 256, stats=True)
run("T D2 convT 3x3 s2 128->64 64^2->128^2 +stats", bf, 128, 64, 64, 64, cg.convT_fwd(3, 2, 1, 1, 64, 64), 128, 128, stats=True)
run("T 1x1 128->128 64^2 +stats", bf, 128, 128, 64, 64, cg.conv_fwd(1, 1, 0, 64, 64), 64, 64, stats=True)
run("VGG conv1_2 64->64 256^2 tf32", f32, 64, 64, 256, 256, cg.conv_fwd(3, 1, 1, 256, 256), 256, 256)
run("VGG conv1_1 vt3 16->64 256^2 tf32", f32, 16, 64, 256, 256, [cg.Launch(256, 256, 1, 1, 0, 0, [(-1, 0), (0, 0), (1, 0)], [(0, 0), (1, 0), (2, 0)], 0)], 256, 256)
run("VGG dgrad conv1_2 64->64 256^2 bf16 +mask", bf, 64, 64, 256, 256, cg.conv_dgrad(3, 1, 1, 256, 256), 256, 256, mask=True)
# thin filter gradients (first / last 9x9 layers after the row fold)
def wg(name, xs, gs, launches, dshape, s_co, s_ci, s_u, s_v, reps=10):
    x = torch.randn(*xs, device='cuda').to(bf); g = torch.randn(*gs, device='cuda').to(bf)
    dw = torch.zeros(*dshape, device='cuda')
    for _ in range(3): ops.wgrad_gather(x, g, launches, dw, s_co, s_ci, s_u, s_v, tensor=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): ops.wgrad_gather(x, g, launches, dw, s_co, s_ci, s_u, s_v, tensor=True)
    e1.record(); torch.cuda.synchronize()
    print(f"{name:44s} {e0.elapsed_time(e1)/reps*1e3:8.1f} us")
wg("wgrad first layer x(264,256,32) g(256,256,32)", (n, 264, 256, 32), (n, 256, 256, 32), vt9, (9, 32, 32), 32, 1, 32 * 32, 0)
lw = [cg.Launch(256, 264, 1, 1, 0, 0, [(d, 0) for d in range(9)], [(d, 0) for d in range(9)], 0)]
wg("wgrad last layer x(264,264,32) g(256,264,32)", (n, 264, 264, 32), (n, 256, 264, 32), lw, (9, 32, 32), 32, 1, 32 * 32, 0)
