"""Representative conv launches of the B=32 256^2 step, timed back to back (CUDA events, rotating buffers > L2).
usage: prof_convs.py [tag]   (AST_CONV_HX=0 disables the halo kernel for an A/B)"""
import sys, torch
sys.path.insert(0, '.')
from artist_style_transfer_b200 import _lib, ops, conv_geometry as cg
torch.manual_seed(0)
n = 32
def run(name, dtype, cin, cout, hin, launches, hout, mask=False, stats=False, reps=10, sets=3):
    bufs = []
    nt = sum(len(l.taps) for l in launches)
    for _ in range(sets):
        x = torch.randn(n, hin, hin, cin, device='cuda').to(dtype)
        y = torch.empty(n, hout, hout, cout, device='cuda', dtype=dtype)
        m = torch.randn(n, hout, hout, cout, device='cuda') if mask else None
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
    print(f"{name:34s} {fam} {us:8.1f} us  {gf/us*1e3/1e3:7.1f} TF/s")
bf, f32 = torch.bfloat16, torch.float32
run("res fwd 128->128 66->64 bf16+stats", bf, 128, 128, 66, cg.conv_fwd(3, 1, 0, 66, 66), 64, stats=True)
run("res dgrad 128->128 64->66 bf16", bf, 128, 128, 64, cg.conv_dgrad(3, 1, 0, 66, 66), 66)
run("vgg conv2_2 128->128 128^2 tf32", f32, 128, 128, 128, cg.conv_fwd(3, 1, 1, 128, 128), 128)
run("vgg conv2_1 64->128 128^2 tf32", f32, 64, 128, 128, cg.conv_fwd(3, 1, 1, 128, 128), 128)
run("vgg conv3_2 256->256 64^2 tf32", f32, 256, 256, 64, cg.conv_fwd(3, 1, 1, 64, 64), 64)
run("vgg conv4_2 512->512 32^2 tf32", f32, 512, 512, 32, cg.conv_fwd(3, 1, 1, 32, 32), 32)
run("vgg dgrad conv4 512->512 bf16+mask", bf, 512, 512, 32, cg.conv_dgrad(3, 1, 1, 32, 32), 32, mask=True)
run("vgg dgrad conv3 256->256 bf16+mask", bf, 256, 256, 64, cg.conv_dgrad(3, 1, 1, 64, 64), 64, mask=True)
run("vgg dgrad conv2_2 128->128 bf16+mask", bf, 128, 128, 128, cg.conv_dgrad(3, 1, 1, 128, 128), 128, mask=True)
# alignment experiment: nine taps that all read the SAME pixel (no shifted descriptors, 8-pixel patch rows = aligned atoms)
z9 = [cg.Launch(64, 64, 1, 1, 0, 0, [(0, 0)] * 9, [(u, v) for u in range(3) for v in range(3)], 0)]
run("9 zero-shift taps 128->128 64^2 bf16", bf, 128, 128, 64, z9, 64)
z9b = [cg.Launch(128, 128, 1, 1, 0, 0, [(0, 0)] * 9, [(u, v) for u in range(3) for v in range(3)], 0)]
run("9 zero-shift taps 128->128 128^2 tf32", f32, 128, 128, 128, z9b, 128)
run("res fwd-like 128->128 64^2 pad1 bf16", bf, 128, 128, 64, cg.conv_fwd(3, 1, 1, 64, 64), 64)
