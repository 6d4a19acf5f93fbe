"""Runs a few single conv layers (B=32) for instrumented builds (AST_B200_LIB=scratch/variants/libast_prof.so)."""
import sys, torch
sys.path.insert(0, '.')
from artist_style_transfer_b200 import ops, conv_geometry as cg
torch.manual_seed(0)
def run(dt, cin, cout, h, w, pad_in=0, reps=4):
    n = 32
    x = torch.randn(n, h + 2 * pad_in, w + 2 * pad_in, cin, device='cuda').to(dt)
    launches = cg.conv_fwd(3, 1, 0 if pad_in else 1, x.shape[1], x.shape[2])
    wp = (torch.randn(9, cout, cin, device='cuda') / (cin * 9) ** 0.5).to(dt)
    y = torch.empty(n, h, w, cout, device='cuda', dtype=dt)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for r in range(reps):
        e0.record()
        ops.conv_gather(x, wp, launches, y, tensor=True)
        e1.record(); torch.cuda.synchronize()
    print(f"{dt} {cin}->{cout} @{h}x{w}: {e0.elapsed_time(e1)*1e3:.1f} us", flush=True)
run(torch.float32, 128, 128, 128, 128)
run(torch.bfloat16, 128, 128, 64, 64, pad_in=1)
run(torch.float32, 256, 256, 64, 64)
