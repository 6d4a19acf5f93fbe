"""A few representative conv launches (B=32) for ncu: conv1_1-like (16->64, 3 vertical taps) and conv1_2 (64->64 3x3)."""
import sys, torch
sys.path.insert(0, '.')
from artist_style_transfer_b200 import ops, conv_geometry as cg
torch.manual_seed(0)
n, h, w = 32, 256, 256
def run(cin, cout, launches, reps=3):
    x = torch.randn(n, h, w, cin, device='cuda')
    nt = sum(len(l.taps) for l in launches)
    wp = torch.randn(nt, cout, cin, device='cuda') / (cin * nt) ** 0.5
    b = torch.randn(cout, device='cuda')
    y = torch.empty(n, h, w, cout, device='cuda')
    for _ in range(reps):
        ops.conv_gather(x, wp, launches, y, bias=b, relu=True, tensor=True, round_tf32=True)
    torch.cuda.synchronize()
run(16, 64, [cg.Launch(h, w, 1, 1, 0, 0, [(-1, 0), (0, 0), (1, 0)], [(0, 0), (1, 0), (2, 0)], 0)])
run(64, 64, cg.conv_fwd(3, 1, 1, h, w))
print("done")
