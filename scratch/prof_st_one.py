"""One block-stacked launch for ncu: case = first | deconv2."""
import sys, torch
sys.path.insert(0, '.')
from artist_style_transfer_b200 import ops, conv_geometry as cg
torch.manual_seed(0)
case = sys.argv[1] if len(sys.argv) > 1 else "first"
n, bf = 32, torch.bfloat16
if case == "first":
    ls = [cg.Launch(256, 256, 1, 1, 0, 0, [(d, 0) for d in range(9)], [(d, 0) for d in range(9)], 0)]
    x = torch.randn(n, 264, 256, 32, device='cuda').to(bf); cout = 32
    y = torch.empty(n, 256, 256, 32, device='cuda', dtype=bf)
    stk = cg.stack_rows(ls[0], 4)
else:
    ls = cg.convT_fwd(3, 2, 1, 1, 128, 128)
    x = torch.randn(n, 128, 128, 64, device='cuda').to(bf); cout = 32
    y = torch.empty(n, 256, 256, 32, device='cuda', dtype=bf)
    stk = cg.stack_phases(ls)
nt = sum(len(l.taps) for l in ls)
wp = (torch.randn(nt, cout, x.shape[3], device='cuda') / (x.shape[3] * nt) ** 0.5).to(bf)
tidx = {wt: l.woff + t for l in ls for t, wt in enumerate(l.wtaps)}
w = ops.stack_filter(lambda pos: wp[tidx[pos]], stk, cout, x.shape[3], bf, 'cuda')
sums = torch.zeros(2 * n * cout, dtype=torch.float64, device='cuda')
for _ in range(3):
    ops.conv_stacked(x, w, stk, y, stats=sums if "nostats" not in sys.argv else None)
torch.cuda.synchronize()
print("ok")
