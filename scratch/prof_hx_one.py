"""Two conv_hx launches for ncu: residual 3x3 forward (bf16, with fused statistics) and VGG conv2_2 (TF32), B=32."""
import sys, torch
sys.path.insert(0, '.')
from artist_style_transfer_b200 import ops, conv_geometry as cg
torch.manual_seed(0)
n = 32
def run(dtype, cin, cout, hin, launches, hout, stats):
    x = torch.randn(n, hin, hin, cin, device='cuda').to(dtype)
    y = torch.empty(n, hout, hout, cout, device='cuda', dtype=dtype)
    nt = sum(len(l.taps) for l in launches)
    wp = (torch.randn(nt, cout, cin, device='cuda') / (cin * nt) ** 0.5).to(dtype)
    sums = torch.zeros(2 * n * cout, dtype=torch.float64, device='cuda') if stats else None
    for _ in range(2):
        ops.conv_gather(x, wp, launches, y, tensor=True, stats=sums)
    torch.cuda.synchronize()
run(torch.bfloat16, 128, 128, 66, cg.conv_fwd(3, 1, 0, 66, 66), 64, True)
run(torch.float32, 128, 128, 128, cg.conv_fwd(3, 1, 1, 128, 128), 128, False)
print("done")
