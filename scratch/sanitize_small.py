"""Tiny launches of every tcgen05 kernel for compute-sanitizer memcheck."""
import sys, torch
sys.path.insert(0, '.')
from artist_style_transfer_b200 import ops, conv_geometry as cg
torch.manual_seed(0)
x = torch.randn(1, 20, 12, 64, device='cuda').bfloat16()
L = cg.conv_fwd(3, 1, 1, 20, 12)
wp = (torch.randn(9, 64, 64, device='cuda') / 24).bfloat16()
y = torch.empty(1, 20, 12, 64, device='cuda', dtype=torch.bfloat16)
ops.conv_gather(x, wp, L, y, tensor=True)                      # conv_ws (weights resident)
x2 = torch.randn(1, 16, 16, 128, device='cuda')
wp2 = torch.randn(9, 256, 128, device='cuda') / 34
y2 = torch.empty(1, 16, 16, 256, device='cuda')
ops.conv_gather(x2, wp2, cg.conv_fwd(3, 1, 1, 16, 16), y2, tensor=True)   # conv_tc (streamed)
g = torch.randn(1, 18, 10, 64, device='cuda').bfloat16()
dw = torch.zeros(64, 64, 3, 3, device='cuda')
ops.wgrad_gather(x, g, cg.conv_fwd(3, 1, 0, 20, 12), dw, 64 * 9, 9, 3, 1, tensor=True)   # contract_tc
f = torch.randn(2, 8, 8, 64, device='cuda')
ops.gram(f, 1.0, tensor=True)
torch.cuda.synchronize()
print("sanitize run ok", float(y.float().abs().sum()), float(y2.abs().sum()), float(dw.abs().sum()))
