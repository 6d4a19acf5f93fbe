"""Thin filter gradient of the first layer (9 vertical taps, 32 x 32 channels) - AST_THIN_TH sweep."""
import sys, torch
sys.path.insert(0, '.')
from artist_style_transfer_b200 import ops, conv_geometry as cg
torch.manual_seed(0)
n, bf = 32, torch.bfloat16
vt9 = [cg.Launch(256, 256, 1, 1, 0, 0, [(d, 0) for d in range(9)], [(d, 0) for d in range(9)], 0)]
sets = [(torch.randn(n, 264, 256, 32, device='cuda').to(bf), torch.randn(n, 256, 256, 32, device='cuda').to(bf)) for _ in range(3)]
dw = torch.zeros(9, 32, 32, device='cuda')
def go(i):
    x, g = sets[i % 3]
    ops.wgrad_gather(x, g, vt9, dw, 32, 1, 32 * 32, 0, tensor=True)
for i in range(3): go(i)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(20): go(i)
e1.record(); torch.cuda.synchronize()
print(sys.argv[1] if len(sys.argv) > 1 else "", f"{e0.elapsed_time(e1) / 20 * 1e3:.1f} us")
