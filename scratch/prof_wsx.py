"""conv_ws / conv_wsx on two representative layers (B=32) for ncu: VGG conv1_2 dgrad (bf16 64->64 + fp32 mask), first-layer-like 32->32."""
import sys, torch
sys.path.insert(0, '.')
from artist_style_transfer_b200 import ops, conv_geometry as cg
torch.manual_seed(0)
n, h, w = 32, 256, 256
def run(cin, cout, mask, reps=3):
    x = torch.randn(n, h, w, cin, device='cuda').bfloat16()
    launches = cg.conv_fwd(3, 1, 1, h, w)
    wp = (torch.randn(9, cout, cin, device='cuda') / (cin * 9) ** 0.5).bfloat16()
    y = torch.empty(n, h, w, cout, device='cuda', dtype=torch.bfloat16)
    m = torch.randn(n, h, w, cout, device='cuda') if mask else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for r in range(reps):
        e0.record()
        ops.conv_gather(x, wp, launches, y, mask=m, tensor=True)
        e1.record(); torch.cuda.synchronize()
    print(f"bf16 {cin}->{cout} mask={mask}: {e0.elapsed_time(e1)*1e3:.1f} us", flush=True)
run(64, 64, True)
run(64, 64, False)
run(32, 32, False)
