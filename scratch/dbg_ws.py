"""A/B the weight-stationary halo kernel (AST_CONV_WS modes) against the SIMT kernel."""
import os, sys, torch
sys.path.insert(0, '.')
from artist_style_transfer_b200 import ops, conv_geometry as cg
def rel(a, b): return float((a.double()-b.double()).norm()/(b.double().norm()+1e-30))
def tf32r(x): return (x.view(torch.int32) + 0x1000 & ~0x1FFF).view(torch.float32)
torch.manual_seed(0)
cases = [("fp32 64->64 3x3 (conv1_2)", torch.float32, 64, 64, 3, 1, 2, 40, 24),
         ("bf16 32->32 vertical 9 taps", torch.bfloat16, 32, 32, 0, 0, 2, 40, 24),
         ("bf16 32->32(3) 9x9", torch.bfloat16, 32, 32, 9, 0, 1, 40, 40),
         ("bf16 64->64 3x3", torch.bfloat16, 64, 64, 3, 1, 2, 33, 21),
         ("fp32 16->64 vertical 3 taps", torch.float32, 16, 64, -3, 0, 2, 32, 32)]
for name, dt, cin, cout, k, pad, n, h, w in cases:
    x = torch.randn(n, h, w, cin, device='cuda')
    x = tf32r(x) if dt == torch.float32 else x.to(dt)
    if k > 0:
        L = cg.conv_fwd(k, 1, pad, h, w)
    elif k == 0:
        L = [cg.Launch(h - 8, w, 1, 1, 0, 0, [(dy, 0) for dy in range(9)], [(dy, 0) for dy in range(9)], 0)]
    else:
        L = [cg.Launch(h, w, 1, 1, 0, 0, [(-1, 0), (0, 0), (1, 0)], [(0, 0), (1, 0), (2, 0)], 0)]
    nt = sum(len(l.taps) for l in L)
    wp = torch.randn(nt, cout, cin, device='cuda') / (cin * nt) ** 0.5
    wp = tf32r(wp) if dt == torch.float32 else wp.to(dt)
    ho, wo = L[0].mi, L[0].mj
    ref = torch.empty(n, ho, wo, cout, device='cuda'); ops.conv_gather(x, wp, L, ref)
    for mode in ("0", "1"):
        os.environ["AST_CONV_WS"] = mode
        y = torch.full((n, ho, wo, cout), float('nan'), device='cuda')
        try:
            ops.conv_gather(x, wp, L, y, tensor=True); torch.cuda.synchronize()
            print(f"{name:32s} mode={mode} rel={rel(y, ref):.3e} nan={bool(torch.isnan(y).any())}")
        except Exception as e:
            print(f"{name:32s} mode={mode} ERROR {str(e)[:100]}"); break
