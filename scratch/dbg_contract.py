import torch, sys
sys.path.insert(0, '.')
from artist_style_transfer_b200 import ops, conv_geometry as cg
torch.manual_seed(0)
def rel(a,b): return float((a.double()-b.double()).norm()/(b.double().norm()+1e-30))
for dt in (torch.bfloat16, torch.float32):
    x = torch.randn(2,16,16,128,device='cuda').to(dt)
    gy = torch.randn(2,16,16,128,device='cuda').to(dt)
    l = cg.conv_fwd(1,1,0,16,16)
    a = torch.zeros(128,128,1,1,device='cuda'); b = torch.zeros_like(a)
    ops.wgrad_gather(x, gy, l, a, 128, 1, 1, 1, tensor=True)
    ops.wgrad_gather(x, gy, l, b, 128, 1, 1, 1)
    torch.cuda.synchronize()
    print('wgrad', dt, 'tc norm', float(a.norm()), 'simt norm', float(b.norm()), 'rel', rel(a,b))
    f = torch.randn(2,16,16,128,device='cuda').to(dt)
    g1 = ops.gram(f, 1.0, tensor=True); g2 = ops.gram(f, 1.0)
    torch.cuda.synchronize()
    print('gram ', dt, 'tc norm', float(g1.norm()), 'simt norm', float(g2.norm()), 'rel', rel(g1,g2))
