"""How much of the fast-mode gradient deviation is bf16 storage noise (SIMT bf16 vs tcgen05 bf16 vs fp32)?"""
import os, sys, torch
sys.path.insert(0, '.')
import artist_style_transfer_b200 as ast
from oracle import weights
def build(mode):
    net = ast.StyleTransfer(device=torch.device('cuda'), precision=mode); net.load_state_dict(weights.transfer_state_dict(2))
    vgg = ast.VGG16(vgg_path=None, precision=mode).cuda(); vgg.load_state_dict(weights.vgg_state_dict(2), strict=False)
    return net, vgg
content = weights.content_batch(2, 64, 2).cuda()
res = {}
for tag, mode, dis, gf in (("fp32", "fp32", "0", "0"), ("bf16_simt", "fast", "0", "1"), ("bf16_tc", "fast", "0", "0")):
    os.environ["AST_DISABLE_TC"] = dis
    os.environ["AST_GRAD_FP32"] = gf
    net, vgg = build(mode)
    style = ast.style_grams_single(vgg, weights.style_image(64, 2).cuda(), 2)
    net.zero_grad(); c, s, t = ast.perceptual_step(net, vgg, content, style)
    res[tag] = (float(t), {n: p.grad.clone() for n, p in net.named_parameters()})
def cos(a, b): return float((a*b).sum()/(a.norm()*b.norm()+1e-30))
print({k: v[0] for k, v in res.items()})
for n in res["fp32"][1]:
    g = res["fp32"][1][n]
    if float(g.norm()) < 1e-9: continue
    print(f"{n:45s} cos(fp32,tc16+g32)={cos(g,res['bf16_simt'][1][n]):.4f} cos(fp32,tc16)={cos(g,res['bf16_tc'][1][n]):.4f} cos(g32,g16)={cos(res['bf16_simt'][1][n],res['bf16_tc'][1][n]):.4f}")
