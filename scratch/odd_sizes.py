"""Robustness: fast (tensor-core) vs strict mode on image sizes that are not multiples of 4 / 8 / 16."""
import sys, torch
sys.path.insert(0, '.')
import artist_style_transfer_b200 as ast
dev = torch.device('cuda')
torch.manual_seed(0)
net_f = ast.StyleTransfer(device=dev, precision='fast')
net_s = ast.StyleTransfer(device=dev, precision='fp32')
net_s.load_state_dict(net_f.state_dict())
vgg_f = ast.VGG16(vgg_path=None, precision='fast').to(dev)
vgg_s = ast.VGG16(vgg_path=None, precision='fp32').to(dev)
vgg_s.load_state_dict(vgg_f.state_dict())
for (h, w) in [(250, 330), (255, 257), (37, 53), (252, 332), (68, 76), (36, 52), (132, 140), (260, 1084)]:
    x = torch.randint(0, 256, (2, 3, h, w), device=dev).float()
    with torch.no_grad():
        yf, ys = net_f(x), net_s(x)
    rel = float((yf - ys).norm() / ys.norm())
    if h % 4 or w % 4:        # the net returns a different size then: the reference's losses raise a shape mismatch too
        print(f"{h}x{w}: forward rel {rel:.2e}  out {tuple(yf.shape)}", flush=True)
        continue
    # one training step (losses only) in both modes
    style = torch.randint(0, 256, (3, h, w), device=dev).float()
    out = []
    for net, vgg in ((net_f, vgg_f), (net_s, vgg_s)):
        sg = ast.style_grams_single(vgg, style, 2)
        net.zero_grad()
        c, s, t = ast.perceptual_step(net, vgg, x, sg)
        g = torch.cat([p.grad.flatten() for p in net.parameters()])
        out.append((float(c), float(s), g))
    grel = float((out[0][2] - out[1][2]).norm() / out[1][2].norm())
    print(f"{h}x{w}: forward rel {rel:.2e}  content {out[0][0]:.4f}/{out[1][0]:.4f}  style {out[0][1]:.4e}/{out[1][1]:.4e}  grad rel {grel:.2e}  out {tuple(yf.shape)}", flush=True)
