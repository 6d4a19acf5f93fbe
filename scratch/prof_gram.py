"""Gram variants per tap shape (B=32, 256^2 step): plain ast_gram (memset + contraction + mirror) vs fused ast_gram_mse."""
import sys, torch
sys.path.insert(0, '.')
from artist_style_transfer_b200 import _lib, ops
dev = torch.device('cuda')
def timed(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
B = 32
for c, s in ((64, 256), (128, 128), (256, 64), (512, 32)):
    x = torch.randn(B, s, s, c, device=dev)
    tgt = torch.randn(c, c, device=dev)
    g = torch.zeros(B, c, c, device=dev); cnt = torch.zeros(B * 16, dtype=torch.int32, device=dev)
    loss = torch.zeros(1, dtype=torch.float64, device=dev); d = torch.empty(B, c, c, device=dev)
    def fused(with_d=True, with_t=True):
        g.zero_(); cnt.zero_()
        ops.gram_mse(x, tgt, g, cnt, loss=loss, loss_scale=1.0, d=d if with_d else None, d_scale=1.0, tensor=True)
    def zero_only():
        g.zero_(); cnt.zero_()
    t_plain = timed(lambda: ops.gram(x, 1.0 / (c * s * s), tensor=True))
    t_fused = timed(fused)
    t_fused_nod = timed(lambda: fused(False))
    t_zero = timed(zero_only)
    mb = x.numel() * 4 / 1e6
    print(f"C={c:4d} HW={s}^2 ({mb:.0f} MB): plain {t_plain:7.1f} us  fused {t_fused:7.1f} us  fused(no D) {t_fused_nod:7.1f} us  zeroing {t_zero:5.1f} us  -> HBM floor {mb/6.4649e3*1e3:6.1f} us")
