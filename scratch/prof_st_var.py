"""Times ast_conv_stacked cases with the library AST_B200_LIB points at (variant A/B runs)."""
import sys, torch
sys.path.insert(0, '.')
from artist_style_transfer_b200 import ops, conv_geometry as cg
torch.manual_seed(0)
n, bf, f32 = 32, torch.bfloat16, torch.float32
tag = sys.argv[1] if len(sys.argv) > 1 else ""


def run(name, dtype, odt, cin, cout, hin, win, launches, hout, wout, mode, stats=False, sets=3, reps=10, pair=False):
    nt = sum(len(l.taps) for l in launches)
    wp = (torch.randn(nt, cout, cin, device='cuda') / (cin * nt) ** 0.5).to(dtype)
    tidx = {wt: l.woff + t for l in launches for t, wt in enumerate(l.wtaps)}
    nb = 128 // cout
    stks = [cg.stack_phases(launches[i:i + nb]) for i in range(0, len(launches), nb)] if mode == "phases" else [cg.stack_rows(launches[0], nb)]
    wst = [ops.stack_filter(lambda pos: wp[tidx[pos]], s, cout, cin, dtype, 'cuda') for s in stks]
    xs = [torch.randn(n, hin, win, cin, device='cuda').to(dtype) for _ in range(sets)]
    ys = [torch.zeros(n, hout, wout, cout, device='cuda', dtype=odt) for _ in range(sets)]
    sums = torch.zeros(2 * n * cout, dtype=torch.float64, device='cuda') if stats else None
    def go(i):
        for s, w in zip(stks, wst):
            ops.conv_stacked(xs[i], w, s, ys[i], stats=sums)
    for i in range(3): go(i % sets)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps): go(i % sets)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    mb = (xs[0].numel() * xs[0].element_size() + ys[0].numel() * ys[0].element_size()) / 1e6
    print(f"{tag:8s} {name:34s} {us:7.1f} us  {mb/us:5.2f} TB/s", flush=True)


vt9 = [cg.Launch(256, 256, 1, 1, 0, 0, [(d, 0) for d in range(9)], [(d, 0) for d in range(9)], 0)]
run("first vt9 32->32 +stats", bf, bf, 32, 32, 264, 256, vt9, 256, 256, "rows", stats=True)
run("first vt9 32->32", bf, bf, 32, 32, 264, 256, vt9, 256, 256, "rows")
run("first vt9 32->32 fp32 out", bf, f32, 32, 32, 264, 256, vt9, 256, 256, "rows")
run("deconv2 64->32 +stats", bf, bf, 64, 32, 128, 128, cg.convT_fwd(3, 2, 1, 1, 128, 128), 256, 256, "phases", stats=True)
run("deconv2 64->32", bf, bf, 64, 32, 128, 128, cg.convT_fwd(3, 2, 1, 1, 128, 128), 256, 256, "phases")
vt3 = [cg.Launch(256, 256, 1, 1, 0, 0, [(-1, 0), (0, 0), (1, 0)], [(0, 0), (1, 0), (2, 0)], 0)]
run("conv1_1 vt3 16->64 tf32", f32, f32, 16, 64, 256, 256, vt3, 256, 256, "rows")
vt3b = [cg.Launch(256, 258, 1, 1, 0, 0, [(1, 0), (0, 0), (-1, 0)], [(0, 0), (1, 0), (2, 0)], 0)]
run("conv1_1 dgrad vt3 64->32 bf16", bf, bf, 64, 32, 256, 258, vt3b, 256, 258, "rows")
