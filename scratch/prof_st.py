"""Block-stacked conv kernel (ast_conv_stacked) vs the kernels the same layers ran on before: parity + timing at B=32."""
import sys, torch
sys.path.insert(0, '.')
from artist_style_transfer_b200 import _lib, ops, conv_geometry as cg
torch.manual_seed(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
bf, f32 = torch.bfloat16, torch.float32


def timeit(fn, sets, reps=10):
    for i in range(3): fn(i % sets)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps): fn(i % sets)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def run(name, dtype, odt, cin, cout, hin, win, launches, hout, wout, mode, stats=False, mask=False, relu=False, bias=False,
        round_tf32=False, sets=3):
    nt = sum(len(l.taps) for l in launches)
    wp = (torch.randn(nt, cout, cin, device='cuda') / (cin * nt) ** 0.5)
    if dtype == f32:
        wp = ops.round_tf32(wp) if hasattr(ops, "round_tf32") else wp
    wp = wp.to(dtype)
    b = torch.randn(cout, device='cuda') if bias else None
    xs, ys, zs, ms = [], [], [], []
    for _ in range(sets):
        x = torch.randn(n, hin, win, cin, device='cuda')
        xs.append(x.to(dtype))
        ys.append(torch.zeros(n, hout, wout, cout, device='cuda', dtype=odt))
        zs.append(torch.zeros(n, hout, wout, cout, device='cuda', dtype=odt))
        ms.append(torch.randn(n, hout, wout, cout, device='cuda').to(bf) if mask else None)
    s_old = torch.zeros(2 * n * cout, dtype=torch.float64, device='cuda') if stats else None
    s_new = torch.zeros(2 * n * cout, dtype=torch.float64, device='cuda') if stats else None
    # tile index of every kernel position in the plain pack
    tidx = {}
    for l in launches:
        for t, wt in enumerate(l.wtaps):
            tidx[wt] = l.woff + t
    if mode == "phases":
        stks = [cg.stack_phases(launches[i:i + 128 // cout]) for i in range(0, len(launches), 128 // cout)]
    else:
        stks = [cg.stack_rows(launches[0], 128 // cout)]
    wst = [ops.stack_filter(lambda pos: wp[tidx[pos]], s, cout, cin, dtype, 'cuda') for s in stks]

    def old(i):
        ops.conv_gather(xs[i], wp, launches, ys[i], mask=ms[i], tensor=True, stats=s_old, relu=relu, bias=b, round_tf32=round_tf32)

    def new(i):
        for s, w in zip(stks, wst):
            ops.conv_stacked(xs[i], w, s, zs[i], mask=ms[i], stats=s_new, relu=relu, bias=b, round_tf32=round_tf32)

    before = _lib.family_stats()
    old(0)
    fam = [k for k, v in _lib.family_delta(before).items() if v[0]]
    if stats: s_old.zero_(); s_new.zero_(); old(0)
    new(0)
    torch.cuda.synchronize()
    a, c = ys[0].float(), zs[0].float()
    err = (a - c).abs().max().item()
    ref = a.abs().max().item()
    serr = ((s_old - s_new).abs().max() / s_old.abs().max()).item() if stats else 0.0
    t_old, t_new = timeit(old, sets), timeit(new, sets)
    gf = 2.0 * n * sum(l.mi * l.mj * len(l.taps) for l in launches) * cin * cout / 1e9
    mb = (xs[0].numel() * xs[0].element_size() + ys[0].numel() * ys[0].element_size()) / 1e6
    print(f"{name:40s} {str(fam):14s} old {t_old:7.1f} us  new {t_new:7.1f} us ({gf/t_new*1e3:6.1f} TF/s {mb/t_new:5.2f} TB/s)  "
          f"max|diff| {err:.3e} of {ref:.2f}  stats rel {serr:.1e}  vt={[len(s.vt) for s in stks]}", flush=True)


vt9 = [cg.Launch(256, 256, 1, 1, 0, 0, [(d, 0) for d in range(9)], [(d, 0) for d in range(9)], 0)]
run("T first layer vt9 32->32 +stats", bf, bf, 32, 32, 264, 256, vt9, 256, 256, "rows", stats=True)
run("T deconv2 3x3 s2 64->32 128^2->256^2 +stats", bf, bf, 64, 32, 128, 128, cg.convT_fwd(3, 2, 1, 1, 128, 128), 256, 256, "phases", stats=True)
run("T conv2 dgrad 64->32 128^2->258^2", bf, bf, 64, 32, 128, 128, cg.conv_dgrad(3, 2, 0, 258, 258), 258, 258, "phases")
vt9b = [cg.Launch(256, 264, 1, 1, 0, 0, [(d, 0) for d in range(9)], [(d, 0) for d in range(9)], 0)]
run("T last fwd vt9 32->32 fp32 out", bf, f32, 32, 32, 264, 264, vt9b, 256, 264, "rows")
vt9c = [cg.Launch(264, 264, 1, 1, 0, 0, [(-d, 0) for d in range(9)], [(d, 0) for d in range(9)], 0)]
run("T last dgrad vt9 32->32", bf, bf, 32, 32, 256, 264, vt9c, 264, 264, "rows")
vt3 = [cg.Launch(256, 256, 1, 1, 0, 0, [(-1, 0), (0, 0), (1, 0)], [(0, 0), (1, 0), (2, 0)], 0)]
run("VGG conv1_1 vt3 16->64 tf32 +relu+bias", f32, f32, 16, 64, 256, 256, vt3, 256, 256, "rows", relu=True, bias=True, round_tf32=True)
vt3b = [cg.Launch(256, 258, 1, 1, 0, 0, [(1, 0), (0, 0), (-1, 0)], [(0, 0), (1, 0), (2, 0)], 0)]
run("VGG conv1_1 dgrad vt3 64->32 bf16", bf, bf, 64, 32, 256, 258, vt3b, 256, 258, "rows")
run("T deconv1 3x3 s2 128->64 64^2->128^2 +stats", bf, bf, 128, 64, 64, 64, cg.convT_fwd(3, 2, 1, 1, 64, 64), 128, 128, "phases", stats=True)
run("T conv3 dgrad 128->64 64^2->130^2", bf, bf, 128, 64, 64, 64, cg.conv_dgrad(3, 2, 0, 130, 130), 130, 130, "phases")
run("VGG conv1_2 64->64 tf32 +relu+bias", f32, f32, 64, 64, 256, 256, cg.conv_fwd(3, 1, 1, 256, 256), 256, 256, "rows", relu=True, bias=True, round_tf32=True)
run("VGG dgrad conv1_2 64->64 bf16 +mask", bf, bf, 64, 64, 256, 256, cg.conv_dgrad(3, 1, 1, 256, 256), 256, 256, "rows", mask=True)
