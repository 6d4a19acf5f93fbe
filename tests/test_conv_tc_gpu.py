"""tcgen05/TMA gather-conv kernel against the SIMT kernel and torch fp64 on identical inputs (B200 only)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from artist_style_transfer_b200 import _lib, conv_geometry as cg, ops
    if not _lib.has_tc_conv():
        pytest.skip("library built without the tcgen05 conv kernel")
    return cg, ops


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def tf32_round(x):
    return (x.view(torch.int32) + 0x1000 & ~0x1FFF).view(torch.float32) if x.dtype == torch.float32 else x


CASES = [
    # dtype, cin, cout, k, stride, pad, n, h, w
    (torch.bfloat16, 128, 128, 3, 1, 1, 2, 16, 16),
    (torch.bfloat16, 64, 128, 3, 1, 1, 1, 8, 16),
    (torch.bfloat16, 128, 128, 1, 1, 0, 2, 16, 8),
    (torch.bfloat16, 32, 64, 3, 1, 1, 1, 16, 16),
    (torch.bfloat16, 64, 32, 3, 1, 1, 1, 20, 12),
    (torch.bfloat16, 32, 64, 3, 2, 0, 2, 18, 18),
    (torch.bfloat16, 64, 128, 3, 2, 0, 1, 34, 18),
    (torch.float32, 64, 64, 3, 1, 1, 2, 16, 16),
    (torch.float32, 128, 256, 3, 1, 1, 1, 16, 16),
    (torch.float32, 256, 512, 3, 1, 1, 1, 8, 8),
    (torch.float32, 512, 512, 3, 1, 1, 2, 4, 4),
    (torch.float32, 64, 128, 3, 1, 1, 1, 13, 21),
]


@pytest.mark.parametrize("dtype,cin,cout,k,s,pad,n,h,w", CASES)
def test_conv_fwd_tc(env, dtype, cin, cout, k, s, pad, n, h, w):
    cg, ops = env
    torch.manual_seed(cin * 7 + cout + k + s)
    x = torch.randn(n, h, w, cin, device="cuda").to(dtype)
    wt = (torch.randn(cout, cin, k, k, device="cuda") / (cin * k * k) ** 0.5)
    bias = torch.randn(cout, device="cuda")
    launches = cg.conv_fwd(k, s, pad, h, w)
    ho, wo = launches[0].mi, launches[0].mj
    wp = ops.pack_weights(wt, launches, cout, cin, cin * k * k, k * k, k, 1, dtype)
    y_tc = torch.full((n, ho, wo, cout), float("nan"), device="cuda", dtype=torch.float32)
    ops.conv_gather(x, wp, launches, y_tc, bias=bias, relu=True, tensor=True)
    y_simt = torch.empty_like(y_tc)
    ops.conv_gather(x, wp, launches, y_simt, bias=bias, relu=True)
    torch.cuda.synchronize()
    xr = x.double().cpu().permute(0, 3, 1, 2)
    wr = wp.double().cpu().view(k, k, cout, cin).permute(2, 3, 0, 1)
    if dtype == torch.float32:   # kind::tf32 truncates the operands to 10 mantissa bits
        tol = 2e-3
    else:
        tol = 1e-5               # same bf16 operands, fp32 accumulate: only summation order differs
    yr = F.relu(F.conv2d(xr, wr, bias.double().cpu(), stride=s, padding=pad)).permute(0, 2, 3, 1)
    assert not torch.isnan(y_tc).any()
    assert rel(y_simt, yr) < 1e-5
    assert rel(y_tc, yr) < tol, rel(y_tc, yr)


def test_conv_tc_tf32_exact_when_prerounded(env):
    """With operands already rounded to TF32 the tensor-core result matches fp32 FFMA to accumulation order."""
    cg, ops = env
    from artist_style_transfer_b200 import _lib
    torch.manual_seed(1)
    n, h, w, cin, cout = 2, 16, 16, 64, 128
    x = tf32_round(torch.randn(n, h, w, cin, device="cuda"))
    wt = torch.randn(cout, cin, 3, 3, device="cuda") / 24
    launches = cg.conv_fwd(3, 1, 1, h, w)
    wp = ops.pack_weights(wt, launches, cout, cin, cin * 9, 9, 3, 1, torch.float32)
    wp = tf32_round(wp)
    y_tc = torch.empty((n, h, w, cout), device="cuda")
    y_simt = torch.empty_like(y_tc)
    ops.conv_gather(x, wp, launches, y_tc, tensor=True)
    ops.conv_gather(x, wp, launches, y_simt)
    assert rel(y_tc, y_simt) < 2e-6


def test_conv_tc_epilogue_add_mask_stride2_out(env):
    """dgrad-style launch: output stride 2 phases, tap-gradient add, ReLU mask, bf16 output."""
    cg, ops = env
    torch.manual_seed(2)
    n, h, w, cin, cout = 2, 8, 8, 128, 64          # convT 3x3 s2: (n,8,8,128) -> (n,16,16,64)
    x = torch.randn(n, h, w, cin, device="cuda").bfloat16()
    wt = torch.randn(cin, cout, 3, 3, device="cuda") / 30
    launches = cg.convT_fwd(3, 2, 1, 1, h, w)
    wp = ops.pack_weights(wt, launches, cout, cin, 9, cout * 9, 3, 1, torch.bfloat16)
    add = torch.randn(n, 16, 16, cout, device="cuda")
    mask = torch.randn(n, 16, 16, cout, device="cuda")
    y_tc = torch.full((n, 16, 16, cout), float("nan"), device="cuda", dtype=torch.bfloat16)
    y_simt = torch.empty_like(y_tc)
    ops.conv_gather(x, wp, launches, y_tc, add=add, mask=mask, tensor=True)
    ops.conv_gather(x, wp, launches, y_simt, add=add, mask=mask)
    assert not torch.isnan(y_tc.float()).any()
    assert rel(y_tc, y_simt) < 5e-3
    yr = F.conv_transpose2d(x.double().cpu().permute(0, 3, 1, 2), wt.bfloat16().double().cpu(), stride=2, padding=1,
                            output_padding=1).permute(0, 2, 3, 1)
    yr = (yr + add.double().cpu()) * (mask.double().cpu() > 0)
    assert rel(y_tc, yr) < 5e-3


def test_conv_tc_per_image_weights(env):
    """Gram backward shape: 1x1 conv with a different C x C matrix per image."""
    cg, ops = env
    torch.manual_seed(3)
    n, h, w, c = 3, 16, 16, 128
    x = torch.randn(n, h, w, c, device="cuda")
    d = torch.randn(n, 1, c, c, device="cuda") / 11
    launches = cg.conv_fwd(1, 1, 0, h, w)
    y_tc = torch.empty((n, h, w, c), device="cuda")
    y_simt = torch.empty_like(y_tc)
    ops.conv_gather(x, d, launches, y_tc, w_img_stride=c * c, tensor=True)
    ops.conv_gather(x, d, launches, y_simt, w_img_stride=c * c)
    assert rel(y_tc, y_simt) < 2e-3


def test_conv_tc_rejects_unsupported(env):
    cg, ops = env
    x = torch.randn(1, 8, 8, 3, device="cuda")
    wp = torch.randn(9, 64, 3, device="cuda")
    y = torch.empty((1, 8, 8, 64), device="cuda")
    with pytest.raises(RuntimeError, match="conv_tc"):
        ops.conv_gather(x, wp, cg.conv_fwd(3, 1, 1, 8, 8), y, tensor=True)


WG_CASES = [
    # dtype, cin, cout, k, stride, n, h, w  (input is the physically padded buffer, like the transform net)
    (torch.bfloat16, 128, 128, 3, 1, 2, 18, 18),
    (torch.bfloat16, 64, 128, 3, 2, 2, 18, 18),
    (torch.bfloat16, 32, 64, 3, 2, 1, 34, 34),
    (torch.bfloat16, 128, 128, 1, 1, 2, 16, 16),
    (torch.bfloat16, 128, 128, 3, 1, 1, 13, 11),
    # <= 32 channels on both sides: the thin kernel (four taps stacked in the M rows, 64-byte swizzle)
    (torch.bfloat16, 32, 32, 3, 1, 2, 18, 18),
    (torch.bfloat16, 32, 32, 5, 1, 2, 21, 19),
    (torch.bfloat16, 32, 32, 3, 2, 1, 35, 33),
    (torch.bfloat16, 16, 32, 3, 1, 3, 40, 70),
    (torch.bfloat16, 32, 24, 2, 1, 1, 9, 9),
]


@pytest.mark.parametrize("dtype,cin,cout,k,s,n,h,w", WG_CASES)
def test_wgrad_tc(env, dtype, cin, cout, k, s, n, h, w):
    cg, ops = env
    torch.manual_seed(cin + cout + k + s + h)
    x = torch.randn(n, h, w, cin, device="cuda").to(dtype)
    launches = cg.conv_fwd(k, s, 0, h, w)
    ho, wo = launches[0].mi, launches[0].mj
    gy = torch.randn(n, ho, wo, cout, device="cuda").to(dtype)
    dw_tc = torch.zeros(cout, cin, k, k, device="cuda")
    dw_simt = torch.zeros_like(dw_tc)
    ops.wgrad_gather(x, gy, launches, dw_tc, cin * k * k, k * k, k, 1, tensor=True)
    ops.wgrad_gather(x, gy, launches, dw_simt, cin * k * k, k * k, k, 1)
    xr = x.double().cpu().permute(0, 3, 1, 2).requires_grad_(False)
    wr = torch.zeros(cout, cin, k, k, dtype=torch.float64, requires_grad=True)
    F.conv2d(xr, wr, stride=s).backward(gy.double().cpu().permute(0, 3, 1, 2))
    assert rel(dw_simt, wr.grad) < 1e-5
    assert rel(dw_tc, wr.grad) < 1e-5, rel(dw_tc, wr.grad)
    # tap-major scratch layout [ky][kx][co][ci] (ci contiguous -> 16-byte vector atomics in the epilogue)
    dw_tm = torch.zeros(k, k, cout, cin, device="cuda")
    ops.wgrad_gather(x, gy, launches, dw_tm, cin, 1, k * cout * cin, cout * cin, tensor=True)
    assert rel(dw_tm.permute(2, 3, 0, 1), wr.grad) < 1e-5


def test_wgrad_tc_convT_phases(env):
    cg, ops = env
    torch.manual_seed(9)
    n, h, w, cin, cout = 2, 8, 8, 128, 64
    x = torch.randn(n, h, w, cin, device="cuda").bfloat16()
    launches = cg.convT_fwd(3, 2, 1, 1, h, w)
    gy = torch.randn(n, 16, 16, cout, device="cuda").bfloat16()
    dw_tc = torch.zeros(cin, cout, 3, 3, device="cuda")
    ops.wgrad_gather(x, gy, launches, dw_tc, 9, cout * 9, 3, 1, tensor=True)
    wr = torch.zeros(cin, cout, 3, 3, dtype=torch.float64, requires_grad=True)
    F.conv_transpose2d(x.double().cpu().permute(0, 3, 1, 2), wr, stride=2, padding=1, output_padding=1).backward(
        gy.double().cpu().permute(0, 3, 1, 2))
    assert rel(dw_tc, wr.grad) < 1e-5


@pytest.mark.parametrize("c,h,w,n", [(64, 32, 32, 2), (128, 16, 16, 3), (256, 8, 8, 2), (512, 4, 4, 2), (64, 13, 9, 1)])
def test_gram_tc(env, c, h, w, n):
    cg, ops = env
    from oracle import port
    torch.manual_seed(c + h)
    f = tf32_round(torch.randn(n, h, w, c, device="cuda") * 3)
    g_tc = ops.gram(f, 1.0 / (c * h * w), tensor=True)
    ref = port.gram(f.double().cpu().permute(0, 3, 1, 2))
    assert rel(g_tc, ref) < 1e-5, rel(g_tc, ref)
    assert float((g_tc - g_tc.transpose(1, 2)).abs().max()) == 0.0


@pytest.mark.parametrize("pool_only", [False, True])
def test_conv_ws_fused_maxpool(env, pool_only):
    """ast_gather_geom.pooled: the weight-stationary kernel also writes MaxPool2d(2,2) of conv+bias+ReLU (VGG conv1_2 ->
    pool, train_cnn.py:54,72-73); with AST_CONV_POOL_ONLY the full-resolution output is left untouched."""
    cg, ops = env
    torch.manual_seed(11)
    n, h, w, c = 2, 20, 24, 64
    x = torch.randn(n, h, w, c, device="cuda").half()          # the fast-mode VGG feeds conv1_2 fp16 activations
    wt = (torch.randn(c, c, 3, 3, device="cuda") / 24).half().float()
    bias = torch.randn(c, device="cuda")
    launches = cg.conv_fwd(3, 1, 1, h, w)
    wp = ops.pack_weights(wt, launches, c, c, c * 9, 9, 3, 1, torch.float32).half()
    y = torch.full((n, h, w, c), -7.0, device="cuda")
    yp = torch.empty(n, h // 2, w // 2, c, device="cuda")
    ops.conv_gather(x, wp, launches, y, bias=bias, relu=True, tensor=True, pooled=yp, pool_only=pool_only)
    ref = F.relu(F.conv2d(x.permute(0, 3, 1, 2).double().cpu(), wt.double().cpu(), bias.double().cpu(), padding=1))
    refp = F.max_pool2d(ref, 2, 2).permute(0, 2, 3, 1)
    assert rel(yp, refp) < 1e-5                      # identical fp16 operands, fp32 accumulation
    if pool_only:
        assert bool((y == -7.0).all())
    else:
        assert rel(y, ref.permute(0, 2, 3, 1)) < 1e-5
        assert torch.equal(yp, F.max_pool2d(y.permute(0, 3, 1, 2), 2, 2).permute(0, 2, 3, 1))   # pooling itself is exact


def test_conv_pooled_rejected_where_unsupported(env):
    cg, ops = env
    x = torch.randn(1, 8, 8, 512, device="cuda")
    wp = torch.randn(9, 512, 512, device="cuda")
    y = torch.empty(1, 8, 8, 512, device="cuda")
    yp = torch.empty(1, 4, 4, 512, device="cuda")
    with pytest.raises(RuntimeError, match="pooled"):
        ops.conv_gather(x, wp, cg.conv_fwd(3, 1, 1, 8, 8), y, tensor=True, pooled=yp)


@pytest.mark.parametrize("c,h,w,n,shared", [(64, 32, 32, 2, True), (128, 16, 24, 3, False), (256, 16, 16, 2, True),
                                            (512, 8, 8, 3, False), (256, 64, 64, 32, True), (512, 32, 32, 32, True),
                                            (64, 13, 9, 1, False)])
def test_gram_mse_fused(env, c, h, w, n, shared):
    """ast_gram_mse (train_cnn.py:321-325 + :103-107 in one kernel): Gram, style-MSE numerator and the Gram-backward
    weights D = d_scale (G - S) against torch fp64, for the split-K / ticket finish (C <= 128) and the single-writer
    register finish (C >= 256 at enough images), shared and per-image targets."""
    cg, ops = env
    from oracle import port
    torch.manual_seed(c + h + n)
    f = tf32_round(torch.randn(n, h, w, c, device="cuda") * 2)
    tgt = torch.randn(c, c, device="cuda") if shared else torch.randn(n, c, c, device="cuda")
    tgt = (tgt + tgt.transpose(-1, -2)) / 2                       # style Grams are symmetric
    g = torch.zeros(n, c, c, device="cuda")
    counters = torch.zeros(n * 16, dtype=torch.int32, device="cuda")
    loss = torch.zeros(1, dtype=torch.float64, device="cuda")
    d = torch.full((n, c, c), float("nan"), device="cuda")
    ops.gram_mse(f, tgt, g, counters, loss=loss, loss_scale=0.5, d=d, d_scale=3.0, tensor=True)
    ref = port.gram(f.double().cpu().permute(0, 3, 1, 2))
    diff = ref - tgt.double().cpu()
    # fp32 TMEM accumulation over K = h*w pixels in ONE CTA (single-writer blocks): the tensor core's accumulate step
    # truncates, so long sums of squares (the diagonal) drift by ~K/8 * 2^-25: 1e-5 up to K = 1024, 1e-4 at K = 4096
    assert rel(g, ref) < (1e-5 if h * w <= 1024 else 1e-4), rel(g, ref)
    assert float((g - g.transpose(1, 2)).abs().max()) == 0.0
    assert abs(float(loss) - 0.5 * float((diff ** 2).sum())) < 1e-4 * 0.5 * float((diff ** 2).sum())
    assert not torch.isnan(d).any()
    assert rel(d, 3.0 * diff) < 1e-3                              # D is rounded to TF32 (it feeds a kind::tf32 conv)
    assert float((d - d.transpose(1, 2)).abs().max()) == 0.0


# ---------------------------------------------------------------------------------------------------------------------
# Block-stacked kernel (ast_conv_stacked, conv_st.cu): same results as the plain gather launches it replaces.
STACKED_CASES = [
    # name, dtype, cin, cout, n, (hin, win), (hout, wout), launches builder
    ("convT 64->32", torch.bfloat16, 64, 32, 2, (9, 13), (18, 26), lambda cg: cg.convT_fwd(3, 2, 1, 1, 9, 13)),
    ("convT 128->64", torch.bfloat16, 128, 64, 1, (16, 16), (32, 32), lambda cg: cg.convT_fwd(3, 2, 1, 1, 16, 16)),
    ("s2 dgrad 64->32 odd", torch.bfloat16, 64, 32, 2, (8, 11), (19, 25), lambda cg: cg.conv_dgrad(3, 2, 0, 19, 25)),
    ("s2 dgrad 128->64", torch.bfloat16, 128, 64, 1, (16, 8), (34, 18), lambda cg: cg.conv_dgrad(3, 2, 0, 34, 18)),
    ("rows 9 taps 32->32", torch.bfloat16, 32, 32, 2, (41, 20), (33, 20),
     lambda cg: [cg.Launch(33, 20, 1, 1, 0, 0, [(d, 0) for d in range(9)], [(d, 0) for d in range(9)], 0)]),
    ("rows 9 taps up 32->32", torch.bfloat16, 32, 32, 1, (20, 17), (28, 17),
     lambda cg: [cg.Launch(28, 17, 1, 1, 0, 0, [(-d, 0) for d in range(9)], [(d, 0) for d in range(9)], 0)]),
    ("rows 3 taps 16->64 tf32", torch.float32, 16, 64, 2, (21, 24), (21, 24),
     lambda cg: [cg.Launch(21, 24, 1, 1, 0, 0, [(-1, 0), (0, 0), (1, 0)], [(0, 0), (1, 0), (2, 0)], 0)]),
    ("rows 3 taps dx -1 64->32", torch.bfloat16, 64, 32, 1, (16, 30), (16, 32),
     lambda cg: [cg.Launch(16, 32, 1, 1, 0, 0, [(1, -1), (0, -1), (-1, -1)], [(0, 0), (1, 0), (2, 0)], 0)]),
    # filters that do not fit in shared memory: the streaming kernel (conv_hx.cu) takes the stacked launch
    ("rows 3x3 64->64 tf32 streamed", torch.float32, 64, 64, 2, (20, 24), (20, 24), lambda cg: cg.conv_fwd(3, 1, 1, 20, 24)),
    ("rows 3x3 64->64 bf16 streamed", torch.bfloat16, 64, 64, 1, (37, 19), (37, 19), lambda cg: cg.conv_dgrad(3, 1, 1, 37, 19)),
    ("rows tall 9 taps 32->32", torch.bfloat16, 32, 32, 1, (272, 8), (264, 8),
     lambda cg: [cg.Launch(264, 8, 1, 1, 0, 0, [(d, 0) for d in range(9)], [(d, 0) for d in range(9)], 0)]),
]


@pytest.mark.parametrize("name,dtype,cin,cout,n,hw_in,hw_out,build", STACKED_CASES, ids=[c[0] for c in STACKED_CASES])
@pytest.mark.parametrize("epi", ["plain", "stats", "bias_relu_mask_add"])
def test_conv_stacked_matches_the_plain_launches(env, name, dtype, cin, cout, n, hw_in, hw_out, build, epi):
    cg, ops = env
    from artist_style_transfer_b200 import _lib, arena
    torch.manual_seed(len(name) + cin)
    launches = build(cg)
    nt = sum(len(l.taps) for l in launches)
    x = tf32_round(torch.randn(n, *hw_in, cin, device="cuda")).to(dtype)
    wp = tf32_round(torch.randn(nt, cout, cin, device="cuda") / (cin * 4) ** 0.5).to(dtype)
    tidx = {wt: l.woff + t for l in launches for t, wt in enumerate(l.wtaps)}
    odt = torch.float32 if dtype == torch.float32 or epi == "plain" else torch.bfloat16
    kw = {}
    if epi == "bias_relu_mask_add":
        kw = dict(bias=torch.randn(cout, device="cuda"), relu=True,
                  mask=torch.randn(n, *hw_out, cout, device="cuda").to(torch.bfloat16),
                  add=torch.randn(n, *hw_out, cout, device="cuda").to(odt))
    ref = torch.full((n, *hw_out, cout), float("nan"), device="cuda", dtype=odt)
    got = torch.full_like(ref, float("nan"))
    s_ref = torch.zeros(n, cout, 2, dtype=torch.float64, device="cuda") if epi == "stats" else None
    s_got = torch.zeros_like(s_ref) if epi == "stats" else None
    ops.conv_gather(x, wp, launches, ref, tensor=True, stats=s_ref, **kw)
    before = _lib.family_stats()
    for stk in arena.stack_groups(launches, cout):
        ws = ops.stack_filter(lambda pos: wp[tidx[pos]], stk, cout, cin, dtype, "cuda")
        ops.conv_stacked(x, ws, stk, got, stats=s_got, **kw)
    delta = _lib.family_delta(before)
    fam = "conv_hx" if "streamed" in name else "conv_st"
    assert delta[fam][0] == (1 if len(launches) == 1 or cout == 32 else 2)
    flops = 2.0 * n * sum(l.mi * l.mj * len(l.taps) for l in launches) * cin * cout
    assert abs(delta[fam][1] - flops) <= 0.1 * flops      # interleaved rows round mi up to a multiple of nblk
    torch.cuda.synchronize()
    assert not torch.isnan(got.float()).any()
    # same operands, fp32 accumulation: only the summation order differs (one bf16 rounding when the output is bf16)
    tol = 1e-5 if odt == torch.float32 else 6e-3
    assert rel(got, ref) < tol, rel(got, ref)
    if odt == torch.bfloat16:
        assert float((got.float() - ref.float()).abs().max()) <= 2 ** -7 * float(ref.float().abs().max())
    if epi == "stats":
        assert rel(s_got, s_ref) < 1e-6


def test_conv_stacked_argument_errors(env):
    cg, ops = env
    stk = cg.stack_phases(cg.convT_fwd(3, 2, 1, 1, 8, 8))
    x = torch.zeros(1, 8, 8, 64, device="cuda", dtype=torch.bfloat16)
    w = torch.zeros(len(stk.vt), 128, 64, device="cuda", dtype=torch.bfloat16)
    with pytest.raises(RuntimeError, match="contiguous channels"):
        ops.conv_stacked(x, w, stk, torch.zeros(1, 16, 16, 64, device="cuda", dtype=torch.bfloat16))
    with pytest.raises(RuntimeError, match="64 bytes or a multiple of 128"):
        ops.conv_stacked(x[..., :48].contiguous(), w[..., :48].contiguous(), stk,
                         torch.zeros(1, 16, 16, 32, device="cuda", dtype=torch.bfloat16))


# ---------------------------------------------------------------------------------------------------------------------
# fp16 operands (AST_F16): the fast-mode VGG forward stores its non-tap activations as IEEE half - the 10 mantissa bits a
# kind::tf32 MMA keeps of an fp32 operand - and runs kind::f16.
F16_CASES = [
    # cin, cout, n, h, w, out dtype, expected kernel family
    (64, 64, 2, 16, 16, torch.float32, "conv_ws"),      # conv1_2 class: fp16 in, fp32 (tap) out
    (64, 128, 1, 16, 24, torch.float16, "conv_hx"),     # conv2_1 class
    (128, 128, 2, 16, 16, torch.float16, "conv_hx"),
    (256, 512, 1, 8, 8, torch.float16, "conv_hx"),
    (512, 512, 1, 8, 8, torch.float32, "conv_hx"),      # conv4_3 class: tap output
]


@pytest.mark.parametrize("cin,cout,n,h,w,odt,family", F16_CASES)
def test_conv_fp16_operands(env, cin, cout, n, h, w, odt, family):
    cg, ops = env
    from artist_style_transfer_b200 import _lib
    torch.manual_seed(cin + cout)
    x = torch.randn(n, h, w, cin, device="cuda").half()
    wt = (torch.randn(cout, cin, 3, 3, device="cuda") / (cin * 9) ** 0.5)
    bias = torch.randn(cout, device="cuda")
    launches = cg.conv_fwd(3, 1, 1, h, w)
    wp = ops.pack_weights(wt, launches, cout, cin, cin * 9, 9, 3, 1, torch.float32).half()
    y = torch.full((n, h, w, cout), float("nan"), device="cuda", dtype=odt)
    before = _lib.family_stats()
    ops.conv_gather(x, wp, launches, y, bias=bias, relu=True, tensor=True)
    assert _lib.family_delta(before)[family][0] == 1
    torch.cuda.synchronize()
    xr = x.double().cpu().permute(0, 3, 1, 2)
    wr = wp.double().cpu().view(3, 3, cout, cin).permute(2, 3, 0, 1)
    yr = F.relu(F.conv2d(xr, wr, bias.double().cpu(), padding=1)).permute(0, 2, 3, 1)
    assert not torch.isnan(y.float()).any()
    # identical fp16 operands, fp32 accumulation: only summation order (and one fp16 rounding of the result) differs
    assert rel(y, yr) < (1e-5 if odt == torch.float32 else 4e-4), rel(y, yr)


def test_fp16_stores_saturate_and_masks_read_fp16(env):
    cg, ops = env
    # a result beyond the fp16 range is stored as +-65504, never inf
    x = torch.full((1, 8, 8, 64), 200.0, device="cuda").half()
    wp = torch.full((9, 64, 64), 1.0, device="cuda").half()
    y = torch.empty(1, 8, 8, 64, device="cuda", dtype=torch.float16)
    ops.conv_gather(x, wp, cg.conv_fwd(3, 1, 1, 8, 8), y, tensor=True)
    torch.cuda.synchronize()
    assert torch.isfinite(y.float()).all() and float(y.float().max()) == 65504.0
    # max-pool writing fp16 (+ window codes) == torch on the same values
    a = torch.randn(2, 8, 12, 64, device="cuda")
    codes = torch.empty(2, 4, 6, 64, dtype=torch.uint8, device="cuda")
    p = ops.maxpool2_fwd(a, codes=codes, out_dtype=torch.float16)
    ref = F.max_pool2d(a.permute(0, 3, 1, 2), 2).permute(0, 2, 3, 1)
    assert p.dtype == torch.float16 and torch.equal(p, ref.half())
    # a bf16 data gradient masked by an fp16 activation == the same with the fp32 copy of that activation
    g = torch.randn(1, 16, 16, 128, device="cuda").bfloat16()
    act = torch.randn(1, 16, 16, 128, device="cuda")
    wd = (torch.randn(9, 128, 128, device="cuda") / 34).bfloat16()
    ls = cg.conv_dgrad(3, 1, 1, 16, 16)
    o16 = torch.empty(1, 16, 16, 128, device="cuda", dtype=torch.bfloat16)
    o32 = torch.empty_like(o16)
    ops.conv_gather(g, wd, ls, o16, mask=act.half(), tensor=True)
    ops.conv_gather(g, wd, ls, o32, mask=act.half().float(), tensor=True)
    torch.cuda.synchronize()
    assert torch.equal(o16, o32)
