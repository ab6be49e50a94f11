"""-m gpu: every non-tensor-core kernel of libnbasr against torch-fp32 / the numpy oracle (through the C ABI)."""
import ctypes as C
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from nb_asr_b200 import _lib  # noqa: E402
from nb_asr_b200._lib import BF16, F32, PAD_L, GConv  # noqa: E402
from nb_asr_b200.model import pad_rule  # noqa: E402
from oracle import decode_np as D  # noqa: E402
from oracle import model_ref as M  # noqa: E402
import gpu_utils as U  # noqa: E402


def _dense_conv_ref(x, w, b, stride):
    # x (B,T,Cin) ; w (Cout,Cin,8) reference layout
    return M.pad_conv_relu(x.permute(0, 2, 1), w, b, 8, 1, stride, 1).permute(0, 2, 1)


@pytest.mark.parametrize('stride', [1, 2])
def test_simt_dense_conv_as_gemm(stride):
    torch.manual_seed(0)
    B, T, Cin, Cout = 2, 37, 24, 40
    x = torch.randn(B, T, Cin)
    w = torch.randn(Cout, Cin, 8) * 0.2
    bias = torch.randn(Cout)
    skip = torch.randn(B, (T + stride - 1) // stride, Cout)
    ref = _dense_conv_ref(x, w, bias, stride) + skip
    To = ref.shape[1]
    xb = U.to_padded(x, F32)
    wp = w.permute(0, 2, 1).contiguous().view(Cout, 8 * Cin).to(U.DEV)       # (Cout, k, Cin)
    out = U.empty_padded(B, To, Cout, F32)
    sk = U.to_padded(skip, F32)
    mask = U.new_mask(out.shape[0], Cout)
    lpad, _ = pad_rule(8, 1, stride)
    epi = U.epilogue(F32, Cout, bias=bias.to(U.DEV), relu=1, adds=[sk], out=out, mask_out=mask)
    U.run_gemm(F32, U.ptr(xb, (PAD_L - lpad) * Cin), U.geo(T) * Cin, stride * Cin, B, To, 8 * Cin, Cout, wp, 8 * Cin,
               PAD_L, U.geo(To), 1, epi)
    got = U.from_padded(out, B, To).cpu()
    assert U.relerr(got, ref) < 1e-5
    z = F.conv1d(F.pad(x.permute(0, 2, 1), pad_rule(8, 1, stride)), w, bias, stride=stride).permute(0, 2, 1)
    m = U.unpack_mask(mask, B, To, Cout).cpu()
    assert torch.equal(m, (z > 0) & (z <= 20))
    # pad rows untouched
    assert float(out[:PAD_L].abs().sum()) == 0.0


def test_simt_wgrad_and_dgrad():
    torch.manual_seed(1)
    B, T, Cin, Cout = 2, 29, 16, 24
    x = torch.randn(B, T, Cin, requires_grad=True)
    w = (torch.randn(Cout, Cin, 8) * 0.2).requires_grad_(True)
    for stride in (1, 2):
        y = F.conv1d(F.pad(x.permute(0, 2, 1), pad_rule(8, 1, stride)), w, None, stride=stride).permute(0, 2, 1)
        To = y.shape[1]
        dy = torch.randn_like(y)
        gx, gw = torch.autograd.grad(y, (x, w), dy)
        xb, dyb = U.to_padded(x.detach(), F32), U.to_padded(dy, F32)
        lpad, _ = pad_rule(8, 1, stride)
        dw = torch.zeros(Cout, 8 * Cin, device=U.DEV)
        U.run_wgrad(F32, U.ptr(dyb, PAD_L * Cout), U.geo(To) * Cout, Cout, U.ptr(xb, (PAD_L - lpad) * Cin), U.geo(T) * Cin,
                    stride * Cin, B, To, Cout, 8 * Cin, dw, 8 * Cin)
        got_w = dw.view(Cout, 8, Cin).permute(0, 2, 1).cpu()
        assert U.relerr(got_w, gw) < 1e-5
        # dgrad through packed weights
        lib = _lib.load()
        wp = w.detach().permute(0, 2, 1).contiguous().to(U.DEV)       # (Cout, 8, Cin)
        dx = U.empty_padded(B, T, Cin, F32)
        specs = [(8, 7, -1, 4, 0, T, 1)] if stride == 1 else [(4, 7, -2, 1, 0, (T + 1) // 2, 2), (4, 6, -2, 0, 1, T // 2, 2)]
        for nq, t0, ts, back, par, nr, ors in specs:
            wd = torch.empty(Cin, nq * Cout, device=U.DEV)
            _lib.check(lib.nbasr_pack_weight(wp.data_ptr(), wd.data_ptr(), F32, Cout, Cin, nq, t0, ts, 8 * Cin, 1, Cin, U.stream()))
            epi = U.epilogue(F32, Cin, out=dx)
            U.run_gemm(F32, U.ptr(dyb, (PAD_L - back) * Cout), U.geo(To) * Cout, Cout, B, nr, nq * Cout, Cin, wd, nq * Cout,
                       PAD_L + par, U.geo(T), ors, epi)
        assert U.relerr(U.from_padded(dx, B, T).cpu(), gx) < 1e-5


@pytest.mark.parametrize('dt', [F32, BF16, 'mma'])
@pytest.mark.parametrize('op', ['conv5', 'conv5d2', 'conv7', 'conv7d2'])
@pytest.mark.parametrize('Cc', [600, 800, 1000, 1200])
def test_gconv_fwd_bwd(dt, op, Cc):
    """F32 / BF16: SIMT kernels.  'mma': bf16 tcgen05 block-diagonal kernels (the product path in bf16 mode)."""
    _gconv_fwd_bwd(dt, op, Cc)


def _gconv_fwd_bwd(dt, op, Cc):
    if dt == BF16 and Cc not in (600, 1000):
        pytest.skip('bf16 SIMT variant sampled on two widths')
    mma = dt == 'mma'
    if mma:
        dt = BF16
    torch.manual_seed(2)
    k, d = M.CONV_EDGE[op]
    B, T, cpg = (3, 300, Cc // 100) if mma else (2, 70, Cc // 100)
    rnd = (lambda t: t.bfloat16().float()) if dt == BF16 else (lambda t: t)
    x = rnd(torch.randn(B, T, Cc)).requires_grad_(True)
    w = rnd(torch.randn(Cc, cpg, k) * 0.3).requires_grad_(True) if mma else (torch.randn(Cc, cpg, k) * 0.3).requires_grad_(True)
    bias = torch.randn(Cc) * 0.1
    skip = rnd(torch.randn(B, T, Cc))
    lp, rp = pad_rule(k, d, 1)
    z = F.conv1d(F.pad(x.permute(0, 2, 1), (lp, rp)), w, bias, dilation=d, groups=100).permute(0, 2, 1)
    ref = M.relu20(z) + skip
    lib = _lib.load()
    xb, sk = U.to_padded(x.detach(), dt), U.to_padded(skip, dt)
    out = U.empty_padded(B, T, Cc, dt)
    mwid = (40 if cpg == 10 else 48) if mma else 32        # mask plane width = the producing kernel's slab
    mask = U.new_mask(out.shape[0], Cc, mwid)
    wg, bg = w.detach().to(U.DEV), bias.to(U.DEV)
    gc = GConv()
    gc.dtype, gc.x, gc.B, gc.T, gc.Tp, gc.C, gc.cpg, gc.ktaps, gc.off0, gc.dstep = dt, xb.data_ptr(), B, T, U.geo(T), Cc, cpg, k, -lp, d
    gc.w = wg.data_ptr()
    if mma:
        ne = int(lib.nbasr_gconv_mma_pack_elems(Cc, cpg, k))
        wpk = torch.zeros(ne, dtype=torch.bfloat16, device=U.DEV)
        wpk_t = torch.zeros(ne, dtype=torch.bfloat16, device=U.DEV)
        _lib.check(lib.nbasr_pack_gconv_mma(wg.data_ptr(), wpk.data_ptr(), BF16, Cc, cpg, k, 0, U.stream()))
        _lib.check(lib.nbasr_pack_gconv_mma(wg.data_ptr(), wpk_t.data_ptr(), BF16, Cc, cpg, k, 1, U.stream()))
        gc.w, gc.w_packed = wpk.data_ptr(), 1
    gc.epi = U.epilogue(dt, Cc, bias=bg, relu=1, adds=[sk], out=out, mask_out=mask, mask_w=mwid)
    _lib.check(lib.nbasr_gconv_fwd(C.byref(gc), U.stream()), 'gconv')
    torch.cuda.synchronize()
    tol = 1e-5 if dt == F32 else 6e-3
    assert U.relerr(U.from_padded(out, B, T).cpu(), ref) < tol
    if dt == F32:
        assert torch.equal(U.unpack_mask(mask, B, T, Cc).cpu(), (z > 0) & (z <= 20))
    if mma:
        assert (U.unpack_mask(mask, B, T, Cc, mwid).cpu() != ((z > 0) & (z <= 20))).float().mean() < 1e-3
    # backward: dz given
    dz = rnd(torch.randn(B, T, Cc))
    gx, gw = torch.autograd.grad(z, (x, w), dz)
    dzb = U.to_padded(dz, dt)
    dw = torch.zeros(Cc, cpg, k, device=U.DEV)
    dbf = torch.zeros(Cc, device=U.DEV)
    _lib.check(lib.nbasr_gconv_wgrad(dt, dzb.data_ptr(), xb.data_ptr(), B, T, U.geo(T), Cc, cpg, k, -lp, d, dw.data_ptr(),
                                     dbf.data_ptr(), U.stream()))
    wt = torch.empty_like(wg)
    _lib.check(lib.nbasr_pack_gconv_dgrad(wg.data_ptr(), wt.data_ptr(), Cc, cpg, k, U.stream()))
    dx = U.empty_padded(B, T, Cc, dt)
    gc2 = GConv()
    gc2.dtype, gc2.x, gc2.B, gc2.T, gc2.Tp, gc2.C, gc2.cpg, gc2.ktaps, gc2.dstep = dt, dzb.data_ptr(), B, T, U.geo(T), Cc, cpg, k, d
    gc2.off0 = lp - (k - 1) * d
    gc2.w = wt.data_ptr()
    if mma:
        gc2.w, gc2.w_packed = wpk_t.data_ptr(), 1
    gc2.epi = U.epilogue(dt, Cc, out=dx)
    _lib.check(lib.nbasr_gconv_fwd(C.byref(gc2), U.stream()), 'gconv dgrad')
    torch.cuda.synchronize()
    assert U.relerr(dw.cpu(), gw) < (1e-4 if dt == F32 else (2e-5 if mma else 1e-2))      # mma: exact bf16 inputs, fp32 accumulate
    assert float(out[:PAD_L].abs().sum()) == 0.0 and float(dx[:PAD_L].abs().sum()) == 0.0
    assert U.relerr(U.from_padded(dx, B, T).cpu(), gx) < tol
    db = torch.zeros(Cc, device=U.DEV)
    _lib.check(lib.nbasr_colsum(dt, dzb.data_ptr(), B, T, U.geo(T), Cc, db.data_ptr(), U.stream()))
    assert U.relerr(db.cpu(), dz.sum((0, 1))) < 1e-4
    assert U.relerr(dbf.cpu(), dz.sum((0, 1))) < 1e-4       # bias gradient fused into the weight-gradient kernel


@pytest.mark.parametrize('dt', [F32, BF16])
@pytest.mark.parametrize('Cc', [600, 1200])
def test_layernorm_fwd_bwd(dt, Cc):
    torch.manual_seed(3)
    B, T = 3, 41
    rnd = (lambda t: t.bfloat16().float()) if dt == BF16 else (lambda t: t)
    x = rnd(torch.randn(B, T, Cc) * 2 + 0.5).requires_grad_(True)
    g = (torch.randn(Cc) * 0.2 + 1).requires_grad_(True)
    b = (torch.randn(Cc) * 0.1).requires_grad_(True)
    y = F.layer_norm(x, (Cc,), g, b, 1e-3)
    dy = rnd(torch.randn_like(y))
    gx, gg, gb = torch.autograd.grad(y, (x, g, b), dy)
    lib = _lib.load()
    xb, yb = U.to_padded(x.detach(), dt), U.empty_padded(B, T, Cc, dt)
    mean = torch.zeros(xb.shape[0], device=U.DEV)
    rstd = torch.zeros(xb.shape[0], device=U.DEV)
    gd, bd = g.detach().to(U.DEV), b.detach().to(U.DEV)
    _lib.check(lib.nbasr_layernorm_fwd(dt, xb.data_ptr(), yb.data_ptr(), B, T, U.geo(T), Cc, gd.data_ptr(), bd.data_ptr(), 1e-3,
                                       mean.data_ptr(), rstd.data_ptr(), 1.0, None, U.stream()))
    tol = 2e-6 if dt == F32 else 4e-3
    assert U.relerr(U.from_padded(yb, B, T).cpu(), y.detach()) < tol
    dyb, dxb = U.to_padded(dy, dt), U.empty_padded(B, T, Cc, dt)
    dg, db = torch.zeros(Cc, device=U.DEV), torch.zeros(Cc, device=U.DEV)
    _lib.check(lib.nbasr_layernorm_bwd(dt, dyb.data_ptr(), xb.data_ptr(), dt, 1.0, mean.data_ptr(), rstd.data_ptr(), gd.data_ptr(), B, T,
                                       U.geo(T), Cc, dxb.data_ptr(), None, None, 1.0, 0, 32, dg.data_ptr(), db.data_ptr(), U.stream()))
    torch.cuda.synchronize()
    assert U.relerr(U.from_padded(dxb, B, T).cpu(), gx) < (1e-4 if dt == F32 else 8e-3)
    assert U.relerr(dg.cpu(), gg) < 1e-4 and U.relerr(db.cpu(), gb) < 1e-4


def test_layernorm_degenerate_rows():
    # all-zero input (all-`zero` no-skip cell): var = 0 -> y = beta, no NaN
    B, T, Cc = 1, 5, 600
    lib = _lib.load()
    xb, yb = U.empty_padded(B, T, Cc, F32), U.empty_padded(B, T, Cc, F32)
    g = torch.ones(Cc, device=U.DEV)
    b = torch.full((Cc,), 0.25, device=U.DEV)
    _lib.check(lib.nbasr_layernorm_fwd(F32, xb.data_ptr(), yb.data_ptr(), B, T, U.geo(T), Cc, g.data_ptr(), b.data_ptr(), 1e-3,
                                       None, None, 1.0, None, U.stream()))
    assert torch.allclose(U.from_padded(yb, B, T), torch.full((B, T, Cc), 0.25, device=U.DEV))


@pytest.mark.parametrize('Cc', [600, 800, 1000, 1200])
def test_layernorm_scaled_fp16_activations(Cc):
    """16-bit mode: x is fp16 holding S*x_true.  With eps*S^2 the normalised value is that of the unscaled tensor; the output is
    written scaled (fp16) plus an unscaled bf16 twin; the backward pass (bf16 gradients) returns d/dx_true."""
    from nb_asr_b200._lib import F16
    S = 32.0
    torch.manual_seed(7)
    B, T = 3, 70
    x = ((torch.randn(B, T, Cc) * 2 + 0.5) * S).half().float() / S             # exactly representable as scaled fp16
    x.requires_grad_(True)
    g = (torch.randn(Cc) * 0.2 + 1).requires_grad_(True)
    b = (torch.randn(Cc) * 0.1).requires_grad_(True)
    y = F.layer_norm(x, (Cc,), g, b, 1e-3)
    dy = torch.randn_like(y).bfloat16().float()
    gx, gg, gb = torch.autograd.grad(y, (x, g, b), dy)
    lib = _lib.load()
    xb = U.to_padded(x.detach() * S, F16)
    yb, y2 = U.empty_padded(B, T, Cc, F16), U.empty_padded(B, T, Cc, BF16)
    mean, rstd = torch.zeros(xb.shape[0], device=U.DEV), torch.zeros(xb.shape[0], device=U.DEV)
    gd, bd = g.detach().to(U.DEV), b.detach().to(U.DEV)
    _lib.check(lib.nbasr_layernorm_fwd(F16, xb.data_ptr(), yb.data_ptr(), B, T, U.geo(T), Cc, gd.data_ptr(), bd.data_ptr(), 1e-3 * S * S,
                                       mean.data_ptr(), rstd.data_ptr(), S, y2.data_ptr(), U.stream()))
    assert U.relerr(U.from_padded(yb, B, T).cpu() / S, y.detach()) < 6e-4         # fp16 output rounding (2^-11)
    assert U.relerr(U.from_padded(y2, B, T).cpu(), y.detach()) < 4e-3            # bf16 twin
    dyb, dxb = U.to_padded(dy, BF16), U.empty_padded(B, T, Cc, BF16)
    dg, db = torch.zeros(Cc, device=U.DEV), torch.zeros(Cc, device=U.DEV)
    _lib.check(lib.nbasr_layernorm_bwd(BF16, dyb.data_ptr(), xb.data_ptr(), F16, S, mean.data_ptr(), rstd.data_ptr(), gd.data_ptr(), B, T,
                                       U.geo(T), Cc, dxb.data_ptr(), None, None, 1.0, 0, 32, dg.data_ptr(), db.data_ptr(), U.stream()))
    torch.cuda.synchronize()
    assert U.relerr(U.from_padded(dxb, B, T).cpu(), gx) < 8e-3
    assert U.relerr(dg.cpu(), gg) < 1e-4 and U.relerr(db.cpu(), gb) < 1e-4


@pytest.mark.parametrize('op', ['conv5', 'conv7d2'])
@pytest.mark.parametrize('Cc', [600, 1000])
def test_gconv_scaled_fp16_forward(op, Cc):
    """16-bit mode forward edge: fp16 activations holding S*x, fp16 block-diagonal weights, epilogue in the scaled domain
    (bias * S, ReLU bound 20 * S, scaled skip tensor), scaled fp16 output plus the unscaled bf16 twin (out2)."""
    from nb_asr_b200._lib import F16
    S = 32.0
    torch.manual_seed(2)
    k, d = M.CONV_EDGE[op]
    B, T, cpg = 3, 300, Cc // 100
    h = lambda t: t.half().float()
    x = h(torch.randn(B, T, Cc) * S) / S
    w = h(torch.randn(Cc, cpg, k) * 0.3)
    bias = torch.randn(Cc) * 0.1
    skip = h(torch.randn(B, T, Cc) * S) / S
    lp, rp = pad_rule(k, d, 1)
    z = F.conv1d(F.pad(x.permute(0, 2, 1), (lp, rp)), w, bias, dilation=d, groups=100).permute(0, 2, 1)
    ref = M.relu20(z) + skip
    lib = _lib.load()
    xb, sk = U.to_padded(x * S, F16), U.to_padded(skip * S, F16)
    out, out2 = U.empty_padded(B, T, Cc, F16), U.empty_padded(B, T, Cc, BF16)
    mwid = 40 if cpg == 10 else 48
    mask = U.new_mask(out.shape[0], Cc, mwid)
    wg, bg = w.to(U.DEV), bias.to(U.DEV)
    ne = int(lib.nbasr_gconv_mma_pack_elems(Cc, cpg, k))
    wpk = torch.zeros(ne, dtype=torch.float16, device=U.DEV)
    _lib.check(lib.nbasr_pack_gconv_mma(wg.data_ptr(), wpk.data_ptr(), F16, Cc, cpg, k, 0, U.stream()))
    for with_skip in (True, False):           # general epilogue path / lean path
        gc = GConv()
        gc.dtype, gc.x, gc.B, gc.T, gc.Tp, gc.C, gc.cpg, gc.ktaps, gc.off0, gc.dstep = F16, xb.data_ptr(), B, T, U.geo(T), Cc, cpg, k, -lp, d
        gc.w, gc.w_packed = wpk.data_ptr(), 1
        gc.epi = U.epilogue(F16, Cc, bias=bg, relu=1, adds=[sk] if with_skip else [], out=out, mask_out=mask, mask_w=mwid, out2=out2,
                            out2_dtype=BF16, scale2=1.0 / S, bias_scale=S, relu_hi=20.0 * S)
        _lib.check(lib.nbasr_gconv_fwd(C.byref(gc), U.stream()), 'gconv f16')
        torch.cuda.synchronize()
        r = ref if with_skip else M.relu20(z)
        assert U.relerr(U.from_padded(out, B, T).cpu() / S, r) < 6e-4
        assert U.relerr(U.from_padded(out2, B, T).cpu(), r) < 4e-3
        assert (U.unpack_mask(mask, B, T, Cc, mwid).cpu() != ((z > 0) & (z <= 20))).float().mean() < 1e-3
        assert float(out[:PAD_L].abs().sum()) == 0.0


@pytest.mark.parametrize('B,T', [(3, 9), (20, 6), (64, 4)])
def test_lstm_fwd_bwd(B, T):
    torch.manual_seed(4)
    H, I = 500, 64
    x = torch.randn(B, T, I)
    w_ih = (torch.randn(4 * H, I) * 0.2).requires_grad_(True)
    w_hh = (torch.randn(4 * H, H) * 0.08).requires_grad_(True)
    b_ih = torch.randn(4 * H) * 0.1
    b_hh = torch.randn(4 * H) * 0.1
    gx = (x @ w_ih.t() + b_ih + b_hh)
    gx_leaf = gx.detach().clone().requires_grad_(True)
    # reference recurrence on the precomputed projection
    h = torch.zeros(B, H)
    c = torch.zeros(B, H)
    outs = []
    for t in range(T):
        g = gx_leaf[:, t] + h @ w_hh.t()
        i, f, gg, o = g.split(H, 1)
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
        h = torch.sigmoid(o) * torch.tanh(c)
        outs.append(h)
    hs = torch.stack(outs, 1)
    dh = torch.randn_like(hs)
    ggx, gwhh = torch.autograd.grad(hs, (gx_leaf, w_hh), dh)
    lib = _lib.load()
    gxd = gx.detach().contiguous().to(U.DEV)
    whh = w_hh.detach().to(U.DEV)
    hseq = torch.zeros(B, T, 512, device=U.DEV)
    gates = torch.zeros(B * T, 4 * H, device=U.DEV)
    cst = torch.zeros(B * T, H, device=U.DEV)
    work = torch.zeros(2 * B * H + 256, device=U.DEV)
    _lib.check(lib.nbasr_lstm_fwd(gxd.data_ptr(), whh.data_ptr(), T, B, H, hseq.data_ptr(), F32, T * 512, 512, 512, gates.data_ptr(),
                                  cst.data_ptr(), None, work.data_ptr(), U.stream()), 'lstm_fwd')
    torch.cuda.synchronize()
    assert U.relerr(hseq[:, :, :H].cpu(), hs.detach()) < 2e-5
    dhd = dh.contiguous().to(U.DEV)
    dgx = torch.zeros(B * T, 4 * H, device=U.DEV)
    _lib.check(lib.nbasr_lstm_bwd(dhd.data_ptr(), T * H, H, H, whh.data_ptr(), gates.data_ptr(), cst.data_ptr(), T, B, H, dgx.data_ptr(),
                                  work.data_ptr(), None, None, U.stream()), 'lstm_bwd')
    torch.cuda.synchronize()
    assert U.relerr(dgx.view(B, T, 4 * H).cpu(), ggx) < 5e-5
    # dW_hh = dgx^T h_{t-1} via the SIMT wgrad over a row-shifted view of h_seq
    hp = torch.zeros(B, T + 1, 512, device=U.DEV)
    hp[:, 1:] = hseq
    dw = torch.zeros(4 * H, H, device=U.DEV)
    U.run_wgrad(F32, dgx.data_ptr(), T * 4 * H, 4 * H, hp.data_ptr(), (T + 1) * 512, 512, B, T, 4 * H, H, dw, H)
    assert U.relerr(dw.cpu(), gwhh) < 5e-5


def test_head_ctc_greedy_per():
    torch.manual_seed(5)
    B, T, K, V, S = 5, 23, 500, 49, 9
    h = torch.randn(B, T, K) * 0.5
    w = (torch.randn(V, K) * 0.1).requires_grad_(True)
    bias = (torch.randn(V) * 0.1).requires_grad_(True)
    hq = h.clone().requires_grad_(True)
    logits = hq @ w.t() + bias
    logp = F.log_softmax(logits, 2)
    alen = torch.tensor([92, 60, 95, 20, 77])          # // 4 -> 23, 15, 23, 5, 19
    tl = torch.tensor([9, 4, 7, 8, 5])                  # utt 3: 8 labels in 5 frames -> infeasible
    tg = torch.randint(1, V, (B, S), dtype=torch.int32)
    for b in range(B):
        tg[b, int(tl[b]):] = 0
    out_len = alen // 4
    loss = M.ctc_loss_ref(logp, out_len, tg, tl)
    gh, gw, gb = torch.autograd.grad(loss, (hq, w, bias))
    lib = _lib.load()
    hd, wd, bd = h.to(U.DEV), w.detach().to(U.DEV), bias.detach().to(U.DEV)
    lg = torch.zeros(B, T, V, device=U.DEV)
    lp = torch.zeros(B, T, V, device=U.DEV)
    _lib.check(lib.nbasr_head_fwd(F32, hd.data_ptr(), T * K, K, B, T, K, V, wd.data_ptr(), bd.data_ptr(), lg.data_ptr(), lp.data_ptr(), U.stream()))
    torch.cuda.synchronize()
    assert U.relerr(lg.cpu(), logits.detach()) < 1e-5 and U.relerr(lp.cpu(), logp.detach()) < 1e-5
    tgd, alend, tld = tg.to(U.DEV), alen.to(U.DEV), tl.to(U.DEV)     # keep device copies alive across the calls
    nll = torch.zeros(B, device=U.DEV)
    lossd = torch.zeros(1, device=U.DEV)
    dl = torch.zeros(B, T, V, device=U.DEV)
    work = torch.zeros(2 * B * T * (2 * S + 1) + 16, device=U.DEV)
    _lib.check(lib.nbasr_ctc(lp.data_ptr(), B, T, V, tgd.data_ptr(), S, alend.data_ptr(), 4, tld.data_ptr(),
                             nll.data_ptr(), lossd.data_ptr(), dl.data_ptr(), work.data_ptr(), U.stream()))
    torch.cuda.synchronize()
    assert abs(lossd.item() - loss.item()) < 1e-5 * abs(loss.item())
    assert nll[3].item() == 0.0                          # zero_infinity
    ref64 = D.ctc_nll(logp.detach().numpy(), out_len.numpy(), tg.numpy(), tl.numpy())
    assert np.allclose(nll.cpu().numpy()[[0, 1, 2, 4]], ref64[[0, 1, 2, 4]], rtol=1e-5)
    # backward through head
    dh = torch.zeros(B, T, K, device=U.DEV)
    dw = torch.zeros(V, K, device=U.DEV)
    db = torch.zeros(V, device=U.DEV)
    _lib.check(lib.nbasr_head_bwd(F32, hd.data_ptr(), T * K, K, B, T, K, V, wd.data_ptr(), dl.data_ptr(), dh.data_ptr(), T * K, K,
                                  dw.data_ptr(), db.data_ptr(), U.stream()))
    torch.cuda.synchronize()
    assert U.relerr(dh.cpu(), gh) < 2e-4 and U.relerr(dw.cpu(), gw) < 2e-4 and U.relerr(db.cpu(), gb) < 2e-4
    # greedy + fold + PER: bit exact against the numpy oracle on the SAME log-probs
    lut = torch.as_tensor(D.FOLD_LUT, dtype=torch.int32, device=U.DEV)
    hyp = torch.zeros(B, T, dtype=torch.int32, device=U.DEV)
    hl = torch.zeros(B, dtype=torch.int32, device=U.DEV)
    dist = torch.zeros(B, dtype=torch.int32, device=U.DEV)
    per = torch.zeros(2, dtype=torch.float64, device=U.DEV)
    iw = torch.zeros(B * (S + 2) + 16, dtype=torch.int32, device=U.DEV)
    _lib.check(lib.nbasr_greedy_per(lp.data_ptr(), B, T, V, alend.data_ptr(), 4, tgd.data_ptr(), S,
                                    tld.data_ptr(), lut.data_ptr(), hyp.data_ptr(), hl.data_ptr(), dist.data_ptr(),
                                    per.data_ptr(), iw.data_ptr(), U.stream()))
    torch.cuda.synchronize()
    rper, rd, _, rh = D.per_batch(lp.cpu().numpy(), out_len.numpy(), tg.numpy(), tl.numpy())
    assert dist.cpu().tolist() == rd.tolist()
    assert per[0].item() == rper
    for b in range(B):
        assert hyp[b, :int(hl[b])].cpu().tolist() == rh[b].tolist()


@pytest.mark.parametrize('T,S', [(64, 20), (700, 270)])
def test_ctc_loss_and_gradient_long_targets(T, S):
    """CTC alone against the torch reference formula (trainer.py:36-44): the alpha and beta recursions run concurrently in the two
    halves of the CTA; S = 270 (2S+1 = 541 > 512 states: cfg 5's 300-label targets) exercises the several-states-per-thread path,
    S = 20 the one-state-per-thread path with software-pipelined emissions.  Mixed lengths, one empty and one infeasible target."""
    torch.manual_seed(11)
    B, V = 4, 49
    logits = (torch.randn(B, T, V) * 2).double().requires_grad_(True)      # fp64 reference
    logp = F.log_softmax(logits, 2)
    alen = torch.tensor([4 * T, 4 * (T - 3), 4 * (T // 2) + 1, 4 * max(2, S // 4)])       # last: fewer frames than labels
    tl = torch.tensor([S, S - 3, 0, S])
    tg = torch.randint(1, V, (B, S), dtype=torch.int32)
    tg[0, 1::2] = tg[0, 0:-1:2][: tg[0, 1::2].numel()]        # many repeated labels (need blanks between them)
    for b in range(B):
        tg[b, int(tl[b]):] = 0
    out_len = alen // 4
    loss = M.ctc_loss_ref(logp, out_len, tg, tl)
    gl, = torch.autograd.grad(loss, logits)
    lib = _lib.load()
    lp = logp.detach().float().to(U.DEV).contiguous()
    tgd, alend, tld = tg.to(U.DEV), alen.to(U.DEV), tl.to(U.DEV)
    nll = torch.zeros(B, device=U.DEV)
    lossd = torch.zeros(1, device=U.DEV)
    dl = torch.full((B, T, V), 7.0, device=U.DEV)
    work = torch.zeros(2 * B * T * (2 * S + 1) + 16, device=U.DEV)
    _lib.check(lib.nbasr_ctc(lp.data_ptr(), B, T, V, tgd.data_ptr(), S, alend.data_ptr(), 4, tld.data_ptr(),
                             nll.data_ptr(), lossd.data_ptr(), dl.data_ptr(), work.data_ptr(), U.stream()))
    torch.cuda.synchronize()
    assert abs(lossd.item() - loss.item()) < 2e-5 * abs(loss.item())
    ref64 = D.ctc_nll(logp.detach().float().numpy(), out_len.numpy(), tg.numpy(), tl.numpy())
    feas = np.isfinite(ref64)
    assert not feas[3] and nll[3].item() == 0.0 and feas[:3].all()         # zero_infinity on the infeasible utterance
    assert np.allclose(nll.cpu().numpy()[feas], ref64[feas], rtol=2e-5)
    assert float(dl[3].abs().max()) == 0.0
    # fp32 log-space recursions against the fp64 reference: the error grows with the number of frames (torch's own fp32 CTC is
    # 2e-5 / 4e-4 off the fp64 one at these two sizes; this kernel measures 3e-5 / 1.3e-3, tools/ctc_err.py)
    assert U.relerr(dl.cpu().double(), gl) < (1e-4 if T <= 64 else 3e-3)
    # loss only (eval): no gradient buffer, beta half idle
    lossd.zero_()
    _lib.check(lib.nbasr_ctc(lp.data_ptr(), B, T, V, tgd.data_ptr(), S, alend.data_ptr(), 4, tld.data_ptr(),
                             nll.data_ptr(), lossd.data_ptr(), None, work.data_ptr(), U.stream()))
    torch.cuda.synchronize()
    assert abs(lossd.item() - loss.item()) < 2e-5 * abs(loss.item())


def test_levenshtein_edge_cases():
    lib = _lib.load()
    V = 49
    cases = [([], [1, 2]), ([11, 9, 20, 20, 5, 14], [19, 9, 20, 20, 9, 14, 7]), ([3, 4, 5], [3, 4, 5]), ([7] * 1, [8] * 12)]
    B, T, S = len(cases), 16, 12
    lp = torch.full((B, T, V), -10.0)
    alen = torch.zeros(B, dtype=torch.int64)
    tg = torch.zeros(B, S, dtype=torch.int32)
    tl = torch.zeros(B, dtype=torch.int64)
    for b, (h, r) in enumerate(cases):
        # frames: label, blank, label, blank ... so greedy reproduces h exactly (incl. repeats)
        seq = []
        for s in h:
            seq += [s, 0]
        for t, s in enumerate(seq):
            lp[b, t, s] = 0.0
        alen[b] = max(len(seq), 1)
        if not seq:
            lp[b, 0, 0] = 0.0
        tg[b, :len(r)] = torch.tensor(r, dtype=torch.int32)
        tl[b] = len(r)
    d = [t.to(U.DEV) for t in (lp, alen, tg, tl)]
    hyp = torch.zeros(B, T, dtype=torch.int32, device=U.DEV)
    hl = torch.zeros(B, dtype=torch.int32, device=U.DEV)
    dist = torch.zeros(B, dtype=torch.int32, device=U.DEV)
    per = torch.zeros(2, dtype=torch.float64, device=U.DEV)
    iw = torch.zeros(64, dtype=torch.int32, device=U.DEV)
    _lib.check(lib.nbasr_greedy_per(d[0].data_ptr(), B, T, V, d[1].data_ptr(), 1, d[2].data_ptr(), S, d[3].data_ptr(), None,
                                    hyp.data_ptr(), hl.data_ptr(), dist.data_ptr(), per.data_ptr(), iw.data_ptr(), U.stream()))
    torch.cuda.synchronize()
    assert dist.cpu().tolist() == [2, 3, 0, 12]
    assert dist.cpu().tolist() == [D.levenshtein(h, r) for h, r in cases]


def test_optim_step_matches_reference_formulas():
    torch.manual_seed(6)
    n = 5000
    p = torch.randn(n)
    g = torch.randn(n) * 3
    segs = [(0, 1000), (1024, 2000)]
    lib = _lib.load()
    pd, gd = p.clone().to(U.DEV), g.clone().to(U.DEV)
    m, v = torch.zeros(n, device=U.DEV), torch.zeros(n, device=U.DEV)
    so = torch.tensor([s[0] for s in segs], dtype=torch.int64, device=U.DEV)
    sl = torch.tensor([s[1] for s in segs], dtype=torch.int64, device=U.DEV)
    state = torch.zeros(8 + 2 + 592 + 2, device=U.DEV)      # header + 2 segments + scratch of the deterministic reductions
    state[1] = 1e-3
    pr, mr, vr = p.clone(), torch.zeros(n), torch.zeros(n)
    for step in range(1, 4):
        gr = g.clone() * step
        gd.copy_(gr)
        for o, l in segs:
            gr[o:o + l] += 0.01 * pr[o:o + l] / pr[o:o + l].norm()
        coef = min(1.0, 5.0 / (float(gr.norm()) + 1e-6))
        gr = gr * coef
        mr = 0.9 * mr + 0.1 * gr
        vr = 0.999 * vr + 0.001 * gr * gr
        pr = pr - (1e-3 / (1 - 0.9 ** step)) * mr / (vr.sqrt() / math.sqrt(1 - 0.999 ** step) + 1e-7)
        _lib.check(lib.nbasr_optim_step(pd.data_ptr(), gd.data_ptr(), m.data_ptr(), v.data_ptr(), n, so.data_ptr(), sl.data_ptr(), 2,
                                        sum((l + 16383) // 16384 for _, l in segs), 0.01, 5.0, 0.9, 0.999, 1e-7, state.data_ptr(),
                                        U.stream()))
        torch.cuda.synchronize()
        assert U.relerr(pd.cpu(), pr) < 1e-6
        assert abs(state[3].item() - coef) < 1e-5 * coef
    assert state[0].item() == 3.0


def test_dropout_statistics_and_mask_consistency():
    B, T, Cc, p = 2, 50, 608, 0.2
    lib = _lib.load()
    src = U.to_padded(torch.ones(B, T, Cc), F32)
    out = U.empty_padded(B, T, Cc, F32)
    mask = U.new_mask(out.shape[0], Cc)
    epi = U.epilogue(F32, Cc, drop_p=p, salt=1234, out=out, mask_out=mask)
    _lib.check(lib.nbasr_eltwise(F32, src.data_ptr(), Cc, B, T, U.geo(T), Cc, C.byref(epi), U.stream()))
    torch.cuda.synchronize()
    o = U.from_padded(out, B, T)
    keep = (o != 0)
    assert abs(keep.float().mean().item() - (1 - p)) < 0.01
    assert torch.allclose(o[keep], torch.full_like(o[keep], 1 / (1 - p)))
    assert torch.equal(U.unpack_mask(mask, B, T, Cc), keep)


@pytest.mark.parametrize('B,T', [(16, 7), (20, 12), (64, 5)])
def test_lstm_cluster_tensor_core_forward(B, T):
    """bf16 mode: 16-CTA cluster kernel (tcgen05 MMA, distributed-shared-memory h exchange) vs fp32 reference."""
    import ctypes as C_
    torch.manual_seed(7)
    H = 500
    gx = torch.randn(B, T, 4 * H) * 0.7
    w_hh = torch.randn(4 * H, H) * 0.08
    w_bf = w_hh.bfloat16().float()
    # reference with the same operand rounding (bf16 W_hh, bf16 h fed back), fp32 state
    h = torch.zeros(B, H)
    c = torch.zeros(B, H)
    outs = []
    for t in range(T):
        g = gx[:, t] + h.bfloat16().float() @ w_bf.t()
        i, f, gg, o = g.split(H, 1)
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
        h = torch.sigmoid(o) * torch.tanh(c)
        outs.append(h)
    hs = torch.stack(outs, 1)
    lib = _lib.load()
    whh = w_hh.to(U.DEV)
    wp = torch.zeros(16 * 128 * 512, dtype=torch.bfloat16, device=U.DEV)
    job = (_lib.PackJob * 1)()
    job[0].kind, job[0].out_dtype, job[0].src, job[0].dst, job[0].n_out = 4, BF16, whh.data_ptr(), wp.data_ptr(), 16 * 128 * 512
    job[0].a[0] = H
    jd = torch.frombuffer(bytearray(bytes(job)), dtype=torch.uint8).to(U.DEV)
    nblk = (16 * 128 * 512 + 4095) // 4096
    bmap = torch.stack([torch.zeros(nblk, dtype=torch.int32), torch.arange(nblk, dtype=torch.int32)], 1).contiguous().to(U.DEV)
    _lib.check(lib.nbasr_pack_batch(jd.data_ptr(), 1, bmap.data_ptr(), nblk, U.stream()))
    gxd = gx.contiguous().to(U.DEV)
    hseq = torch.zeros(B, T, 512, dtype=torch.bfloat16, device=U.DEV)
    gates = torch.zeros(B * T, 4 * H, device=U.DEV)
    cst = torch.zeros(B * T, H, device=U.DEV)
    work = torch.zeros(2 * B * H + 256, device=U.DEV)
    _lib.check(lib.nbasr_lstm_fwd(gxd.data_ptr(), whh.data_ptr(), T, B, H, hseq.data_ptr(), BF16, T * 512, 512, 512, gates.data_ptr(),
                                  cst.data_ptr(), wp.data_ptr(), work.data_ptr(), U.stream()), 'lstm_fwd cluster')
    torch.cuda.synchronize()
    assert U.relerr(hseq[:, :, :H].float().cpu(), hs) < 6e-3          # bf16 output rounding
    assert U.relerr(cst.view(B, T, H)[:, -1].cpu(), c) < 2e-3
    assert float(hseq[:, :, H:].float().abs().sum()) == 0.0
    # backward through the cluster kernel vs the fp32 SIMT kernel on the SAME saved gates / cell states
    dh = torch.randn(B, T, H, device=U.DEV)
    dgx_ref = torch.zeros(B * T, 4 * H, device=U.DEV)
    _lib.check(lib.nbasr_lstm_bwd(dh.data_ptr(), T * H, H, H, whh.data_ptr(), gates.data_ptr(), cst.data_ptr(), T, B, H,
                                  dgx_ref.data_ptr(), work.data_ptr(), None, None, U.stream()), 'lstm_bwd simt')
    dgx = torch.zeros(B * T, 4 * H, device=U.DEV)
    dgx16 = torch.zeros(B * T, 4 * H, dtype=torch.bfloat16, device=U.DEV)
    _lib.check(lib.nbasr_lstm_bwd(dh.data_ptr(), T * H, H, H, whh.data_ptr(), gates.data_ptr(), cst.data_ptr(), T, B, H,
                                  dgx.data_ptr(), work.data_ptr(), wp.data_ptr(), dgx16.data_ptr(), U.stream()), 'lstm_bwd cluster')
    torch.cuda.synchronize()
    assert U.relerr(dgx.cpu(), dgx_ref.cpu()) < 1.5e-2, U.relerr(dgx.cpu(), dgx_ref.cpu())     # bf16 W_hh / dg / partial sums
    assert U.relerr(dgx16.float().cpu(), dgx.cpu()) < 4e-3


@pytest.mark.parametrize('B,T', [(32, 125), (64, 125), (32, 750)])
def test_lstm_cluster_long_sequences_vs_torch_autograd(B, T):
    """The cluster recurrence at the lengths the configs use (cfg 2: T' = 125, cfg 5: T' = 750), forward against the
    step-by-step torch recurrence with the same operand rounding, BACKWARD against torch autograd through that recurrence
    (straight-through for the bf16 rounding of h) -- not against another kernel of this library."""
    torch.manual_seed(8)
    H = 500
    gx = (torch.randn(B, T, 4 * H) * 0.7).requires_grad_(True)
    w_hh = torch.randn(4 * H, H) * 0.05
    w_bf = w_hh.bfloat16().float()
    h = torch.zeros(B, H)
    c = torch.zeros(B, H)
    outs = []
    for t in range(T):
        hr = h + (h.bfloat16().float() - h).detach()               # bf16 operand, identity gradient
        g = gx[:, t] + hr @ w_bf.t()
        i, f, gg, o = g.split(H, 1)
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
        h = torch.sigmoid(o) * torch.tanh(c)
        outs.append(h)
    hs = torch.stack(outs, 1)
    dh = torch.randn(B, T, H) * (1.0 / T)
    (dgx_ref,) = torch.autograd.grad(hs, gx, dh)
    lib = _lib.load()
    whh = w_hh.to(U.DEV)
    wp = torch.zeros(16 * 128 * 512, dtype=torch.bfloat16, device=U.DEV)
    job = (_lib.PackJob * 1)()
    job[0].kind, job[0].out_dtype, job[0].src, job[0].dst, job[0].n_out = 4, BF16, whh.data_ptr(), wp.data_ptr(), 16 * 128 * 512
    job[0].a[0] = H
    jd = torch.frombuffer(bytearray(bytes(job)), dtype=torch.uint8).to(U.DEV)
    nblk = (16 * 128 * 512 + 4095) // 4096
    bmap = torch.stack([torch.zeros(nblk, dtype=torch.int32), torch.arange(nblk, dtype=torch.int32)], 1).contiguous().to(U.DEV)
    _lib.check(lib.nbasr_pack_batch(jd.data_ptr(), 1, bmap.data_ptr(), nblk, U.stream()))
    gxd = gx.detach().contiguous().to(U.DEV)
    hseq = torch.zeros(B, T, 512, dtype=torch.bfloat16, device=U.DEV)
    gates = torch.zeros(B * T, 4 * H, device=U.DEV)
    cst = torch.zeros(B * T, H, device=U.DEV)
    work = torch.zeros(2 * B * H + 256, device=U.DEV)
    _lib.check(lib.nbasr_lstm_fwd(gxd.data_ptr(), whh.data_ptr(), T, B, H, hseq.data_ptr(), BF16, T * 512, 512, 512, gates.data_ptr(),
                                  cst.data_ptr(), wp.data_ptr(), work.data_ptr(), U.stream()), 'lstm_fwd cluster')
    torch.cuda.synchronize()
    assert U.relerr(hseq[:, :, :H].float().cpu(), hs.detach()) < 8e-3          # bf16 output rounding, no drift over T steps
    assert U.relerr(cst.view(B, T, H)[:, -1].cpu(), c.detach()) < 5e-3
    dhd = dh.contiguous().to(U.DEV)
    dgx = torch.zeros(B * T, 4 * H, device=U.DEV)
    dgx16 = torch.zeros(B * T, 4 * H, dtype=torch.bfloat16, device=U.DEV)
    _lib.check(lib.nbasr_lstm_bwd(dhd.data_ptr(), T * H, H, H, whh.data_ptr(), gates.data_ptr(), cst.data_ptr(), T, B, H,
                                  dgx.data_ptr(), work.data_ptr(), wp.data_ptr(), dgx16.data_ptr(), U.stream()), 'lstm_bwd cluster')
    torch.cuda.synchronize()
    err = U.relerr(dgx.view(B, T, 4 * H).cpu(), dgx_ref)
    assert err < 2e-2, err                                                     # bf16 W_hh / dg operands through T steps
    assert U.relerr(dgx16.float().cpu(), dgx.cpu()) < 4e-3
