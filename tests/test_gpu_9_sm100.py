"""-m gpu: the tcgen05/TMEM/TMA GEMM kernels (bf16) against a torch fp32 reference on the same bf16-rounded inputs."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from nb_asr_b200._lib import BF16, F16, F32, PAD_L  # noqa: E402
from nb_asr_b200.model import pad_rule  # noqa: E402
from oracle import model_ref as M  # noqa: E402
import gpu_utils as U  # noqa: E402


def _bf(t):
    return t.bfloat16().float()


@pytest.mark.parametrize('nb,nr,K,N', [(1, 128, 128, 64), (1, 128, 64, 256), (3, 500, 600, 600), (2, 125, 1200, 1200),
                                       (16, 500, 640, 800), (4, 250, 2000, 1000)])
def test_gemm_tn_plain(nb, nr, K, N):
    torch.manual_seed(10)
    x = _bf(torch.randn(nb, nr, K))
    w = _bf(torch.randn(N, K) * 0.1)
    ref = x @ w.t()
    xb = U.to_padded(x, BF16)
    out = U.empty_padded(nb, nr, N, F32)
    epi = U.epilogue(BF16, N, out=out, out_dtype=F32)
    U.run_gemm(BF16, U.ptr(xb, PAD_L * K), U.geo(nr) * K, K, nb, nr, K, N, w.bfloat16().to(U.DEV), K, PAD_L, U.geo(nr), 1, epi)
    got = U.from_padded(out, nb, nr).cpu()
    assert U.relerr(got, ref) < 1e-5, U.relerr(got, ref)
    assert float(out[:PAD_L].abs().sum()) == 0.0


@pytest.mark.parametrize('stride', [1, 2])
def test_gemm_tn_conv_epilogue(stride):
    torch.manual_seed(11)
    B, T, Cin, Cout = 3, 300, 80, 600
    x = _bf(torch.randn(B, T, Cin))
    w = _bf(torch.randn(Cout, Cin, 8) * 0.05)
    bias = torch.randn(Cout) * 0.5
    To = (T + stride - 1) // stride
    skip = _bf(torch.randn(B, To, Cout))
    z = F.conv1d(F.pad(x.permute(0, 2, 1), pad_rule(8, 1, stride)), w, bias, stride=stride).permute(0, 2, 1)
    ref = M.relu20(z) + skip
    xb = U.to_padded(x, BF16)
    wp = w.permute(0, 2, 1).contiguous().view(Cout, 8 * Cin).bfloat16().to(U.DEV)
    out = U.empty_padded(B, To, Cout, BF16)
    sk = U.to_padded(skip, BF16)
    mask = U.new_mask(out.shape[0], Cout)
    lpad, _ = pad_rule(8, 1, stride)
    epi = U.epilogue(BF16, Cout, bias=bias.to(U.DEV), relu=1, adds=[sk], out=out, mask_out=mask)
    U.run_gemm(BF16, U.ptr(xb, (PAD_L - lpad) * Cin), U.geo(T) * Cin, stride * Cin, B, To, 8 * Cin, Cout, wp, 8 * Cin, PAD_L,
               U.geo(To), 1, epi)
    got = U.from_padded(out, B, To).cpu()
    assert U.relerr(got, ref) < 4e-3, U.relerr(got, ref)      # bf16 output rounding only
    m = U.unpack_mask(mask, B, To, Cout).cpu()
    exp = (z > 0) & (z <= 20)
    assert (m != exp).float().mean() < 1e-4                    # accumulation-order ties at z ~ 0 only


@pytest.mark.parametrize('stride', [1, 2])
def test_gemm_tn_scaled_fp16_forward(stride):
    """16-bit mode forward GEMM: fp16 operands (activations hold S*x), scaled-domain epilogue, fp16 output + bf16 twin; and the
    un-scaling epilogue (acc_scale = 1/S, fp32 output) the LSTM input projection uses."""
    torch.manual_seed(11)
    S = 32.0
    h = lambda t: t.half().float()
    B, T, Cin, Cout = 3, 300, 80, 600
    x = h(torch.randn(B, T, Cin) * S) / S
    w = h(torch.randn(Cout, Cin, 8) * 0.05)
    bias = torch.randn(Cout) * 0.5
    To = (T + stride - 1) // stride
    skip = h(torch.randn(B, To, Cout) * S) / S
    z = F.conv1d(F.pad(x.permute(0, 2, 1), pad_rule(8, 1, stride)), w, bias, stride=stride).permute(0, 2, 1)
    ref = M.relu20(z) + skip
    xb = U.to_padded(x * S, F16)
    wp = w.permute(0, 2, 1).contiguous().view(Cout, 8 * Cin).half().to(U.DEV)
    out, out2 = U.empty_padded(B, To, Cout, F16), U.empty_padded(B, To, Cout, BF16)
    sk = U.to_padded(skip * S, F16)
    mask = U.new_mask(out.shape[0], Cout)
    lpad, _ = pad_rule(8, 1, stride)
    epi = U.epilogue(F16, Cout, bias=bias.to(U.DEV), relu=1, adds=[sk], out=out, mask_out=mask, out2=out2, out2_dtype=BF16,
                     scale2=1.0 / S, bias_scale=S, relu_hi=20.0 * S)
    U.run_gemm(F16, U.ptr(xb, (PAD_L - lpad) * Cin), U.geo(T) * Cin, stride * Cin, B, To, 8 * Cin, Cout, wp, 8 * Cin, PAD_L,
               U.geo(To), 1, epi)
    assert U.relerr(U.from_padded(out, B, To).cpu() / S, ref) < 6e-4          # fp16 output rounding only
    assert U.relerr(U.from_padded(out2, B, To).cpu(), ref) < 4e-3             # bf16 twin
    m = U.unpack_mask(mask, B, To, Cout).cpu()
    assert (m != ((z > 0) & (z <= 20))).float().mean() < 1e-4
    # un-scaling epilogue: fp32 output of true values
    o32 = U.empty_padded(B, To, Cout, F32)
    epi = U.epilogue(F16, Cout, bias=bias.to(U.DEV), out=o32, out_dtype=F32, acc_scale=1.0 / S)
    U.run_gemm(F16, U.ptr(xb, (PAD_L - lpad) * Cin), U.geo(T) * Cin, stride * Cin, B, To, 8 * Cin, Cout, wp, 8 * Cin, PAD_L,
               U.geo(To), 1, epi)
    assert U.relerr(U.from_padded(o32, B, To).cpu(), z) < 1e-5


@pytest.mark.parametrize('nb,nr,M_,N,ldx', [(1, 64, 128, 256, 256), (2, 100, 64, 64, 64), (4, 500, 600, 640, 640),
                                             (3, 125, 2000, 500, 512)])
def test_gemm_wgrad(nb, nr, M_, N, ldx):
    # ldx > N: the LSTM h_seq buffer keeps 500 hidden units in rows of 512 (16-byte aligned pitch for TMA)
    torch.manual_seed(12)
    dy = _bf(torch.randn(nb, nr, M_))
    x = _bf(torch.randn(nb, nr, ldx))
    ref = torch.einsum('brm,brn->mn', dy, x[:, :, :N])
    dyb, xb = U.to_padded(dy, BF16), U.to_padded(x, BF16)
    dw = torch.zeros(M_, N, device=U.DEV)
    db = torch.zeros(M_, device=U.DEV)
    args = (BF16, U.ptr(dyb, PAD_L * M_), U.geo(nr) * M_, M_, U.ptr(xb, PAD_L * ldx), U.geo(nr) * ldx, ldx, nb, nr, M_, N, dw, N)
    U.run_wgrad(*args, dbias=db)
    assert U.relerr(dw.cpu(), ref) < 1e-5, U.relerr(dw.cpu(), ref)
    assert U.relerr(db.cpu(), dy.sum((0, 1))) < 1e-5       # fused bias gradient (ones operand)
    U.run_wgrad(*args)                                      # accumulates (+=)
    assert U.relerr(dw.cpu(), 2 * ref) < 1e-5


def test_gemm_wgrad_conv_view():
    torch.manual_seed(13)
    B, T, Cin, Cout, stride = 2, 260, 80, 128, 2
    x = _bf(torch.randn(B, T, Cin)).requires_grad_(True)
    w = _bf(torch.randn(Cout, Cin, 8) * 0.1).requires_grad_(True)
    y = F.conv1d(F.pad(x.permute(0, 2, 1), pad_rule(8, 1, stride)), w, None, stride=stride).permute(0, 2, 1)
    To = y.shape[1]
    dy = _bf(torch.randn_like(y))
    (gw,) = torch.autograd.grad(y, (w,), dy)
    xb, dyb = U.to_padded(x.detach(), BF16), U.to_padded(dy, BF16)
    lpad, _ = pad_rule(8, 1, stride)
    dw = torch.zeros(Cout, 8 * Cin, device=U.DEV)
    U.run_wgrad(BF16, U.ptr(dyb, PAD_L * Cout), U.geo(To) * Cout, Cout, U.ptr(xb, (PAD_L - lpad) * Cin), U.geo(T) * Cin, stride * Cin,
                B, To, Cout, 8 * Cin, dw, 8 * Cin)
    got = dw.view(Cout, 8, Cin).permute(0, 2, 1).cpu()
    assert U.relerr(got, gw) < 1e-5, U.relerr(got, gw)
