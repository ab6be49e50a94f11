"""Helpers for the -m gpu parity tests: padded-buffer conversion and thin torch wrappers over the C ABI."""
import ctypes as C

import torch

from nb_asr_b200 import _lib
from nb_asr_b200._lib import BF16, F16, F32, PAD_L, PAD_R, Epilogue, GConv, Gemm, Wgrad

DEV = 'cuda:0'


def tdt(dt):
    return {BF16: torch.bfloat16, F16: torch.float16}.get(dt, torch.float32)


def geo(T):
    tp = T + PAD_L + PAD_R
    return tp + (tp & 1)


def to_padded(x, dt):
    """x (B,T,C) -> zero padded (B*Tp+8, C) buffer on the GPU."""
    B, T, Cc = x.shape
    Tp = geo(T)
    buf = torch.zeros(B * Tp + 8, Cc, dtype=tdt(dt), device=DEV)
    buf[:B * Tp].view(B, Tp, Cc)[:, PAD_L:PAD_L + T] = x.to(DEV).to(tdt(dt))
    return buf


def from_padded(buf, B, T):
    Tp = geo(T)
    return buf[:B * Tp].view(B, Tp, -1)[:, PAD_L:PAD_L + T].float()


def empty_padded(B, T, Cc, dt):
    return torch.zeros(B * geo(T) + 8, Cc, dtype=tdt(dt), device=DEV)


def stream():
    return torch.cuda.current_stream().cuda_stream


def ptr(t, off=0):
    return t.data_ptr() + off * t.element_size()


def new_mask(rows, Cc, w=32):
    """plane-major gate-bit mask buffer (see include/nbasr.h): (planes, rows, entry bytes) uint8"""
    eb = 4 if w == 32 else 8
    return torch.zeros((Cc + w - 1) // w, rows, eb, dtype=torch.uint8, device=DEV)


def epilogue(dt, ld, bias=None, relu=0, drop_p=0.0, salt=0, adds=(), out=None, out_dtype=None, mask_out=None, out2=None,
             mask2=None, scale2=1.0, mask_w=32, mask2_w=32, accumulate=0, out2_dtype=None, acc_scale=0.0, bias_scale=0.0,
             relu_hi=0.0):
    e = Epilogue()
    e.bias = bias.data_ptr() if bias is not None else None
    e.relu20, e.drop_p, e.drop_seed, e.drop_step = relu, drop_p, salt, None
    e.n_add = len(adds)
    for i, a in enumerate(adds):
        e.add[i] = a.data_ptr()
    e.add_dtype = dt
    e.out = out.data_ptr() if out is not None else None
    e.out_dtype = dt if out_dtype is None else out_dtype
    e.ld_out = ld
    e.mask_out = mask_out.data_ptr() if mask_out is not None else None
    e.out2 = out2.data_ptr() if out2 is not None else None
    e.out2_dtype = dt if out2_dtype is None else out2_dtype
    e.acc_scale, e.bias_scale, e.relu_hi = acc_scale, bias_scale, relu_hi
    e.mask2 = mask2.data_ptr() if mask2 is not None else None
    e.scale2, e.accumulate = scale2, accumulate
    e.mask_w, e.mask2_w = mask_w, mask2_w
    m_any = mask_out if mask_out is not None else mask2
    e.mask_rows = m_any.shape[1] if m_any is not None else 0
    return e


def run_gemm(dt, a_ptr, a_bs, a_rs, nb, nr, K, N, w, ldw, o_r0, o_bs, o_rs, epi):
    lib = _lib.load()
    g = Gemm()
    g.dtype, g.a, g.a_bs, g.a_rs, g.nb, g.nr, g.K, g.N = dt, a_ptr, a_bs, a_rs, nb, nr, K, N
    g.w, g.ldw, g.o_r0, g.o_bs, g.o_rs, g.epi = w.data_ptr(), ldw, o_r0, o_bs, o_rs, epi
    _lib.check(lib.nbasr_gemm_tn(C.byref(g), stream()), 'gemm_tn')
    torch.cuda.synchronize()


def run_wgrad(dt, dy_ptr, dy_bs, dy_rs, x_ptr, x_bs, x_rs, nb, nr, M, N, dw, ldw, dbias=None):
    lib = _lib.load()
    w = Wgrad()
    w.dtype, w.dy, w.dy_bs, w.dy_rs, w.x, w.x_bs, w.x_rs = dt, dy_ptr, dy_bs, dy_rs, x_ptr, x_bs, x_rs
    w.nb, w.nr, w.M, w.N, w.dw, w.ldw = nb, nr, M, N, dw.data_ptr(), ldw
    w.dbias = dbias.data_ptr() if dbias is not None else None
    _lib.check(lib.nbasr_gemm_wgrad(C.byref(w), stream()), 'gemm_wgrad')
    torch.cuda.synchronize()


def relerr(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-300))


def unpack_mask(mask, B, T, Cc, w=32):
    """(planes, rows, eb) uint8 plane-major bit mask -> bool (B,T,C)."""
    Tp = geo(T)
    planes, rows, eb = mask.shape
    m = mask[:, :B * Tp].view(planes, B, Tp, eb)[:, :, PAD_L:PAD_L + T]          # (planes, B, T, eb)
    bits = ((m.unsqueeze(-1).to(torch.int32) >> torch.arange(8, device=m.device, dtype=torch.int32)) & 1).bool()
    bits = bits.reshape(planes, B, T, eb * 8)[..., :w]                              # (planes, B, T, w)
    return bits.permute(1, 2, 0, 3).reshape(B, T, planes * w)[:, :, :Cc]
