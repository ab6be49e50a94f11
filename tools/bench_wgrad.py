"""Micro-benchmark of the dense weight-gradient GEMM (tcgen05, CTA pairs) on the shapes of the train step, with / without the
fused bias gradient."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
from nb_asr_b200 import _lib
from nb_asr_b200._lib import BF16, PAD_L, Wgrad
import gpu_utils as U
lib = _lib.load()
B = 64
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=U.DEV)
# (name, T_out, Cin, stride, taps, Cout)
shapes = [('conv1 600->800', 500, 600, 1, 8, 800), ('conv2 800->1000 s2', 250, 800, 2, 8, 1000), ('conv3 1000->1200 s2', 125, 1000, 2, 8, 1200),
          ('linear 800', 500, 800, 1, 1, 800), ('linear 1200', 125, 1200, 1, 1, 1200), ('lstm W_ih 1200->2000', 125, 1200, 1, 1, 2000)]
for name, To, Cin, s, taps, Cout in shapes:
    Tin = To * s
    x = U.to_padded(torch.randn(B, Tin, Cin), BF16)
    dy = U.to_padded(torch.randn(B, To, Cout), BF16)
    dw = torch.zeros(Cout, taps * Cin, device=U.DEV)
    db = torch.zeros(Cout, device=U.DEV)
    for with_bias in (1, 0):
        w = Wgrad()
        w.dtype, w.dy, w.dy_bs, w.dy_rs = BF16, U.ptr(dy, PAD_L * Cout), U.geo(To) * Cout, Cout
        w.x, w.x_bs, w.x_rs = U.ptr(x, (PAD_L - (3 if taps > 1 and s == 1 else (5 if taps > 1 else 0))) * Cin), U.geo(Tin) * Cin, s * Cin
        w.nb, w.nr, w.M, w.N, w.dw, w.ldw = B, To, Cout, taps * Cin, dw.data_ptr(), taps * Cin
        w.dbias = db.data_ptr() if with_bias else None
        ts = []
        for it in range(6):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); _lib.check(lib.nbasr_gemm_wgrad(C.byref(w), U.stream())); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        t = sorted(ts[1:])[2]
        fl = 2.0 * B * To * taps * Cin * Cout
        print(f'{name:24s} M={Cout:5d} N={taps*Cin:5d} frames={B*To:6d} bias={with_bias}  {t*1e3:8.1f} us  {fl/t/1e9:7.1f} TFLOP/s', flush=True)
