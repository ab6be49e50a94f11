"""Micro-benchmark of gemm_tn (tcgen05) on the shapes of the train step: dense convs, linear edges, LSTM projection."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
from nb_asr_b200 import _lib
from nb_asr_b200._lib import BF16, PAD_L, Gemm
import gpu_utils as U
lib = _lib.load()
B = 64
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=U.DEV)
# (name, T_in, Cin, stride, taps, Cout, epilogue kind)
shapes = [('conv0 80->600 s1', 500, 80, 1, 8, 600, 'relu'), ('conv1 600->800 s1', 500, 600, 1, 8, 800, 'relu'),
          ('conv2 800->1000 s2', 500, 800, 2, 8, 1000, 'relu'), ('conv3 1000->1200 s2', 250, 1000, 2, 8, 1200, 'relu'),
          ('linear 600', 500, 600, 1, 1, 600, 'relu'), ('linear 800', 500, 800, 1, 1, 800, 'relu'),
          ('linear 1000', 250, 1000, 1, 1, 1000, 'relu'), ('linear 1200', 125, 1200, 1, 1, 1200, 'relu'),
          ('linear 800 +2 skips', 500, 800, 1, 1, 800, 'skips'), ('lstm proj 1200->2000 f32 out', 125, 1200, 1, 1, 2000, 'f32'),
          ('dgrad 2000->1200', 125, 2000, 1, 1, 1200, 'plain')]
only = os.environ.get('ONLY')
for name, T, Cin, s, taps, Cout, kind in shapes:
    if only and only not in name:
        continue
    To = (T + s - 1) // s
    x = U.to_padded(torch.randn(B, T, Cin), BF16)
    w = (torch.randn(Cout, taps * Cin, device=U.DEV) * 0.05).bfloat16()
    bias = torch.randn(Cout, device=U.DEV)
    out = U.empty_padded(B, To, Cout, BF16) if kind != 'f32' else torch.zeros(B * U.geo(To) + 8, Cout, device=U.DEV)
    mask = U.new_mask(out.shape[0], Cout)
    sk = [U.to_padded(torch.randn(B, To, Cout), BF16) for _ in range(2)]
    if kind == 'relu':
        epi = U.epilogue(BF16, Cout, bias=bias, relu=1, out=out, mask_out=mask)
    elif kind == 'skips':
        epi = U.epilogue(BF16, Cout, bias=bias, relu=1, adds=sk, out=out, mask_out=mask)
    elif kind == 'f32':
        epi = U.epilogue(BF16, Cout, bias=bias, out=out, out_dtype=0)
    else:
        epi = U.epilogue(BF16, Cout, out=out)
    g = Gemm()
    lpad = 0 if taps == 1 else (3 if s == 1 else 5)
    g.dtype, g.a, g.a_bs, g.a_rs, g.nb, g.nr, g.K, g.N = BF16, U.ptr(x, (PAD_L - lpad) * Cin), U.geo(T) * Cin, s * Cin, B, To, taps * Cin, Cout
    g.w, g.ldw, g.o_r0, g.o_bs, g.o_rs, g.epi = w.data_ptr(), taps * Cin, PAD_L, U.geo(To), 1, epi
    ts = []
    for it in range(2 if only else 6):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); _lib.check(lib.nbasr_gemm_tn(C.byref(g), U.stream())); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = sorted(ts[1:])[len(ts[1:]) // 2]
    fl = 2.0 * B * To * taps * Cin * Cout
    print(f'{name:30s} M={B*To:6d} K={taps*Cin:5d} N={Cout:5d}  {t*1e3:8.1f} us  {fl/t/1e9:7.1f} TFLOP/s', flush=True)
