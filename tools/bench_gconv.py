"""Micro-benchmark of the tcgen05 grouped-conv forward kernel: which part of the epilogue costs what."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import torch
from nb_asr_b200 import _lib
from nb_asr_b200._lib import BF16, GConv
import gpu_utils as U
lib = _lib.load()
B, T = 64, 500
for Cc, k, d in ((1000, 5, 1), (600, 5, 1), (1200, 7, 2)):
    cpg = Cc // 100
    x = U.to_padded(torch.randn(B, T, Cc), BF16)
    sk = U.to_padded(torch.randn(B, T, Cc), BF16)
    out = U.empty_padded(B, T, Cc, BF16)
    out2 = U.empty_padded(B, T, Cc, BF16)
    mwid = 40 if cpg == 10 else 48
    mask = U.new_mask(out.shape[0], Cc, mwid)
    w = torch.randn(Cc, cpg, k, device=U.DEV) * 0.3
    bias = torch.randn(Cc, device=U.DEV)
    ne = int(lib.nbasr_gconv_mma_pack_elems(Cc, cpg, k))
    wp = torch.zeros(ne, dtype=torch.bfloat16, device=U.DEV)
    _lib.check(lib.nbasr_pack_gconv_mma(w.data_ptr(), wp.data_ptr(), BF16, Cc, cpg, k, 0, U.stream()))
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=U.DEV)
    variants = {
        'fwd: bias+relu+out+mask': dict(bias=bias, relu=1, out=out, mask_out=mask, mask_w=mwid),
        'no mask': dict(bias=bias, relu=1, out=out),
        'no out (mask only)': dict(bias=bias, relu=1, mask_out=mask, mask_w=mwid),
        'nothing stored': dict(bias=bias, relu=1),
        'fwd + 1 skip': dict(bias=bias, relu=1, adds=[sk], out=out, mask_out=mask, mask_w=mwid),
        'dgrad: out+out2+mask2': dict(out=out, out2=out2, mask2=mask, mask2_w=mwid),
    }
    for name, kw in variants.items():
        gc = GConv()
        gc.dtype, gc.x, gc.B, gc.T, gc.Tp, gc.C, gc.cpg, gc.ktaps, gc.off0, gc.dstep = BF16, x.data_ptr(), B, T, U.geo(T), Cc, cpg, k, 0 if d == 1 else -8, d
        gc.w, gc.w_packed = wp.data_ptr(), 1
        gc.epi = U.epilogue(BF16, Cc, **kw)
        ts = []
        for it in range(6):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _lib.check(lib.nbasr_gconv_fwd(C.byref(gc), U.stream()))
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts = sorted(ts[1:])
        t = ts[len(ts) // 2]
        el = B * T * Cc * 2
        print(f'C={Cc} k={k} d={d} {name:28s} {t*1e3:8.1f} us   in+out = {2*el/t/1e6:7.0f} GB/s', flush=True)
