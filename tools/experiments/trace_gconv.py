"""Timeline of the grouped-conv forward kernel roles (globaltimer stamps, first 8 CTAs)."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch  # noqa: E402
from nb_asr_b200 import _lib  # noqa: E402
from nb_asr_b200._lib import BF16, GConv  # noqa: E402
import gpu_utils as U  # noqa: E402

lib = _lib.load()
lib.nbasr_dbg_gconv_trace.argtypes = [C.c_void_p]
B, T, Cc, k, d = 64, 500, 1000, 5, 1
cpg = Cc // 100
x = U.to_padded(torch.randn(B, T, Cc), BF16)
out = U.empty_padded(B, T, Cc, BF16)
mwid = 40 if cpg == 10 else 48
mask = U.new_mask(out.shape[0], Cc, mwid)
w = torch.randn(Cc, cpg, k, device=U.DEV) * 0.3
bias = torch.randn(Cc, device=U.DEV)
ne = int(lib.nbasr_gconv_mma_pack_elems(Cc, cpg, k))
wp = torch.zeros(ne, dtype=torch.bfloat16, device=U.DEV)
_lib.check(lib.nbasr_pack_gconv_mma(w.data_ptr(), wp.data_ptr(), Cc, cpg, k, 0, U.stream()))
for name, kw in (('full', dict(bias=bias, relu=1, out=out, mask_out=mask, mask_w=mwid)), ('nostore', dict(bias=bias, relu=1))):
    gc = GConv()
    gc.dtype, gc.x, gc.B, gc.T, gc.Tp, gc.C, gc.cpg, gc.ktaps, gc.off0, gc.dstep = BF16, x.data_ptr(), B, T, U.geo(T), Cc, cpg, k, 0, d
    gc.w, gc.w_packed = wp.data_ptr(), 1
    gc.epi = U.epilogue(BF16, Cc, **kw)
    for i in range(2):
        _lib.check(lib.nbasr_gconv_fwd(C.byref(gc), U.stream()))
    torch.cuda.synchronize()
    buf = torch.zeros(8 * 64 * 8, dtype=torch.int64, device=U.DEV)
    if os.environ.get('COLD'):
        torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=U.DEV).zero_()     # evict the inputs from L2
        torch.cuda.synchronize()
    lib.nbasr_dbg_gconv_trace(buf.data_ptr())
    _lib.check(lib.nbasr_gconv_fwd(C.byref(gc), U.stream()))
    torch.cuda.synchronize()
    lib.nbasr_dbg_gconv_trace(None)
    t = buf.view(8, 64, 8).cpu()
    t0 = int(t[t > 0].min())
    print('=== variant', name)
    if os.environ.get('NBASR_GCONV_FRAG') is not None:
        # fragment kernel (gconv_frag_sm100.cu): clock64 stamps of one SM -> cycles
        for cta in (0, 1):
            print(f'CTA {cta}: item | prod_issue | wait_begin full_ok mma_done epi_done bar_done store_issued  (cycles since the CTA\'s first stamp)')
            tc = t[cta]
            c0 = int(tc[tc > 0].min())
            for it in range(8, 16):
                r = tc[it]
                if r[1] == 0:
                    break
                print(f'  {it:3d} | ' + ' '.join(f'{int(v) - c0:8d}' if v > 0 else '     -  ' for v in r[:7]))
        continue
    for cta in (0, 1):
        print(f'CTA {cta}: item | bar2_done mma_full_ok mma_issued | epi_tfull tmem_released bar1_done staged store_issued  (us since first stamp; v1 kernel: see gconv_sm100.cu)')
        for it in range(0, 24):
            r = t[cta, it]
            if r[1] == 0:
                break
            print(f'  {it:3d} | ' + ' '.join(f'{(int(v) - t0) / 1e3:8.2f}' if v > 0 else '     -  ' for v in r[:8]))
