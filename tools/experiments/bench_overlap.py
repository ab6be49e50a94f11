"""Can the grouped-conv input-gradient (tensor-issue bound) and weight-gradient (data-movement bound) kernels of one node
overlap?  Sequential at 2 CTAs/SM each (product) vs concurrently on two streams at 1 CTA/SM each."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
from nb_asr_b200 import _lib
from nb_asr_b200._lib import BF16, GConv
import gpu_utils as U
lib = _lib.load()
B = 64
for Cc, T, k, d in ((800, 500, 5, 1), (1000, 250, 5, 1), (1200, 125, 5, 1), (600, 500, 5, 1)):
    cpg = Cc // 100
    x = U.to_padded(torch.randn(B, T, Cc), BF16)
    dz = U.to_padded(torch.randn(B, T, Cc), BF16)
    out, out2 = U.empty_padded(B, T, Cc, BF16), U.empty_padded(B, T, Cc, BF16)
    mwid = 40 if cpg == 10 else 48
    mask = U.new_mask(out.shape[0], Cc, mwid); mask.random_(0, 255)
    w = torch.randn(Cc, cpg, k, device=U.DEV) * 0.3
    dw = torch.zeros(Cc, cpg, k, device=U.DEV); db = torch.zeros(Cc, device=U.DEV)
    ne = int(lib.nbasr_gconv_mma_pack_elems(Cc, cpg, k))
    wp = torch.zeros(ne, dtype=torch.bfloat16, device=U.DEV)
    _lib.check(lib.nbasr_pack_gconv_mma(w.data_ptr(), wp.data_ptr(), Cc, cpg, k, 1, U.stream()))
    gc = GConv()
    gc.dtype, gc.x, gc.B, gc.T, gc.Tp, gc.C, gc.cpg, gc.ktaps, gc.off0, gc.dstep = BF16, dz.data_ptr(), B, T, U.geo(T), Cc, cpg, k, -4, d
    gc.w, gc.w_packed = wp.data_ptr(), 1
    gc.epi = U.epilogue(BF16, Cc, out=out, out2=out2, mask2=mask, mask2_w=mwid)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=U.DEV)

    def wgrad(st):
        _lib.check(lib.nbasr_gconv_wgrad(BF16, dz.data_ptr(), x.data_ptr(), B, T, U.geo(T), Cc, cpg, k, 0, d, dw.data_ptr(), db.data_ptr(), st))

    def dgrad(st):
        _lib.check(lib.nbasr_gconv_fwd(C.byref(gc), st))

    def run(mode):
        ts = []
        for it in range(6):
            flush.zero_()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            cur = torch.cuda.current_stream()
            e0.record(cur)
            if mode == 'seq':
                wgrad(cur.cuda_stream); dgrad(cur.cuda_stream)
            else:
                s1.wait_stream(cur); s2.wait_stream(cur)
                wgrad(s1.cuda_stream); dgrad(s2.cuda_stream)
                cur.wait_stream(s1); cur.wait_stream(s2)
            e1.record(cur)
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return sorted(ts[1:])[2] * 1e3
    lib.nbasr_dbg_gconv_slots(2)
    t_seq = run('seq')
    t_par2 = run('par')
    lib.nbasr_dbg_gconv_slots(1)
    t_seq1 = run('seq')
    t_par1 = run('par')
    lib.nbasr_dbg_gconv_slots(2)
    print(f'C={Cc} T={T}: sequential 2 CTAs/SM {t_seq:6.1f} us | two streams 2 CTAs/SM {t_par2:6.1f} | sequential 1 CTA/SM {t_seq1:6.1f} | two streams 1 CTA/SM {t_par1:6.1f}', flush=True)
