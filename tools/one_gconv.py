"""Launch the grouped-conv forward kernel a few times on one shape (ncu target)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
from nb_asr_b200 import _lib
from nb_asr_b200._lib import BF16, GConv
import gpu_utils as U
lib = _lib.load()
B, T, Cc, k, d = 64, 500, int(os.environ.get('C', 1000)), int(os.environ.get('K', 5)), int(os.environ.get('D', 1))
cpg = Cc // 100
x = U.to_padded(torch.randn(B, T, Cc), BF16)
out = U.empty_padded(B, T, Cc, BF16)
mwid = 40 if cpg == 10 else 48
mask = U.new_mask(out.shape[0], Cc, mwid)
w = torch.randn(Cc, cpg, k, device=U.DEV) * 0.3
bias = torch.randn(Cc, device=U.DEV)
ne = int(lib.nbasr_gconv_mma_pack_elems(Cc, cpg, k))
wp = torch.zeros(ne, dtype=torch.bfloat16, device=U.DEV)
_lib.check(lib.nbasr_pack_gconv_mma(w.data_ptr(), wp.data_ptr(), BF16, Cc, cpg, k, 0, U.stream()))
gc = GConv()
gc.dtype, gc.x, gc.B, gc.T, gc.Tp, gc.C, gc.cpg, gc.ktaps, gc.off0, gc.dstep = BF16, x.data_ptr(), B, T, U.geo(T), Cc, cpg, k, 0 if d == 1 else -8, d
gc.w, gc.w_packed = wp.data_ptr(), 1
gc.epi = U.epilogue(BF16, Cc, bias=bias, relu=1, out=out, mask_out=mask, mask_w=mwid)
for i in range(4):
    _lib.check(lib.nbasr_gconv_fwd(C.byref(gc), U.stream()))
torch.cuda.synchronize()
