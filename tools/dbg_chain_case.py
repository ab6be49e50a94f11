import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
from nb_asr_b200 import _lib
from nb_asr_b200._lib import GConv
import gpu_utils as U
import test_gpu_8_gconv_chain as TC
lib = _lib.load()
case = int(os.environ.get('CASE', 0))
Cc, B, T, ops, skips, dt, backward = TC.CASES[case]
n = len(ops)
ref_nodes, ref_outs, k1 = TC._build(lib, Cc, B, T, ops, skips, dt, backward, seed=case)
nodes, outs, k2 = TC._build(lib, Cc, B, T, ops, skips, dt, backward, seed=case)
for gc in ref_nodes:
    _lib.check(lib.nbasr_gconv_fwd(C.byref(gc), U.stream()))
wb = int(lib.nbasr_gconv_chain_work_bytes(B, T, Cc, Cc // 100, 3))
work = torch.zeros(wb // 4, dtype=torch.int32, device=U.DEV)
arr = (GConv * n)(*nodes)
_lib.check(lib.nbasr_gconv_chain(arr, n, 1, work.data_ptr(), wb, U.stream()))
torch.cuda.synchronize()
print('work', work[:4].tolist())
for i, (a, b) in enumerate(zip(outs, ref_outs)):
    if a.dtype == torch.uint8:
        nb = (40 if Cc // 100 == 10 else 48) // 8
        d = (a[..., :nb] ^ b[..., :nb]).to(torch.int32)
        nbits = sum(((d >> j) & 1).sum().item() for j in range(8))
        print(i, 'mask bits differing', nbits, 'of', a[..., :nb].numel() * 8, 'set in ref', sum(((b[..., :nb].to(torch.int32) >> j) & 1).sum().item() for j in range(8)))
    else:
        af, bf = a.float(), b.float()
        print(i, a.dtype, 'relerr', U.relerr(af, bf), 'maxabs', float((af - bf).abs().max()), 'ref absmax', float(bf.abs().max()), 'nan', bool(torch.isnan(af).any()))
