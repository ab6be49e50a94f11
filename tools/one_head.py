"""Head forward (classifier + log-softmax) at the bench shape: ncu target / timing."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
from nb_asr_b200 import _lib
from nb_asr_b200._lib import BF16
import gpu_utils as U
lib = _lib.load()
B, T, K, V, HP = 64, 125, 500, 49, 512
h = torch.randn(B, T, HP, device=U.DEV).bfloat16().contiguous()
w = torch.randn(V, K, device=U.DEV) * 0.1
b = torch.randn(V, device=U.DEV)
logits = torch.empty(B, T, V, device=U.DEV); logp = torch.empty(B, T, V, device=U.DEV)
for it in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _lib.check(lib.nbasr_head_fwd(BF16, h.data_ptr(), T * HP, HP, B, T, K, V, w.data_ptr(), b.data_ptr(), logits.data_ptr(), logp.data_ptr(), U.stream()))
    e1.record(); torch.cuda.synchronize()
    print(f'head_fwd {e0.elapsed_time(e1)*1e3:.1f} us')
ref = torch.log_softmax(h[:, :, :K].float() @ w.t() + b, -1)
print('max err', float((logp - ref).abs().max()))
