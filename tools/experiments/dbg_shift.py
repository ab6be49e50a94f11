"""GPU micro-experiment: row-shifted UMMA descriptors inside a 128B-swizzled tile (see csrc/dbg.cu)."""
import ctypes as C
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nb_asr_b200 import _lib
lib = _lib.load()
f = lib.nbasr_dbg_shift
f.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
f.restype = C.c_int
torch.manual_seed(0)
dev = 'cuda:0'
for mode in (0, 1):
    if mode == 0:
        A = torch.randn(144, 64, device=dev).bfloat16()
        B = torch.randn(64, 64, device=dev).bfloat16()
    else:
        A = torch.randn(64, 128, device=dev).bfloat16()     # Y[k][m]
        B = torch.randn(80, 64, device=dev).bfloat16()      # X[k][n]
    for bo in (0, 1):
        res = []
        for shift in range(0, 13):
            D = torch.zeros(128, 64, device=dev)
            rc = f(A.data_ptr(), B.data_ptr(), D.data_ptr(), mode, shift, bo, None)
            torch.cuda.synchronize()
            if mode == 0:
                ref = A[shift:shift + 128].float() @ B.float().t()
            else:
                ref = A.float().t() @ B[shift:shift + 64].float()
            err = float((D - ref).norm() / ref.norm())
            res.append(f'{shift}:{err:.1e}')
        print(f'mode {mode} base_offset_mode {bo}:', ' '.join(res), flush=True)

# mode 2: no-swizzle K-major B operand (N=16, K=64); "shift" selects the LBO/SBO assignment under test
A = torch.randn(128, 64, device=dev).bfloat16()
Bm = torch.randn(16, 64, device=dev).bfloat16()
for variant in (0, 1):
    buf = torch.zeros(128 * 64 + 16 * 32, device=dev)          # D followed by the bf16 B matrix
    buf[128 * 64:].view(torch.bfloat16)[:16 * 64] = Bm.flatten()
    rc = f(A.data_ptr(), None, buf.data_ptr(), 2, variant, 0, None)
    torch.cuda.synchronize()
    D = buf[:128 * 64].view(128, 64)[:, :16]
    ref = A.float() @ Bm.float().t()
    print(f'mode 2 (no-swizzle B) variant {variant} [0: LBO=K-stride 256, SBO=row-group 128 | 1: swapped]: err {float((D - ref).norm() / ref.norm()):.2e}', flush=True)

# modes 3/4/5: mixed operand formats inside kind::f16 (A,B) = (f16,bf16) / (bf16,f16) / (f16,f16)
for mode, (ta, tb) in ((3, (torch.float16, torch.bfloat16)), (4, (torch.bfloat16, torch.float16)), (5, (torch.float16, torch.float16))):
    A = torch.randn(144, 64, device=dev).to(ta)
    B = torch.randn(64, 64, device=dev).to(tb)
    D = torch.zeros(128, 64, device=dev)
    rc = f(A.data_ptr(), B.data_ptr(), D.data_ptr(), mode, 0, 0, None)
    torch.cuda.synchronize()
    ref = A[:128].float() @ B.float().t()
    print(f'mode {mode} A={ta} B={tb}: err {float((D - ref).norm() / ref.norm()):.2e}', flush=True)
