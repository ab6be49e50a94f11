"""LayerNorm fwd/bwd at the benchmark shape (ncu target / micro-benchmark): time per launch with a cold L2."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
from nb_asr_b200 import _lib
from nb_asr_b200._lib import BF16
import gpu_utils as U
lib = _lib.load()
B, T = int(os.environ.get("B", 64)), 500
once = bool(os.environ.get('ONCE'))
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=U.DEV)
for Cc in ((800,) if once else (600, 800, 1000, 1200)):
    x = U.to_padded(torch.randn(B, T, Cc), BF16)
    dy = U.to_padded(torch.randn(B, T, Cc), BF16)
    y, dx, dx2 = (U.empty_padded(B, T, Cc, BF16) for _ in range(3))
    mean = torch.zeros(x.shape[0], device=U.DEV); rstd = torch.zeros(x.shape[0], device=U.DEV)
    g = torch.randn(Cc, device=U.DEV); b = torch.randn(Cc, device=U.DEV)
    dg = torch.zeros(Cc, device=U.DEV); db = torch.zeros(Cc, device=U.DEV)
    mwid = 40 if Cc == 1000 else 48
    mask = U.new_mask(x.shape[0], Cc, mwid)
    mask.random_(0, 255)
    fwd = lambda: lib.nbasr_layernorm_fwd(BF16, x.data_ptr(), y.data_ptr(), B, T, U.geo(T), Cc, g.data_ptr(), b.data_ptr(), 1e-3, mean.data_ptr(), rstd.data_ptr(), 1.0, None, U.stream())
    bwd1 = lambda: lib.nbasr_layernorm_bwd(BF16, dy.data_ptr(), x.data_ptr(), BF16, 1.0, mean.data_ptr(), rstd.data_ptr(), g.data_ptr(), B, T, U.geo(T), Cc, dx.data_ptr(), None, None, 1.0, 0, 32, dg.data_ptr(), db.data_ptr(), U.stream())
    bwd2 = lambda: lib.nbasr_layernorm_bwd(BF16, dy.data_ptr(), x.data_ptr(), BF16, 1.0, mean.data_ptr(), rstd.data_ptr(), g.data_ptr(), B, T, U.geo(T), Cc, dx.data_ptr(), dx2.data_ptr(), mask.data_ptr(), 1.0, mask.shape[1], mwid, dg.data_ptr(), db.data_ptr(), U.stream())
    el = B * T * Cc * 2
    for name, fn, passes in (('fwd', fwd, 2), ('bwd dx', bwd1, 3), ('bwd dx+dx2', bwd2, 4)):
        ts = []
        for it in range(2 if once else 6):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); _lib.check(fn()); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        t = sorted(ts[1:])[len(ts[1:]) // 2]
        print(f'C={Cc} {name:12s} {t*1e3:7.1f} us  {passes*el/t/1e6:6.0f} GB/s ({passes} passes)', flush=True)
