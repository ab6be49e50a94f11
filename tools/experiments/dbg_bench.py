"""GPU micro-benchmarks (csrc/dbg.cu): tcgen05.mma cycles vs N, tcgen05.ld throughput, shuffle throughput."""
import ctypes as C
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nb_asr_b200 import _lib
lib = _lib.load()
f = lib.nbasr_dbg_bench
f.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
f.restype = C.c_int
dev = 'cuda:0'
out = torch.zeros(4096, dtype=torch.int64, device=dev)


def run(what, N, iters, variant, ctas=1):
    out.zero_()
    for _ in range(2):
        rc = f(out.data_ptr(), what, N, iters, variant, ctas, None)
        assert rc == 0
        torch.cuda.synchronize()
    return out.cpu()


iters = 3000
for variant in (0, 1, 2, 3):
    for N in (16, 32, 48, 64, 96, 128, 192, 240, 256):
        if (variant & 1) and N > 256:
            continue
        o = run(0, N, iters, variant)
        print(f'mma 128x{N}x16 variant {variant} (bit0 alt-acc, bit1 tap/k walk): {o[0].item() / iters:7.1f} cyc/mma total, {o[1].item() / iters:6.1f} cyc/mma issue', flush=True)
for nw in (1, 4, 8, 16):
    o = run(1, nw, 2000, 0)
    cyc = o[:nw].max().item() / 2000
    print(f'tmem_ld 32x32b.x32 with {nw} warps: {cyc:6.1f} cyc per round ({nw * 4096 / cyc:6.1f} B/clk/SM)', flush=True)
for nw in (1, 4, 8, 16):
    o = run(2, nw, 500, 0)
    cyc = o[:nw].max().item() / (500 * 32)
    print(f'shfl+fadd with {nw} warps: {cyc:6.2f} cyc per warp-shuffle per warp -> {nw / cyc:5.2f} warp-shuffles/clk/SM', flush=True)
