"""One search cell's grouped-conv chain on one shape: node-by-node launches vs nbasr_gconv_chain (timing + ncu target).
   env: C (800), OPS (conv5,conv5,conv5), BWD (0), SKIPS (0: none, 1: every earlier output), REPS (5)"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
from nb_asr_b200 import _lib
from nb_asr_b200._lib import BF16, F16, GConv
import gpu_utils as U
import test_gpu_8_gconv_chain as TC
lib = _lib.load()
Cc = int(os.environ.get('C', 800))
ops = os.environ.get('OPS', 'conv5,conv5,conv5').split(',')
bwd = bool(int(os.environ.get('BWD', 0)))
reps = int(os.environ.get('REPS', 5))
B, T = 64, 500 if Cc <= 800 else (250 if Cc == 1000 else 125)
skips = [list(range(i + 1)) for i in range(len(ops))] if int(os.environ.get('SKIPS', 0)) else [[] for _ in ops]
nodes, outs, keep = TC._build(lib, Cc, B, T, ops, skips, BF16 if bwd else F16, bwd, seed=0)
n = len(nodes)
arr = (GConv * n)(*nodes)
wb = int(lib.nbasr_gconv_chain_work_bytes(B, T, Cc, Cc // 100, 3))
work = torch.zeros(wb // 4, dtype=torch.int32, device=U.DEV)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=U.DEV)
st = U.stream()


def timed(fn):
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return sorted(ts)[len(ts) // 2]


def old():
    for gc in nodes:
        _lib.check(lib.nbasr_gconv_fwd(C.byref(gc), st))


def chain():
    _lib.check(lib.nbasr_gconv_chain(arr, n, 1, work.data_ptr(), wb, st))


print(f'C={Cc} T={T} ops={ops} bwd={bwd}: node-by-node {timed(old):.1f} us, chain {timed(chain):.1f} us, work[2]={int(work[2])}')
