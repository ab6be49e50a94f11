"""Time the optimiser tail pieces (CUDA events): nbasr_optim_step vs the operand re-pack."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import nb_asr_b200 as nb
from nb_asr_b200 import _lib
nb.set_seed(1235)
model = nb.get_model([[1, 0], [1, 0, 0], [1, 0, 0, 0]], use_rnn=True, dropout_rate=0.0, gpu=0, precision='bf16')
eng = model.engine
eng.bind()
st = torch.cuda.current_stream().cuda_stream


def timeit(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def optim_only():
    _lib.check(eng.lib.nbasr_optim_step(eng.flat_p.data_ptr(), eng.flat_g.data_ptr(), eng.adam_m.data_ptr(), eng.adam_v.data_ptr(),
                                        eng.n_flat, eng.seg_off.data_ptr(), eng.seg_len.data_ptr(), int(eng.seg_off.numel()),
                                        eng.seg_chunks, 0.01, 5.0, 0.9, 0.999, 1e-7, eng.opt_state.data_ptr(), st))


print(f'params (flat, padded): {eng.n_flat/1e6:.2f} M; pack jobs {eng.pack_njobs}, blocks {eng.pack_blocks}')
print(f'nbasr_optim_step : {timeit(optim_only):8.1f} us')
print(f'refresh_packs    : {timeit(lambda: eng.refresh_packs(force=True)):8.1f} us')
