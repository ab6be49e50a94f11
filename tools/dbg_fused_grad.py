"""Per-parameter gradient difference between the fused-chain plan (NBASR_GCONV_CHAIN=1) and the default plan."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import nb_asr_b200 as nb
arch = [[4, 1], [1, 0, 1], [2, 1, 0, 1]] if not os.environ.get('ARCH') else eval(os.environ['ARCH'])
batch = nb.data.make_batch(3, 300, seed=3, min_len=150)
res = {}
MODE = os.environ.get('MODE', '1')
for fused in ('0', MODE):
    os.environ['NBASR_GCONV_CHAIN'] = fused
    nb.set_seed(1235)
    model = nb.get_model(arch, use_rnn=True, dropout_rate=0.0, gpu=0, precision='bf16')
    tr = nb.get_trainer((nb.PhonemeEncoder(48), None, None, None), nb.get_loss(), gpus=[0], save_dir=None, verbose=False)
    tr.model = tr._model = model
    tr.optimizer = nb.trainer.FusedAdam(model, lr=1e-4)
    model.train()
    l0, lp0, _ = tr.step(batch, training=True)
    res[fused] = (l0.item(), model.engine.flat_g.clone(), dict(model.engine.slices))
a, b = res[MODE], res['0']
print('loss', a[0], b[0])
rows = []
for name, (off, n) in b[2].items():
    ga, gb = a[1][off:off + n].double(), b[1][off:off + n].double()
    rel = float((ga - gb).norm() / gb.norm().clamp_min(1e-300))
    rows.append((rel, name, float(gb.norm())))
for rel, name, nrm in sorted(rows, reverse=True)[:int(os.environ.get('TOP', 6))]:
    print(f'{rel:10.3e}  |g|={nrm:9.3e}  {name}')
print('flat relerr', float((a[1].double() - b[1].double()).norm() / b[1].double().norm()))
