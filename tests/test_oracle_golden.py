"""Pin the oracle (oracle/) against fixtures produced by the REAL reference (oracle/make_golden.py)."""
import hashlib
import json

import numpy as np
import pytest
import torch

from conftest import golden_path
from oracle import decode_np as D
from oracle import model_ref as M


def _digest(sd):
    h = hashlib.sha256()
    for k in sd:
        h.update(k.encode())
        h.update(sd[k].detach().contiguous().numpy().tobytes())
    return h.hexdigest()


SMALL = ['default', 'c7d2_skips', 'linear_skips', 'mixed', 'zero_mix', 'c5_c7']


@pytest.mark.parametrize('name', SMALL)
def test_init_and_forward_small(name, golden_meta):
    meta = golden_meta[f'small_{name}']
    sd = M.build_state_dict(meta['arch'], seed=1235)
    assert _digest(sd) == meta['sd_digest']          # same-seed init is bit-identical to the reference
    assert sum(v.numel() for v in sd.values()) == meta['n_params']
    g = np.load(golden_path(f'eval_small_{name}.npz'))
    audio, alen, targets, tl = M.make_batch(meta['B'], meta['T'], seed=0, min_len=meta['T'] // 2)
    with torch.no_grad():
        loss, logp, out_len, logits = M.eval_step(sd, meta['arch'], audio, alen, targets, tl)
    ref = torch.from_numpy(g['logits'])
    rel = (logits.double() - ref.double()).norm() / ref.double().norm()   # fp64: default arch logits are ~1e-31
    assert rel < 1e-5, rel
    assert abs(loss.item() - float(g['loss'])) < 1e-5 * abs(float(g['loss']))
    assert np.array_equal(out_len.numpy(), g['out_len'])
    # numpy fp64 CTC restatement agrees with the reference's loss
    l2 = D.ctc_mean_loss(logp.numpy(), out_len.numpy(), targets.numpy(), tl.numpy())
    assert abs(l2 - float(g['loss'])) < 2e-5 * abs(float(g['loss']))


def test_forward_survey_fixture(golden_meta):
    name = 'c7d2_skips'
    meta = golden_meta[f'survey_{name}']
    sd = M.build_state_dict(meta['arch'], seed=1235)
    assert _digest(sd) == meta['sd_digest']
    g = np.load(golden_path(f'eval_survey_{name}.npz'))
    audio, alen, targets, tl = M.make_batch(8, 500, seed=0, min_len=250)
    with torch.no_grad():
        loss, logp, out_len, logits = M.eval_step(sd, meta['arch'], audio, alen, targets, tl)
    ref = torch.from_numpy(g['logits'])
    assert (logits - ref).norm() / ref.norm() < 1e-5
    assert abs(loss.item() - 3.168433) < 1e-5           # SURVEY.md §8c KAT
    assert out_len.tolist() == [125, 93, 65, 79, 111, 101, 111, 109]
    per, dists, rl, hyps = D.per_batch(logp.numpy(), out_len.numpy(), targets.numpy(), tl.numpy())
    assert abs(per - 4.490172) < 1e-6                     # SURVEY.md §8c KAT (greedy PER)
    assert [len(h) for h in D.greedy_decode(logp.numpy(), out_len.numpy())] == [109, 86, 58, 71, 89, 88, 96, 100]


@pytest.mark.parametrize('name', ['c7d2_skips', 'mixed', 'zero_mix'])
def test_train_step(name, golden_meta):
    meta = golden_meta[f'train_{name}']
    arch = meta['arch']
    sd = M.build_state_dict(arch, seed=1235)
    audio, alen, targets, tl = M.make_batch(meta['B'], meta['T'], seed=0, min_len=meta['T'] // 2)
    loss0, _, raw, total, sd1, st, _ = M.train_step(sd, arch, audio, alen, targets, tl, None)
    assert abs(loss0.item() - meta['loss0']) < 1e-5 * abs(meta['loss0'])
    coef = min(1.0, 5.0 / (total + 1e-6))
    for k, s in meta['grads_clipped'].items():
        n = float((raw[k].double() * coef).norm())
        assert abs(n - s['norm']) <= 2e-4 * s['norm'] + 1e-9, (k, n, s['norm'])
    for k, s in meta['params0'].items():
        assert abs(float(sd1[k].double().sum()) - s['sum']) <= 1e-4 * abs(s['sum']) + 1e-3, k
        assert np.allclose(sd1[k].double().flatten()[:8].numpy(), s['head'], rtol=1e-4, atol=1e-6), k
    loss1, *_ = M.train_step(sd1, arch, audio, alen, targets, tl, st)
    assert abs(loss1.item() - meta['loss1']) < 5e-4 * abs(meta['loss1'])


def test_fold_lut():
    g = json.load(open(golden_path('fold_lut.json')))
    assert g['lut'] == D.FOLD_LUT.tolist()
    assert D.chained_fold_lut({int(k): v for k, v in g['idx_mapping'].items()}).tolist() == g['lut']


def test_infeasible_ctc():
    g = np.load(golden_path('ctc_infeasible.npz'))
    l = D.ctc_mean_loss(g['logp'], g['out_len'], g['targets'], g['tgt_len'])
    assert abs(l - float(g['loss'])) < 1e-5 * abs(float(g['loss']))
    nll = D.ctc_nll(g['logp'], g['out_len'], g['targets'], g['tgt_len'])
    assert np.isinf(nll[0]) and np.isfinite(nll[1])


def test_levenshtein_kats():
    assert D.levenshtein([], [1, 2]) == 2
    assert D.levenshtein([11, 9, 20, 20, 5, 14], [19, 9, 20, 20, 9, 14, 7]) == 3
    assert D.levenshtein([1, 2, 3], []) == 3
    assert D.levenshtein([], []) == 0
