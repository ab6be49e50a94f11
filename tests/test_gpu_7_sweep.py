"""-m gpu: the architecture-sweep path (BASELINE.json configs[3]): device-side initialisation, the TIMIT-shaped set,
one row per candidate.  Reference pieces: search_space.py:32-47, graph_utils.py:145-180, model/torch/__init__.py:13-29."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

import nb_asr_b200 as nb  # noqa: E402
from nb_asr_b200 import data, sweep  # noqa: E402
from oracle import decode_np as D  # noqa: E402

DEV = 'cuda:0'


def test_device_init_has_the_reference_distributions():
    arch = [[1, 0], [0, 1, 0], [5, 0, 1, 1]]
    model = nb.get_model(arch, use_rnn=True, dropout_rate=0.0, gpu=0, init='device', seed=7, conv_gain=3.0)
    model.engine.bind()
    ref = nb.get_model(arch, use_rnn=True, dropout_rate=0.0, gpu=0)                # reference procedure: same keys / shapes
    sd, rsd = model.state_dict(), ref.state_dict()
    assert list(sd.keys()) == list(rsd.keys())
    for k, v in sd.items():
        assert v.shape == rsd[k].shape and v.is_cuda, k
        leaf = k.rsplit('.', 1)[-1]
        if v.dim() >= 2:
            rf = math.prod(v.shape[2:])
            bound = math.sqrt(6.0 / (v.shape[1] * rf + v.shape[0] * rf)) * (3.0 if ('.nodes.' in k and k.endswith('.conv.weight')) else 1.0)
            assert float(v.abs().max()) <= bound * (1 + 1e-6), k
            # uniform(-a, a): std = a / sqrt(3), mean ~ 0
            assert abs(float(v.float().std()) / (bound / math.sqrt(3)) - 1) < 0.05, k
            assert abs(float(v.float().mean())) < 0.05 * bound, k
        elif leaf == 'weight':
            assert bool((v == 1).all()), k                                        # LayerNorm gamma
        else:
            assert bool((v == 0).all()), k                                        # biases, LayerNorm beta
    # same seed -> same weights; other seed -> different
    m2 = nb.get_model(arch, use_rnn=True, dropout_rate=0.0, gpu=0, init='device', seed=7, conv_gain=3.0)
    m3 = nb.get_model(arch, use_rnn=True, dropout_rate=0.0, gpu=0, init='device', seed=8, conv_gain=3.0)
    m2.engine.bind(); m3.engine.bind()
    assert torch.equal(m2.engine.flat_p, model.engine.flat_p) and not torch.equal(m3.engine.flat_p, model.engine.flat_p)


def test_sweep_row_equals_a_manual_eval_and_calibrated_gain_gives_nontrivial_per():
    batches = data.timit_shaped_eval_set(n_utt=96, batch_size=32, seed=0)
    n_ref = sum(int(tl.sum()) for _, (t, tl) in batches)
    dev_batches = [((a.to(DEV), al.to(DEV)), (t.to(DEV), tl.to(DEV))) for (a, al), (t, tl) in batches]
    arch = [[2, 1], [3, 0, 1], [0, 1, 0, 1]]
    row = sweep.evaluate_arch(arch, dev_batches, 0, 'fp32', seed=1235, init='reference', n_ref=n_ref)
    # the same numbers by hand: reference init, eval step, numpy-oracle decode on our log-probs
    nb.set_seed(1235)
    model = nb.get_model(arch, use_rnn=True, dropout_rate=0.0, gpu=0, precision='fp32').eval()
    tr = nb.get_trainer((nb.PhonemeEncoder(48), None, None, None), nb.get_loss(), gpus=[0], verbose=False)
    tr.model = tr._model = model
    losses, pers, dist = [], [], 0
    for b in batches:
        loss, logp, ol = tr.step(b, training=False)
        per, d, _, _ = D.per_batch(logp.cpu().numpy(), ol.cpu().numpy(), b[1][0].numpy(), b[1][1].numpy())
        losses.append(loss.item()); pers.append(per); dist += int(d.sum())
    assert abs(row['loss'] - sum(losses) / len(losses)) < 1e-6
    assert row['per'] == pytest.approx(sum(pers) / len(pers), abs=1e-12)
    assert row['per_corpus'] == pytest.approx(dist / n_ref, abs=1e-12)
    # skip-free conv arch: all-blank (PER = 1) at the reference init, non-trivial with the variance-preserving gain
    default = [[1, 0], [1, 0, 0], [1, 0, 0, 0]]
    r0 = sweep.evaluate_arch(default, dev_batches, 0, 'bf16', seed=1235, init='device', conv_gain=1.0, n_ref=n_ref)
    r1 = sweep.evaluate_arch(default, dev_batches, 0, 'bf16', seed=1235, init='device', conv_gain=101 ** 0.5, n_ref=n_ref)
    assert r0['per'] == 1.0
    assert r1['per'] != 1.0 and math.isfinite(r1['loss'])
