"""-m gpu: the whole candidate train/eval step on the engine against golden fixtures of the REAL reference
(tests/golden, oracle/make_golden.py) and against the CPU oracle on the same seeded inputs."""
import json

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import nb_asr_b200 as nb  # noqa: E402
from conftest import golden_path  # noqa: E402
from oracle import decode_np as D  # noqa: E402
from oracle import model_ref as M  # noqa: E402

DEV = 'cuda:0'
SMALL = ['default', 'c7d2_skips', 'linear_skips', 'mixed', 'zero_mix', 'c5_c7']


def rel64(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm())


# Gradient tolerance.  Every backward kernel is checked in isolation at 1e-4..1e-5 (test_gpu_1_kernels).
# Through the WHOLE model a per-tensor bound of 2e-2 is the meaningful one: fp32 pre-activations differ from
# the CPU library's at the 1e-6 level (summation order), which flips the ReLU20 gate of a ~1e-6 fraction f of
# the elements whose pre-activation is ~0; each flip switches a gradient term on/off, so a weight-gradient
# (a random-sign sum) moves by ~sqrt(f) ~ 1e-3..1e-2 relative.  Direction and global norm are pinned tightly.
GRAD_TOL = 2e-2
BF16_LOGIT_TOL = 2e-2          # north_star: logits within 2e-2 relative in the 16-bit mode


def check_grads(model, raw, total):
    dot = n1 = n2 = 0.0
    for k, p in model.named_parameters():
        gref = raw[k].double()
        got = p.grad.double().cpu()
        err = float((got - gref).norm())
        assert err <= GRAD_TOL * float(gref.norm()) + 1e-7 * total, (k, err / max(float(gref.norm()), 1e-30))
        dot += float((got * gref).sum())
        n1 += float((got * got).sum())
        n2 += float((gref * gref).sum())
    assert dot / (n1 ** 0.5 * n2 ** 0.5) > 1 - 1e-4          # cosine of the full gradient vector
    assert abs(n1 ** 0.5 - n2 ** 0.5) < 2e-3 * n2 ** 0.5     # global gradient norm (what clip_grad_norm_ uses)


def build(arch, precision, dropout=0.0):
    nb.set_seed(1235)
    return nb.get_model(arch, use_rnn=True, dropout_rate=dropout, gpu=0, precision=precision)


def trainer(model):
    tr = nb.get_trainer((nb.PhonemeEncoder(48), None, None, None), nb.get_loss(), gpus=[0], save_dir=None, verbose=False)
    tr.model = tr._model = model
    tr.lr = 1e-4
    return tr


@pytest.mark.parametrize('name', SMALL)
def test_fp32_forward_matches_reference(name, golden_meta):
    meta = golden_meta[f'small_{name}']
    g = np.load(golden_path(f'eval_small_{name}.npz'))
    model = build(meta['arch'], 'fp32').eval()
    (audio, alen), (tg, tl) = nb.data.make_batch(meta['B'], meta['T'], seed=0, min_len=meta['T'] // 2)
    with torch.no_grad():
        logits = model(audio.to(DEV))
    ref = torch.from_numpy(g['logits'])
    assert rel64(logits, ref) < 1e-4, rel64(logits, ref)       # north_star: 1e-4 relative in fp32
    tr = trainer(model)
    loss, logp, out_len = tr.step(((audio, alen), (tg, tl)), training=False)
    assert abs(loss.item() - float(g['loss'])) < 1e-4 * abs(float(g['loss']))
    assert out_len.cpu().tolist() == g['out_len'].tolist()
    # decode on OUR log-probs: kernel vs numpy oracle, bit exact
    per = tr.decode(logp, out_len, ((audio, alen), (tg, tl)))
    rper, rd, _, rh = D.per_batch(logp.cpu().numpy(), out_len.cpu().numpy(), tg.numpy(), tl.numpy())
    hyp, hl, dist = tr.last_hyp
    assert dist.cpu().tolist() == rd.tolist()
    assert per.item() == rper
    for b in range(meta['B']):
        assert hyp[b, :int(hl[b])].cpu().tolist() == rh[b].tolist()


def test_fp32_survey_fixture_kat(golden_meta):
    meta = golden_meta['survey_c7d2_skips']
    g = np.load(golden_path('eval_survey_c7d2_skips.npz'))
    model = build(meta['arch'], 'fp32').eval()
    batch = nb.data.make_batch(8, 500, seed=0, min_len=250)
    tr = trainer(model)
    loss, logp, out_len = tr.step(batch, training=False)
    assert abs(loss.item() - 3.168433) < 2e-4
    with torch.no_grad():
        assert rel64(model(batch[0][0].to(DEV)), torch.from_numpy(g['logits'])) < 1e-4
    per = tr.decode(logp, out_len, batch)
    assert abs(per.item() - 4.490172) < 1e-6               # SURVEY.md §8c KAT: labels/PER bit-exact
    hyp, hl, _ = tr.last_hyp
    assert hl.cpu().tolist() == [109, 86, 58, 71, 89, 88, 96, 100]


@pytest.mark.parametrize('name', ['c7d2_skips', 'linear_skips', 'mixed', 'zero_mix', 'c5_c7', 'default'])
def test_fp32_train_step_matches_reference(name, golden_meta):
    meta = golden_meta[f'train_{name}']
    arch = meta['arch']
    model = build(arch, 'fp32').train()
    batch = nb.data.make_batch(meta['B'], meta['T'], seed=0, min_len=meta['T'] // 2)
    tr = trainer(model)
    tr.optimizer = nb.trainer.FusedAdam(model, lr=1e-4)
    # oracle gradients (CPU autograd over the restated forward)
    sd0 = M.build_state_dict(arch, seed=1235)
    (audio, alen), (tg, tl) = batch
    _, _, raw, total, sd1, _, _ = M.train_step(sd0, arch, audio, alen, tg, tl, None)
    eng = model.engine
    loss0, _, _ = tr.step(batch, training=True)
    assert abs(loss0.item() - meta['loss0']) < 1e-4 * abs(meta['loss0'])
    # flat_g now holds raw grads + regulariser (what clip_grad_norm_ saw); compare per-parameter
    coef = eng.opt_state[3].item()
    assert abs(coef - min(1.0, 5.0 / (total + 1e-6))) < 1e-3 * coef
    check_grads(model, raw, total)
    # parameters after Adam vs the REAL reference
    sd = model.state_dict()
    for k, s in meta['params0'].items():
        v = sd[k].double().cpu()
        assert abs(float(v.sum()) - s['sum']) <= 2e-4 * abs(s['sum']) + 2e-3, k
        assert np.allclose(v.flatten()[:8].numpy(), s['head'], rtol=2e-4, atol=2e-6), k
    loss1, _, _ = tr.step(batch, training=True)
    assert abs(loss1.item() - meta['loss1']) < 2e-3 * abs(meta['loss1'])


@pytest.mark.parametrize('name', ['c7d2_skips', 'linear_skips', 'mixed', 'c5_c7'])
def test_bf16_forward_within_tolerance(name, golden_meta):
    meta = golden_meta[f'small_{name}']
    g = np.load(golden_path(f'eval_small_{name}.npz'))
    model = build(meta['arch'], 'bf16').eval()
    (audio, alen), (tg, tl) = nb.data.make_batch(meta['B'], meta['T'], seed=0, min_len=meta['T'] // 2)
    with torch.no_grad():
        logits = model(audio.to(DEV))
    ref = torch.from_numpy(g['logits'])
    # north_star asks 2e-2 relative in the 16-bit mode.  bf16 storage of every activation cannot meet it on this 80-layer
    # net (rounding the reference's OWN activations / operands to bf16 on the CPU gives 2.1e-2 .. 3.7e-2), so
    # forward activations and their weight operands are fp16 (scaled by 32), gradients bf16: 0.4e-2 .. 1.0e-2 simulated.
    assert rel64(logits, ref) < BF16_LOGIT_TOL, rel64(logits, ref)
    tr = trainer(model)
    loss, logp, out_len = tr.step(((audio, alen), (tg, tl)), training=False)
    assert abs(loss.item() - float(g['loss'])) < 2e-3 * abs(float(g['loss']))
    per = tr.decode(logp, out_len, ((audio, alen), (tg, tl)))
    rper, rd, _, _ = D.per_batch(logp.cpu().numpy(), out_len.cpu().numpy(), tg.numpy(), tl.numpy())
    assert per.item() == rper and tr.last_hyp[2].cpu().tolist() == rd.tolist()


def test_bf16_train_step_close_to_fp32(golden_meta):
    meta = golden_meta['train_mixed']
    model = build(meta['arch'], 'bf16').train()
    batch = nb.data.make_batch(meta['B'], meta['T'], seed=0, min_len=meta['T'] // 2)
    tr = trainer(model)
    tr.optimizer = nb.trainer.FusedAdam(model, lr=1e-4)
    l0, _, _ = tr.step(batch, training=True)
    l1, _, _ = tr.step(batch, training=True)
    assert abs(l0.item() - meta['loss0']) < 2e-2 * abs(meta['loss0'])
    assert abs(l1.item() - meta['loss1']) < 5e-2 * abs(meta['loss1'])


def test_autograd_path_and_state_dict_roundtrip(golden_meta, tmp_path):
    arch = golden_meta['small_mixed']['arch']
    model = build(arch, 'fp32').train()
    (audio, alen), (tg, tl) = nb.data.make_batch(3, 70, seed=0, min_len=35)
    logits = model(audio.to(DEV))
    logp = torch.log_softmax(logits, 2)
    loss = nb.get_loss()(logp, (alen // 4).to(DEV), tg.to(DEV), tl.to(DEV))
    loss.backward()
    sd0 = M.build_state_dict(arch, seed=1235)
    p = {k: v.clone().requires_grad_(True) for k, v in sd0.items()}
    lr = M.ctc_loss_ref(torch.log_softmax(M.forward(p, arch, audio), 2), alen // 4, tg, tl)
    lr.backward()
    assert abs(loss.item() - lr.item()) < 1e-4 * abs(lr.item())
    raw = {k: (p[k].grad if p[k].grad is not None else torch.zeros_like(p[k])) for k in p}
    check_grads(model, raw, float(sum((g.double() ** 2).sum() for g in raw.values())) ** 0.5)
    # checkpoint interchange: reference-shaped keys and shapes
    tr = trainer(model)
    tr.optimizer = nb.trainer.FusedAdam(model, lr=1e-4)
    tr.save(tmp_path / 'x.ckpt')
    st = torch.load(tmp_path / 'x.ckpt', map_location='cpu')
    assert list(st['model'].keys()) == list(sd0.keys())
    assert all(st['model'][k].shape == sd0[k].shape for k in sd0)
    for k in sd0:
        assert torch.equal(st['model'][k], sd0[k]), k


def test_dropout_training_runs_and_differs():
    model = build([[0, 1], [2, 1, 0], [5, 1, 0, 1]], 'bf16', dropout=0.2).train()
    batch = nb.data.make_batch(2, 64, seed=1, min_len=40)
    tr = trainer(model)
    tr.optimizer = nb.trainer.FusedAdam(model, lr=1e-4)
    l0, lp0, _ = tr.step(batch, training=True)
    l1, lp1, _ = tr.step(batch, training=True)
    assert torch.isfinite(l0) and torch.isfinite(l1)
    model.eval()
    a, b, _ = tr.step(batch, training=False)
    c, d, _ = tr.step(batch, training=False)
    assert torch.equal(b, d)                                   # eval is deterministic (no dropout)


def test_train_loop_checkpoints_and_resume(tmp_path):
    """Trainer.train (trainer.py:80-206): warm-up epochs, validation with greedy PER, best/latest checkpoints, resume."""
    loaders = nb.get_dataloaders(batch_size=4, frames=96, n_train=2, n_val=1, n_test=1)
    nb.set_seed(1235)
    model = nb.get_model([[0, 1], [1, 1, 0], [5, 0, 1, 1]], use_rnn=True, dropout_rate=0.1, gpu=0)
    tr = nb.get_trainer(loaders, nb.get_loss(), gpus=[0], save_dir=tmp_path, verbose=False)
    val_scores, test_loss, test_per = tr.train(model, epochs=1, lr=1e-4, model_name='m')
    assert len(val_scores) == 1 and np.isfinite(test_loss) and np.isfinite(test_per)
    assert (tmp_path / 'm' / 'latest.ckpt').exists() and (tmp_path / 'm' / 'best.ckpt').exists()
    st = torch.load(tmp_path / 'm' / 'latest.ckpt', map_location='cpu')
    assert set(st) == {'model', 'optim'} and len(st['optim']['state']) == len(list(model.parameters()))
    # resume: a fresh model picks the checkpoint up (trainer.py:110-120)
    nb.set_seed(7)
    model2 = nb.get_model([[0, 1], [1, 1, 0], [5, 0, 1, 1]], use_rnn=True, dropout_rate=0.1, gpu=0)
    tr2 = nb.get_trainer(loaders, nb.get_loss(), gpus=[0], save_dir=tmp_path, verbose=False)
    tr2.train(model2, epochs=0, lr=1e-4, model_name='m')
    for k, v in model2.state_dict().items():
        assert torch.equal(v.cpu(), st['model'][k]), k


def test_long_mixed_length_utterances_fp32():
    """cfg 5 shape class: long utterances, mixed lengths, one infeasible alignment (zero_infinity), odd frame counts."""
    arch = [[3, 1], [0, 0, 1], [2, 1, 0, 0]]
    B, T = 3, 1203
    (audio, alen), (tg, tl) = nb.data.make_batch(B, T, seed=3, min_len=300, tgt_lo=40, tgt_hi=90)
    alen[1] = 301                                   # 75 output frames for >= 40 labels: may be infeasible with repeats
    tl[2] = min(int(tl[2]), tg.shape[1])
    alen[2] = 4 * 20                                # 20 frames < target length -> infeasible -> zero loss / zero grad
    model = build(arch, 'fp32').train()
    tr = trainer(model)
    tr.optimizer = nb.trainer.FusedAdam(model, lr=1e-4)
    sd0 = M.build_state_dict(arch, seed=1235)
    loss_ref, _, raw, total, _, _, logits_ref = M.train_step(sd0, arch, audio, alen, tg, tl, None)
    with torch.no_grad():
        model.eval()
        assert rel64(model(audio.to(DEV)), logits_ref) < 1e-4
        assert model(audio.to(DEV)).shape == (B, 301, 49)        # 1203 -> 602 -> 301
        model.train()
    loss, logp, out_len = tr.step(((audio, alen), (tg, tl)), training=True)
    assert abs(loss.item() - loss_ref.item()) < 1e-4 * abs(loss_ref.item())
    check_grads(model, raw, total)


@pytest.mark.parametrize('arch', [[[1, 0], [1, 0, 0], [1, 0, 0, 0]], [[4, 1], [0, 1, 1], [2, 0, 1, 1]]])
def test_gradient_buckets_partition_the_backward_pass(arch):
    """Data-parallel exchange: the backward plan is cut into segments after which a contiguous range of the flat
    gradient is final.  The ranges must tile the buffer, and a range must not change once its segment has run."""
    model = build(arch, 'bf16').train()
    eng = model.engine
    (audio, alen), (tg, tl) = nb.data.make_batch(4, 200, seed=3, min_len=120)
    tr = trainer(model)
    pl = eng.forward(audio.to(DEV), training=True)
    ws = tr._plan_ws(pl, torch.device(DEV), 4, tg.shape[1])
    tr._ctc(eng, pl, tg.to(DEV).int().contiguous(), alen.to(DEV), tl.to(DEV), True, ws)
    marks = [m for m, _, _ in pl.buckets]
    assert marks == sorted(marks) and marks[-1] == len(pl.bwd) and len(marks) == 5
    spans = sorted((lo, hi) for _, lo, hi in pl.buckets)
    assert spans[0][0] == 0 and spans[-1][1] == eng.n_flat
    assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    eng.backward(pl)
    torch.cuda.synchronize()
    full = eng.flat_g.clone()
    assert float(full.abs().sum()) > 0
    # segment by segment: after segment k, bucket k already holds its final value
    eng.flat_g.zero_()
    prev = 0
    for m, lo, hi in pl.buckets:
        eng._run(pl.bwd[prev:m])
        prev = m
        torch.cuda.synchronize()
        got, ref = eng.flat_g[lo:hi], full[lo:hi]
        # fp32 atomics make the weight-gradient sums run-to-run different in the last bits
        assert float((got - ref).norm()) <= 1e-4 * float(ref.norm()) + 1e-12, (m, lo, hi)


def test_async_bucket_allreduce_single_process_is_a_noop():
    from nb_asr_b200.distributed import allreduce_mean_async
    t = torch.arange(8, dtype=torch.float32, device=DEV)
    w = allreduce_mean_async(t)
    w.wait()
    assert torch.equal(t.cpu(), torch.arange(8, dtype=torch.float32))
