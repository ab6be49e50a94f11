"""Log-mel front end: oracle pinned to the reference transform chain, GPU kernel against both."""
import os

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(__file__), 'golden', 'frontend.npz')


def _fixture():
    d = np.load(GOLD)
    wavs = [d['wav'][i, :n] for i, n in enumerate(d['lens'])]
    return d, wavs


def test_oracle_matches_reference_transform_chain():
    """oracle/frontend_np.py vs torchaudio MelSpectrogram -> log -> reference normalisation (make_golden_frontend.py)"""
    from oracle import frontend_np as Fn
    d, wavs = _fixture()
    out, frames = Fn.logmel_batch(wavs, d['mean'], d['var'])
    assert frames.tolist() == d['frames'].tolist() == [1 + n // 160 for n in d['lens']]
    assert out.shape == d['logmel'].shape
    assert np.abs(out - d['logmel']).max() < 5e-5          # fp64 oracle vs the reference's fp32 FFT
    # zero padding beyond each utterance (collate_fn, timit.py:104)
    for i, f in enumerate(frames):
        assert np.all(out[i, :, f:] == 0.0)


def test_oracle_operands_match_host_side_tables():
    """the DFT / mel operands the GPU path uploads are the oracle's, to fp32 rounding"""
    from oracle import frontend_np as Fn
    from nb_asr_b200 import frontend as F
    assert np.abs(F._dft_matrix().numpy() - Fn.dft_matrix()).max() < 1e-6
    fb = F._mel_fb().numpy()
    assert fb.shape == (80, 208) and np.abs(fb[:, :201] - Fn.mel_filterbank().T).max() < 1e-6 and np.all(fb[:, 201:] == 0)


@pytest.mark.gpu
def test_gpu_logmel_matches_reference_and_oracle():
    from nb_asr_b200.frontend import LogMelFrontend
    from oracle import frontend_np as Fn
    d, wavs = _fixture()
    fe = LogMelFrontend('cuda:0', stats=(d['mean'], d['var']))
    audio, alen = fe(torch.from_numpy(d['wav']), torch.from_numpy(d['lens']))
    torch.cuda.synchronize()
    got = audio.cpu().numpy()
    assert alen.tolist() == d['frames'].tolist()
    assert got.shape == d['logmel'].shape
    # fp32 400-term DFT sums vs FFT: tolerance on the log-mel features (values span about [-4, 4])
    assert np.abs(got - d['logmel']).max() < 2e-3
    ref, _ = Fn.logmel_batch(wavs, d['mean'], d['var'])
    assert np.linalg.norm(got - ref) / np.linalg.norm(ref) < 1e-4
    for i, f in enumerate(d['frames']):
        assert np.all(got[i, :, f:] == 0.0)
    # list-of-waveforms entry point (what collate_wav_batch uses) gives the same features
    audio2, alen2 = fe([torch.from_numpy(w) for w in wavs], None)
    assert torch.equal(audio2, audio) and alen2.tolist() == alen.tolist()


@pytest.mark.gpu
def test_gpu_logmel_feeds_the_eval_step():
    """wav -> front end -> model eval step runs end to end on the GPU (shapes / dtypes of Trainer.step inputs)"""
    import nb_asr_b200 as nb
    from nb_asr_b200.frontend import LogMelFrontend, collate_wav_batch
    d, wavs = _fixture()
    fe = LogMelFrontend('cuda:0', stats=(d['mean'], d['var']))
    batch = collate_wav_batch(fe, [(torch.from_numpy(w), [3, 7, 11, 2]) for w in wavs])
    (audio, alen), (tg, tl) = batch
    assert audio.shape[1] == 80 and audio.dtype == torch.float32 and tg.dtype == torch.int32
    nb.set_seed(1235)
    model = nb.get_model([[2, 1], [3, 0, 1], [0, 1, 0, 1]], use_rnn=True, dropout_rate=0.0, gpu=0, precision='bf16')
    model.eval()
    tr = nb.get_trainer((nb.PhonemeEncoder(48), None, None, None), nb.get_loss(), gpus=[0], verbose=False)
    tr.model = tr._model = model
    loss, logp, out_len = tr.step(batch, training=False)
    assert torch.isfinite(loss) and out_len.tolist() == (alen // 4).tolist()
