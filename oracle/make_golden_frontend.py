"""Generate tests/golden/frontend.npz with the REAL transform chain of the reference (timit.py:90-95):
torchaudio MelSpectrogram -> torch.log -> reference normalisation statistics.  Run in the build container
(needs torchaudio and /root/reference); the fixture travels to the GPU box, this script does not need to.

    python oracle/make_golden_frontend.py
torchaudio here is 2.11 (the reference pins 0.7.0, setup.py:46, which cannot be installed offline); the MelSpectrogram
defaults used by the reference are identical in both (see oracle/frontend_np.py)."""
import os
import sys

import numpy as np
import torch
import torchaudio

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get('NBASR_REFERENCE', '/root/reference')


def main():
    stats = np.load(os.path.join(REF, 'nasbench_asr', 'training', 'timit_train_stats.npz'))
    mean, var = stats['moving_mean'].astype(np.float32), stats['moving_variance'].astype(np.float32)
    g = torch.Generator().manual_seed(7)
    lens = [9731, 4000, 12800, 481]
    wavs = []
    for i, n in enumerate(lens):
        t = torch.arange(n) / 16000.0
        w = 0.3 * torch.sin(2 * np.pi * (220.0 * (i + 1)) * t) + 0.05 * torch.randn(n, generator=g) + 0.1 * torch.sin(2 * np.pi * 3100.0 * t)
        wavs.append(w.float())
    tf = torchaudio.transforms.MelSpectrogram(sample_rate=16000, win_length=400, hop_length=160, n_mels=80)
    feats = []
    for w in wavs:
        f = torch.log(tf(w[None]))                                            # (1, 80, T)
        feats.append(((f - torch.from_numpy(mean)[None, :, None]) / (torch.from_numpy(var)[None, :, None] + 0.001))[0])
    T = max(f.shape[-1] for f in feats)
    out = torch.zeros(len(wavs), 80, T)
    for i, f in enumerate(feats):
        out[i, :, :f.shape[-1]] = f                                           # pad_sequence_bft, padding_value 0.0
    L = max(lens)
    wav = np.zeros((len(lens), L), np.float32)
    for i, w in enumerate(wavs):
        wav[i, :lens[i]] = w.numpy()
    np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'frontend.npz'), wav=wav, lens=np.array(lens, np.int64),
                        logmel=out.numpy(), frames=np.array([f.shape[-1] for f in feats], np.int64), mean=mean, var=var,
                        torchaudio_version=np.array(torchaudio.__version__))
    print('wrote tests/golden/frontend.npz', out.shape, 'torchaudio', torchaudio.__version__)


if __name__ == '__main__':
    main()
