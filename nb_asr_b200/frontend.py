"""GPU log-mel front end + collate: the step either side of the hot path.

Mirrors the transform chain the reference builds in ``get_dataloaders`` (training/torch/timit.py:78-106):
MelSpectrogram(16 kHz, n_fft = win = 400, hop 160, 80 mels) -> log -> (x - mean) / (variance + 1e-3) -> zero padding to
the longest utterance of the batch, and returns the same ``(audio[B,80,T] fp32, audio_len[B])`` pair ``Trainer.step``
consumes -- but computed on the B200 by ``nbasr_logmel`` (csrc/frontend.cu: two fp32 GEMMs + glue kernels), so raw
waveforms can be fed to the train/eval step without a CPU feature pipeline.
"""
import math
import os

import numpy as np
import torch

from . import _lib

SAMPLE_RATE, NFFT, HOP, NMEL = 16000, 400, 160, 80


def _dft_matrix():
    n = torch.arange(NFFT // 2 + 1, dtype=torch.float64)[:, None]
    k = torch.arange(NFFT, dtype=torch.float64)[None, :]
    w = (0.5 - 0.5 * torch.cos(2.0 * math.pi * torch.arange(NFFT, dtype=torch.float64) / NFFT))[None, :]   # periodic Hann
    ang = 2.0 * math.pi * n * k / NFFT
    return torch.cat([w * torch.cos(ang), -w * torch.sin(ang)], 0).float().contiguous()                    # (402, 400)


def _mel_fb():
    """torchaudio create_fb_matrix(201, 0, 8000, 80, 16000, norm=None), HTK scale -> transposed, K padded to 208"""
    n_freqs = NFFT // 2 + 1
    all_freqs = torch.linspace(0, SAMPLE_RATE // 2, n_freqs, dtype=torch.float64)
    m_max = 2595.0 * math.log10(1.0 + (SAMPLE_RATE / 2.0) / 700.0)
    m_pts = torch.linspace(0.0, m_max, NMEL + 2, dtype=torch.float64)
    f_pts = 700.0 * (10.0 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts[None, :] - all_freqs[:, None]
    fb = torch.clamp(torch.minimum(-slopes[:, :-2] / f_diff[:-1], slopes[:, 2:] / f_diff[1:]), min=0.0)   # (201, 80)
    out = torch.zeros(NMEL, 208, dtype=torch.float32)
    out[:, :n_freqs] = fb.t().float()
    return out.contiguous()


class LogMelFrontend:
    """wav batch -> (audio (B, 80, T) fp32 on the GPU, audio_len (B,) int64 frames).

    ``stats``: path of the reference's ``timit_train_stats.npz`` (keys moving_mean / moving_variance), a
    ``(mean, variance)`` pair, or None for no normalisation shift (mean 0, variance 1 - eps)."""

    def __init__(self, device='cuda:0', stats=None, eps=1e-3):
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise RuntimeError('nb_asr_b200.frontend runs on a CUDA device only (no CPU fallback)')
        self.lib = _lib.load()
        self.eps = float(eps)
        if stats is None:
            mean, var = np.zeros(NMEL, np.float32), np.full(NMEL, 1.0 - eps, np.float32)
        elif isinstance(stats, (str, os.PathLike)):
            d = np.load(stats)
            mean, var = d['moving_mean'], d['moving_variance']
        else:
            mean, var = stats
        self.mean = torch.as_tensor(np.asarray(mean, np.float32)).to(self.device)
        self.var = torch.as_tensor(np.asarray(var, np.float32)).to(self.device)
        self.dft = _dft_matrix().to(self.device)
        self.melfb = _mel_fb().to(self.device)
        self._work = None

    def __call__(self, wav, lengths):
        """wav: (B, L) fp32 zero padded (any device) or a list of 1-D tensors; lengths: samples per utterance."""
        if isinstance(wav, (list, tuple)):
            lengths = torch.tensor([int(w.numel()) for w in wav], dtype=torch.int64)
            L = int(lengths.max())
            buf = torch.zeros(len(wav), L, dtype=torch.float32)
            for i, w in enumerate(wav):
                buf[i, :w.numel()] = w.reshape(-1).float()
            wav = buf
        lengths = torch.as_tensor(lengths, dtype=torch.int64)
        if int(lengths.min()) <= NFFT // 2:
            raise ValueError(f'utterances must be longer than {NFFT // 2} samples (reflect padding), got {int(lengths.min())}')
        wav = wav.to(self.device, dtype=torch.float32, non_blocking=True).contiguous()
        len_d = lengths.to(self.device, non_blocking=True)
        B, L = wav.shape
        T = 1 + int(lengths.max()) // HOP
        need = int(self.lib.nbasr_logmel_work_floats(B, L))
        if self._work is None or self._work.numel() < need:
            self._work = torch.empty(need, dtype=torch.float32, device=self.device)
        out = torch.empty(B, NMEL, 1 + L // HOP, dtype=torch.float32, device=self.device)
        st = torch.cuda.current_stream().cuda_stream
        _lib.check(self.lib.nbasr_logmel(wav.data_ptr(), len_d.data_ptr(), B, L, self.dft.data_ptr(), self.melfb.data_ptr(),
                                         self.mean.data_ptr(), self.var.data_ptr(), self.eps, out.data_ptr(), out.shape[2],
                                         self._work.data_ptr(), self._work.numel(), st), 'logmel')
        return out[:, :, :T], 1 + lengths // HOP


def collate_wav_batch(frontend, batch):
    """Reference collate_fn (timit.py:97-105) on raw waveforms: batch = [(wav 1-D tensor, [phoneme ids]), ...] ->
    ((audio, audio_len), (targets int32 zero padded, targets_len))."""
    wavs = [b[0] for b in batch]
    sents = [list(b[1]) for b in batch]
    audio, alen = frontend(wavs, None)
    tl = torch.tensor([len(s) for s in sents])
    S = int(tl.max())
    tg = torch.tensor([s + [0] * (S - len(s)) for s in sents], dtype=torch.int32)
    return (audio, alen), (tg, tl)
