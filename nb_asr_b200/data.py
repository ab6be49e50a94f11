"""Synthetic batches with the structure of the reference collate_fn (training/torch/timit.py:99-106):
((audio (B,80,T) f32 zero padded, audio_len (B,)), (targets (B,S) i32 zero padded, targets_len (B,)))."""
import torch

from .encoder import PhonemeEncoder


def make_batch(B, T, seed=0, min_len=None, tgt_lo=10, tgt_hi=30, pin=False):
    g = torch.Generator().manual_seed(seed)
    audio = torch.randn(B, 80, T, generator=g)
    lo = T // 2 if min_len is None else min_len
    alen = torch.randint(lo, T + 1, (B,), generator=g)
    alen[0] = T
    for b in range(B):
        audio[b, :, int(alen[b]):] = 0.0
    tl = torch.randint(tgt_lo, tgt_hi, (B,), generator=g)
    targets = torch.randint(1, 49, (B, int(tl.max())), generator=g, dtype=torch.int32)
    for b in range(B):
        targets[b, int(tl[b]):] = 0
    if pin and torch.cuda.is_available():
        audio, alen, targets, tl = (t.pin_memory() for t in (audio, alen, targets, tl))
    return (audio, alen), (targets, tl)


class SyntheticLoader:
    def __init__(self, n_batches, B, T, seed=0, **kw):
        self.batches = [make_batch(B, T, seed=seed + i, **kw) for i in range(n_batches)]

    def __iter__(self):
        return iter(self.batches)

    def __len__(self):
        return len(self.batches)


def synthetic_dataloaders(batch_size=64, frames=500, n_train=4, n_val=2, n_test=2, seed=0):
    enc = PhonemeEncoder(48)
    return (enc, SyntheticLoader(n_train, batch_size, frames, seed),
            SyntheticLoader(n_val, batch_size, frames, seed + 1000),
            SyntheticLoader(n_test, batch_size, frames, seed + 2000))
