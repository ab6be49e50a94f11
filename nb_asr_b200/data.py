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


def timit_shaped_eval_set(n_utt=1344, batch_size=64, seed=0, pad_multiple=64, mean_frames=306.0, lo=92, hi=778):
    """The fixed synthetic evaluation set of BASELINE.json configs[3]: `n_utt` utterances (the size of
    TIMIT's test set without the SA sentences) with clipped log-normal lengths (mean ~306 frames in [92, 778], 10 ms hop),
    target lengths ~ frames / 8 in [10, 75], N(0, 1) log-mel stand-in, U{1..48} labels.  Utterances are sorted by length and
    cut into batches of `batch_size` (what a bucketing loader does); each batch is zero padded, like the reference
    collate_fn (timit.py:99-106), to its longest utterance rounded up to `pad_multiple` frames so that the sweep meets only
    a handful of distinct shapes.  Returns a list of ((audio, audio_len), (targets, targets_len)) CPU batches."""
    g = torch.Generator().manual_seed(seed)
    sigma = 0.45
    mu = torch.log(torch.tensor(mean_frames)) - sigma * sigma / 2          # E[lognormal] = exp(mu + sigma^2 / 2)
    lens = torch.exp(mu + sigma * torch.randn(n_utt, generator=g)).round().clamp(lo, hi).long()
    lens, _ = torch.sort(lens)
    batches = []
    for i in range(0, n_utt, batch_size):
        al = lens[i:i + batch_size].clone()
        B = al.numel()
        T = int((int(al.max()) + pad_multiple - 1) // pad_multiple * pad_multiple)
        audio = torch.randn(B, 80, T, generator=g)
        for b in range(B):
            audio[b, :, int(al[b]):] = 0.0
        tl = (al.float() / 8).round().clamp(10, 75).long()
        tl = torch.minimum(tl, al // 4 // 2).clamp(min=1)            # keep every alignment feasible
        tg = torch.randint(1, 49, (B, int(tl.max())), generator=g, dtype=torch.int32)
        for b in range(B):
            tg[b, int(tl[b]):] = 0
        batches.append(((audio, al), (tg, tl)))
    return batches
