"""CTC prefix beam search: oracle sanity on CPU, CUDA kernel bit-exact against the oracle."""
import numpy as np
import pytest
import torch

from oracle import decode_np as D


def _logp(B, T, V, seed, peaky=3.0):
    g = np.random.default_rng(seed)
    x = g.normal(size=(B, T, V)) * peaky
    x = x - np.log(np.exp(x).sum(-1, keepdims=True))
    return x.astype(np.float32)


def test_beam_width_one_without_cutoff_is_not_worse_than_greedy_and_kats():
    # a peaked distribution: beam search and greedy agree
    T, V = 12, 6
    path = [0, 2, 2, 0, 3, 3, 3, 0, 0, 1, 0, 4]
    lp = np.full((1, T, V), np.log(1e-4), np.float32)
    for t, c in enumerate(path):
        lp[0, t, c] = np.log(1 - 5e-4)
    ol = np.array([T])
    assert D.beam_search(lp, ol, 12, 40)[0].tolist() == D.greedy_decode(lp, ol)[0].tolist() == [2, 3, 1, 4]
    # the textbook case where prefix merging beats the best single path: P(a) summed over alignments > P(blank path)
    lp = np.log(np.array([[[0.6, 0.4], [0.6, 0.4]]], np.float32))     # greedy: blank,blank -> ""; P("")=.36, P("a")=.64
    assert D.greedy_decode(lp, np.array([2]))[0].tolist() == []
    assert D.beam_search(lp, np.array([2]), 4, 0)[0].tolist() == [1]
    # repeated label needs a blank in between: "aa" only via a-blank-a
    lp = np.log(np.array([[[0.1, 0.9], [0.8, 0.2], [0.1, 0.9]]], np.float32))
    assert D.beam_search(lp, np.array([3]), 4, 0)[0].tolist() == [1, 1]
    # empty utterance
    assert D.beam_search(lp, np.array([0]), 4, 0)[0].tolist() == []


def test_beam_probability_mass_is_exact_on_a_tiny_alphabet():
    """with a beam wider than the number of prefixes the search is exact: compare with brute-force path enumeration"""
    import itertools
    T, V = 5, 3
    lp = _logp(1, T, V, 3, peaky=1.0)
    probs = {}
    for path in itertools.product(range(V), repeat=T):
        p = float(np.exp(sum(np.float64(lp[0, t, c]) for t, c in enumerate(path))))
        lab, prev = [], -1
        for c in path:
            if c != prev and c != 0:
                lab.append(c)
            prev = c
        probs[tuple(lab)] = probs.get(tuple(lab), 0.0) + p
    best = max(probs.items(), key=lambda kv: kv[1])[0]
    assert tuple(D.beam_search(lp, np.array([T]), 16, 0)[0].tolist()) == best


@pytest.mark.gpu
@pytest.mark.parametrize('B,T,V,width,topn,peaky', [(6, 40, 49, 12, 40, 3.0), (4, 125, 49, 12, 40, 1.0), (3, 30, 49, 4, 0, 2.0),
                                                    (2, 17, 7, 16, 3, 0.5)])
def test_gpu_beam_search_bit_exact_vs_oracle(B, T, V, width, topn, peaky):
    import ctypes as C
    from nb_asr_b200 import _lib
    lib = _lib.load()
    dev = 'cuda:0'
    lp = _logp(B, T, V, 11 + T, peaky)
    g = np.random.default_rng(5)
    ol = g.integers(T // 2, T + 1, size=B)
    ol[0] = T
    S = 20
    tl = g.integers(3, S, size=B)
    tg = np.zeros((B, S), np.int32)
    for b in range(B):
        tg[b, :tl[b]] = g.integers(1, V, size=tl[b])
    lut = torch.from_numpy(D.FOLD_LUT[:V].copy() if V == 49 else np.arange(V, dtype=np.int32)).to(dev)
    d_lp, d_ol = torch.from_numpy(lp).to(dev), torch.from_numpy(ol.astype(np.int64)).to(dev)
    d_tg, d_tl = torch.from_numpy(tg).to(dev), torch.from_numpy(tl.astype(np.int64)).to(dev)
    raw = torch.zeros(B, T, dtype=torch.int32, device=dev); raw_len = torch.zeros(B, dtype=torch.int32, device=dev)
    hyp = torch.zeros(B, T, dtype=torch.int32, device=dev); hyp_len = torch.zeros(B, dtype=torch.int32, device=dev)
    dist = torch.zeros(B, dtype=torch.int32, device=dev); per = torch.zeros(2, dtype=torch.float64, device=dev)
    work = torch.zeros(B * (S + 2) + 8, dtype=torch.int32, device=dev)
    _lib.check(lib.nbasr_beam_per(d_lp.data_ptr(), B, T, V, d_ol.data_ptr(), 1, width, topn, d_tg.data_ptr(), S, d_tl.data_ptr(),
                                  lut.data_ptr(), raw.data_ptr(), raw_len.data_ptr(), hyp.data_ptr(), hyp_len.data_ptr(),
                                  dist.data_ptr(), per.data_ptr(), work.data_ptr(), torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    ref = D.beam_search(lp, ol, width, topn)
    for b in range(B):
        assert raw[b, :int(raw_len[b])].cpu().tolist() == ref[b].tolist(), b
    if V == 49:
        rper, rd, _, fh = D.per_from_hyps(ref, tg, tl)
        assert dist.cpu().tolist() == rd.tolist() and float(per[0]) == rper
        for b in range(B):
            assert hyp[b, :int(hyp_len[b])].cpu().tolist() == fh[b].tolist()


@pytest.mark.gpu
def test_trainer_decode_beam_option():
    import nb_asr_b200 as nb
    B, T, V = 4, 50, 49
    lp = torch.from_numpy(_logp(B, T, V, 2)).cuda()
    ol = torch.full((B,), T)
    tg = torch.randint(1, 49, (B, 12), dtype=torch.int32); tl = torch.full((B,), 12)
    tr = nb.get_trainer((nb.PhonemeEncoder(48), None, None, None), nb.get_loss(), gpus=[0], verbose=False)
    per_g = tr.decode(lp, ol, (None, (tg, tl)))
    per_b = tr.decode(lp, ol, (None, (tg, tl)), beam_width=12)
    ref = D.beam_search(lp.cpu().numpy(), ol.numpy(), 12, 40)
    assert per_b.item() == D.per_from_hyps(ref, tg.numpy(), tl.numpy())[0]
    assert per_g.item() == D.per_batch(lp.cpu().numpy(), ol.numpy(), tg.numpy(), tl.numpy())[0]
