"""CPU: host-side logic of the package (builder, init parity, state_dict layout, encoder, search space)."""
import hashlib
import json

import numpy as np
import pytest
import torch

import nb_asr_b200 as nb
from conftest import golden_path
from nb_asr_b200 import search_space as ss
from nb_asr_b200.model import pad_rule
from oracle import model_ref as M


def _digest(sd):
    h = hashlib.sha256()
    for k in sd:
        h.update(k.encode())
        h.update(sd[k].detach().contiguous().numpy().tobytes())
    return h.hexdigest()


@pytest.mark.parametrize('name', ['default', 'linear_skips', 'zero_mix'])
def test_same_seed_init_is_bit_identical_to_reference(name, golden_meta):
    meta = golden_meta[f'small_{name}']
    nb.set_seed(1235)
    model = nb.get_model(meta['arch'], use_rnn=True, dropout_rate=0.0)
    assert _digest(model.state_dict()) == meta['sd_digest']
    assert sum(p.numel() for p in model.parameters()) == meta['n_params']
    sd = M.build_state_dict(meta['arch'], seed=1235)
    assert list(sd.keys()) == list(model.state_dict().keys())


def test_model_attributes_and_errors():
    m = nb.get_model([[1, 0], [1, 0, 0], [1, 0, 0, 0]], use_rnn=True, dropout_rate=0.2)
    assert m.backend == 'b200' and m.num_classes == 48 and m.use_rnn and m.use_norm and m.dropout_rate == 0.2
    assert m.arch_desc == [['conv5', 0], ['conv5', 0, 0], ['conv5', 0, 0, 0]]
    assert len(m.model) == 29
    assert sum(isinstance(l, nb.PadConvRelu) for l in m.modules()) == 4 + 54
    with pytest.raises(ValueError):
        nb.get_model([[6, 0], [1, 0, 0], [1, 0, 0, 0]], True, 0.0)
    with pytest.raises(ValueError):
        nb.get_model([[1, 2], [1, 0, 0], [1, 0, 0, 0]], True, 0.0)
    with pytest.raises(RuntimeError):          # no CPU fallback: the product path fails loudly
        m(torch.zeros(1, 80, 16))
    m2 = nb.get_model([[5, 0], [5, 0, 0], [5, 0, 0, 0]], use_rnn=False, dropout_rate=0.0)
    assert sum(p.numel() for p in m2.parameters()) < 24_000_000


def test_pad_rule_matches_reference_table():
    assert pad_rule(8, 1, 1) == (3, 4) and pad_rule(8, 1, 2) == (5, 2)
    assert [pad_rule(k, d, 1) for k, d in ((5, 1), (5, 2), (7, 1), (7, 2))] == [(0, 4), (4, 4), (2, 4), (8, 4)]


def test_encoder_fold_lut_and_tables():
    g = json.load(open(golden_path('fold_lut.json')))
    enc = nb.PhonemeEncoder(48)
    assert enc.fold_lut(39).tolist() == g['lut']
    assert enc.get_vocab(inc_blank=True) == g['vocab48']
    assert {int(k): v for k, v in g['idx_mapping'].items()} == enc.idx_mappings[1][2]
    x = torch.arange(49, dtype=torch.int32)
    assert enc.fold_encoded(x, 39).tolist() == g['lut']
    assert len(enc.get_vocab(num_classes=61)) == 61 and len(enc.get_vocab(num_classes=39)) == 39


def test_search_space_enumeration():
    archs = list(ss.get_all_architectures())
    assert len(archs) == 13824
    assert archs[0] == [[0, 0], [0, 0, 0], [0, 0, 0, 0]] and archs[1] == [[1, 0], [0, 0, 0], [0, 0, 0, 0]]
    assert archs[-1] == [[5, 1], [5, 1, 1], [5, 1, 1, 1]]
    assert ss.arch_vec_to_names([[1, 0], [4, 0, 1], [5, 1, 0, 1]]) == [['conv5', 0], ['conv7d2', 0, 1], ['zero', 1, 0, 1]]


def test_avg_meter_matches_reference_formula():
    m = nb.AvgMeter()
    vals = [0.5, 0.25, 1.0, 0.125]
    for v in vals:
        m.update(v)
    ref = vals[0]
    for n, v in enumerate(vals[1:], 1):
        ref = ref * (n / (n + 1)) + v / (n + 1)
    assert m.get() == ref


def test_graph_hash_matches_reference_golden_and_kats():
    """graph_utils.get_model_hash vs hashes produced by the real reference (oracle/make_golden_graph.py) + README KATs"""
    import json
    import os
    from nb_asr_b200 import graph_utils as G
    gold = json.load(open(os.path.join(os.path.dirname(__file__), 'golden', 'graph_hashes.json')))
    for row in gold['rows']:
        assert G.get_model_hash(row['arch']) == row['hash'], row['arch']
    assert G.get_model_hash([[1, 0], [1, 0, 0], [1, 0, 0, 0]]) == '36855332a5778e0df5114305bc3ce238'      # README.md:61
    uniq = G.get_unique_architectures()
    assert gold['n_all'] == 13824 and len(uniq) == gold['n_unique'] == 8242                                # graph_utils.py:365-379
    assert sum(1 for _, a in uniq if all(n[0] != 5 for n in a)) == 8000
    # isomorphic pair: a `zero` node cuts the chain, the skip keeps the path -> same graph as dropping that node's op
    assert G.get_model_hash([[5, 1], [1, 0, 0], [1, 0, 0, 0]]) != G.get_model_hash([[1, 0], [1, 0, 0], [1, 0, 0, 0]])


def test_timit_shaped_eval_set():
    """BASELINE.json configs[3] / SURVEY.md 8d: 1 344 utterances, clipped log-normal lengths (mean ~306 in [92, 778]),
    targets ~ frames / 8 in [10, 75], sorted into batches of 64 padded to a multiple of 64 frames; all alignments feasible."""
    from nb_asr_b200 import data
    bs = data.timit_shaped_eval_set()
    assert len(bs) == 21 and sum(b[0][1].numel() for b in bs) == 1344
    lens = torch.cat([b[0][1] for b in bs]).float()
    assert 92 <= lens.min() and lens.max() <= 778 and 290 < lens.mean() < 325
    assert torch.equal(lens, lens.sort().values)                        # bucketed by length
    assert len({b[0][0].shape[2] for b in bs}) <= 16                    # a handful of distinct padded shapes
    for (audio, al), (tg, tl) in bs:
        assert audio.shape[2] % 64 == 0 and audio.shape[2] >= int(al.max())
        assert float(audio[0, :, int(al[0]):].abs().sum()) == 0.0       # zero padded like collate_fn (timit.py:104)
        assert int(tl.min()) >= 1 and int(tl.max()) <= 75
        assert bool((2 * tl <= al // 4).all())                          # feasible even if every label repeats
        assert tg.dtype == torch.int32 and int(tg.max()) <= 48
    again = data.timit_shaped_eval_set()
    assert all(torch.equal(a[0][0], b[0][0]) and torch.equal(a[1][0], b[1][0]) for a, b in zip(bs, again))      # fixed set
