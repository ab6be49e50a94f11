"""Phoneme encoder: TIMIT 61 -> 48 -> 39 folding tables and the class-index fold used inside PER.

Mirrors nasbench_asr/training/torch/encoder.py (PhonemeEncoder :10-84).  The folding table is the
standard Lee & Hon TIMIT map that the reference ships as training/timit_folding.txt (data, one
line per 61-set phoneme: `p61<TAB>p48<TAB>p39`, `q` is dropped).
"""
import numpy as np

_FOLDING = """aa aa aa|ae ae ae|ah ah ah|ao ao aa|aw aw aw|ax ax ah|ax-h ax ah|axr er er|ay ay ay|b b b|bcl vcl sil|ch ch ch|d d d|dcl vcl sil|dh dh dh|dx dx dx|eh eh eh|el el l|em m m|en en n|eng ng ng|epi epi sil|er er er|ey ey ey|f f f|g g g|gcl vcl sil|h# sil sil|hh hh hh|hv hh hh|ih ih ih|ix ix ih|iy iy iy|jh jh jh|k k k|kcl cl sil|l l l|m m m|n n n|ng ng ng|nx n n|ow ow ow|oy oy oy|p p p|pau sil sil|pcl cl sil|q  |r r r|s s s|sh sh sh|t t t|tcl cl sil|th th th|uh uh uh|uw uw uw|ux uw uw|v v v|w w w|y y y|z z z|zh zh sh"""


def _table():
    rows = []
    for line in _FOLDING.split('|'):
        parts = line.split(' ')
        parts += [''] * (3 - len(parts))
        rows.append(parts[:3])
    return rows


class PhonemeEncoder:
    all_encodings = [61, 48, 39]

    def __init__(self, num_classes, remove_folded=True):
        assert num_classes in self.all_encodings
        self.num_classes = num_classes
        self.class_idx = self.all_encodings.index(num_classes)
        self.remove_folded = remove_folded
        rows = _table()
        n = len(self.all_encodings)
        self.mappings = {s: {d: {r[s]: r[d] for r in rows} for d in range(s + 1, n)} for s in range(n - 1)}
        self.to_delete = {s: {d: {p for p, v in self.mappings[s][d].items() if not v} for d in range(s + 1, n)}
                          for s in range(n - 1)}
        self.encodeds = [sorted({r[i] for r in rows if r[i]}) for i in range(n)]
        self.idx_mappings = {}
        for s in range(n - 1):
            self.idx_mappings[s] = {}
            for d in range(s + 1, n):
                mp = {0: 0}
                for si, ph in enumerate(self.encodeds[s]):
                    dp = self.mappings[s][d][ph]
                    mp[si + 1] = self.encodeds[d].index(dp) + 1 if dp else 0
                self.idx_mappings[s][d] = mp

    def get_vocab(self, inc_blank=False, num_classes=None):
        ci = self.all_encodings.index(num_classes) if num_classes is not None else self.class_idx
        v = list(self.encodeds[ci])
        return ['_'] + v if inc_blank else v

    def fold_lut(self, num_classes):
        """Effective class-index LUT of fold_encoded: the reference remaps sequentially IN PLACE
        (encoder.py:71-72 `x[x==old]=new` for old = 0..N), so remaps chain; reproduce that."""
        n = self.num_classes + 1
        lut = np.arange(n, dtype=np.int32)
        if num_classes >= self.num_classes:
            return lut
        if num_classes not in self.all_encodings:
            raise ValueError(num_classes)
        for old, new in self.idx_mappings[self.class_idx][self.all_encodings.index(num_classes)].items():
            lut[lut == old] = new
        return lut

    def fold_encoded(self, encodeds, num_classes):
        """In-place fold of an integer tensor (torch or numpy), same result as the reference."""
        if num_classes >= self.num_classes:
            return encodeds
        lut = self.fold_lut(num_classes)
        try:
            import torch
            if isinstance(encodeds, torch.Tensor):
                t = torch.as_tensor(lut, device=encodeds.device, dtype=encodeds.dtype)
                encodeds.copy_(t[encodeds.long()])
                return encodeds
        except ImportError:
            pass
        encodeds[...] = lut[encodeds]
        return encodeds

    def _fold(self, phonemes, dst_class_idx=None):
        d = self.class_idx if dst_class_idx is None else dst_class_idx
        if d == 0:
            return phonemes
        return [self.mappings[0][d][p] for p in phonemes if not self.remove_folded or p not in self.to_delete[0][d]]

    def encode(self, phonemes):
        folded = self._fold(phonemes)
        return [self.encodeds[self.class_idx].index(p) + 1 if p else 0 for p in folded]

    def decode(self, encodeds):
        return [self.encodeds[self.class_idx][i - 1] if i else '' for i in encodeds]
