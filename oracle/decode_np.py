"""numpy restatement of CTC loss (log-space alpha), greedy CTC decode, 48->39 fold and PER.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Integer paths are bit-exact by construction.

Restated from /root/reference/nasbench_asr/:
  greedy decode ... training/tf/metrics/ctc.py:76-81 (argmax per frame, merge repeats, drop blank;
                    blank is class 0 in the torch layout, trainer.py:71,242)
  fold 48->39 ..... training/torch/encoder.py:64-74 -- remaps are applied sequentially IN PLACE,
                    so they chain; FOLD_LUT below is the *effective* table (pinned by
                    tests/golden/fold_lut.json, generated from the reference's PhonemeEncoder)
  PER ............. training/torch/trainer.py:245-246: compute_wer(hyp, ref, hyp_len, ref_len,
                    blank=[0], sep=[]) = Levenshtein(hyp\\blank, ref\\blank) / ref_len per
                    utterance (torch-edit-distance's documented behaviour), then batch mean
  AvgMeter ........ training/torch/trainer.py:16-33
"""
import numpy as np

FOLD_LUT = np.array([0, 1, 2, 3, 1, 4, 3, 5, 6, 7, 24, 8, 9, 10, 11, 15, 17, 24, 12, 13, 14, 15, 16, 17,
                     17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29, 30, 31, 32, 33, 34, 35, 36, 31,
                     37, 38, 39, 30], dtype=np.int32)


def chained_fold_lut(idx_mapping, n=49):
    """Effective LUT of encoder.py:71-72 (`x[x==old]=new` for old in mapping order)."""
    lut = np.arange(n, dtype=np.int32)
    for old, new in idx_mapping.items():
        lut[lut == old] = new
    return lut


def greedy_decode(logp, out_len):
    """logp (B,T,C) float; out_len (B,) -> list of int32 arrays (collapsed, blank-free)."""
    hyps = []
    for b in range(logp.shape[0]):
        best = np.argmax(logp[b, :int(out_len[b])], axis=1)      # first max on ties, like torch
        prev = -1
        seq = []
        for s in best:
            if s != prev and s != 0:
                seq.append(int(s))
            prev = s
        hyps.append(np.array(seq, dtype=np.int32))
    return hyps


def levenshtein(a, b):
    a = [int(x) for x in a]
    b = [int(x) for x in b]
    prev = list(range(len(b) + 1))
    for i, ca in enumerate(a, 1):
        cur = [i] + [0] * len(b)
        for j, cb in enumerate(b, 1):
            cur[j] = min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (ca != cb))
        prev = cur
    return prev[len(b)]


def per_batch(logp, out_len, targets, targets_len, fold=True):
    """-> (mean PER as python float (fp64 sequential mean), dists, ref_lens, folded hyps)."""
    hyps = greedy_decode(np.asarray(logp), out_len)
    dists, rlens, fh = [], [], []
    for b, h in enumerate(hyps):
        ref = np.asarray(targets[b, :int(targets_len[b])], dtype=np.int32)
        if fold:
            h = FOLD_LUT[h]
            ref = FOLD_LUT[ref]
        h = h[h != 0]
        ref = ref[ref != 0]
        dists.append(levenshtein(h, ref))
        rlens.append(int(targets_len[b]))
        fh.append(h)
    acc = 0.0
    for d, r in zip(dists, rlens):
        acc += float(d) / float(r)
    return acc / len(dists), np.array(dists, np.int32), np.array(rlens, np.int32), fh


def ctc_nll(logp, out_len, targets, targets_len):
    """Per-utterance CTC negative log-likelihood in float64 (blank 0). inf if infeasible."""
    B = logp.shape[0]
    out = np.zeros(B, dtype=np.float64)
    for b in range(B):
        T = int(out_len[b])
        lab = [int(x) for x in targets[b, :int(targets_len[b])]]
        ext = [0]
        for l in lab:
            ext += [l, 0]
        S = len(ext)
        lp = np.asarray(logp[b, :T], dtype=np.float64)
        alpha = np.full(S, -np.inf)
        if T > 0:
            alpha[0] = lp[0, 0]
            if S > 1:
                alpha[1] = lp[0, ext[1]]
        for t in range(1, T):
            new = np.full(S, -np.inf)
            for s in range(S):
                c = [alpha[s]]
                if s >= 1:
                    c.append(alpha[s - 1])
                if s >= 2 and ext[s] != 0 and ext[s] != ext[s - 2]:
                    c.append(alpha[s - 2])
                m = max(c)
                if m > -np.inf:
                    new[s] = m + np.log(sum(np.exp(x - m) for x in c)) + lp[t, ext[s]]
            alpha = new
        tail = [alpha[S - 1]] + ([alpha[S - 2]] if S > 1 else [])
        m = max(tail)
        out[b] = np.inf if m == -np.inf else -(m + np.log(sum(np.exp(x - m) for x in tail)))
    return out


def ctc_mean_loss(logp, out_len, targets, targets_len):
    """trainer.py:36-44 with zero_infinity=True."""
    nll = ctc_nll(logp, out_len, targets, targets_len)
    nll = np.where(np.isinf(nll), 0.0, nll)
    return float(np.mean(nll / np.asarray(out_len, dtype=np.float64)))


class AvgMeter:
    """trainer.py:16-33 -- unweighted running mean with the reference's update formula."""

    def __init__(self):
        self.avg, self.n = 0, 0

    def update(self, a):
        if not self.n:
            self.avg, self.n = a, 1
        else:
            self.avg = self.avg * (self.n / (self.n + 1)) + (a / (self.n + 1))
            self.n += 1

    def get(self):
        return self.avg


# ---------------------------------------------------------------------------------------------------------------------
# CTC prefix beam search -- what Trainer.decode really calls (trainer.py:71,236): ctcdecode.CTCBeamDecoder(labels,
# beam_width=12, log_probs_input=True).  The library is a third-party dependency that is NOT under /root/reference
# (parlance/ctcdecode, git 9a20e00f34d8f605f4a8501cc42b1a53231f1597, setup.py:49); its published algorithm (the
# PaddlePaddle DeepSpeech prefix beam search without a language model: defaults cutoff_top_n=40, cutoff_prob=1.0,
# blank_id=0) is restated below.  PARITY UNPINNED: the reference holds no test or golden vector for this boundary and
# the library cannot be run here.  Deliberate differences, shared with the CUDA kernel so the two agree bit-exactly:
# scores are fp64 (ctcdecode: fp32), ties are broken on (last label, candidate slot), and an extension whose
# probability is exactly zero is not created.
# ---------------------------------------------------------------------------------------------------------------------
def _logadd(a, b):
    if a == -np.inf:
        return b
    if b == -np.inf:
        return a
    m = max(a, b)
    return m + np.log(np.exp(a - m) + np.exp(b - m))


def beam_search(logp, out_len, beam_width=12, cutoff_top_n=40):
    """logp (B,T,V) log-probs, out_len (B,) -> list of int32 label arrays (best prefix per utterance, unfolded)."""
    logp = np.asarray(logp, dtype=np.float64)
    B, T, V = logp.shape
    top_n = cutoff_top_n if cutoff_top_n and cutoff_top_n > 0 else V
    res = []
    for b in range(B):
        # beam entry: (labels tuple, p_b, p_nb, score)
        beam = [((), 0.0, -np.inf, 0.0)]
        for t in range(int(out_len[b])):
            lp = logp[b, t]
            rank = np.array([sum(1 for c in range(V) if lp[c] > lp[v] or (lp[c] == lp[v] and c < v)) for v in range(V)])
            keep = rank < top_n
            n = len(beam)
            slots = {}                                   # slot index -> [labels, p_b, p_nb, parent, char]
            for i, (lab, pb, pnb, sc) in enumerate(beam):
                last = lab[-1] if lab else -1
                npb = sc + lp[0] if keep[0] else -np.inf
                npnb = lp[last] + pnb if (last > 0 and keep[last]) else -np.inf
                slots[i] = [lab, npb, npnb, i, -1]
            index = {e[0]: i for i, e in enumerate(beam)}
            for i, (lab, pb, pnb, sc) in enumerate(beam):
                last = lab[-1] if lab else -1
                for c in range(1, V):
                    if not keep[c]:
                        continue
                    if c == last:
                        l = -np.inf if pb == -np.inf else lp[c] + pb
                    else:
                        l = lp[c] + sc
                    if l == -np.inf:
                        continue
                    new = lab + (c,)
                    j = index.get(new)
                    if j is not None:                    # prefix already in the beam: merge
                        slots[j][2] = _logadd(slots[j][2], l)
                    else:
                        slots[n + i * V + c] = [new, -np.inf, l, i, c]
            cand = []
            for s, (lab, pb, pnb, par, ch) in slots.items():
                sc = _logadd(pb, pnb)
                if sc == -np.inf:
                    continue
                lastc = ch if ch >= 0 else (beam[par][0][-1] if beam[par][0] else -1)
                cand.append((-sc, lastc, s, lab, pb, pnb, sc))
            cand.sort(key=lambda x: (x[0], x[1], x[2]))
            beam = [(x[3], x[4], x[5], x[6]) for x in cand[:beam_width]]
        res.append(np.array(beam[0][0] if beam else (), dtype=np.int32))
    return res


def per_from_hyps(hyps, targets, targets_len, fold=True):
    """PER of given (unfolded) hypotheses; same conventions as per_batch."""
    dists, rlens, fh = [], [], []
    for b, h in enumerate(hyps):
        ref = np.asarray(targets[b, :int(targets_len[b])], dtype=np.int32)
        h = np.asarray(h, dtype=np.int32)
        if fold:
            h = FOLD_LUT[h]
            ref = FOLD_LUT[ref]
        h = h[h != 0]
        ref = ref[ref != 0]
        dists.append(levenshtein(h, ref))
        rlens.append(int(targets_len[b]))
        fh.append(h)
    acc = 0.0
    for d, r in zip(dists, rlens):
        acc += float(d) / float(r)
    return acc / len(dists), np.array(dists, np.int32), np.array(rlens, np.int32), fh
