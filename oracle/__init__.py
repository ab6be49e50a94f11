"""CPU oracle for the NAS-Bench-ASR candidate train/eval step.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and there only as the checker.

Pinning status (see DESIGN.md "Oracle"):
  * model forward / CTC loss / gradients / Adam step: PINNED against outputs of
    the real reference (``/root/reference`` imported with four shims by
    ``oracle/make_golden.py``; fixtures committed under ``tests/golden/``).
  * greedy CTC decode + 48->39 fold + Levenshtein PER: the reference delegates
    decode to ``ctcdecode`` (git 9a20e00f34d8f605f4a8501cc42b1a53231f1597,
    setup.py:49) and scoring to ``torch-edit-distance`` (unpinned, setup.py:50);
    neither source is in /root/reference nor installed, and the reference holds
    no tests -> **parity unpinned** at that boundary.  The restatement follows
    training/tf/metrics/ctc.py:76-81 (greedy), training/torch/encoder.py:64-74
    (fold, effective chained LUT) and trainer.py:245-246 (PER = edit distance /
    reference length, batch mean).  The fold LUT itself IS pinned (generated
    from the reference's PhonemeEncoder by make_golden.py).
"""
