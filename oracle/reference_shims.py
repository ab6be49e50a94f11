"""Import the REAL reference package (unmodified) behind the four shims of SURVEY.md §8c.

TEST / BASELINE INFRASTRUCTURE (see oracle/__init__.py).  Used by oracle/make_golden*.py (reference at
/root/reference, build container only) and by bench.py's `--impl reference` / `cpu_baseline` legs (reference installed
by `oracle/install_reference.sh` into baseline/_ref/, which is git-ignored but travels to the GPU box).

Shims (the reference's own files stay byte-identical):
  1. torchaudio.set_audio_backend no-op      (training/torch/timit.py:11 calls an API torchaudio 2.x removed)
  2. stub module torch_edit_distance         (training/torch/trainer.py:8; third-party, not installed)
  3. stub module ctcdecode.CTCBeamDecoder    (training/torch/trainer.py:9,71; third-party, not installed)
  4. torch.clamp_max_ -> out-of-place        (model/torch/ops.py:28,47: the in-place clamp on ReLU's output makes
                                              autograd raise on torch >= 1.8; forward values are identical)
"""
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INSTALLED = os.path.join(ROOT, 'baseline', '_ref')


def find_reference():
    """Path that contains the `nasbench_asr` package: $NBASR_REFERENCE, baseline/_ref (pip --target), /root/reference."""
    for p in (os.environ.get('NBASR_REFERENCE'), INSTALLED, '/root/reference'):
        if p and os.path.isdir(os.path.join(p, 'nasbench_asr')):
            return p
    return None


class StubEncoder:
    """Stands in for PhonemeEncoder(48) where the reference's data file (training/timit_folding.txt, not packaged by the
    reference's setup.py) is absent: Trainer.__init__ only asks it for the 49-symbol vocabulary (trainer.py:71)."""

    def get_vocab(self, inc_blank=True):
        return ['_'] + [f'p{i}' for i in range(48)]


def import_reference(path=None):
    import torch
    path = path or find_reference()
    if path is None:
        raise ImportError('reference package not found (run oracle/install_reference.sh in the build container)')
    try:
        import torchaudio
        torchaudio.set_audio_backend = lambda *a, **k: None
    except ImportError:                       # timit.py imports torchaudio at module level
        ta = types.ModuleType('torchaudio')
        ta.set_audio_backend = lambda *a, **k: None
        ta.transforms = types.ModuleType('torchaudio.transforms')
        sys.modules['torchaudio'] = ta
        sys.modules['torchaudio.transforms'] = ta.transforms
    sys.modules.setdefault('torch_edit_distance', types.ModuleType('torch_edit_distance'))
    cd = types.ModuleType('ctcdecode')

    class CTCBeamDecoder:
        def __init__(self, *a, **k):
            pass
    cd.CTCBeamDecoder = CTCBeamDecoder
    sys.modules.setdefault('ctcdecode', cd)
    torch.clamp_max_ = lambda x, m: torch.clamp_max(x, m)
    if path not in sys.path:
        sys.path.insert(0, path)
    import nasbench_asr
    nasbench_asr.set_default_backend('torch')
    return nasbench_asr


def reference_trainer(nb, model, lr=1e-4, encoder=None):
    """The reference's own Trainer around `model`, on the CPU (gpus=[]), with the optimiser Trainer.train would create
    (trainer.py:84)."""
    import torch
    tr = nb.get_trainer((encoder or StubEncoder(), None, None, None), nb.get_loss(), gpus=[], save_dir=None, verbose=False)
    tr.model = tr._model = model
    tr.optimizer = torch.optim.Adam(model.parameters(), lr=lr, eps=1e-7)
    return tr
