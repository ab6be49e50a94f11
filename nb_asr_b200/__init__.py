"""nb_asr_b200 -- B200-native drop-in for the NAS-Bench-ASR candidate train/eval step.

Facade with the names of nasbench_asr/__init__.py:11-52 for the hot path:
    set_default_backend, get_backend_name, set_seed, prepare_devices, get_model, get_loss,
    get_trainer, get_dataloaders (synthetic stand-in; the TIMIT file readers are out of scope, the log-mel
    front end is nb_asr_b200.frontend).
Everything executes through libnbasr.so (hand-written sm_100a CUDA, include/nbasr.h).
"""
import random

import numpy
import torch

from . import data, graph_utils, search_space
from .encoder import PhonemeEncoder
from .model import ASRModel, PadConvRelu, get_model, print_model_summary
from .trainer import AvgMeter, Trainer, get_loss, get_trainer, set_time_limit

__version__ = '0.1.0'
BACKEND = 'b200'


def set_default_backend(backend=None):
    if backend not in (None, 'b200', 'torch'):
        raise ValueError(f'Unknown backend: {backend}')
    return BACKEND, BACKEND


def get_backend_name():
    return BACKEND, BACKEND


def set_seed(seed):
    """training/torch/__init__.py:9-14"""
    random.seed(seed)
    numpy.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed(seed)
        torch.cuda.manual_seed_all(seed)


def prepare_devices(devices):
    pass


def get_dataloaders(timit_root=None, batch_size=64, **synthetic):
    """(encoder, train, val, test). The TIMIT reader/featuriser is out of scope;
    this returns synthetic log-mel loaders of the same batch structure."""
    from .data import synthetic_dataloaders
    return synthetic_dataloaders(batch_size=batch_size, **synthetic)
