"""arch_vec -> torch module whose forward/backward run on the B200 engine (nb_asr_b200.engine).

Drop-in for nasbench_asr.model.torch (model/torch/__init__.py:7-35 get_model, model.py:62-135
ASRModel): same constructor order (so same-seed initial weights are bit-identical), same
state_dict keys/shapes, same attributes (arch_desc, num_classes, use_rnn, use_norm,
dropout_rate, backend, get_prunable_copy). The sub-modules are PARAMETER CONTAINERS: all compute
goes through libnbasr's CUDA kernels; there is no torch/CPU execution path.
"""
import torch
import torch.nn as nn

from . import search_space as ss

FEATURES = 80
FILTERS = [600, 800, 1000, 1200]
TR_KERNEL = 8
TR_STRIDES = [1, 1, 2, 2]
CELLS_PER_BLOCK = [3, 4, 5, 6]
HIDDEN = 500
GROUPS = 100
# name -> (kernel, dilation) of the grouped-conv edges (ops.py:73-76)
CONV_EDGES = {'conv5': (5, 1), 'conv5d2': (5, 2), 'conv7': (7, 1), 'conv7d2': (7, 2)}


def pad_rule(kernel, dilation, stride, context=4):
    """(left, right) zero padding of PadConvRelu (ops.py:12-17)."""
    if int(context / stride) >= kernel * dilation - stride:
        return 0, kernel * dilation - stride
    right = int(context / stride)
    return int((kernel - 1) * dilation - right), right


class _Container(nn.Module):
    def forward(self, *a, **k):
        raise RuntimeError(f'{type(self).__name__} is a parameter container of the B200 engine; '
                           'call the ASRModel, not its sub-modules')


class PadConvRelu(_Container):
    """Holds `conv` (nn.Conv1d) like ops.py:7-30; the trainer's regulariser finds it by type."""

    def __init__(self, in_channels, out_channels, kernel_size, dilation, strides, groups=1, dropout_rate=0, name='PadConvRelu'):
        super().__init__()
        self.name = name
        self.kernel_size, self.dilation, self.strides, self.groups = kernel_size, dilation, strides, groups
        self.dropout_rate = dropout_rate
        self.lpad, self.rpad = pad_rule(kernel_size, dilation, strides)
        self.conv = nn.Conv1d(in_channels, out_channels, kernel_size, stride=strides, dilation=dilation, groups=groups)


class Linear(_Container):
    def __init__(self, in_features, out_features, dropout_rate=0, name='Linear'):
        super().__init__()
        self.name = name
        self.dropout_rate = dropout_rate
        self.linear = nn.Linear(in_features, out_features)


class Zero(_Container):
    def __init__(self, name='zero'):
        super().__init__()
        self.name = name


class Identity(_Container):
    def __init__(self, name='Identity'):
        super().__init__()
        self.name = name


def _make_op(name, filters, dropout_rate):
    if name == 'linear':
        return Linear(filters, filters, dropout_rate=dropout_rate)
    if name in CONV_EDGES:
        k, d = CONV_EDGES[name]
        return PadConvRelu(filters, filters, k, d, 1, groups=GROUPS, dropout_rate=dropout_rate, name=name)
    if name == 'zero':
        return Zero()
    raise ValueError(f'Operation "{name}" is not implemented')


class Node(_Container):
    def __init__(self, filters, op_name, branches, dropout_rate=0.0):
        super().__init__()
        for b in branches:
            if b not in (0, 1):
                raise ValueError(f'Invalid branch operations: {branches}, expected is a vector of 0 (no skip-con.) '
                                 'and 1 (skip-con. present)')
        self.op_name = op_name
        self.branches = list(branches)
        self.op = _make_op(op_name, filters, dropout_rate)
        # plain list, as in the reference (model.py:11): branch ops own no parameters
        self.branch_ops = [Identity() if b else Zero() for b in branches]


class SearchCell(_Container):
    def __init__(self, filters, node_configs, dropout_rate=0.0, use_norm=True):
        super().__init__()
        self.nodes = nn.ModuleList()
        for cfg in node_configs:
            name, *branches = cfg
            self.nodes.append(Node(filters, name, branches, dropout_rate))
        self.use_norm = use_norm
        if use_norm:
            self.norm_layer = nn.LayerNorm(filters, eps=0.001)


class ASRModel(nn.Module):
    def __init__(self, arch_desc, num_classes=48, use_rnn=False, use_norm=True, dropout_rate=0.0, **kwargs):
        super().__init__()
        self.arch_desc = arch_desc
        self.num_classes = num_classes
        self.use_rnn = use_rnn
        self.use_norm = use_norm
        self.dropout_rate = dropout_rate
        layers = nn.ModuleList()
        for i in range(4):
            cin = FEATURES if i == 0 else FILTERS[i - 1]
            layers.append(PadConvRelu(cin, FILTERS[i], TR_KERNEL, 1, TR_STRIDES[i], groups=1, name=f'conv_{i}'))
            layers.append(nn.LayerNorm(FILTERS[i], eps=0.001))
            for _ in range(CELLS_PER_BLOCK[i]):
                layers.append(SearchCell(FILTERS[i], arch_desc, dropout_rate=dropout_rate, use_norm=use_norm))
        if use_rnn:
            layers.append(nn.Dropout(dropout_rate))
            layers.append(nn.LSTM(input_size=FILTERS[-1], hidden_size=HIDDEN, batch_first=True, dropout=0.0))
            layers.append(nn.Linear(HIDDEN, num_classes + 1))
        else:
            layers.append(nn.Linear(FILTERS[-1], num_classes + 1))
        self.model = layers
        self._engine = None
        self._device_init = None
        self.precision = kwargs.get('precision', 'bf16')

    # -- drop-in surface -------------------------------------------------------------------
    @property
    def backend(self):
        return 'b200'

    def get_prunable_copy(self, bn=False, masks=None):
        new = ASRModel(self.arch_desc, num_classes=self.num_classes, use_rnn=self.use_rnn, use_norm=bn,
                       dropout_rate=self.dropout_rate, precision=self.precision)
        new.load_state_dict(self.state_dict(), strict=False)
        dev = next(self.parameters()).device
        new.to(dev)
        new.train()
        return new

    # -- engine ----------------------------------------------------------------------------
    @property
    def engine(self):
        if self._engine is None:
            from .engine import Engine
            self._engine = Engine(self, precision=self.precision)
        return self._engine

    def set_precision(self, precision):
        assert precision in ('bf16', 'fp32')
        if precision != self.precision:
            self.precision = precision
            self._engine = None
        return self

    def forward(self, input):
        """(B, 80, T) fp32 on a CUDA device -> logits (B, ceil(ceil(T/2)/2), 49) fp32."""
        from .engine import ModelFunction
        if not input.is_cuda:
            raise RuntimeError('nb_asr_b200 runs on a B200 only: move the model and the input to cuda '
                               '(there is deliberately no CPU fallback)')
        return ModelFunction.apply(self, input, *self.engine.params)


def get_model(arch_vec, use_rnn, dropout_rate, gpu=None, precision='bf16', init='reference', seed=None, conv_gain=1.0):
    """model/torch/__init__.py:7-35: build, re-initialise (xavier / zeros), move to cuda:{gpu}.

    init='reference' (default): the reference's own procedure on the host RNG -- same seed, bit-identical weights.
    init='device': for architecture sweeps (thousands of candidates; the host procedure costs 0.3-1.1 s each).  The
    module tree is built without storage, materialised on cuda:{gpu} and initialised there by the engine from a device
    generator seeded with `seed`: the same DISTRIBUTIONS as model/torch/__init__.py:13-29 (xavier_uniform weights, zero
    biases, LayerNorm 1 / 0), not the same stream.  `conv_gain` multiplies the xavier bound of the grouped-conv edges
    (sqrt(1 + groups) makes them variance-preserving: SURVEY finding 5, activations vanish through skip-free conv
    cells at the reference's init)."""
    ss.validate_arch(arch_vec)
    arch_desc = ss.arch_vec_to_names(arch_vec)
    if init == 'device':
        if gpu is None:
            raise ValueError("init='device' needs gpu=<index>")
        with torch.device('meta'):
            model = ASRModel(arch_desc, use_rnn=use_rnn, dropout_rate=dropout_rate, precision=precision)
        # no storage yet: the engine gives every parameter a view of its flat device buffer and fills it there
        model._device_init = dict(seed=1235 if seed is None else int(seed), conv_gain=float(conv_gain),
                                  device=torch.device('cuda', int(gpu)))
        return model
    if init != 'reference':
        raise ValueError(f'unknown init {init!r}')
    model = ASRModel(arch_desc, use_rnn=use_rnn, dropout_rate=dropout_rate, precision=precision)

    def init_weights(m):
        if isinstance(m, (nn.Linear, nn.Conv1d)):
            nn.init.xavier_uniform_(m.weight)
            nn.init.zeros_(m.bias)
        elif isinstance(m, nn.LSTM):
            for layer in range(m.num_layers):
                nn.init.xavier_uniform_(getattr(m, f'weight_ih_l{layer}'))
                nn.init.xavier_uniform_(getattr(m, f'weight_hh_l{layer}'))
                nn.init.zeros_(getattr(m, f'bias_ih_l{layer}'))
                nn.init.zeros_(getattr(m, f'bias_hh_l{layer}'))

    model.apply(init_weights)
    if gpu is not None:
        model.to(device=f'cuda:{gpu}')
    return model


def print_model_summary(model):
    print(model)
    print('======================')

    def walk(m, level=0):
        for n, child in m.named_children():
            print('  ' * level + type(child).__name__, ' ', n, ' ', sum(p.numel() for p in child.parameters()))
            walk(child, level + 1)
    walk(model.model)
    print('======================')
    n = sum(p.numel() for p in model.parameters())
    print('Trainable parameters:', f'{n:,}'.replace(',', ' '))
