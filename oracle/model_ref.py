"""Torch-fp32 CPU restatement of the reference model, loss and train step.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Functional style over a plain
``state_dict`` with the reference's key names; no nn.Module forward is reused.

Reference locations restated here (all under /root/reference/nasbench_asr/):
  arch_vec -> op names ............ search_space.py:77-93
  padding rule / PadConvRelu ...... model/torch/ops.py:7-30
  Linear edge ..................... model/torch/ops.py:33-50
  Zero / Identity branches ........ model/torch/ops.py:53-83
  Node / SearchCell ............... model/torch/model.py:7-59
  ASRModel layout + forward ....... model/torch/model.py:62-131
  init (xavier / zeros) ........... model/torch/__init__.py:13-31
  loss ............................ training/torch/trainer.py:36-44
  step (reg, clip, Adam) .......... training/torch/trainer.py:208-227
"""
import math

import torch
import torch.nn.functional as F

ALL_OPS = ['linear', 'conv5', 'conv5d2', 'conv7', 'conv7d2', 'zero']   # search_space.py:6
FILTERS = [600, 800, 1000, 1200]          # model.py:74
STRIDES = [1, 1, 2, 2]                    # model.py:76
CELLS = [3, 4, 5, 6]                      # model.py:77
FEATURES = 80
HIDDEN = 500
NUM_CLASSES = 48                          # +1 blank -> 49 logits
GROUPS = 100
# (kernel, dilation) of the grouped conv edges, ops.py:73-76
CONV_EDGE = {'conv5': (5, 1), 'conv5d2': (5, 2), 'conv7': (7, 1), 'conv7d2': (7, 2)}


def arch_vec_to_names(arch_vec):
    return [[ALL_OPS[v[0]]] + list(v[1:]) for v in arch_vec]


def pad_rule(kernel, dilation, stride, context=4):
    """ops.py:12-17 -> (lpad, rpad)."""
    if int(context / stride) >= (kernel * dilation - stride):
        return 0, kernel * dilation - stride
    rpad = int(context / stride)
    return int((kernel - 1) * dilation - rpad), rpad


def module_order(arch_vec, use_rnn=True):
    """Parametrised leaf modules in construction order, as (prefix, kind, shape-args).

    Mirrors ASRModel.__init__ (model.py:79-106): top-level ModuleList index i
    counts conv, LN and every cell; node ops live at model.{i}.nodes.{n}.op.
    """
    names = arch_vec_to_names(arch_vec)
    out = []
    idx = 0
    for b in range(4):
        cin = FEATURES if b == 0 else FILTERS[b - 1]
        c = FILTERS[b]
        out.append((f'model.{idx}.conv', 'conv', (cin, c, 8, 1)))
        idx += 1
        out.append((f'model.{idx}', 'ln', (c,)))
        idx += 1
        for _ in range(CELLS[b]):
            for n, node in enumerate(names):
                op = node[0]
                if op == 'linear':
                    out.append((f'model.{idx}.nodes.{n}.op.linear', 'linear', (c, c)))
                elif op in CONV_EDGE:
                    k, _ = CONV_EDGE[op]
                    out.append((f'model.{idx}.nodes.{n}.op.conv', 'conv', (c, c, k, GROUPS)))
                elif op != 'zero':
                    raise ValueError(op)
            out.append((f'model.{idx}.norm_layer', 'ln', (c,)))
            idx += 1
    if use_rnn:
        idx += 1  # nn.Dropout holds no parameters but takes an index (model.py:99)
        out.append((f'model.{idx}', 'lstm', (FILTERS[3], HIDDEN)))
        idx += 1
        out.append((f'model.{idx}', 'linear', (HIDDEN, NUM_CLASSES + 1)))
    else:
        out.append((f'model.{idx}', 'linear', (FILTERS[3], NUM_CLASSES + 1)))
    return out


def build_state_dict(arch_vec, seed=None, use_rnn=True):
    """Same-seed initial weights as the reference's get_model.

    The reference first runs the default torch.nn constructors (which consume
    the global RNG) and then re-initialises in module order
    (model/torch/__init__.py:13-31).  We consume the RNG identically by
    instantiating the same leaf constructors in the same order.
    """
    if seed is not None:
        torch.manual_seed(seed)
    mods = []
    for prefix, kind, a in module_order(arch_vec, use_rnn):
        if kind == 'conv':
            m = torch.nn.Conv1d(a[0], a[1], a[2], groups=a[3])
        elif kind == 'linear':
            m = torch.nn.Linear(a[0], a[1])
        elif kind == 'ln':
            m = torch.nn.LayerNorm(a[0], eps=0.001)
        else:
            m = torch.nn.LSTM(input_size=a[0], hidden_size=a[1], batch_first=True)
        mods.append((prefix, kind, m))
    sd = {}
    for prefix, kind, m in mods:
        if kind in ('conv', 'linear'):
            torch.nn.init.xavier_uniform_(m.weight)
            torch.nn.init.zeros_(m.bias)
        elif kind == 'lstm':
            torch.nn.init.xavier_uniform_(m.weight_ih_l0)
            torch.nn.init.xavier_uniform_(m.weight_hh_l0)
            torch.nn.init.zeros_(m.bias_ih_l0)
            torch.nn.init.zeros_(m.bias_hh_l0)
        for k, v in m.state_dict().items():
            sd[f'{prefix}.{k}'] = v.detach().clone()
    return sd


def relu20(x):
    # ReLU then clamp_max 20 (ops.py:27-28); out-of-place so autograd works.
    return torch.clamp(torch.relu(x), max=20.0)


def pad_conv_relu(x, w, b, kernel, dilation, stride, groups):
    lp, rp = pad_rule(kernel, dilation, stride)
    x = F.pad(x, (lp, rp))
    return relu20(F.conv1d(x, w, b, stride=stride, dilation=dilation, groups=groups))


def layer_norm_ch(x, w, b):
    """LayerNorm over channels of a (B,C,T) tensor, eps 1e-3 (model.py:92,125-128)."""
    return F.layer_norm(x.permute(0, 2, 1), (x.shape[1],), w, b, 1e-3).permute(0, 2, 1)


def lstm_ref(x, w_ih, w_hh, b_ih, b_hh):
    """Single-layer LSTM, zero initial state, gate order i,f,g,o. x: (B,T,I) -> (B,T,H)."""
    B, T, _ = x.shape
    H = w_hh.shape[1]
    h = x.new_zeros(B, H)
    c = x.new_zeros(B, H)
    gx = x @ w_ih.t() + b_ih + b_hh
    outs = []
    for t in range(T):
        g = gx[:, t] + h @ w_hh.t()
        i, f, gg, o = g.split(H, dim=1)
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
        h = torch.sigmoid(o) * torch.tanh(c)
        outs.append(h)
    return torch.stack(outs, dim=1)


def forward(sd, arch_vec, audio, use_rnn=True, use_norm=True, collect=None):
    """(B,80,T) fp32 -> logits (B,T',49).  Dropout is identity (p=0 / eval)."""
    names = arch_vec_to_names(arch_vec)
    x = audio
    idx = 0
    for b in range(4):
        x = pad_conv_relu(x, sd[f'model.{idx}.conv.weight'], sd[f'model.{idx}.conv.bias'], 8, 1, STRIDES[b], 1)
        idx += 1
        x = layer_norm_ch(x, sd[f'model.{idx}.weight'], sd[f'model.{idx}.bias'])
        idx += 1
        for _ in range(CELLS[b]):
            outs = [x]
            for n, node in enumerate(names):
                op, branches = node[0], node[1:]
                assert len(branches) == len(outs)
                src = outs[-1]
                p = f'model.{idx}.nodes.{n}.op'
                if op == 'linear':
                    y = relu20(F.linear(src.permute(0, 2, 1), sd[p + '.linear.weight'], sd[p + '.linear.bias'])).permute(0, 2, 1)
                elif op in CONV_EDGE:
                    k, d = CONV_EDGE[op]
                    y = pad_conv_relu(src, sd[p + '.conv.weight'], sd[p + '.conv.bias'], k, d, 1, GROUPS)
                else:
                    y = torch.zeros_like(src)
                # python sum(): 0 + y + branch_0 + branch_1 ... left to right (model.py:16-22)
                acc = y
                for i, bit in enumerate(branches):
                    acc = acc + (outs[i] if bit else torch.zeros_like(outs[i]))
                outs.append(acc)
            x = outs[-1]
            if use_norm:
                x = layer_norm_ch(x, sd[f'model.{idx}.norm_layer.weight'], sd[f'model.{idx}.norm_layer.bias'])
            idx += 1
            if collect is not None:
                collect.append(x)
    if use_rnn:
        idx += 1
        p = f'model.{idx}'
        h = lstm_ref(x.permute(0, 2, 1), sd[p + '.weight_ih_l0'], sd[p + '.weight_hh_l0'],
                     sd[p + '.bias_ih_l0'], sd[p + '.bias_hh_l0'])
        idx += 1
    else:
        h = x.permute(0, 2, 1)
    return F.linear(h, sd[f'model.{idx}.weight'], sd[f'model.{idx}.bias'])


def ctc_loss_ref(logp, output_len, targets, targets_len):
    """trainer.py:36-44: per-utterance CTC NLL / output_len, batch mean, zero_infinity."""
    loss = F.ctc_loss(logp.permute(1, 0, 2), targets, output_len, targets_len,
                      reduction='none', zero_infinity=True)
    return (loss / output_len).mean()


def conv_reg_keys(sd):
    """Weights of every PadConvRelu (trainer.py:221): 3-D '.conv.weight' tensors."""
    return [k for k in sd if k.endswith('.conv.weight')]


def eval_step(sd, arch_vec, audio, audio_len, targets, targets_len, use_rnn=True):
    logits = forward(sd, arch_vec, audio, use_rnn)
    logp = F.log_softmax(logits, dim=2)
    out_len = audio_len // 4
    loss = ctc_loss_ref(logp, out_len, targets, targets_len)
    return loss, logp, out_len, logits


def train_step(sd, arch_vec, audio, audio_len, targets, targets_len, opt_state, lr=1e-4, use_rnn=True):
    """One reference training step; returns (loss, logp, grads(after clip), new sd, new opt_state).

    opt_state: dict(step=int, m={k:t}, v={k:t}) or None.  Adam(lr, eps=1e-7, betas .9/.999),
    regulariser 0.01*sum ||W_conv||_F, clip_grad_norm_ 5 (trainer.py:84,221-225).
    """
    p = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()}
    logits = forward(p, arch_vec, audio, use_rnn)
    logp = F.log_softmax(logits, dim=2)
    out_len = audio_len // 4
    loss = ctc_loss_ref(logp, out_len, targets, targets_len)
    reg = loss + 0.01 * sum(torch.norm(p[k]) for k in conv_reg_keys(p))
    reg.backward()
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in p.items()}
    raw = {k: g.clone() for k, g in grads.items()}
    total = math.sqrt(sum(float((g.double() ** 2).sum()) for g in grads.values()))
    coef = min(1.0, 5.0 / (total + 1e-6))
    if opt_state is None:
        opt_state = dict(step=0, m={k: torch.zeros_like(v) for k, v in sd.items()},
                         v={k: torch.zeros_like(v) for k, v in sd.items()})
    t = opt_state['step'] + 1
    b1, b2, eps = 0.9, 0.999, 1e-7
    new_sd, m_new, v_new = {}, {}, {}
    for k in sd:
        g = grads[k] * coef
        m = opt_state['m'][k] * b1 + (1 - b1) * g
        v = opt_state['v'][k] * b2 + (1 - b2) * g * g
        denom = v.sqrt() / math.sqrt(1 - b2 ** t) + eps
        new_sd[k] = sd[k] - (lr / (1 - b1 ** t)) * m / denom
        m_new[k], v_new[k] = m, v
    return (loss.detach(), logp.detach(), raw, total, new_sd,
            dict(step=t, m=m_new, v=v_new), logits.detach())


def make_batch(B, T, seed=0, min_len=None, tgt_lo=10, tgt_hi=30):
    """Synthetic fixture of SURVEY.md §8c / §8d (N(0,1) log-mel stand-in, U{1..48} labels)."""
    g = torch.Generator().manual_seed(seed)
    audio = torch.randn(B, FEATURES, T, generator=g)
    lo = T // 2 if min_len is None else min_len
    alen = torch.randint(lo, T + 1, (B,), generator=g)
    alen[0] = T
    for b in range(B):
        audio[b, :, int(alen[b]):] = 0.0
    tl = torch.randint(tgt_lo, tgt_hi, (B,), generator=g)
    targets = torch.randint(1, 49, (B, int(tl.max())), generator=g, dtype=torch.int32)
    for b in range(B):
        targets[b, int(tl[b]):] = 0
    return audio, alen, targets, tl
