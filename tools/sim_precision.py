"""CPU simulation of where bf16 rounding is applied (tools only; uses the oracle as the fp32 truth).

Policies (flags):
  w      weights rounded to bf16 (MMA operand)
  opnd   the input of every op (MMA A operand) rounded to bf16
  node   node outputs (after the skip sum) STORED in bf16 (=> skips and LN input see rounded values)
  ln     LayerNorm outputs stored in bf16
  z      time-reduction conv outputs stored in bf16 before their LN
  head   LSTM input/h operands + W_ih/W_hh in bf16
"""
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from oracle import model_ref as M

ARCHS = {'default': [[1, 0], [1, 0, 0], [1, 0, 0, 0]], 'c7d2_skips': [[4, 1], [4, 1, 1], [4, 1, 1, 1]],
         'linear_skips': [[0, 1], [0, 1, 1], [0, 1, 1, 1]], 'mixed': [[2, 1], [3, 0, 1], [0, 1, 0, 1]],
         'c5_c7': [[1, 0], [3, 1, 0], [1, 0, 0, 1]]}


def r(x):
    return x.to(torch.bfloat16).float()


def forward(sd, arch_vec, audio, pol):
    names = M.arch_vec_to_names(arch_vec)
    W = (lambda t: r(t)) if 'w' in pol else (lambda t: t)
    OP = (lambda t: r(t)) if 'opnd' in pol else (lambda t: t)
    ND = (lambda t: r(t)) if 'node' in pol else (lambda t: t)
    LN = (lambda t: r(t)) if 'ln' in pol else (lambda t: t)
    Z = (lambda t: r(t)) if 'z' in pol else (lambda t: t)
    x = audio
    idx = 0
    for b in range(4):
        x = Z(M.pad_conv_relu(OP(x), W(sd[f'model.{idx}.conv.weight']), sd[f'model.{idx}.conv.bias'], 8, 1, M.STRIDES[b], 1))
        idx += 1
        x = LN(M.layer_norm_ch(x, sd[f'model.{idx}.weight'], sd[f'model.{idx}.bias']))
        idx += 1
        for _ in range(M.CELLS[b]):
            outs = [x]
            for n, node in enumerate(names):
                op, branches = node[0], node[1:]
                src = OP(outs[-1])
                p = f'model.{idx}.nodes.{n}.op'
                if op == 'linear':
                    y = M.relu20(F.linear(src.permute(0, 2, 1), W(sd[p + '.linear.weight']), sd[p + '.linear.bias'])).permute(0, 2, 1)
                elif op in M.CONV_EDGE:
                    k, d = M.CONV_EDGE[op]
                    y = M.pad_conv_relu(src, W(sd[p + '.conv.weight']), sd[p + '.conv.bias'], k, d, 1, M.GROUPS)
                else:
                    y = torch.zeros_like(src)
                acc = y
                for i, bit in enumerate(branches):
                    if bit:
                        acc = acc + outs[i]
                outs.append(ND(acc))
            x = LN(M.layer_norm_ch(outs[-1], sd[f'model.{idx}.norm_layer.weight'], sd[f'model.{idx}.norm_layer.bias']))
            idx += 1
    idx += 1
    p = f'model.{idx}'
    if 'head' in pol:
        B, T = x.shape[0], x.shape[2]
        H = M.HIDDEN
        xx = r(x.permute(0, 2, 1))
        gx = xx @ r(sd[p + '.weight_ih_l0']).t() + sd[p + '.bias_ih_l0'] + sd[p + '.bias_hh_l0']
        h = xx.new_zeros(B, H); c = xx.new_zeros(B, H)
        whh = r(sd[p + '.weight_hh_l0'])
        outs = []
        for t in range(T):
            g = gx[:, t] + r(h) @ whh.t()
            i, f, gg, o = g.split(H, dim=1)
            c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
            h = torch.sigmoid(o) * torch.tanh(c)
            outs.append(h)
        h = r(torch.stack(outs, 1))
    else:
        h = M.lstm_ref(x.permute(0, 2, 1), sd[p + '.weight_ih_l0'], sd[p + '.weight_hh_l0'], sd[p + '.bias_ih_l0'], sd[p + '.bias_hh_l0'])
    idx += 1
    return F.linear(h, sd[f'model.{idx}.weight'], sd[f'model.{idx}.bias'])


if __name__ == '__main__':
    B, T = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (3, 70)
    pols = [('current', {'w', 'opnd', 'node', 'ln', 'z', 'head'}),
            ('w only', {'w'}),
            ('w+head', {'w', 'head'}),
            ('opnd only', {'opnd'}),
            ('w+opnd+head (fp32 storage)', {'w', 'opnd', 'head'}),
            ('w+opnd+ln+z+head (fp32 node)', {'w', 'opnd', 'ln', 'z', 'head'}),
            ('w+opnd+node+head (fp32 ln/z)', {'w', 'opnd', 'node', 'head'}),
            ]
    torch.set_num_threads(8)
    for name, arch in ARCHS.items():
        sd = M.build_state_dict(arch, seed=1235)
        audio, alen, tg, tl = M.make_batch(B, T, seed=0, min_len=T // 2)
        with torch.no_grad():
            ref = forward(sd, arch, audio, set())
            for pn, pol in pols:
                out = forward(sd, arch, audio, pol)
                rel = float((out.double() - ref.double()).norm() / ref.double().norm())
                print(f'{name:14s} {pn:34s} {rel:.3e}', flush=True)
