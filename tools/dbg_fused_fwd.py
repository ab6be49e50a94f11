"""Which forward buffers differ between the fused-chain plan (NBASR_GCONV_CHAIN=2: forward chains fused) and the default plan?
Compares the two plans' arenas (same allocation order) after ONE training-mode forward pass, per tensor written by a chain."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import nb_asr_b200 as nb
arch = [[4, 1], [1, 0, 1], [2, 1, 0, 1]]
B, T = 3, 300
batch = nb.data.make_batch(B, T, seed=3, min_len=150)
plans = {}
for fused in ('0', '2'):
    os.environ['NBASR_GCONV_CHAIN'] = fused
    nb.set_seed(1235)
    model = nb.get_model(arch, use_rnn=True, dropout_rate=0.0, gpu=0, precision='bf16')
    model.train()
    pl = model.engine.forward(batch[0][0].cuda(), training=True, grad=True)
    torch.cuda.synchronize()
    plans[fused] = (model, pl)


def locate(pl, ptr):
    for bi, blk in enumerate(pl.arena.blocks):
        base = blk.data_ptr()
        if base <= ptr < base + blk.numel():
            return bi, ptr - base
    return None


(m0, p0), (m1, p1) = plans['0'], plans['2']
lib = m0.engine.lib
nchain = 0
for (fn0, a0), (fn1, a1) in zip(p0.fwd, p1.fwd):
    if fn0.__name__ != 'nbasr_gconv_chain':
        continue
    for i in range(a0[1]):
        g0, g1 = a0[0][i], a1[0][i]
        rows = g0.B * g0.Tp
        for what, q0, q1, nbytes in (('out', g0.epi.out, g1.epi.out, rows * g0.C * 2), ('out2', g0.epi.out2, g1.epi.out2, rows * g0.C * 2),
                                     ('mask', g0.epi.mask_out, g1.epi.mask_out, -(-g0.C // g0.epi.mask_w) * g0.epi.mask_rows * 8 if g0.epi.mask_out else 0)):
            if not q0:
                continue
            (b0, o0), (b1, o1) = locate(p0, q0), locate(p1, q1)
            t0 = p0.arena.blocks[b0][o0:o0 + nbytes]
            t1 = p1.arena.blocks[b1][o1:o1 + nbytes]
            if what == 'mask':
                x0 = t0.view(-1, 8)[:, :(40 if g0.cpg == 10 else 48) // 8]
                x1 = t1.view(-1, 8)[:, :(40 if g0.cpg == 10 else 48) // 8]
                nd = int((x0 != x1).sum())
                if nd:
                    bad = (x0 != x1).any(1).nonzero().flatten()
                    print(f'chain {nchain} node {i} C={g0.C} MASK: {nd} bytes differ; entries {bad[:6].tolist()} .. rows-in-plane {(bad[:6] % g0.epi.mask_rows).tolist()}')
            else:
                dt = torch.float16 if (what == 'out') else torch.bfloat16
                x0, x1 = t0.view(dt).float().view(rows, g0.C), t1.view(dt).float().view(rows, g0.C)
                d = (x0 - x1).abs()
                rel = float(d.norm() / x0.norm().clamp_min(1e-30))
                if rel > 1e-4:
                    r, c = divmod(int(d.argmax()), g0.C)
                    print(f'chain {nchain} node {i} C={g0.C} {what}: relerr {rel:.3e}, worst at row {r} (utt {r // g0.Tp}, t {r % g0.Tp - 8}) col {c}: {float(x0[r, c])} vs {float(x1[r, c])}; rows with diffs: {int((d.sum(1) > 0).sum())}')
    nchain += 1
print('chains compared', nchain)
