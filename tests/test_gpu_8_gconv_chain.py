"""nbasr_gconv_chain (several grouped-conv edges of a cell in one launch) == the same edges launched one by one.

The chain kernel runs the same tile engine as nbasr_gconv_fwd (itself pinned to F.conv1d(groups=100) + autograd in
test_gpu_1_kernels.py): input-gradient chains are compared BIT-exactly, forward chains to 16-bit rounding (the fused
kernel evaluates ReLU20 as fma.sat on z / hi); what is under test is the cross-CTA tile dependency protocol
(flags / epoch in the work buffer), the node-boundary hand-over of the weight tile, skip-sum operands produced inside the
chain, repeated launches on one work buffer and CUDA-graph replay.  Reference: model.py:13-22,49-59."""
import ctypes as C

import pytest
import torch

import gpu_utils as U
from nb_asr_b200 import _lib
from nb_asr_b200._lib import BF16, F16, GConv
from nb_asr_b200.model import CONV_EDGES, pad_rule

pytestmark = pytest.mark.gpu


def _pack(lib, w, Cc, cpg, k, transposed, dt):
    ne = int(lib.nbasr_gconv_mma_pack_elems(Cc, cpg, k))
    out = torch.zeros(ne, dtype=U.tdt(dt), device=U.DEV)
    _lib.check(lib.nbasr_pack_gconv_mma(w.data_ptr(), out.data_ptr(), dt, Cc, cpg, k, transposed, U.stream()))
    return out


def _same(a, b, Cc, exact=True):
    """exact: bit equality.  Otherwise (forward chains: the fused kernel's ReLU20 runs as fma.sat on z / hi, see tile_fast in
    gconv_chain_sm100.cu) outputs agree to 16-bit rounding and gate bits differ on < 1e-3 of the elements.
    Of a gate-bit plane entry (8 bytes per row) only the slab's 6 (or 5) bytes are defined."""
    if a.dtype == torch.uint8:
        nb = (40 if Cc // 100 == 10 else 48) // 8
        x, y = a[..., :nb], b[..., :nb]
        if exact:
            return torch.equal(x, y)
        diff = (x ^ y).to(torch.int32)
        nbits = sum(((diff >> i) & 1).sum().item() for i in range(8))
        return nbits < 1e-3 * x.numel() * 8
    if exact:
        return torch.equal(a, b)
    return U.relerr(a.float(), b.float()) < 2e-3


def _build(lib, Cc, B, T, ops, skips, dt, backward, seed):
    """-> (list of GConv, list of output tensors, keep-alive list).  forward: relu + bias + skips + bf16 twin (out2) + gate
    bits; backward: input-gradient chain (x of node i+1 = out2 of node i, gated by a random mask, skip-sums of earlier outs)."""
    g = torch.Generator().manual_seed(seed)
    cpg = Cc // 100
    mw = 40 if cpg == 10 else 48
    keep, nodes, outs = [], [], []
    x0 = U.to_padded(torch.randn(B, T, Cc, generator=g), dt)
    src = x0
    produced = [x0]
    for i, op in enumerate(ops):
        k, d = CONV_EDGES[op]
        lp, _ = pad_rule(k, d, 1)
        w = (torch.randn(Cc, cpg, k, generator=g) * 0.3).to(U.DEV)
        wp = _pack(lib, w, Cc, cpg, k, 1 if backward else 0, dt)
        bias = (torch.randn(Cc, generator=g) * 0.1).to(U.DEV)
        o = U.empty_padded(B, T, Cc, dt)
        o2 = U.empty_padded(B, T, Cc, BF16 if not backward else dt)
        adds = [produced[j] for j in skips[i]]
        gc = GConv()
        gc.dtype, gc.x, gc.B, gc.T, gc.Tp, gc.C, gc.cpg = dt, src.data_ptr(), B, T, U.geo(T), Cc, cpg
        gc.ktaps, gc.dstep, gc.w, gc.w_packed = k, d, wp.data_ptr(), 1
        if not backward:
            gc.off0 = -lp
            mask = U.new_mask(o.shape[0], Cc, mw)
            gc.epi = U.epilogue(dt, Cc, bias=bias, relu=1, adds=adds, out=o, mask_out=mask, mask_w=mw, out2=o2, out2_dtype=BF16,
                                scale2=0.5)
            gc.epi.mask_rows = o.shape[0]
            keep += [mask]
            src = o
            outs += [o, o2, mask]
        else:
            gc.off0 = lp - (k - 1) * d
            m2 = torch.randint(0, 256, U.new_mask(o.shape[0], Cc, mw).shape, generator=g, dtype=torch.uint8).to(U.DEV)
            gc.epi = U.epilogue(dt, Cc, adds=adds, out=o, out2=o2, mask2=m2, mask2_w=mw, scale2=1.25)
            gc.epi.mask_rows = o.shape[0]
            keep += [m2]
            src = o2
            outs += [o, o2]
        produced.append(o)
        keep += [w, wp, bias, o, o2]
        nodes.append(gc)
    keep.append(x0)
    return nodes, outs, keep


CASES = [
    # Cc, B, T, ops, skips (indices into [x0, out_0, out_1, ...] added by each node), dtype, backward
    (800, 16, 500, ['conv5', 'conv5', 'conv5'], [[], [], []], F16, False),            # the default arch, ~4 tiles per CTA
    (600, 4, 500, ['conv7d2', 'conv5', 'conv5d2'], [[0], [0, 1], [0, 1, 2]], F16, False),
    (1000, 3, 257, ['conv7', 'conv7d2', 'conv5'], [[], [1], [0, 2]], F16, False),      # 40-channel slabs (cpg 10)
    (1200, 7, 130, ['conv5d2', 'conv7d2'], [[0], [1]], F16, False),                    # chain of two
    (800, 5, 333, ['conv5', 'conv5', 'conv5'], [[], [], []], BF16, True),
    (1000, 3, 257, ['conv7d2', 'conv5', 'conv7'], [[], [1], [1, 2]], BF16, True),
    (600, 64, 100, ['conv7d2', 'conv7d2', 'conv7d2'], [[0], [0, 1], [0, 1, 2]], BF16, True),   # one tile per utterance, many CTAs
    (1200, 2, 40, ['conv5', 'conv7'], [[], []], BF16, True),                           # fewer tiles than CTA slots: 1 tile per CTA
]


@pytest.mark.parametrize('case', range(len(CASES)))
def test_chain_equals_node_by_node(case):
    Cc, B, T, ops, skips, dt, backward = CASES[case]
    lib = _lib.load()
    n = len(ops)
    ref_nodes, ref_outs, k1 = _build(lib, Cc, B, T, ops, skips, dt, backward, seed=case)
    nodes, outs, k2 = _build(lib, Cc, B, T, ops, skips, dt, backward, seed=case)
    for gc in ref_nodes:
        _lib.check(lib.nbasr_gconv_fwd(C.byref(gc), U.stream()), 'gconv')
    wb = int(lib.nbasr_gconv_chain_work_bytes(B, T, Cc, Cc // 100, 3))
    work = torch.zeros(wb // 4, dtype=torch.int32, device=U.DEV)
    arr = (GConv * n)(*nodes)
    for rep in range(3):          # the work buffer is reused without clearing: the epoch advances by one per launch
        if rep:
            for o in outs:
                o.zero_()
        _lib.check(lib.nbasr_gconv_chain(arr, n, 1, work.data_ptr(), wb, U.stream()), 'gconv_chain')
        torch.cuda.synchronize()
        assert int(work[2]) == 0, 'a tile dependency timed out'
        assert int(work[0]) == rep + 1 and int(work[1]) == 0
        for a, b in zip(outs, ref_outs):
            assert _same(a, b, Cc, exact=backward), (case, rep)
    assert float(ref_outs[0].float().abs().sum()) > 0


def test_chain_graph_replay_and_shared_work_buffer():
    """Two chains of different geometry share one work buffer; captured in ONE CUDA graph and replayed."""
    lib = _lib.load()
    specs = [CASES[1], CASES[5]]
    built = [(_build(lib, *s[:4], s[4], s[5], s[6], seed=10 + i), _build(lib, *s[:4], s[4], s[5], s[6], seed=10 + i))
             for i, s in enumerate(specs)]
    wb = max(int(lib.nbasr_gconv_chain_work_bytes(s[1], s[2], s[0], s[0] // 100, 3)) for s in specs)
    work = torch.zeros(wb // 4, dtype=torch.int32, device=U.DEV)
    for (ref_nodes, _, _), _ in built:
        for gc in ref_nodes:
            _lib.check(lib.nbasr_gconv_fwd(C.byref(gc), U.stream()), 'gconv')
    arrs = [(GConv * len(b[1][0]))(*b[1][0]) for b in built]
    st = torch.cuda.Stream()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(st):
        for a in arrs:       # warm-up launch outside the capture (attribute opt-in, descriptor cache)
            _lib.check(lib.nbasr_gconv_chain(a, len(a), 1, work.data_ptr(), wb, st.cuda_stream))
        st.synchronize()
        with torch.cuda.graph(graph, stream=st):
            for a in arrs:
                _lib.check(lib.nbasr_gconv_chain(a, len(a), 1, work.data_ptr(), wb, st.cuda_stream))
    for rep in range(3):
        for b in built:
            for o in b[1][1]:
                o.zero_()
        graph.replay()
        torch.cuda.synchronize()
        assert int(work[2]) == 0
        for ((_, ref_outs, _), (_, outs, _)), bwd in zip(built, [sp[6] for sp in specs]):
            for a, b in zip(outs, ref_outs):
                assert _same(a, b, b.shape[-1] if b.dtype != torch.uint8 else (1000 if b.shape[0] == 25 else 600), exact=bwd), rep


@pytest.mark.parametrize('graph', [False, True])
def test_engine_with_fused_chains_matches_node_by_node(graph, monkeypatch):
    """NBASR_GCONV_CHAIN=1 (the engine passes fused = 1): one train step + one eval step of a conv architecture with skips, 16-bit
    mode, eager and CUDA-graph replay, against the default plan (one launch per node).  Forward agrees to 16-bit rounding, the
    gradient within the net's sensitivity to that; no tile dependency times out."""
    import nb_asr_b200 as nb
    arch = [[4, 1], [1, 0, 1], [2, 1, 0, 1]]          # conv7d2 / conv5 / conv5d2 edges, mixed skips: one chain of three per cell
    batch = nb.data.make_batch(3, 300, seed=3, min_len=150)
    out = {}
    for fused in ('0', '1'):
        monkeypatch.setenv('NBASR_GCONV_CHAIN', fused)
        nb.set_seed(1235)
        model = nb.get_model(arch, use_rnn=True, dropout_rate=0.0, gpu=0, precision='bf16')
        assert model.engine.fuse_chains == int(fused)
        tr = nb.get_trainer((nb.PhonemeEncoder(48), None, None, None), nb.get_loss(), gpus=[0], save_dir=None, verbose=False)
        tr.model = tr._model = model
        tr.optimizer = nb.trainer.FusedAdam(model, lr=1e-4)
        tr.use_graph = graph
        model.train()
        l0, lp0, _ = tr.step(batch, training=True)
        grad = model.engine.flat_g.clone()
        l1, _, _ = tr.step(batch, training=True)
        model.eval()
        le, lpe, _ = tr.step(batch, training=False)
        torch.cuda.synchronize()
        for pl in model.engine.plans.values():
            assert int(pl.chain_work[2]) == 0, 'a tile dependency timed out'
        if fused == '1':
            assert any(int(pl.chain_work[0]) > 0 for pl in model.engine.plans.values()), 'no fused chain was launched'
        out[fused] = (l0.item(), lp0.float().clone(), grad, l1.item(), le.item(), lpe.float().clone())
    a, b = out['1'], out['0']
    assert abs(a[0] - b[0]) < 1e-3 * abs(b[0]) and U.relerr(a[1], b[1]) < 2e-3
    # The fused forward differs from the unfused one by 16-bit rounding flips only (tools/dbg_fused_fwd.py: nothing above 1e-4 in
    # block 0, 7e-4 by the last block; fused input-gradient chains reproduce the gradient to 1.4e-7), but at initialisation this
    # net amplifies a 7e-4 forward perturbation to 2-4 % of the gradient (the LayerNorm bias at the input carries |g| = 500),
    # so this is a sanity bound; bit-level parity of the kernel is checked above.
    assert U.relerr(a[2], b[2]) < 8e-2
    # after two Adam updates (+-lr steps: chaotic in the sign of small gradients) the two runs have drifted a little
    assert abs(a[3] - b[3]) < 1e-2 * abs(b[3])
    assert abs(a[4] - b[4]) < 1e-2 * abs(b[4]) and U.relerr(a[5], b[5]) < 5e-2
