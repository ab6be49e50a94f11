"""Worker of tests/test_gpu_6_graph_dp.py::test_two_rank_nccl_step_equals_single_process_step (run under torchrun, 2 ranks).

Each rank first runs the SINGLE-process step on the concatenated batch (no process group yet), then the process group is
created and the same model takes the same steps data-parallel on its half of the utterances (CUDA graphs + bucketed NCCL
all-reduce, the path bench.py --gpus N times).  Rank 0 writes the comparison to <out>/result.json."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import nb_asr_b200 as nb  # noqa: E402


def make(arch, precision, graph, local):
    nb.set_seed(1235)
    model = nb.get_model(arch, use_rnn=True, dropout_rate=0.0, gpu=local, precision=precision)
    tr = nb.get_trainer((nb.PhonemeEncoder(48), None, None, None), nb.get_loss(), gpus=[local], save_dir=None, verbose=False)
    tr.model = tr._model = model
    tr.optimizer = nb.trainer.FusedAdam(model, lr=1e-4)
    tr.use_graph = graph
    model.train()
    return model, tr


def main():
    out, precision = sys.argv[1], sys.argv[2]
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    arch = [[4, 1], [0, 1, 1], [2, 0, 1, 1]]
    b, T = 3, 140
    (audio, alen), (tg, tl) = nb.data.make_batch(world * b, T, seed=0, min_len=T // 2)
    full = ((audio, alen), (tg, tl))
    sl = slice(rank * b, (rank + 1) * b)
    shard = ((audio[sl], alen[sl]), (tg[sl], tl[sl]))

    # ---- single process, whole batch (eager path, already pinned to the reference by test_gpu_5_model)
    model, tr = make(arch, precision, False, local)
    init_params = None
    ref_losses, ref_grad = [], None
    for i in range(2):
        if i == 0:
            model.engine.bind()
            init_params = model.engine.flat_p.clone()
        ref_losses.append(tr.step(full, training=True)[0].item())
        if i == 0:
            ref_grad = model.engine.flat_g.clone()      # averaged raw gradient + regulariser gradient of the first step
    ref_params = model.engine.flat_p.clone()
    ref_losses.append(tr.step(full, training=False)[0].item())
    del model, tr

    # ---- data parallel
    dist.init_process_group('nccl', device_id=dev)
    model, tr = make(arch, precision, True, local)
    losses, grad = [], None
    for i in range(2):
        l = tr.step(shard, training=True)[0].clone()
        if i == 0:
            grad = model.engine.flat_g.clone()
        dist.all_reduce(l, op=dist.ReduceOp.AVG)        # equal shards: mean of shard means = global mean (trainer.py:41)
        losses.append(l.item())
    params = model.engine.flat_p.clone()
    l = tr.step(shard, training=False)[0].clone()
    dist.all_reduce(l, op=dist.ReduceOp.AVG)
    losses.append(l.item())
    lo, hi = params.clone(), params.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    pl = model.engine.plan(b, T, True)
    # Adam's first steps are -lr * g / (|g| + eps): wherever the true gradient is ~0 (e.g. every bias in front of a LayerNorm)
    # the step is +-lr with the sign of fp32 summation noise, so the update is compared on the elements whose reference
    # gradient is significant (and the fraction of such elements is reported)
    # (16-bit mode: the exchanged gradient agrees to ~1e-4 of its norm, so 'significant' starts further above the noise)
    sig = ref_grad.abs() > (1e-3 if precision == 'fp32' else 5e-2) * ref_grad.abs().mean()
    upd, ref_upd = (params - init_params).double(), (ref_params - init_params).double()
    res = dict(ranks_equal=bool(torch.equal(lo, hi)),
               loss0_rel=abs(losses[0] - ref_losses[0]) / abs(ref_losses[0]),
               loss1_rel=abs(losses[1] - ref_losses[1]) / abs(ref_losses[1]),
               loss2_rel=abs(losses[2] - ref_losses[2]) / abs(ref_losses[2]),
               params_rel=float((params.double() - ref_params.double()).norm() / ref_params.double().norm()),
               update_rel=float((params.double() - ref_params.double()).norm() / (ref_params.double() - init_params.double()).norm()),
               update_rel_sig=float((upd[sig] - ref_upd[sig]).norm() / ref_upd[sig].norm()), sig_frac=float(sig.double().mean()),
               grad_rel=float((grad.double() - ref_grad.double()).norm() / ref_grad.double().norm()),
               used_graph=bool(tr.use_graph), buckets=len(pl.buckets), losses=losses, ref_losses=ref_losses)
    if rank == 0:
        json.dump(res, open(os.path.join(out, 'result.json'), 'w'))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
