"""Architecture-sharded sweep evaluation (BASELINE.json configs[3]): every candidate arch_vec is built from the
same seed, run through the eval step (forward + CTC loss + greedy decode + PER) on a fixed synthetic set and
reported as one row.  Candidates are independent: rank r evaluates its shard (nb_asr_b200.distributed.shard_archs),
no collective touches the data path, rows are gathered on rank 0.

    torchrun --nproc-per-node 8 -m nb_asr_b200.sweep --limit 64
The reference has no sweep driver; enumeration follows search_space.get_all_architectures and, by
default, keeps the first arch_vec of every graph-isomorphism class (graph_utils.get_model_hash: 13 824 -> 8 242).
``--out-pickle`` writes an nb-asr style table (README "Dataset format": pickle.dump(header) then pickle.dump(rows),
row order = header['columns']).
"""
import argparse
import json
import os
import time

import torch

import pickle

from . import PhonemeEncoder, data, distributed, get_loss, get_model, get_trainer, graph_utils, search_space, set_seed


def evaluate_arch(arch, batches, gpu, precision='bf16', seed=1235):
    set_seed(seed)
    model = get_model(arch, use_rnn=True, dropout_rate=0.0, gpu=gpu, precision=precision)
    model.eval()
    tr = get_trainer((PhonemeEncoder(48), None, None, None), get_loss(), gpus=[gpu], verbose=False)
    tr.model = tr._model = model
    losses, pers = [], []
    for batch in batches:
        loss, logp, out_len = tr.step(batch, training=False)
        per = tr.decode(logp, out_len, batch)
        losses.append(loss.double())
        pers.append(per.double())
    res = torch.stack([torch.stack(losses).mean(), torch.stack(pers).mean()]).tolist()
    model._engine = None
    return dict(arch=arch, loss=res[0], per=res[1])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--limit', type=int, default=16)
    ap.add_argument('--arch-file', default=None, help='JSON list of arch_vecs (e.g. the 8242 unique ones)')
    ap.add_argument('--batch', type=int, default=64)
    ap.add_argument('--frames', type=int, default=320)
    ap.add_argument('--n-batches', type=int, default=2)
    ap.add_argument('--precision', default='bf16')
    ap.add_argument('--balance', default='lpt', choices=['lpt', 'rr'])
    ap.add_argument('--out', default=None)
    ap.add_argument('--out-pickle', default=None, help='nb-asr style table: header, then rows')
    ap.add_argument('--all', action='store_true', help='evaluate every arch_vec, not one per isomorphism class')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    # Ranks map round-robin onto the visible GPUs; more ranks than GPUs is allowed and useful: a candidate costs
    # 0.3-1.1 s of host time (the reference's same-seed CPU initialisation) against tens of ms of GPU time, so several
    # processes per GPU keep it busy.  Rows are only gathered at the end -> gloo, no NCCL communicator needed.
    n_gpus = max(1, torch.cuda.device_count())
    local = local % n_gpus
    torch.cuda.set_device(local)
    if world > 1:
        torch.distributed.init_process_group('gloo')
    if args.arch_file:
        archs = json.load(open(args.arch_file))
    elif args.all:
        archs = list(search_space.get_all_architectures())
    else:
        archs = [a for _, a in graph_utils.get_unique_architectures()]
    archs = archs[:args.limit] if args.limit else archs
    mine = distributed.shard_archs(archs, rank, world, balance=args.balance, frames=args.frames)
    batches = [data.make_batch(args.batch, args.frames, seed=100 + i, min_len=args.frames // 3) for i in range(args.n_batches)]
    dev = torch.device('cuda', local)
    batches = [((a.to(dev), al.to(dev)), (t.to(dev), tl.to(dev))) for (a, al), (t, tl) in batches]
    t0 = time.time()
    rows = []
    for i in mine:
        r = evaluate_arch(archs[i], batches, local, args.precision)
        r['index'] = i
        r['hash'] = graph_utils.get_model_hash(archs[i])
        rows.append(r)
    torch.cuda.synchronize()
    dt = time.time() - t0
    rows = distributed.gather_rows(rows)
    if rank == 0:
        rows.sort(key=lambda r: r['index'])
        audio_s = sum(float(al.sum()) for (a, al), _ in batches) / 100.0 * len(archs)
        summary = dict(n_archs=len(archs), world=world, seconds=dt, archs_per_s=len(archs) / dt, audio_s_per_s=audio_s / dt)
        print(json.dumps(summary))
        if args.out:
            json.dump(dict(summary=summary, rows=rows), open(args.out, 'w'))
        if args.out_pickle:
            header = dict(dataset_type='b200-sweep-eval', version=1, search_space=search_space.get_search_space(),
                          ops=search_space.all_ops, columns=['model_hash', 'arch_vec', 'ctc_loss', 'per'], seed=1235,
                          precision=args.precision, batch=args.batch, frames=args.frames, n_batches=args.n_batches,
                          data='synthetic', decode='greedy')
            with open(args.out_pickle, 'wb') as f:
                pickle.dump(header, f)
                pickle.dump([[r['hash'], r['arch'], r['loss'], r['per']] for r in rows], f)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == '__main__':
    main()
