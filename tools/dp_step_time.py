"""Data-parallel train-step time under torchrun (one rank per GPU): the timed loop of bench.py without its other legs.
Usage: [NCCL_MAX_CTAS=8] [NBASR_DP_BUCKETS=0] python -m torch.distributed.run --nproc-per-node N ... tools/dp_step_time.py [steps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import nb_asr_b200 as nb  # noqa: E402

rank, world, local = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
if world > 1:
    os.environ.setdefault('NCCL_DEBUG', 'WARN')
    dist.init_process_group('nccl', device_id=dev)
nb.set_seed(1235)
model = nb.get_model([[1, 0], [1, 0, 0], [1, 0, 0, 0]], use_rnn=True, dropout_rate=0.0, gpu=local, precision='bf16')
model.train()
tr = nb.get_trainer((nb.PhonemeEncoder(48), None, None, None), nb.get_loss(), gpus=[local], verbose=False)
tr.model = tr._model = model
tr.optimizer = nb.trainer.FusedAdam(model, lr=1e-4)
tr.use_graph = True
(audio, alen), (tg, tl) = nb.data.make_batch(64, 500, seed=rank, min_len=500, tgt_lo=20, tgt_hi=50, pin=True)
batch = ((audio.to(dev), alen.to(dev)), (tg.to(dev), tl.to(dev)))
for _ in range(5):
    tr.step(batch, training=True)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    tr.step(batch, training=True)
e1.record()
torch.cuda.synchronize()
t = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f'world {world} buckets={os.environ.get("NBASR_DP_BUCKETS", "1")} NCCL_MAX_CTAS={os.environ.get("NCCL_MAX_CTAS")} '
          f'NCCL_ALGO={os.environ.get("NCCL_ALGO")}: {float(t):.3f} ms/step (max over ranks), {world * 64 / float(t) * 1e3:.0f} utt/s', flush=True)
if world > 1:
    dist.destroy_process_group()
