#!/usr/bin/env python
"""Headline benchmark: NAS-Bench-ASR candidate TRAIN step (fwd + CTC + bwd + reg + clip + Adam) throughput.

    python bench.py --gpus N --steps K --warmup W            # ours (one process per GPU under torchrun for N>1)
    python bench.py --impl reference --gpus N --steps K ...   # the reference algorithm on the host CPU cores

Workload (BASELINE.json configs[1]): arch [[1,0],[1,0,0],[1,0,0,0]], batch 64 x 500 frames x 80 log-mel per GPU,
bf16 conv/GEMM operands with fp32 accumulation, fp32 LayerNorm statistics / LSTM state / CTC.  Synthetic N(0,1)
log-mel, U{1..48} labels, weights from seed 1235.  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DEFAULT_ARCH = [[1, 0], [1, 0, 0], [1, 0, 0, 0]]
ARCHS = {'default': DEFAULT_ARCH, 'c7d2_skips': [[4, 1], [4, 1, 1], [4, 1, 1, 1]],
         'linear_skips': [[0, 1], [0, 1, 1], [0, 1, 1, 1]], 'mixed': [[2, 1], [3, 0, 1], [0, 1, 0, 1]]}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--arch', default='default')
    ap.add_argument('--batch', type=int, default=64)
    ap.add_argument('--frames', type=int, default=500)
    ap.add_argument('--precision', default='bf16', choices=['bf16', 'fp32'])
    ap.add_argument('--dropout', type=float, default=0.0)
    ap.add_argument('--profile', action='store_true', help='print the per-kernel-class breakdown to stderr')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-graph', action='store_true')
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d['hbm_gbs'], tf_burst=d['bf16_tflops'], tf_sust=d['bf16_tflops_sustained'], src='measured')
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src='fallback')


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = 'clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-lms', '100',
                                          '-i', str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(',')])

    def stop(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['nvidia-smi unavailable'])
        self.proc.terminate()
        time.sleep(0.05)
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[3:7]):
                    if v.lower().startswith('active'):
                        reasons.add(name)
            except (ValueError, IndexError):
                pass
        sm.sort()
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=mx, reasons=sorted(reasons), samples=len(sm))


def cpu_train_step_time(arch, B, T, steps, warmup):
    """The reference algorithm (oracle/model_ref.py torch-fp32 restatement, pinned to the real reference by
    tests/golden) on all host cores: one full train step on a B-utterance sample."""
    import torch
    from oracle import model_ref as M
    torch.set_num_threads(os.cpu_count() or 1)
    sd = M.build_state_dict(arch, seed=1235)
    audio, alen, tg, tl = M.make_batch(B, T, seed=0, min_len=T)
    st, times = None, []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        _, _, _, _, sd, st, _ = M.train_step(sd, arch, audio, alen, tg, tl, st)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    return sum(times) / len(times)


def run_reference(args, arch, rank):
    if rank != 0:
        return
    B = 8
    t = cpu_train_step_time(arch, B, args.frames, max(1, args.steps), max(0, args.warmup))
    val = B / t
    cores = os.cpu_count() or 1
    sample = f'each step = one full train step on {B} of the {args.batch} utterances ({B}x{args.frames}x80), fp32, all host threads'
    print(json.dumps({
        'impl': 'reference', 'metric': 'train_utterances_per_sec', 'value': val, 'unit': 'utt/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * t * args.batch / B, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': workload_config(args, arch, 1),
        'cpu_baseline': {'value': val, 'unit': 'utt/s', 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': val, 'unit': 'utt/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}))


def workload_config(args, arch, world):
    return {'workload': f'train step (fwd+CTC+bwd+reg+clip+Adam), arch {arch}, batch {args.batch}x{args.frames} frames x80 log-mel per GPU',
            'arch_vec': arch, 'per_gpu_batch': args.batch, 'global_batch': args.batch * world, 'frames': args.frames,
            'parallelism': f'dp{world} (utterance-sharded; the flat fp32 gradient is all-reduced by NCCL in 5 buckets overlapped with the backward pass)' if world > 1 else 'single GPU',
            'l2': 'per-step working set (activations + weights) is > 2 GB, far above the 126 MB L2; no flush needed',
            'dropout': args.dropout}


def main():
    args = parse()
    arch = ARCHS.get(args.arch) or json.loads(args.arch)
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    if args.impl == 'reference':
        run_reference(args, arch, rank)
        return
    import torch
    import torch.distributed as dist
    import nb_asr_b200 as nb
    from nb_asr_b200 import profiling
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        # keep stdout to the one JSON line: NCCL's banner / warnings go to stderr
        os.environ['NCCL_DEBUG'] = os.environ.get('NBASR_NCCL_DEBUG', 'WARN')
        os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')
        dist.init_process_group('nccl', device_id=dev)
    B, T = args.batch, args.frames
    nb.set_seed(1235)
    model = nb.get_model(arch, use_rnn=True, dropout_rate=args.dropout, gpu=local, precision=args.precision)
    model.train()
    tr = nb.get_trainer((nb.PhonemeEncoder(48), None, None, None), nb.get_loss(), gpus=[local], verbose=False)
    tr.model = tr._model = model
    tr.optimizer = nb.trainer.FusedAdam(model, lr=1e-4)
    tr.use_graph = not args.no_graph
    (audio, alen), (tg, tl) = nb.data.make_batch(B, T, seed=rank, min_len=T, tgt_lo=20, tgt_hi=50, pin=True)
    host_batch = ((audio, alen), (tg, tl))
    dev_batch = ((audio.to(dev), alen.to(dev)), (tg.to(dev), tl.to(dev)))
    eng = model.engine

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        timed.launches = eng.launches
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        timed.launches = eng.launches - timed.launches
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    # ---- device-resident throughput (value)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms = timed(lambda: tr.step(dev_batch, training=True), args.steps, max(3, args.warmup))
    clocks = sampler.stop() if rank == 0 else None
    launches_total = timed.launches
    ms_step = ms / args.steps
    value = world * B * args.steps / (ms / 1e3)

    # ---- end to end: pinned host inputs -> step -> loss on host, every step
    def e2e_step():
        loss, _, _ = tr.step(host_batch, training=True)
        return loss.item()
    ms_e2e = timed(e2e_step, args.steps, 3)
    e2e_val = world * B * args.steps / (ms_e2e / 1e3)
    h2d = sum(t.numel() * t.element_size() for t in (audio, alen, tg, tl))

    # ---- eval step (fwd + CTC loss + greedy decode + PER), audio-seconds per second
    model.eval()

    def eval_step():
        loss, logp, out_len = tr.step(dev_batch, training=False)
        return tr.decode(logp, out_len, dev_batch)
    ms_eval = timed(eval_step, args.steps, 3)
    audio_s = float(alen.sum()) / 100.0
    eval_val = world * audio_s * args.steps / (ms_eval / 1e3)
    model.train()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- live per-kernel timing (CUDA events on the launching stream) for the roofline entry
    pl = eng.plan(B, T, True)
    prof = profiling.profile_ops(eng, pl.fwd + pl.bwd, iters=3)
    # optimiser tail (regulariser + clip + Adam + operand re-pack), timed as one unit on the launching stream
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    eng._optimizer_launch()
    torch.cuda.synchronize()
    ev[0].record()
    for _ in range(3):
        eng._optimizer_launch()
    ev[1].record()
    torch.cuda.synchronize()
    n_par = float(eng.n_flat)
    prof['optim+repack'] = dict(ms=ev[0].elapsed_time(ev[1]) / 3, flops=0.0, bytes=n_par * (4 * 7 + 2 * 2), n=1.0)
    if args.profile:
        sys.stderr.write(profiling.format_profile(prof) + '\n')
    fam = {}
    for tag, d in prof.items():
        f = fam.setdefault(tag.split(' ')[0], dict(ms=0.0, flops=0.0, bytes=0.0, n=0, floor_cycles=0.0))
        for k in ('ms', 'flops', 'bytes', 'n', 'floor_cycles'):
            f[k] += d.get(k, 0.0)
    tot_ms = sum(f['ms'] for f in fam.values())
    dom_name, dom = max(fam.items(), key=lambda kv: kv[1]['ms'])
    pk = peaks()
    tensor_bound = dom_name in ('gemm_tn', 'gemm_wgrad')
    if tensor_bound:
        achieved = dom['flops'] / dom['ms'] / 1e9
        peak, unit = pk['tf_sust'], 'TFLOP/s'
    else:
        achieved = dom['bytes'] / dom['ms'] / 1e6
        peak, unit = pk['hbm'], 'GB/s'
    # DRAM traffic of the dominant kernel: per-launch dram read+write bytes of an `ncu --set full` capture of this same
    # command (profiles/r1_ncu_step_full.json, written by tools/ncu_summary.py); null when no capture is committed
    traffic, traffic_note = None, None
    fam_kernel = {'gconv': 'gconv_mma_fwd_kernel', 'gconv_wgrad': 'gconv_mma_wgrad_kernel', 'gemm_tn': 'gemm_tn_pair_kernel',
                  'gemm_wgrad': 'gemm_wgrad_pair_kernel', 'ln_bwd': 'layernorm_bwd_bf16_kernel', 'ln_fwd': 'layernorm_fwd_bf16_kernel'}
    ncu_path = os.path.join(ROOT, 'profiles', 'r1_ncu_step_full.json')
    if os.path.exists(ncu_path) and dom_name in fam_kernel:
        nc = json.load(open(ncu_path)).get('kernels', {}).get(fam_kernel[dom_name])
        if nc and 'avg_dram_bytes_per_launch' in nc:
            traffic = nc['avg_dram_bytes_per_launch']
            traffic_note = (f"dram read+write bytes per launch, ncu --set full over {nc['launches']} launches of "
                            f"{fam_kernel[dom_name]} (profiles/r1_ncu_step_full.json)")
    roofline = {'kernel': dom_name, 'bound': 'tensor' if tensor_bound else 'hbm', 'achieved': achieved, 'peak': peak, 'unit': unit,
                'frac': achieved / peak, 'traffic': traffic, 'traffic_note': traffic_note,
                'algorithmic_bytes_per_launch': dom['bytes'] / max(dom['n'], 1), 'algorithmic_flops_per_launch': dom['flops'] / max(dom['n'], 1), 'peak_source': pk['src'] + (' sustained' if tensor_bound else ''),
                'launches_per_step': dom['n'], 'avg_launch_ms': dom['ms'] / max(dom['n'], 1),
                'share_of_step_kernel_time': dom['ms'] / tot_ms,
                # the grouped-conv kernel is bound by tensor-pipe ISSUE, not by HBM or FLOPs: time of its
                # 3k MMAs per tile at 88 cycles each on every SM at the sampled clock, against the measured time
                'mma_issue_floor': ({'ms': dom['floor_cycles'] / torch.cuda.get_device_properties(dev).multi_processor_count /
                                     ((clocks or {}).get('sm_mhz') or 1965.0) / 1e3,
                                     'frac_of_measured': dom['floor_cycles'] / torch.cuda.get_device_properties(dev).multi_processor_count /
                                     ((clocks or {}).get('sm_mhz') or 1965.0) / 1e3 / dom['ms'],
                                     'note': '3k tcgen05.mma (N=48, operands in shared memory, >= 88 cycles each) per 128x48 tile'}
                                    if dom.get('floor_cycles') else None),
                'families': {k: {'ms': round(v['ms'], 4), 'n': v['n'],
                                 'tflops': round(v['flops'] / v['ms'] / 1e9, 1) if v['ms'] else 0,
                                 'gbs': round(v['bytes'] / v['ms'] / 1e6, 0) if v['ms'] else 0} for k, v in fam.items()}}

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        Bc = 8
        t = cpu_train_step_time(arch, Bc, T, steps=2, warmup=1)
        cpu = {'value': Bc / t, 'unit': 'utt/s', 'cores': os.cpu_count() or 1, 'kind': 'port',
               'sample': f'2 timed train steps (after 1 warm-up) on {Bc} of the {B} utterances ({Bc}x{T}x80), fp32 torch CPU '
                         f'restatement of the reference step (oracle/model_ref.py), all host threads'}

    out = {'metric': 'train_utterances_per_sec', 'value': value, 'unit': 'utt/s', 'n_gpus': world, 'steps': args.steps,
           'warmup': max(3, args.warmup), 'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
           'dtype': 'bf16' if args.precision == 'bf16' else 'f32', 'data': 'synthetic', 'config': workload_config(args, arch, world),
           'clocks': clocks,
           'e2e': {'value': e2e_val, 'unit': 'utt/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': 4, 'ms_per_step': ms_e2e / args.steps},
           'gpu_launches': launches_total,
           'eval': {'metric': 'eval_audio_seconds_per_sec', 'value': eval_val, 'unit': 'audio-s/s', 'ms_per_step': ms_eval / args.steps},
           'roofline': roofline, 'cpu_baseline': cpu}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
