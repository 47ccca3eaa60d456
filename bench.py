#!/usr/bin/env python
"""bench.py -- IST-GCN training throughput (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl istgcn|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one full training iteration (forward, cross-entropy, backward, SGD-nesterov step)
of the composite IST-GCN (net.ist_gcn: 'ntu-rgb+d_sym' + 'spatial_3_sym', Inception GCN,
1x1-bottleneck Inception TCN, 10 blocks, dropout 0.5) on one batch of 64 synthetic NTU-shape
clips (3, 300, 25, 2) per GPU -- BASELINE.json configs[1] / configs[3].  Weak scaling: every
rank gets its own 64 clips, BatchNorm statistics stay per rank, gradients are averaged with
bucketed NCCL all-reduces overlapped with the backward pass.

Rank 0 prints ONE JSON line:
  value        clips/s over all ranks with the batch resident in HBM (CUDA events, max over ranks)
  e2e          the same step driven from pinned HOST buffers: H2D copy of the batch and a D2H
               read of the loss inside the timed region, every step
  roofline     the dominant kernel (largest share of the step): algorithmic bytes of its launches
               / their CUDA-event time, against the measured HBM copy bandwidth
  cpu_baseline the oracle port of the reference (plain PyTorch fp32, all host threads) timed on
               a bounded sample of the same workload (rank 0, N=1 only)
``--impl reference`` times only that CPU port (the reference is pure PyTorch; /root/reference
itself does not exist on the GPU box) and prints the same line with "impl": "reference".
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(ROOT, 'ist-gcn_b200'), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import torch.nn.functional as F  # noqa: E402

WORKLOADS = {
    'ntu': dict(graph_args=dict(layout='ntu-rgb+d_sym', strategy='spatial_3_sym'), num_class=60,
                shape=(3, 300, 25, 2), name='IST-GCN fwd+bwd+SGD, NTU-RGB+D shape (3,300,25,2)'),
    'kinetics': dict(graph_args=dict(layout='openpose_sym', strategy='spatial_3_sym'), num_class=400,
                     shape=(3, 300, 18, 2), name='IST-GCN fwd+bwd+SGD, Kinetics-skeleton shape (3,300,18,2)'),
}
BLOCKS = ((3, 64, 1, 0), (64, 64, 1, 1), (64, 64, 1, 1), (64, 64, 1, 1), (64, 128, 2, 2),
          (128, 128, 1, 1), (128, 128, 1, 1), (128, 256, 2, 2), (256, 256, 1, 1), (256, 256, 1, 1))


def measured_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.isfile(path):
        with open(path) as f:
            d = json.load(f)
        return float(d.get('hbm_gbs', 6650.0)), 'measured (MEASURED_PEAKS.json)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    QUERY = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.proc, self.lines = None, []
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(index), '--query-gpu=' + self.QUERY,
                 '--format=csv,noheader,nounits', '-lms', '200'],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap')
        for line in self.lines:
            parts = [s.strip() for s in line.split(',')]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx = float(parts[1])
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': mx,
                'reasons': sorted(reasons), 'samples': len(sm)}


# ------------------------------------------------------------------------------------------
def algorithmic_bytes(kernel, batch, shape):
    """Algorithmic (compulsory) HBM bytes of every launch of ``kernel`` in one training step:
    per-row figures of DESIGN.md section 4 x rows of each block (fp32 activations)."""
    C, T, V, M = shape
    NM = batch * M
    per_launch = []
    t = T
    for cin, cout, s, res in BLOCKS:
        tout = (t - 1) // s + 1
        r_in, r_out = NM * t * V, NM * tout * V
        if kernel == 'gcn_fwd':          # residual 1x1 conv (mma.sync engine): reads x, writes r
            if res == 2:
                per_launch.append(4 * r_out * (cin + cout))
        elif kernel == 'gcn_tc':         # forward: reads x, writes z (block 0 runs gcn_small_*)
            if cin >= 32:
                per_launch.append(4 * r_in * (cin + cout))
        elif kernel == 'gcn_tc_bwd':     # input gradient: reads dz, writes gin (reduce-add on top
            if cin >= 32:                # of the residual gradient: read + write)
                per_launch.append(4 * r_in * (cout + cin * (2 if res else 1)))
        elif kernel == 'bn_back_colsum':  # reads g1, z, writes dz (+ frame sums)
            if cin >= 32:
                per_launch.append(4 * r_in * 3 * cout)
        elif kernel == 'bn_back_apply':   # residual branch of the two strided blocks: reads go, rres, writes dyr
            if res == 2:
                per_launch.append(4 * r_out * 3 * cout)
        elif kernel == 'gcn_pair_grads':  # weight + adjacency gradient in one pass: reads dz, x
            if cin >= 32:
                per_launch.append(4 * r_in * (cout + cin))
        elif kernel == 'gcn_tc_dvals':   # reads dz, x
            if cin >= 32:
                per_launch.append(4 * r_in * (cout + cin))
        elif kernel == 'gcn_tc_dw':      # reads dz, x (the bias-term column sums ride on bn_back_colsum)
            if cin >= 32:
                per_launch.append(4 * r_in * (cout + cin))
        elif kernel == 'gcn_small_fwd':  # block 0: reads x (3 channels), writes z
            if cin < 32:
                per_launch.append(4 * r_in * (cin + cout))
        elif kernel == 'gcn_small_bwd':  # block 0: reads g1, z, x, writes dx
            if cin < 32:
                per_launch.append(4 * r_in * (2 * cout + 2 * cin))
        elif kernel == 'gcn_bwd_x':      # residual conv only: reads go, rres, read-modify-write gin
            if res == 2:
                per_launch.append(4 * r_out * (2 * cout + 2 * cin))
        elif kernel == 'gcn_bwd_w':
            if res == 2:
                per_launch.append(4 * r_out * (2 * cout + cin))
        elif kernel == 'tcn2_down':      # reads z, writes h1
            bp = 8 if int(cout ** 0.5) <= 8 else 16
            per_launch.append(4 * r_in * (cout + bp))
        elif kernel == 'tcn2_conv':      # reads h1, writes h2
            bp = 8 if int(cout ** 0.5) <= 8 else 16
            per_launch.append(4 * (r_in + r_out) * bp)
        elif kernel == 'tcn2_up':        # reads h2, writes u
            bp = 8 if int(cout ** 0.5) <= 8 else 16
            per_launch.append(4 * r_out * (cout + bp))
        elif kernel == 'tcn2_bwd_up':    # reads go, u, h2, writes dh2
            bp = 8 if int(cout ** 0.5) <= 8 else 16
            per_launch.append(4 * r_out * (2 * cout + 2 * bp))
            if cin < 32:                 # block 0's graph-conv backward runs on the same kernel:
                per_launch.append(4 * r_in * (2 * cout + 2 * 16))      # reads g1, z, X', writes G
        elif kernel == 'joint_colsum':   # block 0: per-joint sums of g1
            if cin < 32:
                per_launch.append(4 * r_in * cout)
        elif kernel == 'gcn_small_bwd_post':   # block 0: reads G (16 wide), x, writes dx
            if cin < 32:
                per_launch.append(4 * r_in * (16 + 2 * cin))
        elif kernel == 'tcn2_bwd_conv':  # reads dh2, h1 (twice: data + weight kernel), writes dh1
            bp = 8 if int(cout ** 0.5) <= 8 else 16
            per_launch.append(4 * (2 * r_in + 2 * r_out) * bp)
        elif kernel == 'tcn2_bwd_down':  # reads z, dh1, writes g1
            bp = 8 if int(cout ** 0.5) <= 8 else 16
            per_launch.append(4 * r_in * (2 * cout + 2 * bp))
        elif kernel == 'tcn_fwd':        # reads z, writes u (+ h1, h2 write, h1 read)
            bp = 8 if int(cout ** 0.5) <= 8 else 16
            per_launch.append(4 * (r_in * (cout + 2 * bp) + r_out * (cout + bp)))
        elif kernel == 'tcn_bwd':        # reads go, u, z, h1, h2 (+dh2, dh1 round trips), writes g1
            bp = 8 if int(cout ** 0.5) <= 8 else 16
            per_launch.append(4 * (r_out * (2 * cout + 3 * bp) + r_in * (2 * cout + 3 * bp)))
        elif kernel == 'block_tail_fwd':
            per_launch.append(4 * r_out * cout * (2 + (1 if res else 0)))
        elif kernel == 'block_tail_bwd':
            per_launch.append(4 * r_out * cout * (4 + (1 if res == 2 else 0)))
        t = tout
    return per_launch


ARCHS = ('ist_gcn', 'st_gcn_mstcn_1x1', 'st_gcn', 'st_gcn_msgcn', 'st_gcn_mstcn')


def graph_args_for(workload, arch):
    """The symmetric 4-partition graph for the Inception-GCN variants, the reference's default
    3-partition 'spatial' graph of the same skeleton for the single-adjacency ones."""
    g = dict(WORKLOADS[workload]['graph_args'])
    if arch not in ('ist_gcn', 'st_gcn_msgcn'):
        g = dict(layout=g['layout'].replace('_sym', ''), strategy='spatial')
    return g


def build_model(workload, device, dropout=0.5, arch='ist_gcn'):
    import importlib
    from istgcn import trainer
    w = WORKLOADS[workload]
    cls = importlib.import_module('net.' + arch).Model
    model = cls(w['shape'][0], w['num_class'], graph_args_for(workload, arch), True, dropout=dropout)
    model.apply(trainer.weights_init)
    return model.to(device)


def run_istgcn(args):
    import istgcn
    from istgcn import _lib, dp, trainer
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device (the product path has no CPU fallback)')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    assert world == args.gpus or world == 1, 'launch with torchrun --nproc-per-node %d' % args.gpus
    istgcn.set_math(args.math)
    w = WORKLOADS[args.workload]
    torch.manual_seed(0)
    model = build_model(args.workload, dev, arch=args.arch)
    dp.broadcast_state(model)
    tr = trainer.Trainer(model, base_lr=0.01, use_graph=not args.no_graph)
    torch.manual_seed(1000 + rank)
    B = args.batch
    x = torch.randn(B, *w['shape'], device=dev)
    y = torch.randint(0, w['num_class'], (B,), device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        tr.step(x, y)
    barrier()
    # launches per step are counted on an eager iteration (a graph replay re-launches the same
    # kernels without passing through the Python binding)
    launches0 = _lib.launch_count
    tr._iteration(x, y, True)
    launches_per_step = _lib.launch_count - launches0
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = tr.step(x, y)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if sampler else None
    launches = launches_per_step * args.steps
    # per-kernel device time: CUDA events around every launch of an eager pass of the same step
    # (events cannot be recorded inside a graph replay)
    # In this pass the weight-gradient kernels run on the main stream like everything else
    # (ISTGCN_PAIR_ASYNC=0): a kernel timed while another stream's kernel shares the SMs would be
    # charged for both.
    _lib.timing = {}
    prof_steps = min(3, args.steps)
    async_env = os.environ.get('ISTGCN_PAIR_ASYNC')
    os.environ['ISTGCN_PAIR_ASYNC'] = '0'
    torch.cuda.synchronize()
    ep0, ep1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ep0.record()
    for _ in range(prof_steps):
        tr._iteration(x, y, True)
    ep1.record()
    barrier()
    if async_env is None:
        os.environ.pop('ISTGCN_PAIR_ASYNC', None)
    else:
        os.environ['ISTGCN_PAIR_ASYNC'] = async_env
    timing, _lib.timing = _lib.timing, None
    eager_ms = ep0.elapsed_time(ep1)
    t_ms = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms = t_ms.item()
    ms_per_step = ms / args.steps
    value = world * B * args.steps / (ms / 1e3)
    final_loss = loss.item()

    ms_per_step_ = ms / args.steps
    # the *_bn entry points are the same kernels with the BatchNorm bookkeeping folded in
    merged = {}
    for k, v in timing.items():
        merged.setdefault(k[:-3] if k.endswith('_bn') else k, []).extend(v)
    timing = merged
    per_kernel = {k: sum(a.elapsed_time(b) for a, b in v) for k, v in timing.items()}
    top = max(per_kernel, key=per_kernel.get)
    peak, peak_src = measured_peaks()
    # the per-launch byte model below describes the IST-GCN block table only
    alg = algorithmic_bytes(top, B, w['shape']) if args.arch == 'ist_gcn' else []
    if alg and top == 'gcn_tc':          # the same entry point serves forward and input gradient
        alg = alg + algorithmic_bytes('gcn_tc_bwd', B, w['shape'])
    n_launch = len(timing[top])
    alg_total = sum(alg) * prof_steps if alg else None
    roof = {'kernel': top, 'bound': 'hbm', 'peak': peak, 'unit': 'GB/s', 'peak_source': peak_src,
            'launches': n_launch, 'share_of_step': per_kernel[top] / prof_steps / ms_per_step_,
            'traffic': None,
            'timed_in': 'eager pass of %d steps, every kernel on the main stream (%.1f ms/step)' % (prof_steps, eager_ms / prof_steps)}
    tpath = os.path.join(ROOT, 'profiles', 'r2_traffic.json')
    if os.path.isfile(tpath) and args.arch == 'ist_gcn':
        with open(tpath) as f:          # DRAM bytes per launch from the committed ncu capture
            tr_rec = json.load(f).get(top)
        if tr_rec and tr_rec.get('workload') == args.workload and tr_rec.get('clips_per_gpu') == B:
            roof['traffic'] = tr_rec['dram_bytes_per_launch']
            roof['traffic_source'] = tr_rec['source']
    if alg_total:
        ach = alg_total / (per_kernel[top] / 1e3) / 1e9
        roof.update(achieved=ach, frac=ach / peak,
                    algorithmic_bytes_per_launch=alg_total / n_launch,
                    avg_launch_ms=per_kernel[top] / n_launch)
    shares = {k: round(v / prof_steps / ms_per_step_, 4)
              for k, v in sorted(per_kernel.items(), key=lambda kv: -kv[1])}
    # every entry point with a byte model: device time per step, algorithmic GB/s, fraction of peak
    table = []
    if args.arch == 'ist_gcn':
        for k, v in sorted(per_kernel.items(), key=lambda kv: -kv[1]):
            ab = algorithmic_bytes(k, B, w['shape'])
            if k == 'gcn_tc':
                ab = ab + algorithmic_bytes('gcn_tc_bwd', B, w['shape'])
            if not ab:
                continue
            gbs = sum(ab) * prof_steps / (v / 1e3) / 1e9
            table.append({'entry_point': k, 'ms_per_step': round(v / prof_steps, 4), 'launches_per_step': len(ab),
                          'algorithmic_GB_per_step': round(sum(ab) / 1e9, 3), 'GBps': round(gbs, 1),
                          'frac_of_hbm_peak': round(gbs / peak, 3)})

    # end to end through the public API from pinned HOST buffers: the input pipeline of the trainer
    # (istgcn.pipeline.DevicePrefetcher: copy stream, two slots -- the H2D copy of step i+1 overlaps
    # step i) feeds Trainer.step; the loss of every step is read back to the host
    from istgcn import pipeline
    ring = [(torch.randn(B, *w['shape']).pin_memory(), torch.randint(0, w['num_class'], (B,)).pin_memory())
            for _ in range(3)]

    def host_loader(n):
        for i in range(n):
            yield ring[i % len(ring)]

    def e2e_run(n):
        pf = pipeline.DevicePrefetcher(host_loader(n), dev)
        last = None
        for xd, yd in pf:
            last = tr.step(xd, yd).item()        # D2H read of the loss, every step
        return pf.h2d_bytes, last

    e2e_run(3)
    barrier()
    t0 = time.perf_counter()
    h2d_total, _ = e2e_run(args.steps)
    barrier()
    dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    e2e_value = world * B * args.steps / dt.item()
    h2d_per_step = h2d_total // args.steps

    # forward-only (eval) throughput, reported next to the training number
    with torch.no_grad():
        for _ in range(2):
            tr.evaluate(x)
        barrier()
        e0.record()
        for _ in range(args.steps):
            tr.evaluate(x)
        e1.record()
        barrier()
        fwd = world * B * args.steps / (e0.elapsed_time(e1) / 1e3)

    replicas_ok = dp.replicas_equal(model) if world > 1 else None
    extras = {}
    if world == 1 and not args.no_extras:
        del tr, model
        torch.cuda.empty_cache()
        extras = run_extras(args, dev)

    line = {
        'metric': '%s_train_clips_per_s' % args.arch, 'value': value, 'unit': 'clips/s',
        'n_gpus': world, 'steps': args.steps, 'warmup': max(3, args.warmup),
        'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'tf32' if args.math == 'tf32' else 'fp32(3xtf32)',
        'data': 'synthetic',
        'config': {'workload': w['name'].replace('IST-GCN', args.arch) if args.arch != 'ist_gcn' else w['name'],
                   'clips_per_gpu': B, 'global_batch': B * world,
                   'dropout': 0.5, 'optimizer': 'SGD(momentum .9, nesterov, wd 1e-4)',
                   'parallelism': 'dp%d' % world, 'activations': 'fp32 channels-last',
                   'cuda_graph': not args.no_graph,
                   'l2_policy': 'working set (>10 GB of activations per step) is far larger than the 126 MB L2'},
        'clocks': clocks, 'gpu_launches': launches,
        'e2e': {'value': e2e_value, 'unit': 'clips/s',
                'h2d_bytes_per_step': h2d_per_step, 'd2h_bytes_per_step': 4,
                'pipeline': 'pinned host batches -> copy stream (double buffer) -> Trainer.step -> loss.item()'},
        'roofline': roof, 'kernel_shares': shares, 'per_entry_point': table, 'fwd_clips_per_s': fwd,
        'loss': final_loss,
    }
    if replicas_ok is not None:
        line['replicas_equal_after_run'] = replicas_ok
        assert replicas_ok, 'data-parallel replicas diverged during the timed loop'
    line.update(extras)
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            line['cpu_baseline'] = cpu_reference(args, steps=1, warmup=1)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()



# ------------------------------------------------------------------------------------------
def _timed_training(workload, arch, math, batch, dev, steps, warmup=3, use_graph=True):
    """clips/s of ``steps`` training iterations (CUDA events) of a freshly built model."""
    import istgcn
    from istgcn import trainer
    w = WORKLOADS[workload]
    old = istgcn.set_math(math)
    try:
        torch.manual_seed(0)
        model = build_model(workload, dev, arch=arch)
        tr = trainer.Trainer(model, base_lr=0.01, use_graph=use_graph)
        x = torch.randn(batch, *w['shape'], device=dev)
        y = torch.randint(0, w['num_class'], (batch,), device=dev)
        for _ in range(max(3, warmup)):
            tr.step(x, y)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            loss = tr.step(x, y)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        out = {'clips_per_s': batch / (ms / 1e3), 'ms_per_step': ms, 'batch': batch, 'loss': loss.item()}
    finally:
        istgcn.set_math(old)
    del tr, model, x, y
    torch.cuda.empty_cache()
    return out


def gpu_eager_reference(args, dev, steps=3):
    """The real competitor: the reference's own algorithm as stock
    PyTorch eager ops (conv2d / einsum / batch_norm -> cuDNN / cuBLAS) on the SAME GPU, same
    batch, fwd + bwd + SGD(nesterov), with TF32 enabled (PyTorch's default for convolutions plus
    matmul TF32) and with full fp32.  Uses the fixture-pinned oracle port of the reference's
    modules; a reported baseline, never part of the product path."""
    from net.utils.graph import Graph
    from oracle import model_ref
    w = WORKLOADS[args.workload]
    g = Graph(**graph_args_for(args.workload, args.arch))
    out = {'kind': 'oracle port of the reference modules, PyTorch %s eager on cuda' % torch.__version__,
           'batch': args.batch}
    for label, flags in (('tf32', (True, True)), ('fp32', (False, False))):
        old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = flags
        try:
            torch.manual_seed(0)
            state = model_ref.make_state(args.arch, w['shape'][0], w['num_class'], g.A,
                                         getattr(g, 'A2', None), getattr(g, 'A3', None), seed=0)
            state = {k: v.to(dev) for k, v in state.items()}
            names = [k for k, v in state.items() if v.is_floating_point() and 'running' not in k
                     and k not in ('A', 'A2', 'A3') and '.gcn.branch.bn.' not in k and '.linear.' not in k]
            params = [state[k].requires_grad_(True) for k in names]
            opt = torch.optim.SGD(params, lr=0.01, momentum=0.9, nesterov=True, weight_decay=1e-4,
                                  foreach=True)
            x = torch.randn(args.batch, *w['shape'], device=dev)
            y = torch.randint(0, w['num_class'], (args.batch,), device=dev)

            def step():
                opt.zero_grad(set_to_none=True)
                loss = F.cross_entropy(model_ref.forward(state, x, args.arch, training=True, dropout=0.5), y)
                loss.backward()
                opt.step()
                return loss

            for _ in range(2):
                step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                step()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            out[label] = {'clips_per_s': args.batch / (ms / 1e3), 'ms_per_step': ms,
                          'peak_mem_gb': torch.cuda.max_memory_allocated(dev) / 2 ** 30}
        finally:
            torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
        del state, params, opt, x, y
        torch.cuda.empty_cache()
    return out


def build_twostream(workload, dev, arch):
    """net.st_gcn_twostream.Model (reference net/st_gcn_twostream.py:11-26: joint stream + motion
    stream, logits summed) -- or, with another --arch, the same wrapper around that network."""
    import importlib
    import net.st_gcn_twostream as ts
    from istgcn import trainer
    w = WORKLOADS[workload]
    args_ = (w['shape'][0], w['num_class'], graph_args_for(workload, arch), True)
    model = ts.Model(*args_)
    if arch != 'st_gcn':
        cls = importlib.import_module('net.' + arch).Model
        model.origin_stream, model.motion_stream = cls(*args_), cls(*args_)
    model.apply(trainer.weights_init)
    return model.to(dev).eval()


def twostream_sweep(args, dev, batches, world=1, rank=0, chunk=512, reps=5):
    """BASELINE.json configs[4]: two-stream inference, batch sweep.  A batch is sharded over the
    ranks (batch < world: fewer ranks active); each rank runs its share in chunks of at most
    ``chunk`` clips through one CUDA graph per chunk shape.  Per point: whole-job clips/s and
    the latency of the batch (device time, max over ranks)."""
    w = WORKLOADS[args.workload]
    model = build_twostream(args.workload, dev, args.ts_arch)
    graphs = {}

    @torch.no_grad()
    def run_chunk(xc):
        n = xc.shape[0]
        if n not in graphs:
            sx = torch.empty_like(xc)
            sx.copy_(xc)
            for _ in range(2):
                model(sx)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                out = model(sx)
            graphs[n] = (g, sx, out)
        g, sx, out = graphs[n]
        sx.copy_(xc, non_blocking=True)
        g.replay()
        return out

    points = []
    for b in batches:
        share = b // world + (1 if rank < b % world else 0)
        x = torch.randn(max(share, 1), *w['shape'], device=dev) if share else None

        def run_batch():
            outs = []
            for i in range(0, share, chunk):
                outs.append(run_chunk(x[i:i + chunk]).clone())
            return outs

        if share:
            run_batch()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            if share:
                run_batch()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
        points.append({'batch': b, 'latency_ms': ms, 'clips_per_s': b / (ms / 1e3),
                       'active_ranks': min(b, world)})
        del x
    del model, graphs
    torch.cuda.empty_cache()
    return points


def tf32_gemm_peak(dev, n=8192, iters=10):
    """cuBLAS TF32 GEMM n^3 on this GPU (TFLOP/s): the library's tensor-pipe rate for the arithmetic
    the tcgen05 kernels use (tools/microbench/mma_rate.cu measures the instruction rate itself)."""
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        a = torch.randn(n, n, device=dev)
        b = torch.randn(n, n, device=dev)
        for _ in range(2):
            a @ b
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            a @ b
        e1.record()
        torch.cuda.synchronize()
        return iters * 2.0 * n ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev


def run_extras(args, dev):
    """Driver-visible datapoints beside the headline (N = 1 only): the <= 1e-4 parity mode's
    throughput, BASELINE.json configs[2] (Kinetics-skeleton, batch 256), the stock-PyTorch-eager
    competitor on this GPU, and three points of the configs[4] two-stream inference sweep."""
    import copy
    out = {}
    steps = max(2, min(args.steps, 5))
    if args.math == 'tf32':
        out['value_3xtf32'] = _timed_training(args.workload, args.arch, '3xtf32', args.batch, dev, steps)
    if args.workload == 'ntu' and args.arch == 'ist_gcn':
        out['kinetics_b256'] = _timed_training('kinetics', args.arch, args.math, 256, dev, steps)
    out['gpu_eager'] = gpu_eager_reference(args, dev)
    out['tf32_gemm_tflops'] = tf32_gemm_peak(dev)
    ts_args = copy.copy(args)
    out['twostream_inference'] = {'stream_arch': args.ts_arch,
                                  'points': twostream_sweep(ts_args, dev, [1, 64, 1024])}
    return out


def run_sweep(args):
    """``--sweep``: the full configs[4] batch sweep 1..4096 (one JSON line)."""
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    import istgcn
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    istgcn.set_math(args.math)
    batches = [1 << i for i in range(13)]
    points = twostream_sweep(args, dev, batches, world, rank)
    if rank == 0:
        w = WORKLOADS[args.workload]
        print(json.dumps({'metric': 'twostream_inference_clips_per_s', 'unit': 'clips/s', 'n_gpus': world,
                          'dtype': args.math, 'data': 'synthetic', 'higher_is_better': True,
                          'config': {'workload': 'two-stream (joint + motion) inference, %s, eval' % w['name'],
                                     'stream_arch': args.ts_arch, 'cuda_graph': True,
                                     'sharding': 'batch over ranks, chunks of <= 512 clips per graph'},
                          'points': points}))
    if world > 1:
        dist.destroy_process_group()

# ------------------------------------------------------------------------------------------
def cpu_reference(args, steps, warmup):
    """The reference's own algorithm (oracle port: plain PyTorch fp32 conv/einsum/batch_norm)
    on the host cores: full training steps on a bounded sample of the workload."""
    from net.utils.graph import Graph
    from oracle import model_ref
    w = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    g = Graph(**graph_args_for(args.workload, args.arch))
    torch.manual_seed(0)
    state = model_ref.make_state(args.arch, w['shape'][0], w['num_class'], g.A, getattr(g, 'A2', None),
                                 getattr(g, 'A3', None), seed=0)
    names = [k for k, v in state.items() if v.is_floating_point() and 'running' not in k
             and k not in ('A', 'A2', 'A3') and '.gcn.branch.bn.' not in k]
    for k in names:
        state[k].requires_grad_(True)
    bs = args.cpu_batch
    x = torch.randn(bs, *w['shape'])
    y = torch.randint(0, w['num_class'], (bs,))
    bufs = [None] * len(names)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        out = model_ref.forward(state, x, args.arch, training=True, dropout=0.5)
        loss = F.cross_entropy(out, y)
        grads = torch.autograd.grad(loss, [state[k] for k in names], allow_unused=True)
        with torch.no_grad():
            ps = [state[k] for k, gr in zip(names, grads) if gr is not None]
            gs = [gr for gr in grads if gr is not None]
            model_ref.sgd_nesterov_step(ps, gs, bufs[:len(ps)], 0.01)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    best = min(times)
    # BASELINE.json configs[0]: baseline ST-GCN (net.st_gcn) forward, (8, 3, 300, 25, 2), eval, CPU
    cfg1 = None
    if args.workload == 'ntu':
        g1 = Graph(layout='ntu-rgb+d', strategy='spatial')
        st1 = model_ref.make_state('st_gcn', 3, 60, g1.A, None, None, seed=0)
        x1 = torch.randn(8, 3, 300, 25, 2)
        t1 = []
        with torch.no_grad():
            for it in range(2):
                t0 = time.perf_counter()
                model_ref.forward(st1, x1, 'st_gcn', training=False)
                t1.append(time.perf_counter() - t0)
        cfg1 = {'metric': 'st_gcn_cpu_fwd_clips_per_s', 'value': 8 / min(t1), 'batch': 8,
                'threads': torch.get_num_threads()}
    return {'value': bs / best, 'unit': 'clips/s', 'cores': cores, 'kind': 'port', 'cfg1_st_gcn_cpu_forward': cfg1,
            'sample': '%d clip(s) of the same shape per step, %d timed step(s), best' % (bs, steps),
            'torch_threads': torch.get_num_threads(), 's_per_step': best}


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    w = WORKLOADS[args.workload]
    steps, warmup = max(1, min(args.steps, 3)), max(1, min(args.warmup, 1))
    base = cpu_reference(args, steps, warmup)
    line = {'impl': 'reference', 'metric': '%s_train_clips_per_s' % args.arch, 'value': base['value'],
            'unit': 'clips/s', 'n_gpus': args.gpus, 'steps': steps, 'warmup': warmup,
            'ms_per_step': base['s_per_step'] * 1e3, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'fp32', 'data': 'synthetic',
            'config': {'workload': w['name'], 'clips_per_step': args.cpu_batch,
                       'note': 'reference = pure PyTorch; its CPU path timed on the host cores'},
            'cpu_baseline': base,
            'e2e': {'value': base['value'], 'unit': 'clips/s', 'h2d_bytes_per_step': 0,
                    'd2h_bytes_per_step': 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='istgcn', choices=['istgcn', 'reference'])
    ap.add_argument('--workload', default='ntu', choices=sorted(WORKLOADS))
    ap.add_argument('--arch', default='ist_gcn', choices=ARCHS,
                    help='network variant (the headline metric is quoted on ist_gcn)')
    ap.add_argument('--batch', type=int, default=64, help='clips per GPU')
    ap.add_argument('--math', default='tf32', choices=['tf32', '3xtf32'])
    ap.add_argument('--cpu-batch', type=int, default=4, help='clips per CPU-baseline step')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-graph', action='store_true', help='eager launches instead of a CUDA graph')
    ap.add_argument('--no-extras', action='store_true',
                    help='skip value_3xtf32 / kinetics_b256 / gpu_eager / twostream_inference')
    ap.add_argument('--sweep', action='store_true',
                    help='two-stream inference batch sweep 1..4096 (BASELINE.json configs[4])')
    ap.add_argument('--ts-arch', default='st_gcn', choices=ARCHS,
                    help='network of each stream of the two-stream model (reference: net.st_gcn)')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    elif args.sweep:
        run_sweep(args)
    else:
        run_istgcn(args)


if __name__ == '__main__':
    main()
