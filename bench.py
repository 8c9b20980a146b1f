#!/usr/bin/env python
"""bench.py -- DSNT head throughput (heatmaps/s, fwd+bwd+regulariser) and fraction of the HBM roofline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[3], the one the metric is quoted on): per GPU a batch of 4096 samples x 16
MPII joints x 64x64 fp32 logits, Euclidean loss + JS regulariser (sigma = 1 px), joint mask; weak scaling
(every rank owns a fixed 4096-sample shard of a 4096*N batch; the three partial sums of masked_average are
exchanged between the ranks inside the finishing kernels over NVLink peer memory, or by NCCL where that is unavailable).

A "step" is one pass of the hot path over one batch: fused forward (coords, loss) + backward (dL/dZ), through the
public autograd API (`dsnt_head(...).loss.backward()`).  --path one-pass (default) lets the forward also write dL/dZ
while the heatmap is on chip (dsnt_head_step: 2*H*W*sizeof algorithmic bytes per heatmap); --path two-kernel is the
forward kernel + the streaming backward kernel (3*H*W*sizeof).  Heatmaps too large for the one-pass kernel
(256x256) take the two-kernel path either way; config.path says which ran.
  value  whole-job heatmaps/s with Z already resident in HBM (CUDA events, barrier + sync both sides, max
         over ranks).  Z is 1 GiB per step, far larger than the 126 MB L2, so no flush is needed.
  e2e    the same step through the public API with HOST buffers: pinned Z/target/mask -> device copies
         in, loss + coords device -> host out, all inside the timed region.
  roofline / cpu_baseline / clocks: see DESIGN.md "Measurement".

--impl reference times the reference's CPU implementation of the same path on this box's host cores.
The reference is pure Python on torch; its tree is not on the GPU box, so this arm executes the committed
restatement oracle/torch_port.py (the only place besides the cpu_baseline leg where bench.py executes oracle/).
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

RESULT_OUT = sys.stdout
METRIC = 'dsnt_head_heatmaps_per_sec'
UNIT = 'heatmaps/s'
WORKLOADS = {
    # name: (batch per GPU, joints, H, W, reg, dtype)
    'cfg4_64x64_f32_js': (4096, 16, 64, 64, 'js', 'f32'),
    'cfg4_64x64_bf16_js': (4096, 16, 64, 64, 'js', 'bf16'),
    'cfg4_64x64_f32_var': (4096, 16, 64, 64, 'var', 'f32'),
    'cfg5_256x256_f32_var': (512, 16, 256, 256, 'var', 'f32'),
    'cfg1_64x64_f32_js': (32, 16, 64, 64, 'js', 'f32'),
    'cfg2_28x28_f32_js': (64, 16, 28, 28, 'js', 'f32'),
}
DEFAULT_WORKLOAD = 'cfg4_64x64_f32_js'
CPU_SAMPLE = (32, 16, 64, 64)          # BASELINE cfg 1 shape: what the reference's CPU path is timed on


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--path', default='one-pass', choices=['one-pass', 'two-kernel'])
    ap.add_argument('--no-graph', dest='graph', action='store_false',
                    help='time eager autograd calls instead of replaying the captured step (CUDA graph)')
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------- helpers
def measured_peak():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md)'


def committed_traffic(workload):
    """dram__bytes_read+write per launch from the committed ncu summary, if one exists for this workload."""
    try:
        with open(os.path.join(ROOT, 'profiles', 'roofline_traffic.json')) as f:
            return json.load(f).get(workload)
    except Exception:
        return None


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons of one GPU while the timed region runs."""

    FIELDS = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
              'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
              'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.FIELDS, '--format=csv,noheader,nounits',
                 '-lms', '100'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        rows = [ln for (ts, ln) in self.lines if t0 - 0.05 <= ts <= t1 + 0.15] or [ln for (_, ln) in self.lines]
        sm, smax, power, reasons = [], [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for ln in rows:
            parts = [p.strip() for p in ln.split(',')]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                smax.append(float(parts[1]))
                power.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(smax) if smax else None,
                'power_w_max': max(power) if power else None, 'samples': len(sm), 'reasons': sorted(reasons)}


def leave(world, dist, torch):
    """End of a multi-rank run: flush, meet the other ranks once, and leave WITHOUT tearing NCCL down.
    destroy_process_group() with a live CUDA graph that captured an NCCL kernel blocked for minutes on the
    2-GPU box (the communicator waits for work the graph still owns); the process is ending anyway."""
    sys.stdout.flush()
    sys.stderr.flush()
    if world > 1:
        torch.cuda.synchronize()
        try:
            dist.barrier()
        except Exception:          # noqa: BLE001
            pass
        os._exit(0)


def start_watchdog(seconds):
    """A bench run that has not finished after `seconds` is stuck: exit loudly instead of holding the GPU box."""
    def bark():
        sys.stderr.write('bench.py: watchdog fired after %d s, exiting\n' % seconds)
        sys.stderr.flush()
        os._exit(3)
    timer = threading.Timer(seconds, bark)
    timer.daemon = True
    timer.start()


def launches_per_step(step, lib):
    """Kernels of ours enqueued by one step (counted on an eager call; a graph replay launches the same set)."""
    before = lib.launch_count
    step()
    return lib.launch_count - before


def cpu_reference_step(tp, torch, z, target, mask, reg):
    """One fwd+bwd of the reference's head on CPU tensors (oracle/torch_port.py restates it op for op)."""
    zz = z.detach().requires_grad_(True)
    loss, _, _, _ = tp.head_loss(zz, target, mask, reg, 1.0, 1.0)
    loss.backward()
    return loss.item()


def run_cpu_baseline(reg, budget_s=12.0, max_iters=10):
    import torch
    from oracle import torch_port as tp
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    b, c, h, w = CPU_SAMPLE
    g = torch.Generator().manual_seed(0)
    z = torch.randn(b, c, h, w, generator=g)
    target = torch.rand(b, c, 2, generator=g) * 1.6 - 0.8
    mask = (torch.rand(b, c, generator=g) > 0.1).float()
    for _ in range(2):
        cpu_reference_step(tp, torch, z, target, mask, reg)
    times = []
    t_start = time.perf_counter()
    while len(times) < max_iters and (time.perf_counter() - t_start < budget_s or len(times) < 3):
        t0 = time.perf_counter()
        cpu_reference_step(tp, torch, z, target, mask, reg)
        times.append(time.perf_counter() - t0)
    times.sort()
    med = times[len(times) // 2]
    return {'value': b * c / med, 'unit': UNIT, 'cores': torch.get_num_threads(), 'kind': 'port',
            'sample': '%d iterations of fwd+bwd on %dx%dx%dx%d fp32 (BASELINE cfg 1 shape), euclid + %s + mask, '
                      'oracle/torch_port.py on torch CPU, median %.1f ms/iter' % (len(times), b, c, h, w, reg,
                                                                                  med * 1e3)}


# ----------------------------------------------------------------------------------------------- reference arm
def main_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return 0
    import torch
    from oracle import torch_port as tp
    bsz, joints, h, w, reg, dtype = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    b, c = CPU_SAMPLE[0], CPU_SAMPLE[1]
    if h * w > 64 * 64:
        b = max(1, b * 64 * 64 // (h * w))
    g = torch.Generator().manual_seed(0)
    z = torch.randn(b, c, h, w, generator=g)
    target = torch.rand(b, c, 2, generator=g) * 1.6 - 0.8
    mask = (torch.rand(b, c, generator=g) > 0.1).float()
    for _ in range(args.warmup):
        cpu_reference_step(tp, torch, z, target, mask, reg)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_reference_step(tp, torch, z, target, mask, reg)
    dt = time.perf_counter() - t0
    value = b * c * args.steps / dt
    sample = ('each step = fwd+bwd on a %dx%dx%dx%d fp32 sample of the workload (bounded so the run ends within '
              'minutes), oracle/torch_port.py restating src/dsnt/nn.py + model.py on torch CPU' % (b, c, h, w))
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': dt / args.steps * 1e3,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': args.workload, 'heatmap': [h, w], 'joints': joints, 'reg': reg,
                   'sample_heatmaps_per_step': b * c, 'device': 'host CPU'},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': torch.get_num_threads(), 'kind': 'port',
                         'sample': sample},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), file=RESULT_OUT)
    RESULT_OUT.flush()
    return 0


# ----------------------------------------------------------------------------------------------- our arm
def main_ours(args):
    import torch
    import torch.distributed as dist
    import dsnt_pose2d_b200 as dp
    from dsnt_pose2d_b200 import _lib
    from dsnt_pose2d_b200.parallel import init_from_env

    world_env = int(os.environ.get('WORLD_SIZE', '1'))
    if args.gpus > 1 and world_env == 1:
        # launched as plain `python bench.py --gpus N`: re-exec under torchrun, one rank per GPU
        cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(args.gpus),
               '--master-addr', '127.0.0.1', '--master-port', str(29500 + os.getpid() % 2000), os.path.abspath(__file__)]
        return subprocess.call(cmd + sys.argv[1:])
    rank, local, world = init_from_env('nccl')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    group = dist.group.WORLD if world > 1 else None

    bsz, joints, h, w, reg, dtype = WORKLOADS[args.workload]
    tdt = torch.float32 if dtype == 'f32' else torch.bfloat16
    esize = 4 if dtype == 'f32' else 2
    n_local = bsz * joints
    torch.manual_seed(1234 + rank)
    z = torch.randn(bsz, joints, h, w, device=dev).to(tdt).requires_grad_(True)
    target = torch.rand(bsz, joints, 2, device=dev) * 1.6 - 0.8
    mask = (torch.rand(bsz, joints, device=dev) > 0.1).float()

    from dsnt_pose2d_b200.head import step_supported
    from dsnt_pose2d_b200.parallel import PeerExchange
    one_pass = args.path == 'one-pass' and step_supported(z, reg)
    exchange_note = ('inside the kernels over NVLink peer memory (no collective launch)'
                     if PeerExchange.get(group, dev) is not None else 'by a 3-float NCCL all-reduce')

    def step():
        z.grad = None
        out = dp.dsnt_head(z, target, mask, reg=reg, hm_sigma=1.0, reg_coeff=1.0, group=group, one_pass=one_pass)
        out.loss.backward()
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- resident-input timing (pass 1: the number of record, nothing but the steps in the region)
    graph = None
    graph_note = 'eager autograd'
    if args.graph:
        # the step has no host synchronisation, so the whole fwd+bwd (and, sharded, the 3-float all-reduce) is
        # captured once and replayed; if capture is refused the eager calls are timed instead (and said so)
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(3):
                    step()
            torch.cuda.current_stream().wait_stream(side)
            z.grad = None
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                step()
            graph_note = 'cuda-graph replay of the public-API step'
        except Exception as e:          # noqa: BLE001
            graph = None
            graph_note = 'eager autograd (graph capture failed: %s)' % (str(e).splitlines()[0][:120],)
            torch.cuda.synchronize()
    run_step = graph.replay if graph is not None else step
    for _ in range(max(args.warmup, 3)):
        run_step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    launches0 = _lib.launch_count
    barrier()
    t_wall0 = time.time()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        run_step()
    ev1.record()
    barrier()
    t_wall1 = time.time()
    launches = (_lib.launch_count - launches0) if graph is None else args.steps * launches_per_step(step, _lib)
    elapsed_ms = ev0.elapsed_time(ev1)
    t = torch.tensor([elapsed_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms = t.item()

    # ---------------- pass 2 (same steps, still inside the clock-sampled window): CUDA events around every launch of
    # our kernels, on the stream they are enqueued on, for the per-kernel roofline
    _lib.event_log = {'dsnt_head_fwd': [], 'dsnt_head_bwd': [], 'dsnt_head_step': [], 'dsnt_head_step_fused': [],
                      'dsnt_head_step_fused_peer': [],
                      'dsnt_finish_loss': [], 'dsnt_mask_count': [], 'dsnt_scale_unless_one': [],
                      'dsnt_finish_loss_peer': [], 'dsnt_mask_count_peer': []}
    for _ in range(args.steps):
        step()
    barrier()
    t_wall1 = time.time()
    logs, _lib.event_log = _lib.event_log, None
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    kernel_ms = {k: (sum(a.elapsed_time(b) for a, b in v) / len(v) if v else None) for k, v in logs.items()}

    # ---------------- end-to-end timing: host buffers in, loss + coords out, copies inside the timed region
    e2e = None
    if not args.no_e2e:
        zh = torch.empty(bsz, joints, h, w, dtype=tdt).pin_memory()
        zh.copy_(z.detach())
        th, mh = target.cpu().pin_memory(), mask.cpu().pin_memory()
        loss_h = torch.empty((), dtype=torch.float32).pin_memory()
        coords_h = torch.empty(bsz, joints, 2, dtype=torch.float32).pin_memory()
        zd = torch.empty_like(z.detach())
        td, md = torch.empty_like(target), torch.empty_like(mask)

        def e2e_step():
            zd.copy_(zh, non_blocking=True)
            td.copy_(th, non_blocking=True)
            md.copy_(mh, non_blocking=True)
            zin = zd.detach().requires_grad_(True)
            out = dp.dsnt_head(zin, td, md, reg=reg, hm_sigma=1.0, reg_coeff=1.0, group=group, one_pass=one_pass)
            out.loss.backward()
            loss_h.copy_(out.loss.detach(), non_blocking=True)
            coords_h.copy_(out.coords.detach(), non_blocking=True)
            return zin.grad

        e2e_steps = max(3, min(args.steps, 20))
        for _ in range(3):
            e2e_step()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(e2e_steps):
            e2e_step()
        e1.record()
        barrier()
        te = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        h2d = zh.numel() * esize + th.numel() * 4 + mh.numel() * 4
        d2h = 4 + coords_h.numel() * 4
        e2e = {'value': n_local * world * e2e_steps / (te.item() * 1e-3), 'unit': UNIT,
               'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h, 'steps': e2e_steps,
               'ms_per_step': te.item() / e2e_steps,
               'note': 'pinned host Z/target/mask -> device, fused fwd+bwd, loss+coords -> pinned host; '
                       'dL/dZ stays on the device for the backbone backward'}

    if rank != 0:
        leave(world, dist, torch)
        return 0

    # ---------------- roofline of the dominant kernel (live CUDA-event durations from the timed region)
    peak, peak_how = measured_peak()
    hw = h * w
    alg = {'dsnt_head_fwd': n_local * (hw * esize + 56),        # read Z; target 8 r, coords 8 w, stats 32 w, terms 8 w
           'dsnt_head_bwd': n_local * (2 * hw * esize + 44),    # read Z, write dZ; stats 32 r, target 8 r, mask 4 r
           'dsnt_head_step': n_local * (2 * hw * esize + 68),   # read Z, write dZ; target 8 r, mask 4 r, coords/stats/terms 56 w
           'dsnt_head_step_fused': n_local * (2 * hw * esize + 60)}   # the same without the terms (8 w); the mask again from L2
    alg['dsnt_head_step_fused_peer'] = alg['dsnt_head_step_fused']     # + 2 x 16 bytes to every peer
    ran = [k for k in alg if kernel_ms.get(k)]
    if not ran:
        raise RuntimeError('bench.py: none of the timed entry points ran (%r)' % (sorted(kernel_ms),))
    dominant = max(ran, key=lambda k: kernel_ms[k])
    per_kernel = {}
    for k in ran:
        ach = alg[k] / (kernel_ms[k] * 1e-3) / 1e9
        per_kernel[k] = {'ms': kernel_ms[k], 'algorithmic_bytes': alg[k], 'achieved_gbs': ach, 'frac': ach / peak}
    small = {k: kernel_ms[k] for k in ('dsnt_finish_loss', 'dsnt_mask_count', 'dsnt_scale_unless_one',
                                       'dsnt_finish_loss_peer', 'dsnt_mask_count_peer') if kernel_ms.get(k)}
    step_bytes = sum(alg[k] for k in ran)
    step_gbs = step_bytes * world / (elapsed_ms / args.steps * 1e-3) / 1e9
    traffic = committed_traffic(args.workload)
    if isinstance(traffic, dict) and 'dsnt_head_step' in traffic:
        for same in ('dsnt_head_step_fused', 'dsnt_head_step_fused_peer'):      # the same kernel (head_step2_kernel)
            traffic.setdefault(same, traffic['dsnt_head_step'])
    roofline = {'bound': 'hbm', 'kernel': dominant, 'achieved': per_kernel[dominant]['achieved_gbs'], 'peak': peak,
                'unit': 'GB/s', 'frac': per_kernel[dominant]['frac'],
                'traffic': (traffic or {}).get(dominant) if isinstance(traffic, dict) else None,
                'peak_source': peak_how, 'kernels': per_kernel, 'small_kernels_ms': small,
                'step': {'algorithmic_bytes_per_gpu': step_bytes, 'achieved_gbs_per_gpu': step_gbs / world,
                         'frac': step_gbs / world / peak}}

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        cpu_baseline = run_cpu_baseline(reg)

    value = n_local * world * args.steps / (elapsed_ms * 1e-3)
    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
        'warmup': max(args.warmup, 3), 'ms_per_step': elapsed_ms / args.steps, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': dtype, 'data': 'synthetic',
        'config': {'workload': args.workload, 'heatmaps_per_gpu': n_local, 'batch_per_gpu': bsz, 'joints': joints,
                   'heatmap': [h, w], 'reg': reg, 'hm_sigma_px': 1.0, 'mask': True,
                   'parallelism': ('batch-sharded x%d; the three partial sums of masked_average are exchanged %s' % (
                       world, exchange_note)) if world > 1 else 'single GPU (batch shard = whole batch)',
                   'l2_policy': 'inputs larger than L2 (%.0f MiB of logits per step vs 126 MB L2)'
                                % (n_local * hw * esize / 2 ** 20),
                   'path': ('one-pass: %s (forward and dL/dZ while the heatmap is in shared memory, 2*H*W*sizeof '
                            'algorithmic bytes); backward only scales in place when d(loss) != 1' % (
                                'dsnt_head_step_fused_peer, ONE launch per rank: mask count, step, loss composition and '
                                'both exchanges between the ranks' if kernel_ms.get('dsnt_head_step_fused_peer') else
                                'dsnt_head_step_fused, ONE launch: mask count, step and loss composition'
                                if kernel_ms.get('dsnt_head_step_fused') else
                                'dsnt_mask_count + dsnt_head_step + dsnt_finish_loss')) if one_pass else
                           'two-kernel: dsnt_head_fwd + dsnt_finish_loss + dsnt_head_bwd (3*H*W*sizeof algorithmic bytes)',
                   'launch': graph_note},
        'roofline': roofline, 'cpu_baseline': cpu_baseline, 'clocks': clocks, 'e2e': e2e,
        'gpu_launches': launches,
    }
    print(json.dumps(line), file=RESULT_OUT)
    RESULT_OUT.flush()
    leave(world, dist, torch)
    return 0


def claim_stdout():
    """Keep file descriptor 1 for the ONE JSON line: libraries (NCCL prints its version banner to stdout) write to
    stderr from here on; returns the stream the result line is printed to."""
    sys.stdout.flush()
    keep = os.dup(1)
    os.dup2(2, 1)
    return os.fdopen(keep, 'w')


if __name__ == '__main__':
    a = parse_args()
    start_watchdog(900)
    RESULT_OUT = claim_stdout()
    sys.exit(main_reference(a) if a.impl == 'reference' else main_ours(a))
