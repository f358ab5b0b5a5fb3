#!/usr/bin/env python
"""bench.py -- DPC-RNN training-step throughput on B200 (BASELINE.json metric: clips/sec, device-timed).

    python bench.py --gpus 1 --steps K --warmup W                 # this repo's CUDA path
    torchrun --nproc-per-node N ... bench.py --gpus N ...         # one rank per GPU, NCCL grad all-reduce
    python bench.py --impl reference ...                          # the reference algorithm on host cores (oracle port)

One "step" = forward + fused NCE loss + backward + gradient all-reduce + Adam over one batch of
synthetic video ([B,8,3,5,img,img] fp32, B clips per GPU; BASELINE config 2: 2d3d-R18, 128^2, B=128).
Prints ONE JSON line (rank 0).  See DESIGN.md §Measurement for every field.
"""
import argparse
import contextlib
import io
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = 'clips/sec (device-timed) DPC-RNN 2d3d-R18 128^2 train step'


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=8)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--net', default='resnet18')
    ap.add_argument('--img_dim', type=int, default=128)
    ap.add_argument('--batch_size', type=int, default=128, help='clips per GPU')
    ap.add_argument('--pred_step', type=int, default=3)
    ap.add_argument('--no_cpu_baseline', action='store_true')
    ap.add_argument('--no_e2e', action='store_true')
    ap.add_argument('--no_stock', action='store_true', help='skip the stock PyTorch-CUDA leg')
    ap.add_argument('--no_pipeline', action='store_true', help='skip the on-device input pipeline leg')
    ap.add_argument('--stock_steps', type=int, default=4)
    return ap.parse_args()


def workload(a, n):
    return {'workload': 'DPC-RNN train step (fwd + NCE loss + bwd + grad all-reduce + Adam), 2d3d-%s, img %d, '
                        'num_seq 8, seq_len 5, pred_step %d, batch %d clips/GPU, synthetic video'
                        % (a.net.replace('resnet', 'R'), a.img_dim, a.pred_step, a.batch_size),
            'network': a.net, 'img_dim': a.img_dim, 'batch_per_gpu': a.batch_size,
            'global_batch': a.batch_size * n, 'parallelism': 'dp%d' % n,
            'l2': 'input batch %.2f GB/step per GPU > 126 MB L2 (no explicit flush needed)'
                  % (a.batch_size * 8 * 3 * 5 * a.img_dim ** 2 * 4 / 1e9)}


def cpu_threads():
    """host threads for the CPU arm: all cores up to 32 -- torch's CPU conv3d backward gets SLOWER beyond
    that on the 128-core GPU hosts (measured 14-94 s/step at 128 threads vs ~1.5 s at 8-32)"""
    return max(1, min(os.cpu_count() or 1, int(os.environ.get('DPC_CPU_THREADS', 32))))


def _out_dims(img, seq_len=5):
    """(T, H, W) of the layer3 feature map: stem /2, pool /2, layer2 /2, layer3 /2 spatially; layer3 halves T (ceil)"""
    h = img // 16
    return ((seq_len + 1) // 2, h, h)


def ncu_evidence():
    """per-kernel numbers of the committed `ncu --set full` captures (profiles/r2_ncu_summary.json): DRAM traffic per
    launch vs algorithmic bytes, tensor-pipe activity -- static evidence, not re-measured by this run"""
    p = os.path.join(ROOT, 'profiles', 'r2_ncu_summary.json')
    if not os.path.exists(p):
        return None
    keep = ('kernel', 'site', 'ncu_duration_ms', 'dram_bytes', 'traffic_over_algorithmic', 'tensor_pipe_active_pct',
            'achieved_tflops_algorithmic', 'frac_of_bf16_burst', 'achieved_gbs_algorithmic', 'frac_of_hbm')
    return {k: {kk: v[kk] for kk in keep if kk in v} for k, v in json.load(open(p)).items()}


def measured_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get('hbm_gbs', 6650.0), d.get('bf16_tflops_sustained', 1400.0), 'measured'
    return 6650.0, 1590.0, 'fallback'


# ---------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock / clock-event reasons of ONE GPU sampled every 100 ms during the timed region: in-process through NVML
    (pynvml; no child process, so every rank can watch its own GPU), else an `nvidia-smi -lms 200` child"""
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')
    BITS = (('hw_slowdown', 0x8), ('hw_thermal_slowdown', 0x40), ('sw_thermal_slowdown', 0x20), ('sw_power_cap', 0x4))

    def __init__(self, index):
        self.lines, self.proc, self.nv = [], None, None
        self._stop = threading.Event()
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            uuid = 'GPU-' + str(torch.cuda.get_device_properties(index).uuid)
            self.h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode() if hasattr(uuid, 'encode') else uuid)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.nv = pynvml
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nv = None
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '200'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _poll(self):
        nv = self.nv
        get_reasons = getattr(nv, 'nvmlDeviceGetCurrentClocksEventReasons', None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self._stop.is_set():
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                r = int(get_reasons(self.h))
                self.lines.append(', '.join([str(sm), str(self.mx)] + ['Active' if r & bit else 'Not Active' for _, bit in self.BITS]))
            except Exception:
                pass
            self._stop.wait(0.1)

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.nv is not None:
            self._stop.set()
            self.t.join(timeout=2)
        elif self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
        else:
            return None
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for ln in self.lines:
            f = [x.strip() for x in ln.split(',')]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nme, v in zip(names, f[2:6]):
                if v == 'Active':
                    reasons.add(nme)
        if not sm:
            return None
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2], 'sm_max_mhz': max(mx), 'reasons': sorted(reasons), 'samples': len(sm)}


# ---------------------------------------------------------------------------------------------
def ref_sample_batch(a):
    """clips per CPU step of the reference arm: BASELINE config 1 (4 clips at R18 / 128^2, ~1 s per step on 32 host
    threads), scaled down for the heavier configs so that `--steps K --warmup W` still ends within a few minutes"""
    cost = (a.img_dim / 128.0) ** 2 * (7.1 if a.net == 'resnet34' else 1.0 if a.net == 'resnet18' else 4.0)
    return max(1, min(4, int(round(4 / cost))))


def cpu_reference_step(a):
    """-> (step fn, kind, impl description) for the CPU arms: the unmodified reference modules staged in baseline/_ref
    (oracle/make_ref.py) driven by main.py's own loop lines, else the oracle port"""
    import torch
    from oracle import ref_harness as H
    step = H.reference_step(a.net, a.img_dim, a.pred_step, torch.device('cpu'))
    if step is not None:
        return step, 'reference', 'unmodified reference modules (baseline/_ref) + main.py loop lines, torch CPU fp32'
    return H.port_step(a.net, a.img_dim, a.pred_step, torch.device('cpu')), 'port', 'oracle port (oracle/dpc_oracle.py), torch CPU fp32'


def time_cpu_steps(a, steps, warmup):
    """-> (seconds per step, clips per step, kind, impl, threads): fwd + CE + top-k + bwd + Adam on the host cores"""
    import torch
    torch.set_num_threads(cpu_threads())
    bs = ref_sample_batch(a)
    step, kind, impl = cpu_reference_step(a)
    g = torch.Generator().manual_seed(1234)
    x = torch.randn(bs, 8, 3, 5, a.img_dim, a.img_dim, generator=g)
    for _ in range(warmup):
        step(x)
    t0 = time.perf_counter()
    for _ in range(steps):
        step(x)
    return (time.perf_counter() - t0) / steps, bs, kind, impl, torch.get_num_threads()


def run_reference(a, rank, world):
    """`--impl reference`: the reference's own CPU implementation of the path on this box's host cores.  Runs EXACTLY
    the requested W warm-up + K timed steps; each step is a bounded sample of the workload (`ref_sample_batch` clips
    instead of batch_per_gpu -- stated in config.reference_sample), so the line's steps / ms_per_step are the real ones."""
    if rank != 0:
        return
    t, bs, kind, impl, threads = time_cpu_steps(a, a.steps, a.warmup)
    v = bs / t
    cfg = workload(a, world)
    cfg['reference_sample'] = {'clips_per_step': bs, 'note': 'each timed step is fwd + CE + top-k + bwd + Adam over %d clips '
                               '(not batch_per_gpu) on %d host threads; value = clips_per_step / seconds per step' % (bs, threads)}
    line = {'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': 'clips/s', 'n_gpus': a.gpus,
            'steps': a.steps, 'warmup': a.warmup, 'ms_per_step': t * 1e3, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'config': cfg,
            'cpu_baseline': {'value': v, 'unit': 'clips/s', 'cores': threads, 'kind': kind,
                             'sample': '%d timed train steps (fwd+CE+top-k+bwd+Adam) of %d clips each, %s img %d, %s'
                                       % (a.steps, bs, a.net, a.img_dim, impl)},
            'e2e': {'value': v, 'unit': 'clips/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line))


def run_stock_cuda(a, world, steps, warmup=2):
    """The reference's real GPU path ("stock PyTorch-CUDA", north_star's 10x denominator): unmodified reference modules
    under nn.DataParallel over `world` GPUs (main.py:65), cudnn.benchmark (main.py:25), torch's default TF32 convs,
    Adam, the driver's loss / top-k lines -- global batch world x batch_per_gpu on device 0, device-timed.
    Runs in THIS process (rank 0) after our arm has released its memory."""
    import torch
    from oracle import ref_harness as H
    torch.backends.cudnn.benchmark = True
    dev = torch.device('cuda', 0)
    ids = list(range(world))
    step = H.reference_step(a.net, a.img_dim, a.pred_step, dev, device_ids=ids)
    impl = 'unmodified reference modules (baseline/_ref) under nn.DataParallel(%d), cudnn.benchmark, TF32 convs' % world
    if step is None:
        if world > 1:
            return {'unavailable': 'baseline/_ref not staged and the oracle port has no DataParallel wrapper'}
        step = H.port_step(a.net, a.img_dim, a.pred_step, dev)
        impl = 'oracle port on CUDA (reference op sequence), cudnn.benchmark, TF32 convs'
    B = a.batch_size * world
    g = torch.Generator().manual_seed(4321)
    x = torch.randn(B, 8, 3, 5, a.img_dim, a.img_dim, generator=g).to(dev)
    for _ in range(warmup):
        step(x)
    for d in ids:
        torch.cuda.synchronize(d)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(steps):
        loss = step(x)
        _ = loss.item()                                             # main.py:224: losses.update(loss.item(), B)
    e1.record()
    for d in ids:
        torch.cuda.synchronize(d)
    wall = (time.perf_counter() - t0) * 1e3 / steps
    ms = e0.elapsed_time(e1) / steps
    out = {'clips_s': B / (ms / 1e3), 'ms_per_step': ms, 'wall_ms_per_step': wall, 'steps': steps, 'warmup': warmup,
           'global_batch': B, 'n_gpus': world, 'impl': impl, 'loss': float(loss.detach()),
           'peak_mem_gb': torch.cuda.max_memory_allocated(dev) / 1e9}
    if world > 1:
        # nn.DataParallel rebuilds the 6-D mask with Python loops in every replica on every step (model_3d.py:86-96: the
        # replicas are throw-away copies, so `self.mask` never survives) and re-broadcasts the parameters: the reference's
        # own multi-GPU path is slower than its single-GPU path.  Also report the single-GPU rate, i.e. what an ideal
        # data-parallel launcher would multiply by `world`.
        del step, x
        torch.cuda.empty_cache()
        one = run_stock_cuda(a, 1, steps, warmup)
        out['single_gpu_clips_s'] = one.get('clips_s')
        out['ideal_ddp_clips_s'] = one['clips_s'] * world if one.get('clips_s') else None
    return out


def conv_family_flops(network, NB, T, H, W):
    """algorithmic FLOPs (2*MAC) of the im2col-GEMM conv family per step: fwd + dgrad + wgrad"""
    from dpc_b200.arch import backbone_spec
    from dpc_b200.engine import _out_extent as e
    Hc, Wc = e(e(H, 7, 2, 3), 3, 2, 1), e(e(W, 7, 2, 3), 3, 2, 1)
    Tc = T
    total = 0
    for b in backbone_spec(network):
        k = (3, 3, 3) if b['is3d'] else (1, 3, 3)
        s = (b['stride'],) * 3 if b['is3d'] else (1, b['stride'], b['stride'])
        p = (1, 1, 1) if b['is3d'] else (0, 1, 1)
        To, Ho, Wo = e(Tc, k[0], s[0], p[0]), e(Hc, k[1], s[1], p[1]), e(Wc, k[2], s[2], p[2])
        rows = NB * To * Ho * Wo
        taps = k[0] * k[1] * k[2]
        if b['block'] == 'bottleneck':                                     # 1x1x1 -> k (strided) -> 1x1x1
            total += 2 * NB * Tc * Hc * Wc * b['planes'] * b['inplanes']
            total += 2 * rows * b['planes'] * taps * b['planes']
            total += 2 * rows * b['outplanes'] * b['planes']
        else:
            total += 2 * rows * b['planes'] * taps * b['inplanes']           # conv1
            total += 2 * rows * b['planes'] * taps * b['planes']             # conv2
        if b['downsample']:
            total += 2 * rows * b['outplanes'] * b['inplanes']
        Tc, Hc, Wc = To, Ho, Wo
    return 3 * total


def run_b200(a, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    import dpc_b200
    from dpc_b200 import engine
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    gloo = None
    if world > 1:
        import datetime
        dist.init_process_group('nccl', device_id=dev)
        gloo = dist.new_group(backend='gloo', timeout=datetime.timedelta(minutes=30))   # host-side waits (stock leg)
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        model = dpc_b200.DPC_RNN(a.img_dim, num_seq=8, seq_len=5, pred_step=a.pred_step, network=a.net)
    model = model.to(dev).train()                                    # dropout on, as in training
    crit = dpc_b200.NCECriterion()
    trainer = dpc_b200.FlatTrainer(model, lr=1e-3, weight_decay=1e-5)
    B = a.batch_size
    shape = (B, 8, 3, 5, a.img_dim, a.img_dim)
    g = torch.Generator().manual_seed(1234 + rank)
    host = [torch.randn(shape, generator=g).pin_memory() for _ in range(2)]
    x_dev = host[0].to(dev)
    L = dpc_b200.lib()

    def step(x):
        trainer.zero_grad()
        score, _ = model(x)
        loss = crit(score)
        loss.backward()
        trainer.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t[0])
        return ms

    # ---- device-resident ("value") -------------------------------------------------------------
    for _ in range(a.warmup):
        step(x_dev)
    barrier()
    sampler = ClockSampler(local_rank)                                # every rank samples its own GPU
    trainer.allreduce_events = [] if world > 1 else None
    allreduce_impl = None if world == 1 else ('dpc_flat_allreduce (library-owned NCCL communicator, compute stream)'
                                              if trainer.comm is not None else 'torch.distributed.all_reduce')
    n0 = L.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        loss = step(x_dev)
    e1.record()
    barrier()
    ms_local = e0.elapsed_time(e1)
    ms_total = max_over_ranks(ms_local)
    rank_ms = rank_diag = None
    launches = L.launch_count() - n0
    clocks = sampler.stop()
    if world > 1:
        # per rank: device time per step, time inside the all-reduce bracket (transfer + waiting for the slowest rank: the
        # rank that arrives last waits least) and its own GPU's median SM clock -- names the weak-scaling limiter
        ar = sum(x.elapsed_time(y) for x, y in trainer.allreduce_events) / max(1, len(trainer.allreduce_events))
        trainer.allreduce_events = None
        mine = {'rank': rank, 'ms_per_step': round(ms_local / a.steps, 3), 'allreduce_bracket_ms': round(ar, 3),
                'compute_ms': round(ms_local / a.steps - ar, 3), 'sm_mhz': (clocks or {}).get('sm_mhz'),
                'reasons': (clocks or {}).get('reasons')}
        allr = [None] * world
        dist.all_gather_object(allr, mine, group=gloo)
        rank_ms = [r['ms_per_step'] for r in allr]
        rank_diag = allr
    ms_step = ms_total / a.steps
    value = world * B / (ms_step / 1e3)
    last_loss = float(loss.detach())

    # ---- end to end: pinned host input -> H2D (prefetched on a copy stream) -> step -> loss.item() ----
    e2e = None
    if not a.no_e2e:
        copy_stream = torch.cuda.Stream()
        bufs = [torch.empty(shape, device=dev) for _ in range(2)]
        ready = [torch.cuda.Event() for _ in range(2)]
        free = [torch.cuda.Event() for _ in range(2)]

        def prefetch(i):
            j = i % 2
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(free[j])
                bufs[j].copy_(host[j], non_blocking=True)
                ready[j].record(copy_stream)

        for j in range(2):
            free[j].record()
        barrier()
        nwarm = 1
        prefetch(0)
        t0 = None
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for i in range(nwarm + a.steps):
            if i == nwarm:
                barrier()
                ev0.record()
                t0 = time.perf_counter()
            prefetch(i + 1)
            torch.cuda.current_stream().wait_event(ready[i % 2])
            loss = step(bufs[i % 2])
            free[i % 2].record()
            _ = loss.item()                                          # the D2H read of the step's result
        ev1.record()
        barrier()
        wall = (time.perf_counter() - t0) * 1e3
        ms_e2e = max_over_ranks(max(ev0.elapsed_time(ev1), 0.0)) / a.steps
        e2e = {'value': world * B / (ms_e2e / 1e3), 'unit': 'clips/s',
               'h2d_bytes_per_step': host[0].numel() * 4, 'd2h_bytes_per_step': 4,
               'ms_per_step': ms_e2e, 'wall_ms_per_step': wall / a.steps}

    # ---- on-device input pipeline: decoded uint8 frames -> H2D -> augment kernel -> step ----------
    # the reference's CPU transform chain (utils/augmentation.py, main.py:125-133) replaced by dpc_b200.augmentation; the
    # random decisions of batch i+1 are drawn on the host while the GPU runs step i
    pipe = None
    if not a.no_pipeline and not a.no_e2e:
        import random as _random
        import numpy as _np
        from dpc_b200 import augmentation as aug
        fw, fh = (200, 150) if a.img_dim <= 128 else (340, 256)      # k400-style frames: short side 150 (256 for 224^2)
        tr = aug.k400_transform(a.img_dim)
        _random.seed(1234 + rank)
        _np.random.seed(1234 + rank)
        hostf = [torch.randint(0, 256, (B, 40, fh, fw, 3), dtype=torch.uint8).pin_memory() for _ in range(2)]
        fbufs = [torch.empty((B, 40, fh, fw, 3), dtype=torch.uint8, device=dev) for _ in range(2)]
        blk = bufs[0]                                                # reuse the e2e leg's block buffer
        fready = [torch.cuda.Event() for _ in range(2)]
        ffree = [torch.cuda.Event() for _ in range(2)]

        def fprefetch(i):
            j = i % 2
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(ffree[j])
                fbufs[j].copy_(hostf[j], non_blocking=True)
                fready[j].record(copy_stream)

        for j in range(2):
            ffree[j].record()
        t_plan = time.perf_counter()
        plans = [tr.plan(40, fw, fh) for _ in range(B)]
        t_plan = (time.perf_counter() - t_plan) * 1e3
        t_pack = time.perf_counter()
        prep = tr.prepare(plans, dev)
        t_pack = (time.perf_counter() - t_pack) * 1e3
        side_bytes = int(prep['tables'].numel() + prep['frame_params'].numel()) * 4     # tables + per-frame decisions per batch
        fprefetch(0)
        torch.cuda.current_stream().wait_event(fready[0])
        ka, kb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        tr(fbufs[0], 8, 5, prepared=prep, out=blk)                   # warm-up
        ka.record()
        for _ in range(5):
            tr(fbufs[0], 8, 5, prepared=prep, out=blk)
        kb.record()
        ffree[0].record()
        barrier()
        aug_ms = ka.elapsed_time(kb) / 5
        fprefetch(0)
        pv0, pv1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for i in range(1 + a.steps):
            if i == 1:
                barrier()
                pv0.record()
            fprefetch(i + 1)
            torch.cuda.current_stream().wait_event(fready[i % 2])
            tr(fbufs[i % 2], 8, 5, prepared=prep, out=blk)
            ffree[i % 2].record()
            loss = step(blk)
            prep = tr.prepare([tr.plan(40, fw, fh) for _ in range(B)], dev)      # next batch's draws + tables, under the running step
            _ = loss.item()
        pv1.record()
        barrier()
        ms_p = max_over_ranks(pv0.elapsed_time(pv1)) / a.steps
        pipe = {'frames': 'uint8 [%d, 40, %d, %d, 3] per GPU (decoded RGB, synthetic)' % (B, fh, fw),
                'recipe': 'k400 (main.py:125-133): RandomSizedCrop, RandomHorizontalFlip, RandomGray, ColorJitter, ToTensor, Normalize',
                'augment_ms_per_batch': aug_ms, 'augment_clips_s': B / (aug_ms / 1e3),
                'augment_gbs': (hostf[0].numel() + blk.numel() * 4) / (aug_ms / 1e3) / 1e9,
                'host_draw_ms_per_batch': t_plan, 'host_tables_ms_per_batch': t_pack,
                'e2e_uint8': {'value': world * B / (ms_p / 1e3), 'unit': 'clips/s', 'ms_per_step': ms_p,
                              'h2d_bytes_per_step': hostf[0].numel() + side_bytes,
                              'd2h_bytes_per_step': 4}}
        del hostf, fbufs

    # ---- roofline of the dominant kernel family (one extra, instrumented step) -------------------
    roof = None
    fam = None
    sites = allreduce_ms = None
    if rank == 0:
        # the instrumented step runs on ONE stream (no wgrad side stream), i.e. with a different allocation pattern: run
        # it twice and keep the second, so that the caching allocator's cudaMalloc stalls (host-side, but inside the
        # event brackets) do not pollute the per-family times
        for _rep in range(2):
            timer = engine.EventTimer()
            engine.set_timer(timer)
            step(x_dev)
            engine.set_timer(None)
        tot = timer.totals()
        allreduce_ms = tot.pop('allreduce', (0, None))[1]
        fam = {k: {'calls': c, 'ms': round(t, 3)} for k, (c, t) in sorted(tot.items(), key=lambda kv: -kv[1][1])}
        conv_ms = sum(t for k, (c, t) in tot.items() if k in ('conv_fwd', 'conv_dgrad', 'conv_wgrad'))
        # per call site, in launch order: [family, (Ci, Co, taps, rows_out, stride-1?), ms]
        sites = [[tg, list(d), round(x.elapsed_time(y), 4)] for tg, x, y, d in timer.records
                 if tg in ('conv_fwd', 'conv_dgrad', 'conv_wgrad') and d is not None]
        flops = conv_family_flops(a.net, B * 8, 5, a.img_dim, a.img_dim)
        hbm, tf, how = measured_peaks()
        ach = flops / (conv_ms / 1e3) / 1e12
        family = {'kernel': 'conv3d implicit-GEMM family (fwd+dgrad+wgrad, all layers)', 'achieved': ach, 'frac': ach / tf,
                  'algorithmic_flops_per_step': flops, 'family_ms_per_step': conv_ms, 'share_of_step': conv_ms / ms_step,
                  'executed_tflops': 3 * ach}
        # the dominant kernel: conv_tc_kernel at the layer3 3x3x3 256->256 stride-1 sites (3 forward launches per step),
        # each launch timed with CUDA events inside this instrumented step
        nb = B * 8
        d3 = _out_dims(a.img_dim)
        rows3 = nb * d3[0] * d3[1] * d3[2]
        l3 = [ms for d, ms in timer.calls('conv_fwd') if d == (256, 256, 27, rows3, True)]
        ev = ncu_evidence() or {}
        if l3:
            ms3 = sum(l3) / len(l3)
            fl3 = 2.0 * rows3 * 256 * 256 * 27
            a3 = fl3 / (ms3 / 1e3) / 1e12
            roof = {'kernel': 'conv_tc_kernel: layer3 256->256 3x3x3 stride-1 conv forward (+ fused BatchNorm statistics)',
                    'bound': 'tensor', 'achieved': a3, 'peak': tf, 'unit': 'TFLOP/s', 'frac': a3 / tf,
                    'traffic': (ev.get('conv_l3') or {}).get('dram_bytes'), 'launches_timed': len(l3), 'ms_per_launch': ms3,
                    'algorithmic_flops_per_launch': fl3, 'algorithmic_bytes_per_launch': 2 * rows3 * 256 * 4 + 27 * 256 * 256 * 4,
                    'peak_source': how + ' bf16_tflops_sustained (kernel timed inside the step)', 'mma_passes': 3,
                    'executed_tflops': 3 * a3, 'family': family, 'ncu': ev}
            # the score matmul (dpc/model_3d.py:79-83) is bound by its fp32 output write: M^2 * 4 bytes (+ the operands)
            sc = [ms for _, ms in timer.calls('score_fwd')]
            if sc:
                Msc = B * a.pred_step * ((a.img_dim + 31) // 32) ** 2
                by = Msc * Msc * 4.0 + 2 * Msc * 256 * 4.0
                gbs = by / (sc[0] / 1e3) / 1e9
                roof['score_matmul'] = {'kernel': 'score matmul forward [M,256] x [256,M] -> fp32 [M,M], fp16-pair operands',
                                        'bound': 'hbm', 'M': Msc, 'ms': sc[0], 'algorithmic_bytes': by, 'achieved': gbs, 'peak': hbm,
                                        'unit': 'GB/s', 'frac': gbs / hbm, 'tflops_algorithmic': 2.0 * Msc * Msc * 256 / (sc[0] / 1e3) / 1e12}
        else:
            roof = dict(family, bound='tensor', peak=tf, unit='TFLOP/s', traffic=None, mma_passes=3,
                        peak_source=how + ' bf16_tflops_sustained', ncu=ev)
    else:
        for _rep in range(2):
            step(x_dev)                                              # keep the collective count equal
    barrier()

    # ---- stock PyTorch-CUDA leg (the reference's own GPU path), after our arm released its memory ----------
    stock = None
    if not a.no_stock:
        import gc
        trainer.close()
        del model, trainer, crit, x_dev, step
        if not a.no_e2e:
            del bufs
        gc.collect()
        torch.cuda.empty_cache()
        barrier()
        if rank == 0:
            try:
                # DataParallel steps cost seconds each (Python mask loops in every replica, serialised by the GIL): keep the
                # multi-GPU leg short so that the whole line stays within minutes at N = 8
                stock = run_stock_cuda(a, world, steps=max(1, min(a.stock_steps if world == 1 else 2, a.steps)),
                                       warmup=2 if world == 1 else 1)
            except Exception as exc:                                 # report, never hide: the leg is evidence, not product
                stock = {'error': '%s: %s' % (type(exc).__name__, str(exc)[:300])}
            torch.cuda.empty_cache()
        if world > 1:
            dist.barrier(group=gloo)                                 # host-side wait: no kernel spinning on the idle GPUs

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        t, bs, kind, impl, threads = time_cpu_steps(a, steps=8, warmup=1)
        cpu = {'value': bs / t, 'unit': 'clips/s', 'cores': threads, 'kind': kind,
               'sample': '8 train steps (fwd+CE+top-k+bwd+Adam) of %d clips each (BASELINE config 1 batch), %s' % (bs, impl)}

    if rank == 0:
        line = {'metric': METRIC, 'value': value, 'unit': 'clips/s', 'n_gpus': world, 'steps': a.steps,
                'warmup': a.warmup, 'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'weak',
                'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'config': workload(a, world),
                'clocks': clocks, 'e2e': e2e, 'gpu_launches': int(launches), 'roofline': roof,
                'cpu_baseline': cpu, 'kernel_families_ms': fam, 'loss': last_loss,
                'allreduce_impl': allreduce_impl, 'allreduce_ms': allreduce_ms, 'rank_ms_per_step': rank_ms, 'rank_diagnostics': rank_diag, 'conv_sites_ms': sites,
                'input_pipeline': pipe,
                'stock_cuda': stock,
                'vs_stock_cuda': (value / stock['clips_s']) if stock and stock.get('clips_s') else None,
                'vs_stock_cuda_ideal_ddp': (value / stock['ideal_ddp_clips_s']) if stock and stock.get('ideal_ddp_clips_s') else None}
        print(json.dumps(line))
    if world > 1:
        if a.no_stock:
            trainer.close()
        dist.destroy_process_group()


def main():
    a = parse()
    rank = int(os.environ.get('RANK', 0))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    if a.impl == 'reference':
        return run_reference(a, rank, world)
    if world != a.gpus:
        if a.gpus > 1 and world == 1:
            sys.exit('bench.py --gpus %d must be launched with torchrun --nproc-per-node %d' % (a.gpus, a.gpus))
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    run_b200(a, rank, local_rank, world)


if __name__ == '__main__':
    main()
