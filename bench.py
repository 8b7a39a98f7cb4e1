#!/usr/bin/env python
"""Throughput benchmark of the batched MultiAgentTracking step (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm (oracle port, all host threads)

A "step" is one pass of the fused step kernel over one batch of environments
(MATE-4v8-9, 65 536 envs per GPU, random joint actions, auto-reset on).  Multi-GPU runs
are launched by ``torch.distributed.run`` (one rank per GPU): environments are independent,
so the batch is sharded with no step-path collective (weak scaling: 65 536 envs per rank);
timing is CUDA events per rank, MAX over ranks.  Rank 0 prints ONE JSON line.
"""

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = 'env-steps/sec, MATE-4v8-9 batched, 1/2/4/8 B200 vs host-CPU reference'
UNIT = 'env-steps/s'


def algorithmic_bytes(nc, nt, no):
    """in + out + state bytes per env-step (fp32 I/O)."""
    dc = 22 + 5 * nt + 4 * no + 7 * nc
    dt = 27 + 7 * nc + 4 * no + 5 * nt
    return 8 * (nc + nt) + 4 * (nc * dc + nt * dt) + 9 + 8 * (2 * nc + 4 * nt + 10) + 4 * (2 * nc + 3 * no + 1)


def measured_peak_gbs():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    try:
        with open(path, encoding='UTF-8') as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    except (OSError, KeyError, ValueError):
        return 6650.0, 'fallback (B200_PROFILING.md)'


def ncu_traffic(workload):
    """dram read+write bytes per launch from the committed ncu --set full capture, with the hash of the kernel
    sources the capture was taken on: {'bytes': ..., 'kernel_sources_sha16': ..., 'capture': ...} or {}."""
    path = os.path.join(ROOT, 'profiles', 'traffic.json')
    try:
        with open(path, encoding='UTF-8') as f:
            entry = json.load(f).get(workload)
    except (OSError, ValueError):
        return {}
    if isinstance(entry, dict):
        return entry
    return {'bytes': entry} if entry else {}


class ClockSampler:
    """Samples SM clocks and throttle reasons while the timed region runs.

    NVML is queried in-process (nvidia_ml_py): forking ``nvidia-smi`` from a thread holds the GIL of the thread
    that launches the kernels, which showed up as +0.1 ms on a 2.9 ms timed region (profiles/r2h_summary.md).
    ``nvidia-smi`` is only the fallback when the NVML binding is missing."""

    QUERY = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')
    NAMES = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']

    def __init__(self, index, period=0.002):
        self.index = index
        self.period = period
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self.source = 'nvml'
        self._stop = threading.Event()
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._nvml = None
        self._handle = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self._nvml = pynvml
            # NVML enumerates physical devices; honour CUDA_VISIBLE_DEVICES when it lists indices
            visible = os.environ.get('CUDA_VISIBLE_DEVICES', '')
            ids = [x for x in visible.split(',') if x.strip().isdigit()]
            physical = int(ids[index]) if index < len(ids) else index
            self._handle = pynvml.nvmlDeviceGetHandleByIndex(physical)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._handle, pynvml.NVML_CLOCK_SM))
        except Exception:  # pylint: disable=broad-except
            self._nvml = None
            self.source = 'nvidia-smi'

    def _sample_nvml(self):
        nv = self._nvml
        self.samples.append(float(nv.nvmlDeviceGetClockInfo(self._handle, nv.NVML_CLOCK_SM)))
        reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(self._handle) if hasattr(nv, 'nvmlDeviceGetCurrentClocksEventReasons') \
            else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._handle)
        table = {
            'hw_slowdown': getattr(nv, 'nvmlClocksThrottleReasonHwSlowdown', 0x8),
            'hw_thermal_slowdown': getattr(nv, 'nvmlClocksThrottleReasonHwThermalSlowdown', 0x40),
            'sw_thermal_slowdown': getattr(nv, 'nvmlClocksThrottleReasonSwThermalSlowdown', 0x20),
            'sw_power_cap': getattr(nv, 'nvmlClocksThrottleReasonSwPowerCap', 0x4),
        }
        for name, bit in table.items():
            if reasons & bit:
                self.reasons.add(name)

    def _sample_smi(self):
        out = subprocess.run(
            ['nvidia-smi', f'--id={self.index}', f'--query-gpu={self.QUERY}', '--format=csv,noheader,nounits'],
            capture_output=True, text=True, timeout=5).stdout.strip().splitlines()
        if out:
            parts = [x.strip() for x in out[0].split(',')]
            self.samples.append(float(parts[0]))
            self.max_mhz = float(parts[1])
            for name, value in zip(self.NAMES, parts[2:]):
                if value.lower().startswith('active'):
                    self.reasons.add(name)

    def _run(self):
        while not self._stop.is_set():
            try:
                if self._nvml is not None:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:  # pylint: disable=broad-except
                pass
            self._stop.wait(self.period if self._nvml is not None else 0.02)

    def __enter__(self):
        self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        self._thread.join(timeout=10)

    def summary(self):
        return {
            'sm_mhz': statistics.median(self.samples) if self.samples else None,
            'sm_max_mhz': self.max_mhz,
            'reasons': sorted(self.reasons),
            'samples': len(self.samples),
            'source': self.source,
        }


def make_actions(cfg, B, device, ring, seed):
    import torch

    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    nc, nt = cfg['num_cameras'], cfg['num_targets']
    scale = torch.tensor([cfg['camera_rotation_step'], cfg['camera_zooming_step']], device=device)
    cams, tgts = [], []
    for _ in range(ring):
        cams.append((torch.rand((B, max(nc, 1), 2), device=device, generator=gen) * 2 - 1) * scale)
        tgts.append((torch.rand((B, nt, 2), device=device, generator=gen) * 2 - 1) * cfg['target_step_size'])
    return cams, tgts


def cpu_arm(cfg, envs, seconds, threads, seed=0, settle=0):
    """Time the CPU oracle (float64 C port of the reference step) on `envs` environments with
    `threads` host threads for about `seconds` seconds.  Returns (env-steps/s, steps, sample)."""
    import numpy as np

    from oracle.oracle import Oracle

    nc, nt = cfg['num_cameras'], cfg['num_targets']
    ref = Oracle(cfg, envs, num_threads=threads)
    ref.reset(seed=seed)
    ref.set_state({'episode_step': np.random.RandomState(1234).randint(0, cfg['max_episode_steps'] + 1, size=envs).astype(np.int32)})
    rng = np.random.RandomState(seed)
    ring = 4
    cams = [rng.uniform(-1, 1, (envs, nc, 2)) * [cfg['camera_rotation_step'], cfg['camera_zooming_step']] for _ in range(ring)]
    tgts = [rng.uniform(-1, 1, (envs, nt, 2)) * cfg['target_step_size'] for _ in range(ring)]
    aux = ref.alloc_aux()
    aux = {k: (v if k in ('coverage', 'num_delivered') else None) for k, v in aux.items()}
    for k in range(settle):   # the same initial condition as the GPU arm: the running distribution, not the reset state
        ref.step(cams[k % ring], tgts[k % ring], seed=seed, auto_reset=True, aux=aux)
    ref.step(cams[0], tgts[0], seed=seed, auto_reset=True, aux=aux)   # warm-up / calibration
    t0 = time.perf_counter()
    ref.step(cams[1], tgts[1], seed=seed, auto_reset=True, aux=aux)
    per_step = max(time.perf_counter() - t0, 1e-6)
    steps = max(3, min(2000, int(seconds / per_step)))   # a bounded sample: about `seconds` of CPU work
    t0 = time.perf_counter()
    for k in range(steps):
        ref.step(cams[k % ring], tgts[k % ring], seed=seed, auto_reset=True, aux=aux)
    elapsed = time.perf_counter() - t0
    return envs * steps / elapsed, steps, elapsed


def run_reference(args, cfg, workload):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    envs = args.cpu_envs
    # K "steps", each a bounded sample: one pass over `envs` environments
    import numpy as np

    from oracle.oracle import Oracle

    nc, nt = cfg['num_cameras'], cfg['num_targets']
    ref = Oracle(cfg, envs, num_threads=threads)
    ref.reset(seed=0)
    # the same steady state as the GPU arm: episode clocks staggered uniformly over one episode length
    ref.set_state({'episode_step': np.random.RandomState(1234).randint(0, cfg['max_episode_steps'] + 1, size=envs).astype(np.int32)})
    rng = np.random.RandomState(0)
    ring = 4
    cams = [rng.uniform(-1, 1, (envs, nc, 2)) * [cfg['camera_rotation_step'], cfg['camera_zooming_step']] for _ in range(ring)]
    tgts = [rng.uniform(-1, 1, (envs, nt, 2)) * cfg['target_step_size'] for _ in range(ring)]
    aux = {k: (v if k in ('coverage', 'num_delivered') else None) for k, v in ref.alloc_aux().items()}
    for k in range(args.settle + args.warmup):
        ref.step(cams[k % ring], tgts[k % ring], seed=0, auto_reset=True, aux=aux)
    t0 = time.perf_counter()
    for k in range(args.steps):
        ref.step(cams[k % ring], tgts[k % ring], seed=0, auto_reset=True, aux=aux)
    elapsed = time.perf_counter() - t0
    value = envs * args.steps / elapsed
    sample = f'{envs} envs x {args.steps} steps, float64 C port of the reference step (oracle/mate_oracle.c), {threads} pthreads'
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * elapsed / args.steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': workload, 'envs_per_gpu': envs, 'actions': 'uniform random joint actions (ring of 4 pre-generated batches)', 'auto_reset': True,
                   'episode_clocks': 'staggered uniformly over one episode: %.1f resets per step' % (envs / (cfg['max_episode_steps'] + 1.0)),
                   'initial_state': '%d untimed steps after the reset, before the W warm-up steps (positions drawn from the running distribution, not from reset)' % args.settle},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': threads, 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'agent_steps_per_s': value * (nc + nt),
    }
    print(json.dumps(line), flush=True)


def kernel_sources_sha16():
    """Identifies the kernel build a committed ncu traffic figure belongs to."""
    import hashlib

    h = hashlib.sha256()
    for name in ('mate_step.cuh', 'mate_common.cuh'):
        with open(os.path.join(ROOT, 'mate_b200', 'csrc', name), 'rb') as f:
            h.update(f.read())
    return h.hexdigest()[:16]


# the other workloads BASELINE.json names (the headline one is --config / --envs): (config, environments per GPU)
OTHER_CONFIGS = [('MATE-4v2-9.yaml', 65536), ('MATE-4v8-0.yaml', 65536), ('MATE-Navigation.yaml', 65536), ('MATE-8v8-9.yaml', 32768)]


def time_workload(config_name, B, steps, warmup, rank, world, local_rank, stagger=True, sample_clocks=False, settle=0):
    """`settle` steps that move the batch from its reset state to the running distribution, then W untimed + K timed
    steps of one workload on this rank's GPU; CUDA events, max over ranks.
    Returns (record, sim, cfg, (cams, tgts))."""
    import numpy as np
    import torch
    import torch.distributed as dist

    from mate_b200 import _abi
    from mate_b200.config import flatten_config, read_config
    from mate_b200.sim import BatchedSim

    cfg = flatten_config(read_config(config_name))
    nc, nt, no = cfg['num_cameras'], cfg['num_targets'], cfg['num_obstacles']
    device = torch.device('cuda', local_rank)
    sim = BatchedSim(cfg, B, device=local_rank, env_index_base=rank * B)
    sim.reset(seed=0)
    # Steady state: under random actions an episode lasts max_episode_steps + 1 steps, so a fresh batch would
    # never reach a reset inside the timed region.  Stagger the episode clocks uniformly over one episode
    # length: B / (max_episode_steps + 1) environments finish and are re-initialised in EVERY step.
    if stagger:
        steps0 = np.random.RandomState(1234 + rank).randint(0, cfg['max_episode_steps'] + 1, size=B).astype(np.int32)
        sim.set_state({'episode_step': steps0})
    ring = 8
    cams, tgts = make_actions(cfg, B, device, ring, seed=rank)
    sim.alloc_aux()
    # per-step info tensors the reference returns in its info dicts: keep only the cheap ones
    for name, ctype, _, _ in _abi.AUX_FIELDS:
        if name not in ('coverage', 'num_delivered'):
            setattr(sim._aux_struct, name, ctype())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    # Initial condition: after reset every target stands at its start and every camera looks where the reset put it; with
    # staggered clocks that never happens again once the batch runs (the first ~60 steps see 51 % non-zero target entries,
    # the running state 29 %, profiles/r2h_summary.md).  `settle` untimed steps draw the state from the running distribution.
    for k in range(settle):
        sim.step(cams[k % ring], tgts[k % ring], auto_reset=True, aux=True)
    for k in range(warmup):
        sim.step(cams[(settle + k) % ring], tgts[(settle + k) % ring], auto_reset=True, aux=True)
    barrier()
    launches0 = sim.launch_count
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(local_rank) if sample_clocks else None
    if sampler is not None:
        sampler.__enter__()
        time.sleep(0.01)   # the sampling thread is up and has taken its first sample before the region starts
    barrier()
    # The timed launches are enqueued behind a short device-side delay (1 ms), so that the host is ahead of the
    # device for the whole region and no launch gap is timed; the events bracket exactly the K step launches.
    torch.cuda._sleep(2_000_000)   # pylint: disable=protected-access
    start.record()
    for k in range(steps):
        sim.step(cams[k % ring], tgts[k % ring], auto_reset=True, aux=True)
    stop.record()
    barrier()
    if sampler is not None:
        sampler.__exit__(None, None, None)
    elapsed_ms = start.elapsed_time(stop)
    launches = sim.launch_count - launches0
    if world > 1:
        t = torch.tensor([elapsed_ms], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    A = algorithmic_bytes(nc, nt, no)
    peak, peak_src = measured_peak_gbs()
    ms = elapsed_ms / max(steps, 1)
    record = {
        'workload': f'{config_name.replace(".yaml", "")} x {B} envs/GPU', 'ms_per_step': ms,
        'env_steps_per_s': world * B * steps / (elapsed_ms * 1e-3), 'steps': steps, 'warmup': warmup,
        'algorithmic_bytes_per_env_step': A, 'achieved_gbs': A * B / (ms * 1e-3) / 1e9,
        'frac': A * B / (ms * 1e-3) / 1e9 / peak, 'kernel': 'mate_step_kernel2<%d,%d,%d>' % (nc, nt, no),
        'gpu_launches': launches, 'elapsed_ms': elapsed_ms, 'peak': peak, 'peak_source': peak_src,
        'clocks': sampler.summary() if sampler is not None else None,
    }
    return record, sim, cfg, (cams, tgts)


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument('--gpus', type=int, default=1)
    parser.add_argument('--steps', type=int, default=2000)
    parser.add_argument('--warmup', type=int, default=20)
    parser.add_argument('--impl', default='mine', choices=['mine', 'reference'])
    parser.add_argument('--config', default='MATE-4v8-9.yaml')
    parser.add_argument('--envs', type=int, default=65536, help='environments per GPU')
    parser.add_argument('--cpu-envs', type=int, default=0, help='environments per CPU-arm step (0 = the same batch as the GPU arm)')
    parser.add_argument('--cpu-seconds', type=float, default=12.0)
    parser.add_argument('--e2e-steps', type=int, default=10)
    parser.add_argument('--other-steps', type=int, default=200, help='timed steps of each of the other BASELINE workloads (`configs` key)')
    parser.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
    parser.add_argument('--no-e2e', action='store_true', help='skip the host-buffer e2e leg')
    parser.add_argument('--no-configs', action='store_true', help='skip the other BASELINE workloads')
    parser.add_argument('--settle', type=int, default=64, help='untimed steps after the reset, before the warm-up: the state is drawn from the running distribution')
    parser.add_argument('--no-stagger', action='store_true', help='do not stagger the episode clocks (no resets in the timed region)')
    args = parser.parse_args()
    if args.cpu_envs <= 0:
        args.cpu_envs = args.envs

    from mate_b200.config import flatten_config, read_config

    cfg = flatten_config(read_config(args.config))
    nc, nt, no = cfg['num_cameras'], cfg['num_targets'], cfg['num_obstacles']
    workload = f'{args.config.replace(".yaml", "")} x {args.envs} envs/GPU'
    if args.impl == 'reference':
        run_reference(args, cfg, workload)
        return

    import torch
    import torch.distributed as dist

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    torch.cuda.set_device(local_rank)
    device = torch.device('cuda', local_rank)
    B = args.envs
    warmup = max(args.warmup, 3)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    head, sim, cfg, (cams, tgts) = time_workload(args.config, B, args.steps, warmup, rank, world, local_rank,
                                                 stagger=not args.no_stagger, sample_clocks=True, settle=args.settle)
    elapsed_ms, launches = head['elapsed_ms'], head['gpu_launches']
    value = head['env_steps_per_s']

    # optional all-reduce of the episode statistics (outside the timed region)
    stats = sim.episode_stats().clone()
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
    stats = stats.cpu().tolist()

    # ---- e2e: the same step through the host-buffer ABI call (pinned host memory) ----
    e2e = None
    if not args.no_e2e:
        host_cam_act = [c.cpu().pin_memory() for c in cams[:2]]
        host_tgt_act = [t.cpu().pin_memory() for t in tgts[:2]]
        out = (torch.zeros((B, nc, sim.dc)).pin_memory(), torch.zeros((B, nt, sim.dt)).pin_memory(),
               torch.zeros((B, 2)).pin_memory(), torch.zeros(B, dtype=torch.uint8).pin_memory())
        # `out` is reused and never written by the host between the calls: MATE_STEP_HOST_ROWS_KEPT
        kept = os.environ.get('MATE_B200_BENCH_ROWS_KEPT', '1') != '0'   # 0: measure the leg that rewrites every byte
        for k in range(3):
            sim.step_host(host_cam_act[k % 2], host_tgt_act[k % 2], out, auto_reset=True, rows_kept=kept)
        barrier()
        t0 = time.perf_counter()
        for k in range(args.e2e_steps):
            sim.step_host(host_cam_act[k % 2], host_tgt_act[k % 2], out, auto_reset=True, rows_kept=kept)
        e2e_s = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([e2e_s], device=device, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_s = float(t.item())
        leg, link_bytes = sim.host_leg_info()
        # the same call for a caller that cannot promise untouched buffers: every byte of the rows is rewritten
        plain_steps = max(3, args.e2e_steps // 2)
        barrier()
        t0 = time.perf_counter()
        for k in range(plain_steps):
            sim.step_host(host_cam_act[k % 2], host_tgt_act[k % 2], out, auto_reset=True, rows_kept=False)
        plain_s = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([plain_s], device=device, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            plain_s = float(t.item())
        plain_leg, _ = sim.host_leg_info()
        legs = ['dense copy', 'rows rebuilt from their non-zero 16-byte chunks', 'changed 64-byte groups patched into the kept rows']
        e2e = {
            'value': world * B * args.e2e_steps / e2e_s, 'unit': UNIT,
            'h2d_bytes_per_step': B * (nc + nt) * 2 * 4,
            'd2h_bytes_per_step': B * (4 * (nc * sim.dc + nt * sim.dt) + 8 + 1),
            'steps': args.e2e_steps,
            'd2h_row_bytes_on_link_last_step': link_bytes,
            'd2h_leg': legs[leg],
            'value_without_rows_kept': world * B * plain_steps / plain_s, 'd2h_leg_without_rows_kept': legs[plain_leg],
            'note': 'mate_b200_step_host: pinned host actions in, observations/rewards/done out (dense rows in host memory), chunked over 4 streams; device -> host leg chosen by the library from the host threads it has: dense copy, or all-zero 16-byte chunks dropped on the device and the dense rows rebuilt by host threads (MATE_B200_HOST_COMPACT); the output buffers are reused and untouched between steps, which the call is told (MATE_STEP_HOST_ROWS_KEPT: only the 64-byte groups that differ from the previous step cross the link and are rewritten)',
        }
        del out, host_cam_act, host_tgt_act
    sim.close()
    del sim, cams, tgts
    torch.cuda.empty_cache()

    # ---- the other workloads BASELINE.json names, same protocol, short runs (device-resident, CUDA events) ----
    others = []
    if not args.no_configs:
        for name, envs in OTHER_CONFIGS:
            if name == args.config and envs == B:
                continue
            rec, osim, _, _ = time_workload(name, envs, args.other_steps, warmup, rank, world, local_rank, stagger=not args.no_stagger, settle=args.settle)
            osim.close()
            del osim
            torch.cuda.empty_cache()
            others.append({k: rec[k] for k in ('workload', 'ms_per_step', 'env_steps_per_s', 'steps', 'warmup', 'frac',
                                               'achieved_gbs', 'algorithmic_bytes_per_env_step', 'kernel')})

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    A = head['algorithmic_bytes_per_env_step']
    traffic = ncu_traffic(args.config.replace('.yaml', ''))
    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
        'warmup': warmup, 'ms_per_step': elapsed_ms / args.steps, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64 decision state, f32 I/O', 'data': 'synthetic',
        'config': {
            'workload': workload, 'envs_per_gpu': B, 'actions': 'uniform random joint actions (ring of 8 pre-generated batches)',
            'auto_reset': True,
            'episode_clocks': 'all zero' if args.no_stagger else 'staggered uniformly over one episode: %.1f resets per step' % (B / (cfg['max_episode_steps'] + 1.0)),
            'l2': 'per-step working set (observations %.0f MB + state) exceeds the 126 MB L2' % (B * 4 * (nc * head_dims(cfg)[0] + nt * head_dims(cfg)[1]) / 1e6),
            'timing': 'CUDA events around the K step launches, enqueued behind a 1 ms device-side delay (no launch gap inside the region)',
            'initial_state': '%d untimed steps after the reset, before the W warm-up steps (positions drawn from the running distribution, not from reset)' % args.settle,
        },
        'agent_steps_per_s': value * (nc + nt),
        'gpu_launches': launches,
        'roofline': {
            'bound': 'hbm', 'achieved': head['achieved_gbs'], 'peak': head['peak'], 'unit': 'GB/s', 'frac': head['frac'],
            'traffic': traffic.get('bytes'), 'traffic_kernel_sources_sha16': traffic.get('kernel_sources_sha16'),
            'traffic_matches_this_build': traffic.get('kernel_sources_sha16') == kernel_sources_sha16() if traffic.get('bytes') else None,
            'kernel': head['kernel'], 'algorithmic_bytes_per_env_step': A,
            'kernel_ms': head['ms_per_step'], 'peak_source': head['peak_source'],
        },
        'clocks': head['clocks'],
        'episode_stats': {'episodes': stats[0], 'sum_return': stats[1], 'sum_length': stats[2], 'env_steps': stats[5],
                          'resets_from_prepared_state': stats[6], 'resets_in_place': stats[7]},
    }
    if others:
        line['configs'] = others
    if e2e is not None:
        line['e2e'] = e2e
    if not args.no_cpu and world == 1:   # the CPU baseline is reported at N = 1 only
        threads = os.cpu_count() or 1
        cpu_value, cpu_steps, cpu_elapsed = cpu_arm(cfg, args.cpu_envs, args.cpu_seconds, threads, settle=args.settle)
        line['cpu_baseline'] = {
            'value': cpu_value, 'unit': UNIT, 'cores': threads, 'kind': 'port',
            'sample': f'{args.cpu_envs} envs x {cpu_steps} steps ({cpu_elapsed:.1f} s) of the same workload, '
                      f'float64 C port of the reference step (oracle/mate_oracle.c), {threads} pthreads',
        }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def head_dims(cfg):
    nc, nt, no = cfg['num_cameras'], cfg['num_targets'], cfg['num_obstacles']
    return 22 + 5 * nt + 4 * no + 7 * nc, 27 + 7 * nc + 4 * no + 5 * nt


if __name__ == '__main__':
    main()
