#!/usr/bin/env python
"""bench.py -- view-tuples/sec of the hot path on synthetic 5-view x 1024-keypoint tuples.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--tuples B]

A *step* is one pass of the hot path over one batch of B synthetic tuples per GPU
(BASELINE.json configs[2]: ScanNet-shape 5-tuple, 1024 kpts, 28-layer matcher, confidence head,
10 x {w8pt + two-view BA}, spanning tree, global GN/LM BA).  One JSON line on rank 0; see
DESIGN.md §measurement for every field.  `--impl reference` times the CPU port of the reference
path (oracle/) on the host cores -- /root/reference does not exist on the GPU box.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

T_VIEWS, N_KPTS = 5, 1024
LAYERS = (['self'] + ['cross'] * 3) * 7
GAIN = 12.0          # final_proj gain of the seeded weights: gives the assignment real structure
METRIC = 'view-tuples/sec @1024 kpts 5-view'

# algorithmic work: attention QK^T + PV FLOPs per (layer launch, tuple)
D = 256
FLOPS_SELF = 4 * N_KPTS * N_KPTS * D * T_VIEWS                       # per tuple, per self layer
FLOPS_CROSS = 4 * N_KPTS * (T_VIEWS - 1) * N_KPTS * D * T_VIEWS       # per tuple, per cross layer
SINKHORN_BYTES_PER_PAIR = 100 * 2 * (N_KPTS + 1) ** 2 * 4 + 2 * (N_KPTS + 1) ** 2 * 4


TF32_PEAK_TFLOPS = 148 * 4096 * 1.965e9 / 1e12      # tcgen05 kind::tf32 issue floor x SMs x max SM clock


def load_traffic(tuples):
    """Per-launch DRAM traffic of the dominant kernels from the committed ncu capture (profiles/ncu_traffic.json),
    valid for the batch size it was captured at."""
    try:
        t = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'profiles', 'ncu_traffic.json')))
        if t.get('tuples_per_step') == tuples:
            return t['attention']['avg_bytes_per_launch'], t['sinkhorn']['avg_bytes_per_launch']
    except Exception:
        pass
    return None, None


def load_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return {'hbm_gbs': d['hbm_gbs'], 'tflops': d.get('bf16_tflops_sustained', d['bf16_tflops']),
                'source': 'measured (MEASURED_PEAKS.json, bf16 sustained / copy)'}
    return {'hbm_gbs': 6650.0, 'tflops': 1400.0, 'source': 'fallback (B200_PROFILING.md)'}


class ClockSampler(threading.Thread):
    """SM clock / throttle reasons sampled DURING the timed region.  In-process NVML (a few microseconds
    per query) -- an `nvidia-smi` subprocess takes ~0.5 s per sample and stalls kernel launches while it
    holds the driver lock, which showed up as a 100 ms hiccup inside the e2e timing; it is only the
    fallback when NVML cannot be loaded."""

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.samples = []
        self.stop_flag = False
        self.nvml = None
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            try:
                self.handle = pynvml.nvmlDeviceGetHandleByUUID('GPU-' + str(torch.cuda.get_device_properties(gpu_index).uuid))
            except Exception:
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        n, h = self.nvml, self.handle
        sm = n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM)
        mx = n.nvmlDeviceGetMaxClockInfo(h, n.NVML_CLOCK_SM)
        try:
            r = n.nvmlDeviceGetCurrentClocksEventReasons(h)
        except Exception:
            r = n.nvmlDeviceGetCurrentClocksThrottleReasons(h)
        act = lambda bit: 'Active' if (r & bit) else 'Not Active'
        # NVML reason bits: SwPowerCap 0x4, HwSlowdown 0x8, SwThermalSlowdown 0x20, HwThermalSlowdown 0x40
        return [str(sm), str(mx), '', act(0x8), act(0x40), act(0x20), act(0x4)]

    def run(self):
        q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')
        while not self.stop_flag:
            try:
                if self.nvml is not None:
                    self.samples.append(self._sample_nvml())
                    time.sleep(0.02)
                    continue
                o = subprocess.run(['nvidia-smi', '-i', str(self.gpu), '--query-gpu=' + q,
                                    '--format=csv,noheader,nounits'], capture_output=True, text=True, timeout=5)
                f = [x.strip() for x in o.stdout.strip().split(',')]
                if len(f) >= 7:
                    self.samples.append(f)
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.samples:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['unavailable']}
        sm = sorted(float(s[0]) for s in self.samples)
        reasons = []
        for i, name in ((3, 'hw_slowdown'), (4, 'hw_thermal_slowdown'), (5, 'sw_thermal_slowdown'), (6, 'sw_power_cap')):
            if any(s[i].lower().startswith('active') for s in self.samples):
                reasons.append(name)
        return {'sm_mhz': sm[len(sm) // 2], 'sm_max_mhz': float(self.samples[0][1]), 'reasons': reasons,
                'samples': len(self.samples), 'source': 'nvml' if self.nvml is not None else 'nvidia-smi'}


# ---------------------------------------------------------------------------------------------
# CPU port of the reference path (oracle) -- cpu_baseline leg and the --impl reference arm
# ---------------------------------------------------------------------------------------------
def cpu_reference_tuple(sd, data_np, cap2=128, cap_ba=32):
    """One tuple through the CPU restatement of the reference: full-size matcher (5 x 1024, 28
    layers, 100 Sinkhorn iterations, confidence head), then the pose stage on a bounded number of
    matches per pair (the reference's dense (6+3n)^2 two-view BA and a dense global BA are
    cubic in n; cap2 / cap_ba say what the sample was)."""
    from oracle.matcher_torch import matcher_forward
    from oracle import pose as P, mvba as M
    one = {k: (v[:1] if isinstance(v, np.ndarray) else v) for k, v in data_np.items()}
    t0 = time.time()
    res = matcher_forward(sd, {'GNN_layers': LAYERS, 'multi_frame_matching': True}, one)
    t1 = time.time()
    K = one['intr0'].astype(np.float64)
    rel, weight, pm = {}, {}, {}
    for b in range(T_VIEWS):
        for a in range(b):
            m = res['matches%d_%d_%d' % (a, a, b)][0]
            c = res['conf_scores_%d_%d' % (a, b)][0, :, 0].astype(np.float64)
            valid = np.nonzero((m >= 0) & (c > 0))[0][:cap2]
            if valid.size < 8:
                continue
            k0 = one['keypoints%d' % a][0][valid].astype(np.float64)[None]
            k1 = one['keypoints%d' % b][0][m[valid]].astype(np.float64)[None]
            cc = c[valid][None, :, None]
            Tw, info = P.estimate_relative_pose_w8pt(k0, k1, K, K, cc, determine_inliers=True)
            cn = info['confidence'].copy()
            cn[~info['pos_depth_mask']] = 0
            ext, vb = P.run_bundle_adjust_2_view(info['kpts0_norm'], info['kpts1_norm'], cn, Tw, 10)
            Tp = ext[0] if vb[0] else Tw[0]
            rel[(a, b)], weight[(a, b)] = Tp, int(valid.size)
            pm[(a, b)] = (info['kpts0_norm'][0][:cap_ba], info['kpts1_norm'][0][:cap_ba], c[valid][:cap_ba])
    n_ok = len(rel)
    if n_ok:
        extr0, _ = M.spanning_tree_extrinsics(T_VIEWS, rel, weight)
        M.solve(M.build_problem(T_VIEWS, pm, extr0))
    t2 = time.time()
    return t1 - t0, t2 - t1, n_ok


def torch_gpu_port(sd, data_np, dev, B, stage_ms):
    import torch
    from oracle.matcher_torch import matcher_forward
    nb = min(B, 4)                                      # the port materialises prob[B,4,N,4N]: keep the batch small
    data = {k: (v[:nb] if isinstance(v, np.ndarray) and v.ndim >= 1 and v.shape[0] == B else v) for k, v in data_np.items()
            if not k.startswith('landmark') and not k.startswith('pose')}
    data = {k: (torch.empty(v.shape, device='meta') if k.startswith('image') else torch.from_numpy(v).to(dev))
            if isinstance(v, np.ndarray) else v for k, v in data.items()}
    out = {}
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    try:
        for name, tf32 in (('tf32_allowed', True), ('fp32', False)):
            torch.backends.cuda.matmul.allow_tf32 = tf32
            torch.backends.cudnn.allow_tf32 = tf32
            for _ in range(2):
                matcher_forward(sd, {'GNN_layers': LAYERS}, data, device=dev, to_numpy=False)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                matcher_forward(sd, {'GNN_layers': LAYERS}, data, device=dev, to_numpy=False)
            e1.record()
            torch.cuda.synchronize()
            out['matcher_tuples_per_s_' + name] = 3 * nb / (e0.elapsed_time(e1) * 1e-3)
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old
    ours = sum(stage_ms.get(k, 0.0) for k in ('gemm', 'attention', 'sinkhorn', 'score_gemm', 'match', 'conf', 'kenc'))
    out['ours_matcher_tuples_per_s'] = B / (ours * 1e-3)
    out['note'] = ('torch port of the reference matcher (oracle/matcher_torch.py) in eager stock PyTorch on the same GPU, '
                   'batch %d; matcher only; tf32_allowed = torch 1.10 defaults' % nb)
    return out


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    from e2e_multi_view_matching_b200.synthetic import make_state_dict, make_scene_tuple_inputs
    sd = make_state_dict(len(LAYERS), seed=0, final_proj_gain=GAIN)
    data = make_scene_tuple_inputs(1000, T_VIEWS, N_KPTS, batch=1)
    cores = os.cpu_count()
    budget = 240.0
    t_m, t_p, _ = cpu_reference_tuple(sd, data)           # warm-up / sizing step
    per = t_m + t_p
    warm = 0 if per * (args.steps + 1) > budget else min(args.warmup, 1)
    steps = max(1, min(args.steps, int(budget / per) - warm - 1))
    for _ in range(warm):
        cpu_reference_tuple(sd, data)
    t0 = time.time()
    for _ in range(steps):
        cpu_reference_tuple(sd, data)
    dt = time.time() - t0
    val = steps / dt
    sample = ('1 tuple/step: matcher full size (5x1024 kpts, 28 layers, 100 Sinkhorn iters, conf head) + pose '
              'stage on <=128 matches/pair (two-view, dense LU like the reference) and <=32 matches/pair (global BA)')
    line = {'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': 'tuples/s', 'n_gpus': args.gpus,
            'steps': steps, 'warmup': warm, 'ms_per_step': 1e3 * dt / steps, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32 (matcher) / f64 (pose)', 'data': 'synthetic',
            'config': {'workload': 'scannet_5tuple_1024kpts_28layers_mvba', 'tuples_per_step': 1},
            'cpu_baseline': {'value': val, 'unit': 'tuples/s', 'cores': cores, 'kind': 'port', 'sample': sample},
            'e2e': {'value': val, 'unit': 'tuples/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--tuples', type=int, default=14, help='tuples per step per GPU')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-torch-gpu', action='store_true', help='skip the informational stock-PyTorch-on-GPU matcher line')
    ap.add_argument('--gemm-tile', type=int, default=256, choices=[128, 256])
    ap.add_argument('--gemm-kernel', default='persistent', choices=['persistent', 'tile'],
                    help='3xTF32 GEMM kernel: persistent (default) or the one-tile-per-CTA kernel (A/B comparison)')
    ap.add_argument('--math-mode', type=int, default=3, choices=[0, 1, 3],
                    help='3 = tcgen05 3xTF32 (fp32-faithful, default), 1 = tcgen05 single-pass TF32, 0 = fp32 CUDA cores')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'ours' else args.warmup
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    if args.impl == 'reference':
        run_reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from e2e_multi_view_matching_b200 import _lib
    from e2e_multi_view_matching_b200.models.multi_view_matcher import MultiViewMatcher
    from e2e_multi_view_matching_b200.pipeline import MultiViewPipeline, pose_auc
    from e2e_multi_view_matching_b200.synthetic import make_state_dict, make_scene_tuple_inputs
    from e2e_multi_view_matching_b200 import sharding

    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    lib = _lib.lib()
    lib.mvm_set_math_mode(args.math_mode)
    lib.mvm_debug_set_gemm_tile(args.gemm_tile)
    lib.mvm_debug_set_gemm_kernel(1 if args.gemm_kernel == 'persistent' else 0)
    B = args.tuples

    sd = make_state_dict(len(LAYERS), seed=0, final_proj_gain=GAIN)
    model = MultiViewMatcher({'GNN_layers': LAYERS}).eval()
    model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})
    model = model.to(dev)
    pipe = MultiViewPipeline(model)

    # tuples 1000 + (rank*B + k): every rank works on its own shard (weak scaling, no data-path collective)
    data_np = make_scene_tuple_inputs(sharding.tuple_shard(rank, world, B)[0], T_VIEWS, N_KPTS, batch=B)
    keys = [k for k, v in data_np.items() if isinstance(v, np.ndarray) and not k.startswith('image')
            and not k.startswith('landmark')]
    host = {k: torch.from_numpy(data_np[k]).pin_memory() for k in keys}
    meta = {k: torch.empty(v.shape, device='meta') for k, v in data_np.items() if k.startswith('image')}
    data_dev = {k: v.to(dev) for k, v in host.items()}
    data_dev.update(meta)
    data_dev['ids'] = data_np['ids']
    h2d_keys = [k for k in keys if not k.startswith('pose')]
    h2d_bytes = sum(host[k].numel() * host[k].element_size() for k in h2d_keys)
    out_host = {'extrinsics': torch.empty(B, T_VIEWS, 4, 4).pin_memory(),
                'T_pair': torch.empty(B, 10, 4, 4).pin_memory()}
    d2h_bytes = sum(v.numel() * v.element_size() for v in out_host.values())
    loss = torch.zeros(1, device=dev)

    last = {}

    def step_device():
        res, pose = pipe(data_dev)
        last['res'] = res
        if world > 1:   # one scalar all-reduce per step (mirrors the val-loss all_reduce, train.py:104-106)
            loss.copy_(pose['ba_cost'][:, 1].sum().float().reshape(1))
            sharding.all_reduce_step_loss(loss)
        return pose

    # End-to-end step through the public API.  Every step's inputs come from pinned host memory: the copy of
    # step k+1 is issued on a copy stream right after step k's kernels are enqueued (double-buffered device
    # inputs), so it overlaps step k's compute; the step's result (poses) is read back to pinned host memory.
    copy_stream = torch.cuda.Stream(device=dev)
    staged = {}

    def stage_inputs():
        with torch.cuda.stream(copy_stream):
            d = {k: host[k].to(dev, non_blocking=True) for k in h2d_keys}
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        staged['d'], staged['ev'] = d, ev

    def step_e2e():
        if 'd' not in staged:
            stage_inputs()
        d, ev = staged.pop('d'), staged.pop('ev')
        torch.cuda.current_stream().wait_event(ev)
        for t in d.values():
            t.record_stream(torch.cuda.current_stream())
        d.update(meta)
        d['ids'] = data_np['ids']
        res, pose = pipe(d)
        stage_inputs()                               # next step's H2D, overlapped with this step's kernels
        out_host['extrinsics'].copy_(pose['extrinsics'], non_blocking=True)
        out_host['T_pair'].copy_(pose['T_pair'], non_blocking=True)
        if world > 1:
            loss.copy_(pose['ba_cost'][:, 1].sum().float().reshape(1))
            sharding.all_reduce_step_loss(loss)
        return pose

    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def timed(fn, steps):
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.time()
        for a, b in evs:
            flush.zero_()                      # L2 flush between timed iterations (outside the events)
            a.record()
            fn()
            b.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        wall = time.time() - t0
        per_step = [a.elapsed_time(b) for a, b in evs]
        last['per_step_ms'] = per_step
        ms = sum(per_step)
        return sharding.max_over_ranks(ms, dev), wall

    for _ in range(args.warmup):
        step_device()
        step_e2e()
    torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    n0 = lib.mvm_launch_count()
    ms_dev, wall_dev = timed(step_device, args.steps)
    launches = lib.mvm_launch_count() - n0
    staged.clear()                 # the first timed step stages its own inputs inside the timed region
    ms_e2e, wall_e2e = timed(step_e2e, args.steps)
    e2e_steps = [round(x, 2) for x in last['per_step_ms']]
    sampler.stop_flag = True

    # secondary line: the same steps in single-pass TF32 (what torch 1.10 ran on Ampere by default)
    tf32 = None
    if args.math_mode == 3:
        lib.mvm_set_math_mode(1)
        for _ in range(2):
            step_device()
        ms_tf32, _ = timed(step_device, args.steps)
        staged.clear()
        ms_tf32_e2e, _ = timed(step_e2e, args.steps)
        lib.mvm_set_math_mode(3)
        tf32 = (ms_tf32, ms_tf32_e2e)

    # live per-kernel-class timing (CUDA events on the launching stream) over two more steps
    lib.mvm_profile_enable(1)
    prof_steps = 2
    for _ in range(prof_steps):
        flush.zero_()
        pose = step_device()
    torch.cuda.synchronize()
    prof = _lib.profile_collect()
    lib.mvm_profile_enable(0)

    if rank == 0:
        peaks = load_peaks()
        total_tuples = B * args.steps * world
        value = sharding.whole_job_throughput(B, args.steps, world, ms_dev)
        e2e = total_tuples / (ms_e2e * 1e-3)
        traffic_att, traffic_sink = load_traffic(B)
        att_ms, att_n = prof['attention']
        n_self, n_cross = LAYERS.count('self'), LAYERS.count('cross')
        att_flops = (n_self * FLOPS_SELF + n_cross * FLOPS_CROSS) * B * prof_steps   # over the profiled steps
        att_tflops = att_flops / (att_ms * 1e-3) / 1e12 if att_ms > 0 else 0.0
        sk_ms, sk_n = prof['sinkhorn']
        sk_gbs = SINKHORN_BYTES_PER_PAIR * 10 * B * prof_steps / (sk_ms * 1e-3) / 1e9 if sk_ms > 0 else 0.0
        stage_ms = {k: round(v[0] / prof_steps, 4) for k, v in prof.items() if v[1] > 0}
        # pose AUC of the engine on its own synthetic tuples (informational; parity is in tests/)
        errs = MultiViewPipeline.pair_errors({k: v for k, v in data_dev.items() if k.startswith('pose')}, pose, T_VIEWS)
        auc = pose_auc([e[0] for e in errs], [5, 10, 20])
        last_res = last['res']
        line = {
            'metric': METRIC, 'value': value, 'unit': 'tuples/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': ms_dev / args.steps, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': {3: 'tf32x3 on tcgen05 (fp32-faithful) / f64 pose kernels', 1: 'tf32 on tcgen05 / f64 pose kernels',
                      0: 'f32 CUDA cores / f64 pose kernels'}[args.math_mode], 'data': 'synthetic',
            'config': {'workload': 'scannet_5tuple_1024kpts_28layers_mvba', 'tuples_per_step_per_gpu': B,
                       'views': T_VIEWS, 'kpts': N_KPTS, 'gnn_layers': len(LAYERS), 'sinkhorn_iters': 100,
                       'pose': '10x(w8pt+10it 2-view BA) + spanning tree + global LM BA (<=50 it)',
                       'l2': 'flushed between timed steps (256 MB write)', 'math_mode': args.math_mode, 'parallelism': 'dp%d' % world},
            'e2e': {'value': e2e, 'unit': 'tuples/s', 'h2d_bytes_per_step': h2d_bytes, 'd2h_bytes_per_step': d2h_bytes,
                    'ms_per_step': ms_e2e / args.steps, 'ms_steps': e2e_steps,
                    'h2d': 'pinned host -> device on a copy stream, step k+1 staged under step k'},
            'gpu_launches': int(launches),
            'clocks': sampler.summary(),
            'roofline': {'kernel': 'attention (QK^T + PV, all views of one GNN layer per launch)', 'bound': 'tensor',
                         'achieved': att_tflops, 'peak': peaks['tflops'], 'unit': 'TFLOP/s',
                         'frac': att_tflops / peaks['tflops'], 'traffic': traffic_att, 'launches_timed': att_n,
                         'peak_source': peaks['source'],
                         # the path computes in tf32 (half the bf16 rate: M128.N.K8 every N/2 cycles = 4096 FLOP/clk/SM)
                         # and needs three passes to stay fp32-faithful: the reachable algorithmic ceiling
                         'ceiling_tf32x3': TF32_PEAK_TFLOPS / (3.0 if args.math_mode == 3 else 1.0),
                         'frac_of_ceiling': att_tflops / (TF32_PEAK_TFLOPS / (3.0 if args.math_mode == 3 else 1.0))},
            'roofline_sinkhorn': {'kernel': 'sinkhorn (10 pairs x B problems per launch)', 'bound': 'hbm',
                                  'achieved': sk_gbs, 'peak': peaks['hbm_gbs'], 'unit': 'GB/s',
                                  'frac': sk_gbs / peaks['hbm_gbs'], 'traffic': traffic_sink, 'launches_timed': sk_n},
            'stage_ms_per_step': stage_ms,
            'wall_s': {'device_resident': wall_dev, 'e2e': wall_e2e},
            'pose_auc_5_10_20': [round(100 * a, 2) for a in auc],
        }
        if tf32 is not None:
            line['tf32_single_pass'] = {'value': total_tuples / (tf32[0] * 1e-3), 'e2e': total_tuples / (tf32[1] * 1e-3),
                                        'unit': 'tuples/s', 'note': 'math mode 1 (tcgen05 kind::tf32, one pass)'}
        # quality of the synthetic assignment: fraction of returned matches that join the same landmark
        hits = tot = 0
        for b_ in range(T_VIEWS):
            for a_ in range(b_):
                m = last_res['matches%d_%d_%d' % (a_, a_, b_)].cpu().numpy()
                la, lb = data_np['landmark%d' % a_], data_np['landmark%d' % b_]
                for i in range(B):
                    v = m[i] >= 0
                    hits += int((la[i][v] == lb[i][m[i][v]]).sum()); tot += int(v.sum())
        line['match_precision'] = round(hits / max(tot, 1), 4)
        if world == 1 and not args.no_torch_gpu:
            # informational: the op-for-op torch port of the reference matcher run by stock PyTorch (cuBLAS / cuDNN
            # eager) on the same GPU -- what a user of the reference gets by moving its model to the B200.  Matcher
            # only (the reference's pose stage is CPU code); our matcher-only rate from the stage timers beside it.
            try:
                line['torch_gpu_port'] = torch_gpu_port(sd, data_np, dev, B, stage_ms)
            except Exception as e:                                  # never let the extra line break the bench
                line['torch_gpu_port'] = {'error': repr(e)[:200]}
        if world == 1 and not args.no_cpu_baseline:
            t_m, t_p, n_ok = cpu_reference_tuple(sd, data_np)
            line['cpu_baseline'] = {
                'value': 1.0 / (t_m + t_p), 'unit': 'tuples/s', 'cores': os.cpu_count(), 'kind': 'port',
                'sample': '1 tuple: matcher full size %.1f s + pose stage on <=128 matches/pair (2-view) and <=32 '
                          'matches/pair (global BA) %.1f s' % (t_m, t_p)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
