#!/usr/bin/env python
"""Benchmark of the DPoser hot path: poses/sec for reverse-SDE sampling + SMPL LBS.

    python bench.py --gpus N --steps K --warmup W [--workload sample_lbs|lbs|sample|completion] [--impl reference]

One "step" = one pass of the hot path over one batch of synthetic input resident in HBM:
  sample_lbs (default): x_T[B,63] -> 1000-step reverse SDE (EM predictor, subVP, config default) ->
                        de-normalise -> SMPL LBS forward (6890 verts / 45 joints).  B = 65536 per GPU
                        (BASELINE.json configs[1] batch; the metric "reverse-SDE sampling + SMPL LBS").
  lbs:     configs[1] alone (SMPL LBS forward, B = 65536).
  sample:  the sampler alone.     completion: configs[2] (imputation sampler, 40960 rows).
Prints ONE JSON line (rank 0).  Multi-GPU: rows are independent, every rank processes its own B rows
(weak scaling, no data-path collective); timing is barrier + CUDA events, max over ranks.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SCORE_FLOP_PER_ROW = 8646656          # 2*(63*1024 + 4*1024^2 + 1024*63), batch-uniform t
LBS_BYTES_PER_POSE = 83560            # verts 6890*12 + joints 45*12 + inputs 85*4
# DRAM traffic from the committed `ncu --set full` capture (profiles/r1_ncu_full_summary_final.md):
#   fused sampler: 1.201 GB for 37 888 rows x 4 steps (mostly write-back of the L2-resident activation scratch)
#   LBS (65 536 poses): fused blend + skinning kernel 5.369 GB written + 0.450 GB read (the output plus operand refills)
SAMPLER_DRAM_BYTES_PER_ROW_STEP = 1.201e9 / (37888 * 4)
LBS_DRAM_BYTES_PER_POSE = (5.369e9 + 0.450e9) / 65536
N_SDE = 1000


def load_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d['hbm_gbs'], tc_burst=d['bf16_tflops'], tc_sustained=d['bf16_tflops_sustained'],
                    src='measured')
    return dict(hbm=6650.0, tc_burst=1590.0, tc_sustained=1400.0, src='fallback')   # B200_PROFILING.md fallback


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                          '-i', str(self.index), '-lms', '200'], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i] == 'Active' for r in self.rows)]
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': reasons, 'samples': len(sm)}


# ------------------------------------------------------------------------------------------- CPU reference arm
def cpu_reference(workload, budget_s=20.0):
    """The reference algorithm on the host cores (oracle port: torch-CPU restatement pinned bit-exactly to the
    real reference by tests/golden).  Bounded sample, linear in rows and steps (all rows independent)."""
    from dposer_b200 import synthetic
    from oracle import lbs_ref, score_ref
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    t_pose = 0.0
    sample = []
    if workload in ('sample_lbs', 'sample', 'completion'):
        sd = score_ref.make_state_dict(42)
        sde = score_ref.SubVP(0.1, 20., N_SDE)
        rows, steps = 512, 8
        gen = torch.Generator().manual_seed(1234)
        x = torch.randn(rows, 63, generator=gen)
        kw = {}
        if workload == 'completion':
            _, mask, obs = synthetic.completion_inputs(n_partial=rows, hypotheses=1)
            kw = dict(observation=obs, mask=mask, task='completion')
        score_ref.pc_sample(sd, sde, x, 1e-3, n_run=1, **kw)                       # warm-up
        t0 = time.perf_counter()
        done = 0
        while time.perf_counter() - t0 < budget_s * 0.6:
            score_ref.pc_sample(sd, sde, x, 1e-3, n_run=steps, **kw)
            done += steps
        dt = time.perf_counter() - t0
        t_pose += dt / done * N_SDE / rows
        sample.append(f'pc_sampler on {rows} rows: {done} Euler-Maruyama steps timed in runs of {steps}, '
                      f'per-step time x {N_SDE} steps')
    if workload in ('sample_lbs', 'lbs'):
        m = synthetic.make_body_tensors('smpl')
        rows = 512
        inp = synthetic.lbs_inputs(rows, 'smpl')
        pose, shape = synthetic.full_pose_from(inp, 'smpl')
        lbs_ref.body_forward(m, shape[:64], pose[:64], inp['trans'][:64])
        t0 = time.perf_counter()
        done = 0
        while time.perf_counter() - t0 < budget_s * 0.4:
            lbs_ref.body_forward(m, shape, pose, inp['trans'], chunk=256)
            done += rows
        t_pose += (time.perf_counter() - t0) / done
        sample.append(f'SMPL LBS {done} poses in chunks of 256')
    return dict(value=1.0 / t_pose, unit='poses/s', cores=cores, kind='port', sample='; '.join(sample))


def run_reference_arm(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    t0 = time.perf_counter()
    vals = []
    for _ in range(max(1, args.warmup)):
        cpu_reference(args.workload, budget_s=4.0)
    for _ in range(args.steps):
        vals.append(cpu_reference(args.workload, budget_s=12.0))
    v = sum(x['value'] for x in vals) / len(vals)
    base = vals[-1]
    base['value'] = v
    B = default_batch(args)
    line = {'impl': 'reference', 'metric': 'poses/sec: DPoser reverse-SDE sampling + SMPL LBS', 'value': v,
            'unit': 'poses/s', 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': 1000.0 * B / v, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32', 'data': 'synthetic', 'config': workload_config(args, B),
            'cpu_baseline': base,
            'e2e': {'value': v, 'unit': 'poses/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0, 'wall_s': time.perf_counter() - t0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------- GPU arm
def default_batch(args):
    if args.batch:
        return args.batch
    return 40960 if args.workload == 'completion' else 65536


def workload_config(args, B):
    names = {'sample_lbs': f'reverse-SDE sampling (subVP, EM predictor, N={args.sde_steps}, eps=1e-3, config default) of '
                           f'{B} poses/GPU + SMPL LBS forward (6890 verts, 24+21 joints, 10 betas) = configs[0] '
                           'pipeline at configs[1] batch',
             'lbs': f'configs[1]: SMPL LBS forward only, batch {B}/GPU',
             'sample': f'reverse-SDE sampling only, N={args.sde_steps}, {B} poses/GPU',
             'completion': f'configs[2]: completion by imputation sampler, {B} rows/GPU, N={args.sde_steps}'}
    return {'workload': names[args.workload], 'batch_per_gpu': B, 'sde_steps': args.sde_steps,
            'l2': 'L2 flushed (256 MiB write) between timed steps', 'engine': args.engine,
            'weights': 'random-init ScoreModelFC (seed 42), synthetic SMPL tensors (seed 7)'}


def run_gpu_arm(args):
    from dposer_b200 import _lib as L
    from dposer_b200 import dist as D
    from dposer_b200 import sampling, sde_lib, synthetic
    from dposer_b200.body_model import BodyModel
    from dposer_b200.misc import Posenormalizer
    import types

    rank, local, world = D.init_from_env()
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    peaks = load_peaks()
    B = default_batch(args)
    wl = args.workload
    engine = {'auto': L.ENGINE_AUTO, 'tc': L.ENGINE_TC, 'fp32': L.ENGINE_FP32}[args.engine]

    model = synthetic.make_score_model(42).to(dev)
    model.engine = engine
    cfg = synthetic.default_config()
    sde = sde_lib.subVPSDE(0.1, 20., args.sde_steps)
    norm = Posenormalizer(None, device=dev, normalize=True, min_max=False, rot_rep='axis')
    body = synthetic.make_body_tensors('smpl')
    bm = BodyModel(body, batch_size=B, model_type='smpl').to(dev)
    fn = sampling.get_sampling_fn(cfg, sde, (B, 63), lambda x: x, 1e-3, device=dev, return_trajs=False)

    gen = torch.Generator().manual_seed(1234 + rank)
    xT_host = torch.randn(B, 63, generator=gen).pin_memory()
    lin = synthetic.lbs_inputs(B, 'smpl', seed=11 + rank)
    lbs_host = {k: v.pin_memory() for k, v in lin.items()}
    comp = None
    if wl == 'completion':
        _, mask, obs = synthetic.completion_inputs(n_partial=B // 10, hypotheses=10, seed=21 + rank)
        comp = (obs.pin_memory(), mask.pin_memory())
    task_args = types.SimpleNamespace(task='completion') if wl == 'completion' else None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def to_dev(d):
        return {k: v.to(dev, non_blocking=True) for k, v in d.items()}

    def hot_path(xT, lbs_in, comp_dev):
        """inputs resident in HBM -> results resident in HBM"""
        res = {}
        if wl in ('sample_lbs', 'sample', 'completion'):
            kw = {}
            if comp_dev is not None:
                kw = dict(observation=comp_dev[0], mask=comp_dev[1], args=task_args)
            _, x0 = fn(model, z=xT, **kw)
            res['poses'] = x0
        if wl == 'sample_lbs':
            # random-init weights drive |x| to ~1e4: squash to a plausible angle range so the
            # LBS input is well-conditioned; the squash is outside the measured kernels' arithmetic
            pose = torch.cat([norm.offline_denormalize(torch.tanh(res['poses'] * 1e-4)),
                              torch.zeros(B, 6, device=dev)], dim=1)
            out = bm(root_orient=lbs_in['root_orient'], pose_body=pose, betas=lbs_in['betas'], trans=lbs_in['trans'])
            res['joints'], res['verts'] = out.Jtr, out.v
        elif wl == 'lbs':
            out = bm(**lbs_in)
            res['joints'], res['verts'] = out.Jtr, out.v
        return res

    results_host = {}

    def step(e2e):
        if e2e:
            xT = xT_host.to(dev, non_blocking=True)
            lbs_in = to_dev(lbs_host)
            comp_dev = None if comp is None else tuple(c.to(dev, non_blocking=True) for c in comp)
        else:
            xT, lbs_in, comp_dev = xT_res, lbs_res, comp_res
        with torch.no_grad():
            res = hot_path(xT, lbs_in, comp_dev)
        if e2e:
            for k in ('poses', 'joints'):
                if k in res:
                    if k not in results_host:
                        results_host[k] = torch.empty(res[k].shape, dtype=res[k].dtype).pin_memory()
                    results_host[k].copy_(res[k], non_blocking=True)
        return res

    xT_res = xT_host.to(dev)
    lbs_res = to_dev(lbs_host)
    comp_res = None if comp is None else tuple(c.to(dev) for c in comp)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def timed(e2e, n_steps):
        total = 0.0
        for _ in range(n_steps):
            flush.fill_(1)
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            step(e2e)
            e1.record()
            torch.cuda.synchronize()
            total += D.max_over_ranks(e0.elapsed_time(e1), dev)
        return total / n_steps

    for _ in range(max(3, args.warmup)):
        step(False)
    torch.cuda.synchronize()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    ms = timed(False, args.steps)
    step(True)                      # untimed: allocates the pinned result buffers the end-to-end steps copy into
    torch.cuda.synchronize()
    ms_e2e = timed(True, max(1, min(args.steps, 3)))
    clk = clocks.stop() if rank == 0 else None

    # per-kernel timing of the two stages for the roofline (same stream, CUDA events, inputs resident)
    def time_stage(f, reps=3):
        f()
        torch.cuda.synchronize()
        best = 1e30
        for _ in range(reps):
            flush.fill_(1)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            f()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return best

    roof, roof_lbs = None, None
    with torch.no_grad():
        if wl in ('sample_lbs', 'sample', 'completion'):
            kw = {} if comp_res is None else dict(observation=comp_res[0], mask=comp_res[1], args=task_args)
            t_s = time_stage(lambda: fn(model, z=xT_res, **kw), reps=2)
            ach = SCORE_FLOP_PER_ROW * B * args.sde_steps / (t_s * 1e-3) / 1e12
            roof = {'kernel': 'fused sampler (score net x N steps)', 'bound': 'tensor', 'achieved': ach,
                    'peak': peaks['tc_sustained'], 'unit': 'TFLOP/s', 'frac': ach / peaks['tc_sustained'],
                    'traffic': SAMPLER_DRAM_BYTES_PER_ROW_STEP * B * args.sde_steps,
                    'traffic_note': 'DRAM bytes, scaled per row-step from the committed ncu capture', 'ms': t_s,
                    'peak_source': peaks['src'] + ' (bf16 sustained)'}
        if wl in ('sample_lbs', 'lbs'):
            t_l = time_stage(lambda: bm(**lbs_res), reps=5)
            ach = LBS_BYTES_PER_POSE * B / (t_l * 1e-3) / 1e9
            roof_lbs = {'kernel': 'SMPL LBS forward', 'bound': 'hbm', 'achieved': ach, 'peak': peaks['hbm'],
                        'unit': 'GB/s', 'frac': ach / peaks['hbm'], 'traffic': LBS_DRAM_BYTES_PER_POSE * B,
                        'traffic_note': 'DRAM bytes, scaled per pose from the committed ncu capture (65536 poses)',
                        'algorithmic_bytes': LBS_BYTES_PER_POSE * B, 'ms': t_l, 'peak_source': peaks['src']}
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
    if rank != 0:
        return
    tc = engine != L.ENGINE_FP32 and B >= 64
    launches = 0
    if wl in ('sample_lbs', 'sample', 'completion'):
        launches += (1 + 1) if tc else (1 + args.sde_steps * 12)        # time table + fused / per-layer kernels
    if wl in ('sample_lbs', 'lbs'):
        launches += 6 if tc else 3          # pose, (featop, blend, skinop, skin | vertex), gather kernels
    h2d = xT_host.numel() * 4 + sum(v.numel() * 4 for v in lbs_host.values()) + \
        (0 if comp is None else sum(c.numel() * 4 for c in comp))
    d2h = sum(v.numel() * 4 for v in results_host.values())
    value = world * B / (ms * 1e-3)
    line = {'metric': 'poses/sec: DPoser reverse-SDE sampling + SMPL LBS', 'value': value, 'unit': 'poses/s',
            'n_gpus': world, 'steps': args.steps, 'warmup': max(3, args.warmup), 'ms_per_step': ms,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f16 operands / f32 accumulate (score net, tcgen05); f32 (LBS)' if tc else 'f32',
            'data': 'synthetic', 'config': workload_config(args, B), 'clocks': clk,
            'e2e': {'value': world * B / (ms_e2e * 1e-3), 'unit': 'poses/s', 'h2d_bytes_per_step': h2d,
                    'd2h_bytes_per_step': d2h, 'ms_per_step': ms_e2e,
                    'note': 'pinned host inputs -> device, hot path, generated poses + joints -> pinned host'},
            'gpu_launches': launches * args.steps,
            'roofline': roof if roof is not None else roof_lbs}
    if roof is not None and roof_lbs is not None:
        line['roofline_lbs'] = roof_lbs
    if not args.no_cpu:
        line['cpu_baseline'] = cpu_reference(wl, budget_s=args.cpu_budget)
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='sample_lbs', choices=['sample_lbs', 'lbs', 'sample', 'completion'])
    ap.add_argument('--batch', type=int, default=0, help='rows per GPU (default: the BASELINE config size)')
    ap.add_argument('--sde-steps', type=int, default=N_SDE)
    ap.add_argument('--engine', default='auto', choices=['auto', 'tc', 'fp32'])
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
    ap.add_argument('--cpu-budget', type=float, default=20.0)
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == '__main__':
    main()
