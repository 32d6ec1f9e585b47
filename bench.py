#!/usr/bin/env python
"""bench.py - Match-Tensor scoring throughput (BASELINE.json configs[1]) on N B200s.

A "step" = one pass of the scoring hot path over one batch of synthetic sessions:
B=128 queries x N=10 candidate docs per GPU, Lq=20, Ld=200, E=300, H=128 (64/dir), F=40, C=50,
all lengths at max (the headline throughput set), V=131072, random-init weights.
Weak scaling: every rank scores its own contiguous slice of B*N*world pairs (doc-parallel) and one
all-gather of the fp32 scores follows inside the timed step.

  python bench.py [--gpus N] [--steps K] [--warmup W]        product arm (libcair.so)
  python bench.py --impl reference ...                       reference arm: the CPU restatement of the
                                                             reference forward (oracle/) on host cores
Prints ONE JSON line (rank 0).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

CFG = dict(model='match_tensor', emsize=300, src_vocab_size=131072, dropout_emb=0.2, rnn_type='LSTM',
           bidirection=True, nlayers=1, dropout_rnn=0.2, featsize=40, nhid_query=128, nhid_doc=128,
           nchannels=50, nfilters=6, match_filter_size=20)
B, N, LQ, LD = 128, 10, 20, 200
WORKLOAD = 'match_tensor cfg2: B=128 N=10 Lq=20 Ld=200 E=300 H=128(64/dir) F=40 C=50 V=131072, full lengths'
METRIC = 'query-doc pairs scored/sec (Match-Tensor, Lq20/Ld200/N10)'
UNIT = 'pairs/s'


def flops_per_pair():
    """Reference-equivalent (direct form) and executed (factorised conv) MFLOP per pair, SURVEY.md 8(d)."""
    E, F, h, C_, nf, M = 300, 40, 64, 50, 6, 20
    proj = 2 * LD * E * F
    lstm = 2 * LD * 2 * 4 * h * (F + h)
    dproj = 2 * LD * 2 * h * C_
    conv_direct = 2 * LQ * LD * (C_ + 1) * nf * (9 + 15 + 21)
    conv_fact = 2 * LQ * LD * (7 * C_) * (3 * nf)
    conv1 = 2 * LQ * LD * 3 * nf * M
    return dict(ref=(proj + lstm + dproj + conv_direct + conv1) / 1e6, interact_ref=(conv_direct + conv1) / 1e6,
                interact_exec=(conv_fact + conv1) / 1e6, lstm=lstm / 1e6)


def bytes_per_pair_folded():
    """Compulsory HBM bytes per pair with the folded [V,F] fp32 table: ids int64."""
    F = 40
    return LD * (8 + F * 4) + (LQ * (8 + F * 4) + 8) / N + 8 + 4


class ClockSampler(threading.Thread):
    """SM clock + throttle reasons sampled DURING the timed regions: NVML in-process (a query takes well under a
    millisecond; nvidia-smi as a subprocess takes longer than the whole timed region), nvidia-smi as the fallback."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag, self.active, self.source = index, [], False, False, 'nvml'
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML enumerates physical devices: map through CUDA_VISIBLE_DEVICES when it is a plain index list
            vis = os.environ.get('CUDA_VISIBLE_DEVICES')
            phys = index
            if vis:
                try:
                    phys = int(vis.split(',')[index])
                except Exception:
                    phys = index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nvml = pynvml
        except Exception:
            self.source = 'nvidia-smi'

    def _sample_nvml(self):
        n = self.nvml
        sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
        mx = n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)
        try:
            r = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
        except Exception:
            r = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
        flags = [bool(r & n.nvmlClocksThrottleReasonHwSlowdown), bool(r & n.nvmlClocksThrottleReasonHwThermalSlowdown),
                 bool(r & n.nvmlClocksThrottleReasonSwThermalSlowdown), bool(r & n.nvmlClocksThrottleReasonSwPowerCap)]
        return [sm, mx] + flags

    def _sample_smi(self):
        q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')
        out = subprocess.run(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + q, '--format=csv,noheader,nounits'],
                             capture_output=True, text=True, timeout=5).stdout
        f = [x.strip() for x in out.strip().split(',')]
        return [float(f[0]), float(f[1])] + [x.lower().startswith('active') for x in f[2:6]]

    def run(self):
        while not self.stop_flag:
            try:
                row = self._sample_nvml() if self.nvml else self._sample_smi()
                if self.active or not self.nvml:
                    self.rows.append(row)
            except Exception:
                pass
            time.sleep(0.001 if self.nvml else 0.05)

    def summary(self):
        if not self.rows:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['no clock samples'], source=self.source)
        sm = sorted(float(r[0]) for r in self.rows)
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for i, n in enumerate(names) if any(r[2 + i] for r in self.rows)]
        return dict(sm_mhz=sm[len(sm) // 2], sm_min_mhz=sm[0], sm_max_mhz=float(self.rows[0][1]), reasons=reasons,
                    samples=len(self.rows), source=self.source + ' (sampled only while a timed region is running)')


def make_batch(seed):
    from context_attentive_ir_b200 import synth
    return synth.ranker_batch(seed, B, N, LQ, LD, CFG['src_vocab_size'], variable=False)


def cpu_oracle_rate(sample_queries, threads=None):
    """pairs/s of the C restatement of the reference forward on a bounded sample of the workload."""
    import torch
    import helpers
    import oracle_lib as ol
    torch.manual_seed(1013)
    small = dict(CFG)
    net = helpers.build_module(small)
    sd = helpers.state_dict_numpy(net)
    batch = make_batch(1236)
    sl = slice(0, sample_queries)
    cores = threads or os.cpu_count()
    os.environ['OMP_NUM_THREADS'] = str(cores)
    cores = int(ol.lib().cair_oracle_set_threads(int(cores)))   # torchrun exports OMP_NUM_THREADS=1; set it explicitly
    t0 = time.perf_counter()
    ol.run_ranker(small, sd, batch['q'][sl], batch['qlen'][sl], batch['d'][sl], batch['dlen'][sl])
    dt = time.perf_counter() - t0
    return sample_queries * N / dt, dt, cores


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cores = os.cpu_count()
    # size the per-step sample from a 1-query probe so the whole run stays within a few minutes
    rate1, dt1, _ = cpu_oracle_rate(1)
    per_query = dt1
    budget = 150.0 / max(1, args.steps + args.warmup)
    sample = int(max(1, min(B, budget / max(per_query / max(1, min(cores, N)), 1e-3) / 4)))
    times = []
    for i in range(args.warmup + args.steps):
        rate, dt, _ = cpu_oracle_rate(sample)
        if i >= args.warmup:
            times.append(dt)
    ms = 1e3 * float(np.mean(times))
    value = sample * N / (ms / 1e3)
    sample_desc = '%d of %d queries x %d docs of the workload per step (C port of the reference forward, OpenMP)' % (sample, B, N)
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic', 'config': {'workload': WORKLOAD, 'sample': sample_desc},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample_desc},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0}))


def run_product(args):
    import torch
    import torch.distributed as dist
    import helpers
    from context_attentive_ir_b200 import lib
    from context_attentive_ir_b200.parallel import gather_scores

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)

    torch.manual_seed(1013)  # identical replicated weights on every rank
    net = helpers.build_module(CFG).to(dev)
    batch = make_batch(1236 + rank)  # this rank's contiguous slice of the global B*world queries
    q, ql, d, dl = helpers.to_dev(batch, dev)
    hq, hql, hd, hdl = [torch.from_numpy(np.ascontiguousarray(batch[k])).pin_memory() for k in ('q', 'qlen', 'd', 'dlen')]
    hout = torch.empty(B, N, dtype=torch.float32).pin_memory()
    flush = torch.empty(192 * 1024 * 1024, dtype=torch.uint8, device=dev)  # 1.5 x the 126 MB L2
    pairs_local, pairs_total = B * N, B * N * world
    stream = torch.cuda.current_stream(dev)

    def step():
        with torch.no_grad():
            s = net(q, ql, d, dl)
            if world > 1:
                s = gather_scores(s.reshape(-1), pairs_total)
        return s

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    h = net._cair_handle
    L = lib.load()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # ---- device-resident timing: K steps, L2 flushed between steps, CUDA events per step ----
    # The host only enqueues inside the timed region (no per-step synchronisation), so the GPU never waits for
    # Python; the library's per-stage CUDA events stay on and are read once after the loop (they are re-recorded by
    # every step, so that read gives the final timed step; a few more profiled steps follow for the stage averages).
    lib.check(L.cair_profile_enable(h, 1))
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    stage_ms = {}

    def read_stages():
        names = C.create_string_buffer(1024)
        ms = (C.c_float * 32)()
        cnt = C.c_int32()
        lib.check(L.cair_profile_read(h, names, 1024, ms, 32, C.byref(cnt)))
        for nm, v in zip(names.value.decode().split(','), list(ms)[:cnt.value]):
            stage_ms.setdefault(nm, []).append(v)

    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    n0 = lib.launch_count()
    sampler.active = True
    for i in range(args.steps):
        flush.fill_(i & 0xff)
        ev[i][0].record(stream)
        step()
        ev[i][1].record(stream)
    torch.cuda.synchronize()
    sampler.active = False
    launches = lib.launch_count() - n0
    if world > 1:
        dist.barrier()
    read_stages()
    for i in range(5):   # stage averages (outside the timed region, same step, L2 flushed)
        flush.fill_(i & 0xff)
        step()
        torch.cuda.synchronize()
        read_stages()
    lib.check(L.cair_profile_enable(h, 0))
    total_ms = sum(a.elapsed_time(b) for a, b in ev)
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = pairs_total / (ms_per_step / 1e3)

    # ---- end-to-end: host ids -> H2D -> score -> D2H scores through the C-ABI host entry points ----
    # A serving loop over three staging slots (submit_host / wait_host): every step copies its ids from pinned host
    # memory, scores them and copies the scores back; up to three steps are in flight (copies of step k+1 and its
    # document encoder overlap the interaction kernel of step k).  The L2 flush of every step is enqueued ahead of the
    # step INSIDE the timed region (it costs ~40 us of GPU time per step).
    houts = [torch.empty(B, N, dtype=torch.float32).pin_memory() for _ in range(3)]
    cstream = torch.cuda.Stream(dev)

    def e2e_loop(k):
        with torch.cuda.stream(cstream):
            for i in range(k):
                if i >= 3:
                    net.wait_host(i % 3)     # scores of step i-3 are on the host; its slot is free again
                flush.fill_(i & 0xff)
                net.submit_host(hq, hql, hd, hdl, out=houts[i % 3], slot=i % 3, device=dev, stream=cstream)
            for i in range(max(0, k - 3), k):
                net.wait_host(i % 3)

    e2e_loop(7)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler.active = True
    t0 = time.perf_counter()
    e2e_loop(args.steps)
    e2e_total = time.perf_counter() - t0
    sampler.active = False
    # the synchronous single-call form (forward_host: one cached CUDA graph per call, returns after the D2H)
    for _ in range(3):
        net.forward_host(hq, hql, hd, hdl, out=hout, device=dev)
    sync_t = []
    for i in range(args.steps):
        flush.fill_(i & 0xff)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        net.forward_host(hq, hql, hd, hdl, out=hout, device=dev)
        sync_t.append(time.perf_counter() - t1)
    e = torch.tensor([e2e_total, sum(sync_t)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e, op=dist.ReduceOp.MAX)
    e2e_value = pairs_total / (float(e[0].item()) / args.steps)
    e2e_sync_value = pairs_total / (float(e[1].item()) / args.steps)
    h2d = (hq.numel() + hql.numel() + hd.numel() + hdl.numel()) * 8 * world
    d2h = hout.numel() * 4 * world
    if rank == 0:
        sampler.stop_flag = True
        sampler.join(timeout=2)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        except Exception:
            pass
        hbm_peak = peaks.get('hbm_gbs', 6650.0)
        tf_peak = peaks.get('bf16_tflops_sustained', 1400.0)
        peak_src = 'measured (MEASURED_PEAKS.json, sustained bf16)' if peaks else 'fallback (B200_PROFILING.md)'
        stages = {k: float(np.mean(v)) for k, v in stage_ms.items()}
        fl = flops_per_pair()
        timed = {k: v for k, v in stages.items() if k not in ('begin', 'join_query_side')}
        top = max(timed, key=timed.get) if timed else None
        roof = None
        if top == 'interact':
            # executed = factorised conv on tcgen05 (bf16x3 => 3 MMA passes per logical FLOP, padded tiles not counted)
            ach = fl['interact_exec'] * 1e6 * pairs_local / (stages[top] / 1e3) / 1e12
            roof = dict(kernel='mt_tc_interact_kernel', bound='tensor', achieved=ach, peak=tf_peak, unit='TFLOP/s',
                        frac=ach / tf_peak, traffic=None, achieved_ref_equiv=fl['interact_ref'] / fl['interact_exec'] * ach,
                        achieved_issued_bf16=3 * ach, ms_per_launch=stages[top], peak_source=peak_src)
        elif top in ('doc_recurrence', 'lstm_recurrence'):
            # algorithmic FLOPs of the doc BiLSTM (input + recurrent projection, both directions); the kernel issues
            # 3 bf16 MMA passes per logical FLOP and is bound by the 200-step recurrence latency, not by the pipe
            ach = 2 * LD * 2 * 4 * 64 * (40 + 64) * pairs_local / (stages[top] / 1e3) / 1e12
            roof = dict(kernel='lstm_tc_kernel', bound='tensor', achieved=ach, peak=tf_peak, unit='TFLOP/s',
                        frac=ach / tf_peak, traffic=None, achieved_issued_bf16=3 * ach, ms_per_launch=stages[top],
                        steps_per_launch=LD, us_per_recurrence_step=1e3 * stages[top] / LD, peak_source=peak_src)
        elif top is not None:
            ach = bytes_per_pair_folded() * pairs_local / (stages[top] / 1e3) / 1e9
            roof = dict(kernel=top, bound='hbm', achieved=ach, peak=hbm_peak, unit='GB/s', frac=ach / hbm_peak,
                        traffic=None, ms_per_launch=stages[top], peak_source=peak_src)
        try:
            tr = json.load(open(os.path.join(ROOT, 'profiles', 'traffic.json')))
            if roof and roof['kernel'] in tr:
                roof['traffic'] = tr[roof['kernel']]['bytes']
                roof['traffic_source'] = tr[roof['kernel']]['capture']
        except Exception:
            pass
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            # bounded sample: whole batches of the workload until >= cpu-sample-seconds of CPU work (about 10-30 s)
            nq_done, t_done, cores = 0, 0.0, os.cpu_count()
            while t_done < args.cpu_sample_seconds:
                rate, dt, cores = cpu_oracle_rate(args.cpu_sample_queries)
                nq_done += args.cpu_sample_queries
                t_done += dt
            cpu = dict(value=nq_done * N / t_done, unit=UNIT, cores=cores, kind='port',
                       sample='%d queries x %d docs of the workload, in chunks of %d queries (%.1f s of the C port of the '
                              'reference forward, OpenMP over %d threads)' % (nq_done, N, args.cpu_sample_queries, t_done, cores))
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3),
            'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32 (tcgen05 bf16x3 split-precision MMA, fp32 accumulate/state)', 'data': 'synthetic',
            'config': {'workload': WORKLOAD, 'per_gpu_pairs': pairs_local, 'global_pairs': pairs_total,
                       'parallelism': 'doc-parallel x%d, one all-gather of scores' % world,
                       'l2': 'flushed between steps (192 MiB fill = 1.5 x L2, outside the per-step events)',
                       'table': 'eval-mode folded [V,40] fp32 table (built once at handle creation)'},
            'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                    'api': 'submit_host/wait_host over 3 staging slots: H2D of step k+1 and its document encoder overlap the '
                           'interaction kernel of step k; L2 flush inside the timed region',
                    'single_call_value': e2e_sync_value,
                    'single_call_api': 'forward_host: H2D + kernels + D2H as one cached CUDA graph, synchronous'},
            'gpu_launches': int(launches), 'clocks': sampler.summary(),
            'roofline': roof, 'stages_ms': stages,
            'flops_per_pair_mflop': fl, 'hbm_bytes_per_pair': bytes_per_pair_folded(),
            'hbm_frac_end_to_end': bytes_per_pair_folded() * value / world / 1e9 / hbm_peak,
        }
        if cpu is not None:
            line['cpu_baseline'] = cpu
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=100)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--cpu-sample-queries', type=int, default=32)
    ap.add_argument('--cpu-sample-seconds', type=float, default=12.0)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_product(args)


if __name__ == '__main__':
    main()
