#!/usr/bin/env python
"""bench.py - Match-Tensor scoring throughput (BASELINE.json configs[1]) on N B200s, plus the other BASELINE configs.

A "step" = BATCHES_PER_STEP passes of the scoring hot path, each over one DISTINCT pre-staged batch of synthetic
sessions: B=128 queries x N=10 candidate docs per GPU, Lq=20, Ld=200, E=300, H=128 (64/dir), F=40, C=50, all lengths
at max (the headline throughput set), V=131072, random-init weights.  Weak scaling: every rank scores
its own batches (doc-parallel, weights replicated) and one all-gather of the fp32 scores follows every batch inside
the timed step.

  python bench.py [--gpus N] [--steps K] [--warmup W]        product arm (libcair.so)
  python bench.py --impl reference ...                       reference arm: the CPU restatement of the reference
                                                             forward (oracle/) on the host cores, bounded sample
Prints ONE JSON line (rank 0).  `other_configs` in that line carries ESM cfg1, DRMM cfg3, DUET cfg5 (N = 10, 50, 100,
500) and CARS cfg4, each measured the same way (device-resident, L2 flushed between steps, collective included).
"""
import argparse
import ctypes as C
import hashlib
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
BATCHES_PER_STEP = 25
WORKLOAD = 'match_tensor cfg2: B=128 N=10 Lq=20 Ld=200 E=300 H=128(64/dir) F=40 C=50 V=131072, full lengths'
METRIC = 'query-doc pairs scored/sec (Match-Tensor, Lq20/Ld200/N10)'
UNIT = 'pairs/s'
# identical in both arms (the driver compares the dicts); everything arm-specific lives under "run"
CONFIG = {'workload': WORKLOAD, 'B': B, 'N': N, 'Lq': LQ, 'Ld': LD, 'batches_per_step': BATCHES_PER_STEP,
          'l2': 'inputs of a step (25 distinct batches, working set ~350 MB each) exceed L2; L2 also flushed between steps'}

CARS_CFG = dict(model='cars', emsize=300, src_vocab_size=131072, tgt_vocab_size=50, dropout_emb=0.2, dropout=0.2, rnn_type='LSTM',
                bidirection=True, nlayers=1, nhid_query=256, nhid_document=256, nhid_click=512, nhid_session_query=512,
                nhid_session_document=512, nhid_decoder=512, query_session_off=False, doc_session_off=False, dropout_rnn=0.2,
                attn_type='general', mlp_nhid=150, pool_type='attn', regularize_coeff=0.1, alpha=0.1, lambda1=0.01,
                lambda2=0.0001, turn_ranker_off=False, turn_recommender_off=False)
DUET_CFG = dict(model='duet', emsize=300, src_vocab_size=131072, dropout_emb=0.2, dropout=0.2, use_word=True, nfilters=300,
                local_filter_size=1, dist_filter_size=3, pool_size=5, max_doc_len=200, max_query_len=20)


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
    """HBM bytes per pair actually compulsory on the product path: an int64 id + the 192-byte pre-split bf16 hi/lo row of
    the folded table per token (48 K slots x 2 x 2 B; SURVEY.md 8d says: count what is moved), + length + score."""
    row = 8 + 192
    return LD * row + (LQ * row + 8) / N + 8 + 4


def bpp_fp32(E, Lq, Ld, n):   # int64 ids, fp32 table rows, one fp32 score
    return Ld * (8 + E * 4) + (Lq * (8 + E * 4) + 8) / n + 12


class ClockSampler(threading.Thread):
    """SM clock + throttle reasons sampled DURING the timed regions: NVML in-process (a query takes well under a
    millisecond), nvidia-smi as the fallback."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag, self.active, self.source = index, [], False, False, 'nvml'
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
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
            time.sleep(0.002 if self.nvml else 0.05)

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


# ---- CPU arm (the only place besides tests/ and smoke() that touches oracle/) -------------------------------------
def cpu_oracle_rate(sample_queries, threads=None):
    """pairs/s of the C restatement of the reference forward on a bounded sample of the workload."""
    import torch
    import helpers
    import oracle_lib as ol
    torch.manual_seed(1013)
    net = helpers.build_module(dict(CFG))
    sd = helpers.state_dict_numpy(net)
    batch = make_batch(1236)
    sl = slice(0, sample_queries)
    cores = threads or os.cpu_count()
    os.environ['OMP_NUM_THREADS'] = str(cores)
    cores = int(ol.lib().cair_oracle_set_threads(int(cores)))   # torchrun exports OMP_NUM_THREADS=1; set it explicitly
    t0 = time.perf_counter()
    ol.run_ranker(dict(CFG), sd, batch['q'][sl], batch['qlen'][sl], batch['d'][sl], batch['dlen'][sl])
    dt = time.perf_counter() - t0
    return sample_queries * N / dt, dt, cores


PORT_NOTE = ('C restatement of the reference forward (fp32 values, double accumulation, OpenMP); the reference itself is pure '
             'Python and cannot travel to the GPU box. SURVEY.md section 6 measured the real reference (torch CPU) at 347 pairs/s '
             'on 8 vCPU: the port is about 2x slower per core, so ratios against it overstate the speed-up by about that factor')


def cpu_baseline_worker(args):
    """Runs in a SUBPROCESS of the product arm so that the product process never loads oracle/libcair_oracle.so."""
    nq_done, t_done, cores = 0, 0.0, os.cpu_count()
    while t_done < args.cpu_sample_seconds:
        rate, dt, cores = cpu_oracle_rate(args.cpu_sample_queries)
        nq_done += args.cpu_sample_queries
        t_done += dt
    print(json.dumps(dict(value=nq_done * N / t_done, unit=UNIT, cores=cores, kind='port',
                          sample='%d queries x %d docs of the workload, in chunks of %d queries (%.1f s of the C port of the '
                                 'reference forward, OpenMP over %d threads, separate process)'
                                 % (nq_done, N, args.cpu_sample_queries, t_done, cores), note=PORT_NOTE)))


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cores = os.cpu_count()
    rate1, dt1, _ = cpu_oracle_rate(1)   # size the per-step sample from a 1-query probe: the whole run stays within minutes
    budget = 150.0 / max(1, args.steps + args.warmup)
    sample = int(max(1, min(B, budget / max(dt1 / max(1, min(cores, N)), 1e-3) / 4)))
    times = []
    for i in range(args.warmup + args.steps):
        rate, dt, _ = cpu_oracle_rate(sample)
        if i >= args.warmup:
            times.append(dt)
    ms = 1e3 * float(np.mean(times))
    value = sample * N / (ms / 1e3)
    sample_desc = '%d of %d queries x %d docs of one batch of the workload per step (C port of the reference forward, OpenMP)' % (sample, B, N)
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic', 'config': CONFIG,
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample_desc, 'note': PORT_NOTE},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0}))


# ---- product arm ---------------------------------------------------------------------------------------------------
def _sha16(path):
    try:
        return hashlib.sha256(open(path, 'rb').read()).hexdigest()[:16]
    except Exception:
        return None


def traffic_for(kernel):
    """dram bytes per launch from the committed ncu capture of `kernel` - only when the kernel source is byte-identical to
    the one that was profiled (profiles/traffic.json records its sha); otherwise None rather than a stale constant."""
    try:
        tr = json.load(open(os.path.join(ROOT, 'profiles', 'traffic.json')))
        e = tr.get(kernel)
        if e and e.get('source_sha16') and e['source_sha16'] == _sha16(os.path.join(ROOT, e['source'])):
            return e['bytes'], e['capture']
    except Exception:
        pass
    return None, None


class Runner:
    """Shared timing harness: device-resident inputs, L2 flush before every step (outside the events), CUDA events per
    step on the launching stream, max over ranks."""

    def make_gather(self, per):
        """(gather(local, total) -> all scores, description): our NVLink peer-memory kernel (cair_allgather_scores) when the
        symmetric-memory rendezvous works, else NCCL."""
        from context_attentive_ir_b200.parallel import P2PScoreGather, gather_scores
        if self.world == 1:
            return gather_scores, 'none (1 GPU)'
        if not self.no_p2p:
            try:
                return P2PScoreGather(per, self.dev), 'cair_allgather_scores: peer stores over NVLink + flags (torch symmetric memory mappings)'
            except Exception as ex:   # noqa: BLE001 - any rendezvous problem: fall back to the library collective
                self.p2p_error = repr(ex)[:200]
        return gather_scores, 'NCCL all_gather_into_tensor'

    def __init__(self, dev, world, rank, no_p2p=False):
        import torch
        self.no_p2p, self.p2p_error = no_p2p, None
        self.torch, self.dev, self.world, self.rank = torch, dev, world, rank
        self.flush = torch.empty(192 * 1024 * 1024, dtype=torch.uint8, device=dev)   # 1.5 x the 126 MB L2
        self.stream = torch.cuda.current_stream(dev)

    def time_steps(self, step, steps, warmup):
        """Returns (ms per step as the max over ranks, libcair kernel launches inside the timed steps)."""
        import torch.distributed as dist
        from context_attentive_ir_b200 import lib
        torch = self.torch
        for _ in range(max(warmup, 3)):
            step()
        torch.cuda.synchronize()
        if self.world > 1:
            dist.barrier()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        n0 = lib.launch_count()
        for i in range(steps):
            self.flush.fill_(i & 0xff)
            ev[i][0].record(self.stream)
            step()
            ev[i][1].record(self.stream)
        torch.cuda.synchronize()
        launches = lib.launch_count() - n0
        if self.world > 1:
            dist.barrier()
        t = torch.tensor([sum(a.elapsed_time(b) for a, b in ev)], dtype=torch.float64, device=self.dev)
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / steps, int(launches)


def read_stages(h, acc):
    from context_attentive_ir_b200 import lib
    L = lib.load()
    names = C.create_string_buffer(2048)
    ms = (C.c_float * 64)()
    cnt = C.c_int32()
    lib.check(L.cair_profile_read(h, names, 2048, ms, 64, C.byref(cnt)))
    seen = set()
    for nm, v in zip(names.value.decode().split(','), list(ms)[:cnt.value]):
        k = nm
        while k in seen:
            k += "'"
        seen.add(k)
        acc.setdefault(k, []).append(v)


def stage_profile(run, net, once, reps=5):
    """Per-stage CUDA-event times (ms) of the library's own profiler over a few single-batch, L2-flushed steps."""
    from context_attentive_ir_b200 import lib
    L = lib.load()
    h = net.__dict__['_cair_handle']
    acc = {}
    lib.check(L.cair_profile_enable(h, 1))
    for i in range(reps):
        run.flush.fill_(i & 0xff)
        once()
        run.torch.cuda.synchronize()
        read_stages(h, acc)
    lib.check(L.cair_profile_enable(h, 0))
    return {k: float(np.mean(v)) for k, v in acc.items()}


def other_configs(run, args, peaks):
    """ESM cfg1, DRMM cfg3, DUET cfg5 (N sweep) and CARS cfg4, weak scaling with the score all-gather inside the step."""
    import torch
    import helpers
    from context_attentive_ir_b200 import synth
    from context_attentive_ir_b200.parallel import gather_scores as nccl_gather
    import torch.distributed as dist
    dev, world, rank = run.dev, run.world, run.rank

    def make_gather(per):
        return run.make_gather(per)[0]
    hbm_peak = peaks.get('hbm_gbs', 6650.0)
    tf_peak = peaks.get('bf16_tflops_sustained', 1400.0)
    out = []

    def ranker_case(name, cfg, b, n, lq, ld, steps, roof, **batch_kw):
        torch.manual_seed(1013)
        net = helpers.build_module(cfg).to(dev)
        batch = synth.ranker_batch(1234 + rank, b, n, lq, ld, cfg['src_vocab_size'], **(batch_kw or dict(variable=False)))
        t = helpers.to_dev(batch, dev)
        total = b * n * world
        gather_scores = make_gather(b * n)

        def step():
            with torch.no_grad():
                s = net(*t)
                if world > 1:
                    s = gather_scores(s.reshape(-1), total)
            return s
        ms, _ = run.time_steps(step, steps, 3)
        st = stage_profile(run, net, step, 3)
        e = dict(name=name, config=dict(model=cfg['model'], B=b, N=n, Lq=lq, Ld=ld, E=cfg['emsize'], V=cfg['src_vocab_size']),
                 pairs_per_s=total / (ms / 1e3), ms_per_step=ms, n_gpus=world, scaling='weak',
                 parallelism='doc-parallel x%d: %d pairs per GPU, one all-gather of scores per step' % (world, b * n),
                 steps=steps, stages_ms={k: round(v, 4) for k, v in st.items() if k != 'begin'})
        e['roofline'] = roof(b * n, ms, st)
        out.append(e)
        del net
        torch.cuda.empty_cache()

    def hbm_roof(kernel, bytes_per_pair, captured=False):
        def f(pairs, ms, st):
            ach = bytes_per_pair * pairs / (ms / 1e3) / 1e9
            tr, src = traffic_for(kernel) if captured else (None, None)   # only where the committed capture is of THIS workload
            return dict(kernel=kernel, bound='hbm', achieved=ach, peak=hbm_peak, unit='GB/s', frac=ach / hbm_peak,
                        traffic=tr, bytes_per_pair=bytes_per_pair, note='whole step (one fused kernel chain) against the measured copy bandwidth')
        return f

    def tensor_roof(kernel, mflop_ref_per_pair, captured=False):
        def f(pairs, ms, st):
            ach = mflop_ref_per_pair * 1e6 * pairs / (ms / 1e3) / 1e12
            tr, src = traffic_for(kernel) if captured else (None, None)   # only where the committed capture is of THIS workload
            cand = [k for k in st if k not in ('begin', 'end')]
            top = max(cand, key=lambda k: st[k]) if cand else None
            return dict(kernel=kernel, bound='tensor', achieved=ach, peak=tf_peak, unit='TFLOP/s', frac=ach / tf_peak, traffic=tr,
                        mflop_ref_per_pair=mflop_ref_per_pair, top_stage=top,
                        note='reference-equivalent FLOPs of the whole step against sustained bf16 (bf16x3 issues 3x that)')
        return f

    # the headline model on the dataset's real length distribution (avg document 63 tokens + BOS / EOS): packed-sequence
    # semantics retire short documents early; reported beside the full-length headline, not instead of it
    ranker_case('match_tensor cfg2, realistic lengths (avg doc 63 + BOS/EOS)', CFG, B, N, LQ, LD, 20,
                tensor_roof('lstm_tc_kernel', flops_per_pair()['ref']), realistic=True)
    # the reference's STOCK Match-Tensor hidden sizes (neuroir/hyparam.py:88-100: nhid_query 30, nhid_doc 140 = 15 / 70 per
    # direction): the 70-wide document encoder is beyond the single-CTA tcgen05 kernel and runs on the cluster-split one
    ranker_case('match_tensor, stock hidden sizes 30/140 (SURVEY 8d)', dict(CFG, nhid_query=30, nhid_doc=140), B, N, LQ, LD, 20,
                tensor_roof('rnn_tc_kernel', 145.3))   # conv 113.1 + projections 7.6 + BiLSTM 24.6 MFLOP per pair (reference-equivalent)
    ranker_case('esm cfg1', dict(model='esm', emsize=64, src_vocab_size=10000), 8, 5, 10, 50, 20,
                hbm_roof('esm_kernel', bpp_fp32(64, 10, 50, 5)))
    ranker_case('esm E=300 (cfg3 shape)', dict(model='esm', emsize=300, src_vocab_size=131072), 256, 10, 20, 200, 10,
                hbm_roof('esm_kernel', bpp_fp32(300, 20, 200, 10)))
    ranker_case('drmm cfg3', dict(model='drmm', emsize=300, src_vocab_size=131072, dropout_emb=0.2, nbins=5), 256, 10, 20, 200, 10,
                hbm_roof('drmm_tc_kernel', bpp_fp32(300, 20, 200, 10), captured=True))
    for n in (10, 50, 100, 500):
        ranker_case('duet cfg5 N=%d' % n, DUET_CFG, 32, n, 20, 200, 10 if n <= 100 else 4, tensor_roof('gemm_tc_kernel', 145.8))

    # CARS cfg4: B = 32 sessions per GPU; the global batch (32 x world sessions, labels replicated - the click-mask width
    # is a batch-global max) is sharded by session, one all-gather of the scores
    torch.manual_seed(1013)
    net = helpers.build_module(CARS_CFG).to(dev)
    Bc, S, Nc, Lq, Ld = 32, 7, 10, 20, 200
    batch = synth.session_batch(1238, Bc * world, S, Nc, Lq, Ld, CARS_CFG['src_vocab_size'], variable=False, max_clicks=2)
    t = helpers.to_dev(batch, dev, ('q', 'qlen', 'd', 'dlen', 'label'))
    total = Bc * world * S * Nc
    gather_scores = make_gather(Bc * S * Nc)

    def cars_step():
        with torch.no_grad():
            s = net.score(*t, session_slice=(rank * Bc, Bc))['scores']
            if world > 1:
                s = gather_scores(s[rank * Bc:(rank + 1) * Bc].reshape(-1), total)
        return s
    ms, _ = run.time_steps(cars_step, 10, 3)
    st = stage_profile(run, net, cars_step, 3)
    e = dict(name='cars cfg4 (ranking path)', config=dict(model='cars', B=Bc, S=S, N=Nc, Lq=Lq, Ld=Ld, E=300, H=256),
             pairs_per_s=total / (ms / 1e3), ms_per_step=ms, n_gpus=world, scaling='weak',
             parallelism='session-parallel x%d: %d sessions per GPU, one all-gather of scores per step' % (world, Bc),
             steps=10, stages_ms={k: round(v, 4) for k, v in st.items() if k != 'begin'})
    e['roofline'] = tensor_roof('rnn_tc_kernel', 205.0, captured=True)(Bc * S * Nc, ms, st)
    out.append(e)
    del net
    torch.cuda.empty_cache()

    # MNSRF ranking path at the cfg4 shape, sharded by session like CARS
    torch.manual_seed(1013)
    mcfg = dict(model='mnsrf', emsize=300, src_vocab_size=CARS_CFG['src_vocab_size'], tgt_vocab_size=50, dropout_emb=0.2, dropout=0.2,
                rnn_type='LSTM', bidirection=True, nlayers=1, nhid_query=256, nhid_document=256, nhid_session=512, dropout_rnn=0.2,
                regularize_coeff=0.1)
    net = helpers.build_module(mcfg).to(dev)

    def mnsrf_step():
        with torch.no_grad():
            s = net.score(*t[:4], session_slice=(rank * Bc, Bc))['scores']
            if world > 1:
                s = gather_scores(s[rank * Bc:(rank + 1) * Bc].reshape(-1), total)
        return s
    ms, _ = run.time_steps(mnsrf_step, 10, 3)
    out.append(dict(name='mnsrf, cfg4 shape (ranking path)', config=dict(model='mnsrf', B=Bc, S=S, N=Nc, Lq=Lq, Ld=Ld, E=300, H=256),
                    pairs_per_s=total / (ms / 1e3), ms_per_step=ms, n_gpus=world, scaling='weak', steps=10,
                    parallelism='session-parallel x%d: %d sessions per GPU, one all-gather of scores per step' % (world, Bc),
                    roofline=tensor_roof('rnn_tc_kernel', 197.0)(Bc * S * Nc, ms, {})))
    del net
    torch.cuda.empty_cache()

    # Match-Tensor TRAINING step at the cfg2 shape: the statement order of the reference's Ranker.update
    # (forward, BCEWithLogitsLoss, zero_grad, backward, clip_grad_norm, SGD step); data-parallel replicas, gradients
    # all-reduced by NCCL at N > 1 (what nn.DataParallel's backward does in the reference)
    torch.manual_seed(1013)
    tnet = helpers.build_module(CFG).to(dev).train()
    tb = synth.ranker_batch(4321 + rank, B, N, LQ, LD, CFG['src_vocab_size'], variable=False)
    tq = helpers.to_dev(tb, dev)
    labels = torch.from_numpy(tb['label']).float().to(dev)
    params = [p for p in tnet.parameters() if p.requires_grad]
    opt = torch.optim.SGD(params, 0.05)
    crit = torch.nn.BCEWithLogitsLoss()

    def train_step():
        loss = crit(tnet(*tq), labels)
        opt.zero_grad()
        loss.backward()
        if world > 1:
            for p in params:
                dist.all_reduce(p.grad)
                p.grad.div_(world)
        torch.nn.utils.clip_grad_norm_(params, 5.0)
        opt.step()
        return loss
    ms, _ = run.time_steps(train_step, 6, 3)
    out.append(dict(name='match_tensor cfg2 TRAINING step (forward + backward + clip + SGD)', config=dict(model='match_tensor', B=B, N=N, Lq=LQ, Ld=LD, dropout_emb=CFG['dropout_emb']),
                    pairs_per_s=B * N * world / (ms / 1e3), ms_per_step=ms, n_gpus=world, scaling='weak', steps=6,
                    parallelism='data-parallel x%d: %d pairs per GPU, gradient all-reduce per step' % (world, B * N),
                    roofline=tensor_roof('lstm_train_bwd_kernel / mt_train_interact_bwd_kernel / gemm_tn_kernel', 3 * flops_per_pair()['ref'])(B * N, ms, {}),
                    note='training row (csrc/train.cu): forward on the tcgen05 kernels (interaction with arg-max, gate-saving lstm_tc, gemm_tc dense layers); BPTT / weight-gradient GEMMs / sparse interaction backward on fp32 CUDA cores'))
    del tnet, opt
    torch.cuda.empty_cache()
    return out


def run_product(args):
    import torch
    import torch.distributed as dist
    import helpers
    from context_attentive_ir_b200 import lib
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    run = Runner(dev, world, rank, args.nccl_collective)

    torch.manual_seed(1013)  # identical replicated weights on every rank
    net = helpers.build_module(CFG).to(dev)
    nb = BATCHES_PER_STEP
    # this rank's batches of the global (B * world queries) x nb workload
    host = [make_batch(1236 + rank * 1000 + i) for i in range(nb)]
    devb = [helpers.to_dev(b, dev) for b in host]
    pinned = [[torch.from_numpy(np.ascontiguousarray(b[k])).pin_memory() for k in ('q', 'qlen', 'd', 'dlen')] for b in host]
    pairs_local, pairs_total = B * N, B * N * world
    stream = run.stream
    gather_scores, collective = run.make_gather(pairs_local)

    def one(i):
        with torch.no_grad():
            s = net(*devb[i])
            if world > 1:
                s = gather_scores(s.reshape(-1), pairs_total)
        return s

    def step():
        for i in range(nb):
            one(i)

    step()
    torch.cuda.synchronize()
    h = net._cair_handle
    L = lib.load()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # ---- device-resident timing: K steps x 25 distinct batches, L2 flushed between steps, CUDA events per step ----
    sampler.active = True
    ms_per_step, launches = run.time_steps(step, args.steps, args.warmup)
    sampler.active = False
    value = pairs_total * nb / (ms_per_step / 1e3)
    stages = stage_profile(run, net, lambda: one(0), 5)

    # ---- end-to-end: host ids -> H2D -> score (-> all-gather) -> D2H scores, every step over its 25 pinned host batches ----
    houts = [torch.empty(B * world, N, dtype=torch.float32).pin_memory() for _ in range(3)]
    cstream = torch.cuda.Stream(dev)
    pipe_gather = None
    if world > 1 and not args.nccl_collective:
        try:   # the all-gather inside the serving pipeline (cair_ranker_set_gather) needs its own peer buffers
            from context_attentive_ir_b200.parallel import P2PScoreGather
            pipe_gather = P2PScoreGather(pairs_local, dev).attach(net)
        except Exception as ex:   # noqa: BLE001
            run.p2p_error = run.p2p_error or repr(ex)[:200]
    if world == 1 or pipe_gather is not None:
        # serving loop over three staging slots of the C ABI (cair_ranker_submit_host / cair_ranker_wait_host): the copies of
        # batch k+1 and its document encoder overlap the interaction kernel of batch k
        def e2e_loop(k):
            with torch.cuda.stream(cstream):
                for i in range(k):
                    if i >= 3:
                        net.wait_host(i % 3)
                    net.submit_host(*pinned[i % nb], out=houts[i % 3], slot=i % 3, device=dev, stream=cstream)
                for i in range(max(0, k - 3), k):
                    net.wait_host(i % 3)
        api = ('cair_ranker_submit_host / cair_ranker_wait_host over 3 staging slots: H2D of batch k+1 and its document encoder '
               'overlap the interaction kernel of batch k; 25 distinct pinned batches per step'
               + ('; the scores of every batch are all-gathered over NVLink peer memory inside the pipeline '
                  '(cair_ranker_set_gather) and every rank copies all B*world*N scores to its host' if world > 1 else ''))
    else:
        # N > 1: the score all-gather sits between the kernels and the D2H copy, so the loop runs on the device entry point:
        # pinned ids -> device staging (3 slots) -> this rank's scores -> NCCL all-gather -> all B*world*N scores to the host
        dq = [[torch.empty_like(t, device=dev) for t in pinned[0]] for _ in range(3)]

        def e2e_loop(k):
            with torch.cuda.stream(cstream), torch.no_grad():
                for i in range(k):
                    s3 = i % 3
                    for dst, src in zip(dq[s3], pinned[i % nb]):
                        dst.copy_(src, non_blocking=True)
                    s = net(*dq[s3])
                    g = gather_scores(s.reshape(-1), pairs_total)
                    houts[s3].view(-1).copy_(g, non_blocking=True)
                cstream.synchronize()
        api = ('device entry point on 3 staging slots: pinned ids -> H2D -> kernels -> all-gather of the scores (see run.collective) -> D2H of all '
               'B*world*N scores, stream-ordered, one synchronisation per step of 25 batches')
    e2e_loop(2 * nb)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler.active = True
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_loop(nb)
    e2e_total = time.perf_counter() - t0
    sampler.active = False
    e = torch.tensor([e2e_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e, op=dist.ReduceOp.MAX)
    e2e_value = pairs_total * nb / (float(e[0].item()) / args.steps)
    # the synchronous single-call form (cair_ranker_forward_host: one cached CUDA graph per call, returns after the D2H)
    e2e_sync_value = None
    if world == 1:
        hout1 = torch.empty(B, N, dtype=torch.float32).pin_memory()
        for _ in range(3):
            net.forward_host(*pinned[0], out=hout1, device=dev)
        t1 = time.perf_counter()
        for i in range(40):
            net.forward_host(*pinned[i % nb], out=hout1, device=dev)
        e2e_sync_value = pairs_total * 40 / (time.perf_counter() - t1)
    h2d = sum(t.numel() for t in pinned[0]) * 8 * nb * world
    d2h = (B * N * 4 if world == 1 else B * world * N * 4) * nb * world
    if rank == 0:
        sampler.stop_flag = True
        sampler.join(timeout=2)

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    others = None
    if not args.no_other_configs:
        others = other_configs(run, args, peaks)

    if rank == 0:
        hbm_peak = peaks.get('hbm_gbs', 6650.0)
        tf_peak = peaks.get('bf16_tflops_sustained', 1400.0)
        peak_src = 'measured (MEASURED_PEAKS.json, sustained bf16)' if peaks else 'fallback (B200_PROFILING.md)'
        fl = flops_per_pair()
        timed = {k: v for k, v in stages.items() if k not in ('begin', 'join_query_side', 'end')}
        top = max(timed, key=timed.get) if timed else None
        roof = None
        if top == 'interact':
            ach = fl['interact_exec'] * 1e6 * pairs_local / (stages[top] / 1e3) / 1e12
            tr, src = traffic_for('mt_tc_interact_kernel')
            roof = dict(kernel='mt_tc_interact_kernel', bound='tensor', achieved=ach, peak=tf_peak, unit='TFLOP/s',
                        frac=ach / tf_peak, traffic=tr, traffic_source=src, achieved_ref_equiv=fl['interact_ref'] / fl['interact_exec'] * ach,
                        achieved_issued_bf16=3 * ach, ms_per_launch=stages[top], peak_source=peak_src)
        elif top in ('doc_recurrence', 'lstm_recurrence'):
            # algorithmic FLOPs of the doc BiLSTM (input + recurrent projection, both directions); the kernel issues
            # 3 bf16 MMA passes per logical FLOP and is bound by the 200-step recurrence latency, not by the pipe
            ach = 2 * LD * 2 * 4 * 64 * (40 + 64) * pairs_local / (stages[top] / 1e3) / 1e12
            tr, src = traffic_for('lstm_tc_kernel')
            roof = dict(kernel='lstm_tc_kernel', bound='tensor', achieved=ach, peak=tf_peak, unit='TFLOP/s',
                        frac=ach / tf_peak, traffic=tr, traffic_source=src, achieved_issued_bf16=3 * ach, ms_per_launch=stages[top],
                        steps_per_launch=LD, us_per_recurrence_step=1e3 * stages[top] / LD, peak_source=peak_src,
                        note='a 200-step dependency chain: bounded by per-step latency (MMA issue + TMEM read + MUFU cell update + '
                             'hand-over), see DESIGN.md 4.2 / profiles/r02_rnn_*')
        elif top is not None:
            ach = bytes_per_pair_folded() * pairs_local / (stages[top] / 1e3) / 1e9
            roof = dict(kernel=top, bound='hbm', achieved=ach, peak=hbm_peak, unit='GB/s', frac=ach / hbm_peak,
                        traffic=None, ms_per_launch=stages[top], peak_source=peak_src)
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            try:
                r = subprocess.run([sys.executable, os.path.abspath(__file__), '--cpu-baseline-worker', '--cpu-sample-queries',
                                    str(args.cpu_sample_queries), '--cpu-sample-seconds', str(args.cpu_sample_seconds)],
                                   capture_output=True, text=True, timeout=600)
                cpu = json.loads(r.stdout.strip().splitlines()[-1])
            except Exception as ex:   # the baseline is a reported extra, never a reason to lose the bench line
                cpu = dict(value=None, unit=UNIT, cores=os.cpu_count(), kind='port', sample='failed: %r' % (ex,))
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3),
            'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32 (tcgen05 bf16x3 split-precision MMA, fp32 accumulate/state)', 'data': 'synthetic',
            'config': CONFIG,
            'run': {'per_gpu_pairs_per_batch': pairs_local, 'global_pairs_per_step': pairs_total * nb, 'ms_per_batch': ms_per_step / nb,
                    'parallelism': 'doc-parallel x%d, one all-gather of scores per batch' % world, 'collective': collective,
                    'collective_fallback_reason': run.p2p_error,
                    'timed_region_s': ms_per_step * args.steps / 1e3,
                    'table': 'eval-mode folded [V,40] table, pre-split into bf16 hi/lo operand rows (192 B per token, built once at handle creation)'},
            'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h, 'api': api,
                    'includes_collective': world > 1, 'single_call_value': e2e_sync_value,
                    'single_call_api': 'cair_ranker_forward_host: H2D + kernels + D2H as one cached CUDA graph, synchronous'},
            'gpu_launches': int(launches), 'clocks': sampler.summary(),
            'roofline': roof, 'stages_ms': stages,
            'flops_per_pair_mflop': fl, 'hbm_bytes_per_pair': bytes_per_pair_folded(),
            'hbm_frac_end_to_end': bytes_per_pair_folded() * value / world / 1e9 / hbm_peak,
        }
        if others is not None:
            line['other_configs'] = others
        if cpu is not None:
            line['cpu_baseline'] = cpu
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--cpu-sample-queries', type=int, default=32)
    ap.add_argument('--cpu-sample-seconds', type=float, default=12.0)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-other-configs', action='store_true')
    ap.add_argument('--nccl-collective', action='store_true', help='use NCCL for the score all-gather instead of cair_allgather_scores')
    ap.add_argument('--cpu-baseline-worker', action='store_true', help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.cpu_baseline_worker:
        cpu_baseline_worker(args)
    elif args.impl == 'reference':
        run_reference(args)
    else:
        run_product(args)


if __name__ == '__main__':
    main()
