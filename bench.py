#!/usr/bin/env python3
"""Benchmark of the Seekmer bulk-infer hot path on B200 (contract: see the repo README/DESIGN).

    python bench.py --gpus N --steps K --warmup W            # this implementation
    python bench.py --impl reference --gpus N ...            # the reference's CPU path

Workload (BASELINE.json configs[1], scaled weakly per GPU): synthetic human-scale
transcriptome (200 000 transcripts, ~300 Mb cDNA, isoform families), indexed on the GPU in
the reference's array layout, + 30 M simulated 2x150 bp pairs per GPU (1 % substitutions,
0.1 % N, 1 % unalignable pairs).  A "step" is one mapping pass over all pairs: reads ->
equivalence classes (device dictionary) + FLD, classes merged across ranks and exported.
EM and the 100-replicate bootstrap are timed separately and reported in `em`.

One JSON line is printed by rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = 'read_pairs_pseudoaligned_per_sec'
UNIT = 'pairs/s'
READ_LEN, FRAG_MEAN, FRAG_SD = 150, 350, 50
SEED_TX, SEED_EXPR, SEED_READS = 2, 3, 10


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--pairs', type=int, default=30_000_000, help='read pairs per GPU')
    ap.add_argument('--transcripts', type=int, default=200_000)
    ap.add_argument('--bootstraps', type=int, default=100)
    ap.add_argument('--cpu-sample', type=int, default=2_000_000, help='pairs timed on the CPU baseline')
    ap.add_argument('--e2e-batch', type=int, default=0, help='pairs per host-buffer call (0 = one call; the library '
                    'double-buffers H2D copies against the kernels internally)')
    ap.add_argument('--no-em', action='store_true')
    ap.add_argument('--no-cpu', action='store_true')
    return ap.parse_args()


def log(*a):
    print('[bench]', *a, file=sys.stderr, flush=True)


# ------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.rows = []   # (host time, csv line)
        self.proc = None
        self.gpu = gpu_index
        self.window = [None, None]

    def start(self):
        """Started BEFORE the warm-up steps (nvidia-smi needs a few 100 ms to come up and a timed
        region of 3 steps is ~0.1 s); samples are time-stamped on arrival."""
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(self.gpu), '--query-gpu=' + self.Q, '--format=csv,noheader,nounits',
                 '-lms', '20'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def mark_begin(self):
        self.window[0] = time.time()

    def mark_end(self):
        self.window[1] = time.time()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            pass
        return self.summarize(self.rows, self.window)

    @staticmethod
    def summarize(stamped_rows, window):
        """Reduce time-stamped nvidia-smi csv lines to the `clocks` object of the bench line.
        Only the samples that arrived inside `window` count when there are any."""
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        t0, t1 = window
        timed = [r for t, r in stamped_rows if t0 is not None and t1 is not None and t0 <= t <= t1 + 0.02]
        # a timed region shorter than the sampling period: fall back to the samples taken while the
        # identical warm-up steps ran just before it
        near = [r for t, r in stamped_rows
                if t0 is not None and t1 is not None and t0 - 0.25 <= t <= t1 + 0.05]
        rows = timed or near or [r for _, r in stamped_rows]
        for r in rows:
            f = [x.strip() for x in r.split(',')]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(names, f[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': float(numpy.median(sm)) if sm else None,
                'sm_max_mhz': float(max(mx)) if mx else None, 'samples': len(sm),
                'samples_in_timed_region': len(timed), 'reasons': sorted(reasons)}


# ------------------------------------------------------------------ workload
def make_workload(args, rank, world, device):
    """Transcriptome (rank 0 generates, broadcast), index built on this GPU, simulator params."""
    import torch
    import torch.distributed as dist
    from seekmer_b200 import index_build, synth
    t0 = time.time()
    if rank == 0:
        tx = synth.make_transcriptome(args.transcripts, seed=SEED_TX, mean_exons=11)
        codes = torch.from_numpy(tx.codes).to(device)
        offsets = torch.from_numpy(tx.offsets).to(device)
        shape = torch.tensor([codes.shape[0], offsets.shape[0]], dtype=torch.int64, device=device)
    else:
        shape = torch.zeros(2, dtype=torch.int64, device=device)
    if world > 1:
        dist.broadcast(shape, 0)
        if rank != 0:
            codes = torch.empty(int(shape[0]), dtype=torch.uint8, device=device)
            offsets = torch.empty(int(shape[1]), dtype=torch.int64, device=device)
        dist.broadcast(codes, 0)
        dist.broadcast(offsets, 0)
    t1 = time.time()
    built = index_build.build_index(codes, offsets, device=device)
    torch.cuda.synchronize()
    t2 = time.time()
    lengths = (offsets[1:] - offsets[:-1]).cpu().numpy()
    expr = synth.make_expression(lengths.shape[0], seed=SEED_EXPR)
    w = expr * numpy.maximum(lengths - FRAG_MEAN + 1, 1)
    w = w / w.sum()
    cum = numpy.cumsum(numpy.floor(w * float(1 << 40)).astype('u8')).astype('u8')
    sim = dict(codes=codes, offsets=offsets, cum=torch.from_numpy(cum.view('i8')).to(device), total=int(cum[-1]),
               n_tx=lengths.shape[0])
    if rank == 0:
        log('transcriptome %.1f Mb in %.1fs; index built in %.1fs: %s' %
            (codes.shape[0] / 1e6, t1 - t0, t2 - t1, built.stats))
    return built, sim, lengths


def synth_reads(sim, first_unit, n_units, out, device_index):
    from seekmer_b200 import _lib
    L = _lib.load()
    step = 8_000_000
    for s in range(0, n_units, step):
        n = min(step, n_units - s)
        _lib.check(L.skm_synth_reads(
            _lib._ptr(sim['codes']), _lib._ptr(sim['offsets']), sim['n_tx'], _lib._ptr(sim['cum']), sim['total'],
            READ_LEN, FRAG_MEAN, FRAG_SD, int(round(0.01 * 65536)), int(round(0.001 * 65536)), 1, SEED_READS, 1,
            first_unit + s, n, out.data_ptr() + s * 2 * READ_LEN, device_index, _lib.current_stream_ptr()))


def algorithmic_bytes_per_pair(orc, oidx, bases, n_pairs):
    """B_read = L + 16 S + 48 K + 8 T + 8 Q from the counting oracle."""
    offs = numpy.arange(2 * n_pairs + 1, dtype='i8') * READ_LEN
    out = orc.map_batch(oidx, bases[:2 * n_pairs * READ_LEN], offs, True, counters=True)
    c = out.counters
    reads = 2.0 * n_pairs
    per_read = (READ_LEN + 16.0 * c['slots'] / reads + 48.0 * c['contig_reads'] / reads
                + 8.0 * (c['map_contig_items'] + c['filter_items']) / reads + 8.0 * c['windows'] / reads)
    return 2.0 * per_read + 8.0, {k: round(v / reads, 3) for k, v in c.items()}


# ------------------------------------------------------------------ our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from seekmer_b200 import _lib, dist as sdist, infer, mapper

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world != args.gpus and world > 1:
        log('warning: --gpus %d but WORLD_SIZE %d' % (args.gpus, world))
    torch.cuda.set_device(local)
    device = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=device)
    built, sim, lengths = make_workload(args, rank, world, device)
    index = _lib.DeviceIndex(built.kmers, built.contigs, built.sequences, built.targets, built.n_transcripts)
    info = index.info()
    mp = _lib.DeviceMapper(index, class_capacity=1 << 22, id_capacity=1 << 26)  # ~4x the classes of this workload

    n_pairs = args.pairs
    first_unit = rank * n_pairs
    d_bases = torch.empty(n_pairs * 2 * READ_LEN, dtype=torch.uint8, device=device)
    synth_reads(sim, first_unit, n_pairs, d_bases, local)
    torch.cuda.synchronize()

    launches = {'n': 0}

    kernel_times = []

    def step_device(timed=False):
        mp.reset()
        mp.map_batch(d_bases, None, n_pairs, True, first_unit=first_unit, fixed_len=READ_LEN)
        launches['n'] += 3  # pack_reads_kernel, map_reads_kernel, tally_units_kernel
        if world > 1:
            # one all-gather of the exported dictionaries, peers merged on the device
            table = sdist.merge_mappers(mp)
            launches['n'] += 4 + 2 * (world - 1)  # 2 exports (2 kernels each) + merge + FLD add per peer
        else:
            table = mp.export_torch()
            launches['n'] += 2  # dict_export_kernel, set_i64_kernel
        if timed:
            # CUDA events the library records around each kernel on the launch stream; the
            # export above has already synchronised that stream
            kernel_times.append(mp.kernel_ms())
        return table

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    sampler.start()
    step_device()  # first step also warms nvidia-smi up
    time.sleep(0.3)
    for _ in range(max(args.warmup - 1, 0)):
        step_device()
    barrier()
    sampler.mark_begin()
    launches['n'] = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(args.steps):
        table = step_device(timed=True)
    e1.record()
    barrier()
    sampler.mark_end()
    clocks = sampler.stop()
    total_ms = e0.elapsed_time(e1)
    kernel_ms = float(numpy.mean([k['map_reads_kernel'] for k in kernel_times]))
    kernels_ms = {name: round(float(numpy.mean([k[name] for k in kernel_times])), 3) for name in kernel_times[0]}
    gpu_launches = launches['n']
    t = torch.tensor([total_ms, kernel_ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, kernel_ms_max = float(t[0]), float(t[1])
    ms_per_step = total_ms / args.steps
    value = world * n_pairs / (ms_per_step / 1e3)
    host_table = sdist.table_to_host(table)

    # ---- end to end through the C ABI with HOST buffers (pinned): H2D copies inside
    h_bases = torch.empty(n_pairs * 2 * READ_LEN, dtype=torch.uint8, pin_memory=True)
    h_bases.copy_(d_bases)
    torch.cuda.synchronize()
    h_np = h_bases.numpy()

    def step_e2e():
        mp.reset()
        batch = args.e2e_batch or n_pairs
        for s in range(0, n_pairs, batch):
            n = min(batch, n_pairs - s)
            mp.map_batch(h_np[s * 2 * READ_LEN:(s + n) * 2 * READ_LEN], None, n, True, first_unit=first_unit + s,
                         fixed_len=READ_LEN)
        tab = sdist.merge_mappers(mp) if world > 1 else mp.export_torch()
        return sdist.table_to_host(tab)

    e2e_steps = max(1, min(args.steps, 2))
    step_e2e()
    barrier()
    w0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_table = step_e2e()
    barrier()
    e2e_s = (time.perf_counter() - w0) / e2e_steps
    t = torch.tensor([e2e_s], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * n_pairs / float(t[0])
    d2h_bytes = int(sum(host_table[k].nbytes for k in ('key_offsets', 'key_ids', 'counts', 'first_unit', 'fld')))
    assert (e2e_table['counts'] == host_table['counts']).all() and (e2e_table['key_ids'] == host_table['key_ids']).all()

    # ---- EM + bootstraps (timed separately; replicates shard across ranks)
    em = None
    if not args.no_em:
        class FakeIndex:
            transcripts = numpy.zeros(lengths.shape[0], dtype=[('length', 'f8')])
        FakeIndex.transcripts['length'] = lengths
        mr = mapper.MapResult(FakeIndex)
        mr.fragment_length_counts = host_table['fld'].astype('i8')
        summ = mapper.summarize_table(host_table, mr)
        barrier()
        w0 = time.perf_counter()
        main, main_iters = None, None
        x = numpy.ones(lengths.shape[0]) / summ.effective_lengths
        x /= x.sum()
        main_x, main_iters = infer.em(x, summ.effective_lengths, summ.class_map, summ.class_count, return_iters=True)
        main = infer._finish(main_x)
        torch.cuda.synchronize()
        em_main_s = time.perf_counter() - w0
        per_rank = (args.bootstraps + world - 1) // world
        r0 = rank * per_rank
        nrep = max(0, min(per_rank, args.bootstraps - r0))
        barrier()
        w0 = time.perf_counter()
        boots, iters = infer.quantify_bootstraps(summ, main, nrep, seed=1234, return_iters=True, first_replicate=r0)
        if world > 1:
            mine = torch.from_numpy(numpy.stack(boots) if boots else numpy.zeros((0, lengths.shape[0]))).to(device)
            pad = torch.zeros(per_rank, lengths.shape[0], dtype=torch.float64, device=device)
            pad[:mine.shape[0]] = mine
            gathered = [torch.zeros_like(pad) for _ in range(world)]
            dist.all_gather(gathered, pad)
        barrier()
        boot_s = time.perf_counter() - w0
        em = {'main_ms': round(em_main_s * 1e3, 3), 'main_iters': int(main_iters),
              'bootstrap_ms': round(boot_s * 1e3, 3), 'bootstraps': args.bootstraps,
              'bootstrap_iters_mean': float(numpy.mean(iters)) if len(iters) else None,
              'n_classes': int(summ.class_count.size), 'nnz': int(summ.class_map.shape[1]),
              'n_transcripts': int(lengths.shape[0]),
              'em_plus_bootstraps_ms': round((em_main_s + boot_s) * 1e3, 3)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline + CPU baseline (rank 0; the oracle is only the checker / the baseline here)
    roofline, cpu_baseline, parity = None, None, None
    try:
        from oracle import oracle as orc
        arrays = built.numpy_arrays()
        oidx = orc.OracleIndex(*arrays)
        sample = min(args.cpu_sample, n_pairs)
        hb = h_np[:sample * 2 * READ_LEN]
        b_pair, per_read = algorithmic_bytes_per_pair(orc, oidx, hb, min(100_000, sample))
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        except Exception:
            pass
        peak = float(peaks.get('hbm_gbs', 6650.0))
        achieved = n_pairs * b_pair / (kernel_ms_max / 1e3) / 1e9
        # dram__bytes_read.sum + dram__bytes_write.sum of map_reads_kernel from the committed
        # ncu --set full capture (profiles/): bytes per pair there x the pairs of one launch here
        traffic, traffic_src = None, None
        try:
            tr = json.load(open(os.path.join(ROOT, 'profiles', 'map_reads_kernel_traffic.json')))
            traffic = round(tr['dram_bytes_per_pair'] * n_pairs)
            traffic_src = tr['source']
        except Exception:
            pass
        roofline = {'bound': 'hbm', 'kernel': 'map_reads_kernel', 'achieved': round(achieved, 2), 'peak': peak,
                    'peak_source': 'measured (MEASURED_PEAKS.json hbm_gbs, burst)' if 'hbm_gbs' in peaks
                    else 'fallback (B200_PROFILING.md)', 'unit': 'GB/s',
                    'frac': round(achieved / peak, 4), 'traffic': traffic, 'traffic_source': traffic_src,
                    'algorithmic_bytes_per_pair': round(b_pair, 1), 'pairs_per_launch': n_pairs,
                    'kernel_ms': round(kernel_ms_max, 3), 'step_kernels_ms': kernels_ms,
                    'kernel_share_of_step': round(kernel_ms_max / ms_per_step, 3),
                    'per_read_accesses': per_read}
        if not args.no_cpu:
            cores = os.cpu_count() or 1
            offs = numpy.arange(2 * sample + 1, dtype='i8') * READ_LEN
            w0 = time.perf_counter()
            aligned, h, cnt, length, fld = orc.map_batch_mt(oidx, hb, offs, True, cores)
            cpu_s = time.perf_counter() - w0
            cpu_baseline = {'value': round(sample / cpu_s, 1), 'unit': UNIT, 'cores': cores, 'kind': 'port',
                            'sample': '%d of the same 2x%d pairs, C oracle (OpenMP), index resident in RAM'
                                      % (sample, READ_LEN)}
            # parity on the sample, from an independent GPU pass with per-unit outputs: FLD, aligned
            # count, every unit's span length, and the partition of units into classes (two units
            # share a GPU dictionary slot exactly when the oracle gives them the same id tuple)
            mp.reset()
            g_cls, g_len = mp.map_batch(d_bases[:sample * 2 * READ_LEN], None, sample, True, first_unit=0,
                                        fixed_len=READ_LEN, per_read=True)
            chk = mp.export()
            g_cls, g_len = g_cls.cpu().numpy().astype('i8'), g_len.cpu().numpy()
            o_key = numpy.where(cnt > 0, h, numpy.uint64(0)).astype('u8')  # unaligned units: one class
            g_key = numpy.where(g_cls >= 0, g_cls + 1, 0)
            pairs_seen = numpy.unique(numpy.stack([g_key.astype('u8'), o_key]), axis=1)
            one_to_one = (numpy.unique(pairs_seen[0]).size == pairs_seen.shape[1]
                          and numpy.unique(pairs_seen[1]).size == pairs_seen.shape[1])
            parity = {'sample_pairs': sample, 'fld_equal': bool((chk['fld'] == fld).all()),
                      'aligned_equal': bool(chk['aligned'] == aligned),
                      'unit_lengths_equal': bool((g_len == length).all()),
                      'unit_classes_equal': bool(one_to_one and ((g_cls >= 0) == (cnt > 0)).all())}
    except Exception as exc:  # the oracle is optional infrastructure for the bench line
        log('oracle leg skipped: %r' % (exc,))

    line = {
        'metric': METRIC, 'value': round(value, 1), 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': round(ms_per_step, 3), 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'u64', 'data': 'synthetic',
        'config': {'workload': 'human-scale synthetic transcriptome (%d transcripts, %.0f Mb cDNA) + %d M 2x%d bp '
                               'pairs per GPU, 1%% subs' % (args.transcripts, sim['codes'].shape[0] / 1e6,
                                                            n_pairs // 1_000_000, READ_LEN),
                   'pairs_per_gpu': n_pairs, 'read_len': READ_LEN, 'index_kmers': info['n_kmers'],
                   'index_table_slots': info['table_slots'], 'index_device_bytes': info['device_bytes'],
                   'n_contigs': info['n_contigs'], 'l2_policy': 'inputs (%.1f GB reads + %.1f GB table) larger than L2'
                   % (d_bases.numel() / 1e9, info['table_slots'] * 16 / 1e9),
                   'parallelism': 'reads sharded x%d, index replicated' % world},
        'clocks': clocks,
        'e2e': {'value': round(e2e_value, 1), 'unit': UNIT, 'h2d_bytes_per_step': int(n_pairs * 2 * READ_LEN),
                'd2h_bytes_per_step': d2h_bytes, 'ms_per_step': round(float(t[0]) * 1e3, 3)},
        'gpu_launches': gpu_launches,
        'classes': {'n_classes': int(host_table['counts'].shape[0]), 'aligned': host_table['aligned'],
                    'unaligned': host_table['unaligned']},
    }
    if roofline:
        line['roofline'] = roofline
    if cpu_baseline:
        line['cpu_baseline'] = cpu_baseline
    if parity:
        line['parity_on_sample'] = parity
    if em:
        line['em'] = em
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------ reference arm
def run_reference(args):
    """The reference's own CPU implementation of the path on this box's host cores."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    import torch
    from oracle import build_ref, oracle as orc
    from seekmer_b200 import _lib
    cores = os.cpu_count() or 1
    sample = min(args.cpu_sample, args.pairs)

    class A:
        pass
    device = torch.device('cuda', 0)
    torch.cuda.set_device(0)
    built, sim, lengths = make_workload(args, 0, 1, device)
    d_bases = torch.empty(sample * 2 * READ_LEN, dtype=torch.uint8, device=device)
    synth_reads(sim, 0, sample, d_bases, 0)
    torch.cuda.synchronize()
    bases = d_bases.cpu().numpy()
    arrays = built.numpy_arrays()
    del built, d_bases
    torch.cuda.empty_cache()
    kind = 'port'
    run = None
    if build_ref.built():
        try:
            from oracle import ref_harness as rh
            rh.load_ref()
            ridx = rh.ref_index_from_arrays(*arrays)
            raw = bases.tobytes()
            bsz = 65536  # common.BUFFER_SIZE
            batches = []
            for s in range(0, sample, bsz):
                n = min(bsz, sample - s)
                reads = [raw[(2 * s + i) * READ_LEN:(2 * s + i + 1) * READ_LEN] for i in range(2 * n)]
                batches.append((n, [b''] * n, reads))
            kind = 'reference'

            def run():
                res = rh.ref_map_threads(ridx, batches, cores)  # mapper.map_reads threading model, -j cores
                return sum(v for k, v in res.counter.items() if k)
        except Exception as exc:
            log('compiled reference unavailable (%r); timing the C port' % (exc,))
    if run is None:
        oidx = orc.OracleIndex(*arrays)
        offs = numpy.arange(2 * sample + 1, dtype='i8') * READ_LEN

        def run():
            return orc.map_batch_mt(oidx, bases, offs, True, cores)[0]
    for _ in range(min(args.warmup, 1)):
        run()
    steps = max(1, min(args.steps, 3))
    w0 = time.perf_counter()
    for _ in range(steps):
        aligned = run()
    dt = (time.perf_counter() - w0) / steps
    value = sample / dt
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': round(value, 1), 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': steps, 'warmup': min(args.warmup, 1), 'ms_per_step': round(dt * 1e3, 3), 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'u64', 'data': 'synthetic',
        'config': {'workload': 'human-scale synthetic transcriptome (%d transcripts, %.0f Mb cDNA) + %d M 2x%d bp '
                               'pairs per GPU, 1%% subs' % (args.transcripts, sim['codes'].shape[0] / 1e6,
                                                            args.pairs // 1_000_000, READ_LEN),
                   'sample_pairs_per_step': sample, 'aligned': int(aligned)},
        'cpu_baseline': {'value': round(value, 1), 'unit': UNIT, 'cores': cores, 'kind': kind,
                         'sample': '%d of the same 2x%d pairs per step; %s' % (
                             sample, READ_LEN,
                             'compiled reference natives (oracle/_ref), mapper.map_reads threading model with '
                             'job_count=%d, pre-materialised feeder batches' % cores if kind == 'reference'
                             else 'C oracle (OpenMP, %d threads)' % cores)},
        'e2e': {'value': round(value, 1), 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line), flush=True)


def main():
    args = parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
