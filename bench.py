#!/usr/bin/env python3
"""Benchmark of the Seekmer bulk-infer hot path on B200 (contract: see the repo README/DESIGN).

    python bench.py --gpus N --steps K --warmup W            # this implementation
    python bench.py --impl reference --gpus N ...            # the reference's CPU path
    python bench.py --scaling strong --total-pairs 200000000 # BASELINE configs[2] (fixed total work)
    python bench.py --config c5                              # BASELINE configs[4] (16 SE75 samples)

Workload (BASELINE.json configs[1], scaled weakly per GPU): synthetic human-scale
transcriptome (200 000 transcripts, ~300 Mb cDNA, isoform families), indexed on the GPU in
the reference's array layout, + 30 M simulated 2x150 bp pairs per GPU (1 % substitutions,
0.1 % N, 1 % unalignable pairs).  A "step" is one mapping pass over all pairs: reads ->
equivalence classes (device dictionary) + FLD, classes merged across ranks and exported.
EM and the 100-replicate bootstrap are timed separately and reported in `em`.

One JSON line is printed by rank 0.
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import tempfile
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
DICT_CLASSES, DICT_IDS = 1 << 21, 1 << 24  # class dictionary of the mapping legs: classes, ids over all classes
C5_READ_LEN, C5_SUB_RATE = 75, 0.02


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--config', default='c2', choices=['c2', 'c5'],
                    help='c2: 2x150 pairs against the human-scale index (the headline); c5: 16 single-end 75 bp samples')
    ap.add_argument('--pairs', type=int, default=30_000_000, help='read pairs per GPU (weak scaling)')
    ap.add_argument('--scaling', default='weak', choices=['weak', 'strong'])
    ap.add_argument('--total-pairs', type=int, default=200_000_000, help='read pairs of the whole job (strong scaling)')
    ap.add_argument('--transcripts', type=int, default=200_000)
    ap.add_argument('--bootstraps', type=int, default=100)
    ap.add_argument('--cpu-sample', type=int, default=2_000_000, help='pairs timed on the CPU baseline')
    ap.add_argument('--cpu-bootstraps', type=int, default=3, help='bootstrap replicates timed on the CPU (extrapolated)')
    ap.add_argument('--fastq-pairs', type=int, default=4_000_000, help='pairs of the FASTQ end-to-end leg (0 = skip)')
    ap.add_argument('--samples', type=int, default=16, help='c5: samples')
    ap.add_argument('--sample-reads', type=int, default=4_000_000, help='c5: single-end reads per sample')
    ap.add_argument('--e2e-batch', type=int, default=0, help='pairs per host-buffer call (0 = one call; the library '
                    'double-buffers H2D copies against the kernels internally)')
    ap.add_argument('--no-em', action='store_true')
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--make-workload', default=None, help=argparse.SUPPRESS)  # child of the reference arm
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


# ------------------------------------------------------------------ host placement
def bind_to_gpu_numa(device_index):
    """Run this rank's host threads (and so first-touch its pinned buffers) on the NUMA node the
    GPU hangs off: concurrent H2D copies from one node are what limits end-to-end scaling."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(device_index).pci_bus_id
        dom = torch.cuda.get_device_properties(device_index).pci_domain_id
        dev = torch.cuda.get_device_properties(device_index).pci_device_id
        path = '/sys/bus/pci/devices/%04x:%02x:%02x.0/numa_node' % (dom, bus, dev)
        node = int(open(path).read().strip())
        if node < 0:
            return {'numa_node': None}
        cpus = set()
        for part in open('/sys/devices/system/node/node%d/cpulist' % node).read().strip().split(','):
            lo, _, hi = part.partition('-')
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if allowed:
            os.sched_setaffinity(0, allowed)
        return {'numa_node': node, 'cpus': len(allowed)}
    except Exception as exc:  # best effort: containers often hide the topology
        return {'numa_node': None, 'why': repr(exc)[:80]}


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
    frag = FRAG_MEAN if getattr(args, 'config', 'c2') == 'c2' else 250
    w = expr * numpy.maximum(lengths - frag + 1, 1)
    w = w / w.sum()
    cum = numpy.cumsum(numpy.floor(w * float(1 << 40)).astype('u8')).astype('u8')
    sim = dict(codes=codes, offsets=offsets, cum=torch.from_numpy(cum.view('i8')).to(device), total=int(cum[-1]),
               n_tx=lengths.shape[0])
    if rank == 0:
        log('transcriptome %.1f Mb in %.1fs; index built in %.1fs: %s' %
            (codes.shape[0] / 1e6, t1 - t0, t2 - t1, built.stats))
    return built, sim, lengths


def synth_reads(sim, first_unit, n_units, out, device_index, read_len=READ_LEN, frag_mean=FRAG_MEAN, frag_sd=FRAG_SD,
                sub_rate=0.01, paired=True, seed=SEED_READS):
    from seekmer_b200 import _lib
    L = _lib.load()
    step = 8_000_000
    per_unit = (2 if paired else 1) * read_len
    for s in range(0, n_units, step):
        n = min(step, n_units - s)
        _lib.check(L.skm_synth_reads(
            _lib._ptr(sim['codes']), _lib._ptr(sim['offsets']), sim['n_tx'], _lib._ptr(sim['cum']), sim['total'],
            read_len, frag_mean, frag_sd, int(round(sub_rate * 65536)), int(round(0.001 * 65536)), 1, seed,
            1 if paired else 0, first_unit + s, n, out.data_ptr() + s * per_unit, device_index,
            _lib.current_stream_ptr()))


def algorithmic_bytes_per_pair(orc, oidx, bases, n_pairs):
    """B_read = L + 16 S + 48 K + 8 T + 8 Q from the counting oracle.  Returns the
    bytes per pair with ASCII reads in (the timed step: pack + map + tally), the bytes per pair of
    the map kernel alone (it reads 2-bit packed reads: L / 4), and the per-read access counts."""
    offs = numpy.arange(2 * n_pairs + 1, dtype='i8') * READ_LEN
    out = orc.map_batch(oidx, bases[:2 * n_pairs * READ_LEN], offs, True, counters=True)
    c = out.counters
    reads = 2.0 * n_pairs
    index_side = (16.0 * c['slots'] / reads + 48.0 * c['contig_reads'] / reads
                  + 8.0 * (c['map_contig_items'] + c['filter_items']) / reads + 8.0 * c['windows'] / reads)
    step = 2.0 * (READ_LEN + index_side) + 8.0
    kernel = 2.0 * (READ_LEN / 4.0 + index_side) + 8.0
    return step, kernel, {k: round(v / reads, 3) for k, v in c.items()}


def table_digest(table):
    """Order-independent identity of a class table: sorted (ordered id tuple, count) pairs, the FLD
    and the unaligned / aligned totals."""
    off = numpy.asarray(table['key_offsets'], dtype='i8')
    ids = numpy.asarray(table['key_ids'], dtype='i4')
    counts = numpy.asarray(table['counts'], dtype='i8')
    n = counts.shape[0]
    lens = off[1:] - off[:-1]
    width = int(lens.max()) if n else 0
    rows = numpy.full((n, width + 2), -1, dtype='i8')
    if n:
        r = numpy.repeat(numpy.arange(n), lens)
        c = numpy.arange(ids.shape[0]) - numpy.repeat(off[:-1], lens)
        rows[r, c + 2] = ids
    rows[:, 0] = lens
    rows[:, 1] = counts
    order = numpy.lexsort(rows.T[::-1])
    h = hashlib.sha256()
    h.update(numpy.ascontiguousarray(rows[order]).tobytes())
    h.update(numpy.ascontiguousarray(table['fld'], dtype='i8').tobytes())
    h.update(('%d %d' % (int(table['unaligned']), int(table['aligned']))).encode())
    return h.hexdigest()[:20]


def em_bytes_per_iteration(n_classes, nnz, n_tx):
    """B_iter = 8 nnz + 20 C + 36 T (one replicate)."""
    return 8.0 * nnz + 20.0 * n_classes + 36.0 * n_tx


def write_fastq_pair(folder, bases, n_pairs, read_len):
    """Two FASTQ files (mates) of fixed-width records written with numpy block copies."""
    paths = []
    name = numpy.frombuffer(b'@r\n', dtype='u1')
    tail = numpy.frombuffer(b'\n+\n' + b'I' * read_len + b'\n', dtype='u1')
    width = name.shape[0] + read_len + tail.shape[0]
    reads = bases[:2 * n_pairs * read_len].reshape(n_pairs, 2, read_len)
    for mate in range(2):
        rec = numpy.empty((n_pairs, width), dtype='u1')
        rec[:, :name.shape[0]] = name
        rec[:, name.shape[0]:name.shape[0] + read_len] = reads[:, mate]
        rec[:, name.shape[0] + read_len:] = tail
        path = os.path.join(folder, 'reads_%d.fq' % (mate + 1))
        rec.tofile(path)
        paths.append(path)
    return paths, 2 * n_pairs * width


# ------------------------------------------------------------------ our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from seekmer_b200 import _lib, common, dist as sdist, infer, mapper

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world != args.gpus and world > 1:
        log('warning: --gpus %d but WORLD_SIZE %d' % (args.gpus, world))
    torch.cuda.set_device(local)
    device = torch.device('cuda', local)
    placement = bind_to_gpu_numa(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=device)
    built, sim, lengths = make_workload(args, rank, world, device)
    index = _lib.DeviceIndex(built.kmers, built.contigs, built.sequences, built.targets, built.n_transcripts)
    info = index.info()
    strong = args.scaling == 'strong'
    n_pairs = args.total_pairs // world if strong else args.pairs
    total_pairs = n_pairs * world
    # dictionary sized for the classes of the whole job (~0.8 M at one rank, ~1.0 M at eight): 2^22
    # slots.  Its arrays (40 B per slot + the id pool) then stay inside the 256 MB the TLB reaches,
    # which is what the random accesses of the merge kernel and the slot scan of the export cost
    # (profiles/r02c_*: one page walk per access beyond that reach).
    mp = _lib.DeviceMapper(index, class_capacity=DICT_CLASSES, id_capacity=DICT_IDS)

    first_unit = rank * n_pairs
    d_bases = torch.empty(n_pairs * 2 * READ_LEN, dtype=torch.uint8, device=device)
    synth_reads(sim, first_unit, n_pairs, d_bases, local)
    torch.cuda.synchronize()

    launches = {'n': 0}
    kernel_times = []
    merge_times = []

    def step_device(timed=False):
        mp.reset()
        mp.map_batch(d_bases, None, n_pairs, True, first_unit=first_unit, fixed_len=READ_LEN)
        launches['n'] += 3  # pack_reads_kernel, map_reads_kernel, tally_units_kernel
        if world > 1:
            # one all-gather of the exported dictionaries, peers merged on the device
            stages = {} if timed else None
            table = sdist.merge_mappers(mp, stages=stages)
            launches['n'] += 4 + 2 * (world - 1)  # 2 exports (2 kernels each) + merge + FLD add per peer
            if timed:
                merge_times.append(stages)
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
    kernels_ms = {name: float(numpy.mean([k[name] for k in kernel_times])) for name in kernel_times[0]}
    gpu_launches = launches['n']
    t = torch.tensor([total_ms, kernels_ms['map_reads_kernel'], sum(kernels_ms.values())], dtype=torch.float64,
                     device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, kernel_ms_max, step_kernels_ms_max = float(t[0]), float(t[1]), float(t[2])
    ms_per_step = total_ms / args.steps
    value = total_pairs / (ms_per_step / 1e3)
    host_table = sdist.table_to_host(table)
    merge_split = None
    if merge_times:
        merge_split = {k: round(float(numpy.mean([m[k] for m in merge_times])), 3) for k in merge_times[0]}

    # ---- the N-rank merged table against ONE mapper over the same global units (rank 0 maps them all)
    merged_equals_single = None
    if world > 1:
        if rank == 0:
            single = _lib.DeviceMapper(index, class_capacity=DICT_CLASSES, id_capacity=DICT_IDS)
            scratch = torch.empty_like(d_bases)
            for r in range(world):
                synth_reads(sim, r * n_pairs, n_pairs, scratch, local)
                single.map_batch(scratch, None, n_pairs, True, first_unit=r * n_pairs, fixed_len=READ_LEN)
            one = single.export()
            merged_equals_single = bool(
                table_digest(one) == table_digest(host_table)
                and (one['key_ids'] == host_table['key_ids']).all()  # the same first-seen class order, too
                and (one['counts'] == host_table['counts']).all())
            single.close()
            del scratch
        barrier()

    # ---- end to end through the C ABI with HOST buffers (pinned): H2D copies inside
    # (at most 30 M pairs of the rank's shard: 9 GB of pinned host memory; a rate, so the slice is enough)
    n_e2e = min(n_pairs, 30_000_000)
    h_bases = torch.empty(n_e2e * 2 * READ_LEN, dtype=torch.uint8, pin_memory=True)
    h_bases.copy_(d_bases[:n_e2e * 2 * READ_LEN])
    torch.cuda.synchronize()
    h_np = h_bases.numpy()

    def step_e2e():
        mp.reset()
        batch = args.e2e_batch or n_e2e
        for s in range(0, n_e2e, batch):
            n = min(batch, n_e2e - s)
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
    e2e_s = float(t[0])
    e2e_value = world * n_e2e / e2e_s
    d2h_bytes = int(sum(e2e_table[k].nbytes for k in ('key_offsets', 'key_ids', 'counts', 'first_unit', 'fld')))
    if n_e2e == n_pairs:
        assert (e2e_table['counts'] == host_table['counts']).all() and (e2e_table['key_ids'] == host_table['key_ids']).all()

    # ---- what the box gives when every rank copies its reads host -> device at the same time
    scratch = torch.empty(h_bases.numel(), dtype=torch.uint8, device=device)
    scratch.copy_(h_bases, non_blocking=True)
    barrier()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record()
    for _ in range(2):
        scratch.copy_(h_bases, non_blocking=True)
    c1.record()
    barrier()
    copy_s = c0.elapsed_time(c1) / 2e3
    t = torch.tensor([copy_s], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    h2d_ceiling_gbs = world * h_bases.numel() / float(t[0]) / 1e9
    del scratch

    # ---- EM + bootstraps: the merged dictionary becomes the EM's class structure on the device
    em = None
    if not args.no_em:
        class FakeIndex:
            transcripts = numpy.zeros(lengths.shape[0], dtype=[('length', 'f8')])
        FakeIndex.transcripts['length'] = lengths
        FakeIndex.default_device = local
        if n_e2e != n_pairs:  # the end-to-end pass only mapped a slice: map the whole shard again
            mp.reset()
            mp.map_batch(d_bases, None, n_pairs, True, first_unit=first_unit, fixed_len=READ_LEN)
            if world > 1:
                sdist.merge_mappers(mp)
        # `mp` holds the (merged) dictionary of the whole job: the same classes, counts and first-seen
        # order as `host_table` (asserted above for the end-to-end pass)
        barrier()
        w0 = time.perf_counter()
        plan = _lib.EmPlan.from_mapper(mp, lengths.shape[0])
        mr = mapper.MapResult(FakeIndex)
        mr.fragment_length_counts = host_table['fld'].astype('i8')
        eff = mr.effective_lengths
        torch.cuda.synchronize()
        plan_s = time.perf_counter() - w0
        w0 = time.perf_counter()
        x = numpy.ones(lengths.shape[0]) / eff
        x /= x.sum()
        out, its = plan.run(eff, x)
        main_iters = int(its[0])
        main = infer._finish(out[0])
        em_main_s = time.perf_counter() - w0
        # device time of the iterations alone (one cooperative launch), for the EM roofline
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        d_eff, d_x = torch.from_numpy(eff).to(device), torch.from_numpy(x).to(device)
        d_out = torch.zeros_like(d_x)
        d_it = torch.zeros(1, dtype=torch.int32, device=device)
        torch.cuda.synchronize()
        ev0.record()
        _lib.check(_lib.load().skm_em_plan_run(plan._h, None, _lib._ptr(d_eff), _lib._ptr(d_x), 1, 0, _lib._ptr(d_out),
                                               _lib._ptr(d_it), 1, _lib.current_stream_ptr()))
        ev1.record()
        torch.cuda.synchronize()
        em_device_ms = ev0.elapsed_time(ev1)
        per_rank = (args.bootstraps + world - 1) // world
        r0 = rank * per_rank
        nrep = max(0, min(per_rank, args.bootstraps - r0))
        barrier()
        w0 = time.perf_counter()
        if nrep:
            boots, iters = plan.bootstrap(eff, main / main.sum(), nrep, 1234, first_replicate=r0)
        else:
            boots, iters = numpy.zeros((0, lengths.shape[0])), numpy.zeros(0, dtype='i4')
        if world > 1:
            pad = torch.zeros(per_rank, lengths.shape[0], dtype=torch.float64, device=device)
            pad[:boots.shape[0]] = torch.from_numpy(boots).to(device)
            gathered = torch.zeros(world * per_rank, lengths.shape[0], dtype=torch.float64, device=device)
            dist.all_gather_into_tensor(gathered, pad)
        barrier()
        boot_s = time.perf_counter() - w0
        t = torch.tensor([plan_s, em_main_s, boot_s, em_device_ms], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        plan_s, em_main_s, boot_s, em_device_ms = (float(v) for v in t)
        n_classes, nnz = plan.n_classes, plan.nnz
        b_iter = em_bytes_per_iteration(n_classes, nnz, lengths.shape[0])
        em = {'plan_ms': round(plan_s * 1e3, 3), 'main_ms': round(em_main_s * 1e3, 3), 'main_iters': main_iters,
              'main_device_ms': round(em_device_ms, 3),
              'bootstrap_ms': round(boot_s * 1e3, 3), 'bootstraps': args.bootstraps,
              'bootstrap_iters_mean': float(numpy.mean(iters)) if len(iters) else None,
              'n_classes': int(n_classes), 'nnz': int(nnz), 'n_transcripts': int(lengths.shape[0]),
              'em_plus_bootstraps_ms': round((plan_s + em_main_s + boot_s) * 1e3, 3),
              'what': 'plan = class structure made from the merged device dictionary (device to device) + effective '
                      'lengths; main = x0 up, one cooperative launch, TPM down; bootstraps = resample + batched EM + '
                      'TPM step on the device, replicates sharded over the ranks and all-gathered',
              'main_x': main, 'eff': eff}
        peak = _peak_hbm()[0]
        gbs = main_iters * b_iter / (em_device_ms / 1e3) / 1e9
        em['roofline'] = {'bound': 'hbm', 'kernel': 'em_loop_kernel', 'bytes_per_iteration': int(b_iter),
                          'iterations': main_iters, 'achieved': round(gbs, 1), 'peak': peak, 'unit': 'GB/s',
                          'frac': round(gbs / peak, 4),
                          'note': 'B_iter = 8 nnz + 20 C + 36 T (SURVEY 8(d)); the structure fits L2, so this is a '
                                  'latency / grid-barrier bound, not an HBM one'}
        plan.close()

    # ---- the product's FASTQ path end to end (rank 0, N = 1): files in the page cache -> classes
    e2e_fastq = None
    if world == 1 and args.fastq_pairs > 0:
        try:
            n_fq = min(args.fastq_pairs, n_e2e)
            with tempfile.TemporaryDirectory(prefix='skm_bench_') as folder:
                paths, fq_bytes = write_fastq_pair(folder, h_np, n_fq, READ_LEN)

                class Idx:
                    transcripts = numpy.zeros(lengths.shape[0], dtype=[('length', 'f8')])
                    default_device = local

                    def device_index(self, device=0):
                        return index
                rates = []
                for _ in range(2):  # the first pass also pulls the files into the page cache
                    w0 = time.perf_counter()
                    res = mapper.map_reads(Idx(), common.feed_pair_ended_reads(*paths))
                    rates.append(n_fq / (time.perf_counter() - w0))
                fq_table = res._table
                mp.reset()
                mp.map_batch(d_bases[:n_fq * 2 * READ_LEN], None, n_fq, True, first_unit=0, fixed_len=READ_LEN)
                want = mp.export()
                e2e_fastq = {'value': round(rates[-1], 1), 'unit': UNIT, 'pairs': n_fq, 'fastq_bytes': int(fq_bytes),
                             'what': 'mapper.map_reads(index, feed_pair_ended_reads(file_1, file_2)): files read from '
                                     'the page cache, raw text to the GPU, parsed and mapped there (skm_map_fastq)',
                             'equals_batch_path': bool((fq_table['counts'] == want['counts']).all()
                                                       and (fq_table['key_ids'] == want['key_ids']).all()
                                                       and (fq_table['fld'] == want['fld']).all())}
                if res._plan is not None:
                    res._plan.close()
        except Exception as exc:
            log('FASTQ leg skipped: %r' % (exc,))

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
        sample = min(args.cpu_sample, n_e2e)
        hb = h_np[:sample * 2 * READ_LEN]
        b_step, b_kernel, per_read = algorithmic_bytes_per_pair(orc, oidx, hb, min(100_000, sample))
        peak, peak_src = _peak_hbm()
        achieved_kernel = n_pairs * b_kernel / (kernel_ms_max / 1e3) / 1e9
        achieved_step = n_pairs * b_step / (step_kernels_ms_max / 1e3) / 1e9
        # dram__bytes_read.sum + dram__bytes_write.sum of map_reads_kernel from the committed
        # ncu --set full capture (profiles/): bytes per pair there x the pairs of one launch here
        traffic, traffic_src = None, None
        try:
            tr = json.load(open(os.path.join(ROOT, 'profiles', 'map_reads_kernel_traffic.json')))
            traffic = round(tr['dram_bytes_per_pair'] * n_pairs)
            traffic_src = tr['source']
        except Exception:
            pass
        roofline = {'bound': 'hbm', 'kernel': 'map_reads_kernel', 'achieved': round(achieved_kernel, 2), 'peak': peak,
                    'peak_source': peak_src, 'unit': 'GB/s',
                    'frac': round(achieved_kernel / peak, 4), 'frac_kernel': round(achieved_kernel / peak, 4),
                    'frac_step': round(achieved_step / peak, 4), 'achieved_step': round(achieved_step, 2),
                    'traffic': traffic, 'traffic_source': traffic_src,
                    'algorithmic_bytes_per_pair': round(b_kernel, 1),
                    'algorithmic_bytes_per_pair_step': round(b_step, 1),
                    'bytes_note': 'kernel: 2-bit packed reads in (L/4 per read) + the index-side bytes of SURVEY 8(d); '
                                  'step (pack + map + tally): ASCII reads in (L per read) + the same index-side bytes',
                    'pairs_per_launch': n_pairs,
                    'kernel_ms': round(kernel_ms_max, 3), 'step_kernels_ms': {k: round(v, 3) for k, v in kernels_ms.items()},
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
            if em is not None and world == 1:  # a baseline of rank 0 at N = 1 (12 s of numpy EM)
                em['cpu_baseline'] = em_cpu_baseline(orc, host_table, em['eff'], em['main_x'], em['main_iters'],
                                                     args.cpu_bootstraps, args.bootstraps, 'port')
    except Exception as exc:  # the oracle is optional infrastructure for the bench line
        log('oracle leg skipped: %r' % (exc,))
    if em is not None:
        em.pop('main_x', None)
        em.pop('eff', None)

    workload = ('human-scale synthetic transcriptome (%d transcripts, %.0f Mb cDNA) + %d M 2x%d bp pairs %s, 1%% subs'
                % (args.transcripts, sim['codes'].shape[0] / 1e6, (total_pairs if strong else n_pairs) // 1_000_000,
                   READ_LEN, 'in total' if strong else 'per GPU'))
    line = {
        'metric': METRIC, 'value': round(value, 1), 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': round(ms_per_step, 3), 'higher_is_better': True,
        'scaling': args.scaling, 'vs_baseline': None, 'dtype': 'u64', 'data': 'synthetic',
        'config': {'workload': workload,
                   'pairs_per_gpu': n_pairs, 'total_pairs': total_pairs, 'read_len': READ_LEN,
                   'index_kmers': info['n_kmers'],
                   'index_table_slots': info['table_slots'], 'index_device_bytes': info['device_bytes'],
                   'n_contigs': info['n_contigs'], 'l2_policy': 'inputs (%.1f GB reads + %.1f GB table) larger than L2'
                   % (d_bases.numel() / 1e9, info['table_slots'] * 16 / 1e9),
                   'parallelism': 'reads sharded x%d, index replicated' % world, 'host_placement': placement},
        'clocks': clocks,
        'e2e': {'value': round(e2e_value, 1), 'unit': UNIT, 'h2d_bytes_per_step': int(n_e2e * 2 * READ_LEN),
                'pairs_per_gpu': n_e2e,
                'd2h_bytes_per_step': d2h_bytes, 'ms_per_step': round(e2e_s * 1e3, 3),
                'h2d_ceiling_gbs': round(h2d_ceiling_gbs, 2),
                'h2d_achieved_gbs': round(world * n_e2e * 2 * READ_LEN / e2e_s / 1e9, 2),
                'frac_of_h2d_ceiling': round(world * n_e2e * 2 * READ_LEN / e2e_s / 1e9 / h2d_ceiling_gbs, 3),
                'ceiling_note': 'all %d ranks copying their pinned read buffer host -> device at the same time '
                                '(cudaMemcpyAsync, CUDA events, slowest rank)' % world},
        'gpu_launches': gpu_launches,
        'classes': {'n_classes': int(host_table['counts'].shape[0]), 'aligned': host_table['aligned'],
                    'unaligned': host_table['unaligned']},
    }
    if merge_split:
        line['merge_ms'] = merge_split
    if merged_equals_single is not None:
        line['merged_equals_single'] = merged_equals_single
    if e2e_fastq:
        line['e2e_fastq'] = e2e_fastq
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


def _peak_hbm():
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    if 'hbm_gbs' in peaks:
        return float(peaks['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs, burst)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


def class_structure_of(host_table):
    """`MapResult.summarize` shape (`mapper.py:77-104`) of an exported table: class_map int64 (2, nnz),
    class_count f8."""
    off = numpy.asarray(host_table['key_offsets'], dtype='i8')
    sizes = off[1:] - off[:-1]
    class_map = numpy.stack([numpy.repeat(numpy.arange(sizes.shape[0], dtype='i8'), sizes),
                             numpy.asarray(host_table['key_ids'], dtype='i8')])
    return class_map, numpy.asarray(host_table['counts'], dtype='f8')


def em_cpu_baseline(orc, host_table, eff, gpu_main, gpu_iters, n_boot, n_total, kind):
    """The numpy restatement of infer.quantify on the host: the main EM once, `n_boot` bootstrap
    replicates, the bootstrap time extrapolated linearly to `n_total` replicates."""
    class_map, class_count = class_structure_of(host_table)
    w0 = time.perf_counter()
    tpm, iters = orc.quantify(eff, class_map, class_count, return_iters=True)
    main_s = time.perf_counter() - w0
    counts = orc.bootstrap_counts(class_count.astype('i8'), n_boot, 1234)
    w0 = time.perf_counter()
    for r in range(n_boot):
        orc.quantify(eff, class_map, counts[r].astype('f8'), x0=tpm)
    boot_s = (time.perf_counter() - w0) / max(n_boot, 1)
    ok = bool(numpy.allclose(gpu_main, tpm, rtol=1e-6, atol=0)) and int(iters) == int(gpu_iters)
    return {'main_ms': round(main_s * 1e3, 1), 'main_iters': int(iters), 'cores': 1, 'kind': kind,
            'bootstrap_ms_per_replicate': round(boot_s * 1e3, 1),
            'bootstrap_ms_extrapolated': round(boot_s * n_total * 1e3, 1),
            'em_plus_bootstraps_ms': round((main_s + boot_s * n_total) * 1e3, 1),
            'sample': 'main EM in full; %d bootstrap replicates timed, x %d / %d for the %d the GPU ran (the '
                      'reference runs them one after the other, infer.py:79-82)' % (n_boot, n_total, n_boot, n_total),
            'gpu_main_matches': ok}


# ------------------------------------------------------------------ config 5
def run_c5(args):
    """BASELINE configs[4]: 16 single-end 75 bp samples at 2 % substitutions, samples dealt to the
    ranks round-robin, every sample mapped and quantified (main EM) where its reads are."""
    import torch
    import torch.distributed as dist
    from seekmer_b200 import _lib, infer, mapper

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    device = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=device)
    built, sim, lengths = make_workload(args, rank, world, device)
    index = _lib.DeviceIndex(built.kmers, built.contigs, built.sequences, built.targets, built.n_transcripts)
    mp = _lib.DeviceMapper(index, class_capacity=1 << 21, id_capacity=1 << 25)
    mine = list(range(rank, args.samples, world))
    n = args.sample_reads
    d_reads = [torch.empty(n * C5_READ_LEN, dtype=torch.uint8, device=device) for _ in mine]
    for buf, s in zip(d_reads, mine):
        synth_reads(sim, 0, n, buf, local, read_len=C5_READ_LEN, frag_mean=250, frag_sd=30, sub_rate=C5_SUB_RATE,
                    paired=False, seed=100 + s)
    torch.cuda.synchronize()

    class FakeIndex:
        transcripts = numpy.zeros(lengths.shape[0], dtype=[('length', 'f8')])
    FakeIndex.transcripts['length'] = lengths
    n_tx = lengths.shape[0]
    stage = {'map_ms': 0.0, 'plan_ms': 0.0, 'em_ms': 0.0}

    def step(keep=False):
        # one sample after the other: map, plan (device to device), main EM on the plan.  The EMs of a
        # rank's samples could iterate side by side (`EmPlan.run_many` / skm_em_plans_run), but for
        # samples of this size that is no faster at one GPU and slower at eight (two samples per GPU:
        # 10.3 against 8.5 ms per sample) - an EM iteration is bound by its L2 gathers, not by launches
        results = []
        for buf in d_reads:
            w0 = time.perf_counter()
            mp.reset()
            mp.map_batch(buf, None, n, False, first_unit=0, fixed_len=C5_READ_LEN)
            sz = mp.sizes()  # synchronises
            w1 = time.perf_counter()
            plan = _lib.EmPlan.from_mapper(mp, n_tx)
            fld = torch.zeros(2000, dtype=torch.int64, device=device)
            _lib.check(_lib.load().skm_classes_export(mp._h, None, None, None, None, None, _lib._ptr(fld), 1,
                                                      _lib.current_stream_ptr()))
            mr = mapper.MapResult(FakeIndex)
            mr.fragment_length_counts = fld.cpu().numpy()
            eff = mr.effective_lengths
            w2 = time.perf_counter()
            x = numpy.ones(n_tx) / eff
            x /= x.sum()
            out, its = plan.run(eff, x)
            tpm = infer._finish(out[0])
            w3 = time.perf_counter()
            plan.close()
            stage['map_ms'] += (w1 - w0) * 1e3
            stage['plan_ms'] += (w2 - w1) * 1e3
            stage['em_ms'] += (w3 - w2) * 1e3
            if keep:
                results.append((sz, int(its[0]), tpm, eff))
        return results

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(max(args.warmup, 1)):
        step()
    barrier()
    for k in stage:
        stage[k] = 0.0
    sampler.mark_begin()
    w0 = time.perf_counter()
    for _ in range(args.steps):
        results = step(keep=True)
    barrier()
    dt = time.perf_counter() - w0
    sampler.mark_end()
    clocks = sampler.stop()
    t = torch.tensor([dt], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    s_per_step = float(t[0]) / args.steps
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # parity of the first sample's mapping + EM against the oracle on a bounded slice (rank 0)
    parity = None
    try:
        from oracle import oracle as orc
        oidx = orc.OracleIndex(*built.numpy_arrays())
        m = min(200_000, n)
        hb = d_reads[0][:m * C5_READ_LEN].cpu().numpy()
        offs = numpy.arange(m + 1, dtype='i8') * C5_READ_LEN
        want = orc.map_batch(oidx, hb, offs, False)
        mp.reset()
        mp.map_batch(d_reads[0][:m * C5_READ_LEN], None, m, False, first_unit=0, fixed_len=C5_READ_LEN)
        got = mp.export()
        cls_ptr, cls_ids, cls_count, una = orc.tally(want.ptr, want.ids)
        parity = {'sample_reads': m, 'classes_equal': bool((got['key_offsets'] == cls_ptr).all()
                                                           and (got['key_ids'] == cls_ids).all()
                                                           and (got['counts'] == cls_count).all()),
                  'fld_equal': bool((got['fld'] == want.fld).all()), 'unaligned_equal': bool(got['unaligned'] == una)}
        eff = orc.effective_lengths(want.fld, lengths.astype('f8'))
        ref_tpm, ref_it = orc.quantify(eff, orc.class_map_from_csr(cls_ptr, cls_ids), cls_count, return_iters=True)
        plan = _lib.EmPlan.from_mapper(mp, n_tx)
        x = numpy.ones(n_tx) / eff
        x /= x.sum()
        out, its = plan.run(eff, x)
        plan.close()
        parity['tpm_rel_1e-6'] = bool(numpy.allclose(infer._finish(out[0]), ref_tpm, rtol=1e-6, atol=0))
        parity['em_iterations_equal'] = bool(int(its[0]) == int(ref_it))
    except Exception as exc:
        log('oracle leg skipped: %r' % (exc,))
    per = args.steps * len(mine)
    line = {
        'metric': 'samples_mapped_and_quantified_per_sec', 'value': round(args.samples / s_per_step, 3),
        'unit': 'samples/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': round(s_per_step * 1e3, 3), 'higher_is_better': True, 'scaling': 'strong',
        'vs_baseline': None, 'dtype': 'u64', 'data': 'synthetic',
        'config': {'workload': 'config 5: %d samples x %d M single-end %d bp reads, %.0f %% substitutions, human-scale '
                               'synthetic index; samples dealt round-robin to the GPUs, each mapped and quantified '
                               '(main EM) on its GPU' % (args.samples, n // 1_000_000, C5_READ_LEN, 100 * C5_SUB_RATE),
                   'reads_per_sample': n, 'samples': args.samples,
                   'l2_policy': 'inputs (%.1f GB reads per sample + the 4.3 GB table) larger than L2'
                   % (n * C5_READ_LEN / 1e9)},
        'reads_per_sec': round(args.samples * n / s_per_step, 1),
        'per_sample_ms': {k: round(v / per, 3) for k, v in stage.items()},
        'first_sample': {'classes': results[0][0]['n_classes'], 'aligned': results[0][0]['aligned'],
                         'unaligned': results[0][0]['unaligned'], 'em_iterations': results[0][1]},
        'clocks': clocks, 'gpu_launches': args.steps * len(mine) * 18,
    }
    if parity:
        line['parity_on_sample'] = parity
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------ reference arm
def make_reference_workload(args):
    """Child process of the reference arm: the workload generator needs this repo's CUDA library
    (index construction, read synthesis, and the class structure of the full 30 M pairs for the EM
    leg); it runs HERE, writes an .npz and exits, so that the timed process never loads it."""
    import torch
    from seekmer_b200 import _lib, mapper
    device = torch.device('cuda', 0)
    torch.cuda.set_device(0)
    built, sim, lengths = make_workload(args, 0, 1, device)
    sample = min(args.cpu_sample, args.pairs)
    d_bases = torch.empty(args.pairs * 2 * READ_LEN, dtype=torch.uint8, device=device)
    synth_reads(sim, 0, args.pairs, d_bases, 0)
    index = _lib.DeviceIndex(built.kmers, built.contigs, built.sequences, built.targets, built.n_transcripts)
    mp = _lib.DeviceMapper(index, class_capacity=1 << 22, id_capacity=1 << 26)
    mp.map_batch(d_bases, None, args.pairs, True, first_unit=0, fixed_len=READ_LEN)
    table = mp.export()

    class FakeIndex:
        transcripts = numpy.zeros(lengths.shape[0], dtype=[('length', 'f8')])
    FakeIndex.transcripts['length'] = lengths
    mr = mapper.MapResult(FakeIndex)
    mr.fragment_length_counts = table['fld'].astype('i8')
    eff = mr.effective_lengths
    kmers, contigs, sequences, targets = built.numpy_arrays()
    numpy.savez(args.make_workload, kmers=kmers, contigs=contigs, sequences=sequences, targets=targets,
                lengths=lengths, bases=d_bases[:sample * 2 * READ_LEN].cpu().numpy(),
                key_offsets=table['key_offsets'], key_ids=table['key_ids'], counts=table['counts'], eff=eff,
                aligned=table['aligned'], unaligned=table['unaligned'], mb=numpy.asarray(sim['codes'].shape[0] / 1e6))


def run_reference(args):
    """The reference's own CPU implementation of the path on this box's host cores: stock
    `mapper.map_reads(index, feeder, job_count=cores)` and `infer.quantify` from oracle/_ref (the
    unmodified reference, compiled / placed by oracle/build_ref.py).  Nothing of this repo's CUDA
    library is loaded into this process."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    from oracle import build_ref
    cores = os.cpu_count() or 1
    sample = min(args.cpu_sample, args.pairs)
    with tempfile.TemporaryDirectory(prefix='skm_ref_') as folder:
        path = os.path.join(folder, 'workload.npz')
        env = dict(os.environ)
        for k in ('RANK', 'WORLD_SIZE', 'LOCAL_RANK', 'MASTER_ADDR', 'MASTER_PORT'):
            env.pop(k, None)
        cmd = [sys.executable, os.path.abspath(__file__), '--make-workload', path, '--pairs', str(args.pairs),
               '--transcripts', str(args.transcripts), '--cpu-sample', str(sample)]
        subprocess.run(cmd, check=True, env=env)
        z = numpy.load(path)
        w = {k: z[k] for k in z.files}
    bases = w['bases']
    kind, run, em_leg = 'port', None, None
    if build_ref.built():
        try:
            import importlib
            from oracle import ref_harness as rh
            rh.load_ref()
            ridx = rh.ref_index_from_arrays(w['kmers'], w['contigs'], w['sequences'], w['targets'])
            raw = bases.tobytes()
            bsz = 65536  # common.BUFFER_SIZE
            batches = []
            for s in range(0, sample, bsz):
                n = min(bsz, sample - s)
                reads = [raw[(2 * s + i) * READ_LEN:(2 * s + i + 1) * READ_LEN] for i in range(2 * n)]
                batches.append((n, [b''] * n, reads))
            if rh.have_python_reference():
                ref_mapper = importlib.import_module('seekmer.mapper')
                ref_infer = importlib.import_module('seekmer.infer')
                kind = 'reference'

                def run():
                    res = ref_mapper.map_reads(ridx, iter(batches), job_count=cores)  # the stock function
                    return sum(v for k, v in res.counter.items() if k)

                def em_leg():
                    class_map, class_count = class_structure_of(w)
                    summ = ref_mapper.SummarizedResult(int(w['aligned']), int(w['unaligned']),
                                                       int(w['aligned']) + int(w['unaligned']), class_map, class_count,
                                                       None, w['eff'])
                    t0 = time.perf_counter()
                    main = ref_infer.quantify(summ)
                    main_s = time.perf_counter() - t0
                    t0 = time.perf_counter()
                    for _ in range(args.cpu_bootstraps):
                        ref_infer.quantify(summ, x0=main, bootstrap=True)
                    boot_s = (time.perf_counter() - t0) / max(args.cpu_bootstraps, 1)
                    return main_s, boot_s, 'stock infer.quantify (numpy bincount EM, scipy multinomial)'
            else:
                kind = 'reference'

                def run():
                    res = rh.ref_map_threads(ridx, batches, cores)  # mapper.map_reads threading model on the natives
                    return sum(v for k, v in res.counter.items() if k)
        except Exception as exc:
            log('compiled reference unavailable (%r); timing the C port' % (exc,))
            run = None
    from oracle import oracle as orc
    if run is None:
        oidx = orc.OracleIndex(w['kmers'], w['contigs'], w['sequences'], w['targets'])
        offs = numpy.arange(2 * sample + 1, dtype='i8') * READ_LEN

        def run():
            return orc.map_batch_mt(oidx, bases, offs, True, cores)[0]
    if em_leg is None:
        def em_leg():
            class_map, class_count = class_structure_of(w)
            t0 = time.perf_counter()
            tpm = orc.quantify(w['eff'], class_map, class_count)
            main_s = time.perf_counter() - t0
            counts = orc.bootstrap_counts(class_count.astype('i8'), args.cpu_bootstraps, 1234)
            t0 = time.perf_counter()
            for r in range(args.cpu_bootstraps):
                orc.quantify(w['eff'], class_map, counts[r].astype('f8'), x0=tpm)
            return main_s, (time.perf_counter() - t0) / max(args.cpu_bootstraps, 1), 'numpy port of infer.quantify'
    for _ in range(min(args.warmup, 1)):
        run()
    steps = max(1, min(args.steps, 3))
    w0 = time.perf_counter()
    for _ in range(steps):
        aligned = run()
    dt = (time.perf_counter() - w0) / steps
    value = sample / dt
    em = None
    if not args.no_em:
        main_s, boot_s, how = em_leg()
        em = {'main_ms': round(main_s * 1e3, 1), 'bootstrap_ms_per_replicate': round(boot_s * 1e3, 1),
              'bootstraps': args.bootstraps, 'bootstrap_ms': round(boot_s * args.bootstraps * 1e3, 1),
              'em_plus_bootstraps_ms': round((main_s + boot_s * args.bootstraps) * 1e3, 1),
              'n_classes': int(w['counts'].shape[0]), 'nnz': int(w['key_ids'].shape[0]),
              'n_transcripts': int(w['lengths'].shape[0]), 'cores': 1,
              'sample': '%s on the class structure of all %d M pairs: main EM in full, %d bootstrap replicates '
                        'timed, x %d / %d (they run one after the other, infer.py:79-82)'
                        % (how, args.pairs // 1_000_000, args.cpu_bootstraps, args.bootstraps, args.cpu_bootstraps)}
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': round(value, 1), 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': steps, 'warmup': min(args.warmup, 1), 'ms_per_step': round(dt * 1e3, 3), 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'u64', 'data': 'synthetic',
        'config': {'workload': 'human-scale synthetic transcriptome (%d transcripts, %.0f Mb cDNA) + %d M 2x%d bp '
                               'pairs per GPU, 1%% subs' % (args.transcripts, float(w['mb']),
                                                            args.pairs // 1_000_000, READ_LEN),
                   'sample_pairs_per_step': sample, 'aligned': int(aligned), 'reference_class': 'cpu'},
        'cpu_baseline': {'value': round(value, 1), 'unit': UNIT, 'cores': cores, 'kind': kind,
                         'sample': '%d of the same 2x%d pairs per step; %s' % (
                             sample, READ_LEN,
                             'unmodified reference (oracle/_ref): stock mapper.map_reads(index, feeder, job_count=%d), '
                             'pre-materialised feeder batches' % cores if kind == 'reference'
                             else 'C oracle (OpenMP, %d threads)' % cores)},
        'e2e': {'value': round(value, 1), 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    if em:
        line['em'] = em
    print(json.dumps(line), flush=True)


def main():
    args = parse_args()
    if args.make_workload:
        make_reference_workload(args)
    elif args.impl == 'reference':
        run_reference(args)
    elif args.config == 'c5':
        run_c5(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
