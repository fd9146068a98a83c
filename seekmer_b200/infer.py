"""Inference workflow — the `seekmer.infer` surface (`infer.py:27-353`).

Mirrors: run, quantify, em, output_results, add_subcommand_parser — same signatures, same
output folder (run_info.json, abundance.tsv, abundance.h5, optional readmap.txt).

What changes underneath: `em` runs as fp64 segmented-reduction kernels (`skm_em`); the
bootstrap resampling of `quantify(..., bootstrap=True)` is an on-device multinomial
(`skm_multinomial`); and `run` batches all bootstrap replicates through one EM call
(`quantify_bootstraps`) instead of the reference's serial loop (`infer.py:79-82`);
`quantify_samples` does the same for many samples with their own class structures
(`skm_em_samples`, used by `impute`).  There is no CPU EM fallback.
"""
import datetime
import json
import pathlib
import shlex
import sys

import numpy

from . import _lib
from . import common
from . import mapper
from ._log import Logger, nvtx_range

__all__ = ['run', 'quantify', 'quantify_bootstraps', 'quantify_samples', 'em', 'output_results',
           'add_subcommand_parser']

_LOG = Logger(__name__)


def run(index_path, output_path, fastq_paths, job_count, save_readmap, single_ended, bootstrap,
        debug, **__):
    """The entrypoint of the inference module (`infer.py:27-85`)."""
    start_time = datetime.datetime.now(datetime.timezone.utc).replace(tzinfo=None)  # utcnow() of `infer.py:48`
    try:
        output_path.mkdir(parents=True)
    except FileExistsError:
        _LOG.warn('The output folder exists. Overriding...')
    readmap = (output_path / 'readmap.txt').open('wt') if save_readmap else None
    _LOG.info('Inferring transcript abundance')
    index = common.KMerIndex.load(index_path)
    _LOG.info('Mapping all reads')
    if single_ended:
        read_feeder = common.feed_single_ended_reads(*fastq_paths)
    else:
        read_feeder = common.feed_pair_ended_reads(*fastq_paths)
    with nvtx_range('seekmer: map reads'):
        map_result = mapper.map_reads(index, read_feeder, job_count=job_count, readmap=readmap,
                                      debug=debug)
    _LOG.info('Mapped all reads')
    mean_fragment_length = map_result.harmonic_mean_fragment_length
    _LOG.info('Estimated fragment length: {:.2f}', mean_fragment_length)
    summarized_results = map_result.summarize()
    _LOG.info('Quantifying transcripts')
    _LOG.info('Aligned {} reads ({:.2%})', summarized_results.aligned,
              summarized_results.aligned / max(summarized_results.total, 1))
    with nvtx_range('seekmer: EM'):
        main_result = quantify(summarized_results)
    _LOG.info('Quantified transcripts')
    # `-j N`: the replicates of `infer.py:79-82` are independent; they are dealt to N GPUs
    with nvtx_range('seekmer: bootstraps'):
        bootstrapped_results = quantify_bootstraps(summarized_results, main_result, bootstrap,
                                                   devices=mapper._devices_for(job_count))
    output_results(output_path, index, start_time, summarized_results, main_result,
                   bootstrapped_results)
    _LOG.info('Wrote results to {}'.format(output_path))


def _csr_from_class_map(class_map, n_classes):
    """(2, nnz) class_map -> (ptr int64[C+1], tx int32[nnz]) keeping nnz order within a class."""
    rows = numpy.asarray(class_map[0], dtype='i8')
    cols = numpy.asarray(class_map[1], dtype='i8')
    if rows.size > 1 and (rows[1:] < rows[:-1]).any():
        order = numpy.argsort(rows, kind='stable')
        rows, cols = rows[order], cols[order]
    ptr = numpy.zeros(n_classes + 1, dtype='i8')
    numpy.cumsum(numpy.bincount(rows, minlength=n_classes), out=ptr[1:])
    return ptr, numpy.ascontiguousarray(cols, dtype='i4')


def _em_device(x0, l, class_map, class_counts, device=0):
    """x0: (R, T) normalised guesses; class_counts: (R, C). Returns (x (R, T), iters (R,))."""
    x0 = numpy.ascontiguousarray(numpy.atleast_2d(x0), dtype='f8')
    counts = numpy.ascontiguousarray(numpy.atleast_2d(class_counts), dtype='f8')
    l = numpy.ascontiguousarray(l, dtype='f8')
    n_rep, n_tx = x0.shape
    n_classes = counts.shape[1]
    ptr, tx = _csr_from_class_map(class_map, n_classes)
    out = numpy.zeros_like(x0)
    iters = numpy.zeros(n_rep, dtype='i4')
    _lib.require_device()
    _lib.check(_lib.load().skm_em(
        _lib._np_ptr(ptr), _lib._np_ptr(tx), n_classes, tx.shape[0], _lib._np_ptr(counts),
        _lib._np_ptr(l), n_tx, _lib._np_ptr(x0), n_rep, 0, _lib._np_ptr(out), _lib._np_ptr(iters),
        0, device, None))
    return out, iters


def em(x, l, class_map, class_count, return_iters=False):
    """Expectation-maximization (`infer.py:133-168`) on the GPU; fp64, same stop rule."""
    out, iters = _em_device(x, l, class_map, numpy.asarray(class_count, dtype='f8'))
    return (out[0], int(iters[0])) if return_iters else out[0]


def _finish(x):
    """TPM post-processing of `infer.py:127-129`."""
    with numpy.errstate(all='ignore'):
        x /= x.sum() / 1000000
        x[x < 0.001] = 0
        x /= x.sum() / 1000000
    return x


def _resample(class_count, n_replicates, seed, first_replicate=0, device=0, method=_lib.RESAMPLE_TREE):
    """Multinomial(n, count / n) for `n_replicates` replicates (`infer.py:108-111`) on the device."""
    counts = numpy.ascontiguousarray(class_count, dtype='i8')
    if not (counts == class_count).all():
        raise ValueError('bootstrap needs integral class counts')
    out = numpy.zeros((n_replicates, counts.shape[0]), dtype='i8')
    _lib.require_device()
    _lib.check(_lib.load().skm_multinomial(_lib._np_ptr(counts), counts.shape[0], n_replicates,
                                           first_replicate, int(seed) & (2 ** 64 - 1), int(method),
                                           _lib._np_ptr(out), 0, device, None))
    return out


def _draw_seed():
    # the reference resamples from numpy's global RNG (`infer.py:111`); take the seed from the
    # same stream so `numpy.random.seed(...)` makes runs repeatable
    return int(numpy.random.randint(0, 2 ** 31 - 1)) | (int(numpy.random.randint(0, 2 ** 31 - 1)) << 31)


def _plan_of(results, device=None):
    """The device-resident class structure of `results` (made by the mapper, device to device),
    or None when the result did not come from the device mapper or sits on another GPU."""
    plan = getattr(results, 'plan', None)
    if plan is None or not plan._h or (device is not None and plan.device != device):
        return None
    return plan


def quantify(results, x0=None, bootstrap=False, seed=None, return_iters=False):
    """Estimate the transcript abundance (`infer.py:88-130`)."""
    transcript_length = results.effective_lengths.astype('f8')
    if results.class_map.size == 0:
        z = numpy.zeros(results.effective_lengths.size).astype('f8')
        return (z, 0) if return_iters else z
    if x0 is None:
        x = numpy.ones(transcript_length.size, dtype='f8') / transcript_length
    else:
        x = x0.copy()
    x /= x.sum()
    plan = _plan_of(results)
    if bootstrap:
        if seed is None:
            seed = _draw_seed()
        if plan is not None:
            out, iters = plan.bootstrap(transcript_length, x, 1, seed)
            return (out[0], int(iters[0])) if return_iters else out[0]
        class_count = _resample(results.class_count, 1, seed)[0].astype('f8')
    else:
        class_count = results.class_count
    if plan is not None and not bootstrap:
        # class structure and counts are already in HBM: nothing but x and the lengths go up
        out, iters = plan.run(transcript_length, x)
        x, iters = out[0], int(iters[0])
    else:
        x, iters = em(x, transcript_length, results.class_map, class_count, return_iters=True)
    x = _finish(x)
    return (x, iters) if return_iters else x


def quantify_bootstraps(results, x0, n_replicates, seed=None, return_iters=False,
                        first_replicate=0, devices=None):
    """`[quantify(results, x0=x0, bootstrap=True) for _ in range(n)]` (`infer.py:79-82`) as one
    batched resample + one batched EM per GPU.  Replicate r is a pure function of (seed, r)
    (counter-based resampler), so dealing contiguous replicate ranges to `devices` gives exactly
    the arrays one GPU would.  Returns a list of n arrays."""
    if n_replicates <= 0:
        return ([], numpy.zeros(0, dtype='i4')) if return_iters else []
    transcript_length = results.effective_lengths.astype('f8')
    if results.class_map.size == 0:
        z = [numpy.zeros(transcript_length.size, dtype='f8') for _ in range(n_replicates)]
        return (z, numpy.zeros(n_replicates, dtype='i4')) if return_iters else z
    if seed is None:
        seed = _draw_seed()
    counts = numpy.ascontiguousarray(results.class_count, dtype='i8')
    if not (counts == results.class_count).all():
        raise ValueError('bootstrap needs integral class counts')
    x = numpy.ascontiguousarray(x0, dtype='f8').copy()
    x /= x.sum()
    _lib.require_device()
    devices = list(devices) if devices else [0]
    devices = devices[:max(1, min(len(devices), n_replicates))]
    per = (n_replicates + len(devices) - 1) // len(devices)
    shares = [(d, k * per, max(0, min(per, n_replicates - k * per))) for k, d in enumerate(devices)]
    xs = numpy.zeros((n_replicates, x.shape[0]), dtype='f8')
    iters = numpy.zeros(n_replicates, dtype='i4')
    csr = []

    def run(device, start, n):
        if n <= 0:
            return
        plan, own = _plan_of(results, device), False
        if plan is None:  # a replica of the class structure on this GPU
            if not csr:
                csr.append(_csr_from_class_map(results.class_map, counts.shape[0]))
            plan, own = _lib.EmPlan.from_csr(csr[0][0], csr[0][1], x.shape[0], device=device), True
        try:
            # resample + EM + TPM step for this share in one device-resident call
            o, it = plan.bootstrap(transcript_length, x, n, seed, first_replicate=first_replicate + start,
                                   counts=None if plan.owns_counts else counts)
        finally:
            if own:
                plan.close()
        xs[start:start + n] = o
        iters[start:start + n] = it

    if len(shares) == 1:
        run(*shares[0])
    else:
        import threading
        _lib.uses_peer_gpus()
        if any(_plan_of(results, d) is None for d in devices):
            csr.append(_csr_from_class_map(results.class_map, counts.shape[0]))
        errors = []

        def guarded(*a):
            try:
                run(*a)
            except BaseException as exc:  # noqa: BLE001 - re-raised below
                errors.append(exc)
        threads = [threading.Thread(target=guarded, args=s) for s in shares]
        for th in threads:
            th.start()
        for th in threads:
            th.join()
        if errors:
            raise errors[0]
    out = list(xs)  # TPM post-processing (`infer.py:127-129`) already applied on the device
    return (out, iters) if return_iters else out


# transcript rows (samples x transcripts) handed to one skm_em_samples call: ~40 B of device
# memory per row, and row indices are 32-bit
_SAMPLE_ROWS_PER_CALL = 1 << 27
# samples with more classes than this (on average) run one after the other on their plans; smaller
# ones (single cells) share launches, where the launch and barrier latency is what they would pay
_BATCH_CLASSES_PER_SAMPLE = 100_000


def quantify_samples(results_list, return_iters=False, device=0):
    """`[quantify(r) for r in results_list]` for samples that share the transcript set (the first
    round of `impute.py:101`): every sample keeps its own class structure, counts, effective
    lengths and stopping point, but all of them iterate in the same launches
    (`skm_em_samples`).  Bit-identical to the loop."""
    n_tx = results_list[0].effective_lengths.size if results_list else 0
    out = numpy.zeros((len(results_list), n_tx), dtype='f8')
    iters = numpy.zeros(len(results_list), dtype='i4')
    live = [i for i, r in enumerate(results_list) if r.class_map.size]  # `infer.py:104-105`
    per_call = max(1, _SAMPLE_ROWS_PER_CALL // max(n_tx, 1))
    # samples whose class structure is still on a GPU as a plan (the mapper left it there,
    # `map_multiple_samples`) iterate on those plans, one call per GPU: nothing but lengths and
    # first guesses goes up; the others take the host CSR route
    by_device = {}
    for i in live:
        plan = _plan_of(results_list[i])
        if plan is not None and plan.owns_counts and plan.n_transcripts == n_tx:
            by_device.setdefault(plan.device, []).append(i)
    planned = set()
    for dev, members in by_device.items():
        for start in range(0, len(members), per_call):
            chunk = members[start:start + per_call]
            lengths = numpy.stack([results_list[i].effective_lengths.astype('f8') for i in chunk])
            x0 = numpy.empty_like(lengths)
            for k in range(len(chunk)):  # row by row: the same arithmetic as `quantify`
                x0[k] = numpy.ones(n_tx, dtype='f8') / lengths[k]
                x0[k] /= x0[k].sum()
            plans = [_plan_of(results_list[i]) for i in chunk]
            if sum(p.n_classes for p in plans) > _BATCH_CLASSES_PER_SAMPLE * len(plans):
                # large samples: an iteration is bound by its gathers, side by side buys nothing
                # (measured: 10.3 against 8.5 ms per sample for two 480 k-class samples)
                runs = [p.run(lengths[k], x0[k]) for k, p in enumerate(plans)]
                xs = numpy.stack([o[0] for o, _ in runs])
                its = numpy.asarray([int(it[0]) for _, it in runs], dtype='i4')
            else:
                xs, its = _lib.EmPlan.run_many(plans, lengths, x0)
            for k, i in enumerate(chunk):
                out[i] = _finish(xs[k])
                iters[i] = its[k]
                planned.add(i)
    live = [i for i in live if i not in planned]
    for start in range(0, len(live), per_call):
        chunk = live[start:start + per_call]
        xs, its = _em_samples_device([results_list[i] for i in chunk], n_tx, device)
        for k, i in enumerate(chunk):
            out[i] = _finish(xs[k])
            iters[i] = its[k]
    return (out, iters) if return_iters else out


def _em_samples_device(samples, n_tx, device=0):
    """One `skm_em_samples` call: the samples' class structures laid end to end.  Returns the
    EM results (samples x transcripts, before the TPM step) and the iteration counts."""
    ptrs, txs, counts, lengths, x0 = [], [], [], [], []
    first_class = numpy.zeros(len(samples) + 1, dtype='i8')
    nnz = 0
    for k, r in enumerate(samples):
        if r.effective_lengths.size != n_tx:
            raise ValueError('quantify_samples: samples must share the transcript set')
        count = numpy.ascontiguousarray(r.class_count, dtype='f8')
        ptr, tx = _csr_from_class_map(r.class_map, count.shape[0])
        ptrs.append(ptr[:-1] + nnz)
        nnz += int(ptr[-1])
        txs.append(tx)
        counts.append(count)
        first_class[k + 1] = first_class[k] + count.shape[0]
        length = r.effective_lengths.astype('f8')
        x = numpy.ones(n_tx, dtype='f8') / length
        x /= x.sum()
        lengths.append(length)
        x0.append(x)
    ptr = numpy.ascontiguousarray(numpy.concatenate(ptrs + [numpy.asarray([nnz], dtype='i8')]), dtype='i8')
    tx = numpy.ascontiguousarray(numpy.concatenate(txs), dtype='i4')
    counts = numpy.ascontiguousarray(numpy.concatenate(counts), dtype='f8')
    lengths = numpy.ascontiguousarray(numpy.stack(lengths), dtype='f8')
    x0 = numpy.ascontiguousarray(numpy.stack(x0), dtype='f8')
    xs = numpy.zeros_like(x0)
    its = numpy.zeros(len(samples), dtype='i4')
    _lib.require_device()
    _lib.check(_lib.load().skm_em_samples(
        _lib._np_ptr(ptr), _lib._np_ptr(tx), _lib._np_ptr(first_class), len(samples), counts.shape[0],
        tx.shape[0], _lib._np_ptr(counts), _lib._np_ptr(lengths), n_tx, _lib._np_ptr(x0), 0,
        _lib._np_ptr(xs), _lib._np_ptr(its), 0, device, None))
    return xs, its


# ---- writers (`infer.py:171-325`) -------------------------------------------------------------
def output_results(output_path, index, start_time, results, main_abundance,
                   bootstrapped_abundance):
    run_info = _generate_run_info(bootstrapped_abundance, index, results, start_time)
    with (output_path / 'run_info.json').open('w') as f:
        json.dump(run_info, f)
    est_counts = _infer_est_counts(index, results, main_abundance)
    _output_abundance_table(output_path, index, results, est_counts, main_abundance)
    _output_hdf5(output_path, index, results, run_info, est_counts, bootstrapped_abundance)


def _generate_run_info(bootstrapped_abundance, index, results, start_time):
    if results.class_map.size:
        class_target_count = numpy.bincount(results.class_map[0],
                                            minlength=results.class_count.size)
        unique_count = results.class_count[class_target_count == 1].sum()
    else:
        unique_count = 0.0
    total = max(results.total, 1)
    return {
        'n_targets': len(index.transcripts),
        'n_bootstraps': len(bootstrapped_abundance),
        'n_processed': results.total,
        'n_pseudoaligned': results.aligned,
        'n_unique': int(unique_count),
        'p_pseudoaligned': results.aligned / total,
        'p_unique': float(unique_count) / total,
        'kallisto_version': '0.44.0',
        'index_version': 9000,
        'start_time': start_time.isoformat(sep=' '),
        'call': ' '.join([shlex.quote(arg) for arg in sys.argv]),
    }


def _infer_est_counts(index, results, main_abundance):
    est_counts = main_abundance * index.transcripts['length']
    with numpy.errstate(all='ignore'):
        est_counts *= results.aligned / est_counts.sum()
    return est_counts


def _output_abundance_table(output_path, index, results, est_counts, main_abundance):
    """Kallisto-style TSV; `%g` floats like pandas' `float_format='%g'` (`infer.py:219-230`)."""
    ids = index.transcripts['transcript_id']
    lengths = index.transcripts['length']
    eff = results.effective_lengths.astype('f4')
    # pandas applies float_format to float columns only: `length` is f8 in an index written by
    # `seekmer index` (`index_builder.py:225-226`) and goes through '%g' like the rest; an index
    # that carries integer lengths gets them printed in full
    def column(values):
        values = numpy.asarray(values)
        if values.dtype.kind != 'f':
            return [str(v) for v in values.tolist()]
        return ['' if v != v else '%g' % v for v in values.tolist()]  # to_csv writes NaN as an empty field

    columns = [[i.decode() for i in ids], column(lengths), column(eff), column(est_counts), column(main_abundance)]
    with (output_path / 'abundance.tsv').open('w') as f:
        f.write('target_id\tlength\teff_length\test_count\ttpm\n')
        f.writelines('\t'.join(row) + '\n' for row in zip(*columns))


def _output_hdf5(output_path, index, results, run_info, est_counts, bootstrapped_abundance):
    """`abundance.h5` (`infer.py:255-325`); needs PyTables.  Where HDF5 is unavailable the same
    arrays are written to `abundance.npz` and a warning is logged."""
    arrays = {
        'aux/call': numpy.frombuffer(run_info['call'].encode(), dtype='S1'),
        'aux/index_version': numpy.asarray([run_info['index_version']]),
        'aux/start_time': numpy.frombuffer(run_info['start_time'].encode(), dtype='S1'),
        'aux/num_bootstrap': numpy.asarray([run_info['n_bootstraps']]),
        'aux/num_processed': numpy.asarray([run_info['n_processed']]),
        'aux/kallisto_version': numpy.frombuffer(run_info['kallisto_version'].encode(), dtype='S1'),
        'aux/ids': index.transcripts['transcript_id'],
        'aux/lengths': index.transcripts['length'],
        'aux/fld': results.fragment_length_frequencies.astype('i4'),
        'aux/eff_lengths': results.effective_lengths.astype('f8'),
        'aux/bias_observed': numpy.ones(4096, dtype='i4'),
        'aux/bias_normalized': numpy.ones(4096, dtype='f8'),
        'est_counts': est_counts.astype('f8'),
    }
    for i, bootstrap in enumerate(bootstrapped_abundance):
        arrays['bootstrap/bs{}'.format(i)] = bootstrap
    try:
        import tables
        if not hasattr(tables, '__version__'):  # a stand-in module without HDF5 behind it
            raise ImportError('not PyTables')
    except ImportError:
        _LOG.warn('PyTables is not installed: writing abundance.npz instead of abundance.h5')
        numpy.savez(str(output_path / 'abundance.npz'),
                    **{k.replace('/', '__'): v for k, v in arrays.items()})
        return
    with tables.open_file(str(output_path / 'abundance.h5'), mode='w',
                          filters=tables.Filters()) as file:
        groups = {}
        for key, value in arrays.items():
            if '/' in key:
                group_name, name = key.split('/')
                if group_name not in groups:
                    groups[group_name] = file.create_group('/', group_name)
                file.create_carray(groups[group_name], name, obj=value)
            else:
                file.create_carray('/', key, obj=value)


def add_subcommand_parser(subparsers):
    """`seekmer infer` arguments (`infer.py:328-353`)."""
    parser = subparsers.add_parser('infer', help='infer transcript abundance')
    parser.add_argument('index_path', type=pathlib.Path, metavar='index',
                        help='specify a Seekmer index file')
    parser.add_argument('output_path', type=pathlib.Path, metavar='output',
                        help='specify a output folder')
    parser.add_argument('fastq_paths', type=pathlib.Path, metavar='fastq', nargs='+',
                        help='specify a FASTQ read file')
    parser.add_argument('-j', '--jobs', type=int, dest='job_count', metavar='N', default=1,
                        help='specify the maximum parallel job number (here: GPUs to use)')
    parser.add_argument('-m', '--save-readmap', action='store_true', dest='save_readmap',
                        help='output an readmap file')
    parser.add_argument('-s', '--single-ended', action='store_true', dest='single_ended',
                        help='specify whether the reads are single-ended')
    parser.add_argument('-b', '--bootstrap', type=int, dest='bootstrap', default=0,
                        help='specify the number of bootstrapped estimation')
