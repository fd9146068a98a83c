"""Single-cell imputation workflow — the `seekmer.impute` surface (`impute.py:19-252`),
SURVEY.md §8(f)3.

Mirrors: add_subcommand_parser, run, _merge_fragment_lengths, _calculate_cell_weights,
_blend_mapping_results — same arguments, same files in the output folder
(initial_gene_table.csv, weight.csv, tpm.csv).

What changes underneath:
  * every cell is mapped by the GPU mapper (`mapper.map_multiple_samples`);
  * the second-round quantification is ONE batched EM.  After blending, all cells share one
    class structure (the concatenation of every cell's classes, `impute.py:238-246`) and
    differ only in their class counts, which is exactly the replicate layout `skm_em` runs
    (`[n_classes][n_cells]`, replicate fastest).  The reference loops `infer.quantify` over
    the cells, N EMs over N-times-larger inputs (`impute.py:110-115`);
  * the two-cluster split of the correlation values is solved exactly (sorted prefix sums)
    instead of by sklearn's randomly initialised Lloyd iteration (`impute.py:213-214`, no
    random_state): deterministic, and equal to what the reference converges to whenever its
    iteration is not trapped in a worse local optimum.
`_calculate_uniquely_mapped_counts` (`impute.py:144-179`) has no caller in the reference and
is not reproduced.  There is no CPU EM fallback.
"""
import pathlib

import numpy

from . import common
from . import infer
from . import mapper
from ._log import Logger

__all__ = ('add_subcommand_parser', 'run', 'impute_cells')

_LOG = Logger(__name__)

# class-count matrix handed to one skm_em call (fp64, cells x blended classes)
_EM_BATCH_BYTES = 2 << 30


def add_subcommand_parser(subparsers):
    """Add the impute command (`impute.py:19-51`)."""
    parser = subparsers.add_parser(
        'impute', help='impute transcript abundance for single-cell data',
        epilog='Demultiplex the reads first. Every two files are one cell; with "-s" every '
               'single file is one cell.')
    parser.add_argument('index_path', type=pathlib.Path, metavar='index',
                        help='specify a Seekmer index file')
    parser.add_argument('output_path', type=pathlib.Path, metavar='output',
                        help='specify a output folder')
    parser.add_argument('fastq_paths', type=pathlib.Path, metavar='fastq', nargs='+',
                        help='specify a FASTQ read file')
    parser.add_argument('-j', '--jobs', type=int, dest='job_count', metavar='N', default=1,
                        help='specify the maximum parallel job number')
    parser.add_argument('-p', '--power', type=int, dest='power', metavar='P', default=16,
                        help='specify the power of the weight matrix')
    parser.add_argument('-m', '--save-readmap', action='store_true', dest='save_readmap',
                        help='output an readmap file')
    parser.add_argument('-s', '--single-ended', action='store_true', dest='single_ended',
                        help='specify whether the reads are single-ended')


def run(index_path, output_path, fastq_paths, job_count, single_ended, debug, power, **__):
    """The entrypoint of the imputation module (`impute.py:54-127`)."""
    import pandas
    for path in fastq_paths:
        if not pathlib.Path(path).exists():
            raise ValueError(f'invalid FastQ file: {path}')
    try:
        output_path.mkdir(parents=True)
    except FileExistsError:
        _LOG.warn('The output folder exists. Overriding...')
    _LOG.info('Inferring transcript abundance')
    index = common.KMerIndex.load(index_path)
    _LOG.info('Mapping all reads')
    if single_ended:
        cell_paths = list(fastq_paths)
        feeders = [common.feed_single_ended_reads(path) for path in cell_paths]
    else:
        groups = list(common.iterate_by_group(fastq_paths, 2))
        cell_paths = list(fastq_paths[::2])
        feeders = [common.feed_pair_ended_reads(*paths) for paths in groups]
    map_results = mapper.map_multiple_samples(index, feeders, job_count=job_count, debug=debug)
    _LOG.info('Mapped all reads.')
    tpm = impute_cells(index, map_results, power=power, output_path=output_path)
    ids = numpy.char.decode(index.transcripts['transcript_id'])
    table = pandas.DataFrame({str(path): row for path, row in zip(cell_paths, tpm)}, index=ids)
    _LOG.info('Writing results to {}...', output_path)
    table.to_csv(output_path / 'tpm.csv')


def impute_cells(index, map_results, power=16, output_path=None, return_stages=False):
    """Everything of `impute.run` between mapping and the final table (`impute.py:99-122`):
    merge the fragment lengths, quantify every cell, weight the cells by the correlation of
    their gene tables, blend the class counts and quantify again.  Returns the cell-by-
    transcript TPM matrix (with `return_stages`: also the first-round matrix and the filtered
    weights before `power`)."""
    _merge_fragment_lengths(map_results)
    summarized = [r.summarize() for r in map_results]
    _LOG.info('First round quantification...')
    base = numpy.asarray([infer.quantify(r) for r in summarized])
    if power is None:
        return (base, base, None) if return_stages else base
    _LOG.info('Weighting cells.')
    weight = _calculate_cell_weights(index, base, output_path)
    blended = weight ** power
    _blend_mapping_results(summarized, blended)
    _LOG.info('Second round quantification...')
    tpm = _quantify_blended(summarized)
    return (tpm, base, weight) if return_stages else tpm


def _merge_fragment_lengths(map_results):
    """One fragment length distribution for all cells (`impute.py:130-142`): the cells of a
    run come from one sequencing batch.  Every result ends up holding the same array."""
    total = numpy.zeros(mapper.MAX_FRAGMENT_LENGTH, dtype='i8')
    for result in map_results:
        total += result.fragment_length_counts
    for result in map_results:
        result.fragment_length_counts = total


def _gene_matrix(index, base_matrix):
    """Cell-by-gene table of `impute.py:198-204`: TPM summed over the transcripts of a gene,
    stored as int64 (the reference assigns the float sums into an 'i8' array, i.e. truncates),
    genes in sorted order, the empty gene id dropped."""
    genes, gene_of = numpy.unique(index.transcripts['gene_id'], return_inverse=True)
    gene_of = numpy.asarray(gene_of).reshape(-1)
    base_matrix = numpy.asarray(base_matrix, dtype='f8')
    sums = numpy.zeros((base_matrix.shape[0], len(genes)), dtype='f8')
    for g in range(len(genes)):  # the reference's own summation: numpy pairwise over the mask
        sums[:, g] = base_matrix[:, gene_of == g].sum(axis=1)
    named = genes != b''
    return sums.astype('i8')[:, named], genes[named]


def _two_means(values):
    """Optimal 2-means of 1-D data: (low centre, high centre).  Every 2-clustering that is
    optimal for the within-cluster sum of squares is a split of the sorted values."""
    v = numpy.sort(numpy.asarray(values, dtype='f8'))
    n = v.size
    if n < 2:
        raise ValueError(f'n_samples={n} should be >= n_clusters=2.')
    prefix = numpy.cumsum(v)
    prefix_sq = numpy.cumsum(v * v)
    k = numpy.arange(1, n)                     # size of the low cluster
    low_sum, low_sq = prefix[:-1], prefix_sq[:-1]
    high_sum, high_sq = prefix[-1] - low_sum, prefix_sq[-1] - low_sq
    cost = (low_sq - low_sum * low_sum / k) + (high_sq - high_sum * high_sum / (n - k))
    best = int(numpy.argmin(cost))
    return low_sum[best] / k[best], high_sum[best] / (n - k[best])


def _calculate_cell_weights(index, base_matrix, output_path):
    """Cell-by-cell weights (`impute.py:182-224`): Pearson correlation of the integer gene
    tables; the correlations other than NaN and exactly 1.0 are split into two clusters and only
    pairs falling nearer the higher centre keep their weight."""
    gene_matrix, names = _gene_matrix(index, base_matrix)
    if output_path is not None:
        import pandas
        pandas.DataFrame(gene_matrix.T, index=numpy.char.decode(names)).to_csv(
            pathlib.Path(output_path) / 'initial_gene_table.csv')
    with numpy.errstate(all='ignore'):
        weights = numpy.atleast_2d(numpy.corrcoef(gene_matrix))
    valid = ~numpy.isnan(weights)
    low, high = _two_means(weights[valid & (weights != 1.0)])
    weights[~valid] = 0.0
    keep = numpy.abs(weights - high) < numpy.abs(weights - low)
    weights = numpy.where(keep, weights, 0.0)
    if output_path is not None:
        import pandas
        pandas.DataFrame(weights).to_csv(pathlib.Path(output_path) / 'weight.csv')
    return weights


def _blend_mapping_results(map_results, weight):
    """Blend the class counts across cells (`impute.py:227-252`).  Afterwards every result
    holds the same `class_map` — all cells' classes, renumbered consecutively — and, for cell i,
    the counts `c_j * weight[i, j] * total_i / sum(c_j)` over the cells j."""
    maps, counts = [], []
    first = 0
    for result in map_results:
        result.class_map[0, :] += first
        first = result.class_map[0, :].max() + 1
        maps.append(result.class_map)
        counts.append(result.class_count)
    shared = numpy.concatenate(maps, axis=1)
    sums = [c.sum() for c in counts]
    for i, result in enumerate(map_results):
        total = result.class_count.sum()
        with numpy.errstate(all='ignore'):
            result.class_count = numpy.concatenate(
                [c * w * total / s for c, w, s in zip(counts, weight[i, :], sums)])
        result.class_map = shared


def _quantify_blended(map_results):
    """`[infer.quantify(r) for r in map_results]` (`impute.py:110-115`) for results that share
    one class_map: one batched device EM (cells are the replicates), then the TPM step."""
    if not map_results:
        return numpy.zeros((0, 0), dtype='f8')
    shared = map_results[0].class_map
    n_tx = map_results[0].effective_lengths.size
    if shared.size == 0:
        return numpy.zeros((len(map_results), n_tx), dtype='f8')
    out = numpy.zeros((len(map_results), n_tx), dtype='f8')
    n_classes = map_results[0].class_count.size
    per_call = max(1, _EM_BATCH_BYTES // max(8 * n_classes, 1))
    start = 0
    while start < len(map_results):
        cells = map_results[start:start + per_call]
        same_lengths = all(numpy.array_equal(c.effective_lengths, cells[0].effective_lengths)
                           for c in cells)
        if not same_lengths or any(c.class_map is not shared for c in cells):
            raise ValueError('_quantify_blended needs results blended by _blend_mapping_results')
        lengths = cells[0].effective_lengths.astype('f8')
        x0 = numpy.ones(n_tx, dtype='f8') / lengths
        x0 /= x0.sum()
        counts = numpy.stack([numpy.asarray(c.class_count, dtype='f8') for c in cells])
        x, _ = infer._em_device(numpy.tile(x0, (len(cells), 1)), lengths, shared, counts)
        for row in x:
            infer._finish(row)
        out[start:start + len(cells)] = x
        start += len(cells)
    return out
