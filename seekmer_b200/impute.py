"""Single-cell imputation workflow — the `seekmer.impute` surface (`impute.py:19-252`),
SURVEY.md §8(f)3.

Mirrors: add_subcommand_parser, run, _merge_fragment_lengths, _calculate_cell_weights,
_blend_mapping_results — same arguments, same files in the output folder
(initial_gene_table.csv, weight.csv, tpm.csv).

What changes underneath:
  * every cell is mapped by the GPU mapper (`mapper.map_multiple_samples`);
  * the second-round quantification is a batched EM.  After blending, all cells share one
    class structure (the concatenation of every cell's classes, `impute.py:238-246`) and
    differ only in their class counts, which is exactly the replicate layout `skm_em` runs
    (`[n_classes][n_cells]`, replicate fastest).  The reference loops `infer.quantify` over
    the cells, N EMs over N-times-larger inputs (`impute.py:110-115`).  Cells drawing on the
    same set of cells are batched into one call over just those cells' classes
    (`_quantify_weighted`); zero-count classes add exactly 0 to the EM sums;
  * the two-cluster split of the correlation values is, by default, the reference's own
    scikit-learn call (`impute.py:213-214`: `KMeans(2)`, unseeded, stopped by tolerance) so
    that a seeded run reproduces the reference's weights entry for entry; `CLUSTERING =
    'exact'` switches to the optimal split by sorted prefix sums (deterministic; differs from
    the iterative result only for values within ~0.01 of the threshold).
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

# how the cell-cell correlations are split in two (`_high_cluster`): 'reference' or 'exact'
CLUSTERING = 'reference'


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


def impute_cells(index, map_results, power=16, output_path=None, return_stages=False,
                 clustering=None):
    """Everything of `impute.run` between mapping and the final table (`impute.py:99-122`):
    merge the fragment lengths, quantify every cell, weight the cells by the correlation of
    their gene tables, blend the class counts and quantify again.  Returns the cell-by-
    transcript TPM matrix (with `return_stages`: also the first-round matrix and the filtered
    weights before `power`)."""
    _merge_fragment_lengths(map_results)
    summarized = [r.summarize() for r in map_results]
    _LOG.info('First round quantification...')
    base = infer.quantify_samples(summarized)
    if power is None:
        return (base, base, None) if return_stages else base
    _LOG.info('Weighting cells.')
    weight = _calculate_cell_weights(index, base, output_path, clustering)
    _LOG.info('Second round quantification...')
    tpm = _quantify_weighted(summarized, weight ** power)
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
    # The reference sums `base_matrix[:, gene_of == g]` gene by gene (genes x transcripts work).
    # Same sums, same rounding: genes with equally many transcripts are gathered into one
    # (cells, genes, k) block and reduced over the contiguous last axis, which is the
    # summation numpy applies to each gene's (cells, k) block.
    order = numpy.argsort(gene_of, kind='stable')          # transcripts of a gene, ascending
    sizes = numpy.bincount(gene_of, minlength=len(genes))
    first = numpy.concatenate([[0], numpy.cumsum(sizes)[:-1]])
    for k in numpy.unique(sizes):
        which = numpy.flatnonzero(sizes == k)
        members = order[first[which][:, None] + numpy.arange(k)[None, :]]
        sums[:, which] = base_matrix[:, members].sum(axis=2)
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


def _high_cluster(values, matrix, clustering):
    """Which entries of `matrix` fall into the higher of the two clusters of `values`.

    'reference': the reference's own procedure (`impute.py:213-218`) — scikit-learn's
    `KMeans(2)` with its defaults, initialised from numpy's global RNG and stopped by its
    tolerance, labels from `predict`.  Seeding numpy reproduces the reference's weights entry
    for entry; unseeded, boundary values can land on either side from run to run, exactly as
    they do in the reference.
    'exact': the optimal split (`_two_means`), deterministic; never a worse clustering than
    the iterative one and the same labels except for values within ~0.01 of the threshold."""
    if clustering == 'exact':
        low, high = _two_means(values)
        return numpy.abs(matrix - high) < numpy.abs(matrix - low)
    if clustering != 'reference':
        raise ValueError(f'unknown clustering: {clustering!r}')
    import sklearn.cluster
    model = sklearn.cluster.KMeans(2)
    model.fit(numpy.asarray(values, dtype='f8').reshape(-1, 1))
    labels = model.predict(matrix.reshape(-1, 1)).reshape(matrix.shape)
    return labels == int(numpy.argmax(model.cluster_centers_.ravel()))


def _calculate_cell_weights(index, base_matrix, output_path, clustering=None):
    """Cell-by-cell weights (`impute.py:182-224`): Pearson correlation of the integer gene
    tables; the correlations other than NaN and exactly 1.0 are split into two clusters and only
    pairs falling into the higher one keep their weight.  `clustering`: see `_high_cluster`
    (default: module constant `CLUSTERING`)."""
    gene_matrix, names = _gene_matrix(index, base_matrix)
    if output_path is not None:
        import pandas
        pandas.DataFrame(gene_matrix.T, index=numpy.char.decode(names)).to_csv(
            pathlib.Path(output_path) / 'initial_gene_table.csv')
    with numpy.errstate(all='ignore'):
        weights = numpy.atleast_2d(numpy.corrcoef(gene_matrix))
    valid = ~numpy.isnan(weights)
    values = weights[valid & (weights != 1.0)]
    weights[~valid] = 0.0
    keep = _high_cluster(values, weights, CLUSTERING if clustering is None else clustering)
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


def _support_groups(weight):
    """Cells grouped by which cells they draw counts from (the non-zero pattern of their weight
    row): cells of one group have zero counts on the same blocks of the blended classes."""
    groups = {}
    for i, row in enumerate(numpy.asarray(weight)):
        groups.setdefault((row != 0).tobytes(), []).append(i)
    return list(groups.values())


def _blended_group(map_results, weight, group):
    """What `_blend_mapping_results` + `_prune_classes` give for the cells of one support group,
    built directly: (class_map over the supporting cells' classes, counts[len(group)][classes]).
    The blocks a group gives no weight to are never materialised, so memory is
    cells-in-group x supported classes instead of cells x all classes per cell."""
    support = numpy.flatnonzero(numpy.asarray(weight)[group[0]] != 0)
    maps, first = [], 0
    for j in support:
        block = numpy.array(map_results[j].class_map, dtype='i8')
        block[0] += first
        first = block[0].max() + 1
        maps.append(block)
    class_map = numpy.concatenate(maps, axis=1)
    sums = [map_results[j].class_count.sum() for j in support]
    counts = numpy.empty((len(group), int(first)), dtype='f8')
    for row, i in enumerate(group):
        total = map_results[i].class_count.sum()
        with numpy.errstate(all='ignore'):  # same operation order as `impute.py:249-251`
            counts[row] = numpy.concatenate(
                [map_results[j].class_count * weight[i, j] * total / s for j, s in zip(support, sums)])
    return class_map, counts


def _quantify_weighted(map_results, weight):
    """Blend + second-round quantification of `impute.py:108-115` without building the
    cells x (all cells' classes) count matrix: one batched device EM per support group."""
    n_tx = map_results[0].effective_lengths.size if map_results else 0
    out = numpy.zeros((len(map_results), n_tx), dtype='f8')
    if not map_results:
        return out
    lengths = map_results[0].effective_lengths.astype('f8')
    x0 = numpy.ones(n_tx, dtype='f8') / lengths
    x0 /= x0.sum()
    for group in _support_groups(weight):
        if not (numpy.asarray(weight)[group[0]] != 0).any():
            # no gene table to correlate (e.g. nothing mapped): the reference fails on such a
            # cell (`impute.py:240`, indexing its empty class_map); here it is reported as zeros
            _LOG.warn('{} cell(s) without any weighted cell: abundances left at zero', len(group))
            continue
        class_map, counts = _blended_group(map_results, weight, group)
        per_call = max(1, _EM_BATCH_BYTES // max(8 * counts.shape[1], 1))
        for start in range(0, len(group), per_call):
            rows = slice(start, start + per_call)
            x, _ = infer._em_device(numpy.tile(x0, (len(group[rows]), 1)), lengths, class_map,
                                    counts[rows])
            for row in x:
                infer._finish(row)
            out[group[rows]] = x
    return out


def _prune_classes(class_map, counts):
    """Drop the classes whose count is zero in every row of `counts` (cells x classes) and
    renumber the rest.  A zero-count class adds exactly 0 to every sum of the EM update
    (`infer.py:153-157`: its `class_inner` is +inf, or NaN where all its transcripts are already
    0 and stay 0), so the fixed point and the iterates are the same without it."""
    active = (counts != 0).any(axis=0)
    if active.all() or not active.any():
        return class_map, counts
    new_id = numpy.cumsum(active) - 1
    rows = numpy.asarray(class_map[0], dtype='i8')
    keep = active[rows]
    pruned = numpy.stack([new_id[rows[keep]], numpy.asarray(class_map[1], dtype='i8')[keep]])
    return pruned, numpy.ascontiguousarray(counts[:, active])


def _quantify_blended(map_results, groups=None):
    """`[infer.quantify(r) for r in map_results]` (`impute.py:110-115`) for results that share
    one class_map: batched device EMs with the cells as replicates, then the TPM step.
    `groups` (lists of cell indices, default: one group) are quantified separately, each over
    only the classes that have a count in at least one of its cells."""
    if not map_results:
        return numpy.zeros((0, 0), dtype='f8')
    shared = map_results[0].class_map
    n_tx = map_results[0].effective_lengths.size
    out = numpy.zeros((len(map_results), n_tx), dtype='f8')
    if shared.size == 0:
        return out
    lengths = map_results[0].effective_lengths.astype('f8')
    for c in map_results:
        if c.class_map is not shared or not numpy.array_equal(c.effective_lengths,
                                                              map_results[0].effective_lengths):
            raise ValueError('_quantify_blended needs results blended by _blend_mapping_results')
    x0 = numpy.ones(n_tx, dtype='f8') / lengths
    x0 /= x0.sum()
    n_classes = map_results[0].class_count.size
    per_call = max(1, _EM_BATCH_BYTES // max(8 * n_classes, 1))
    if groups is None:
        groups = [list(range(len(map_results)))]
    for group in groups:
        for start in range(0, len(group), per_call):
            cells = group[start:start + per_call]
            counts = numpy.stack([numpy.asarray(map_results[i].class_count, dtype='f8')
                                  for i in cells])
            class_map, counts = _prune_classes(shared, counts)
            x, _ = infer._em_device(numpy.tile(x0, (len(cells), 1)), lengths, class_map, counts)
            for row in x:
                infer._finish(row)
            out[cells] = x
    return out
