"""Deterministic synthetic transcriptomes and reads.

No network, no ENSEMBL: benchmarks and parity tests run on random transcriptomes with
shared-exon isoform families (to create multi-mapping) and simulated reads with substitution
errors.  Everything that has a device twin (``csrc/synth.cu``) is integer-only and driven by
the counter-based Philox4x32-10 generator keyed by the *global* read index, so any slice of
any shard can be regenerated bit-identically on the host (for the CPU oracle) or on any GPU.

This module is host-side numpy; it is workload generation, not the product hot path.
"""
import numpy

__all__ = ['make_transcriptome', 'make_expression', 'ReadSimulator', 'philox4x32',
           'codes_to_ascii', 'ascii_to_codes', 'reverse_complement_ascii']

_ASCII = numpy.frombuffer(b'ACGT', dtype='u1')
_COMP = bytes.maketrans(b'ACGTacgt', b'TGCAtgca')


def codes_to_ascii(codes):
    return _ASCII[numpy.asarray(codes, dtype='u1')]


def ascii_to_codes(ascii_bases):
    lut = numpy.zeros(256, dtype='u1')
    for i, c in enumerate(b'ACGT'):
        lut[c] = i
        lut[c + 32] = i
    return lut[numpy.asarray(ascii_bases, dtype='u1')]


def reverse_complement_ascii(seq):
    return bytes(seq).translate(_COMP)[::-1]


# ------------------------------------------------------------------ Philox4x32-10
_M0 = numpy.uint64(0xD2511F53)
_M1 = numpy.uint64(0xCD9E8D57)
_W0 = 0x9E3779B9
_W1 = 0xBB67AE85
_MASK = numpy.uint64(0xFFFFFFFF)
_S32 = numpy.uint64(32)


def philox4x32(c0, c1, c2, c3, k0, k1):
    """Philox4x32-10 (Salmon et al. 2011). Counter words broadcastable; returns 4 uint64 arrays
    holding 32-bit values."""
    c0 = numpy.asarray(c0, dtype='u8') & _MASK
    c1 = numpy.asarray(c1, dtype='u8') & _MASK
    c2 = numpy.asarray(c2, dtype='u8') & _MASK
    c3 = numpy.asarray(c3, dtype='u8') & _MASK
    c0, c1, c2, c3 = numpy.broadcast_arrays(c0, c1, c2, c3)
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = _M0 * c0
        p1 = _M1 * c2
        c0, c1, c2, c3 = ((p1 >> _S32) ^ c1 ^ numpy.uint64(k0), p1 & _MASK,
                          (p0 >> _S32) ^ c3 ^ numpy.uint64(k1), p0 & _MASK)
        k0 = (k0 + _W0) & 0xFFFFFFFF
        k1 = (k1 + _W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def _mulhi64(a, b):
    """(a * b) >> 64 for uint64 arrays a and python int / uint64 b."""
    a = numpy.asarray(a, dtype='u8')
    b = numpy.asarray(b, dtype='u8')
    a_lo, a_hi = a & _MASK, a >> _S32
    b_lo, b_hi = b & _MASK, b >> _S32
    ll = a_lo * b_lo
    lh = a_lo * b_hi
    hl = a_hi * b_lo
    hh = a_hi * b_hi
    mid = (ll >> _S32) + (lh & _MASK) + (hl & _MASK)
    return hh + (lh >> _S32) + (hl >> _S32) + (mid >> _S32)


# ------------------------------------------------------------------ transcriptome
class Transcriptome:
    """codes: uint8 (0..3) of all transcripts concatenated; offsets int64[T+1]."""

    def __init__(self, codes, offsets, gene_of, strand_flipped):
        self.codes = codes
        self.offsets = offsets
        self.gene_of = gene_of
        self.strand_flipped = strand_flipped

    @property
    def n_transcripts(self):
        return self.offsets.shape[0] - 1

    @property
    def lengths(self):
        return self.offsets[1:] - self.offsets[:-1]

    def ids(self):
        return [b'TX%07d' % i for i in range(self.n_transcripts)]

    def gene_ids(self):
        return [b'GN%07d' % g for g in self.gene_of]

    def sequence(self, t):
        return codes_to_ascii(self.codes[self.offsets[t]:self.offsets[t + 1]]).tobytes()

    def sequences(self):
        asc = codes_to_ascii(self.codes).tobytes()
        o = self.offsets
        return [asc[o[i]:o[i + 1]] for i in range(self.n_transcripts)]


def make_transcriptome(n_transcripts, seed=1, mean_exons=10, median_exon=150, min_exon=30,
                       max_isoforms=12, min_length=400):
    """Random isoform families.

    Each gene family draws E exons (uniform random ACGT, lognormal lengths) and 1..max_isoforms
    isoforms as ordered exon subsets, so isoforms of a family share exons (multi-mapping) and
    exon junctions differ (contig boundaries).  Half of the genes are emitted reverse-
    complemented so that both signs of `entry` occur in the index.
    """
    rng = numpy.random.Generator(numpy.random.PCG64(seed))
    # --- family structure
    n_iso_all = []
    total = 0
    while total < n_transcripts:
        k = int(rng.integers(1, max_isoforms + 1))
        k = min(k, n_transcripts - total)
        n_iso_all.append(k)
        total += k
    n_genes = len(n_iso_all)
    n_iso = numpy.asarray(n_iso_all, dtype='i8')
    n_exons = rng.integers(max(3, mean_exons - 6), mean_exons + 7, size=n_genes)
    exon_gene = numpy.repeat(numpy.arange(n_genes), n_exons)
    n_exon_total = int(n_exons.sum())
    exon_len = numpy.maximum(
        min_exon, rng.lognormal(numpy.log(median_exon), 0.6, size=n_exon_total).astype('i8'))
    gene_exon_start = numpy.zeros(n_genes + 1, dtype='i8')
    numpy.cumsum(n_exons, out=gene_exon_start[1:])
    # every gene is at least min_length long with all exons kept (first exon absorbs the deficit)
    gene_len = numpy.bincount(exon_gene, weights=exon_len, minlength=n_genes).astype('i8')
    exon_len[gene_exon_start[:-1]] += numpy.maximum(min_length - gene_len, 0)
    exon_off = numpy.zeros(n_exon_total + 1, dtype='i8')
    numpy.cumsum(exon_len, out=exon_off[1:])
    pool = rng.integers(0, 4, size=int(exon_off[-1]), dtype='u1')
    flipped_gene = rng.random(n_genes) < 0.5
    # --- isoforms: keep mask per (transcript, exon of its gene)
    tx_gene = numpy.repeat(numpy.arange(n_genes), n_iso)
    tx_nex = n_exons[tx_gene]
    seg_tx = numpy.repeat(numpy.arange(n_transcripts), tx_nex)
    seg_first = numpy.zeros(n_transcripts + 1, dtype='i8')
    numpy.cumsum(tx_nex, out=seg_first[1:])
    seg_local = numpy.arange(seg_tx.shape[0]) - seg_first[seg_tx]
    seg_exon = gene_exon_start[tx_gene[seg_tx]] + seg_local
    keep = rng.random(seg_tx.shape[0]) < 0.7
    # first isoform of every gene keeps all exons; too-short isoforms keep all exons too
    first_tx = numpy.zeros(n_transcripts, dtype=bool)
    first_tx[numpy.cumsum(n_iso) - n_iso] = True
    keep |= first_tx[seg_tx]
    tlen = numpy.bincount(seg_tx, weights=exon_len[seg_exon] * keep, minlength=n_transcripts)
    keep |= (tlen < min_length)[seg_tx]
    seg_tx, seg_exon = seg_tx[keep], seg_exon[keep]
    seg_len = exon_len[seg_exon]
    tlen = numpy.bincount(seg_tx, weights=seg_len, minlength=n_transcripts).astype('i8')
    offsets = numpy.zeros(n_transcripts + 1, dtype='i8')
    numpy.cumsum(tlen, out=offsets[1:])
    # --- gather bases: position p in output -> pool index
    seg_out = numpy.zeros(seg_len.shape[0] + 1, dtype='i8')
    numpy.cumsum(seg_len, out=seg_out[1:])
    n_total = int(seg_out[-1])
    src = numpy.repeat(exon_off[seg_exon] - seg_out[:-1], seg_len) + numpy.arange(n_total)
    codes = pool[src]
    del src
    # --- reverse-complement flipped genes' transcripts in place
    flipped_tx = flipped_gene[tx_gene]
    if flipped_tx.any():
        pos = numpy.arange(n_total, dtype='i8')
        tx_of_pos = numpy.repeat(numpy.arange(n_transcripts), tlen)
        fl = flipped_tx[tx_of_pos]
        mirror = offsets[tx_of_pos] + offsets[tx_of_pos + 1] - 1 - pos
        src2 = numpy.where(fl, mirror, pos)
        codes = numpy.where(fl, 3 - codes[src2], codes[src2]).astype('u1')
    return Transcriptome(codes, offsets, tx_gene, flipped_tx)


def make_expression(n_transcripts, seed=3):
    """lognormal(0, 2) with 30 % zeros."""
    rng = numpy.random.Generator(numpy.random.PCG64(seed))
    e = rng.lognormal(0.0, 2.0, size=n_transcripts)
    e[rng.random(n_transcripts) < 0.3] = 0.0
    return e


# ------------------------------------------------------------------ reads
class ReadSimulator:
    """Integer-only paired/single read simulator with a bit-identical CUDA twin.

    Pair i (global index) is a pure function of (seed, i):
      Philox(ctr=(i_lo, i_hi, 0, 0)) -> x0..x3:
        transcript: u = ((x1<<32|x0) * W) >> 64 over integer cumulative weights (upper bound)
        x2: bit0 = swap mates, bits 8.. : (x2 >> 8) % 100 == 0 -> pure-random pair
        x3: uniform start
      Philox(ctr=(i_lo, i_hi, 1, 0)) -> eight 16-bit uniforms summed (Irwin-Hall ~ normal)
        fragment = mu + (S*sd)//53510 - (262140*sd)//53510, clipped to [L, transcript length]
      Philox(ctr=(i_lo, i_hi, 2 + j//8, 0)), j = base index within the pair (mate1 then mate2):
        16 bits v per base: v < sub_thresh -> substitute by (b + 1 + v % 3) & 3;
        v >= 65536 - n_thresh -> 'N'.
      Pure-random pairs take their bases from the low 2 bits of the same 16-bit lanes
      (>> 2 so the error bits stay independent).
    mate 1 = first L bases of the fragment, mate 2 = reverse complement of the last L bases.
    """

    IRWIN_MEAN = 262140
    IRWIN_SD = 53510

    def __init__(self, transcriptome, expression, read_length, frag_mean, frag_sd,
                 sub_rate=0.01, n_rate=0.001, random_rate_pct=1, seed=10, paired=True):
        self.tx = transcriptome
        self.L = int(read_length)
        self.mu = int(frag_mean)
        self.sd = int(frag_sd)
        self.sub_thresh = int(round(sub_rate * 65536))
        self.n_thresh = int(round(n_rate * 65536))
        self.random_pct = int(random_rate_pct)
        self.seed = int(seed)
        self.paired = bool(paired)
        lengths = transcriptome.lengths
        if int(lengths.min()) < self.L:
            raise ValueError('transcripts shorter than the read length')
        w = numpy.asarray(expression, dtype='f8') * numpy.maximum(lengths - self.mu + 1, 1)
        w = w / w.sum()
        wi = numpy.floor(w * float(1 << 40)).astype('u8')
        self.cum_weights = numpy.cumsum(wi).astype('u8')
        self.total_weight = int(self.cum_weights[-1])

    def _fragments(self, idx):
        lo, hi = idx & _MASK, idx >> _S32
        x0, x1, x2, x3 = philox4x32(lo, hi, 0, 0, self.seed & 0xFFFFFFFF, self.seed >> 32)
        u = _mulhi64((x1 << _S32) | x0, self.total_weight)
        t = numpy.searchsorted(self.cum_weights, u, side='right').astype('i8')
        swap = (x2 & numpy.uint64(1)).astype(bool)
        is_random = ((x2 >> numpy.uint64(8)) % numpy.uint64(100)) < numpy.uint64(self.random_pct)
        y0, y1, y2, y3 = philox4x32(lo, hi, 1, 0, self.seed & 0xFFFFFFFF, self.seed >> 32)
        s = numpy.zeros(idx.shape, dtype='u8')
        for y in (y0, y1, y2, y3):
            s += (y & numpy.uint64(0xFFFF)) + (y >> numpy.uint64(16))
        frag = (self.mu + (s * numpy.uint64(self.sd)) // numpy.uint64(self.IRWIN_SD)).astype('i8') \
            - (self.IRWIN_MEAN * self.sd) // self.IRWIN_SD
        tlen = self.tx.lengths[t]
        fmin = self.L
        frag = numpy.minimum(numpy.maximum(frag, fmin), tlen)
        if not self.paired:
            frag = numpy.full_like(frag, self.L)
        span = (tlen - frag + 1).astype('u8')
        start = ((x3 * span) >> _S32).astype('i8')
        return t, start, frag, swap, is_random

    def generate(self, first, count):
        """ASCII reads for global units [first, first+count).

        Returns (bases uint8[count * reads_per_unit * L], truth dict). Reads of a pair are
        interleaved (mate1, mate2), fixed length L.
        """
        L = self.L
        idx = numpy.arange(first, first + count, dtype='u8')
        t, start, frag, swap, is_random = self._fragments(idx)
        codes = self.tx.codes
        toff = self.tx.offsets[t]
        ar = numpy.arange(L, dtype='i8')
        m1 = codes[(toff + start)[:, None] + ar[None, :]]
        nm = 2 if self.paired else 1
        if self.paired:
            end = toff + start + frag
            m2 = 3 - codes[(end - 1)[:, None] - ar[None, :]]
            sw = swap[:, None]
            a = numpy.where(sw, m2, m1)
            b = numpy.where(sw, m1, m2)
            pair = numpy.concatenate([a, b], axis=1)
        else:
            # single-end: strand chosen by the swap bit
            m1r = 3 - codes[(toff + start + L - 1)[:, None] - ar[None, :]]
            pair = numpy.where(swap[:, None], m1r, m1)
        nb = nm * L
        ncall = (nb + 7) // 8
        lo, hi = idx & _MASK, idx >> _S32
        calls = numpy.arange(ncall, dtype='u8') + numpy.uint64(2)
        z = philox4x32(lo[:, None], hi[:, None], calls[None, :], 0,
                       self.seed & 0xFFFFFFFF, self.seed >> 32)
        lanes = numpy.empty((count, ncall, 8), dtype='u8')
        for k in range(4):
            lanes[:, :, 2 * k] = z[k] & numpy.uint64(0xFFFF)
            lanes[:, :, 2 * k + 1] = z[k] >> numpy.uint64(16)
        v = lanes.reshape(count, ncall * 8)[:, :nb]
        base = numpy.where(is_random[:, None], (v >> numpy.uint64(2)) & numpy.uint64(3),
                           pair.astype('u8'))
        sub = v < numpy.uint64(self.sub_thresh)
        base = numpy.where(sub, (base + numpy.uint64(1) + v % numpy.uint64(3)) & numpy.uint64(3),
                           base)
        out = _ASCII[base.astype('u1')]
        out[v >= numpy.uint64(65536 - self.n_thresh)] = ord('N')
        truth = {'transcript': t, 'start': start, 'fragment': frag, 'swap': swap,
                 'random': is_random}
        return numpy.ascontiguousarray(out.reshape(-1)), truth

    def offsets(self, count):
        nm = 2 if self.paired else 1
        return numpy.arange(count * nm + 1, dtype='i8') * self.L

    def batches(self, first, count, batch=65536):
        """Feeder-shaped batches `(read_count, names, reads)` (`common.py:126-197`)."""
        nm = 2 if self.paired else 1
        done = 0
        while done < count:
            n = min(batch, count - done)
            bases, _ = self.generate(first + done, n)
            raw = bases.tobytes()
            L = self.L
            reads = [raw[i * L:(i + 1) * L] for i in range(n * nm)]
            names = [b'r%d' % (first + done + i) for i in range(n)]
            yield n, names, reads
            done += n
