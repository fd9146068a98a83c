"""Adversarial read sets: N's, lowercase, homopolymers, reads ending at
contig junctions, mates on the same strand, one mate random, L = 25, ragged lengths.
Deterministic; shared by the golden generator and the tests."""
import numpy

from seekmer_b200 import synth


def _rc(s):
    return synth.reverse_complement_ascii(s)


def make_reads(tx, paired):
    rng = numpy.random.Generator(numpy.random.PCG64(99))
    seqs = tx.sequences()
    reads = []

    def frag(t, start, length):
        return seqs[t][start:start + length]

    def mutate(s, rate):
        a = bytearray(s)
        for i in range(len(a)):
            if rng.random() < rate:
                a[i] = b'ACGT'[(b'ACGT'.index(a[i]) + int(rng.integers(1, 4))) % 4] if a[i] in b'ACGT' else a[i]
        return bytes(a)

    units = []
    for k in range(160):
        t = int(rng.integers(0, len(seqs)))
        L = int(rng.choice([25, 26, 30, 49, 50, 51, 75, 100, 101, 150, 151, 200]))
        L = min(L, len(seqs[t]))
        flen = min(len(seqs[t]), max(L, int(rng.integers(L, L + 300))))
        start = int(rng.integers(0, len(seqs[t]) - flen + 1))
        m1 = frag(t, start, L)
        m2 = _rc(frag(t, start + flen - L, L))
        kind = k % 16
        if kind == 0:
            m1, m2 = mutate(m1, 0.05), mutate(m2, 0.05)
        elif kind == 1:
            m1 = m1.lower()
        elif kind == 2:
            a = bytearray(m1)
            for i in rng.integers(0, len(a), size=3):
                a[int(i)] = ord('N')
            m1 = bytes(a)
        elif kind == 3:
            m1 = b'A' * L
        elif kind == 4:
            m2 = _rc(m2)           # mates on the same strand
        elif kind == 5:
            m2 = bytes(rng.choice(list(b'ACGT'), size=L).astype('u1'))  # one mate random
        elif kind == 6:
            m1, m2 = m2, m1        # swapped
        elif kind == 7:
            m1 = bytes(rng.choice(list(b'ACGT'), size=L).astype('u1'))
        elif kind == 8:
            a = bytearray(m1)
            a[0] = ord('N'); a[-1] = ord('n')
            m1 = bytes(a)
        elif kind == 9:
            m1 = mutate(m1, 0.15)
        elif kind == 10:
            # chimeric read: halves from two transcripts
            t2 = int(rng.integers(0, len(seqs)))
            h = L // 2
            m1 = m1[:h] + seqs[t2][:L - h]
        elif kind == 11:
            m1 = b'ACGT' * (L // 4) + b'A' * (L % 4)
        elif kind == 12:
            a = bytearray(m1)
            a[len(a) // 2] = ord('R')   # IUPAC
            m1 = bytes(a)
        elif kind == 13:
            m1 = mutate(m1[:25], 0.0) + mutate(m1[25:], 0.08)
        elif kind == 14:
            # deletion / insertion in the middle
            h = L // 2
            if L > 30:
                m1 = m1[:h] + m1[h + 1:] + b'C'
        elif kind == 15:
            h = L // 2
            if L > 30:
                m1 = (m1[:h] + b'G' + m1[h:])[:L]
        units.append((m1, m2))
    for m1, m2 in units:
        if paired:
            if len(m2) < 25:
                m2 = m2 + b'A' * (25 - len(m2))
            reads += [m1, m2]
        else:
            reads.append(m1)
    return reads
