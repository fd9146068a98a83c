"""Host logic of the FASTQ text path (no GPU): chunk carry-over and end-of-stream handling."""
import io

import numpy

from seekmer_b200 import common


def lines_of(c):
    return bytes(c.buf[:c.n]).split(b'\n')


def test_carry_finishes_streams_like_the_reference_line_logic():
    rec = b'@h\nACGT\n+\nIIII\n'
    for tail, n_reads in ((b'', 2), (b'@x\nAAAA', 3), (b'@x\nAAAA\n', 3), (b'@x\nAAAA\n+', 3), (b'@x\nAAAA\n+\nII', 3),
                          (b'@x', 2), (b'@x\n', 2)):
        c = common._Carry(1 << 16)
        h = io.BytesIO(rec * 2 + tail)
        c.fill(h)
        c.fill(h)
        assert c.eof
        c.finish()
        text = bytes(c.buf[:c.n])
        assert text.count(b'\n') == 4 * n_reads, (tail, text)
        seqs = text.split(b'\n')[1::4]
        assert len(seqs) == n_reads and all(s in (b'ACGT', b'AAAA') for s in seqs)


def test_carry_keeps_the_unconsumed_tail_in_front():
    c = common._Carry(1 << 16)
    data = bytes(range(65, 91)) * 10000
    h = io.BytesIO(data)
    c.fill(h)
    assert c.n == 1 << 16 and bytes(c.buf[:8]) == data[:8]
    c.drop(60000)
    assert c.n == (1 << 16) - 60000 and bytes(c.buf[:c.n]) == data[60000:1 << 16]
    c.fill(h)
    assert bytes(c.buf[:c.n]) == data[60000:60000 + c.n]


def test_sources_iterate_like_the_reference_feeders(tmp_path):
    p = tmp_path / 's.fastq'
    p.write_bytes(b'@a x\nACGT\n+\nIIII\n@b\nGGCC\n+\nIIII\n')
    (count, names, reads), = list(common.feed_single_ended_reads(p))
    assert (count, names, reads) == (2, [b'a x', b'b'], [b'ACGT', b'GGCC'])
    src = common.feed_pair_ended_reads(p, p)
    (count, names, reads), = list(src)
    assert count == 2 and reads == [b'ACGT', b'ACGT', b'GGCC', b'GGCC']
    chunks = list(src.text_chunks(1 << 16))
    assert len(chunks) == 1 and chunks[0][4] is True and chunks[0][1] == p.stat().st_size
