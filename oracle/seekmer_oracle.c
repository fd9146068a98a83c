/*
 * seekmer_oracle.c — CPU restatement of the Seekmer read-mapping hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity oracle for the CUDA path
 * and the "port" CPU baseline of bench.py.  Nothing under seekmer_b200/ links,
 * imports or calls it; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may.
 *
 * Parity status: PINNED.  tests/test_oracle_pinned.py checks this file against
 *   (a) the compiled, unmodified reference (oracle/_ref, built by
 *       oracle/build_ref.py) on the reference's own chr21 fixture and on
 *       simulated / adversarial reads, whenever oracle/_ref is present, and
 *   (b) committed golden vectors generated from that reference
 *       (tests/golden/, script tests/golden/make_golden.py).
 *
 * Every function cites the reference file:line it follows
 * (paths relative to /root/reference/seekmer/).  The restatement keeps the
 * reference's data model on purpose (ASCII reads, malloc'd coordinate lists,
 * 16-byte AoS hash slots, SipHash-variant) so that it is an independent check
 * of the GPU implementation, which uses none of those.
 *
 * Deliberate differences from the reference (none affect defined behaviour):
 *   - 64-bit arithmetic where the reference truncates int64 contig offsets to
 *     C int (_common.pyx:124,165-166); identical below 2 GiB of pooled sequence.
 *   - reads shorter than k=25 are undefined behaviour in the reference
 *     (_kmer.pxd:65 reads past the buffer); here a unit with such a read is
 *     DEFINED to be unaligned with span length 0 (map_unit below).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define SKMO_K 25
#define SKMO_MAX_FRAGMENT_LENGTH 2000 /* _mapper.pyx:18 */
#define SKMO_ALIGN_LENGTH 8           /* _mapper.pyx:22 */
#define SKMO_MAX_OFFSET 2             /* _mapper.pyx:24 */
#define SKMO_MAX_DISTANCE 4           /* _mapper.pyx:26 */
#define SKMO_INVALID_SHIFT 0x7FFF     /* _mapper.pyx:28 */

/* ---- struct layouts: _common.pxd:15-35, _coordinate.pxd:8-10 ------------ */
typedef struct { int32_t entry; int32_t offset; } skmo_coord;
typedef struct { uint64_t kmer; skmo_coord position; } skmo_slot;
typedef struct {
    int64_t offset, length;
    uint64_t first_kmer, last_kmer;
    int64_t target_offset, target_length;
} skmo_contig;
typedef struct { int size; skmo_coord *items; } skmo_list; /* _coordinate_array.pxd:11-13 */
typedef struct { int32_t begin, end; skmo_coord anchor; skmo_list targets; } skmo_span;
typedef struct { int length; const char *bases; } skmo_seq; /* _sequence.pxd:7-9 */

typedef struct {
    const skmo_slot *kmers;   int64_t n_slots;
    const skmo_contig *contigs; int64_t n_contigs;
    const char *sequences;    int64_t n_bases;
    const skmo_coord *targets; int64_t n_targets;
} skmo_index;

/* optional access counters (SURVEY.md §8(d) "counting oracle"); enabled per call */
typedef struct {
    int64_t map_kmer_calls, slots, map_contig_calls, map_contig_items;
    int64_t filter_calls, filter_items, windows, tail_kmers, contig_reads;
} skmo_counters;
static __thread skmo_counters *g_cnt = NULL;
#define CNT(field, n) do { if (g_cnt) g_cnt->field += (n); } while (0)

/* ---- _kmer.pxd ------------------------------------------------------------ */
static inline uint64_t kmer_invalid(void) { return 0xFFFFFFFFFFFFFFFFULL; } /* :20-28 */
static inline uint64_t kmer_mask(void) { return ~(kmer_invalid() << (SKMO_K * 2)); } /* :31-39 */

static inline uint64_t two_bit_encode(char base) /* _kmer.pxd:253-273 */
{
    if (base == 'T' || base == 't') return 3;
    if (base == 'G' || base == 'g') return 2;
    if (base == 'C' || base == 'c') return 1;
    return 0;
}

uint64_t skmo_encode(const char *sequence, int offset) /* _kmer.pxd:46-68 */
{
    uint64_t kmer = 0;
    for (int i = offset; i < offset + SKMO_K; ++i) {
        kmer <<= 2;
        kmer |= two_bit_encode(sequence[i]);
    }
    return kmer;
}

static inline uint64_t kmer_append(uint64_t kmer, char base) /* _kmer.pxd:71-87 */
{
    return ((kmer << 2) | two_bit_encode(base)) & kmer_mask();
}

static inline uint64_t kmer_prepend(uint64_t kmer, char base) /* _kmer.pxd:90-106 */
{
    return (kmer >> 2) | (two_bit_encode(base) << (SKMO_K * 2 - 2));
}

uint64_t skmo_reverse_complement(uint64_t kmer) /* _kmer.pxd:146-171 */
{
    kmer = ((kmer >> 2) & 0x3333333333333333ULL) | ((kmer & 0x3333333333333333ULL) << 2);
    kmer = ((kmer >> 4) & 0x0f0f0f0f0f0f0f0fULL) | ((kmer & 0x0f0f0f0f0f0f0f0fULL) << 4);
    kmer = ((kmer >> 8) & 0x00ff00ff00ff00ffULL) | ((kmer & 0x00ff00ff00ff00ffULL) << 8);
    kmer = ((kmer >> 16) & 0x0000ffff0000ffffULL) | ((kmer & 0x0000ffff0000ffffULL) << 16);
    kmer = (kmer >> 32) | (kmer << 32);
    kmer = kmer >> (64 - SKMO_K * 2);
    return ~kmer & kmer_mask();
}

static inline void sip_half_round(uint64_t *a, uint64_t *b, uint64_t *c, uint64_t *d,
                                  int s, int t) /* _kmer.pxd:222-231 */
{
    *a += *b;
    *c += *d;
    *b = ((*b << s) | (*b >> (64 - s))) ^ *a;
    *d = ((*d << t) | (*d >> (64 - t))) ^ *c;
    *a = (*a << 32) | (*a >> 32);
}

/* _kmer.pxd:174-219.  SipHash-2-4 on one word with the reference's deviation
 * (`v0 ^= 0` at :209 where the standard xors the length block).  Returned as
 * the full 64-bit value; callers keep only the low log2(n_slots) bits exactly
 * as `<int>(...) & (size - 1)` does (_kmer.pxd:219, _common.pyx:77). */
uint64_t skmo_hash(uint64_t kmer)
{
    uint64_t k0 = 5381, k1 = 42, b = 8ULL << 56;
    uint64_t v0 = k0 ^ 0x736f6d6570736575ULL;
    uint64_t v1 = k1 ^ 0x646f72616e646f6dULL;
    uint64_t v2 = k0 ^ 0x6c7967656e657261ULL;
    uint64_t v3 = k1 ^ 0x7465646279746573ULL;
    uint64_t mi = kmer;
    v3 ^= mi;
    sip_half_round(&v0, &v1, &v2, &v3, 13, 16);
    sip_half_round(&v2, &v1, &v0, &v3, 17, 21);
    sip_half_round(&v0, &v1, &v2, &v3, 13, 16);
    sip_half_round(&v2, &v1, &v0, &v3, 17, 21);
    v0 ^= mi;
    v3 ^= b;
    sip_half_round(&v0, &v1, &v2, &v3, 13, 16);
    sip_half_round(&v2, &v1, &v0, &v3, 17, 21);
    sip_half_round(&v0, &v1, &v2, &v3, 13, 16);
    sip_half_round(&v2, &v1, &v0, &v3, 17, 21);
    v0 ^= 0; /* sic, :209 */
    v2 ^= 0xff;
    for (int r = 0; r < 4; ++r) {
        sip_half_round(&v0, &v1, &v2, &v3, 13, 16);
        sip_half_round(&v2, &v1, &v0, &v3, 17, 21);
    }
    return (v0 ^ v1) ^ (v2 ^ v3);
}

/* ---- _coordinate.pxd -------------------------------------------------------- */
static inline skmo_coord coord_invalid(void) { skmo_coord c = {0, -1}; return c; } /* :13-24 */
static inline skmo_coord coord_rc(skmo_coord c) { c.entry = ~c.entry; return c; }     /* :65-81 */
static inline int coord_valid(skmo_coord c) { return c.offset >= 0; }                  /* :84-99 */

/* ---- _coordinate_array.pxd -------------------------------------------------- */
static inline skmo_list list_empty(void) { skmo_list l = {0, NULL}; return l; }
static inline skmo_list list_create(int size)
{
    skmo_list l;
    l.size = size;
    l.items = (skmo_coord *)malloc(sizeof(skmo_coord) * (size > 0 ? size : 1));
    return l;
}
static inline void list_free(skmo_list *l) { free(l->items); *l = list_empty(); }

/* ---- _common.pyx: KMerIndex ------------------------------------------------- */
skmo_coord skmo_map_kmer(const skmo_index *ix, uint64_t kmer) /* _common.pyx:54-97 */
{
    uint64_t rc_kmer = skmo_reverse_complement(kmer);
    int64_t size = ix->n_slots;
    int64_t offset = (int64_t)(skmo_hash(kmer < rc_kmer ? kmer : rc_kmer) & (uint64_t)(size - 1));
    CNT(map_kmer_calls, 1);
    for (int64_t i = offset; i < size; ++i) {
        CNT(slots, 1);
        if (ix->kmers[i].kmer == kmer_invalid()) return coord_invalid();
        if (ix->kmers[i].kmer == kmer) return ix->kmers[i].position;
        if (ix->kmers[i].kmer == rc_kmer) return coord_rc(ix->kmers[i].position);
    }
    for (int64_t i = 0; i < offset; ++i) {
        CNT(slots, 1);
        if (ix->kmers[i].kmer == kmer_invalid()) return coord_invalid();
        if (ix->kmers[i].kmer == kmer) return ix->kmers[i].position;
        if (ix->kmers[i].kmer == rc_kmer) return coord_rc(ix->kmers[i].position);
    }
    return coord_invalid();
}

/* _sequence.pxd:53-75: in-place reverse complement, only uppercase ACGT swap */
static void seq_rc(char *bases, int length)
{
    for (int i = 0; i < length / 2; ++i) {
        char t = bases[length - i - 1];
        bases[length - i - 1] = bases[i];
        bases[i] = t;
    }
    for (int i = 0; i < length; ++i) {
        if (bases[i] == 'A') bases[i] = 'T';
        else if (bases[i] == 'T') bases[i] = 'A';
        else if (bases[i] == 'C') bases[i] = 'G';
        else if (bases[i] == 'G') bases[i] = 'C';
    }
}

/* _common.pyx:103-137.  `out` holds |length|+1 bytes, NUL-terminated like
 * _sequence.create (_sequence.pxd:25-29). */
static int get_contig_sequence(const skmo_index *ix, skmo_coord c, int length, char *out)
{
    int64_t index = c.entry;
    if (index < 0) index = ~index;
    int64_t offset = ix->contigs[index].offset + c.offset;
    CNT(windows, 1); CNT(contig_reads, 1);
    if (c.entry >= 0) offset += length > 0 ? length : SKMO_K;
    else offset += length > 0 ? SKMO_K : -length;
    if (length < 0) length = -length;
    for (int64_t i = offset - length; i < offset; ++i)
        out[i - offset + length] = ix->sequences[i];
    out[length] = 0;
    if (c.entry < 0) seq_rc(out, length);
    return length;
}

static skmo_list map_contig(const skmo_index *ix, skmo_coord c) /* _common.pyx:143-179 */
{
    int64_t index = c.entry;
    int forward = index >= 0;
    if (!forward) index = ~index;
    int64_t start = ix->contigs[index].target_offset;
    int64_t length = ix->contigs[index].target_length;
    skmo_list targets = list_create((int)length);
    CNT(map_contig_calls, 1); CNT(map_contig_items, length); CNT(contig_reads, 1);
    if (forward) {
        for (int64_t i = 0; i < length; ++i) targets.items[i] = ix->targets[start + i];
    } else {
        int64_t i = 0;
        for (int64_t j = start + length - 1; j > start - 1; --j, ++i)
            targets.items[i] = coord_rc(ix->targets[j]);
    }
    return targets;
}

static int filter_on_contig(const skmo_index *ix, skmo_span *span) /* _common.pyx:185-235 */
{
    if (span->targets.size == 0) return 1;
    int64_t contig_id = span->anchor.entry;
    int forward = contig_id >= 0;
    if (!forward) contig_id = ~contig_id;
    int64_t start = ix->contigs[contig_id].target_offset;
    int64_t length = ix->contigs[contig_id].target_length;
    int read_index = 0, write_index = 0;
    int64_t track_index = forward ? start : start + length - 1;
    int64_t track_bound = forward ? start + length : start - 1;
    int64_t step = forward ? 1 : -1;
    CNT(filter_calls, 1); CNT(filter_items, length); CNT(contig_reads, 1);
    while (read_index != span->targets.size && track_index != track_bound) {
        int32_t target_entry = span->targets.items[read_index].entry;
        int32_t index_entry = ix->targets[track_index].entry;
        if (!forward) index_entry = ~index_entry;
        if (target_entry == index_entry) {
            span->targets.items[write_index] = span->targets.items[read_index];
            read_index += 1;
            write_index += 1;
            track_index += step;
        } else if (target_entry < index_entry) {
            read_index += 1;
        } else {
            track_index += step;
        }
    }
    if (write_index == 0) return 0;
    span->targets.size = write_index;
    return 1;
}

static uint64_t get_tail_kmer(const skmo_index *ix, skmo_coord c) /* _common.pyx:241-266 */
{
    int64_t index = c.entry;
    if (index < 0) index = ~index;
    uint64_t kmer = c.offset == 0 ? ix->contigs[index].first_kmer : ix->contigs[index].last_kmer;
    CNT(tail_kmers, 1); CNT(contig_reads, 1);
    if (c.entry < 0) kmer = skmo_reverse_complement(kmer);
    return kmer;
}

/* ---- _mapper.pyx ------------------------------------------------------------ */
static inline int match_base(int reference, int query) /* _mapper.pyx:500-501 */
{
    return reference == query || !(query == 'A' || query == 'C' || query == 'G' || query == 'T');
}

int skmo_sift4_align_left(const char *ref, int ref_length, const char *query, int query_length,
                          int offset) /* _mapper.pyx:404-445 */
{
    (void)query_length;
    int reference_cursor = ref_length - 1;
    int query_cursor = offset + SKMO_ALIGN_LENGTH - 1;
    query_cursor -= 1;
    int distance = 0;
    while (reference_cursor >= 0 && query_cursor >= offset) {
        if (match_base(ref[reference_cursor], query[query_cursor])) {
            reference_cursor -= 1;
            query_cursor -= 1;
            continue;
        }
        if (reference_cursor != query_cursor - offset) {
            int m = query_cursor - offset;
            reference_cursor = m < reference_cursor ? m : reference_cursor;
            query_cursor = reference_cursor + offset;
        }
        for (int i = 0; i < SKMO_MAX_OFFSET; ++i) {
            if (query_cursor - i >= offset - 1 && query_cursor - i >= 0
                && match_base(ref[reference_cursor], query[query_cursor - i])) {
                distance += i - 1;
                query_cursor -= i - 1;
                reference_cursor += 1;
                break;
            }
            if (reference_cursor - i >= 0
                && match_base(ref[reference_cursor - i], query[query_cursor])) {
                distance += i - 1;
                query_cursor += 1;
                reference_cursor -= i - 1;
                break;
            }
        }
        distance += 1;
        query_cursor -= 1;
        reference_cursor -= 1;
        if (distance > SKMO_MAX_DISTANCE) return SKMO_INVALID_SHIFT;
    }
    if (reference_cursor >= 0) return reference_cursor + 1;
    if (query_cursor >= offset) return -1 - query_cursor + offset;
    return 0;
}

int skmo_sift4_align_right(const char *ref, int ref_length, const char *query, int query_length,
                           int offset) /* _mapper.pyx:452-493 */
{
    int reference_cursor = 0;
    int query_cursor = offset;
    int distance = 0;
    while (reference_cursor < ref_length && query_cursor < offset + SKMO_ALIGN_LENGTH) {
        if (match_base(ref[reference_cursor], query[query_cursor])) {
            reference_cursor += 1;
            query_cursor += 1;
            continue;
        }
        if (reference_cursor != query_cursor - offset) {
            int m = query_cursor - offset;
            reference_cursor = m > reference_cursor ? m : reference_cursor;
            query_cursor = reference_cursor + offset;
        }
        for (int i = 0; i < SKMO_MAX_OFFSET; ++i) {
            if (query_cursor + i < offset + SKMO_ALIGN_LENGTH + 1
                && query_cursor + i < query_length
                && match_base(ref[reference_cursor], query[query_cursor + i])) {
                distance += i - 1;
                query_cursor += i - 1;
                reference_cursor -= 1;
                break;
            }
            if (reference_cursor + i < ref_length
                && match_base(ref[reference_cursor + i], query[query_cursor])) {
                distance += i - 1;
                query_cursor -= 1;
                reference_cursor += i - 1;
                break;
            }
        }
        distance += 1;
        query_cursor += 1;
        reference_cursor += 1;
        if (distance > SKMO_MAX_DISTANCE) return SKMO_INVALID_SHIFT;
    }
    if (reference_cursor < ref_length) return ref_length - reference_cursor;
    if (query_cursor < offset + SKMO_ALIGN_LENGTH) return query_cursor - offset - SKMO_ALIGN_LENGTH;
    return 0;
}

static void find_first_kmer(const skmo_index *ix, skmo_seq read, skmo_span *span) /* :199-216 */
{
    uint64_t kmer = skmo_encode(read.bases, span->begin);
    span->anchor = skmo_map_kmer(ix, kmer);
    if (span->anchor.offset >= 0) {
        span->end = span->begin;
        span->targets = map_contig(ix, span->anchor);
        return;
    }
    for (int i = span->begin + SKMO_K; i < read.length; ++i) {
        kmer = kmer_append(kmer, read.bases[i]);
        span->anchor = skmo_map_kmer(ix, kmer);
        if (span->anchor.offset < 0) continue;
        span->begin = i + 1 - SKMO_K;
        span->end = span->begin;
        span->targets = map_contig(ix, span->anchor);
        return;
    }
}

static void filter_targets_to_left(const skmo_index *ix, skmo_seq read, skmo_span *span) /* :222-275 */
{
    uint64_t kmer;
    int forward = span->anchor.entry >= 0;
    int32_t contig_index = forward ? span->anchor.entry : ~span->anchor.entry;
    int contig_length = (int)ix->contigs[contig_index].length;
    int move = forward ? span->anchor.offset : contig_length - span->anchor.offset - SKMO_K;
    char contig[SKMO_ALIGN_LENGTH + 1];
    int shift;
    CNT(contig_reads, 1);
    while (span->begin > move) {
        span->begin -= move;
        span->anchor.offset -= forward ? move : -move;
        get_contig_sequence(ix, span->anchor, SKMO_ALIGN_LENGTH, contig);
        shift = skmo_sift4_align_left(contig, SKMO_ALIGN_LENGTH, read.bases, read.length, span->begin);
        if (shift == SKMO_INVALID_SHIFT || shift + 1 + move <= 0) {
            list_free(&span->targets);
            return;
        }
        span->begin -= shift + 1;
        if (span->begin < 0) {
            span->begin = 0;
            return;
        }
        kmer = kmer_prepend(get_tail_kmer(ix, span->anchor), read.bases[span->begin]);
        span->anchor = skmo_map_kmer(ix, kmer);
        if (!coord_valid(span->anchor) || !filter_on_contig(ix, span)) {
            if (span->begin < SKMO_K) {
                span->begin = 0;
                return;
            }
            span->begin -= SKMO_K;
            kmer = skmo_encode(read.bases, span->begin);
            span->anchor = skmo_map_kmer(ix, kmer);
            if (!coord_valid(span->anchor) || !filter_on_contig(ix, span)) {
                list_free(&span->targets);
                return;
            }
        }
        forward = span->anchor.entry >= 0;
        contig_index = forward ? span->anchor.entry : ~span->anchor.entry;
        contig_length = (int)ix->contigs[contig_index].length;
        CNT(contig_reads, 1);
        move = forward ? span->anchor.offset : contig_length - span->anchor.offset - SKMO_K;
    }
    span->anchor.offset -= forward ? span->begin : -span->begin;
    get_contig_sequence(ix, span->anchor, SKMO_ALIGN_LENGTH, contig);
    shift = skmo_sift4_align_left(contig, SKMO_ALIGN_LENGTH, read.bases, read.length, 0);
    if (shift == SKMO_INVALID_SHIFT) list_free(&span->targets);
}

static void filter_targets_to_right(const skmo_index *ix, skmo_seq read, skmo_span *span) /* :281-343 */
{
    uint64_t kmer = skmo_encode(read.bases, span->end);
    span->anchor = skmo_map_kmer(ix, kmer);
    int forward = span->anchor.entry >= 0;
    int32_t contig_index = forward ? span->anchor.entry : ~span->anchor.entry;
    int contig_length = (int)ix->contigs[contig_index].length;
    int move = forward ? contig_length - span->anchor.offset - SKMO_K : span->anchor.offset;
    char contig[SKMO_ALIGN_LENGTH + 1];
    int shift;
    CNT(contig_reads, 1);
    while (read.length - span->end - SKMO_K > move) {
        span->end += move;
        span->anchor.offset += forward ? move : -move;
        get_contig_sequence(ix, span->anchor, -SKMO_ALIGN_LENGTH, contig);
        shift = skmo_sift4_align_right(contig, SKMO_ALIGN_LENGTH, read.bases, read.length,
                                       span->end + SKMO_K - SKMO_ALIGN_LENGTH);
        if (shift == SKMO_INVALID_SHIFT || shift + 1 + move <= 0) {
            list_free(&span->targets);
            return;
        }
        span->end += shift + 1;
        if (span->end + SKMO_K > read.length) {
            span->end = read.length - SKMO_K;
            return;
        }
        kmer = kmer_append(get_tail_kmer(ix, span->anchor), read.bases[span->end + SKMO_K - 1]);
        span->anchor = skmo_map_kmer(ix, kmer);
        if (!coord_valid(span->anchor) || !filter_on_contig(ix, span)) {
            list_free(&span->targets);
            return;
        }
        /* :316-329 — unreachable in practice (same test as above just passed and
         * _filter_on_contig is idempotent); kept to mirror the reference. */
        if (!coord_valid(span->anchor) || !filter_on_contig(ix, span)) {
            if (span->end > read.length - 2 * SKMO_K) {
                span->end = read.length - SKMO_K;
                return;
            }
            span->end += SKMO_K;
            kmer = skmo_encode(read.bases, span->end);
            span->anchor = skmo_map_kmer(ix, kmer);
            if (!coord_valid(span->anchor) || !filter_on_contig(ix, span)) {
                list_free(&span->targets);
                return;
            }
        }
        forward = span->anchor.entry >= 0;
        contig_index = forward ? span->anchor.entry : ~span->anchor.entry;
        contig_length = (int)ix->contigs[contig_index].length;
        CNT(contig_reads, 1);
        move = forward ? contig_length - span->anchor.offset - SKMO_K : span->anchor.offset;
    }
    if (forward) span->anchor.offset += read.length - span->end - SKMO_K;
    else span->anchor.offset -= read.length - span->end - SKMO_K;
    get_contig_sequence(ix, span->anchor, -SKMO_ALIGN_LENGTH, contig);
    shift = skmo_sift4_align_right(contig, SKMO_ALIGN_LENGTH, read.bases, read.length,
                                   read.length - SKMO_ALIGN_LENGTH);
    if (shift == SKMO_INVALID_SHIFT) list_free(&span->targets);
}

static skmo_span map_read(const skmo_index *ix, skmo_seq read) /* _mapper.pyx:151-193 */
{
    skmo_span span;
    span.anchor = coord_invalid();
    span.begin = 0;
    span.end = span.begin;
    span.targets = list_empty();
    find_first_kmer(ix, read, &span);
    if (span.targets.size == 0) return span;
    if (span.begin > 0) filter_targets_to_left(ix, read, &span);
    if (span.targets.size != 0 && span.end < read.length - SKMO_K)
        filter_targets_to_right(ix, read, &span);
    if (span.targets.size != 0) return span;
    span.anchor = coord_invalid();
    span.targets = list_empty();
    span.begin += SKMO_K;
    if (span.begin + SKMO_K > read.length) span.begin = read.length - SKMO_K;
    span.end = span.begin;
    find_first_kmer(ix, read, &span);
    if (span.targets.size == 0) return span;
    if (span.begin > 0) filter_targets_to_left(ix, read, &span);
    if (span.targets.size != 0 && span.end < read.length - SKMO_K)
        filter_targets_to_right(ix, read, &span);
    return span;
}

static int intersect(skmo_span *r1, skmo_span *r2) /* _mapper.pyx:350-397 */
{
    if (r1->targets.size == 0) return 1;
    if (r2->targets.size == 0) return 0;
    int cursor1_read = 0, cursor1_write = 0;
    int cursor2 = r2->targets.size - 1;
    while (cursor1_read != r1->targets.size && cursor2 != -1) {
        int entry1 = r1->targets.items[cursor1_read].entry;
        int entry2 = ~r2->targets.items[cursor2].entry;
        if (entry1 == entry2) {
            r1->targets.items[cursor1_write] = r1->targets.items[cursor1_read];
            cursor1_read += 1;
            cursor1_write += 1;
            cursor2 -= 1;
        } else if (entry1 < entry2) {
            cursor1_read += 1;
        } else {
            cursor2 -= 1;
        }
    }
    if (cursor1_write == 0) return 0;
    r1->targets.size = cursor1_write;
    return 1;
}

static skmo_span map_read_pair(const skmo_index *ix, skmo_seq read1, skmo_seq read2) /* :111-145 */
{
    skmo_span span1 = map_read(ix, read1);
    skmo_span span2 = map_read(ix, read2);
    int interval = 0;
    if (!intersect(&span1, &span2)) {
        list_free(&span1.targets);
        span1.begin = 0;
        span1.end = span1.begin - SKMO_K;
    } else if (span1.anchor.entry != ~span2.anchor.entry) {
        span1.begin = 0;
        span1.end = span1.begin - SKMO_K;
    } else {
        span1.end = read1.length - SKMO_K;
        span2.end = read2.length - SKMO_K;
        interval = span2.anchor.offset - span1.anchor.offset;
        if (span1.anchor.entry < 0) interval = -interval;
        span1.end += interval + span2.end - span2.begin;
    }
    list_free(&span2.targets);
    return span1;
}

/* ---- batch driver: ReadMapper.__call__ (_mapper.pyx:59-105) ------------------
 *
 * bases: all reads concatenated; offsets[n_reads + 1]; reads of a pair are
 * interleaved (2i, 2i+1) as in :86-88.  Per unit (read or pair) it returns the
 * ordered id tuple of _get_ids (:528-537) in CSR form, the raw
 * `span.end - span.begin + k` value (:90) and accumulates the FLD (:91-94).
 *
 * out_ptr[n_units + 1]; out_ids capacity out_cap; returns total ids, or
 * -(needed) if out_cap was too small (ptr/len/fld are still complete).
 */
/* A read shorter than k is undefined behaviour in the reference (_kmer.pxd:46-68 reads past its
 * end).  The behaviour is DEFINED here, for the oracle and the CUDA path alike: a unit with such
 * a read is unaligned (empty tuple) with span length 0, whatever its mate maps to. */
static skmo_span map_unit(const skmo_index *ix, const char *bases, const int64_t *offsets, int64_t i, int paired)
{
    if (!paired) {
        skmo_seq r = {(int)(offsets[i + 1] - offsets[i]), bases + offsets[i]};
        if (r.length >= SKMO_K) return map_read(ix, r);
    } else {
        skmo_seq r1 = {(int)(offsets[2 * i + 1] - offsets[2 * i]), bases + offsets[2 * i]};
        skmo_seq r2 = {(int)(offsets[2 * i + 2] - offsets[2 * i + 1]), bases + offsets[2 * i + 1]};
        if (r1.length >= SKMO_K && r2.length >= SKMO_K) return map_read_pair(ix, r1, r2);
    }
    skmo_span span;
    memset(&span, 0, sizeof(span));
    span.begin = 0;
    span.end = -SKMO_K;
    return span;
}

int64_t skmo_map_batch(const skmo_index *ix, const char *bases, const int64_t *offsets,
                       int64_t n_units, int paired, int64_t *out_ptr, int32_t *out_ids,
                       int64_t out_cap, int32_t *out_length, int64_t *fld, skmo_counters *counters)
{
    int64_t total = 0;
    g_cnt = counters;
    out_ptr[0] = 0;
    for (int64_t i = 0; i < n_units; ++i) {
        skmo_span span = map_unit(ix, bases, offsets, i, paired);
        int length = span.end - span.begin + SKMO_K;
        if (out_length) out_length[i] = length;
        if (length > 0) {
            if (length >= SKMO_MAX_FRAGMENT_LENGTH) length = SKMO_MAX_FRAGMENT_LENGTH - 1;
            if (fld) fld[length] += 1;
        }
        for (int j = 0; j < span.targets.size; ++j) {
            int32_t entry = span.targets.items[j].entry;
            if (entry < 0) entry = ~entry;
            if (total + j < out_cap) out_ids[total + j] = entry;
        }
        total += span.targets.size;
        out_ptr[i + 1] = total;
        list_free(&span.targets);
    }
    g_cnt = NULL;
    return total <= out_cap ? total : -total;
}

/* Multi-threaded driver used only for the CPU baseline timing: units are split
 * into contiguous chunks (the reference's data-parallel threads, mapper.py:174-189,
 * minus the GIL).  Produces per-unit tuple hash + length instead of CSR so no
 * cross-thread compaction is timed; fld is merged at the end like
 * merge_fragment_lengths (mapper.py:106-115). */
#ifdef _OPENMP
#include <omp.h>
#endif
static inline uint64_t mix64(uint64_t x)
{
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return x;
}
int64_t skmo_map_batch_mt(const skmo_index *ix, const char *bases, const int64_t *offsets,
                          int64_t n_units, int paired, int n_threads, uint64_t *out_hash,
                          int32_t *out_count, int32_t *out_length, int64_t *fld)
{
    int64_t aligned = 0;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel reduction(+ : aligned)
    {
        int64_t local_fld[SKMO_MAX_FRAGMENT_LENGTH];
        memset(local_fld, 0, sizeof(local_fld));
#pragma omp for schedule(dynamic, 4096)
        for (int64_t i = 0; i < n_units; ++i) {
            skmo_span span = map_unit(ix, bases, offsets, i, paired);
            int length = span.end - span.begin + SKMO_K;
            if (out_length) out_length[i] = length;
            if (length > 0) {
                if (length >= SKMO_MAX_FRAGMENT_LENGTH) length = SKMO_MAX_FRAGMENT_LENGTH - 1;
                local_fld[length] += 1;
            }
            uint64_t h = 0x9E3779B97F4A7C15ULL;
            for (int j = 0; j < span.targets.size; ++j) {
                int32_t entry = span.targets.items[j].entry;
                if (entry < 0) entry = ~entry;
                h = mix64(h ^ (uint64_t)(uint32_t)entry) + 0x632BE59BD9B4E019ULL;
            }
            if (out_hash) out_hash[i] = h;
            if (out_count) out_count[i] = span.targets.size;
            aligned += span.targets.size != 0;
            list_free(&span.targets);
        }
#pragma omp critical
        {
            if (fld) for (int k = 0; k < SKMO_MAX_FRAGMENT_LENGTH; ++k) fld[k] += local_fld[k];
        }
    }
    return aligned;
}

/* ---- MapResult.update/summarize tally (mapper.py:60-104) ---------------------
 *
 * Counter keyed by the ordered id tuple; class order = first insertion
 * (what `collections.Counter` iteration gives at job_count=1).  The empty tuple
 * (unaligned) is counted separately like `counter.pop((), 0)` (mapper.py:87).
 *
 * In: CSR of per-unit tuples.  Out: CSR of distinct classes + counts.
 * Returns number of classes; *unaligned gets the () count.
 */
typedef struct { int64_t start; int32_t len; int64_t count; uint64_t hash; } tally_ent;

int64_t skmo_tally(const int64_t *ptr, const int32_t *ids, int64_t n_units,
                   int64_t *cls_ptr, int32_t *cls_ids, int64_t *cls_count, int64_t *unaligned)
{
    int64_t cap = 1024;
    while (cap < n_units * 2) cap <<= 1;
    int64_t *table = (int64_t *)malloc(sizeof(int64_t) * cap);
    tally_ent *ents = (tally_ent *)malloc(sizeof(tally_ent) * (n_units > 0 ? n_units : 1));
    for (int64_t i = 0; i < cap; ++i) table[i] = -1;
    int64_t n_cls = 0, una = 0;
    for (int64_t u = 0; u < n_units; ++u) {
        int32_t len = (int32_t)(ptr[u + 1] - ptr[u]);
        if (len == 0) { una += 1; continue; }
        const int32_t *t = ids + ptr[u];
        uint64_t h = 0x9E3779B97F4A7C15ULL;
        for (int j = 0; j < len; ++j) h = mix64(h ^ (uint64_t)(uint32_t)t[j]) + 0x632BE59BD9B4E019ULL;
        int64_t s = (int64_t)(h & (uint64_t)(cap - 1));
        for (;;) {
            int64_t e = table[s];
            if (e < 0) {
                ents[n_cls].start = ptr[u]; ents[n_cls].len = len; ents[n_cls].count = 1; ents[n_cls].hash = h;
                table[s] = n_cls++;
                break;
            }
            if (ents[e].hash == h && ents[e].len == len
                && memcmp(ids + ents[e].start, t, sizeof(int32_t) * len) == 0) {
                ents[e].count += 1;
                break;
            }
            s = (s + 1) & (cap - 1);
        }
    }
    int64_t pos = 0;
    cls_ptr[0] = 0;
    for (int64_t c = 0; c < n_cls; ++c) {
        memcpy(cls_ids + pos, ids + ents[c].start, sizeof(int32_t) * ents[c].len);
        pos += ents[c].len;
        cls_ptr[c + 1] = pos;
        cls_count[c] = ents[c].count;
    }
    *unaligned = una;
    free(table);
    free(ents);
    return n_cls;
}

/* MapResult.effective_lengths (mapper.py:134-141): accumulated in i order. */
void skmo_effective_lengths(const int64_t *fld, const double *length, int64_t n, double *out)
{
    double total = 0;
    /* numpy int64 sum then true-divide per element: p[i] = fld[i] / sum */
    int64_t s = 0;
    for (int i = 0; i < SKMO_MAX_FRAGMENT_LENGTH; ++i) s += fld[i];
    total = (double)s;
    for (int64_t t = 0; t < n; ++t) out[t] = 0.0;
    for (int i = 0; i < SKMO_MAX_FRAGMENT_LENGTH; ++i) {
        double p = (double)fld[i] / total;
        for (int64_t t = 0; t < n; ++t) {
            double v = length[t] - (double)i;
            if (v < 1.0) v = 1.0;
            out[t] += v * p;
        }
    }
}

/* Sequential open-addressing insert in the reference layout
 * (find_slot + store, _index_builder.pyx:181-198,313-342) — used by the test-side
 * index builder for synthetic transcriptomes on the CPU. `kmers` are stored
 * as given (contig-forward orientation); table must be pre-filled with empties. */
int64_t skmo_table_insert(skmo_slot *table, int64_t n_slots, const uint64_t *kmers,
                          const int32_t *entry, const int32_t *offset, int64_t n)
{
    for (int64_t q = 0; q < n; ++q) {
        uint64_t kmer = kmers[q];
        uint64_t rc = skmo_reverse_complement(kmer);
        int64_t i = (int64_t)(skmo_hash(kmer < rc ? kmer : rc) & (uint64_t)(n_slots - 1));
        int64_t probes = 0;
        while (table[i].kmer != kmer_invalid()) {
            if (table[i].kmer == kmer || table[i].kmer == rc) return -(q + 1); /* duplicate */
            i = (i + 1) & (n_slots - 1);
            if (++probes > n_slots) return -(q + 1);
        }
        table[i].kmer = kmer;
        table[i].position.entry = entry[q];
        table[i].position.offset = offset[q];
    }
    return n;
}
