// cfb_vlc.cu -- host side of SURVEY 8f rank 1: the run-length / VLC stream of a band straight from the sparse
// transfer format (no CUDA in this file; it is a .cu only so that the one-line build picks it up).
//
// Replaces the walk of the reference's coder over a DENSE band:
//   Codec/encoder.c:5386-5700 EncodeQuantLongRuns   rows of `width` coefficients, zero runs carried across the pitch
//                                                   gap (:5653) and across rows, pending run flushed at the end (:5671)
//   Codec/vlc.c:366 PutZeroRun (inlined :5493-5545) greedy split of a run: entry min(count, length - 1), count -= entry.count
//   Codec/vlc.c:188 PutVlcByte (inlined :5553-5568) value clamped to +-(VALUE_TABLE_LENGTH / 2 - 1), negative values index
//                                                   from the top of the table
//   Codec/bitstream.c:819 PutBits (inlined)         32-bit buffer, flushed big-endian only when a code does NOT fit
// The sparse format already holds the positions of the non-zero coefficients (two-level bitmaps) and their values, so
// a zero run is the distance between consecutive set bits of the flat coded region -- the pitch gap and the 64-byte band
// alignment are zero words of that region -- and the host reads ~2 MB per 4K frame instead of 33 MB.
#include <stdint.h>
#include <string.h>

#include <new>
#include <vector>

#include "cfb_host.h"
#include "cfb_sparse_format.h"

using namespace cfb;

namespace {

// the reference's bit buffer, widened: the low `n` bits of `acc` are pending (n <= 32 between codes).  A word leaves the
// buffer only when a code does not fit any more (n + size > 32), exactly as PutBits does, so the state left behind is the
// reference's.  Whether a word leaves is data dependent and close to a coin flip per code, so put() has no branch on it:
// the candidate word is stored at `cur` after every code and `cur` advances by 0 or 4.  (Up to 4 bytes behind the
// returned position are therefore scratch, inside [cur, end).)
struct Bits {
    uint64_t acc;
    int n;
    uint8_t *cur, *end, *start;
    int64_t bytes0;
    bool overflow;
    explicit Bits(const cfb_bitwriter &bw) : acc(bw.bits_free >= 32 ? 0 : (bw.buffer & (0xffffffffu >> bw.bits_free))), n(32 - bw.bits_free),
                                             cur(bw.cur), end(bw.end), start(bw.cur), bytes0(bw.bytes), overflow(false) {}
    // `bits` has no set bit above `size`
    inline void put_masked(uint32_t bits, int size)
    {
        acc = (acc << size) | bits;
        n += size;
        if (__builtin_expect(cur + 4 > end, 0)) {
            if (n > 32) { overflow = true; n -= 32; }
            return;
        }
        const uint32_t be = __builtin_bswap32((uint32_t)(acc >> ((n - 32) & 63)));      // big-endian in the stream (bitstream.c PutLong)
        memcpy(cur, &be, 4);
        const int full = n > 32;
        cur += 4 * full;
        n -= 32 * full;
    }
    inline void put(uint32_t bits, int size) { put_masked(bits & (0xffffffffu >> (32 - size)), size); }
    void store(cfb_bitwriter *bw) const
    {
        bw->cur = cur; bw->bytes = bytes0 + (cur - start);
        bw->bits_free = 32 - n;
        bw->buffer = n ? (uint32_t)(acc & (0xffffffffull >> (32 - n))) : 0u;
    }
};

// (short zero run, small value) pairs are most of a band's tokens: their run codes + value code are pre-joined into one
// code word of at most 32 bits.  Writing the joined word leaves the bit buffer in the same state as writing its parts one
// after the other (a word is flushed exactly when the pending count passes 32 either way, and two flushes cannot occur
// within 32 bits).
constexpr int kFastRuns = 32, kFastValues = 32;
struct FastPairs {
    uint64_t key = 0;
    uint64_t e[kFastRuns][2 * kFastValues + 1];     // (size << 32) | bits, 0 = not available
};

uint64_t book_key(const cfb_vlc_codebook &b)
{
    uint64_t h = 1469598103934665603ull ^ (uint64_t)b.run_length ^ ((uint64_t)b.value_length << 32);
    auto mix = [&](uint64_t x) { h ^= x; h *= 1099511628211ull; h ^= h >> 31; };
    for (int i = 1; i < kFastRuns && i < b.run_length; i++) mix(((uint64_t)b.run_bits[i] << 24) ^ ((uint64_t)b.run_size[i] << 16) ^ b.run_count[i]);
    const int half = b.value_length >> 1;
    for (int v = -kFastValues; v <= kFastValues; v++) {
        if (v <= -half || v >= half) continue;          // a table shorter than the fast range
        const int idx = v < 0 ? b.value_length + v : v;
        mix(((uint64_t)b.value_bits[idx] << 8) ^ b.value_size[idx]);
    }
    return h | 1;
}

const FastPairs &fast_pairs(const cfb_vlc_codebook &b)
{
    static thread_local FastPairs cache[4];
    static thread_local int next = 0;
    const uint64_t key = book_key(b);
    for (FastPairs &f : cache) if (f.key == key) return f;
    FastPairs &f = cache[next++ & 3];
    f.key = key;
    const int half = b.value_length >> 1;
    for (int r = 0; r < kFastRuns; r++) {
        // the greedy split of a run of r zeros (vlc.c:366)
        uint64_t rbits = 0; int rsize = 0; bool ok = true;
        for (int count = r; count > 0 && ok; ) {
            const int i = count < b.run_length - 1 ? count : b.run_length - 1;
            if (rsize + b.run_size[i] > 32) { ok = false; break; }
            rbits = (rbits << b.run_size[i]) | (b.run_bits[i] & (0xffffffffu >> (32 - b.run_size[i])));
            rsize += b.run_size[i];
            count -= (int)b.run_count[i];
        }
        for (int v = -kFastValues; v <= kFastValues; v++) {
            uint64_t &e = f.e[r][v + kFastValues];
            e = 0;
            if (!ok || v == 0 || v <= -half || v >= half) continue;
            const int idx = v < 0 ? b.value_length + v : v;
            const int size = rsize + b.value_size[idx];
            if (size > 32) continue;
            const uint64_t bits = (rbits << b.value_size[idx]) | (b.value_bits[idx] & (0xffffffffu >> (32 - b.value_size[idx])));
            e = ((uint64_t)size << 32) | bits;
        }
    }
    return f;
}

struct Coder {
    const cfb_vlc_codebook &b;
    Bits &out;
    const int half;
    const FastPairs &fast;
    Coder(const cfb_vlc_codebook &book, Bits &o, const FastPairs &f) : b(book), out(o), half(book.value_length >> 1), fast(f) {}
    Coder(const cfb_vlc_codebook &book, Bits &o) : Coder(book, o, fast_pairs(book)) {}
    inline void token(uint64_t zeros, int v)
    {
        if (zeros < (uint64_t)kFastRuns && v >= -kFastValues && v <= kFastValues) {
            const uint64_t e = fast.e[zeros][v + kFastValues];
            if (e) { out.put_masked((uint32_t)e, (int)(e >> 32)); return; }
        }
        run(zeros);
        value(v);
    }
    inline void run(uint64_t count)
    {
        const uint64_t last = (uint64_t)b.run_length - 1;
        while (count > 0) {
            const uint64_t i = count < last ? count : last;
            out.put(b.run_bits[i], b.run_size[i]);
            count -= b.run_count[i];
        }
    }
    inline void value(int v)
    {
        int idx;
        if (v < 0) { if (v <= -half) v = -(half - 1); idx = b.value_length + v; }
        else { if (v >= half) v = half - 1; idx = v; }
        out.put(b.value_bits[idx], b.value_size[idx]);
    }
};

bool book_ok(const cfb_vlc_codebook *b)
{
    if (!b || b->run_length < 2 || b->value_length < 4 || (b->value_length & 1) || !b->run_bits || !b->run_size || !b->run_count ||
        !b->value_bits || !b->value_size) return false;
    // entry i must not cover more zeros than the run it is chosen for (the reference's loop would run past zero otherwise)
    for (int i = 1; i < b->run_length; i++)
        if (b->run_count[i] < 1 || b->run_count[i] > (uint32_t)i || b->run_size[i] < 1 || b->run_size[i] > 31) return false;
    return true;
}

bool writer_ok(const cfb_bitwriter *bw) { return bw && bw->cur && bw->end && bw->cur <= bw->end && bw->bits_free >= 0 && bw->bits_free <= 32; }

// flat word range of a band inside the coded region
cfb_error band_range(const cfb_layout *L, int channel, int level, int band, size_t *w0, size_t *w1, const cfb_band_layout **bl)
{
    if (!L || channel < 0 || channel >= L->num_channels || level < 0 || level >= CFB_NUM_LEVELS || band < 0 || band >= CFB_NUM_BANDS) {
        set_error("bad band (%d, %d, %d)", channel, level, band);
        return CFB_ERROR_INVALID_ARGUMENT;
    }
    if (band == 0 && level != CFB_NUM_LEVELS - 1) { set_error("LL of level %d is not in the coded region", level + 1); return CFB_ERROR_INVALID_ARGUMENT; }
    const cfb_band_layout &b = L->band[channel][level][band];
    *w0 = (size_t)b.offset / 2;
    *w1 = *w0 + (size_t)(b.pitch / 2) * b.height;
    if (bl) *bl = &b;
    return CFB_OK;
}

// Positions (word offsets inside the block, ascending) of a chunk's non-zero words -> pos[]; false = the bitmaps do not
// add up to the table entry's counts.  pos[] has room for a whole block plus one group.
bool chunk_positions(const SparseChunk &c, uint16_t *pos)
{
    const uint64_t *l1w = (const uint64_t *)c.l1;           // chunks are 16-byte aligned
    unsigned gi = 0, k = 0;
    for (unsigned q = 0; q < kSparseBlockGroups / 64; q++) {
        uint64_t bits = l1w[q];
        while (bits) {
            const unsigned g = q * 64 + (unsigned)__builtin_ctzll(bits);
            bits &= bits - 1;
            if (gi >= c.groups) return false;
            unsigned m = c.masks[gi++];
            if (!m) return false;
            const unsigned wbase = g * kSparseGroupWords;
            do {
                pos[k++] = (uint16_t)(wbase + (unsigned)__builtin_ctz(m));
                m &= m - 1;
            } while (m);
        }
    }
    return gi == c.groups && k == c.values;
}

// The non-zero words of the flat range [w0, w1) in raster order, one call per block:
//   f(base, pos, values, wide, n): word base + pos[i] holds values[i], or the next entry of wide[] if values[i] is -128.
// Two tight loops per block (bitmaps -> positions here, positions + value bytes -> codes in the caller) instead of one with
// everything live at once: the single loop kept its counters on the stack and ran at the store-forwarding latency.
template <class F>
cfb_error walk_blocks(const cfb_layout *L, const void *sparse, size_t w0, size_t w1, F &&f)
{
    const unsigned nwords = (unsigned)(L->coded_bytes / 2);
    const unsigned *h = (const unsigned *)sparse;
    const unsigned nblocks = sparse_nblocks(nwords);
    // total_bytes is bounded by the worst case, i.e. by the size every sparse buffer is required to have (cfb_sparse_max_bytes)
    if (!sparse || h[0] != kSparseMagic || h[1] != nwords || h[3] != nblocks || h[2] < sparse_chunks_off(nblocks) || (h[2] & 15) ||
        h[2] > sparse_chunks_off(nblocks) + (size_t)nblocks * kSparseMaxChunk) {
        set_error("bad sparse header");
        return CFB_ERROR_BADFORMAT;
    }
    const size_t total = h[2];
    uint16_t pos[kSparseBlockWords + kSparseGroupWords];
    for (size_t b = w0 / kSparseBlockWords; b * kSparseBlockWords < w1 && b < nblocks; b++) {
        SparseChunk c;
        if (!sparse_chunk_open(sparse, total, (unsigned)b, &c)) { set_error("sparse block %u out of bounds", (unsigned)b); return CFB_ERROR_BADFORMAT; }
        if (!c.groups) continue;
        if (!chunk_positions(c, pos)) { set_error("sparse block %u: bitmaps and counts disagree", (unsigned)b); return CFB_ERROR_BADFORMAT; }
        const size_t base = b * kSparseBlockWords;
        // the part of the block inside [w0, w1)
        unsigned i0 = 0, i1 = c.values;
        if (w0 > base) { const unsigned lo = (unsigned)(w0 - base); while (i0 < i1 && pos[i0] < lo) i0++; }
        if (w1 < base + kSparseBlockWords) { const unsigned hi = (unsigned)(w1 - base); while (i1 > i0 && pos[i1 - 1] >= hi) i1--; }
        unsigned ei = 0, ne = 0;
        for (unsigned k = 0; k < i0; k++) ei += (c.bytes[k] == -128);
        for (unsigned k = i0; k < i1; k++) ne += (c.bytes[k] == -128);
        if (ei + ne > c.escapes) { set_error("sparse block %u: escape overrun", (unsigned)b); return CFB_ERROR_BADFORMAT; }
        if (i1 > i0) f(base, pos + i0, c.bytes + i0, c.wide + ei, i1 - i0);
    }
    return CFB_OK;
}

// f(pos, value) for every non-zero word of [w0, w1)
template <class F>
cfb_error walk(const cfb_layout *L, const void *sparse, size_t w0, size_t w1, F &&f)
{
    return walk_blocks(L, sparse, w0, w1, [&](size_t base, const uint16_t *pos, const signed char *vb, const int16_t *wide, unsigned n) {
        for (unsigned i = 0; i < n; i++) {
            int v = vb[i];
            if (__builtin_expect(v == -128, 0)) v = *wide++;
            f(base + pos[i], v);
        }
    });
}

}  // namespace

extern "C" {

cfb_error cfb_sparse_vlc_band(const cfb_layout *L, const void *sparse, int channel, int level, int band,
                              const cfb_vlc_codebook *book, cfb_bitwriter *bw)
{
    size_t w0, w1;
    cfb_error e = band_range(L, channel, level, band, &w0, &w1, nullptr);
    if (e) return e;
    if (!book_ok(book)) { set_error("bad code book"); return CFB_ERROR_INVALID_ARGUMENT; }
    if (!writer_ok(bw)) { set_error("bad bit writer"); return CFB_ERROR_INVALID_ARGUMENT; }
    Bits bits(*bw);
    const FastPairs &fast = fast_pairs(*book);
    size_t next = w0;               // first word not yet accounted for
    e = walk_blocks(L, sparse, w0, w1, [&](size_t base, const uint16_t *pos, const signed char *vb, const int16_t *wide, unsigned n) {
        Bits local = bits;          // bit buffer and run state in registers for the loop
        Coder coder(*book, local, fast);
        size_t nx = next;
        for (unsigned i = 0; i < n; i++) {
            int v = vb[i];
            if (__builtin_expect(v == -128, 0)) v = *wide++;
            const size_t p = base + pos[i];
            coder.token(p - nx, v);
            nx = p + 1;
        }
        next = nx;
        bits = local;
    });
    if (e) return e;
    Coder coder(*book, bits, fast);
    coder.run(w1 - next);           // pending run, incl. the last row's pitch gap (encoder.c:5671)
    if (bits.overflow) { set_error("bit writer out of space"); return CFB_ERROR_OUTOFMEMORY; }
    bits.store(bw);
    return CFB_OK;
}

cfb_error cfb_dense_vlc_band(const int16_t *image, int width, int height, int pitch_bytes, const cfb_vlc_codebook *book, cfb_bitwriter *bw)
{
    if (!image || width <= 0 || height <= 0 || pitch_bytes < 2 * width || (pitch_bytes & 1)) { set_error("bad band geometry"); return CFB_ERROR_INVALID_ARGUMENT; }
    if (!book_ok(book)) { set_error("bad code book"); return CFB_ERROR_INVALID_ARGUMENT; }
    if (!writer_ok(bw)) { set_error("bad bit writer"); return CFB_ERROR_INVALID_ARGUMENT; }
    Bits bits(*bw);
    Coder coder(*book, bits);
    const int pitch = pitch_bytes / 2, gap = pitch - width;
    uint64_t count = 0;
    for (int r = 0; r < height; r++) {
        const int16_t *row = image + (size_t)r * pitch;
        for (int x = 0; x < width; x++) {
            if (row[x] == 0) { count++; continue; }
            coder.token(count, row[x]);
            count = 0;
        }
        count += (uint64_t)gap;
    }
    coder.run(count);
    if (bits.overflow) { set_error("bit writer out of space"); return CFB_ERROR_OUTOFMEMORY; }
    bits.store(bw);
    return CFB_OK;
}

cfb_error cfb_sparse_band_nonzeros(const cfb_layout *L, const void *sparse, int channel, int level, int band, uint32_t *count)
{
    size_t w0, w1;
    cfb_error e = band_range(L, channel, level, band, &w0, &w1, nullptr);
    if (e) return e;
    if (!count) return CFB_ERROR_INVALID_ARGUMENT;
    uint32_t n = 0;
    e = walk(L, sparse, w0, w1, [&](size_t, int) { n++; });
    *count = n;
    return e;
}

cfb_error cfb_sparse_expand_band(const cfb_layout *L, const void *sparse, int channel, int level, int band, int16_t *out, int pitch_bytes)
{
    size_t w0, w1;
    const cfb_band_layout *bl = nullptr;
    cfb_error e = band_range(L, channel, level, band, &w0, &w1, &bl);
    if (e) return e;
    if (!out || pitch_bytes < 2 * bl->width || (pitch_bytes & 1)) { set_error("bad output pitch"); return CFB_ERROR_INVALID_ARGUMENT; }
    for (int r = 0; r < bl->height; r++) memset((unsigned char *)out + (size_t)r * pitch_bytes, 0, (size_t)bl->width * 2);
    const size_t pitch = (size_t)bl->pitch / 2;
    return walk(L, sparse, w0, w1, [&](size_t pos, int v) {
        const size_t r = (pos - w0) / pitch, x = (pos - w0) % pitch;
        if (x < (size_t)bl->width) *(int16_t *)((unsigned char *)out + r * (size_t)pitch_bytes + 2 * x) = (int16_t)v;
    });
}

// ---------------------------------------------------------------------------------------------------------------
// decoder side: tokens -> sparse

struct cfb_sparse_writer {
    cfb_layout layout;
    unsigned nwords = 0, nblocks = 0;
    unsigned char *out = nullptr;
    size_t capacity = 0, off = 0;
    size_t pos = 0;                 // next word of the flat coded region
    size_t band_end = 0;            // end of the band being written
    unsigned cur_block = 0;
    bool open = false;
    // the block being written, as the pieces of its chunk: tokens arrive in raster order, so every piece only grows
    // (round 2's first version kept a dense 16 KB scratch block and compacted it: a memset and a scan per block)
    int cur_group = -1;             // group of the mask being built, -1 = none
    unsigned cur_mask = 0;
    unsigned G = 0, V = 0, E = 0;
    unsigned char l1[kSparseL1Bytes];
    unsigned masks[kSparseBlockGroups];
    signed char vb[kSparseBlockWords];
    int16_t wide[kSparseBlockWords];
};

namespace {

inline void writer_close_group(cfb_sparse_writer *w)
{
    if (w->cur_group < 0) return;
    w->l1[w->cur_group >> 3] |= (unsigned char)(1u << (w->cur_group & 7));
    w->masks[w->G++] = w->cur_mask;
    w->cur_group = -1;
}

// close every block before `block`: the current one from its pieces, the ones in between are empty
cfb_error writer_flush_until(cfb_sparse_writer *w, unsigned block)
{
    unsigned *tab = (unsigned *)(w->out + kSparseHeaderBytes);
    while (w->cur_block < block && w->cur_block < w->nblocks) {
        const unsigned b = w->cur_block;
        unsigned cb = 0;
        writer_close_group(w);
        if (w->G) {
            if (w->off + kSparseMaxChunk > w->capacity) { set_error("sparse writer: output buffer too small"); return CFB_ERROR_OUTOFMEMORY; }
            cb = sparse_emit_chunk(w->out + w->off, w->l1, w->masks, w->G, w->vb, w->V, w->wide, w->E);
        }
        tab[4 * b] = (unsigned)w->off; tab[4 * b + 1] = w->G; tab[4 * b + 2] = w->V; tab[4 * b + 3] = w->E;
        if (w->G) { memset(w->l1, 0, sizeof(w->l1)); w->G = w->V = w->E = 0; }
        w->off += cb;
        w->cur_block++;
    }
    return CFB_OK;
}

inline cfb_error writer_advance(cfb_sparse_writer *w, size_t n)
{
    if (__builtin_expect(w->pos + n > w->band_end, 0)) { set_error("sparse writer: tokens run past the end of the band"); return CFB_ERROR_BADFORMAT; }
    w->pos += n;
    const unsigned blk = (unsigned)(w->pos / kSparseBlockWords);
    return __builtin_expect(blk > w->cur_block, 0) ? writer_flush_until(w, blk) : CFB_OK;
}

inline cfb_error writer_value(cfb_sparse_writer *w, int value)
{
    if (__builtin_expect(w->pos >= w->band_end, 0)) { set_error("sparse writer: tokens run past the end of the band"); return CFB_ERROR_BADFORMAT; }
    const int v = (int16_t)value;
    if (v) {
        const unsigned o = (unsigned)(w->pos % kSparseBlockWords);
        const int g = (int)(o / kSparseGroupWords);
        if (g != w->cur_group) { writer_close_group(w); w->cur_group = g; w->cur_mask = 0; }
        w->cur_mask |= 1u << (o % kSparseGroupWords);
        if (v < -127 || v > 127) { w->vb[w->V++] = -128; w->wide[w->E++] = (int16_t)v; } else w->vb[w->V++] = (signed char)v;
    }
    return writer_advance(w, 1);
}

}  // namespace

cfb_error cfb_sparse_writer_create(const cfb_layout *L, cfb_sparse_writer **out)
{
    if (!L || !out) { set_error("null argument"); return CFB_ERROR_INVALID_ARGUMENT; }
    cfb_sparse_writer *w = new (std::nothrow) cfb_sparse_writer();
    if (!w) return CFB_ERROR_OUTOFMEMORY;
    w->layout = *L;
    w->nwords = (unsigned)(L->coded_bytes / 2);
    w->nblocks = sparse_nblocks(w->nwords);
    *out = w;
    return CFB_OK;
}

void cfb_sparse_writer_destroy(cfb_sparse_writer *w) { delete w; }

cfb_error cfb_sparse_writer_begin(cfb_sparse_writer *w, void *sparse, size_t capacity)
{
    if (!w || !sparse) { set_error("null argument"); return CFB_ERROR_INVALID_ARGUMENT; }
    const size_t chunks = sparse_chunks_off(w->nblocks);
    if (capacity < chunks) { set_error("sparse writer: output buffer too small"); return CFB_ERROR_OUTOFMEMORY; }
    w->out = (unsigned char *)sparse; w->capacity = capacity; w->off = chunks;
    memset(w->out, 0, chunks);
    memset(w->l1, 0, sizeof(w->l1));
    w->G = w->V = w->E = 0; w->cur_group = -1; w->cur_mask = 0;
    w->pos = 0; w->band_end = 0; w->cur_block = 0; w->open = true;
    return CFB_OK;
}

cfb_error cfb_sparse_writer_band(cfb_sparse_writer *w, int channel, int level, int band)
{
    if (!w || !w->open) { set_error("sparse writer not begun"); return CFB_ERROR_INVALID_ARGUMENT; }
    size_t w0, w1;
    cfb_error e = band_range(&w->layout, channel, level, band, &w0, &w1, nullptr);
    if (e) return e;
    if (w0 < w->pos) { set_error("sparse writer: band (%d, %d, %d) is not the next one in the coded region", channel, level, band); return CFB_ERROR_INVALID_ARGUMENT; }
    w->band_end = w0;
    e = writer_advance(w, w0 - w->pos);         // whatever lies between two bands is zero
    w->band_end = w1;
    return e;
}

cfb_error cfb_sparse_writer_run(cfb_sparse_writer *w, uint32_t zeros) { return writer_advance(w, zeros); }

cfb_error cfb_sparse_writer_value(cfb_sparse_writer *w, int value) { return writer_value(w, value); }

cfb_error cfb_sparse_writer_dense_band(cfb_sparse_writer *w, int channel, int level, int band, const int16_t *rows, int pitch_bytes)
{
    cfb_error e = cfb_sparse_writer_band(w, channel, level, band);
    if (e) return e;
    const cfb_band_layout &bl = w->layout.band[channel][level][band];
    if (!rows || pitch_bytes < 2 * bl.width) { set_error("bad band rows"); return CFB_ERROR_INVALID_ARGUMENT; }
    for (int r = 0; r < bl.height && !e; r++) {
        const int16_t *row = (const int16_t *)((const unsigned char *)rows + (size_t)r * pitch_bytes);
        for (int x = 0; x < bl.width && !e; x++) e = writer_value(w, row[x]);
        if (!e) e = writer_advance(w, (size_t)(bl.pitch / 2 - bl.width));
    }
    return e;
}

cfb_error cfb_sparse_writer_end(cfb_sparse_writer *w, size_t *bytes)
{
    if (!w || !w->open) { set_error("sparse writer not begun"); return CFB_ERROR_INVALID_ARGUMENT; }
    w->band_end = w->nwords;
    cfb_error e = writer_flush_until(w, w->nblocks);
    if (e) return e;
    unsigned *h = (unsigned *)w->out;
    h[0] = kSparseMagic; h[1] = w->nwords; h[2] = (unsigned)w->off; h[3] = w->nblocks; h[4] = h[5] = h[6] = h[7] = 0;
    if (bytes) *bytes = w->off;
    w->open = false;
    return CFB_OK;
}

// ---- table-driven band parser -----------------------------------------------------------------------------------
struct cfb_vlc_decoder {
    static constexpr int kPrimaryBits = 12;
    struct Token { uint8_t kind; int32_t arg; };
    struct Node { int32_t child[2]; int32_t token; };          // token >= 0: leaf (index into tokens)
    // One look-up of the next kPrimaryBits bits decodes as many whole code words as fit: zero runs add up, a coefficient ends
    // the entry, the end-of-band code is only ever an entry of its own.  len 0: no code word ends inside the window ->
    // continue in the trie at `node` (node < 0: no code word starts with these bits).
    struct Fast { uint8_t len; uint8_t has_value; uint8_t end; uint32_t zeros; int32_t value; int32_t node; };
    // second level for code words longer than the window: per trie node reached after kPrimaryBits bits, a table over the
    // next kSecondBits bits (one code word per look-up); what is longer still walks the trie bit by bit
    static constexpr int kSecondBits = 8;
    struct Second { uint8_t len; int32_t token; int32_t node; };     // len 0: not resolved -> continue at `node` (< 0: no code)
    std::vector<Token> tokens;
    std::vector<Node> nodes;
    std::vector<Fast> fast;
    std::vector<int32_t> second_of_node;        // node -> first entry of its table in `second`, -1 = none
    std::vector<Second> second;
};

cfb_error cfb_vlc_decoder_create(const cfb_vlc_decodebook *book, cfb_vlc_decoder **out)
{
    if (!book || !out || book->count < 2 || !book->bits || !book->size || !book->kind || !book->arg) { set_error("bad decode book"); return CFB_ERROR_INVALID_ARGUMENT; }
    cfb_vlc_decoder *d = new (std::nothrow) cfb_vlc_decoder();
    if (!d) return CFB_ERROR_OUTOFMEMORY;
    d->nodes.push_back({{-1, -1}, -1});
    bool has_end = false;
    for (int i = 0; i < book->count; i++) {
        const int n = book->size[i];
        if (n < 1 || n > 31 || book->kind[i] > 2 || (book->kind[i] == 1 && book->arg[i] < 1)) { delete d; set_error("decode book entry %d malformed", i); return CFB_ERROR_INVALID_ARGUMENT; }
        has_end = has_end || book->kind[i] == 2;
        int node = 0;
        for (int k = n - 1; k >= 0; k--) {
            if (d->nodes[node].token >= 0) { delete d; set_error("decode book is not prefix free (entry %d)", i); return CFB_ERROR_INVALID_ARGUMENT; }
            const int bit = (book->bits[i] >> k) & 1;
            if (d->nodes[node].child[bit] < 0) { d->nodes[node].child[bit] = (int32_t)d->nodes.size(); d->nodes.push_back({{-1, -1}, -1}); }
            node = d->nodes[node].child[bit];
        }
        if (d->nodes[node].token >= 0 || d->nodes[node].child[0] >= 0 || d->nodes[node].child[1] >= 0) {
            delete d; set_error("decode book is not prefix free (entry %d)", i); return CFB_ERROR_INVALID_ARGUMENT;
        }
        d->nodes[node].token = (int32_t)d->tokens.size();
        d->tokens.push_back({book->kind[i], book->arg[i]});
    }
    if (!has_end) { delete d; set_error("decode book has no end-of-band code"); return CFB_ERROR_INVALID_ARGUMENT; }
    const int P = cfb_vlc_decoder::kPrimaryBits;
    d->fast.resize((size_t)1 << P);
    for (unsigned p = 0; p < (1u << P); p++) {
        cfb_vlc_decoder::Fast f = {0, 0, 0, 0u, 0, -1};
        int pos = 0;
        for (;;) {
            int node = 0, len = pos;
            while (len < P && node >= 0 && d->nodes[node].token < 0) { node = d->nodes[node].child[(p >> (P - 1 - len)) & 1]; len++; }
            if (node < 0 || d->nodes[node].token < 0) {         // no (whole) code word from `pos` inside the window
                if (pos == 0) f.node = node;
                break;
            }
            const cfb_vlc_decoder::Token &t = d->tokens[d->nodes[node].token];
            if (t.kind == 2) { if (pos == 0) { f.end = 1; pos = len; } break; }
            if (t.kind == 1) {
                if ((uint64_t)f.zeros + (uint32_t)t.arg > 0x7fffffffu) break;
                f.zeros += (uint32_t)t.arg; pos = len;
                continue;
            }
            f.has_value = 1; f.value = t.arg; pos = len;
            break;
        }
        f.len = (uint8_t)pos;
        d->fast[p] = f;
    }
    const int S = cfb_vlc_decoder::kSecondBits;
    d->second_of_node.assign(d->nodes.size(), -1);
    for (unsigned p = 0; p < (1u << P); p++) {
        const int start = d->fast[p].len ? -1 : d->fast[p].node;
        if (start < 0 || d->second_of_node[start] >= 0) continue;
        d->second_of_node[start] = (int32_t)d->second.size();
        for (unsigned q = 0; q < (1u << S); q++) {
            int node = start, len = 0;
            while (len < S && node >= 0 && d->nodes[node].token < 0) { node = d->nodes[node].child[(q >> (S - 1 - len)) & 1]; len++; }
            if (node >= 0 && d->nodes[node].token >= 0) d->second.push_back({(uint8_t)len, d->nodes[node].token, -1});
            else d->second.push_back({0, -1, node});
        }
    }
    *out = d;
    return CFB_OK;
}

void cfb_vlc_decoder_destroy(cfb_vlc_decoder *d) { delete d; }

cfb_error cfb_vlc_decode_band(const cfb_vlc_decoder *d, cfb_sparse_writer *w, int channel, int level, int band,
                              const uint8_t *stream, size_t stream_bytes, int quant, size_t *consumed)
{
    if (!d || !w || !stream) { set_error("null argument"); return CFB_ERROR_INVALID_ARGUMENT; }
    cfb_error e = cfb_sparse_writer_band(w, channel, level, band);
    if (e) return e;
    const int P = cfb_vlc_decoder::kPrimaryBits;
    uint64_t acc = 0;           // the next `have` bits of the stream, left-aligned at bit 63
    int have = 0;
    size_t rd = 0, bitpos = 0;  // bytes fetched, bits consumed
    auto fill = [&]() {
        if (have <= 32 && rd + 4 <= stream_bytes) {         // one big-endian word at a time while the stream lasts
            uint32_t be;
            memcpy(&be, stream + rd, 4);
            acc |= (uint64_t)__builtin_bswap32(be) << (32 - have);
            have += 32; rd += 4;
        }
        while (have <= 56 && rd < stream_bytes && rd + 4 > stream_bytes) { acc |= (uint64_t)stream[rd++] << (56 - have); have += 8; }
    };
    for (;;) {
        fill();
        if (have <= 0) { set_error("band stream ends without an end-of-band code"); return CFB_ERROR_BADFORMAT; }
        const cfb_vlc_decoder::Fast &f = d->fast[(size_t)(acc >> (64 - P))];
        if (__builtin_expect(f.len != 0, 1)) {
            if (f.len > have) { set_error("band stream truncated inside a code word"); return CFB_ERROR_BADFORMAT; }
            acc <<= f.len; have -= f.len; bitpos += (size_t)f.len;
            if (f.end) break;
            if (f.zeros) { e = writer_advance(w, f.zeros); if (e) return e; }
            if (f.has_value) { e = writer_value(w, (int)(int16_t)(f.value * quant)); if (e) return e; }
            continue;
        }
        // a code word longer than the window: second-level table, then the trie
        int node = f.node, len = P;
        if (node >= 0) {
            const int S = cfb_vlc_decoder::kSecondBits;
            const cfb_vlc_decoder::Second &s2 = d->second[(size_t)d->second_of_node[node] + (size_t)((acc >> (64 - P - S)) & ((1u << S) - 1))];
            if (s2.len) { node = -2; len = P + s2.len; }
            else { node = s2.node; len = P + S; }
            if (node == -2) {
                if (len > have) { set_error("band stream truncated inside a code word"); return CFB_ERROR_BADFORMAT; }
                acc <<= len; have -= len; bitpos += (size_t)len;
                const cfb_vlc_decoder::Token &t2 = d->tokens[s2.token];
                if (t2.kind == 2) break;
                e = (t2.kind == 1) ? writer_advance(w, (uint32_t)t2.arg) : writer_value(w, (int)(int16_t)(t2.arg * quant));
                if (e) return e;
                continue;
            }
        }
        while (node >= 0 && d->nodes[node].token < 0 && len < 32) { node = d->nodes[node].child[(acc >> (63 - len)) & 1]; len++; }
        if (node < 0 || d->nodes[node].token < 0) { set_error("band stream: no code word matches at bit %zu", bitpos); return CFB_ERROR_BADFORMAT; }
        if (len > have) { set_error("band stream truncated inside a code word"); return CFB_ERROR_BADFORMAT; }
        acc <<= len; have -= len; bitpos += (size_t)len;
        const cfb_vlc_decoder::Token &t = d->tokens[d->nodes[node].token];
        if (t.kind == 2) break;
        e = (t.kind == 1) ? writer_advance(w, (uint32_t)t.arg) : writer_value(w, (int)(int16_t)(t.arg * quant));
        if (e) return e;
    }
    if (consumed) *consumed = (bitpos + 7) / 8;
    return CFB_OK;
}

}  // extern "C"
