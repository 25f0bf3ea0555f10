// cfb_sparse_format.h -- constants and accessors of the 'CFS2' sparse transfer format (see cfb_sparse.cu for the
// description).  Shared by the device kernels, the host-side conversions and the host VLC walker (cfb_vlc.cu).
#pragma once
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#if defined(__SSE2__) && !defined(__CUDA_ARCH__)
#include <emmintrin.h>
#endif

#ifdef __CUDACC__
#define CFB_HD __host__ __device__ __forceinline__
#else
#define CFB_HD inline
#endif

namespace cfb {

constexpr unsigned kSparseMagic = 0x32534643u;      // 'CFS2'
constexpr unsigned kSparseHeaderBytes = 32;
constexpr unsigned kSparseTableEntry = 16;          // {u32 chunk offset, u32 groups, u32 values, u32 escapes}
constexpr unsigned kSparseBlockWords = 8192;        // int16 words per block (one CTA: 256 threads x 4 pieces x 8 words)
constexpr unsigned kSparseGroupWords = 32;
constexpr unsigned kSparseBlockGroups = kSparseBlockWords / kSparseGroupWords;      // 256
constexpr unsigned kSparseL1Bytes = kSparseBlockGroups / 8;                         // 32
// worst case: every word non-zero and outside [-127, 127]
constexpr unsigned kSparseMaxChunk = kSparseL1Bytes + 4 * kSparseBlockGroups + kSparseBlockWords + 2 * kSparseBlockWords;     // 25632

CFB_HD unsigned sparse_nblocks(unsigned nwords) { return (nwords + kSparseBlockWords - 1) / kSparseBlockWords; }
CFB_HD unsigned sparse_chunks_off(unsigned nblocks) { return (kSparseHeaderBytes + kSparseTableEntry * nblocks + 15u) & ~15u; }
// bytes of a block's chunk (a multiple of 16; 0 for an all-zero block)
CFB_HD unsigned sparse_chunk_bytes(unsigned groups, unsigned values, unsigned escapes)
{
    if (!groups) return 0;
    return (kSparseL1Bytes + 4 * groups + ((values + 3u) & ~3u) + ((2 * escapes + 3u) & ~3u) + 15u) & ~15u;
}

// host view of one block
struct SparseChunk {
    unsigned groups, values, escapes;
    const unsigned char *l1;        // 32 bytes (null when the block is empty)
    const unsigned *masks;
    const signed char *bytes;
    const int16_t *wide;
};

// bounds-checked against `total` (the buffer's size from its header); false = damaged table
inline bool sparse_chunk_open(const void *sparse, size_t total, unsigned block, SparseChunk *c)
{
    const unsigned char *base = (const unsigned char *)sparse;
    const unsigned *e = (const unsigned *)(base + kSparseHeaderBytes + (size_t)block * kSparseTableEntry);
    c->groups = e[1]; c->values = e[2]; c->escapes = e[3];
    c->l1 = nullptr; c->masks = nullptr; c->bytes = nullptr; c->wide = nullptr;
    if (c->groups > kSparseBlockGroups || c->values > kSparseBlockWords || c->escapes > c->values || (c->groups == 0) != (c->values == 0)) return false;
    if (!c->groups) return true;
    const size_t off = e[0];
    if ((off & 15) || off + sparse_chunk_bytes(c->groups, c->values, c->escapes) > total) return false;
    c->l1 = base + off;
    c->masks = (const unsigned *)(base + off + kSparseL1Bytes);
    c->bytes = (const signed char *)(base + off + kSparseL1Bytes + 4 * (size_t)c->groups);
    c->wide = (const int16_t *)(base + off + kSparseL1Bytes + 4 * (size_t)c->groups + ((c->values + 3u) & ~3u));
    return true;
}

// ---- host side: the one place that lays a chunk out (shared by the compaction below and the sparse writer) ----
// pieces of one block -> `chunk`; returns its size (0 for an empty block, nothing written)
inline unsigned sparse_emit_chunk(unsigned char *chunk, const unsigned char *l1, const unsigned *masks, unsigned G,
                                  const signed char *vb, unsigned V, const int16_t *wide, unsigned E)
{
    const unsigned cb = sparse_chunk_bytes(G, V, E);
    if (!G) return 0;
    memset(chunk, 0, cb);
    memcpy(chunk, l1, kSparseL1Bytes);
    memcpy(chunk + kSparseL1Bytes, masks, 4 * (size_t)G);
    memcpy(chunk + kSparseL1Bytes + 4 * (size_t)G, vb, V);
    memcpy(chunk + kSparseL1Bytes + 4 * (size_t)G + ((V + 3) & ~3u), wide, 2 * (size_t)E);
    return cb;
}

// ---- host side: compaction of one dense block (cfb_sparse_compact*) ----
// One block of 8192 words -> its chunk at `chunk`; returns the chunk size and the table entry's counts
inline unsigned sparse_compact_block(const int16_t *in, unsigned nvalid, unsigned char *chunk, unsigned *G_out, unsigned *V_out, unsigned *E_out)
{
    unsigned char l1[kSparseL1Bytes] = {0};
    unsigned masks[kSparseBlockGroups];
    signed char vb[kSparseBlockWords];
    int16_t wide[kSparseBlockWords];
    unsigned G = 0, V = 0, E = 0;
    for (unsigned g = 0; g * kSparseGroupWords < nvalid; g++) {
        const int16_t *grp = in + (size_t)g * kSparseGroupWords;
        const unsigned n = nvalid - g * kSparseGroupWords < kSparseGroupWords ? nvalid - g * kSparseGroupWords : kSparseGroupWords;
        unsigned m = 0;
#if defined(__SSE2__) && !defined(__CUDA_ARCH__)
        if (n == kSparseGroupWords) {           // most groups are empty: the non-zero mask of 32 words from four 16-byte compares
            const __m128i zero = _mm_setzero_si128();
            const __m128i a = _mm_loadu_si128((const __m128i *)grp), b = _mm_loadu_si128((const __m128i *)(grp + 8));
            const __m128i c = _mm_loadu_si128((const __m128i *)(grp + 16)), d = _mm_loadu_si128((const __m128i *)(grp + 24));
            const unsigned z0 = (unsigned)_mm_movemask_epi8(_mm_packs_epi16(_mm_cmpeq_epi16(a, zero), _mm_cmpeq_epi16(b, zero)));
            const unsigned z1 = (unsigned)_mm_movemask_epi8(_mm_packs_epi16(_mm_cmpeq_epi16(c, zero), _mm_cmpeq_epi16(d, zero)));
            m = ~(z0 | (z1 << 16));
            if (!m) continue;
            for (unsigned left = m; left; left &= left - 1) {
                const int v = grp[__builtin_ctz(left)];
                if (v < -127 || v > 127) { vb[V++] = -128; wide[E++] = (int16_t)v; } else vb[V++] = (signed char)v;
            }
            l1[g >> 3] |= (unsigned char)(1u << (g & 7)); masks[G++] = m;
            continue;
        }
#else
        if (n == kSparseGroupWords) {           // most groups are empty: eight 64-bit tests
            uint64_t any = 0, q[8];
            memcpy(q, grp, sizeof(q));
            for (int k = 0; k < 8; k++) any |= q[k];
            if (!any) continue;
        }
#endif
        for (unsigned k = 0; k < n; k++) {
            const int v = grp[k];
            if (!v) continue;
            m |= 1u << k;
            if (v < -127 || v > 127) { vb[V++] = -128; wide[E++] = (int16_t)v; } else vb[V++] = (signed char)v;
        }
        if (m) { l1[g >> 3] |= (unsigned char)(1u << (g & 7)); masks[G++] = m; }
    }
    *G_out = G; *V_out = V; *E_out = E;
    return sparse_emit_chunk(chunk, l1, masks, G, vb, V, wide, E);
}


}  // namespace cfb
