// cfb_sparse.cu -- lossless sparse transfer format for the coded region.
//
// After quantisation ~90 % of the highpass coefficients are zero, and the dense int16 bands (33 MB per
// 4K 4:2:2 frame) are what limits the host<->device path (PCIe), not the kernels.  The host entropy coder
// only ever needs (zero run, value) sequences (Codec/encoder.c:5386-5847 EncodeQuantLongRuns walks the band
// counting zeros, incl. the pitch gap :5811), so the natural wire format is
//
//     header  : u32 magic 'CFSP', u32 nwords, u32 nvalues, u32 reserved
//     bitmap  : nwords bits, bit i set <=> int16 word i of the coded region is non-zero (LSB-first in u32s)
//     values  : the nvalues non-zero int16 words in raster (word index) order
//
// over the flat coded region [0, coded_bytes) exactly as laid out by cfb_layout (pitch padding included, it is
// zero).  Compaction and expansion run on the GPU (three small kernels each: count per CTA of 8192 words,
// per-frame exclusive scan of the ~2000 CTA counts, scatter / gather with the position inside the CTA recomputed); the host helpers cfb_sparse_expand / cfb_sparse_compact are
// pure format conversions for callers that want dense bands.
#include "cfb_host.h"

namespace cfb {

constexpr int kSeg = 256;           // words per segment (one warp, 8 words per lane)
constexpr int kBlockSegs = 32;      // segments per CTA (32 warps): the unit of the cross-CTA prefix sum

struct SparseParams {
    int nframes;
    unsigned nwords;                // int16 words in the coded region
    unsigned nseg;
    unsigned nblocks;               // ceil(nseg / kBlockSegs)
    unsigned bitmap_off, values_off;            // byte offsets inside a sparse buffer
    const unsigned char *dense[kMaxBatch];      // pyramids (coded region at offset 0)
    unsigned char *sparse[kMaxBatch];
    unsigned *counts[kMaxBatch];                // nblocks + 1 entries: per-CTA counts, then exclusive offsets after the scan
};

__device__ __forceinline__ unsigned nonzero_mask8(const uint4 &w) {
    unsigned m = 0;
    m |= (w.x & 0xffffu) ? 1u : 0u;   m |= (w.x >> 16) ? 2u : 0u;
    m |= (w.y & 0xffffu) ? 4u : 0u;   m |= (w.y >> 16) ? 8u : 0u;
    m |= (w.z & 0xffffu) ? 16u : 0u;  m |= (w.z >> 16) ? 32u : 0u;
    m |= (w.w & 0xffffu) ? 64u : 0u;  m |= (w.w >> 16) ? 128u : 0u;
    return m;
}

__device__ __forceinline__ unsigned warp_incl_scan(unsigned v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned t = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += t; }
    return v;
}

// exclusive offset of this warp's segment inside its CTA + the CTA total (all 32 warps call it; one barrier)
__device__ __forceinline__ unsigned block_exclusive(unsigned warp_total, int lane, int wid, unsigned *smem32, unsigned *cta_total) {
    if (lane == 0) smem32[wid] = warp_total;
    __syncthreads();
    const unsigned mine = smem32[lane];
    const unsigned incl = warp_incl_scan(mine, lane);
    if (cta_total) *cta_total = __shfl_sync(0xffffffffu, incl, 31);
    return __shfl_sync(0xffffffffu, incl - mine, wid);
}

// A: per segment the bitmap words, per CTA (32 segments) the number of non-zero words
__global__ void __launch_bounds__(1024) k_sparse_count(const __grid_constant__ SparseParams p)
{
    __shared__ unsigned wtot[32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const unsigned seg = blockIdx.x * kBlockSegs + wid;
    const int f = blockIdx.y;
    const unsigned w0 = seg * kSeg + lane * 8;
    uint4 w = make_uint4(0, 0, 0, 0);
    if (seg < p.nseg && w0 < p.nwords) w = __ldg(reinterpret_cast<const uint4 *>(p.dense[f] + (size_t)w0 * 2));
    const unsigned m8 = nonzero_mask8(w);
    unsigned m = m8 << ((lane & 3) * 8);
    m |= __shfl_xor_sync(0xffffffffu, m, 1);
    m |= __shfl_xor_sync(0xffffffffu, m, 2);
    if ((lane & 3) == 0 && seg < p.nseg && w0 < p.nwords)
        reinterpret_cast<unsigned *>(p.sparse[f] + p.bitmap_off)[seg * 8 + (lane >> 2)] = m;
    unsigned c = __popc(m8);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    unsigned total;
    block_exclusive(c, lane, wid, wtot, &total);
    if (threadIdx.x == 0) p.counts[f][blockIdx.x] = total;
}

// A': per-CTA counts from an uploaded bitmap (32 segments = 256 bitmap words per CTA)
__global__ void __launch_bounds__(256) k_sparse_count_bitmap(const __grid_constant__ SparseParams p)
{
    __shared__ unsigned wtot[8];
    const int f = blockIdx.y;
    const unsigned word = blockIdx.x * (kBlockSegs * 8) + threadIdx.x;          // index of a 32-bit bitmap word
    unsigned c = 0;
    if (word < p.nwords / 32) c = __popc(__ldg(reinterpret_cast<const unsigned *>(p.sparse[f] + p.bitmap_off) + word));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) wtot[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned t = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) t += wtot[i];
        p.counts[f][blockIdx.x] = t;
    }
}

// B: per-frame exclusive scan of the per-CTA counts (one CTA per frame; a 4K 4:2:2 frame has 2026 of them);
// writes the total into the header
__global__ void __launch_bounds__(1024) k_sparse_scan(const __grid_constant__ SparseParams p, int write_header)
{
    __shared__ unsigned wsum[32];
    const int f = blockIdx.x;
    unsigned *cnt = p.counts[f];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    unsigned carry = 0;
    for (unsigned base = 0; base < p.nblocks; base += 1024 * 4) {
        unsigned v[4], s = 0;
        const unsigned i0 = base + tid * 4;
#pragma unroll
        for (int k = 0; k < 4; k++) { v[k] = (i0 + k < p.nblocks) ? cnt[i0 + k] : 0u; s += v[k]; }
        const unsigned incl = warp_incl_scan(s, lane);
        unsigned chunk_total;
        const unsigned wexcl = block_exclusive(__shfl_sync(0xffffffffu, incl, 31), lane, wid, wsum, &chunk_total);
        unsigned excl = carry + wexcl + (incl - s);
#pragma unroll
        for (int k = 0; k < 4; k++) { if (i0 + k < p.nblocks) cnt[i0 + k] = excl; excl += v[k]; }
        carry += chunk_total;
        __syncthreads();
    }
    if (tid == 0) {
        cnt[p.nblocks] = carry;
        if (write_header) {
            unsigned *h = reinterpret_cast<unsigned *>(p.sparse[f]);
            h[0] = 0x50534643u; h[1] = p.nwords; h[2] = carry; h[3] = 0;
        }
    }
}

// C: scatter the non-zero words in raster order (the position inside the CTA is recomputed from the data)
__global__ void __launch_bounds__(1024) k_sparse_scatter(const __grid_constant__ SparseParams p)
{
    __shared__ unsigned wtot[32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const unsigned seg = blockIdx.x * kBlockSegs + wid;
    const int f = blockIdx.y;
    const unsigned w0 = seg * kSeg + lane * 8;
    uint4 w = make_uint4(0, 0, 0, 0);
    if (seg < p.nseg && w0 < p.nwords) w = __ldg(reinterpret_cast<const uint4 *>(p.dense[f] + (size_t)w0 * 2));
    const unsigned m8 = nonzero_mask8(w);
    const unsigned c = __popc(m8);
    const unsigned incl = warp_incl_scan(c, lane);
    const unsigned wexcl = block_exclusive(__shfl_sync(0xffffffffu, incl, 31), lane, wid, wtot, nullptr);
    unsigned pos = p.counts[f][blockIdx.x] + wexcl + incl - c;
    unsigned short *vals = reinterpret_cast<unsigned short *>(p.sparse[f] + p.values_off);
    const unsigned ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const unsigned short x = (unsigned short)((k & 1) ? (ws[k >> 1] >> 16) : (ws[k >> 1] & 0xffffu));
        if (x) vals[pos++] = x;
    }
}

// C': gather back into the dense coded region
__global__ void __launch_bounds__(1024) k_sparse_gather(const __grid_constant__ SparseParams p)
{
    __shared__ unsigned wtot[32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const unsigned seg = blockIdx.x * kBlockSegs + wid;
    const int f = blockIdx.y;
    const unsigned w0 = seg * kSeg + lane * 8;
    const bool on = (seg < p.nseg) && (w0 < p.nwords);          // nwords is a multiple of 8 (bands are 64-byte aligned)
    unsigned m8 = 0;
    if (on) {
        const unsigned bw = __ldg(reinterpret_cast<const unsigned *>(p.sparse[f] + p.bitmap_off) + seg * 8 + (lane >> 2));
        m8 = (bw >> ((lane & 3) * 8)) & 0xffu;
    }
    const unsigned c = __popc(m8);
    const unsigned incl = warp_incl_scan(c, lane);
    const unsigned wexcl = block_exclusive(__shfl_sync(0xffffffffu, incl, 31), lane, wid, wtot, nullptr);
    if (!on) return;
    unsigned pos = p.counts[f][blockIdx.x] + wexcl + incl - c;
    const unsigned short *vals = reinterpret_cast<const unsigned short *>(p.sparse[f] + p.values_off);
    unsigned out[4] = {0, 0, 0, 0};
#pragma unroll
    for (int k = 0; k < 8; k++) {
        if (m8 & (1u << k)) {
            const unsigned x = vals[pos++];
            out[k >> 1] |= (k & 1) ? (x << 16) : x;
        }
    }
    *reinterpret_cast<uint4 *>(const_cast<unsigned char *>(p.dense[f]) + (size_t)w0 * 2) = make_uint4(out[0], out[1], out[2], out[3]);
}

cudaError_t launch_sparse_compact(const SparseParams &p, cudaStream_t stream)
{
    dim3 grid(p.nblocks, p.nframes);
    k_sparse_count<<<grid, 1024, 0, stream>>>(p);
    k_sparse_scan<<<p.nframes, 1024, 0, stream>>>(p, 1);
    k_sparse_scatter<<<grid, 1024, 0, stream>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_sparse_expand(const SparseParams &p, cudaStream_t stream)
{
    dim3 grid(p.nblocks, p.nframes);
    k_sparse_count_bitmap<<<grid, 256, 0, stream>>>(p);
    k_sparse_scan<<<p.nframes, 1024, 0, stream>>>(p, 0);
    k_sparse_gather<<<grid, 1024, 0, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace cfb

// ---------------------------------------------------------------------------
// C ABI
using namespace cfb;

static inline unsigned sp_bitmap_off() { return 16u; }
static inline unsigned sp_values_off(unsigned nwords) { return (16u + nwords / 8u + 15u) & ~15u; }

static cfb_error sparse_prepare(cfb_codec *cd, SparseParams &p, int n)
{
    const cfb_layout &L = cd->layout;
    p.nframes = n;
    p.nwords = (unsigned)(L.coded_bytes / 2);
    p.nseg = (p.nwords + kSeg - 1) / kSeg;
    p.nblocks = (p.nseg + kBlockSegs - 1) / kBlockSegs;
    p.bitmap_off = sp_bitmap_off();
    p.values_off = sp_values_off(p.nwords);
    // each staging buffer under its own check: a failed allocation leaves the others usable for the retry
    cd->sparse_stride = (cfb_sparse_max_bytes(&L) + 255) & ~(size_t)255;
    if (!cd->d_sparse) CFB_CUDA(cudaMalloc((void **)&cd->d_sparse, cd->sparse_stride * cd->max_batch));
    if (!cd->d_counts) CFB_CUDA(cudaMalloc((void **)&cd->d_counts, sizeof(unsigned) * (size_t)(p.nblocks + 1) * cd->max_batch));
    if (!cd->h_headers) CFB_CUDA(cudaHostAlloc((void **)&cd->h_headers, 16 * (size_t)cd->max_batch, cudaHostAllocPortable));
    for (int i = 0; i < n; i++) {
        p.dense[i] = cd->d_pyramids + cd->pyramid_stride * i;
        p.sparse[i] = cd->d_sparse + cd->sparse_stride * i;
        p.counts[i] = cd->d_counts + (size_t)(p.nblocks + 1) * i;
    }
    return CFB_OK;
}

namespace cfb {

unsigned sparse_initial_guess(const cfb_codec *cd) { return (unsigned)(cd->layout.coded_bytes / 2) / 8; }

unsigned sparse_next_guess(const cfb_codec *cd, unsigned max_values)
{
    const unsigned nwords = (unsigned)(cd->layout.coded_bytes / 2);
    unsigned g = (max_values + max_values / 8 + 4096 + 63u) & ~63u;
    if (g > nwords) g = nwords;           // clamp AFTER the rounding: the copy must stay inside cfb_sparse_max_bytes
    return g;
}

cfb_error sparse_compact_device(cfb_codec *cd, int n)
{
    SparseParams sp;
    cfb_error err = sparse_prepare(cd, sp, n);
    if (err) return err;
    CFB_CUDA(launch_sparse_compact(sp, cd->ctx->stream));
    cd->ctx->kernel_launches += 3;
    return CFB_OK;
}

cfb_error sparse_expand_device(cfb_codec *cd, int n)
{
    SparseParams sp;
    cfb_error err = sparse_prepare(cd, sp, n);
    if (err) return err;
    CFB_CUDA(launch_sparse_expand(sp, cd->ctx->stream));
    cd->ctx->kernel_launches += 3;
    return CFB_OK;
}

cfb_error sparse_download(cfb_codec *cd, int n, void *const *h_sparse, unsigned guess, cudaStream_t s)
{
    SparseParams sp;
    cfb_error err = sparse_prepare(cd, sp, n);
    if (err) return err;
    if (guess > sp.nwords) guess = sp.nwords;       // never more than the caller's buffer holds (cfb_sparse_max_bytes)
    for (int i = 0; i < n; i++) {
        if (!h_sparse[i]) { set_error("null host buffer %d", i); return CFB_ERROR_INVALID_ARGUMENT; }
        const size_t bytes = (size_t)sp.values_off + (size_t)guess * 2;
        CFB_CUDA(cudaMemcpyAsync(h_sparse[i], sp.sparse[i], bytes, cudaMemcpyDeviceToHost, s));
        cd->ctx->d2h_bytes += (uint64_t)bytes;
    }
    return CFB_OK;
}

cfb_error stage_fwd_tail(cfb_codec *cd, int n, void *const *h_sparse, unsigned guess, cudaStream_t s, size_t *sizes,
                         unsigned *max_values, bool *more)
{
    SparseParams sp;
    cfb_error err = sparse_prepare(cd, sp, n);
    if (err) return err;
    if (guess > sp.nwords) guess = sp.nwords;
    unsigned maxv = 0;
    *more = false;
    for (int i = 0; i < n; i++) {
        const unsigned nv = ((const unsigned *)h_sparse[i])[2];
        if (nv > sp.nwords) { set_error("sparse header %d corrupt", i); return CFB_ERROR_UNEXPECTED; }
        if (nv > maxv) maxv = nv;
        if (nv > guess) {
            const size_t off = (size_t)sp.values_off + (size_t)guess * 2, rest = (size_t)(nv - guess) * 2;
            CFB_CUDA(cudaMemcpyAsync((unsigned char *)h_sparse[i] + off, sp.sparse[i] + off, rest, cudaMemcpyDeviceToHost, s));
            cd->ctx->d2h_bytes += (uint64_t)rest;
            *more = true;
        }
        if (sizes) sizes[i] = (size_t)sp.values_off + (size_t)nv * 2;
    }
    if (max_values) *max_values = maxv;
    return CFB_OK;
}

cfb_error sparse_upload(cfb_codec *cd, int n, const void *const *h_sparse, cudaStream_t s)
{
    SparseParams sp;
    cfb_error err = sparse_prepare(cd, sp, n);
    if (err) return err;
    for (int i = 0; i < n; i++) {
        if (!h_sparse[i]) { set_error("null host buffer %d", i); return CFB_ERROR_INVALID_ARGUMENT; }
        const unsigned *h = (const unsigned *)h_sparse[i];
        if (h[0] != 0x50534643u || h[1] != sp.nwords || h[2] > sp.nwords) { set_error("sparse buffer %d: bad header", i); return CFB_ERROR_BADFORMAT; }
        const size_t bytes = (size_t)sp.values_off + (size_t)h[2] * 2;
        CFB_CUDA(cudaMemcpyAsync(sp.sparse[i], h_sparse[i], bytes, cudaMemcpyHostToDevice, s));
        cd->ctx->h2d_bytes += (uint64_t)bytes;
    }
    return CFB_OK;
}

}  // namespace cfb

extern "C" {

size_t cfb_sparse_max_bytes(const cfb_layout *L)
{
    if (!L) return 0;
    const unsigned nwords = (unsigned)(L->coded_bytes / 2);
    return (size_t)sp_values_off(nwords) + (size_t)nwords * 2;
}

size_t cfb_sparse_bytes(const void *sparse)
{
    if (!sparse) return 0;
    const unsigned *h = (const unsigned *)sparse;
    if (h[0] != 0x50534643u) return 0;
    return (size_t)sp_values_off(h[1]) + (size_t)h[2] * 2;
}

cfb_error cfb_forward_host_sparse(cfb_codec *cd, int n, const void *const *h_frames, int frame_pitch,
                                  const cfb_quant *quant, void *const *h_sparse, size_t *sparse_bytes)
{
    if (!cd || !h_frames || !quant || !h_sparse) { set_error("null argument"); return CFB_ERROR_INVALID_ARGUMENT; }
    if (n < 1 || n > cd->max_batch) { set_error("batch %d exceeds codec max_batch %d", n, cd->max_batch); return CFB_ERROR_INVALID_ARGUMENT; }
    cfb_context *ctx = cd->ctx;
    for (int i = 0; i < n; i++) if (!h_sparse[i]) { set_error("null host buffer %d", i); return CFB_ERROR_INVALID_ARGUMENT; }
    cfb_error err = stage_fwd_upload(cd, n, h_frames, frame_pitch, ctx->stream);
    if (!err) err = stage_fwd_compute(cd, n, quant, true);
    if (err) return err;
    // Speculative single-pass D2H: copy header + bitmap + as many values as recent frames needed (+12 %) right behind the
    // kernels, without a host round trip; only if a frame turns out to hold more values is the remainder fetched.
    const unsigned guess = cd->value_guess ? cd->value_guess : sparse_initial_guess(cd);
    err = stage_fwd_download(cd, n, h_sparse, true, guess, ctx->stream);
    if (err) return err;
    CFB_CUDA(stream_wait(ctx));
    unsigned maxv = 0;
    bool more = false;
    size_t sizes[kMaxBatch];
    err = stage_fwd_tail(cd, n, h_sparse, guess, ctx->stream, sizes, &maxv, &more);
    if (err) return err;
    if (more) CFB_CUDA(stream_wait(ctx));
    if (sparse_bytes) for (int i = 0; i < n; i++) sparse_bytes[i] = sizes[i];
    cd->value_guess = sparse_next_guess(cd, maxv);
    return CFB_OK;
}

cfb_error cfb_inverse_host_sparse(cfb_codec *cd, int n, const void *const *h_sparse, const cfb_quant *quant,
                                  int out_format, void *const *h_frames, int frame_pitch)
{
    if (!cd || !h_sparse || !quant || !h_frames) { set_error("null argument"); return CFB_ERROR_INVALID_ARGUMENT; }
    cfb_context *ctx = cd->ctx;
    cfb_error err = stage_inv_upload(cd, n, h_sparse, true, ctx->stream);
    if (!err) err = stage_inv_compute(cd, n, quant, out_format, true);
    if (!err) err = stage_inv_download(cd, n, h_frames, frame_pitch, out_format, ctx->stream);
    if (err) return err;
    CFB_CUDA(stream_wait(ctx));
    return CFB_OK;
}

cfb_error cfb_sparse_expand(const cfb_layout *L, const void *sparse, void *dense_coded)
{
    if (!L || !sparse || !dense_coded) return CFB_ERROR_INVALID_ARGUMENT;
    const unsigned *h = (const unsigned *)sparse;
    const unsigned nwords = (unsigned)(L->coded_bytes / 2);
    if (h[0] != 0x50534643u || h[1] != nwords) { set_error("bad sparse header"); return CFB_ERROR_BADFORMAT; }
    const unsigned *bm = (const unsigned *)((const unsigned char *)sparse + sp_bitmap_off());
    const int16_t *vals = (const int16_t *)((const unsigned char *)sparse + sp_values_off(nwords));
    int16_t *out = (int16_t *)dense_coded;
    size_t pos = 0;
    for (unsigned w = 0; w < nwords; w += 32) {
        unsigned m = bm[w >> 5];
        for (int k = 0; k < 32 && w + k < nwords; k++) {
            if ((m >> k) & 1u) {
                if (pos >= h[2]) { set_error("sparse bitmap has more set bits than the header's value count"); return CFB_ERROR_BADFORMAT; }
                out[w + k] = vals[pos++];
            } else out[w + k] = 0;
        }
    }
    if (pos != h[2]) { set_error("sparse value count mismatch"); return CFB_ERROR_BADFORMAT; }
    return CFB_OK;
}

cfb_error cfb_sparse_compact(const cfb_layout *L, const void *dense_coded, void *sparse, size_t *bytes)
{
    if (!L || !sparse || !dense_coded) return CFB_ERROR_INVALID_ARGUMENT;
    const unsigned nwords = (unsigned)(L->coded_bytes / 2);
    unsigned *h = (unsigned *)sparse;
    unsigned *bm = (unsigned *)((unsigned char *)sparse + sp_bitmap_off());
    int16_t *vals = (int16_t *)((unsigned char *)sparse + sp_values_off(nwords));
    const int16_t *in = (const int16_t *)dense_coded;
    size_t pos = 0;
    for (unsigned w = 0; w < nwords; w += 32) {
        unsigned m = 0;
        for (int k = 0; k < 32 && w + k < nwords; k++) if (in[w + k]) { m |= 1u << k; vals[pos++] = in[w + k]; }
        bm[w >> 5] = m;
    }
    h[0] = 0x50534643u; h[1] = nwords; h[2] = (unsigned)pos; h[3] = 0;
    if (bytes) *bytes = (size_t)sp_values_off(nwords) + pos * 2;
    return CFB_OK;
}

}  // extern "C"
