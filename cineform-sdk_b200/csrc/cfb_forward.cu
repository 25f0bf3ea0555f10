// cfb_forward.cu -- forward 2-6 wavelet level + fused quantisation, sm_100a.
//
// Replaces (reference, per level and channel):
//   Codec/spatial.c:10026 FilterSpatialQuant16s      -> k_fwd_plane<0>
//   Codec/spatial.c:12942 FilterSpatialV210Quant16s  -> k_fwd_plane<2>
//   Codec/spatial.c:14726 FilterSpatialYUVQuant16s   -> k_fwd_422 (+ Codec/convert.c:4667 unpack)
//   Codec/quantize.c:1395 QuantizeRow16sTo16s         -> fused into the band stores
//
// Design: no shared memory, no block barriers.  One WARP owns a
// strip of 128 output columns x TH output rows.  Each lane loads 16 bytes of an
// input row straight into registers (coalesced 128-bit loads), does the
// horizontal lifting for its 4 output columns exchanging one value with each
// neighbour lane by warp shuffle, and keeps only three values per column of
// vertical state in registers.  With S_j = row(2j)+row(2j+1) and D_j = row(2j)-row(2j+1)
// the vertical 2-6 filter is   low_j = S_j,  high_j = ((S_{j+1} - S_{j-1} + 4) >> 3) + D_j,
// the top/bottom 6-tap border filters are (-3 S0 + 8 D0 + 4 S1 - S2 + 4) >> 3 and
// (3 S_n + 8 D_n - 4 S_{n-1} + S_{n-2} + 4) >> 3 (same identities horizontally).
// Every input sample is read from global memory once (+ 1 halo row pair per strip
// block, served by L2) and every coefficient is written once with 64-bit stores.
//
// Work split inside one launch: "main" warps run a compact loop that emits every
// LL/LH row and the interior HL/HH rows; the two border rows of HL/HH (first and
// last output row, which use the 6-tap border filters) are produced by dedicated
// border warps in an extra CTA row, so the hot loop carries no border code.
#include "cfb_common.cuh"
#include "cfb_tma.cuh"
#include <cstdlib>
#include <cstring>
#include <type_traits>

namespace cfb {

// ----------------------------------------------------------------------------
struct LaneInfo {
    unsigned amask;         // lanes of this warp that hold image columns
    bool left_border;       // lane owns output column 0
    bool right_border;      // lane owns the last output column
    bool has_border;        // warp-uniform: this strip touches the left or right image border
    bool use_lh, use_rh;    // lane 0 / last lane need a halo word from the neighbouring strip
};

template <int NC> struct VState {
    int llp[2 * NC];    // S_{j-2}
    int llc[2 * NC];    // S_{j-1}
    int dc[2 * NC];     // D_{j-1}
};

template <int NC> struct VecStore;
template <> struct VecStore<4> {
    static __device__ __forceinline__ void st(unsigned char *p, unsigned a, unsigned b) {
        *reinterpret_cast<uint2 *>(p) = make_uint2(a, b);
    }
};
template <> struct VecStore<2> {
    static __device__ __forceinline__ void st(unsigned char *p, unsigned a, unsigned) {
        *reinterpret_cast<unsigned *>(p) = a;
    }
};
template <int NC>
__device__ __forceinline__ void store_raw(unsigned char *p, const int *v) {
    if (NC == 4) VecStore<4>::st(p, pack_lo(v[0], v[1]), pack_lo(v[2], v[3]));
    else VecStore<2>::st(p, pack_lo(v[0], v[1]), 0u);
}
template <int NC>
__device__ __forceinline__ void store_quant(unsigned char *p, const int *v, const QuantParam &q) {
    if (NC == 4)
        VecStore<4>::st(p, pack_hi(quant1(v[0], q), quant1(v[1], q)), pack_hi(quant1(v[2], q), quant1(v[3], q)));
    else
        VecStore<2>::st(p, pack_hi(quant1(v[0], q), quant1(v[1], q)), 0u);
}

// predicated 64/32-bit global stores: the address and the value are computed unconditionally and only the store is
// guarded, so the hot loop carries no divergence-safe branch (BSSY/BSYNC/BRA) around its band stores
__device__ __forceinline__ void st_pred(unsigned char *p, unsigned a, unsigned b, bool on) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %3, 0;\n\t@p st.global.v2.u32 [%0], {%1, %2};\n\t}"
                 :: "l"(p), "r"(a), "r"(b), "r"((int)on) : "memory");
}
__device__ __forceinline__ void st_pred(unsigned char *p, unsigned a, bool on) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %2, 0;\n\t@p st.global.u32 [%0], %1;\n\t}"
                 :: "l"(p), "r"(a), "r"((int)on) : "memory");
}
template <int NC>
__device__ __forceinline__ void store_raw_if(unsigned char *p, const int *v, bool on) {
    if (NC == 4) st_pred(p, pack_lo(v[0], v[1]), pack_lo(v[2], v[3]), on);
    else st_pred(p, pack_lo(v[0], v[1]), on);
}
template <int NC>
__device__ __forceinline__ void store_quant_if(unsigned char *p, const int *v, const QuantParam &q, bool on) {
    if (NC == 4)
        st_pred(p, pack_hi(quant1(v[0], q), quant1(v[1], q)), pack_hi(quant1(v[2], q), quant1(v[3], q)), on);
    else
        st_pred(p, pack_hi(quant1(v[0], q), quant1(v[1], q)), on);
}

// One vertical step on the horizontal outputs of rows 2j (a) and 2j+1 (b); [0,NC) = low, [NC,2NC) = high.
// emit_low : store LL/LH of output row j        at byte offset off
// emit_high: store HL/HH of output row j-1      at byte offset off - pitch   (interior formula)
// Both flags are warp-uniform.  QLL: 0 = LL is never quantised (the 4:2:2 level-1 filter, spatial.c:14726),
// 1 = decided at run time by g.quant_ll (planar filter with an LL divisor > 1, spatial.c:10480).
template <int NC, int QLL>
__device__ __forceinline__ void vstep(VState<NC> &s, const int *a, const int *b, const PlaneGeom &g, unsigned char *out,
                                      unsigned off, bool emit_low, bool emit_high)
{
    int v[2 * NC], dn[2 * NC];
#pragma unroll
    for (int i = 0; i < 2 * NC; i++) { v[i] = a[i] + b[i]; dn[i] = a[i] - b[i]; }
    if (QLL && g.quant_ll) store_quant_if<NC>(out + (g.band_off[0] + off), v, g.q[0], emit_low);
    else store_raw_if<NC>(out + (g.band_off[0] + off), v, emit_low);
    store_quant_if<NC>(out + (g.band_off[1] + off), v + NC, g.q[1], emit_low);
    {
        int h[2 * NC];
#pragma unroll
        for (int i = 0; i < 2 * NC; i++) h[i] = ((v[i] - s.llp[i] + 4) >> 3) + s.dc[i];
        const unsigned offh = off - (unsigned)g.out_pitch;
        store_quant_if<NC>(out + (g.band_off[2] + offh), h, g.q[2], emit_high);
        store_quant_if<NC>(out + (g.band_off[3] + offh), h + NC, g.q[3], emit_high);
    }
#pragma unroll
    for (int i = 0; i < 2 * NC; i++) { s.llp[i] = s.llc[i]; s.llc[i] = v[i]; s.dc[i] = dn[i]; }
}

// Border rows of HL/HH from three consecutive pairs (S,D of each): spatial.c:10166-10208 / :10516-10558.
//   top   : pairs 0,1,2      -> row 0     = clamp((-3 S0 + 8 D0 + 4 S1 - S2 + 4) >> 3)
//   bottom: pairs n-2,n-1,n  -> row n     = clamp(( 3 Sn + 8 Dn - 4 S(n-1) + S(n-2) + 4) >> 3)
template <int NC>
__device__ __forceinline__ void border_emit(const int *s0, const int *s1, const int *s2, const int *dsel, bool bottom,
                                            const PlaneGeom &g, unsigned char *out, unsigned off)
{
    int h[2 * NC];
#pragma unroll
    for (int i = 0; i < 2 * NC; i++)
        h[i] = bottom ? clamp16((3 * s2[i] + 8 * dsel[i] - 4 * s1[i] + s0[i] + 4) >> 3)
                      : clamp16((-3 * s0[i] + 8 * dsel[i] + 4 * s1[i] - s2[i] + 4) >> 3);
    store_quant<NC>(out + (g.band_off[2] + off), h, g.q[2]);
    store_quant<NC>(out + (g.band_off[3] + off), h + NC, g.q[3]);
}

// ----------------------------------------------------------------------------
// int16 plane input
struct RawPlaneRow {
    uint4 v;            // 8 samples of this lane
    unsigned halo;      // lane 0: samples [-2,-1] of the strip; last lane: samples [+256,+257]
};

__device__ __forceinline__ void load_plane_row(const unsigned char *p, const LaneInfo &L, RawPlaneRow &r)
{
    r.v = __ldg(reinterpret_cast<const uint4 *>(p));
    // ONE predicated halo load per row (two loads into the same register would serialise on its scoreboard)
    r.halo = 0u;
    if (L.use_lh | L.use_rh) r.halo = __ldg(reinterpret_cast<const unsigned *>(p + (L.use_lh ? -4 : 16)));
}

// ---- packed 16-bit RGB (RG48) input: a lane's 8 pixels are 48 contiguous bytes; one channel (word SEL of each
// pixel: 0 = R, 1 = G, 2 = B) is extracted and reduced to the codec precision (>> shift), as
// Codec/frame.c:5968 ConvertRGB48ToFrame16s (default branch :6130-6164) does on the host.
struct RawRG48Row {
    uint4 a, b, c;      // 24 words = 8 pixels x 3
    unsigned halo;      // channel samples of pixels [-2,-1] (lane 0) or [+8,+9] (last lane), already packed
};

template <int SEL>
__device__ __forceinline__ void load_rg48_row(const unsigned char *p, const LaneInfo &L, RawRG48Row &r)
{
    r.a = __ldg(reinterpret_cast<const uint4 *>(p));
    r.b = __ldg(reinterpret_cast<const uint4 *>(p + 16));
    r.c = __ldg(reinterpret_cast<const uint4 *>(p + 32));
    r.halo = 0u;
    if (L.use_lh | L.use_rh) {
        const unsigned char *h = p + (L.use_lh ? -12 : 48) + 2 * SEL;
        r.halo = (unsigned)__ldg(reinterpret_cast<const unsigned short *>(h)) |
                 ((unsigned)__ldg(reinterpret_cast<const unsigned short *>(h + 6)) << 16);
    }
}

// word index w (0..23) of the 48-byte group as a (register, half) pair -> PRMT selector nibble pair
template <int SEL>
__device__ __forceinline__ void rg48_extract(const RawRG48Row &r, int shift, RawPlaneRow &o)
{
    const unsigned w[12] = {r.a.x, r.a.y, r.a.z, r.a.w, r.b.x, r.b.y, r.b.z, r.b.w, r.c.x, r.c.y, r.c.z, r.c.w};
    unsigned out[4];
#pragma unroll
    for (int m = 0; m < 4; m++) {
        const int w0 = 3 * (2 * m) + SEL, w1 = 3 * (2 * m + 1) + SEL;      // word indices of samples 2m, 2m+1
        const unsigned lo = (w0 & 1) ? (w[w0 >> 1] >> 16) : (w[w0 >> 1] & 0xffffu);
        const unsigned hi = (w1 & 1) ? (w[w1 >> 1] & 0xffff0000u) : (w[w1 >> 1] << 16);
        out[m] = lo | hi;
    }
    const unsigned mask = (0xffffu >> shift) * 0x00010001u;
    o.v = make_uint4((out[0] >> shift) & mask, (out[1] >> shift) & mask, (out[2] >> shift) & mask, (out[3] >> shift) & mask);
    o.halo = (r.halo >> shift) & mask;
}

template <int PRESCALE>
__device__ __forceinline__ int tap(int x) { return PRESCALE ? ((x + 3) >> 2) : x; }

// horizontal 2-6 for the lane's 4 output columns: o[0..3] = low, o[4..7] = high
// (Codec/spatial.c:253 FilterHorizontalRow16s / :3669 FilterHorizontalRow10bit16s)
// dp2a with unsigned 16-bit halves (a) and signed byte coefficients (b): lo16(a) * b0 + hi16(a) * b1 + c
__device__ __forceinline__ int dp2a_lo_us(unsigned a, unsigned b, int c) {
    int d;
    asm("dp2a.lo.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
// PRESCALE = 3: the prescaled filter (PRESCALE = 2 arithmetic) for planes known to be NON-NEGATIVE, which every level-2 / 3
// input of the codec pyramid is (an LL band of an unsigned source).  Both taps of a packed pair are then prescaled in the
// packed word -- (w + 0x00030003) >> 2, masked -- and tap sum, tap difference and the lowpass (x0 + x1 + 3) >> 2 are one
// dp2a each: 7 instructions per pair against 10 (the prescaled level is issue-bound: profiles/r02_prof_fwdplane_summary.csv)
constexpr unsigned kOnes2 = 0x0101u, kPlusMinus2 = 0xff01u;       // dp2a byte coefficients (+1, +1) and (+1, -1)
__device__ __forceinline__ unsigned prescale_pair_nonneg(unsigned w) { return ((w + 0x00030003u) >> 2) & 0x3fff3fffu; }

template <int PRESCALE>
__device__ __forceinline__ void hfilter_plane(const RawPlaneRow &r, const LaneInfo &L, int *o)
{
    const unsigned w[4] = {r.v.x, r.v.y, r.v.z, r.v.w};
    int S[4], d[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        if (PRESCALE == 3) {
            const unsigned t = prescale_pair_nonneg(w[k]);
            S[k] = dp2a_lo_us(t, kOnes2, 0);
            d[k] = dp2a_lo_us(t, kPlusMinus2, 0);
            o[k] = dp2a_lo_us(w[k], kOnes2, 3) >> 2;
        } else {
            const int x0 = lo16(w[k]), x1 = hi16(w[k]);
            const int t0 = tap<PRESCALE>(x0), t1 = tap<PRESCALE>(x1);
            S[k] = t0 + t1;
            d[k] = t0 - t1;
            o[k] = PRESCALE ? ((x0 + x1 + 3) >> 2) : S[k];
        }
    }
    int Sp = __shfl_up_sync(L.amask, S[3], 1);
    int Sn = __shfl_down_sync(L.amask, S[0], 1);
    const int hs = (PRESCALE == 3) ? dp2a_lo_us(prescale_pair_nonneg(r.halo), kOnes2, 0)
                                   : tap<PRESCALE>(lo16(r.halo)) + tap<PRESCALE>(hi16(r.halo));
    Sp = L.use_lh ? hs : Sp;
    Sn = L.use_rh ? hs : Sn;
    o[4] = ((S[1] - Sp + 4) >> 3) + d[0];
    o[5] = ((S[2] - S[0] + 4) >> 3) + d[1];
    o[6] = ((S[3] - S[1] + 4) >> 3) + d[2];
    o[7] = ((Sn - S[2] + 4) >> 3) + d[3];
    if (L.has_border) {     // warp-uniform: only the first and last strip of a row carry a border lane
        if (L.left_border) o[4] = clamp16((-3 * S[0] + 8 * d[0] + 4 * S[1] - S[2] + 4) >> 3);
        if (L.right_border) o[7] = clamp16((3 * S[3] + 8 * d[3] - 4 * S[2] + S[1] + 4) >> 3);
    }
}

__device__ __forceinline__ bool lane_setup(int strip, int width, int lane, LaneInfo &L)
{
    // A lane is active only when all of its 8 input columns exist.  When the width is not a multiple of 8 the last
    // 2-6 columns (1-3 output columns, the right border among them) are produced by k_fwd_plane_edge; the last full
    // lane then takes its right neighbour value from a halo word like the last lane of an interior strip does.
    const int col0 = strip * kStripIn + lane * 8;
    const bool active = col0 + 8 <= width;
    L.amask = __ballot_sync(0xffffffffu, active);
    if (!active) return false;
    L.left_border = (col0 == 0);
    L.right_border = (col0 + 8 == width);
    L.has_border = (strip == 0) || ((strip + 1) * kStripIn >= width);
    L.use_lh = (lane == 0) && (strip > 0);
    L.use_rh = (col0 + 8 < width) && (lane == 31 || col0 + 16 > width);
    return true;
}

template <int PRESCALE>
__global__ void __launch_bounds__(128) k_fwd_plane(const __grid_constant__ FwdParams p)
{
    const int lane = threadIdx.x;
    const int f = blockIdx.z / p.nchan, c = blockIdx.z - f * p.nchan;
    const PlaneGeom &g = p.ch[c];
    const int strip = blockIdx.x;
    if (strip * kStripIn >= g.width) return;
    const int oh = g.height >> 1;
    LaneInfo L;
    if (!lane_setup(strip, g.width, lane, L)) return;
    const unsigned colbyte = (unsigned)((strip * kStripOut + lane * 4) * 2);
    const unsigned char *in = p.in_base[f] + g.in_off + (strip * kStripIn + lane * 8) * 2;
    unsigned char *out = p.out_base[f];

    if (blockIdx.y == gridDim.y - 1) {
        // ---- border warps: warp 0 -> first HL/HH row, warp 1 -> last HL/HH row ----
        if (threadIdx.y > 1) return;
        const bool bottom = (threadIdx.y == 1);
        const int j0 = bottom ? oh - 3 : 0;
        int s[3][8], dsel[8];
#pragma unroll
        for (int k = 0; k < 3; k++) {
            RawPlaneRow r0, r1;
            int a[8], b[8];
            load_plane_row(in + (long long)(2 * (j0 + k)) * g.in_pitch, L, r0);
            load_plane_row(in + (long long)(2 * (j0 + k) + 1) * g.in_pitch, L, r1);
            hfilter_plane<PRESCALE>(r0, L, a);
            hfilter_plane<PRESCALE>(r1, L, b);
#pragma unroll
            for (int i = 0; i < 8; i++) {
                s[k][i] = a[i] + b[i];
                if (k == (bottom ? 2 : 0)) dsel[i] = a[i] - b[i];
            }
        }
        border_emit<4>(s[0], s[1], s[2], dsel, bottom, g, out, (unsigned)((bottom ? oh - 1 : 0) * g.out_pitch) + colbyte);
        return;
    }

    const int y0 = (blockIdx.y * blockDim.y + threadIdx.y) * p.th;
    if (y0 >= oh) return;
    const int y1 = min(y0 + p.th, oh);
    const int jfirst = max(y0 - 1, 0), jlast = min(y1, oh - 1);
    const int hlo = max(y0, 1);     // first HL/HH row this warp emits (row 0 belongs to the border warp)

    VState<4> st;
#pragma unroll
    for (int i = 0; i < 8; i++) { st.llp[i] = st.llc[i] = st.dc[i] = 0; }

    const unsigned char *rp = in + (long long)(2 * jfirst) * g.in_pitch;
    RawPlaneRow c0, c1, n0, n1;
    load_plane_row(rp, L, c0);
    load_plane_row(rp + g.in_pitch, L, c1);
    n0 = c0; n1 = c1;
    unsigned off = (unsigned)(jfirst * g.out_pitch) + colbyte;
    for (int j = jfirst; j <= jlast; j++) {
        rp += 2 * g.in_pitch;
        if (j < jlast) {
            load_plane_row(rp, L, n0);
            load_plane_row(rp + g.in_pitch, L, n1);
        }
        if (j + 2 < jlast) { prefetch_l2(rp + 4 * g.in_pitch); prefetch_l2(rp + 5 * g.in_pitch); }
        int a[8], b[8];
        hfilter_plane<PRESCALE>(c0, L, a);
        hfilter_plane<PRESCALE>(c1, L, b);
        vstep<4, 1>(st, a, b, g, out, off, j >= y0 && j < y1, j - 1 >= hlo);
        off += (unsigned)g.out_pitch;
        c0 = n0; c1 = n1;
    }
}

// ----------------------------------------------------------------------------
// Ragged widths: output columns [4 * (width / 8), width / 2) of a plane level, one thread per coefficient position,
// written exactly as the formulas read (spatial.c:253-570 rows, :10166-10558 columns).  At most 3 columns per row, so
// the cost is nil; it keeps the lane-granular main kernel free of partial-lane cases.
template <int PRESCALE>
__global__ void __launch_bounds__(128) k_fwd_plane_edge(const __grid_constant__ FwdParams p)
{
    const int f = blockIdx.z / p.nchan, c = blockIdx.z - f * p.nchan;
    const PlaneGeom &g = p.ch[c];
    const int w = g.width, ow = w >> 1, oh = g.height >> 1;
    const int col = (w >> 3) * 4 + blockIdx.y;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= ow || j >= oh) return;
    const unsigned char *in = p.in_base[f] + g.in_off;
    auto px = [&](int r, int i) { return (int)*reinterpret_cast<const short *>(in + (long long)r * g.in_pitch + 2 * i); };
    auto hrow = [&](int r, int &lo, int &hi) {
        auto t = [&](int i) { return tap<PRESCALE>(px(r, i)); };
        const int i0 = 2 * col;
        lo = PRESCALE ? ((px(r, i0) + px(r, i0 + 1) + 3) >> 2) : (px(r, i0) + px(r, i0 + 1));
        if (col == ow - 1)
            hi = clamp16((11 * t(w - 2) - 5 * t(w - 1) - 4 * t(w - 3) - 4 * t(w - 4) + t(w - 5) + t(w - 6) + 4) >> 3);
        else
            hi = ((-t(i0 - 2) - t(i0 - 1) + t(i0 + 2) + t(i0 + 3) + 4) >> 3) + t(i0) - t(i0 + 1);
    };
    int l[6], h[6];
    const int r0 = (j == 0) ? 0 : ((j == oh - 1) ? g.height - 6 : 2 * j - 2);
#pragma unroll
    for (int k = 0; k < 6; k++) hrow(r0 + k, l[k], h[k]);
    int ll, lh, hl, hh;
    if (j == 0) {
        ll = l[0] + l[1]; lh = h[0] + h[1];
        hl = clamp16((5 * l[0] - 11 * l[1] + 4 * l[2] + 4 * l[3] - l[4] - l[5] + 4) >> 3);
        hh = clamp16((5 * h[0] - 11 * h[1] + 4 * h[2] + 4 * h[3] - h[4] - h[5] + 4) >> 3);
    } else if (j == oh - 1) {
        ll = l[4] + l[5]; lh = h[4] + h[5];
        hl = clamp16((11 * l[4] - 5 * l[5] - 4 * l[3] - 4 * l[2] + l[1] + l[0] + 4) >> 3);
        hh = clamp16((11 * h[4] - 5 * h[5] - 4 * h[3] - 4 * h[2] + h[1] + h[0] + 4) >> 3);
    } else {
        ll = l[2] + l[3]; lh = h[2] + h[3];
        hl = ((-l[0] - l[1] + l[4] + l[5] + 4) >> 3) + l[2] - l[3];
        hh = ((-h[0] - h[1] + h[4] + h[5] + 4) >> 3) + h[2] - h[3];
    }
    unsigned char *out = p.out_base[f] + (long long)j * g.out_pitch + 2 * col;
    *reinterpret_cast<short *>(out + g.band_off[0]) = (short)(g.quant_ll ? (quant1(ll, g.q[0]) >> 16) : ll);
    *reinterpret_cast<short *>(out + g.band_off[1]) = (short)(quant1(lh, g.q[1]) >> 16);
    *reinterpret_cast<short *>(out + g.band_off[2]) = (short)(quant1(hl, g.q[2]) >> 16);
    *reinterpret_cast<short *>(out + g.band_off[3]) = (short)(quant1(hh, g.q[3]) >> 16);
}

// ----------------------------------------------------------------------------
// level 1 of one channel of a packed RG48 frame (prescale 0, Codec/spatial.c:10026 on the 12-bit plane that
// ConvertRGB48ToFrame16s would have produced).  One launch per channel: SEL picks the word of each pixel.
template <int SEL>
__global__ void __launch_bounds__(128) k_fwd_rg48(const __grid_constant__ FwdParams p)
{
    const int lane = threadIdx.x;
    const int f = blockIdx.z;
    const PlaneGeom &g = p.ch[0];
    const int strip = blockIdx.x;
    if (strip * kStripIn >= g.width) return;
    const int oh = g.height >> 1;
    LaneInfo L;
    if (!lane_setup(strip, g.width, lane, L)) return;
    const unsigned colbyte = (unsigned)((strip * kStripOut + lane * 4) * 2);
    const unsigned char *in = p.in_base[f] + g.in_off + (long long)(strip * kStripIn + lane * 8) * 6;
    unsigned char *out = p.out_base[f];
    const int shift = p.shift;          // 16 - precision

    if (blockIdx.y == gridDim.y - 1) {
        if (threadIdx.y > 1) return;
        const bool bottom = (threadIdx.y == 1);
        const int j0 = bottom ? oh - 3 : 0;
        int s[3][8], dsel[8];
#pragma unroll
        for (int k = 0; k < 3; k++) {
            RawRG48Row q0, q1;
            RawPlaneRow r0, r1;
            int a[8], b[8];
            load_rg48_row<SEL>(in + (long long)(2 * (j0 + k)) * g.in_pitch, L, q0);
            load_rg48_row<SEL>(in + (long long)(2 * (j0 + k) + 1) * g.in_pitch, L, q1);
            rg48_extract<SEL>(q0, shift, r0);
            rg48_extract<SEL>(q1, shift, r1);
            hfilter_plane<0>(r0, L, a);
            hfilter_plane<0>(r1, L, b);
#pragma unroll
            for (int i = 0; i < 8; i++) {
                s[k][i] = a[i] + b[i];
                if (k == (bottom ? 2 : 0)) dsel[i] = a[i] - b[i];
            }
        }
        border_emit<4>(s[0], s[1], s[2], dsel, bottom, g, out, (unsigned)((bottom ? oh - 1 : 0) * g.out_pitch) + colbyte);
        return;
    }

    const int y0 = (blockIdx.y * blockDim.y + threadIdx.y) * p.th;
    if (y0 >= oh) return;
    const int y1 = min(y0 + p.th, oh);
    const int jfirst = max(y0 - 1, 0), jlast = min(y1, oh - 1);
    const int hlo = max(y0, 1);

    VState<4> st;
#pragma unroll
    for (int i = 0; i < 8; i++) { st.llp[i] = st.llc[i] = st.dc[i] = 0; }

    const unsigned char *rp = in + (long long)(2 * jfirst) * g.in_pitch;
    RawRG48Row c0, c1, n0, n1;
    load_rg48_row<SEL>(rp, L, c0);
    load_rg48_row<SEL>(rp + g.in_pitch, L, c1);
    n0 = c0; n1 = c1;
    unsigned off = (unsigned)(jfirst * g.out_pitch) + colbyte;
    for (int j = jfirst; j <= jlast; j++) {
        rp += 2 * g.in_pitch;
        if (j < jlast) {
            load_rg48_row<SEL>(rp, L, n0);
            load_rg48_row<SEL>(rp + g.in_pitch, L, n1);
        }
        RawPlaneRow r0, r1;
        int a[8], b[8];
        rg48_extract<SEL>(c0, shift, r0);
        rg48_extract<SEL>(c1, shift, r1);
        hfilter_plane<0>(r0, L, a);
        hfilter_plane<0>(r1, L, b);
        vstep<4, 1>(st, a, b, g, out, off, j >= y0 && j < y1, j - 1 >= hlo);
        off += (unsigned)g.out_pitch;
        c0 = n0; c1 = n1;
    }
}

// ----------------------------------------------------------------------------
// level 1 of one channel of a 16-bit Bayer frame (BYR4, curve already applied): the four half-resolution planes
// G = (g1+g2)>>1, RG = (r-G+4096)>>1, BG = (b-G+4096)>>1, DG = (g1-g2+4096)>>1 at 12 bits
// (Codec/frame.c:4993 ConvertBYR4ToFrame16s, encode_curve_preset branch :5040-5200) are formed on the fly from the
// two Bayer lines of each plane row.  blockIdx.x = strip * 4 + channel, so the four channel jobs of a strip run
// next to each other and share the Bayer lines through L1/L2.
struct RawBYR4Row {
    uint4 a0, a1;       // Bayer line 2r   : 16 pixels of this lane
    uint4 b0, b1;       // Bayer line 2r+1
    uint2 ha, hb;       // 4 pixels of each line just outside the strip (lane 0: left, last lane: right)
};

__device__ __forceinline__ void load_byr4_row(const unsigned char *p, int line_pitch, const LaneInfo &L, RawBYR4Row &r)
{
    r.a0 = __ldg(reinterpret_cast<const uint4 *>(p));
    r.a1 = __ldg(reinterpret_cast<const uint4 *>(p + 16));
    r.b0 = __ldg(reinterpret_cast<const uint4 *>(p + line_pitch));
    r.b1 = __ldg(reinterpret_cast<const uint4 *>(p + line_pitch + 16));
    r.ha = make_uint2(0u, 0u); r.hb = make_uint2(0u, 0u);
    if (L.use_lh | L.use_rh) {
        const unsigned char *h = p + (L.use_lh ? -8 : 32);
        r.ha = __ldg(reinterpret_cast<const uint2 *>(h));
        r.hb = __ldg(reinterpret_cast<const uint2 *>(h + line_pitch));
    }
}

// one plane sample from the quad (w1 = two pixels of the first line, w2 = of the second line)
// LUT: the encode curve of Codec/frame.c:5208-5330 (default: log base 90), indexed by the 14 most significant bits
// (MAX_INPUT_PRECISION, frame.c:4843); without it the frame is taken as already curved (`>> shift`, encode_curve_preset).
template <bool LUT>
__device__ __forceinline__ int byr4_sample(unsigned w1, unsigned w2, int shift, int fmt, int chan, const unsigned short *lut)
{
    int q0, q1, q2, q3;
    if (LUT) {
        q0 = __ldg(lut + ((w1 & 0xffffu) >> 2)); q1 = __ldg(lut + (w1 >> 18));
        q2 = __ldg(lut + ((w2 & 0xffffu) >> 2)); q3 = __ldg(lut + (w2 >> 18));
    } else {
        q0 = (int)((w1 & 0xffffu) >> shift); q1 = (int)((w1 >> 16) >> shift);
        q2 = (int)((w2 & 0xffffu) >> shift); q3 = (int)((w2 >> 16) >> shift);
    }
    const bool g_second = (fmt == 0) || (fmt == 3);             // RED_GRN / BLU_GRN: green is the 2nd pixel of line 1
    const int g1 = g_second ? q1 : q0, g2 = g_second ? q2 : q3;
    if (chan == 3) return (g1 - g2 + 4096) >> 1;
    const int gg = (g1 + g2) >> 1;
    if (chan == 0) return gg;
    const int r = (fmt == 0) ? q0 : (fmt == 1) ? q1 : (fmt == 2) ? q2 : q3;
    const int b = (fmt == 0) ? q3 : (fmt == 1) ? q2 : (fmt == 2) ? q1 : q0;
    return (((chan == 1) ? r : b) - gg + 4096) >> 1;
}

template <bool LUT>
__device__ __forceinline__ void byr4_extract(const RawBYR4Row &r, int shift, int fmt, int chan, const unsigned short *lut, RawPlaneRow &o)
{
    const unsigned l1[8] = {r.a0.x, r.a0.y, r.a0.z, r.a0.w, r.a1.x, r.a1.y, r.a1.z, r.a1.w};
    const unsigned l2[8] = {r.b0.x, r.b0.y, r.b0.z, r.b0.w, r.b1.x, r.b1.y, r.b1.z, r.b1.w};
    unsigned out[4];
#pragma unroll
    for (int m = 0; m < 4; m++)
        out[m] = (unsigned)byr4_sample<LUT>(l1[2 * m], l2[2 * m], shift, fmt, chan, lut) |
                 ((unsigned)byr4_sample<LUT>(l1[2 * m + 1], l2[2 * m + 1], shift, fmt, chan, lut) << 16);
    o.v = make_uint4(out[0], out[1], out[2], out[3]);
    o.halo = (unsigned)byr4_sample<LUT>(r.ha.x, r.hb.x, shift, fmt, chan, lut) | ((unsigned)byr4_sample<LUT>(r.ha.y, r.hb.y, shift, fmt, chan, lut) << 16);
}

// The same plane samples with the channel known at compile time and the Bayer phase folded into byte-permute selectors
// (warp-uniform registers): g1 always sits on the first line of a quad and g2 on the second, red / blue on either.
//   selg1 / selg2: halfword of the line word holding g1 / g2;  selx: halfword holding the channel's colour sample,
//   xline: 0 = first line, 1 = second line
struct BayerSel { unsigned selg1, selg2, selx; int xline; };

__device__ __forceinline__ BayerSel bayer_sel(int fmt, int chan)
{
    // quad layout (line 1: q0 q1, line 2: q2 q3): 0 RED_GRN r g / g b, 1 GRN_RED g r / b g, 2 GRN_BLU g b / r g, 3 BLU_GRN b g / g r
    const unsigned lo = 0x4410u, hi = 0x4432u;      // PRMT selectors: low / high halfword, upper half zero (second operand = 0)
    BayerSel s;
    const bool g_second = (fmt == 0) || (fmt == 3);
    s.selg1 = g_second ? hi : lo;
    s.selg2 = g_second ? lo : hi;
    const int rpos = (fmt == 0) ? 0 : (fmt == 1) ? 1 : (fmt == 2) ? 2 : 3;
    const int bpos = 3 - rpos;
    const int xpos = (chan == 1) ? rpos : bpos;
    s.selx = (xpos & 1) ? hi : lo;
    s.xline = xpos >> 1;
    return s;
}

template <bool LUT, int CHAN>
__device__ __forceinline__ int byr4_sample_c(unsigned w1, unsigned w2, int shift, const BayerSel &s, const unsigned short *lut)
{
    unsigned g1 = __byte_perm(w1, 0u, s.selg1), g2 = __byte_perm(w2, 0u, s.selg2);
    if (LUT) { g1 = __ldg(lut + (g1 >> 2)); g2 = __ldg(lut + (g2 >> 2)); } else { g1 >>= shift; g2 >>= shift; }
    if (CHAN == 3) return (int)(g1 - g2 + 4096u) >> 1;
    const unsigned gg = (g1 + g2) >> 1;
    if (CHAN == 0) return (int)gg;
    unsigned x = __byte_perm(s.xline ? w2 : w1, 0u, s.selx);
    if (LUT) x = __ldg(lut + (x >> 2)); else x >>= shift;
    return (int)(x - gg + 4096u) >> 1;
}

template <bool LUT, int CHAN>
__device__ __forceinline__ void byr4_extract_c(const RawBYR4Row &r, int shift, const BayerSel &s, const unsigned short *lut, RawPlaneRow &o)
{
    const unsigned l1[8] = {r.a0.x, r.a0.y, r.a0.z, r.a0.w, r.a1.x, r.a1.y, r.a1.z, r.a1.w};
    const unsigned l2[8] = {r.b0.x, r.b0.y, r.b0.z, r.b0.w, r.b1.x, r.b1.y, r.b1.z, r.b1.w};
    unsigned out[4];
#pragma unroll
    for (int m = 0; m < 4; m++)
        out[m] = pack_lo(byr4_sample_c<LUT, CHAN>(l1[2 * m], l2[2 * m], shift, s, lut),
                         byr4_sample_c<LUT, CHAN>(l1[2 * m + 1], l2[2 * m + 1], shift, s, lut));
    o.v = make_uint4(out[0], out[1], out[2], out[3]);
    o.halo = pack_lo(byr4_sample_c<LUT, CHAN>(r.ha.x, r.hb.x, shift, s, lut), byr4_sample_c<LUT, CHAN>(r.ha.y, r.hb.y, shift, s, lut));
}

template <bool LUT>
__global__ void __launch_bounds__(128) k_fwd_byr4(const __grid_constant__ FwdParams p)
{
    const int lane = threadIdx.x;
    const int f = blockIdx.z;
    const int c = blockIdx.x & 3, strip = blockIdx.x >> 2;
    const PlaneGeom &g = p.ch[c];
    if (strip * kStripIn >= g.width) return;
    const int oh = g.height >> 1;
    LaneInfo L;
    if (!lane_setup(strip, g.width, lane, L)) return;
    const unsigned colbyte = (unsigned)((strip * kStripOut + lane * 4) * 2);
    const int line_pitch = g.in_pitch;                      // bytes per Bayer line
    const long long row_pitch = 2LL * line_pitch;           // one plane row = two Bayer lines
    const unsigned char *in = p.in_base[f] + (long long)(strip * kStripIn + lane * 8) * 4;      // 2 pixels x 2 bytes per plane sample
    unsigned char *out = p.out_base[f];
    const int shift = p.shift, fmt = p.uyvy;                // uyvy field reused as the Bayer phase (0..3)

    if (blockIdx.y == gridDim.y - 1) {
        if (threadIdx.y > 1) return;
        const bool bottom = (threadIdx.y == 1);
        const int j0 = bottom ? oh - 3 : 0;
        int s[3][8], dsel[8];
#pragma unroll
        for (int k = 0; k < 3; k++) {
            RawBYR4Row q0, q1;
            RawPlaneRow r0, r1;
            int a[8], b[8];
            load_byr4_row(in + (long long)(2 * (j0 + k)) * row_pitch, line_pitch, L, q0);
            load_byr4_row(in + (long long)(2 * (j0 + k) + 1) * row_pitch, line_pitch, L, q1);
            byr4_extract<LUT>(q0, shift, fmt, c, p.lut, r0);
            byr4_extract<LUT>(q1, shift, fmt, c, p.lut, r1);
            hfilter_plane<0>(r0, L, a);
            hfilter_plane<0>(r1, L, b);
#pragma unroll
            for (int i = 0; i < 8; i++) {
                s[k][i] = a[i] + b[i];
                if (k == (bottom ? 2 : 0)) dsel[i] = a[i] - b[i];
            }
        }
        border_emit<4>(s[0], s[1], s[2], dsel, bottom, g, out, (unsigned)((bottom ? oh - 1 : 0) * g.out_pitch) + colbyte);
        return;
    }

    const int y0 = (blockIdx.y * blockDim.y + threadIdx.y) * p.th;
    if (y0 >= oh) return;
    const int y1 = min(y0 + p.th, oh);
    const int jfirst = max(y0 - 1, 0), jlast = min(y1, oh - 1);
    const int hlo = max(y0, 1);

    VState<4> st;
#pragma unroll
    for (int i = 0; i < 8; i++) { st.llp[i] = st.llc[i] = st.dc[i] = 0; }

    const unsigned char *rp = in + (long long)(2 * jfirst) * row_pitch;
    RawBYR4Row c0, c1, n0, n1;
    load_byr4_row(rp, line_pitch, L, c0);
    load_byr4_row(rp + row_pitch, line_pitch, L, c1);
    n0 = c0; n1 = c1;
    unsigned off = (unsigned)(jfirst * g.out_pitch) + colbyte;
    for (int j = jfirst; j <= jlast; j++) {
        rp += 2 * row_pitch;
        if (j < jlast) {
            load_byr4_row(rp, line_pitch, L, n0);
            load_byr4_row(rp + row_pitch, line_pitch, L, n1);
        }
        RawPlaneRow r0, r1;
        int a[8], b[8];
        byr4_extract<LUT>(c0, shift, fmt, c, p.lut, r0);
        byr4_extract<LUT>(c1, shift, fmt, c, p.lut, r1);
        hfilter_plane<0>(r0, L, a);
        hfilter_plane<0>(r1, L, b);
        vstep<4, 1>(st, a, b, g, out, off, j >= y0 && j < y1, j - 1 >= hlo);
        off += (unsigned)g.out_pitch;
        c0 = n0; c1 = n1;
    }
}

// ----------------------------------------------------------------------------
// packed 8-bit 4:2:2 input: one warp produces the Y strip (128 columns) and the matching
// U and V strips (64 columns each) from a single read of the packed rows.
struct Raw422Row {
    uint4 v;        // 8 luma + 4 U + 4 V of this lane
    uint2 halo;     // lane 0: previous 8 bytes; last lane: next 8 bytes
};

__device__ __forceinline__ void load_422_row(const unsigned char *p, const LaneInfo &L, Raw422Row &r)
{
    r.v = __ldg(reinterpret_cast<const uint4 *>(p));
    // ONE predicated halo load per row (two loads into the same register would serialise on its scoreboard)
    r.halo = make_uint2(0u, 0u);
    if (L.use_lh | L.use_rh) r.halo = __ldg(reinterpret_cast<const uint2 *>(p + (L.use_lh ? -8 : 16)));
}

struct Sel422 {     // dp4a coefficient words (already scaled by 1 << shift)
    int ysum, ydif, u, v;
};

// Y: oy[0..3] low, oy[4..7] high.  U/V: o[0..1] low, o[2..3] high.
__device__ __forceinline__ void hfilter_422(const Raw422Row &r, const Sel422 &sel, const LaneInfo &L, int *oy, int *ou, int *ov)
{
    const unsigned w[4] = {r.v.x, r.v.y, r.v.z, r.v.w};
    int S[4], d[4], cu[4], cv[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        S[k] = dp4a_us(w[k], sel.ysum, 0);
        d[k] = dp4a_us(w[k], sel.ydif, 0);
        cu[k] = dp4a_us(w[k], sel.u, 0);
        cv[k] = dp4a_us(w[k], sel.v, 0);
        oy[k] = S[k];
    }
    const int Su[2] = {cu[0] + cu[1], cu[2] + cu[3]}, du[2] = {cu[0] - cu[1], cu[2] - cu[3]};
    const int Sv[2] = {cv[0] + cv[1], cv[2] + cv[3]}, dv[2] = {cv[0] - cv[1], cv[2] - cv[3]};
    int Sp = __shfl_up_sync(L.amask, S[3], 1), Sn = __shfl_down_sync(L.amask, S[0], 1);
    int Sup = __shfl_up_sync(L.amask, Su[1], 1), Sun = __shfl_down_sync(L.amask, Su[0], 1);
    int Svp = __shfl_up_sync(L.amask, Sv[1], 1), Svn = __shfl_down_sync(L.amask, Sv[0], 1);
    if (L.use_lh | L.use_rh) {      // lane 0 / 31 of strips with a neighbour strip (divergent but tiny)
        // left halo: luma pair of the later word (.y), chroma pair = both words; right halo: luma pair of word .x
        const int hy = dp4a_us(L.use_lh ? r.halo.y : r.halo.x, sel.ysum, 0);
        const int hu = dp4a_us(r.halo.y, sel.u, dp4a_us(r.halo.x, sel.u, 0));
        const int hv = dp4a_us(r.halo.y, sel.v, dp4a_us(r.halo.x, sel.v, 0));
        if (L.use_lh) { Sp = hy; Sup = hu; Svp = hv; } else { Sn = hy; Sun = hu; Svn = hv; }
    }
    oy[4] = ((S[1] - Sp + 4) >> 3) + d[0];
    oy[5] = ((S[2] - S[0] + 4) >> 3) + d[1];
    oy[6] = ((S[3] - S[1] + 4) >> 3) + d[2];
    oy[7] = ((Sn - S[2] + 4) >> 3) + d[3];
    ou[0] = Su[0]; ou[1] = Su[1];
    ou[2] = ((Su[1] - Sup + 4) >> 3) + du[0];
    ou[3] = ((Sun - Su[0] + 4) >> 3) + du[1];
    ov[0] = Sv[0]; ov[1] = Sv[1];
    ov[2] = ((Sv[1] - Svp + 4) >> 3) + dv[0];
    ov[3] = ((Svn - Sv[0] + 4) >> 3) + dv[1];
    if (L.has_border) {
        if (L.left_border) {
            oy[4] = clamp16((-3 * S[0] + 8 * d[0] + 4 * S[1] - S[2] + 4) >> 3);
            ou[2] = clamp16((-3 * Su[0] + 8 * du[0] + 4 * Su[1] - Sun + 4) >> 3);
            ov[2] = clamp16((-3 * Sv[0] + 8 * dv[0] + 4 * Sv[1] - Svn + 4) >> 3);
        }
        if (L.right_border) {
            oy[7] = clamp16((3 * S[3] + 8 * d[3] - 4 * S[2] + S[1] + 4) >> 3);
            ou[3] = clamp16((3 * Su[1] + 8 * du[1] - 4 * Su[0] + Sup + 4) >> 3);
            ov[3] = clamp16((3 * Sv[1] + 8 * dv[1] - 4 * Sv[0] + Svp + 4) >> 3);
        }
    }
}

// channel numbering of the reference: 0 = Y, 1 = V, 2 = U (Codec/convert.c:4793)
__global__ void __launch_bounds__(128) k_fwd_422(const __grid_constant__ FwdParams p)
{
    const int lane = threadIdx.x;
    const int f = blockIdx.z;
    const PlaneGeom &gy = p.ch[0];
    const PlaneGeom &gv = p.ch[1];
    const PlaneGeom &gu = p.ch[2];
    const int strip = blockIdx.x;
    if (strip * kStripIn >= gy.width) return;
    const int oh = gy.height >> 1;
    LaneInfo L;
    if (!lane_setup(strip, gy.width, lane, L)) return;
    const unsigned colbyte_y = (unsigned)((strip * kStripOut + lane * 4) * 2);
    const unsigned colbyte_c = (unsigned)((strip * (kStripOut / 2) + lane * 2) * 2);
    const unsigned char *in = p.in_base[f] + gy.in_off + (strip * kStripIn + lane * 8) * 2;
    unsigned char *out = p.out_base[f];

    Sel422 sel;
    {
        const int m = 1 << p.shift;
        const int neg = (-m) & 0xff;
        if (!p.uyvy) {          // Y0 U Y1 V
            sel.ysum = m | (m << 16); sel.ydif = m | (neg << 16); sel.u = m << 8; sel.v = m << 24;
        } else {                // U Y0 V Y1
            sel.ysum = (m << 8) | (m << 24); sel.ydif = (m << 8) | (neg << 24); sel.u = m; sel.v = m << 16;
        }
    }

    if (blockIdx.y == gridDim.y - 1) {
        // ---- border warps: warp 0 -> first HL/HH row, warp 1 -> last HL/HH row (all three channels) ----
        if (threadIdx.y > 1) return;
        const bool bottom = (threadIdx.y == 1);
        const int j0 = bottom ? oh - 3 : 0;
        int sy[3][8], su[3][4], sv[3][4], dy[8], du[4], dv[4];
#pragma unroll
        for (int k = 0; k < 3; k++) {
            Raw422Row r0, r1;
            int ay[8], by[8], au[4], bu[4], av[4], bv[4];
            load_422_row(in + (long long)(2 * (j0 + k)) * gy.in_pitch, L, r0);
            load_422_row(in + (long long)(2 * (j0 + k) + 1) * gy.in_pitch, L, r1);
            hfilter_422(r0, sel, L, ay, au, av);
            hfilter_422(r1, sel, L, by, bu, bv);
            const bool keep = (k == (bottom ? 2 : 0));
#pragma unroll
            for (int i = 0; i < 8; i++) { sy[k][i] = ay[i] + by[i]; if (keep) dy[i] = ay[i] - by[i]; }
#pragma unroll
            for (int i = 0; i < 4; i++) {
                su[k][i] = au[i] + bu[i]; sv[k][i] = av[i] + bv[i];
                if (keep) { du[i] = au[i] - bu[i]; dv[i] = av[i] - bv[i]; }
            }
        }
        const int row = bottom ? oh - 1 : 0;
        border_emit<4>(sy[0], sy[1], sy[2], dy, bottom, gy, out, (unsigned)(row * gy.out_pitch) + colbyte_y);
        border_emit<2>(su[0], su[1], su[2], du, bottom, gu, out, (unsigned)(row * gu.out_pitch) + colbyte_c);
        border_emit<2>(sv[0], sv[1], sv[2], dv, bottom, gv, out, (unsigned)(row * gv.out_pitch) + colbyte_c);
        return;
    }

    const int y0 = (blockIdx.y * blockDim.y + threadIdx.y) * p.th;
    if (y0 >= oh) return;
    const int y1 = min(y0 + p.th, oh);
    const int jfirst = max(y0 - 1, 0), jlast = min(y1, oh - 1);
    const int hlo = max(y0, 1);

    VState<4> sy;
    VState<2> su, sv;
#pragma unroll
    for (int i = 0; i < 8; i++) { sy.llp[i] = sy.llc[i] = sy.dc[i] = 0; }
#pragma unroll
    for (int i = 0; i < 4; i++) { su.llp[i] = su.llc[i] = su.dc[i] = 0; sv.llp[i] = sv.llc[i] = sv.dc[i] = 0; }

    const unsigned char *rp = in + (long long)(2 * jfirst) * gy.in_pitch;
    Raw422Row c0, c1, n0, n1;
    load_422_row(rp, L, c0);
    load_422_row(rp + gy.in_pitch, L, c1);
    n0 = c0; n1 = c1;
    unsigned offy = (unsigned)(jfirst * gy.out_pitch) + colbyte_y;
    unsigned offc = (unsigned)(jfirst * gu.out_pitch) + colbyte_c;
    for (int j = jfirst; j <= jlast; j++) {
        rp += 2 * gy.in_pitch;
        if (j < jlast) {
            load_422_row(rp, L, n0);
            load_422_row(rp + gy.in_pitch, L, n1);
        }
        if (j + 2 < jlast) { prefetch_l2(rp + 4 * gy.in_pitch); prefetch_l2(rp + 5 * gy.in_pitch); }
        int ay[8], by[8], au[4], bu[4], av[4], bv[4];
        hfilter_422(c0, sel, L, ay, au, av);
        hfilter_422(c1, sel, L, by, bu, bv);
        const bool emit_low = (j >= y0) && (j < y1), emit_high = (j - 1 >= hlo);
        vstep<4, 0>(sy, ay, by, gy, out, offy, emit_low, emit_high);
        vstep<2, 0>(su, au, bu, gu, out, offc, emit_low, emit_high);
        vstep<2, 0>(sv, av, bv, gv, out, offc, emit_low, emit_high);
        offy += (unsigned)gy.out_pitch;
        offc += (unsigned)gu.out_pitch;
        c0 = n0; c1 = n1;
    }
}

// ----------------------------------------------------------------------------
// Building blocks of the second-generation level-1 kernel (k_fwd_422_tma, cfb_forward_tma.inl).  The first version
// (k_fwd_422 above, kept selectable with CFB_FWD422=r1 for the A/B in profiles/) spent ~12 % of its issue slots on
// register moves (vertical state shuffle llp <- llc <- v, row double buffer c <- n), ~4 % on constant reloads (LDC)
// and a few per cent on divergence-safe branches around the border code.  Here the vertical state is two values per
// column instead of three, and strips with an image border run their own instantiation of the row loop, so interior
// strips carry no border code (the choice is warp-uniform and made once).  Results are bit-identical.
// (Measured and rejected, profiles/r02_ab_fwd422.txt: unrolling the row loop by two to rotate register roles instead
// of moving values -- fewer instructions but 168-214 registers, 173 us against 160 us.)
// Vertical state per column: two values instead of three.  With t_j = 8 D_j - S_{j-1} the interior highpass row is
//   high_{j-1} = ((S_j - S_{j-2} + 4) >> 3) + D_{j-1} = (S_j + t_{j-1} + 4) >> 3      (8 D is a multiple of 8: exact)
// so a step needs t_{j-1} and S_{j-1} only (to form t_j); S_{j-2} and D_{j-1} are never kept separately.
template <int NC> struct RotState {
    int t[2 * NC];      // t_{j-1} = 8 D_{j-1} - S_{j-2}
    int s[2 * NC];      // S_{j-1}
};

template <int NC, int QLL>
__device__ __forceinline__ void vstep_rot(RotState<NC> &st, const int *a, const int *b, const PlaneGeom &g, unsigned char *out,
                                          unsigned off, bool emit_low, bool emit_high)
{
    int v[2 * NC], h[2 * NC];
#pragma unroll
    for (int i = 0; i < 2 * NC; i++) {
        v[i] = a[i] + b[i];
        h[i] = (v[i] + st.t[i] + 4) >> 3;
        st.t[i] = ((a[i] - b[i]) << 3) - st.s[i];
        st.s[i] = v[i];
    }
    if (QLL && g.quant_ll) store_quant_if<NC>(out + (g.band_off[0] + off), v, g.q[0], emit_low);
    else store_raw_if<NC>(out + (g.band_off[0] + off), v, emit_low);
    store_quant_if<NC>(out + (g.band_off[1] + off), v + NC, g.q[1], emit_low);
    const unsigned offh = off - (unsigned)g.out_pitch;
    store_quant_if<NC>(out + (g.band_off[2] + offh), h, g.q[2], emit_high);
    store_quant_if<NC>(out + (g.band_off[3] + offh), h + NC, g.q[3], emit_high);
}

// hfilter_422 with the border decision lifted to a template parameter
template <bool BORDER>
__device__ __forceinline__ void hfilter_422_t(const Raw422Row &r, const Sel422 &sel, const LaneInfo &L, int *oy, int *ou, int *ov)
{
    const unsigned w[4] = {r.v.x, r.v.y, r.v.z, r.v.w};
    int S[4], d[4], cu[4], cv[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        S[k] = dp4a_us(w[k], sel.ysum, 0);
        d[k] = dp4a_us(w[k], sel.ydif, 0);
        cu[k] = dp4a_us(w[k], sel.u, 0);
        cv[k] = dp4a_us(w[k], sel.v, 0);
        oy[k] = S[k];
    }
    const int Su[2] = {cu[0] + cu[1], cu[2] + cu[3]}, du[2] = {cu[0] - cu[1], cu[2] - cu[3]};
    const int Sv[2] = {cv[0] + cv[1], cv[2] + cv[3]}, dv[2] = {cv[0] - cv[1], cv[2] - cv[3]};
    int Sp = __shfl_up_sync(L.amask, S[3], 1), Sn = __shfl_down_sync(L.amask, S[0], 1);
    int Sup = __shfl_up_sync(L.amask, Su[1], 1), Sun = __shfl_down_sync(L.amask, Su[0], 1);
    int Svp = __shfl_up_sync(L.amask, Sv[1], 1), Svn = __shfl_down_sync(L.amask, Sv[0], 1);
    if (L.use_lh | L.use_rh) {
        const int hy = dp4a_us(L.use_lh ? r.halo.y : r.halo.x, sel.ysum, 0);
        const int hu = dp4a_us(r.halo.y, sel.u, dp4a_us(r.halo.x, sel.u, 0));
        const int hv = dp4a_us(r.halo.y, sel.v, dp4a_us(r.halo.x, sel.v, 0));
        if (L.use_lh) { Sp = hy; Sup = hu; Svp = hv; } else { Sn = hy; Sun = hu; Svn = hv; }
    }
    oy[4] = ((S[1] - Sp + 4) >> 3) + d[0];
    oy[5] = ((S[2] - S[0] + 4) >> 3) + d[1];
    oy[6] = ((S[3] - S[1] + 4) >> 3) + d[2];
    oy[7] = ((Sn - S[2] + 4) >> 3) + d[3];
    ou[0] = Su[0]; ou[1] = Su[1];
    ou[2] = ((Su[1] - Sup + 4) >> 3) + du[0];
    ou[3] = ((Sun - Su[0] + 4) >> 3) + du[1];
    ov[0] = Sv[0]; ov[1] = Sv[1];
    ov[2] = ((Sv[1] - Svp + 4) >> 3) + dv[0];
    ov[3] = ((Svn - Sv[0] + 4) >> 3) + dv[1];
    if (BORDER) {
        if (L.left_border) {
            oy[4] = clamp16((-3 * S[0] + 8 * d[0] + 4 * S[1] - S[2] + 4) >> 3);
            ou[2] = clamp16((-3 * Su[0] + 8 * du[0] + 4 * Su[1] - Sun + 4) >> 3);
            ov[2] = clamp16((-3 * Sv[0] + 8 * dv[0] + 4 * Sv[1] - Svn + 4) >> 3);
        }
        if (L.right_border) {
            oy[7] = clamp16((3 * S[3] + 8 * d[3] - 4 * S[2] + S[1] + 4) >> 3);
            ou[3] = clamp16((3 * Su[1] + 8 * du[1] - 4 * Su[0] + Sup + 4) >> 3);
            ov[3] = clamp16((3 * Sv[1] + 8 * dv[1] - 4 * Sv[0] + Svp + 4) >> 3);
        }
    }
}

// ----------------------------------------------------------------------------
// 10-bit packed RGB (RG30 / AB10 / AR10 / R210 / DPX0: one 32-bit word per pixel).  The reference transforms these
// frames directly (Codec/encoder.c:3158-3176 -> wavelet.c:3597 TransformForwardSpatialRGB30 ->
// spatial.c:2080 FilterHorizontalRowRGB30_16s): the 10-bit fields are filtered after `<< (precision - 10)`, planes in
// the order G, R, B like RG48.  One launch per channel: p.pad = bit position of the channel's field, p.uyvy != 0 when
// the word is stored byte-swapped (R210, DPX0), p.shift = precision - 10.
struct RawRGB30Row {
    uint4 a, b;         // 8 pixels
    uint2 halo;         // pixels [-2,-1] (lane 0) or [+8,+9] (last lane)
};

__device__ __forceinline__ void load_rgb30_row(const unsigned char *p, const LaneInfo &L, RawRGB30Row &r)
{
    r.a = __ldg(reinterpret_cast<const uint4 *>(p));
    r.b = __ldg(reinterpret_cast<const uint4 *>(p + 16));
    r.halo = make_uint2(0u, 0u);
    if (L.use_lh | L.use_rh) r.halo = __ldg(reinterpret_cast<const uint2 *>(p + (L.use_lh ? -8 : 32)));
}

__device__ __forceinline__ unsigned rgb30_field(unsigned w, int swap, int pos, int shift)
{
    if (swap) w = __byte_perm(w, 0u, 0x0123);
    return ((w >> pos) & 0x3ffu) << shift;
}

__device__ __forceinline__ void rgb30_extract(const RawRGB30Row &r, int swap, int pos, int shift, RawPlaneRow &o)
{
    const unsigned w[8] = {r.a.x, r.a.y, r.a.z, r.a.w, r.b.x, r.b.y, r.b.z, r.b.w};
    unsigned out[4];
#pragma unroll
    for (int m = 0; m < 4; m++)
        out[m] = rgb30_field(w[2 * m], swap, pos, shift) | (rgb30_field(w[2 * m + 1], swap, pos, shift) << 16);
    o.v = make_uint4(out[0], out[1], out[2], out[3]);
    o.halo = rgb30_field(r.halo.x, swap, pos, shift) | (rgb30_field(r.halo.y, swap, pos, shift) << 16);
}

__global__ void __launch_bounds__(128) k_fwd_rgb30(const __grid_constant__ FwdParams p)
{
    const int lane = threadIdx.x;
    const int f = blockIdx.z;
    const PlaneGeom &g = p.ch[0];
    const int strip = blockIdx.x;
    if (strip * kStripIn >= g.width) return;
    const int oh = g.height >> 1;
    LaneInfo L;
    if (!lane_setup(strip, g.width, lane, L)) return;
    const unsigned colbyte = (unsigned)((strip * kStripOut + lane * 4) * 2);
    const unsigned char *in = p.in_base[f] + g.in_off + (long long)(strip * kStripIn + lane * 8) * 4;
    unsigned char *out = p.out_base[f];
    const int shift = p.shift, swap = p.uyvy, pos = p.pad;

    if (blockIdx.y == gridDim.y - 1) {
        if (threadIdx.y > 1) return;
        const bool bottom = (threadIdx.y == 1);
        const int j0 = bottom ? oh - 3 : 0;
        int s[3][8], dsel[8];
#pragma unroll
        for (int k = 0; k < 3; k++) {
            RawRGB30Row q0, q1;
            RawPlaneRow r0, r1;
            int a[8], b[8];
            load_rgb30_row(in + (long long)(2 * (j0 + k)) * g.in_pitch, L, q0);
            load_rgb30_row(in + (long long)(2 * (j0 + k) + 1) * g.in_pitch, L, q1);
            rgb30_extract(q0, swap, pos, shift, r0);
            rgb30_extract(q1, swap, pos, shift, r1);
            hfilter_plane<0>(r0, L, a);
            hfilter_plane<0>(r1, L, b);
#pragma unroll
            for (int i = 0; i < 8; i++) {
                s[k][i] = a[i] + b[i];
                if (k == (bottom ? 2 : 0)) dsel[i] = a[i] - b[i];
            }
        }
        border_emit<4>(s[0], s[1], s[2], dsel, bottom, g, out, (unsigned)((bottom ? oh - 1 : 0) * g.out_pitch) + colbyte);
        return;
    }

    const int y0 = (blockIdx.y * blockDim.y + threadIdx.y) * p.th;
    if (y0 >= oh) return;
    const int y1 = min(y0 + p.th, oh);
    const int jfirst = max(y0 - 1, 0), jlast = min(y1, oh - 1);
    const int hlo = max(y0, 1);
    VState<4> st;
#pragma unroll
    for (int i = 0; i < 8; i++) { st.llp[i] = st.llc[i] = st.dc[i] = 0; }
    const unsigned char *rp = in + (long long)(2 * jfirst) * g.in_pitch;
    RawRGB30Row c0, c1, n0, n1;
    load_rgb30_row(rp, L, c0);
    load_rgb30_row(rp + g.in_pitch, L, c1);
    n0 = c0; n1 = c1;
    unsigned off = (unsigned)(jfirst * g.out_pitch) + colbyte;
    for (int j = jfirst; j <= jlast; j++) {
        rp += 2 * g.in_pitch;
        if (j < jlast) {
            load_rgb30_row(rp, L, n0);
            load_rgb30_row(rp + g.in_pitch, L, n1);
        }
        RawPlaneRow r0, r1;
        int a[8], b[8];
        rgb30_extract(c0, swap, pos, shift, r0);
        rgb30_extract(c1, swap, pos, shift, r1);
        hfilter_plane<0>(r0, L, a);
        hfilter_plane<0>(r1, L, b);
        vstep<4, 1>(st, a, b, g, out, off, j >= y0 && j < y1, j - 1 >= hlo);
        off += (unsigned)g.out_pitch;
        c0 = n0; c1 = n1;
    }
}

// ----------------------------------------------------------------------------
// Interlaced sources: level 1 is the frame (field) transform, Codec/wavelet.c:6076 TransformForwardFrameYUV
// (Codec/filter.c:273 FilterFrameQuant16s is the planar form of the same transform):
//   t_low = even + odd, t_high = odd - even (Codec/temporal.c:1568), then the horizontal 2-6 filter on both;
//   LL = low(t_low), LH = Q(high(t_low)), HH = Q(high(t_high)) and HL = Q'(low(t_high)) difference coded along the
//   row (Codec/spatial.c:5327: Q' uses the midpoint divisor/g without the "-1", out[i] = q[i] - q[i-1]).
// The temporal step is linear in the packed bytes, so it is applied to the dp4a sums of the two rows before the
// (non-linear) rounding of the highpass filter.  No vertical neighbourhood: no border warps, no carried state.
struct Lin422 {
    int S[4], d[4], cu[4], cv[4];
    int hy, hu, hv;     // halo sums (lane 0 / last lane of strips with a neighbour strip)
};

__device__ __forceinline__ void hlinear_422(const Raw422Row &r, const Sel422 &sel, const LaneInfo &L, Lin422 &o)
{
    const unsigned w[4] = {r.v.x, r.v.y, r.v.z, r.v.w};
#pragma unroll
    for (int k = 0; k < 4; k++) {
        o.S[k] = dp4a_us(w[k], sel.ysum, 0);
        o.d[k] = dp4a_us(w[k], sel.ydif, 0);
        o.cu[k] = dp4a_us(w[k], sel.u, 0);
        o.cv[k] = dp4a_us(w[k], sel.v, 0);
    }
    o.hy = dp4a_us(L.use_lh ? r.halo.y : r.halo.x, sel.ysum, 0);
    o.hu = dp4a_us(r.halo.y, sel.u, dp4a_us(r.halo.x, sel.u, 0));
    o.hv = dp4a_us(r.halo.y, sel.v, dp4a_us(r.halo.x, sel.v, 0));
}

// a + sgn * b on every member
__device__ __forceinline__ void lin_combine(const Lin422 &a, const Lin422 &b, int sgn, Lin422 &o)
{
#pragma unroll
    for (int k = 0; k < 4; k++) {
        o.S[k] = b.S[k] + sgn * a.S[k]; o.d[k] = b.d[k] + sgn * a.d[k];
        o.cu[k] = b.cu[k] + sgn * a.cu[k]; o.cv[k] = b.cv[k] + sgn * a.cv[k];
    }
    o.hy = b.hy + sgn * a.hy; o.hu = b.hu + sgn * a.hu; o.hv = b.hv + sgn * a.hv;
}

// horizontal 2-6 on the (already temporally combined) sums.  Y: oy[0..3] low, oy[4..7] high; U/V: o[0..1], o[2..3].
__device__ __forceinline__ void hfinish_422(const Lin422 &t, const LaneInfo &L, int *oy, int *ou, int *ov)
{
    const int *S = t.S, *d = t.d;
    const int Su[2] = {t.cu[0] + t.cu[1], t.cu[2] + t.cu[3]}, du[2] = {t.cu[0] - t.cu[1], t.cu[2] - t.cu[3]};
    const int Sv[2] = {t.cv[0] + t.cv[1], t.cv[2] + t.cv[3]}, dv[2] = {t.cv[0] - t.cv[1], t.cv[2] - t.cv[3]};
    int Sp = __shfl_up_sync(L.amask, S[3], 1), Sn = __shfl_down_sync(L.amask, S[0], 1);
    int Sup = __shfl_up_sync(L.amask, Su[1], 1), Sun = __shfl_down_sync(L.amask, Su[0], 1);
    int Svp = __shfl_up_sync(L.amask, Sv[1], 1), Svn = __shfl_down_sync(L.amask, Sv[0], 1);
    if (L.use_lh) { Sp = t.hy; Sup = t.hu; Svp = t.hv; }
    if (L.use_rh) { Sn = t.hy; Sun = t.hu; Svn = t.hv; }
#pragma unroll
    for (int k = 0; k < 4; k++) oy[k] = S[k];
    oy[4] = ((S[1] - Sp + 4) >> 3) + d[0];
    oy[5] = ((S[2] - S[0] + 4) >> 3) + d[1];
    oy[6] = ((S[3] - S[1] + 4) >> 3) + d[2];
    oy[7] = ((Sn - S[2] + 4) >> 3) + d[3];
    ou[0] = Su[0]; ou[1] = Su[1];
    ou[2] = ((Su[1] - Sup + 4) >> 3) + du[0];
    ou[3] = ((Sun - Su[0] + 4) >> 3) + du[1];
    ov[0] = Sv[0]; ov[1] = Sv[1];
    ov[2] = ((Sv[1] - Svp + 4) >> 3) + dv[0];
    ov[3] = ((Svn - Sv[0] + 4) >> 3) + dv[1];
    if (L.has_border) {
        if (L.left_border) {
            oy[4] = clamp16((-3 * S[0] + 8 * d[0] + 4 * S[1] - S[2] + 4) >> 3);
            ou[2] = clamp16((-3 * Su[0] + 8 * du[0] + 4 * Su[1] - Sun + 4) >> 3);
            ov[2] = clamp16((-3 * Sv[0] + 8 * dv[0] + 4 * Sv[1] - Svn + 4) >> 3);
        }
        if (L.right_border) {
            oy[7] = clamp16((3 * S[3] + 8 * d[3] - 4 * S[2] + S[1] + 4) >> 3);
            ou[3] = clamp16((3 * Su[1] + 8 * du[1] - 4 * Su[0] + Sup + 4) >> 3);
            ov[3] = clamp16((3 * Sv[1] + 8 * dv[1] - 4 * Sv[0] + Svp + 4) >> 3);
        }
    }
}

// quantise NC lowpass values of t_high and difference-code them along the row; prev_raw = the lowpass value of the
// column left of the strip (halo), used by lane 0 of strips > 0
template <int NC>
__device__ __forceinline__ void store_diffq(unsigned char *p, const int *v, int prev_raw, const QuantParam &q, const LaneInfo &L)
{
    int Q[NC];
#pragma unroll
    for (int i = 0; i < NC; i++) Q[i] = quant1(v[i], q) >> 16;
    int prev = __shfl_up_sync(L.amask, Q[NC - 1], 1);
    if (L.use_lh) prev = quant1(prev_raw, q) >> 16;
    if (L.left_border) prev = 0;
    int o[NC];
#pragma unroll
    for (int i = 0; i < NC; i++) { o[i] = Q[i] - prev; prev = Q[i]; }
    store_raw<NC>(p, o);
}

__global__ void __launch_bounds__(128) k_fwd_422_fields(const __grid_constant__ FwdParams p)
{
    const int lane = threadIdx.x;
    const int f = blockIdx.z;
    const PlaneGeom &gy = p.ch[0];
    const PlaneGeom &gv = p.ch[1];
    const PlaneGeom &gu = p.ch[2];
    const int strip = blockIdx.x;
    if (strip * kStripIn >= gy.width) return;
    const int oh = gy.height >> 1;
    LaneInfo L;
    if (!lane_setup(strip, gy.width, lane, L)) return;
    const unsigned colbyte_y = (unsigned)((strip * kStripOut + lane * 4) * 2);
    const unsigned colbyte_c = (unsigned)((strip * (kStripOut / 2) + lane * 2) * 2);
    const unsigned char *in = p.in_base[f] + gy.in_off + (strip * kStripIn + lane * 8) * 2;
    unsigned char *out = p.out_base[f];

    Sel422 sel;
    {
        const int m = 1 << p.shift;
        const int neg = (-m) & 0xff;
        if (!p.uyvy) { sel.ysum = m | (m << 16); sel.ydif = m | (neg << 16); sel.u = m << 8; sel.v = m << 24; }
        else { sel.ysum = (m << 8) | (m << 24); sel.ydif = (m << 8) | (neg << 24); sel.u = m; sel.v = m << 16; }
    }

    const int y0 = (blockIdx.y * blockDim.y + threadIdx.y) * p.th;
    if (y0 >= oh) return;
    const int y1 = min(y0 + p.th, oh);
    const unsigned char *rp = in + (long long)(2 * y0) * gy.in_pitch;
    Raw422Row c0, c1, n0, n1;
    load_422_row(rp, L, c0);
    load_422_row(rp + gy.in_pitch, L, c1);
    n0 = c0; n1 = c1;
    unsigned offy = (unsigned)(y0 * gy.out_pitch) + colbyte_y;
    unsigned offc = (unsigned)(y0 * gu.out_pitch) + colbyte_c;
    for (int j = y0; j < y1; j++) {
        rp += 2 * gy.in_pitch;
        if (j + 1 < y1) {
            load_422_row(rp, L, n0);
            load_422_row(rp + gy.in_pitch, L, n1);
        }
        if (j + 3 < y1) { prefetch_l2(rp + 4 * gy.in_pitch); prefetch_l2(rp + 5 * gy.in_pitch); }
        Lin422 e, o, t;
        hlinear_422(c0, sel, L, e);
        hlinear_422(c1, sel, L, o);
        int ay[8], au[4], av[4];
        lin_combine(e, o, +1, t);               // temporal lowpass: even + odd
        hfinish_422(t, L, ay, au, av);
        store_raw<4>(out + (gy.band_off[0] + offy), ay);
        store_quant<4>(out + (gy.band_off[1] + offy), ay + 4, gy.q[1]);
        store_raw<2>(out + (gu.band_off[0] + offc), au);
        store_quant<2>(out + (gu.band_off[1] + offc), au + 2, gu.q[1]);
        store_raw<2>(out + (gv.band_off[0] + offc), av);
        store_quant<2>(out + (gv.band_off[1] + offc), av + 2, gv.q[1]);
        lin_combine(e, o, -1, t);               // temporal highpass: odd - even
        hfinish_422(t, L, ay, au, av);
        store_diffq<4>(out + (gy.band_off[2] + offy), ay, t.hy, gy.q[2], L);
        store_quant<4>(out + (gy.band_off[3] + offy), ay + 4, gy.q[3]);
        store_diffq<2>(out + (gu.band_off[2] + offc), au, t.hu, gu.q[2], L);
        store_quant<2>(out + (gu.band_off[3] + offc), au + 2, gu.q[3]);
        store_diffq<2>(out + (gv.band_off[2] + offc), av, t.hv, gv.q[2], L);
        store_quant<2>(out + (gv.band_off[3] + offc), av + 2, gv.q[3]);
        offy += (unsigned)gy.out_pitch;
        offc += (unsigned)gu.out_pitch;
        c0 = n0; c1 = n1;
    }
}

// ----------------------------------------------------------------------------
// 16-bit packed 4:2:2 sources (YU64: Y0 C1 Y1 C3, 16 bits each).  The reference converts them to 10-bit planes on the
// host first (Codec/frame.c:1556 ConvertYU64ToFrame16s: `(word >> 6) & 0x03ff03ff`, convert.c:3345, then
// convert.c:14370 de-interleave: position 1 -> channel 1, position 3 -> channel 2) and runs the planar level-1 filter on
// each plane (Codec/encoder.c:3180-3193 TransformForwardSpatial -> spatial.c:10026 FilterSpatialQuant16s).  Here the
// conversion is fused into the load of the same one-pass kernel structure as k_fwd_422.
struct RawYU64Row {
    uint4 a, b;         // 8 luma + 4 + 4 chroma samples of this lane (32 bytes)
    uint4 halo;         // lane 0: previous 16 bytes; last lane: next 16 bytes
};

struct SrcYU64 {
    typedef RawYU64Row Row;
    // byte offset of global lane lg (8 luma pixels = 4 groups of Y0 C1 Y1 C3, 8 bytes each) inside a row
    static __device__ __forceinline__ long long offset(int lg) { return (long long)lg * 32; }
    static __device__ __forceinline__ void load(const unsigned char *p, int, const LaneInfo &L, Row &r) {
        r.a = __ldg(reinterpret_cast<const uint4 *>(p));
        r.b = __ldg(reinterpret_cast<const uint4 *>(p + 16));
        r.halo = make_uint4(0u, 0u, 0u, 0u);
        if (L.use_lh | L.use_rh) r.halo = __ldg(reinterpret_cast<const uint4 *>(p + (L.use_lh ? -16 : 32)));
    }
    // cu = the position-1 chroma sample, cv = the position-3 sample of every 4-sample group
    static __device__ __forceinline__ void linear(const Row &r, int shift, const LaneInfo &L, Lin422 &o) {
        const unsigned M = (0xffffu >> shift) * 0x00010001u;
        const unsigned w[8] = {r.a.x, r.a.y, r.a.z, r.a.w, r.b.x, r.b.y, r.b.z, r.b.w};
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const unsigned t0 = (w[2 * k] >> shift) & M, t1 = (w[2 * k + 1] >> shift) & M;
            const int y0 = (int)(t0 & 0xffffu), y1 = (int)(t1 & 0xffffu);
            o.S[k] = y0 + y1; o.d[k] = y0 - y1;
            o.cu[k] = (int)(t0 >> 16); o.cv[k] = (int)(t1 >> 16);
        }
        const unsigned h0 = (r.halo.x >> shift) & M, h1 = (r.halo.y >> shift) & M, h2 = (r.halo.z >> shift) & M, h3 = (r.halo.w >> shift) & M;
        o.hy = L.use_lh ? (int)((h2 & 0xffffu) + (h3 & 0xffffu)) : (int)((h0 & 0xffffu) + (h1 & 0xffffu));
        o.hu = (int)((h0 >> 16) + (h2 >> 16));
        o.hv = (int)((h1 >> 16) + (h3 >> 16));
    }
};

// 10-bit packed 4:2:2 (V210): the row is a stream of 10-bit components Cb Y Cr Y Cb Y ... packed three per 32-bit
// word at bits 0, 10, 20 (Codec/convert.c:3365 ConvertYUVRowToV210 shows the layout; rows are padded to 128 bytes).  The
// reference unpacks it on the host into 10-bit planes (Codec/encoder.c:2518-2534 ConvertV210ToFrame16s: first chroma ->
// channel 2, second chroma -> channel 1, values unshifted) and runs the planar filter per plane (encoder.c:3180-3193).
// A lane's 8 pixels are 16 consecutive components starting at component 16 * lg, i.e. at word (16 * lg) / 3 with a
// phase of lg % 3 components into that word: six words always cover them.
struct RawV210Row {
    unsigned w[6];      // words (16 * lg) / 3 ... + 5
    unsigned hw[4];     // halo: the words holding the 8 components before (lane 0) / after (last lane) this lane's
    int phase;          // (16 * lg) % 3
};

// 60 useful bits of two consecutive words
__device__ __forceinline__ unsigned long long v210_pair(unsigned a, unsigned b) {
    return (unsigned long long)(a & 0x3fffffffu) | ((unsigned long long)(b & 0x3fffffffu) << 30);
}
// drop `sh` (0, 10 or 20) bits from the front of the 60-bit pair lo, refilling from the next pair hi
__device__ __forceinline__ unsigned long long v210_shift(unsigned long long lo, unsigned long long hi, int sh) {
    const unsigned long long m60 = (1ull << 60) - 1;
    return sh ? (((lo >> sh) | (hi << (60 - sh))) & m60) : lo;
}
__device__ __forceinline__ int v210_field(unsigned long long x, int i) { return (int)((x >> (10 * i)) & 0x3ffu); }

struct SrcV210 {
    typedef RawV210Row Row;
    static __device__ __forceinline__ long long offset(int lg) { return (long long)((16 * lg) / 3) * 4; }
    static __device__ __forceinline__ void load(const unsigned char *p, int lg, const LaneInfo &L, Row &r) {
        const unsigned *wp = reinterpret_cast<const unsigned *>(p);
#pragma unroll
        for (int i = 0; i < 6; i++) r.w[i] = __ldg(wp + i);
        r.phase = (16 * lg) % 3;
#pragma unroll
        for (int i = 0; i < 4; i++) r.hw[i] = 0u;
        if (L.use_lh | L.use_rh) {
            // the 8 components before this lane start at component 16 lg - 8, the 8 after it at 16 lg + 16
            const int c0 = 16 * lg + (L.use_lh ? -8 : 16);
            const unsigned *hp = wp + (c0 / 3 - (16 * lg) / 3);
#pragma unroll
            for (int i = 0; i < 4; i++) r.hw[i] = __ldg(hp + i);
        }
    }
    // cu = the chroma that goes to channel 1 (the SECOND chroma component, Cr), cv = the one for channel 2 (Cb)
    static __device__ __forceinline__ void linear(const Row &r, int, const LaneInfo &L, Lin422 &o) {
        const int sh = 10 * r.phase;
        const unsigned long long A = v210_pair(r.w[0], r.w[1]), B = v210_pair(r.w[2], r.w[3]), C = v210_pair(r.w[4], r.w[5]);
        const unsigned long long a = v210_shift(A, B, sh), b = v210_shift(B, C, sh), c = v210_shift(C, 0ull, sh);
        int comp[16];
#pragma unroll
        for (int i = 0; i < 6; i++) { comp[i] = v210_field(a, i); comp[6 + i] = v210_field(b, i); }
#pragma unroll
        for (int i = 0; i < 4; i++) comp[12 + i] = v210_field(c, i);
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int y0 = comp[4 * k + 1], y1 = comp[4 * k + 3];
            o.S[k] = y0 + y1; o.d[k] = y0 - y1;
            o.cv[k] = comp[4 * k]; o.cu[k] = comp[4 * k + 2];
        }
        // halo: 8 components starting at phase (phase + 1) % 3 of hw[0] (16 lg - 8 and 16 lg + 16 are both = 16 lg + 1 mod 3)
        const int hsh = 10 * ((r.phase + 1) % 3);
        const unsigned long long HA = v210_pair(r.hw[0], r.hw[1]), HB = v210_pair(r.hw[2], r.hw[3]);
        const unsigned long long ha = v210_shift(HA, HB, hsh), hb = v210_shift(HB, 0ull, hsh);
        int hc[8];
#pragma unroll
        for (int i = 0; i < 6; i++) hc[i] = v210_field(ha, i);
        hc[6] = v210_field(hb, 0); hc[7] = v210_field(hb, 1);
        // components: Cb Y Cr Y | Cb Y Cr Y ; the luma pair adjacent to this lane is the second group on the left
        // side and the first group on the right side
        o.hy = L.use_lh ? (hc[5] + hc[7]) : (hc[1] + hc[3]);
        o.hv = hc[0] + hc[4];
        o.hu = hc[2] + hc[6];
    }
};

// Generic one-pass level 1 of a packed 4:2:2 source: SRC supplies the row load and the linear (pre-rounding) sums.
// p.ch[0] = luma, p.ch[1] receives the position-1 chroma, p.ch[2] the position-3 chroma.
template <class SRC>
__global__ void __launch_bounds__(128) k_fwd_422_src(const __grid_constant__ FwdParams p)
{
    const int lane = threadIdx.x;
    const int f = blockIdx.z;
    const PlaneGeom &gy = p.ch[0];
    const PlaneGeom &g1 = p.ch[1];
    const PlaneGeom &g3 = p.ch[2];
    const int strip = blockIdx.x;
    if (strip * kStripIn >= gy.width) return;
    const int oh = gy.height >> 1;
    LaneInfo L;
    if (!lane_setup(strip, gy.width, lane, L)) return;
    const unsigned colbyte_y = (unsigned)((strip * kStripOut + lane * 4) * 2);
    const unsigned colbyte_c = (unsigned)((strip * (kStripOut / 2) + lane * 2) * 2);
    const int lg = strip * 32 + lane;           // global lane index: 8 luma pixels each
    const unsigned char *in = p.in_base[f] + gy.in_off + SRC::offset(lg);
    unsigned char *out = p.out_base[f];
    const int shift = p.shift;

    if (blockIdx.y == gridDim.y - 1) {      // border warps: first / last HL,HH row of all three channels
        if (threadIdx.y > 1) return;
        const bool bottom = (threadIdx.y == 1);
        const int j0 = bottom ? oh - 3 : 0;
        int sy[3][8], s1[3][4], s3[3][4], dy[8], d1[4], d3[4];
#pragma unroll
        for (int k = 0; k < 3; k++) {
            typename SRC::Row r0, r1;
            Lin422 t;
            int ay[8], by[8], a1[4], b1[4], a3[4], b3[4];
            SRC::load(in + (long long)(2 * (j0 + k)) * gy.in_pitch, lg, L, r0);
            SRC::load(in + (long long)(2 * (j0 + k) + 1) * gy.in_pitch, lg, L, r1);
            SRC::linear(r0, shift, L, t); hfinish_422(t, L, ay, a1, a3);
            SRC::linear(r1, shift, L, t); hfinish_422(t, L, by, b1, b3);
            const bool keep = (k == (bottom ? 2 : 0));
#pragma unroll
            for (int i = 0; i < 8; i++) { sy[k][i] = ay[i] + by[i]; if (keep) dy[i] = ay[i] - by[i]; }
#pragma unroll
            for (int i = 0; i < 4; i++) {
                s1[k][i] = a1[i] + b1[i]; s3[k][i] = a3[i] + b3[i];
                if (keep) { d1[i] = a1[i] - b1[i]; d3[i] = a3[i] - b3[i]; }
            }
        }
        const int row = bottom ? oh - 1 : 0;
        border_emit<4>(sy[0], sy[1], sy[2], dy, bottom, gy, out, (unsigned)(row * gy.out_pitch) + colbyte_y);
        border_emit<2>(s1[0], s1[1], s1[2], d1, bottom, g1, out, (unsigned)(row * g1.out_pitch) + colbyte_c);
        border_emit<2>(s3[0], s3[1], s3[2], d3, bottom, g3, out, (unsigned)(row * g3.out_pitch) + colbyte_c);
        return;
    }

    const int y0 = (blockIdx.y * blockDim.y + threadIdx.y) * p.th;
    if (y0 >= oh) return;
    const int y1 = min(y0 + p.th, oh);
    const int jfirst = max(y0 - 1, 0), jlast = min(y1, oh - 1);
    const int hlo = max(y0, 1);
    VState<4> sy;
    VState<2> s1, s3;
#pragma unroll
    for (int i = 0; i < 8; i++) { sy.llp[i] = sy.llc[i] = sy.dc[i] = 0; }
#pragma unroll
    for (int i = 0; i < 4; i++) { s1.llp[i] = s1.llc[i] = s1.dc[i] = 0; s3.llp[i] = s3.llc[i] = s3.dc[i] = 0; }
    const unsigned char *rp = in + (long long)(2 * jfirst) * gy.in_pitch;
    typename SRC::Row c0, c1, n0, n1;
    SRC::load(rp, lg, L, c0);
    SRC::load(rp + gy.in_pitch, lg, L, c1);
    n0 = c0; n1 = c1;
    unsigned offy = (unsigned)(jfirst * gy.out_pitch) + colbyte_y;
    unsigned offc = (unsigned)(jfirst * g1.out_pitch) + colbyte_c;
    for (int j = jfirst; j <= jlast; j++) {
        rp += 2 * gy.in_pitch;
        if (j < jlast) {
            SRC::load(rp, lg, L, n0);
            SRC::load(rp + gy.in_pitch, lg, L, n1);
        }
        Lin422 t;
        int ay[8], by[8], a1[4], b1[4], a3[4], b3[4];
        SRC::linear(c0, shift, L, t); hfinish_422(t, L, ay, a1, a3);
        SRC::linear(c1, shift, L, t); hfinish_422(t, L, by, b1, b3);
        const bool emit_low = (j >= y0) && (j < y1), emit_high = (j - 1 >= hlo);
        vstep<4, 1>(sy, ay, by, gy, out, offy, emit_low, emit_high);
        vstep<2, 1>(s1, a1, b1, g1, out, offc, emit_low, emit_high);
        vstep<2, 1>(s3, a3, b3, g3, out, offc, emit_low, emit_high);
        offy += (unsigned)gy.out_pitch;
        offc += (unsigned)g1.out_pitch;
        c0 = n0; c1 = n1;
    }
}

// Interlaced (field) level 1 of the packed 16-bit / 10-bit 4:2:2 sources: the reference converts them to planes and runs
//   Codec/filter.c:273 FilterFrameQuant16s: temporal.c FilterTemporalRow16s (even + odd, odd - even), then
//   spatial.c:5826 FilterHorizontalRowQuant16s on the temporal lowpass -- LL (quantised only when its divisor > 1) and LH,
//   both with the midpoint divisor / 2 (filter.c:352 / spatial.c:5856; the packed 8-bit path rounds LH with
//   divisor / 2 - 1) in the columns its 16-sample SSE2 loop produces and WITHOUT a midpoint in the columns of its scalar
//   tail and in the last column, which it redoes with the border filter (spatial.c:6192-6266) -- and spatial.c:5327
//   ...DifferenceFiltered + QuantizeRow16sTo16s on the temporal highpass, as the packed path.  Same structure as
//   k_fwd_422_fields; SRC supplies the row load and the linear sums.  (LL is quantised by the same routine when its divisor
//   exceeds 1, which no schedule of the reference produces at level 1: the host side rejects such a table.)
//   p.ch[1] receives the position-1 chroma, p.ch[2] the position-3 chroma (as k_fwd_422_src).
// LH of the planar field transform: per column, midpoint divisor / 2 or none (see k_fwd_422_fields_src)
template <int NC>
__device__ __forceinline__ void store_quant_lh_planar(unsigned char *ptr, const int *v, const QuantParam &q, int col0, int width_out)
{
    // columns >= tail belong to the scalar tail of a (2 * width_out)-sample row; the last column is the border column
    const int tail = (2 * width_out - (2 * width_out) % 16) / 2;
    int o[NC];
#pragma unroll
    for (int i = 0; i < NC; i++) {
        const int col = col0 + i;
        const bool nomid = (col >= tail) || (col == width_out - 1);
        o[i] = (v[i] * q.m + (v[i] < 0 ? (nomid ? 65535 : q.cneg) : (nomid ? 0 : q.cpos))) >> 16;
    }
    store_raw<NC>(ptr, o);
}

template <class SRC>
__global__ void __launch_bounds__(128) k_fwd_422_fields_src(const __grid_constant__ FwdParams p)
{
    const int lane = threadIdx.x;
    const int f = blockIdx.z;
    const PlaneGeom &gy = p.ch[0];
    const PlaneGeom &g1 = p.ch[1];
    const PlaneGeom &g3 = p.ch[2];
    const int strip = blockIdx.x;
    if (strip * kStripIn >= gy.width) return;
    const int oh = gy.height >> 1;
    LaneInfo L;
    if (!lane_setup(strip, gy.width, lane, L)) return;
    const unsigned colbyte_y = (unsigned)((strip * kStripOut + lane * 4) * 2);
    const unsigned colbyte_c = (unsigned)((strip * (kStripOut / 2) + lane * 2) * 2);
    const int lg = strip * 32 + lane;
    const unsigned char *in = p.in_base[f] + gy.in_off + SRC::offset(lg);
    unsigned char *out = p.out_base[f];
    const int shift = p.shift;
    const int y0 = (blockIdx.y * blockDim.y + threadIdx.y) * p.th;
    if (y0 >= oh) return;
    const int y1 = min(y0 + p.th, oh);
    const unsigned char *rp = in + (long long)(2 * y0) * gy.in_pitch;
    typename SRC::Row c0, c1, n0, n1;
    SRC::load(rp, lg, L, c0);
    SRC::load(rp + gy.in_pitch, lg, L, c1);
    n0 = c0; n1 = c1;
    unsigned offy = (unsigned)(y0 * gy.out_pitch) + colbyte_y;
    unsigned offc = (unsigned)(y0 * g1.out_pitch) + colbyte_c;
    for (int j = y0; j < y1; j++) {
        rp += 2 * gy.in_pitch;
        if (j + 1 < y1) {
            SRC::load(rp, lg, L, n0);
            SRC::load(rp + gy.in_pitch, lg, L, n1);
        }
        Lin422 e, o, t;
        SRC::linear(c0, shift, L, e);
        SRC::linear(c1, shift, L, o);
        int ay[8], a1[4], a3[4];
        lin_combine(e, o, +1, t);               // temporal lowpass: even + odd
        hfinish_422(t, L, ay, a1, a3);
        store_raw<4>(out + (gy.band_off[0] + offy), ay);
        store_quant_lh_planar<4>(out + (gy.band_off[1] + offy), ay + 4, gy.q[1], strip * kStripOut + lane * 4, gy.width >> 1);
        store_raw<2>(out + (g1.band_off[0] + offc), a1);
        store_quant_lh_planar<2>(out + (g1.band_off[1] + offc), a1 + 2, g1.q[1], strip * (kStripOut / 2) + lane * 2, g1.width >> 1);
        store_raw<2>(out + (g3.band_off[0] + offc), a3);
        store_quant_lh_planar<2>(out + (g3.band_off[1] + offc), a3 + 2, g3.q[1], strip * (kStripOut / 2) + lane * 2, g3.width >> 1);
        lin_combine(e, o, -1, t);               // temporal highpass: odd - even
        hfinish_422(t, L, ay, a1, a3);
        store_diffq<4>(out + (gy.band_off[2] + offy), ay, t.hy, gy.q[2], L);
        store_quant<4>(out + (gy.band_off[3] + offy), ay + 4, gy.q[3]);
        store_diffq<2>(out + (g1.band_off[2] + offc), a1, t.hu, g1.q[2], L);
        store_quant<2>(out + (g1.band_off[3] + offc), a1 + 2, g1.q[3]);
        store_diffq<2>(out + (g3.band_off[2] + offc), a3, t.hv, g3.q[2], L);
        store_quant<2>(out + (g3.band_off[3] + offc), a3 + 2, g3.q[3]);
        offy += (unsigned)gy.out_pitch;
        offc += (unsigned)g1.out_pitch;
        c0 = n0; c1 = n1;
    }
}

#include "cfb_forward_tma.inl"

// ----------------------------------------------------------------------------
// host-side launchers (called from cfb_api.cu).  gridDim.y = row blocks + 1 border CTA row.
static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// A/B switch (profiles/r02_ab_fwdplane.txt): CFB_FWDPLANE = r1 forces the round-1 kernels (direct LDG into registers)
// everywhere, = tma the TMA-fed kernel everywhere.  Default: TMA where several channels share one read of the source
// (RG48: 294 -> 174 us per 8 4K frames, BYR4), round-1 kernels for single planes (levels 2 and 3: 66.5 / 22.1 us against
// 74.0 / 28.9 us with the TMA ring, whose start-up is not amortised over the 8-16 row pairs of a CTA).
static int fwdplane_variant()
{
    static int v = -1;
    if (v < 0) { const char *e = getenv("CFB_FWDPLANE"); v = !e ? 0 : !strcmp(e, "r1") ? 1 : !strcmp(e, "tma") ? 2 : 0; }
    return v;
}

cudaError_t launch_fwd_plane(const FwdParams &p, int prescale, cudaStream_t stream)
{
    int maxw = 0, maxoh = 0;
    bool ragged = false, tma_ok = (fwdplane_variant() == 2) && (p.nframes * p.nchan <= kMaxBatch * kMaxChannels);
    for (int c = 0; c < p.nchan; c++) {
        maxw = max(maxw, p.ch[c].width); maxoh = max(maxoh, p.ch[c].height / 2);
        ragged = ragged || (p.ch[c].width & 7);
        tma_ok = tma_ok && !(p.ch[c].in_pitch & 15) && !(p.ch[c].in_off & 15);
    }
    for (int i = 0; i < p.nframes; i++) tma_ok = tma_ok && !((uintptr_t)p.in_base[i] & 15);
    if (ragged) {       // the 1-3 output columns right of the last full lane (they include the right border)
        dim3 eblock(128), egrid(ceil_div(maxoh, 128), 3, p.nframes * p.nchan);
        if (prescale) k_fwd_plane_edge<2><<<egrid, eblock, 0, stream>>>(p); else k_fwd_plane_edge<0><<<egrid, eblock, 0, stream>>>(p);
    }
    dim3 block(32, 4);
    // p.pad != 0: the caller vouches that the planes are non-negative (LL bands of an unsigned source); CFB_FWDPLANE_NN=0
    // keeps the generic prescaled kernel for the A/B
    static const bool nn_on = !(getenv("CFB_FWDPLANE_NN") && !strcmp(getenv("CFB_FWDPLANE_NN"), "0"));
    const bool nonneg = prescale && p.pad && nn_on;
    if (!tma_ok) {      // round-1 path (also: plane pointers / pitches that are not 16-byte aligned cannot be described to the TMA)
        dim3 grid(ceil_div(maxw, kStripIn), ceil_div(ceil_div(maxoh, p.th), (int)block.y) + 1, p.nframes * p.nchan);
        if (nonneg) k_fwd_plane<3><<<grid, block, 0, stream>>>(p);
        else if (prescale) k_fwd_plane<2><<<grid, block, 0, stream>>>(p);
        else k_fwd_plane<0><<<grid, block, 0, stream>>>(p);
        return cudaGetLastError();
    }
    FwdTmaPlaneMaps tm;
    for (int i = 0; i < p.nframes; i++)
        for (int c = 0; c < p.nchan; c++) {
            const PlaneGeom &g = p.ch[c];
            cudaError_t e = tmap_encode_2d(&tm.in_map[i * p.nchan + c], p.in_base[i] + g.in_off, (uint64_t)g.width * 2, (uint64_t)g.height,
                                           (uint64_t)g.in_pitch, SrcPlane16<0>::kRowBytes, 2);
            if (e != cudaSuccess) return e;
        }
    dim3 tgrid(ceil_div(maxw, kStripIn), ceil_div(maxoh, p.th), p.nframes * p.nchan), tblock(32, 1);
    const size_t smem = kTmaStages * SrcPlane16<0>::kStageBytes + 2 * kTmaStages * 8;
    if (prescale) k_fwd_tma<SrcPlane16<2>, 8><<<tgrid, tblock, smem, stream>>>(p, tm);
    else k_fwd_tma<SrcPlane16<0>, 8><<<tgrid, tblock, smem, stream>>>(p, tm);
    // first / last HL,HH row: the border CTA row of the round-1 kernel, alone
    dim3 bgrid(ceil_div(maxw, kStripIn), 1, p.nframes * p.nchan);
    if (prescale) k_fwd_plane<2><<<bgrid, block, 0, stream>>>(p);
    else k_fwd_plane<0><<<bgrid, block, 0, stream>>>(p);
    return cudaGetLastError();
}

// all three channels of packed RG48 frames from ONE read of the pixel groups: p.ch[0..2] = G, R, B
cudaError_t launch_fwd_rg48_all(const FwdParams &p, cudaStream_t stream)
{
    FwdTmaPlaneMaps tm;
    const PlaneGeom &g = p.ch[0];
    for (int i = 0; i < p.nframes; i++) {
        cudaError_t e = tmap_encode_2d(&tm.in_map[i], p.in_base[i] + g.in_off, (uint64_t)g.width * 6, (uint64_t)g.height, (uint64_t)g.in_pitch,
                                       SrcRG48::kRowBytes, 2, 8);
        if (e != cudaSuccess) return e;
    }
    dim3 tgrid(ceil_div(g.width, kStripIn), ceil_div(g.height / 2, p.th), p.nframes), tblock(32, 3);
    k_fwd_tma<SrcRG48, 5><<<tgrid, tblock, kTmaStages * SrcRG48::kStageBytes + 2 * kTmaStages * 8, stream>>>(p, tm);
    // border rows per channel (round-1 kernel on its border CTA row): p.ch[0] must describe the channel
    static const int sel_of_channel[3] = {1, 0, 2};
    dim3 block(32, 4), bgrid(ceil_div(g.width, kStripIn), 1, p.nframes);
    for (int c = 0; c < 3; c++) {
        FwdParams q = p;
        q.nchan = 1; q.ch[0] = p.ch[c];
        if (sel_of_channel[c] == 0) k_fwd_rg48<0><<<bgrid, block, 0, stream>>>(q);
        else if (sel_of_channel[c] == 1) k_fwd_rg48<1><<<bgrid, block, 0, stream>>>(q);
        else k_fwd_rg48<2><<<bgrid, block, 0, stream>>>(q);
    }
    return cudaGetLastError();
}

// sel: word of each RGB pixel feeding this channel (0 = R, 1 = G, 2 = B); p.ch[0] describes the channel
cudaError_t launch_fwd_rg48(const FwdParams &p, int sel, cudaStream_t stream)
{
    dim3 block(32, 4);
    dim3 grid(ceil_div(p.ch[0].width, kStripIn), ceil_div(ceil_div(p.ch[0].height / 2, p.th), (int)block.y) + 1, p.nframes);
    if (sel == 0) k_fwd_rg48<0><<<grid, block, 0, stream>>>(p);
    else if (sel == 1) k_fwd_rg48<1><<<grid, block, 0, stream>>>(p);
    else k_fwd_rg48<2><<<grid, block, 0, stream>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_fwd_rgb30(const FwdParams &p, cudaStream_t stream)
{
    dim3 block(32, 4);
    dim3 grid(ceil_div(p.ch[0].width, kStripIn), ceil_div(ceil_div(p.ch[0].height / 2, p.th), (int)block.y) + 1, p.nframes);
    k_fwd_rgb30<<<grid, block, 0, stream>>>(p);
    return cudaGetLastError();
}

// all four Bayer-derived channels; p.uyvy carries the Bayer phase
cudaError_t launch_fwd_byr4(const FwdParams &p, cudaStream_t stream)
{
    dim3 block(32, 4);
    const PlaneGeom &g = p.ch[0];
    const bool tma_ok = (fwdplane_variant() != 1) && !(g.in_pitch & 15);
    if (!tma_ok) {
        dim3 grid(ceil_div(g.width, kStripIn) * 4, ceil_div(ceil_div(g.height / 2, p.th), (int)block.y) + 1, p.nframes);
        if (p.lut) k_fwd_byr4<true><<<grid, block, 0, stream>>>(p); else k_fwd_byr4<false><<<grid, block, 0, stream>>>(p);
        return cudaGetLastError();
    }
    // one read of the Bayer lines feeds the four channel warps of a CTA (plane width = half the Bayer width)
    FwdTmaPlaneMaps tm;
    for (int i = 0; i < p.nframes; i++) {
        cudaError_t e = tmap_encode_2d(&tm.in_map[i], p.in_base[i], (uint64_t)g.width * 4, (uint64_t)g.height * 2, (uint64_t)g.in_pitch,
                                       SrcBYR4<false>::kRowBytes, 4, 8);
        if (e != cudaSuccess) return e;
    }
    dim3 tgrid(ceil_div(g.width, kStripIn), ceil_div(g.height / 2, p.th), p.nframes), tblock(32, 4);
    const size_t smem = kTmaStages * SrcBYR4<false>::kStageBytes + 2 * kTmaStages * 8;
    if (p.lut) k_fwd_tma<SrcBYR4<true>, 3><<<tgrid, tblock, smem, stream>>>(p, tm);
    else k_fwd_tma<SrcBYR4<false>, 3><<<tgrid, tblock, smem, stream>>>(p, tm);
    dim3 bgrid(ceil_div(g.width, kStripIn) * 4, 1, p.nframes);
    if (p.lut) k_fwd_byr4<true><<<bgrid, block, 0, stream>>>(p); else k_fwd_byr4<false><<<bgrid, block, 0, stream>>>(p);
    return cudaGetLastError();
}

// CFB_FWD422 = r1 selects the first-generation kernel (direct LDG into registers; kept for the A/B evidence in profiles/)
static int fwd422_variant()
{
    static int v = -1;
    if (v < 0) { const char *e = getenv("CFB_FWD422"); v = (e && !strcmp(e, "r1")) ? 1 : 0; }
    return v;
}

cudaError_t launch_fwd_422(const FwdParams &p, cudaStream_t stream)
{
    dim3 block(32, 4);
    dim3 grid(ceil_div(p.ch[0].width, kStripIn), ceil_div(ceil_div(p.ch[0].height / 2, p.th), (int)block.y) + 1, p.nframes);
    if (fwd422_variant() == 1) { k_fwd_422<<<grid, block, 0, stream>>>(p); return cudaGetLastError(); }
    // one tensor map per frame of the batch: rows of 2 * width bytes, `height` rows, the caller's pitch
    FwdTmaMaps tm;
    for (int i = 0; i < p.nframes; i++) {
        cudaError_t e = tmap_encode_2d(&tm.in_map[i], p.in_base[i] + p.ch[0].in_off, (uint64_t)p.ch[0].width * 2, (uint64_t)p.ch[0].height,
                                       (uint64_t)p.ch[0].in_pitch, kTmaRowBytes, 2);
        if (e != cudaSuccess) return e;
    }
    k_fwd_422_tma<3><<<grid, block, 4 * kTmaWarpBytes + 4 * kTmaStages * 8, stream>>>(p, tm);
    return cudaGetLastError();
}

cudaError_t launch_fwd_yu64(const FwdParams &p, cudaStream_t stream)
{
    dim3 block(32, 4);
    dim3 grid(ceil_div(p.ch[0].width, kStripIn), ceil_div(ceil_div(p.ch[0].height / 2, p.th), (int)block.y) + 1, p.nframes);
    k_fwd_422_src<SrcYU64><<<grid, block, 0, stream>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_fwd_v210(const FwdParams &p, cudaStream_t stream)
{
    dim3 block(32, 4);
    dim3 grid(ceil_div(p.ch[0].width, kStripIn), ceil_div(ceil_div(p.ch[0].height / 2, p.th), (int)block.y) + 1, p.nframes);
    k_fwd_422_src<SrcV210><<<grid, block, 0, stream>>>(p);
    return cudaGetLastError();
}

// sel: 0 = YU64, 1 = V210
cudaError_t launch_fwd_422_fields_src(const FwdParams &p, int sel, cudaStream_t stream)
{
    dim3 block(32, 4);
    dim3 grid(ceil_div(p.ch[0].width, kStripIn), ceil_div(ceil_div(p.ch[0].height / 2, p.th), (int)block.y), p.nframes);
    if (sel) k_fwd_422_fields_src<SrcV210><<<grid, block, 0, stream>>>(p); else k_fwd_422_fields_src<SrcYU64><<<grid, block, 0, stream>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_fwd_422_fields(const FwdParams &p, cudaStream_t stream)
{
    dim3 block(32, 4);
    dim3 grid(ceil_div(p.ch[0].width, kStripIn), ceil_div(ceil_div(p.ch[0].height / 2, p.th), (int)block.y), p.nframes);
    k_fwd_422_fields<<<grid, block, 0, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace cfb
