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
// Every input sample is read from global memory once (+ 2 halo row pairs per strip
// block, served by L2) and every coefficient is written once with 64-bit stores.
#include "cfb_common.cuh"

namespace cfb {

constexpr unsigned kFull = 0xffffffffu;

// ----------------------------------------------------------------------------
// vertical state of NC low + NC high columns held by one lane
template <int NC>
struct VState {
    int llp[2 * NC];    // S_{j-2}
    int llc[2 * NC];    // S_{j-1}
    int dc[2 * NC];     // D_{j-1}
    int d0[2 * NC];     // D_0 (top border only)
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

// One vertical step: rows 2j (a) and 2j+1 (b) of the horizontal outputs, [0,NC) = low, [NC,2NC) = high.
template <int NC>
__device__ __forceinline__ void vstep(VState<NC> &s, int j, const int *a, const int *b, int y0, int y1, int oh,
                                      const PlaneGeom &g, unsigned char *out, int colbyte, bool active)
{
    int v[2 * NC], dn[2 * NC];
#pragma unroll
    for (int i = 0; i < 2 * NC; i++) { v[i] = a[i] + b[i]; dn[i] = a[i] - b[i]; }

    if (active && j >= y0 && j < y1) {
        unsigned char *pll = out + g.band_off[0] + (long long)j * g.out_pitch + colbyte;
        unsigned char *plh = out + g.band_off[1] + (long long)j * g.out_pitch + colbyte;
        if (g.quant_ll) store_quant<NC>(pll, v, g.q[0]); else store_raw<NC>(pll, v);
        store_quant<NC>(plh, v + NC, g.q[1]);
    }
    const int r = j - 1;
    if (active && r >= y0 && r < y1 && r >= 1) {
        int h[2 * NC];
#pragma unroll
        for (int i = 0; i < 2 * NC; i++) h[i] = ((v[i] - s.llp[i] + 4) >> 3) + s.dc[i];
        store_quant<NC>(out + g.band_off[2] + (long long)r * g.out_pitch + colbyte, h, g.q[2]);
        store_quant<NC>(out + g.band_off[3] + (long long)r * g.out_pitch + colbyte, h + NC, g.q[3]);
    }
    if (j == 2 && y0 == 0 && active) {          // top border row (spatial.c:10166-10208)
        int h[2 * NC];
#pragma unroll
        for (int i = 0; i < 2 * NC; i++)
            h[i] = clamp16((-3 * s.llp[i] + 8 * s.d0[i] + 4 * s.llc[i] - v[i] + 4) >> 3);
        store_quant<NC>(out + g.band_off[2] + colbyte, h, g.q[2]);
        store_quant<NC>(out + g.band_off[3] + colbyte, h + NC, g.q[3]);
    }
    if (j == oh - 1 && y1 == oh && active) {    // bottom border row (spatial.c:10516-10558)
        int h[2 * NC];
#pragma unroll
        for (int i = 0; i < 2 * NC; i++)
            h[i] = clamp16((3 * v[i] + 8 * dn[i] - 4 * s.llc[i] + s.llp[i] + 4) >> 3);
        store_quant<NC>(out + g.band_off[2] + (long long)(oh - 1) * g.out_pitch + colbyte, h, g.q[2]);
        store_quant<NC>(out + g.band_off[3] + (long long)(oh - 1) * g.out_pitch + colbyte, h + NC, g.q[3]);
    }
#pragma unroll
    for (int i = 0; i < 2 * NC; i++) {
        s.llp[i] = s.llc[i]; s.llc[i] = v[i]; s.dc[i] = dn[i];
        if (j == 0) s.d0[i] = dn[i];
    }
}

// rows of pairs [jb, je] are needed to emit output rows [y0, y1)
__device__ __forceinline__ void pair_range(int y0, int y1, int oh, int &jb, int &je) {
    jb = max(y0 - 1, 0);
    je = min(y1, oh - 1);
    if (y1 == oh) jb = min(jb, max(oh - 3, 0));
    if (y0 == 0) je = max(je, min(2, oh - 1));
}

// ----------------------------------------------------------------------------
// int16 plane input
struct RawPlaneRow {
    uint4 v;        // 8 samples of this lane
    unsigned lh;    // samples [-2,-1] of the strip (lane 0 only)
    unsigned rh;    // samples [256,257] of the strip (lane 31 only)
};

__device__ __forceinline__ void load_plane_row(const unsigned char *in, int pitch, int row, int col0, bool active,
                                               bool use_lh, bool use_rh, RawPlaneRow &r)
{
    const unsigned char *p = in + (long long)row * pitch + (long long)col0 * 2;
    r.v = active ? __ldg(reinterpret_cast<const uint4 *>(p)) : make_uint4(0, 0, 0, 0);
    r.lh = use_lh ? __ldg(reinterpret_cast<const unsigned *>(p - 4)) : 0u;
    r.rh = use_rh ? __ldg(reinterpret_cast<const unsigned *>(p + 16)) : 0u;
}

template <int PRESCALE>
__device__ __forceinline__ int tap(int x) { return PRESCALE ? ((x + 3) >> 2) : x; }

// horizontal 2-6 for the lane's 4 output columns: o[0..3] = low, o[4..7] = high
// (Codec/spatial.c:253 FilterHorizontalRow16s / :3669 FilterHorizontalRow10bit16s)
template <int PRESCALE>
__device__ __forceinline__ void hfilter_plane(const RawPlaneRow &r, bool left_border, bool right_border,
                                              bool use_lh, bool use_rh, int *o)
{
    const unsigned w[4] = {r.v.x, r.v.y, r.v.z, r.v.w};
    int S[4], d[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int x0 = lo16(w[k]), x1 = hi16(w[k]);
        const int t0 = tap<PRESCALE>(x0), t1 = tap<PRESCALE>(x1);
        S[k] = t0 + t1;
        d[k] = t0 - t1;
        o[k] = PRESCALE ? ((x0 + x1 + 3) >> 2) : S[k];
    }
    int Sp = __shfl_up_sync(kFull, S[3], 1);
    int Sn = __shfl_down_sync(kFull, S[0], 1);
    if (use_lh) Sp = tap<PRESCALE>(lo16(r.lh)) + tap<PRESCALE>(hi16(r.lh));
    if (use_rh) Sn = tap<PRESCALE>(lo16(r.rh)) + tap<PRESCALE>(hi16(r.rh));
    o[4] = ((S[1] - Sp + 4) >> 3) + d[0];
    o[5] = ((S[2] - S[0] + 4) >> 3) + d[1];
    o[6] = ((S[3] - S[1] + 4) >> 3) + d[2];
    o[7] = ((Sn - S[2] + 4) >> 3) + d[3];
    if (left_border) o[4] = clamp16((-3 * S[0] + 8 * d[0] + 4 * S[1] - S[2] + 4) >> 3);
    if (right_border) o[7] = clamp16((3 * S[3] + 8 * d[3] - 4 * S[2] + S[1] + 4) >> 3);
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
    const int y0 = (blockIdx.y * blockDim.y + threadIdx.y) * p.th;
    if (y0 >= oh) return;
    const int y1 = min(y0 + p.th, oh);

    const int col0 = strip * kStripIn + lane * 8;
    const bool active = col0 < g.width;
    const bool left_border = (col0 == 0);
    const bool right_border = (col0 + 8 == g.width);
    const bool use_lh = (lane == 0) && (strip > 0);
    const bool use_rh = (lane == 31) && (col0 + 8 < g.width);
    const unsigned char *in = p.in_base[f] + g.in_off;
    unsigned char *out = p.out_base[f];
    const int colbyte = (strip * kStripOut + lane * 4) * 2;

    int jb, je;
    pair_range(y0, y1, oh, jb, je);

    VState<4> st;
#pragma unroll
    for (int i = 0; i < 8; i++) { st.llp[i] = st.llc[i] = st.dc[i] = st.d0[i] = 0; }

    RawPlaneRow c0, c1, n0, n1;
    load_plane_row(in, g.in_pitch, 2 * jb, col0, active, use_lh, use_rh, c0);
    load_plane_row(in, g.in_pitch, 2 * jb + 1, col0, active, use_lh, use_rh, c1);
    for (int j = jb; j <= je; j++) {
        if (j < je) {
            load_plane_row(in, g.in_pitch, 2 * j + 2, col0, active, use_lh, use_rh, n0);
            load_plane_row(in, g.in_pitch, 2 * j + 3, col0, active, use_lh, use_rh, n1);
        }
        int a[8], b[8];
        hfilter_plane<PRESCALE>(c0, left_border, right_border, use_lh, use_rh, a);
        hfilter_plane<PRESCALE>(c1, left_border, right_border, use_lh, use_rh, b);
        vstep<4>(st, j, a, b, y0, y1, oh, g, out, colbyte, active);
        c0 = n0; c1 = n1;
    }
}

// ----------------------------------------------------------------------------
// packed 8-bit 4:2:2 input: one warp produces the Y strip (128 columns) and the matching
// U and V strips (64 columns each) from a single read of the packed rows.
struct Raw422Row {
    uint4 v;        // 8 luma + 4 U + 4 V of this lane
    uint2 lh;       // previous 8 bytes (lane 0 only)
    uint2 rh;       // next 8 bytes (lane 31 only)
};

__device__ __forceinline__ void load_422_row(const unsigned char *in, int pitch, int row, int byte0, bool active,
                                             bool use_lh, bool use_rh, Raw422Row &r)
{
    const unsigned char *p = in + (long long)row * pitch + byte0;
    r.v = active ? __ldg(reinterpret_cast<const uint4 *>(p)) : make_uint4(0, 0, 0, 0);
    r.lh = use_lh ? __ldg(reinterpret_cast<const uint2 *>(p - 8)) : make_uint2(0, 0);
    r.rh = use_rh ? __ldg(reinterpret_cast<const uint2 *>(p + 16)) : make_uint2(0, 0);
}

struct Sel422 {     // dp4a coefficient words (already scaled by 1 << shift)
    int ysum, ydif, u, v;
};

// Y: oy[0..3] low, oy[4..7] high.  U/V: o[0..1] low, o[2..3] high.
__device__ __forceinline__ void hfilter_422(const Raw422Row &r, const Sel422 &sel, bool left_border, bool right_border,
                                            bool use_lh, bool use_rh, int *oy, int *ou, int *ov)
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
    int Su[2] = {cu[0] + cu[1], cu[2] + cu[3]}, du[2] = {cu[0] - cu[1], cu[2] - cu[3]};
    int Sv[2] = {cv[0] + cv[1], cv[2] + cv[3]}, dv[2] = {cv[0] - cv[1], cv[2] - cv[3]};
    int Sp = __shfl_up_sync(kFull, S[3], 1), Sn = __shfl_down_sync(kFull, S[0], 1);
    int Sup = __shfl_up_sync(kFull, Su[1], 1), Sun = __shfl_down_sync(kFull, Su[0], 1);
    int Svp = __shfl_up_sync(kFull, Sv[1], 1), Svn = __shfl_down_sync(kFull, Sv[0], 1);
    if (use_lh) {
        Sp = dp4a_us(r.lh.y, sel.ysum, 0);
        Sup = dp4a_us(r.lh.y, sel.u, dp4a_us(r.lh.x, sel.u, 0));
        Svp = dp4a_us(r.lh.y, sel.v, dp4a_us(r.lh.x, sel.v, 0));
    }
    if (use_rh) {
        Sn = dp4a_us(r.rh.x, sel.ysum, 0);
        Sun = dp4a_us(r.rh.y, sel.u, dp4a_us(r.rh.x, sel.u, 0));
        Svn = dp4a_us(r.rh.y, sel.v, dp4a_us(r.rh.x, sel.v, 0));
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
    if (left_border) {
        oy[4] = clamp16((-3 * S[0] + 8 * d[0] + 4 * S[1] - S[2] + 4) >> 3);
        ou[2] = clamp16((-3 * Su[0] + 8 * du[0] + 4 * Su[1] - Sun + 4) >> 3);
        ov[2] = clamp16((-3 * Sv[0] + 8 * dv[0] + 4 * Sv[1] - Svn + 4) >> 3);
    }
    if (right_border) {
        oy[7] = clamp16((3 * S[3] + 8 * d[3] - 4 * S[2] + S[1] + 4) >> 3);
        ou[3] = clamp16((3 * Su[1] + 8 * du[1] - 4 * Su[0] + Sup + 4) >> 3);
        ov[3] = clamp16((3 * Sv[1] + 8 * dv[1] - 4 * Sv[0] + Svp + 4) >> 3);
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
    const int y0 = (blockIdx.y * blockDim.y + threadIdx.y) * p.th;
    if (y0 >= oh) return;
    const int y1 = min(y0 + p.th, oh);

    const int col0 = strip * kStripIn + lane * 8;       // luma sample index
    const bool active = col0 < gy.width;
    const bool left_border = (col0 == 0);
    const bool right_border = (col0 + 8 == gy.width);
    const bool use_lh = (lane == 0) && (strip > 0);
    const bool use_rh = (lane == 31) && (col0 + 8 < gy.width);
    const unsigned char *in = p.in_base[f] + gy.in_off;
    unsigned char *out = p.out_base[f];
    const int colbyte_y = (strip * kStripOut + lane * 4) * 2;
    const int colbyte_c = (strip * (kStripOut / 2) + lane * 2) * 2;
    const int byte0 = col0 * 2;

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

    int jb, je;
    pair_range(y0, y1, oh, jb, je);

    VState<4> sy;
    VState<2> su, sv;
#pragma unroll
    for (int i = 0; i < 8; i++) { sy.llp[i] = sy.llc[i] = sy.dc[i] = sy.d0[i] = 0; }
#pragma unroll
    for (int i = 0; i < 4; i++) { su.llp[i] = su.llc[i] = su.dc[i] = su.d0[i] = 0; sv.llp[i] = sv.llc[i] = sv.dc[i] = sv.d0[i] = 0; }

    Raw422Row c0, c1, n0, n1;
    load_422_row(in, gy.in_pitch, 2 * jb, byte0, active, use_lh, use_rh, c0);
    load_422_row(in, gy.in_pitch, 2 * jb + 1, byte0, active, use_lh, use_rh, c1);
    for (int j = jb; j <= je; j++) {
        if (j < je) {
            load_422_row(in, gy.in_pitch, 2 * j + 2, byte0, active, use_lh, use_rh, n0);
            load_422_row(in, gy.in_pitch, 2 * j + 3, byte0, active, use_lh, use_rh, n1);
        }
        int ay[8], by[8], au[4], bu[4], av[4], bv[4];
        hfilter_422(c0, sel, left_border, right_border, use_lh, use_rh, ay, au, av);
        hfilter_422(c1, sel, left_border, right_border, use_lh, use_rh, by, bu, bv);
        vstep<4>(sy, j, ay, by, y0, y1, oh, gy, out, colbyte_y, active);
        vstep<2>(su, j, au, bu, y0, y1, oh, gu, out, colbyte_c, active);
        vstep<2>(sv, j, av, bv, y0, y1, oh, gv, out, colbyte_c, active);
        c0 = n0; c1 = n1;
    }
}

// ----------------------------------------------------------------------------
// host-side launchers (called from cfb_api.cu)
static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

cudaError_t launch_fwd_plane(const FwdParams &p, int prescale, cudaStream_t stream)
{
    int maxw = 0, maxoh = 0;
    for (int c = 0; c < p.nchan; c++) { maxw = max(maxw, p.ch[c].width); maxoh = max(maxoh, p.ch[c].height / 2); }
    dim3 block(32, 4);
    dim3 grid(ceil_div(maxw, kStripIn), ceil_div(ceil_div(maxoh, p.th), (int)block.y), p.nframes * p.nchan);
    if (prescale) k_fwd_plane<2><<<grid, block, 0, stream>>>(p);
    else k_fwd_plane<0><<<grid, block, 0, stream>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_fwd_422(const FwdParams &p, cudaStream_t stream)
{
    dim3 block(32, 4);
    dim3 grid(ceil_div(p.ch[0].width, kStripIn), ceil_div(ceil_div(p.ch[0].height / 2, p.th), (int)block.y), p.nframes);
    k_fwd_422<<<grid, block, 0, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace cfb
