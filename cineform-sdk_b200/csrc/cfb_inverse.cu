// cfb_inverse.cu -- inverse 2-6 wavelet level with fused dequantisation, sm_100a.
//
// Replaces (reference):
//   Codec/spatial.c:21877 InvertSpatialQuant16s + Codec/InvertHorizontalStrip16s.c:459   -> k_inv_plane<0>
//   Codec/spatial.c:22414 InvertSpatialQuantDescale16s + InvertHorizontalStrip16s.c:1700 -> k_inv_plane<2>
//   Codec/spatial.c:31341/:31511/:31975 InvertSpatial{Top,Middle,Bottom}Row16sToOutput +
//   Codec/InvertHorizontalStrip16s.c:3770/:5025 InvertHorizontalStrip16sToYUYV/ToUYVY   -> k_inv_422
//   Codec/decoder.c:20551 DeQuantFSM (coefficient * quant)                               -> fused into the loads
//
// Same structure as the forward kernels: no shared memory, one warp per strip, registers only.
// A lane owns 4 band columns; per band row it loads 8 bytes from each of the four bands, keeps a
// three-row window of the two vertically-lowpass bands (LL, LH) in registers, produces the even/odd
// intermediate rows, exchanges one value with each neighbour lane by shuffle for the horizontal
// stage and writes 8 output samples per row with one 128-bit store.  Lanes 0 and 31 of a warp are
// halo lanes (their columns belong to the neighbouring strips), so a strip covers 120 band columns.
#include "cfb_common.cuh"

namespace cfb {

constexpr unsigned kFullMask = 0xffffffffu;

template <int NC>
__device__ __forceinline__ void load_cols(const unsigned char *band, int pitch, int row, int colbyte, int dq, bool active, int *v)
{
    if (NC == 4) {
        uint2 w = active ? __ldg(reinterpret_cast<const uint2 *>(band + (long long)row * pitch + colbyte)) : make_uint2(0, 0);
        v[0] = lo16(w.x) * dq; v[1] = hi16(w.x) * dq; v[2] = lo16(w.y) * dq; v[3] = hi16(w.y) * dq;
    } else {
        unsigned w = active ? __ldg(reinterpret_cast<const unsigned *>(band + (long long)row * pitch + colbyte)) : 0u;
        v[0] = lo16(w) * dq; v[1] = hi16(w) * dq;
    }
}

// vertical inverse for NC columns: rows (p, c, n) of the low band and row c of the high band
template <int NC>
__device__ __forceinline__ void vinv_mid(const int *p, const int *c, const int *n, const int *h, int *e, int *o)
{
#pragma unroll
    for (int i = 0; i < NC; i++) {
        e[i] = (((p[i] - n[i] + 4) >> 3) + c[i] + h[i]) >> 1;
        o[i] = (((n[i] - p[i] + 4) >> 3) + c[i] - h[i]) >> 1;
    }
}
// top border: a0,a1,a2 = rows 0,1,2 ; bottom border: call with a0,a1,a2 = rows H-1,H-2,H-3 and swap=true
template <int NC>
__device__ __forceinline__ void vinv_border(const int *a0, const int *a1, const int *a2, const int *h, bool bottom, int *e, int *o)
{
#pragma unroll
    for (int i = 0; i < NC; i++) {
        const int x = (11 * a0[i] - 4 * a1[i] + a2[i] + 4) >> 3;
        const int y = (5 * a0[i] + 4 * a1[i] - a2[i] + 4) >> 3;
        e[i] = ((bottom ? y : x) + h[i]) >> 1;
        o[i] = ((bottom ? x : y) - h[i]) >> 1;
    }
}

// horizontal inverse for NC columns -> 2*NC samples t (BEFORE the final >>1 / <<1):
//   t[2i] = ((l[i-1] - l[i+1] + 4) >> 3) + l[i] + h[i],  t[2i+1] = ((l[i+1] - l[i-1] + 4) >> 3) + l[i] - h[i]
template <int NC>
__device__ __forceinline__ void hinv(const int *l, const int *h, bool left_border, bool right_border, int *t)
{
    const int lp = __shfl_up_sync(kFullMask, l[NC - 1], 1);
    const int ln = __shfl_down_sync(kFullMask, l[0], 1);
#pragma unroll
    for (int i = 0; i < NC; i++) {
        const int a = (i == 0) ? lp : l[i - 1];
        const int b = (i == NC - 1) ? ln : l[i + 1];
        t[2 * i] = ((a - b + 4) >> 3) + l[i] + h[i];
        t[2 * i + 1] = ((b - a + 4) >> 3) + l[i] - h[i];
    }
    if (left_border) {
        const int l2 = (NC > 2) ? l[2] : ln;
        t[0] = ((11 * l[0] - 4 * l[1] + l2 + 4) >> 3) + h[0];
        t[1] = ((5 * l[0] + 4 * l[1] - l2 + 4) >> 3) - h[0];
    }
    if (right_border) {
        const int k = NC - 1;
        const int l2 = (NC > 2) ? l[k - 2] : lp;
        t[2 * k] = ((5 * l[k] + 4 * l[k - 1] - l2 + 4) >> 3) + h[k];
        t[2 * k + 1] = ((11 * l[k] - 4 * l[k - 1] + l2 + 4) >> 3) - h[k];
    }
}

__device__ __forceinline__ unsigned pack_sat16(int lo, int hi) {
    unsigned d;
    asm("cvt.pack.sat.s16.s32 %0, %1, %2;" : "=r"(d) : "r"(hi), "r"(lo));
    return d;
}

// ----------------------------------------------------------------------------
// generic level: 4 bands -> int16 plane (2W x 2H)
template <int DESCALE>
__global__ void __launch_bounds__(128) k_inv_plane(const __grid_constant__ InvParams p)
{
    const int lane = threadIdx.x;
    const int f = blockIdx.z / p.nchan, c = blockIdx.z - f * p.nchan;
    const InvGeom &g = p.ch[c];
    const int strip = blockIdx.x;
    if (strip * kInvStrip >= g.width) return;
    const int H = g.height;
    const int y0 = (blockIdx.y * blockDim.y + threadIdx.y) * p.th;
    if (y0 >= H) return;
    const int y1 = min(y0 + p.th, H);

    const int col0 = strip * kInvStrip - 4 + lane * 4;          // first band column of this lane
    const bool active = (col0 >= 0) && (col0 < g.width);
    const bool writer = active && lane >= 1 && lane <= 30;
    const bool left_border = (col0 == 0);
    const bool right_border = (col0 + 4 == g.width);
    const int colbyte = col0 * 2;
    const unsigned char *in = p.in_base[f];
    const unsigned char *bll = in + g.band_off[0], *blh = in + g.band_off[1];
    const unsigned char *bhl = in + g.band_off[2], *bhh = in + g.band_off[3];
    unsigned char *out = p.out_base[f] + g.out_off + (long long)col0 * 4;

    int lp[4], lc[4], ln[4], hp[4], hc[4], hn[4];      // LL and LH rows r-1, r, r+1
    int r = y0;
    load_cols<4>(bll, g.pitch, max(r - 1, 0), colbyte, g.dq[0], active, lp);
    load_cols<4>(blh, g.pitch, max(r - 1, 0), colbyte, g.dq[1], active, hp);
    load_cols<4>(bll, g.pitch, r, colbyte, g.dq[0], active, lc);
    load_cols<4>(blh, g.pitch, r, colbyte, g.dq[1], active, hc);
    for (; r < y1; r++) {
        int vhl[4], vhh[4];
        const int rn = min(r + 1, H - 1);
        load_cols<4>(bll, g.pitch, rn, colbyte, g.dq[0], active, ln);
        load_cols<4>(blh, g.pitch, rn, colbyte, g.dq[1], active, hn);
        load_cols<4>(bhl, g.pitch, r, colbyte, g.dq[2], active, vhl);
        load_cols<4>(bhh, g.pitch, r, colbyte, g.dq[3], active, vhh);
        int el[4], ol[4], eh[4], oh[4];
        if (r == 0) {
            int l2[4], h2[4];
            load_cols<4>(bll, g.pitch, 2, colbyte, g.dq[0], active, l2);
            load_cols<4>(blh, g.pitch, 2, colbyte, g.dq[1], active, h2);
            vinv_border<4>(lc, ln, l2, vhl, false, el, ol);
            vinv_border<4>(hc, hn, h2, vhh, false, eh, oh);
        } else if (r == H - 1) {
            int l2[4], h2[4];
            load_cols<4>(bll, g.pitch, H - 3, colbyte, g.dq[0], active, l2);
            load_cols<4>(blh, g.pitch, H - 3, colbyte, g.dq[1], active, h2);
            vinv_border<4>(lc, lp, l2, vhl, true, el, ol);
            vinv_border<4>(hc, hp, h2, vhh, true, eh, oh);
        } else {
            vinv_mid<4>(lp, lc, ln, vhl, el, ol);
            vinv_mid<4>(hp, hc, hn, vhh, eh, oh);
        }
        int te[8], to[8];
        hinv<4>(el, eh, left_border, right_border, te);
        hinv<4>(ol, oh, left_border, right_border, to);
        if (writer) {
            uint4 a, b;
            if (DESCALE) {
                a = make_uint4(pack_sat16(te[0] << 1, te[1] << 1), pack_sat16(te[2] << 1, te[3] << 1),
                               pack_sat16(te[4] << 1, te[5] << 1), pack_sat16(te[6] << 1, te[7] << 1));
                b = make_uint4(pack_sat16(to[0] << 1, to[1] << 1), pack_sat16(to[2] << 1, to[3] << 1),
                               pack_sat16(to[4] << 1, to[5] << 1), pack_sat16(to[6] << 1, to[7] << 1));
            } else {
                a = make_uint4(pack_sat16(te[0] >> 1, te[1] >> 1), pack_sat16(te[2] >> 1, te[3] >> 1),
                               pack_sat16(te[4] >> 1, te[5] >> 1), pack_sat16(te[6] >> 1, te[7] >> 1));
                b = make_uint4(pack_sat16(to[0] >> 1, to[1] >> 1), pack_sat16(to[2] >> 1, to[3] >> 1),
                               pack_sat16(to[4] >> 1, to[5] >> 1), pack_sat16(to[6] >> 1, to[7] >> 1));
            }
            *reinterpret_cast<uint4 *>(out + (long long)(2 * r) * g.out_pitch) = a;
            *reinterpret_cast<uint4 *>(out + (long long)(2 * r + 1) * g.out_pitch) = b;
        }
#pragma unroll
        for (int i = 0; i < 4; i++) { lp[i] = lc[i]; lc[i] = ln[i]; hp[i] = hc[i]; hc[i] = hn[i]; }
    }
}

// ----------------------------------------------------------------------------
// final level of a 4:2:2 frame: 12 bands -> packed 8-bit YUYV / UYVY.
// 8-bit reduction: the reference computes v = max(t, 0) >> 1 (10-bit) and out = sat_u8((v + d) >> 2) with
// d = rand() & 1 per position (InvertHorizontalStrip16s.c:3807-3892) - not reproducible.  We use the
// deterministic ordered dither d = (x ^ y) & 1, i.e. out = sat_u8((t + 2d) >> 3), which stays inside the
// reference's envelope {(v) >> 2, (v + 1) >> 2} at every pixel.
struct Col3 { int p[4], c[4], n[4]; };

__device__ __forceinline__ unsigned pack_u8x4(int a, int b, int c, int d) {
    // bytes (LSB first): a, b, c, d, each saturated to [0,255].
    // cvt.pack.sat.u8.s32.b32 r, x, y, z  ->  r = (z << 16) | (sat(x) << 8) | sat(y)
    unsigned t, r;
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(t) : "r"(d), "r"(c), "r"(0));
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(b), "r"(a), "r"(t));
    return r;
}

template <int NC>
__device__ __forceinline__ void inv_rows(const InvGeom &g, const unsigned char *in, int r, int H, int colbyte, bool active,
                                         int *lp, int *lc, int *ln, int *hp, int *hc, int *hn,
                                         bool left_border, bool right_border, int *te, int *to)
{
    const unsigned char *bll = in + g.band_off[0], *blh = in + g.band_off[1];
    const unsigned char *bhl = in + g.band_off[2], *bhh = in + g.band_off[3];
    int vhl[NC], vhh[NC];
    const int rn = min(r + 1, H - 1);
    load_cols<NC>(bll, g.pitch, rn, colbyte, g.dq[0], active, ln);
    load_cols<NC>(blh, g.pitch, rn, colbyte, g.dq[1], active, hn);
    load_cols<NC>(bhl, g.pitch, r, colbyte, g.dq[2], active, vhl);
    load_cols<NC>(bhh, g.pitch, r, colbyte, g.dq[3], active, vhh);
    int el[NC], ol[NC], eh[NC], oh[NC];
    if (r == 0) {
        int l2[NC], h2[NC];
        load_cols<NC>(bll, g.pitch, 2, colbyte, g.dq[0], active, l2);
        load_cols<NC>(blh, g.pitch, 2, colbyte, g.dq[1], active, h2);
        vinv_border<NC>(lc, ln, l2, vhl, false, el, ol);
        vinv_border<NC>(hc, hn, h2, vhh, false, eh, oh);
    } else if (r == H - 1) {
        int l2[NC], h2[NC];
        load_cols<NC>(bll, g.pitch, H - 3, colbyte, g.dq[0], active, l2);
        load_cols<NC>(blh, g.pitch, H - 3, colbyte, g.dq[1], active, h2);
        vinv_border<NC>(lc, lp, l2, vhl, true, el, ol);
        vinv_border<NC>(hc, hp, h2, vhh, true, eh, oh);
    } else {
        vinv_mid<NC>(lp, lc, ln, vhl, el, ol);
        vinv_mid<NC>(hp, hc, hn, vhh, eh, oh);
    }
    hinv<NC>(el, eh, left_border, right_border, te);
    hinv<NC>(ol, oh, left_border, right_border, to);
#pragma unroll
    for (int i = 0; i < NC; i++) { lp[i] = lc[i]; lc[i] = ln[i]; hp[i] = hc[i]; hc[i] = hn[i]; }
}

__global__ void __launch_bounds__(128) k_inv_422(const __grid_constant__ InvParams p)
{
    const int lane = threadIdx.x;
    const int f = blockIdx.z;
    const InvGeom &gy = p.ch[0];
    const InvGeom &gv = p.ch[1];
    const InvGeom &gu = p.ch[2];
    const int strip = blockIdx.x;
    if (strip * kInvStrip >= gy.width) return;
    const int H = gy.height;
    const int y0 = (blockIdx.y * blockDim.y + threadIdx.y) * p.th;
    if (y0 >= H) return;
    const int y1 = min(y0 + p.th, H);

    const int col0 = strip * kInvStrip - 4 + lane * 4;      // luma band column
    const int ccol0 = col0 >> 1;                            // chroma band column (2 per lane)
    const bool active = (col0 >= 0) && (col0 < gy.width);
    const bool writer = active && lane >= 1 && lane <= 30;
    const bool left_border = (col0 == 0);
    const bool right_border = (col0 + 4 == gy.width);
    const unsigned char *in = p.in_base[f];
    unsigned char *out = p.out_base[f] + gy.out_off + (long long)col0 * 4;     // 2 bytes per luma sample, 2 samples per column

    int ylp[4], ylc[4], yln[4], yhp[4], yhc[4], yhn[4];
    int ulp[2], ulc[2], uln[2], uhp[2], uhc[2], uhn[2];
    int vlp[2], vlc[2], vln[2], vhp[2], vhc[2], vhn[2];
    int r = y0;
    {
        const int rp = max(r - 1, 0);
        load_cols<4>(in + gy.band_off[0], gy.pitch, rp, col0 * 2, gy.dq[0], active, ylp);
        load_cols<4>(in + gy.band_off[1], gy.pitch, rp, col0 * 2, gy.dq[1], active, yhp);
        load_cols<4>(in + gy.band_off[0], gy.pitch, r, col0 * 2, gy.dq[0], active, ylc);
        load_cols<4>(in + gy.band_off[1], gy.pitch, r, col0 * 2, gy.dq[1], active, yhc);
        load_cols<2>(in + gu.band_off[0], gu.pitch, rp, ccol0 * 2, gu.dq[0], active, ulp);
        load_cols<2>(in + gu.band_off[1], gu.pitch, rp, ccol0 * 2, gu.dq[1], active, uhp);
        load_cols<2>(in + gu.band_off[0], gu.pitch, r, ccol0 * 2, gu.dq[0], active, ulc);
        load_cols<2>(in + gu.band_off[1], gu.pitch, r, ccol0 * 2, gu.dq[1], active, uhc);
        load_cols<2>(in + gv.band_off[0], gv.pitch, rp, ccol0 * 2, gv.dq[0], active, vlp);
        load_cols<2>(in + gv.band_off[1], gv.pitch, rp, ccol0 * 2, gv.dq[1], active, vhp);
        load_cols<2>(in + gv.band_off[0], gv.pitch, r, ccol0 * 2, gv.dq[0], active, vlc);
        load_cols<2>(in + gv.band_off[1], gv.pitch, r, ccol0 * 2, gv.dq[1], active, vhc);
    }
    const int sh = p.shift + 1;     // final >>1 of the filter merged with the >> (precision-8) reduction
    for (; r < y1; r++) {
        int ye[8], yo[8], ue[4], uo[4], ve[4], vo[4];
        inv_rows<4>(gy, in, r, H, col0 * 2, active, ylp, ylc, yln, yhp, yhc, yhn, left_border, right_border, ye, yo);
        inv_rows<2>(gu, in, r, H, ccol0 * 2, active, ulp, ulc, uln, uhp, uhc, uhn, left_border, right_border, ue, uo);
        inv_rows<2>(gv, in, r, H, ccol0 * 2, active, vlp, vlc, vln, vhp, vhc, vhn, left_border, right_border, ve, vo);
        if (writer) {
#pragma unroll
            for (int rr = 0; rr < 2; rr++) {
                const int *yy = rr ? yo : ye, *uu = rr ? uo : ue, *vv = rr ? vo : ve;
                const int row = 2 * r + rr;
                // ordered dither: d = (x ^ y) & 1 on the sample's own column index, scaled to the merged shift
                const int d0 = ((row & 1) ? 1 : 0) << (sh - 2), d1 = ((row & 1) ? 0 : 1) << (sh - 2);
                unsigned w[4];
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int ya = (yy[2 * k] + d0) >> sh, yb = (yy[2 * k + 1] + d1) >> sh;
                    const int cu = (uu[k] + ((k & 1) ? d1 : d0)) >> sh, cv = (vv[k] + ((k & 1) ? d1 : d0)) >> sh;
                    w[k] = p.uyvy ? pack_u8x4(cu, ya, cv, yb) : pack_u8x4(ya, cu, yb, cv);
                }
                *reinterpret_cast<uint4 *>(out + (long long)row * gy.out_pitch) = make_uint4(w[0], w[1], w[2], w[3]);
            }
        }
    }
}

// ----------------------------------------------------------------------------
static inline int ceil_div_i(int a, int b) { return (a + b - 1) / b; }

cudaError_t launch_inv_plane(const InvParams &p, int descale, cudaStream_t stream)
{
    int maxw = 0, maxh = 0;
    for (int c = 0; c < p.nchan; c++) { maxw = max(maxw, p.ch[c].width); maxh = max(maxh, p.ch[c].height); }
    dim3 block(32, 4);
    dim3 grid(ceil_div_i(maxw, kInvStrip), ceil_div_i(ceil_div_i(maxh, p.th), (int)block.y), p.nframes * p.nchan);
    if (descale) k_inv_plane<2><<<grid, block, 0, stream>>>(p);
    else k_inv_plane<0><<<grid, block, 0, stream>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_inv_422(const InvParams &p, cudaStream_t stream)
{
    dim3 block(32, 4);
    dim3 grid(ceil_div_i(p.ch[0].width, kInvStrip), ceil_div_i(ceil_div_i(p.ch[0].height, p.th), (int)block.y), p.nframes);
    k_inv_422<<<grid, block, 0, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace cfb
