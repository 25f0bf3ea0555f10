// cfb_inverse_tma.inl -- final inverse level of a 4:2:2 frame with the twelve band rows staged in shared memory by TMA
// (included by cfb_inverse.cu, inside namespace cfb).
//
// Same arithmetic as k_inv_422 (Expand / vinv_mid / hinv / emit_422 are shared); what changes is how the coefficients
// reach the registers.  k_inv_422 issues twelve global loads per lane and band row (LL, LH, HL, HH of Y, V, U), each with
// its own 64-bit address, plus an L2 prefetch; here every warp owns a private ring of NS stages in shared memory, one
// stage = R band rows of all twelve bands, and lane 0 keeps the ring full with six cp.async.bulk.tensor copies per stage:
// per channel one 2-D box for LL (it lives in the scratch region) and one 3-D box for LH, HL, HH, which are equally
// spaced in the coded region (cfb_layout_compute), so the band index is the third tensor dimension.  Columns left of
// the image and right of it are zero-filled by the copy engine (the `active` predicate of k_inv_422), the strip halo
// is part of the box, and the lanes read their 8 (luma) / 4 (chroma) bytes per band with conflict-free LDS at
// immediate offsets from one base register.  A box has to START on a 16-byte boundary of global memory (measured:
// tools/probes/tma3d_probe.cu -- a start 8 bytes off raises "illegal instruction"), while a strip starts at byte
// 240 * strip - 8 (luma) / 120 * strip - 4 (chroma): the boxes are 272 / 144 bytes wide, start at the 16-byte boundary
// below, and the lanes add the remainder (8 for luma, 12 or 4 for chroma of an even or odd strip).  Memory-level
// parallelism is (NS - 1) * R band rows per warp, independent of the occupancy the register count allows.
//
// Row schedule of a warp: it owns band rows [y0, y1) (1 <= y0, y1 <= H - 1; rows 0 and H - 1 belong to the border
// warps, which stay on the global-load path).  Rows q = y0 - 1 ... y1 are streamed once each; at row q the vertical
// window (q - 2, q - 1, q) of LL / LH is complete, so output band row q - 1 is produced with the HL / HH row that was
// held back (still packed) from the previous step.
struct alignas(64) InvTmaMaps {
    CUtensorMap m[kMaxBatch][6];        // [frame][2 * channel + (0: LL, 1: LH,HL,HH)]
};

template <int NC> struct InvRot {
    int lp[NC], lc[NC], hp[NC], hc[NC];     // LL / LH rows q - 2, q - 1 (expanded)
    RawCols<NC> phl, phh;                   // HL / HH row q - 1 (packed)
};

constexpr int kInvBoxY = 272, kInvBoxC = 144;      // bytes per band row in a stage (luma / chroma)

template <int R>
struct InvStage {
    static constexpr int a128(int x) { return (x + 127) & ~127; }       // every box lands on a 128-byte boundary
    static constexpr int kYLL = 0, kYHP = a128(R * kInvBoxY), kYEnd = kYHP + a128(3 * R * kInvBoxY);
    static constexpr int kCLL = 0, kCHP = a128(R * kInvBoxC), kCEnd = kCHP + a128(3 * R * kInvBoxC);
    static constexpr int kBytes = kYEnd + 2 * kCEnd;
    static constexpr int kTx = 4 * R * kInvBoxY + 8 * R * kInvBoxC;     // bytes the six copies of a stage deliver
    // byte offset of (channel, band, row i) inside a stage; band 0 = LL, 1..3 = LH, HL, HH
    static __device__ __forceinline__ unsigned y(int band, int i) { return band == 0 ? kYLL + i * kInvBoxY : kYHP + ((band - 1) * R + i) * kInvBoxY; }
    static __device__ __forceinline__ unsigned c(int chan, int band, int i) {       // chan 1 = V, 2 = U
        const unsigned base = kYEnd + (chan - 1) * kCEnd;
        return base + (band == 0 ? kCLL + i * kInvBoxC : kCHP + ((band - 1) * R + i) * kInvBoxC);
    }
};

template <int NC, bool SMALLDQ>
__device__ __forceinline__ void inv_rot_fill(InvRot<NC> &s, const InvGeom &g, const RawCols<NC> &ll, const RawCols<NC> &lh,
                                             const RawCols<NC> &hl, const RawCols<NC> &hh)
{
#pragma unroll
    for (int i = 0; i < NC; i++) { s.lp[i] = s.lc[i]; s.hp[i] = s.hc[i]; }
    Expand<SMALLDQ, NC>::ll(ll, s.lc);
    Expand<SMALLDQ, NC>::hp(lh, g.dq[1], s.hc);
    s.phl = hl; s.phh = hh;
}

template <int NC, bool SMALLDQ>
__device__ __forceinline__ void inv_rot_step(InvRot<NC> &s, const InvGeom &g, const RawCols<NC> &ll, const RawCols<NC> &lh,
                                             const RawCols<NC> &hl, const RawCols<NC> &hh,
                                             bool has_border, bool left_border, bool right_border, int *te, int *to)
{
    int ln[NC], hn[NC], vhl[NC], vhh[NC];
    Expand<SMALLDQ, NC>::ll(ll, ln);
    Expand<SMALLDQ, NC>::hp(lh, g.dq[1], hn);
    Expand<SMALLDQ, NC>::hp(s.phl, g.dq[2], vhl);
    Expand<SMALLDQ, NC>::hp(s.phh, g.dq[3], vhh);
    int el[NC], ol[NC], eh[NC], oh[NC];
    vinv_mid<NC>(s.lp, s.lc, ln, vhl, el, ol);
    vinv_mid<NC>(s.hp, s.hc, hn, vhh, eh, oh);
    hinv<NC>(el, eh, has_border, left_border, right_border, te);
    hinv<NC>(ol, oh, has_border, left_border, right_border, to);
#pragma unroll
    for (int i = 0; i < NC; i++) { s.lp[i] = s.lc[i]; s.lc[i] = ln[i]; s.hp[i] = s.hc[i]; s.hc[i] = hn[i]; }
    s.phl = hl; s.phh = hh;
}

template <bool SMALLDQ, bool OUT16, int R, int NS, int MINB>
__global__ void __launch_bounds__(128, MINB) k_inv_422_tma(const __grid_constant__ InvParams p, const __grid_constant__ InvTmaMaps tm)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    typedef InvStage<R> ST;
    const int lane = threadIdx.x, warp = threadIdx.y;
    const int f = blockIdx.z;
    const InvGeom &gy = p.ch[0];
    const InvGeom &gv = p.ch[1];
    const InvGeom &gu = p.ch[2];
    const int strip = blockIdx.x;
    if (strip * kInvStrip >= gy.width) return;
    const int H = gy.height;

    const int col0 = strip * kInvStrip - 4 + lane * 4;      // luma band column
    const bool active = (col0 >= 0) && (col0 < gy.width);
    const bool writer = active && lane >= 1 && lane <= 30;
    const bool left_border = (col0 == 0);
    const bool right_border = (col0 + 4 == gy.width);
    const bool has_border = (strip == 0) || ((strip + 1) * kInvStrip + 4 >= gy.width);
    unsigned char *out = p.out_base[f] + gy.out_off + (long long)col0 * (OUT16 ? 8 : 4);

    if (blockIdx.y == gridDim.y - 1) {          // border warps: band rows 0 and H-1, straight from global memory
        if (warp > 1) return;
        const unsigned ycol = (unsigned)(col0 * 2), ccol = (unsigned)col0;
        const unsigned char *in = p.in_base[f];
        const bool bottom = (warp == 1);
        int ye[8], yo[8], ue[4], uo[4], ve[4], vo[4];
        inv_border_row<4>(gy, in, bottom, H, ycol, active, has_border, left_border, right_border, ye, yo);
        inv_border_row<2>(gu, in, bottom, H, ccol, active, has_border, left_border, right_border, ue, uo);
        inv_border_row<2>(gv, in, bottom, H, ccol, active, has_border, left_border, right_border, ve, vo);
        if (writer) emit_422<OUT16>(p, out, col0, bottom ? H - 1 : 0, ye, yo, ue, uo, ve, vo);
        return;
    }
    const int y0 = max((int)(blockIdx.y * blockDim.y + warp) * p.th, 1);
    const int y1 = min((int)(blockIdx.y * blockDim.y + warp + 1) * p.th, H - 1);
    if (y0 >= y1) return;
    const int qfirst = y0 - 1;
    const int nrows = y1 - qfirst + 1;                      // rows y0 - 1 ... y1

    // ---- this warp's ring ----
    const unsigned ring = smem_u32(smem_raw) + warp * (NS * ST::kBytes);
    const unsigned bars = smem_u32(smem_raw) + 4 * (NS * ST::kBytes) + warp * (NS * 8);
    const CUtensorMap *maps = tm.m[f];
    // 32-bit element coordinates of the boxes: the 16-byte boundary below luma column strip * 120 - 4 (byte 240 * strip - 8)
    // and below chroma column strip * 60 - 2 (byte 120 * strip - 4)
    const int xy = strip * (kInvStrip / 2) - 4;
    const int cbyte = strip * kInvStrip - 4;
    const int xc = (cbyte & ~15) / 4;
    auto issue = [&](int stage, int q0) {
        const unsigned bar = bars + 8 * stage, dst = ring + stage * ST::kBytes;
        mbar_expect_tx(bar, ST::kTx);
        tma_load_2d(dst + ST::y(0, 0), &maps[0], xy, q0, bar);
        tma_load_3d(dst + ST::y(1, 0), &maps[1], xy, q0, 0, bar);
        tma_load_2d(dst + ST::c(1, 0, 0), &maps[2], xc, q0, bar);
        tma_load_3d(dst + ST::c(1, 1, 0), &maps[3], xc, q0, 0, bar);
        tma_load_2d(dst + ST::c(2, 0, 0), &maps[4], xc, q0, bar);
        tma_load_3d(dst + ST::c(2, 1, 0), &maps[5], xc, q0, 0, bar);
    };
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < NS; s++) mbar_init(bars + 8 * s, 1);
        mbar_fence_init();
#pragma unroll
        for (int s = 0; s < NS; s++)
            if (s * R < nrows) issue(s, qfirst + s * R);
    }
    __syncwarp();

    InvRot<4> sy;
    InvRot<2> su, sv;
#pragma unroll
    for (int i = 0; i < 4; i++) { sy.lp[i] = sy.lc[i] = sy.hp[i] = sy.hc[i] = 0; }
#pragma unroll
    for (int i = 0; i < 2; i++) { su.lp[i] = su.lc[i] = su.hp[i] = su.hc[i] = 0; sv.lp[i] = sv.lc[i] = sv.hp[i] = sv.hc[i] = 0; }
    sy.phl.w = make_uint2(0u, 0u); sy.phh.w = make_uint2(0u, 0u);
    su.phl.w = su.phh.w = sv.phl.w = sv.phh.w = 0u;

    const unsigned ly = 8u + (unsigned)lane * 8u, lc = (unsigned)(cbyte & 15) + (unsigned)lane * 4u;
    auto run = [&](auto border_tag) {
        constexpr bool HB = decltype(border_tag)::value;
        int stage = 0;
        unsigned parity = 0;
        int q = qfirst;
#pragma unroll 1
        for (int k = 0; k * R < nrows; k++) {
            const unsigned sb = ring + stage * ST::kBytes;
            mbar_wait(bars + 8 * stage, parity);
#pragma unroll
            for (int i = 0; i < R; i++, q++) {
                if (i > 0 && q > y1) break;
                RawCols<4> yll, ylh, yhl, yhh;
                RawCols<2> vll, vlh, vhl, vhh, ull, ulh, uhl, uhh;
                yll.w = lds64(sb + ST::y(0, i) + ly); ylh.w = lds64(sb + ST::y(1, i) + ly);
                yhl.w = lds64(sb + ST::y(2, i) + ly); yhh.w = lds64(sb + ST::y(3, i) + ly);
                vll.w = lds32(sb + ST::c(1, 0, i) + lc); vlh.w = lds32(sb + ST::c(1, 1, i) + lc);
                vhl.w = lds32(sb + ST::c(1, 2, i) + lc); vhh.w = lds32(sb + ST::c(1, 3, i) + lc);
                ull.w = lds32(sb + ST::c(2, 0, i) + lc); ulh.w = lds32(sb + ST::c(2, 1, i) + lc);
                uhl.w = lds32(sb + ST::c(2, 2, i) + lc); uhh.w = lds32(sb + ST::c(2, 3, i) + lc);
                if (q <= y0) {
                    inv_rot_fill<4, SMALLDQ>(sy, gy, yll, ylh, yhl, yhh);
                    inv_rot_fill<2, SMALLDQ>(su, gu, ull, ulh, uhl, uhh);
                    inv_rot_fill<2, SMALLDQ>(sv, gv, vll, vlh, vhl, vhh);
                    __syncwarp();
                } else {
                    int ye[8], yo[8], ue[4], uo[4], ve[4], vo[4];
                    inv_rot_step<4, SMALLDQ>(sy, gy, yll, ylh, yhl, yhh, HB && has_border, HB && left_border, HB && right_border, ye, yo);
                    inv_rot_step<2, SMALLDQ>(su, gu, ull, ulh, uhl, uhh, HB && has_border, HB && left_border, HB && right_border, ue, uo);
                    inv_rot_step<2, SMALLDQ>(sv, gv, vll, vlh, vhl, vhh, HB && has_border, HB && left_border, HB && right_border, ve, vo);
                    if (writer) emit_422<OUT16>(p, out, col0, q - 1, ye, yo, ue, uo, ve, vo);
                }
                // The PREVIOUS stage has been consumed by every lane: its last HL / HH row was expanded for the output row
                // above, and the warp-wide shuffles of the horizontal stage (or the __syncwarp of the fill path) order
                // every lane's shared-memory reads of that stage before this point.
                if (i == 0 && lane == 0 && k >= 1 && (k - 1 + NS) * R < nrows)
                    issue(stage == 0 ? NS - 1 : stage - 1, qfirst + (k - 1 + NS) * R);
            }
            if (++stage == NS) { stage = 0; parity ^= 1; }
        }
    };
    if (has_border) run(std::true_type{}); else run(std::false_type{});
}
