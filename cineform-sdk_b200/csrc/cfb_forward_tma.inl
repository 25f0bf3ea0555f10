// cfb_forward_tma.inl -- level-1 forward of packed 8-bit 4:2:2 with the input rows staged in shared memory by TMA
// (included by cfb_forward.cu, inside namespace cfb).
//
// Same arithmetic and the same register-resident vertical state as k_fwd_422_rot; what changes is how the packed rows
// reach the registers.  Every warp owns a private ring of kTmaStages stages in shared memory; one stage = the two rows
// of a row pair, 16 B left halo + 512 B strip + 16 B right halo each.  Lane 0 keeps the ring full with
// cp.async.bulk.tensor.2d (UTMALDG) against a per-frame tensor map -- out-of-image halo columns are zero-filled by
// the hardware, and no lane computes a global address, issues a halo LDG or an L2 prefetch any more; all lanes wait on
// the stage's mbarrier and read their 16 bytes (+ 8 halo bytes on the two edge lanes) with conflict-free LDS.128.
// Memory-level parallelism (kTmaStages - 1 row pairs in flight per warp) no longer costs registers, so it does not
// depend on occupancy.
constexpr int kTmaStages = 4;
constexpr int kTmaRowBytes = 544;           // 16 + 512 + 16
constexpr int kTmaStageBytes = 1152;        // 2 rows = 1088 B, rounded up to a multiple of 128 B (TMA destination alignment)
constexpr int kTmaWarpBytes = kTmaStages * kTmaStageBytes;

struct alignas(64) FwdTmaMaps {
    CUtensorMap in_map[kMaxBatch];
};
struct alignas(64) FwdTmaPlaneMaps {            // planar sources: one map per (frame, channel)
    CUtensorMap in_map[kMaxBatch * kMaxChannels];
};

template <int MINB>
__global__ void __launch_bounds__(128, MINB) k_fwd_422_tma(const __grid_constant__ FwdParams p, const __grid_constant__ FwdTmaMaps tm)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x, warp = threadIdx.y;
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
    unsigned char *out = p.out_base[f];

    Sel422 sel;
    {
        const int m = 1 << p.shift;
        const int neg = (-m) & 0xff;
        if (!p.uyvy) { sel.ysum = m | (m << 16); sel.ydif = m | (neg << 16); sel.u = m << 8; sel.v = m << 24; }
        else { sel.ysum = (m << 8) | (m << 24); sel.ydif = (m << 8) | (neg << 24); sel.u = m; sel.v = m << 16; }
    }

    if (blockIdx.y == gridDim.y - 1) {
        // ---- border warps (first / last HL,HH row): six rows straight from global memory, as in k_fwd_422_rot ----
        if (warp > 1) return;
        const unsigned char *in = p.in_base[f] + gy.in_off + (strip * kStripIn + lane * 8) * 2;
        const bool bottom = (warp == 1);
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

    const int y0 = (blockIdx.y * blockDim.y + warp) * p.th;
    if (y0 >= oh) return;
    const int y1 = min(y0 + p.th, oh);
    const int jfirst = max(y0 - 1, 0), jlast = min(y1, oh - 1);
    const int hlo = max(y0, 1);

    // ---- this warp's ring ----
    const unsigned ring = smem_u32(smem_raw) + warp * kTmaWarpBytes;
    const unsigned bars = smem_u32(smem_raw) + 4 * kTmaWarpBytes + warp * (kTmaStages * 8);
    const void *map = &tm.in_map[f];
    const int x0 = strip * (kStripIn * 2 / 4) - 4;          // element (32-bit) coordinate of the strip's first byte, minus the halo
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < kTmaStages; s++) mbar_init(bars + 8 * s, 1);
        mbar_fence_init();
#pragma unroll
        for (int s = 0; s < kTmaStages; s++)
            if (jfirst + s <= jlast) {
                mbar_expect_tx(bars + 8 * s, 2 * kTmaRowBytes);
                tma_load_2d(ring + s * kTmaStageBytes, map, x0, 2 * (jfirst + s), bars + 8 * s);
            }
    }
    __syncwarp();

    RotState<4> sy;
    RotState<2> su, sv;
#pragma unroll
    for (int i = 0; i < 8; i++) { sy.t[i] = sy.s[i] = 0; }
#pragma unroll
    for (int i = 0; i < 4; i++) { su.t[i] = su.s[i] = 0; sv.t[i] = sv.s[i] = 0; }
    unsigned offy = (unsigned)(jfirst * gy.out_pitch) + colbyte_y;
    unsigned offc = (unsigned)(jfirst * gu.out_pitch) + colbyte_c;
    const unsigned lane_off = 16u + (unsigned)lane * 16u;
    const unsigned halo_off = lane_off + (L.use_lh ? -8 : 16);
    const bool has_halo = L.use_lh | L.use_rh;

    auto step = [&](int jj, int stage, unsigned parity, auto border_tag) {
        constexpr bool BORDER = decltype(border_tag)::value;
        const unsigned sb = ring + stage * kTmaStageBytes;
        mbar_wait(bars + 8 * stage, parity);
        Raw422Row r0, r1;
        r0.v = lds128(sb + lane_off);
        r1.v = lds128(sb + kTmaRowBytes + lane_off);
        r0.halo = make_uint2(0u, 0u); r1.halo = make_uint2(0u, 0u);
        if (has_halo) { r0.halo = lds64(sb + halo_off); r1.halo = lds64(sb + kTmaRowBytes + halo_off); }
        int ay[8], by[8], au[4], bu[4], av[4], bv[4];
        hfilter_422_t<BORDER>(r0, sel, L, ay, au, av);
        hfilter_422_t<BORDER>(r1, sel, L, by, bu, bv);
        // the stage has been consumed into registers by every lane (the shuffles above are warp-wide): refill it
        if (lane == 0 && jj + kTmaStages <= jlast) {
            mbar_expect_tx(bars + 8 * stage, 2 * kTmaRowBytes);
            tma_load_2d(sb, map, x0, 2 * (jj + kTmaStages), bars + 8 * stage);
        }
        const bool emit_low = (jj >= y0) && (jj < y1), emit_high = (jj - 1 >= hlo);
        vstep_rot<4, 0>(sy, ay, by, gy, out, offy, emit_low, emit_high);
        vstep_rot<2, 0>(su, au, bu, gu, out, offc, emit_low, emit_high);
        vstep_rot<2, 0>(sv, av, bv, gv, out, offc, emit_low, emit_high);
        offy += (unsigned)gy.out_pitch;
        offc += (unsigned)gu.out_pitch;
    };
    auto run = [&](auto border_tag) {
        int stage = 0;
        unsigned parity = 0;
#pragma unroll 1
        for (int j = jfirst; j <= jlast; j++) {
            step(j, stage, parity, border_tag);
            if (++stage == kTmaStages) { stage = 0; parity ^= 1; }
        }
    };
    if (L.has_border) run(std::true_type{}); else run(std::false_type{});
}


// ------------------------------------------------------------------------------------------------------------------
// Generic TMA-fed forward level for sources whose channels are separated while loading:
//   SrcPlane16<PRESCALE>  one int16 plane per channel (levels 2 and 3 of every format, PLANAR16 level 1, cfb_level_*)
//   SrcRG48               packed 16-bit R,G,B -> planes G, R, B: ONE read of the 48-byte pixel groups feeds all three
//                         channel warps of the CTA (round 1 launched one kernel per channel: 3 x the input traffic)
//   SrcBYR4<LUT>          16-bit Bayer quads -> planes G, R-G, B-G, dG: the two Bayer lines of a plane row are loaded
//                         once for the four channel warps
// A CTA = SRC::kWarps warps, all on the same strip and the same row block, one channel each.  Warp 0 / lane 0 keeps a
// CTA-wide ring of kTmaStages stages full (cp.async.bulk.tensor.2d, one box of SRC::kBoxRows rows per row pair); a stage
// is handed back through an "empty" mbarrier on which every channel warp arrives once it has pulled its samples into
// registers.  The first / last HL,HH row (6-tap border filters) is left to the round-1 kernels launched on their border
// CTA row alone (gridDim.y = 1), so this kernel has no border-row code.
template <int PRESCALE>
struct SrcPlane16 {
    static constexpr int kWarps = 1, kBoxRows = 2, kRowBytes = 544, kElem = 4, kPrescale = PRESCALE, kQll = 1;
    static constexpr int kStageBytes = 1152;
    static __device__ __forceinline__ int x0(int strip) { return strip * (kStripIn * 2 / 4) - 4; }
    static __device__ __forceinline__ int map_index(int f, int c, int nchan) { return f * nchan + c; }
    static __device__ __forceinline__ void extract(unsigned sb, int row, int lane, int, const LaneInfo &L, const FwdParams &, RawPlaneRow &r) {
        const unsigned a = sb + row * kRowBytes + 16u + (unsigned)lane * 16u;
        r.v = lds128(a);
        r.halo = 0u;
        if (L.use_lh | L.use_rh) r.halo = lds32(a + (L.use_lh ? -4 : 16));
    }
};

struct SrcRG48 {
    static constexpr int kWarps = 3, kBoxRows = 2, kRowBytes = 1568, kElem = 8, kPrescale = 0, kQll = 1;
    static constexpr int kStageBytes = 3200;        // 2 x 1568 rounded up to a multiple of 128
    static __device__ __forceinline__ int x0(int strip) { return strip * (kStripIn * 6 / 8) - 2; }
    static __device__ __forceinline__ int map_index(int f, int, int) { return f; }
    // channel 0 = G (word 1 of a pixel), 1 = R (word 0), 2 = B (word 2): Codec/frame.c:6155-6157
    static __device__ __forceinline__ void extract(unsigned sb, int row, int lane, int chan, const LaneInfo &L, const FwdParams &p, RawPlaneRow &r) {
        const unsigned a = sb + row * kRowBytes + 16u + (unsigned)lane * 48u;
        RawRG48Row q;
        q.a = lds128(a); q.b = lds128(a + 16); q.c = lds128(a + 32);
        q.halo = 0u;
        const int sel = (chan == 0) ? 1 : (chan == 1 ? 0 : 2);
        if (L.use_lh | L.use_rh) {
            const unsigned h = a + (L.use_lh ? -12 : 48) + 2 * sel;
            q.halo = lds_u16(h) | (lds_u16(h + 6) << 16);
        }
        if (sel == 0) rg48_extract<0>(q, p.shift, r); else if (sel == 1) rg48_extract<1>(q, p.shift, r); else rg48_extract<2>(q, p.shift, r);
    }
};

template <bool LUT>
struct SrcBYR4 {
    static constexpr int kWarps = 4, kBoxRows = 4, kRowBytes = 1056, kElem = 8, kPrescale = 0, kQll = 1;
    static constexpr int kStageBytes = 4224;        // 4 Bayer lines x 1056
    static __device__ __forceinline__ int x0(int strip) { return strip * (kStripIn * 4 / 8) - 2; }
    static __device__ __forceinline__ int map_index(int f, int, int) { return f; }
    static __device__ __forceinline__ void extract(unsigned sb, int row, int lane, int chan, const LaneInfo &L, const FwdParams &p, RawPlaneRow &r) {
        const unsigned a = sb + (2 * row) * kRowBytes + 16u + (unsigned)lane * 32u;     // first Bayer line of the plane row
        RawBYR4Row q;
        q.a0 = lds128(a); q.a1 = lds128(a + 16);
        q.b0 = lds128(a + kRowBytes); q.b1 = lds128(a + kRowBytes + 16);
        q.ha = make_uint2(0u, 0u); q.hb = make_uint2(0u, 0u);
        if (L.use_lh | L.use_rh) {
            const unsigned h = a + (L.use_lh ? -8 : 32);
            q.ha = lds64(h); q.hb = lds64(h + kRowBytes);
        }
        const BayerSel s = bayer_sel(p.uyvy, chan);         // warp-uniform (hoisted out of the row loop by the compiler)
        if (chan == 0) byr4_extract_c<LUT, 0>(q, p.shift, s, p.lut, r);
        else if (chan == 1) byr4_extract_c<LUT, 1>(q, p.shift, s, p.lut, r);
        else if (chan == 2) byr4_extract_c<LUT, 2>(q, p.shift, s, p.lut, r);
        else byr4_extract_c<LUT, 3>(q, p.shift, s, p.lut, r);
    }
};

template <class SRC, int MINB>
__global__ void __launch_bounds__(32 * SRC::kWarps, MINB) k_fwd_tma(const __grid_constant__ FwdParams p, const __grid_constant__ FwdTmaPlaneMaps tm)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int NW = SRC::kWarps;
    const int lane = threadIdx.x, warp = threadIdx.y;
    // planar sources: blockIdx.z = frame * nchan + channel; interleaved sources: blockIdx.z = frame, channel = warp
    const int f = (NW == 1) ? blockIdx.z / p.nchan : blockIdx.z;
    const int c = (NW == 1) ? blockIdx.z - f * p.nchan : warp;
    const PlaneGeom &g = p.ch[c];
    const int strip = blockIdx.x;
    if (strip * kStripIn >= g.width) return;
    const int oh = g.height >> 1;
    LaneInfo L;
    const bool lane_on = lane_setup(strip, g.width, lane, L);
    const int y0 = blockIdx.y * p.th;
    if (y0 >= oh) return;
    const int y1 = min(y0 + p.th, oh);
    const int jfirst = max(y0 - 1, 0), jlast = min(y1, oh - 1);
    const int hlo = max(y0, 1);

    const unsigned ring = smem_u32(smem_raw);
    const unsigned full = ring + kTmaStages * SRC::kStageBytes, empty = full + kTmaStages * 8;
    const void *map = &tm.in_map[SRC::map_index(f, c, p.nchan)];
    const int x0 = SRC::x0(strip);
    constexpr unsigned kTx = SRC::kBoxRows * SRC::kRowBytes;
    const bool producer = (warp == 0 && lane == 0);
    if (producer) {
#pragma unroll
        for (int s = 0; s < kTmaStages; s++) { mbar_init(full + 8 * s, 1); mbar_init(empty + 8 * s, NW); }
        mbar_fence_init();
    }
    if (NW > 1) __syncthreads(); else __syncwarp();
    if (producer) {
#pragma unroll
        for (int s = 0; s < kTmaStages; s++)
            if (jfirst + s <= jlast) {
                mbar_expect_tx(full + 8 * s, kTx);
                tma_load_2d(ring + s * SRC::kStageBytes, map, x0, SRC::kBoxRows * (jfirst + s), full + 8 * s);
            }
    }

    unsigned char *out = p.out_base[f];
    const unsigned colbyte = (unsigned)((strip * kStripOut + lane * 4) * 2);
    RotState<4> st;
#pragma unroll
    for (int i = 0; i < 8; i++) { st.t[i] = st.s[i] = 0; }
    unsigned off = (unsigned)(jfirst * g.out_pitch) + colbyte;
    int stage = 0;
    unsigned parity = 0;
#pragma unroll 1
    for (int j = jfirst; j <= jlast; j++) {
        const unsigned sb = ring + stage * SRC::kStageBytes;
        mbar_wait(full + 8 * stage, parity);
        RawPlaneRow r0, r1;
        r0.v = make_uint4(0u, 0u, 0u, 0u); r0.halo = 0u; r1 = r0;
        if (lane_on) {
            SRC::extract(sb, 0, lane, c, L, p, r0);
            SRC::extract(sb, 1, lane, c, L, p, r1);
        }
        __syncwarp();
        if (NW > 1) {
            // hand the stage back: every channel warp arrives once; the producer refills it when all have
            if (lane == 0) mbar_arrive(empty + 8 * stage);
            if (producer && j + kTmaStages <= jlast) {
                mbar_wait(empty + 8 * stage, parity);
                mbar_expect_tx(full + 8 * stage, kTx);
                tma_load_2d(sb, map, x0, SRC::kBoxRows * (j + kTmaStages), full + 8 * stage);
            }
        } else if (producer && j + kTmaStages <= jlast) {
            mbar_expect_tx(full + 8 * stage, kTx);
            tma_load_2d(sb, map, x0, SRC::kBoxRows * (j + kTmaStages), full + 8 * stage);
        }
        if (lane_on) {
            int a[8], b[8];
            hfilter_plane<SRC::kPrescale>(r0, L, a);
            hfilter_plane<SRC::kPrescale>(r1, L, b);
            vstep_rot<4, SRC::kQll>(st, a, b, g, out, off, j >= y0 && j < y1, j - 1 >= hlo);
        }
        off += (unsigned)g.out_pitch;
        if (++stage == kTmaStages) { stage = 0; parity ^= 1; }
    }
}
