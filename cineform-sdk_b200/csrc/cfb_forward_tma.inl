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
