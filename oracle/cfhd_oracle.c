/* cfhd_oracle.c -- TEST INFRASTRUCTURE ONLY (see cfhd_oracle.h).
 *
 * Scalar C restatement of the reference's 2-6 wavelet + quantiser, including the
 * places where the reference's SSE2 loops (saturating 16-bit chains) and its
 * scalar tails/borders (int32 + clamp, or wrapping 16-bit) give different
 * results once an intermediate leaves int16.  Written from the behaviour of the
 * cited functions; no reference source is included or linked here.
 */
#include "cfhd_oracle.h"
#include <stdlib.h>
#include <string.h>

/* ---- 16-bit primitives (semantics of SSE2 adds/subs_epi16, srai_epi16) ---- */
static inline int16_t sat16(int32_t v) { return (int16_t)(v < -32768 ? -32768 : (v > 32767 ? 32767 : v)); }
static inline int16_t adds(int16_t a, int16_t b) { return sat16((int32_t)a + b); }
static inline int16_t subs(int16_t a, int16_t b) { return sat16((int32_t)a - b); }
static inline int16_t wrap16(int32_t v) { return (int16_t)(uint16_t)(uint32_t)v; }
static inline int16_t sra16(int16_t a, int s) { return (int16_t)(a >> s); }

static uint8_t lowpass_px(int16_t c, int shift, int unsigned_shift)
{
    int v = unsigned_shift ? ((int)(uint16_t)c >> shift) : ((int)c >> shift);
    return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

void orc_lowpass_to_422(const int16_t *y, int y_pitch, const int16_t *v, int v_pitch, const int16_t *u, int u_pitch,
                        int width, int height, int shift, int unsigned_shift, int uyvy,
                        uint8_t *out, int out_pitch)
{
    for (int r = 0; r < height; r++) {
        const int16_t *yr = (const int16_t *)((const uint8_t *)y + (size_t)r * y_pitch);
        const int16_t *vr = (const int16_t *)((const uint8_t *)v + (size_t)r * v_pitch);
        const int16_t *ur = (const int16_t *)((const uint8_t *)u + (size_t)r * u_pitch);
        uint8_t *o = out + (size_t)r * out_pitch;
        for (int x = 0; x < width; x += 2) {
            const uint8_t y0 = lowpass_px(yr[x], shift, unsigned_shift), y1 = lowpass_px(yr[x + 1], shift, unsigned_shift);
            const uint8_t cu = lowpass_px(ur[x / 2], shift, unsigned_shift), cv = lowpass_px(vr[x / 2], shift, unsigned_shift);
            if (uyvy) { o[2 * x] = cu; o[2 * x + 1] = y0; o[2 * x + 2] = cv; o[2 * x + 3] = y1; }
            else      { o[2 * x] = y0; o[2 * x + 1] = cu; o[2 * x + 2] = y1; o[2 * x + 3] = cv; }
        }
    }
}

int orc_version(void) { return 1; }

/* ------------------------------------------------------------------------- */
/* Forward horizontal 2-6 on one row.
 * Codec/spatial.c:253-570 (prescale 0) and :3669-4000 (prescale 2).
 * Column split: outputs 1 .. post/2-1 come from the 16-pixel SSE2 loop
 * (saturating chain in the order written there); post/2 .. m-2 from the scalar
 * tail (int32, one clamp); 0 and m-1 from the 6-tap border filters.          */
void orc_fwd_row(const int16_t *x, int16_t *low, int16_t *high, int width, int prescale)
{
    const int m = width / 2;
    const int last_column = width - 2;
    const int post = last_column - (last_column % 16);
    int i;
#define P(v) (prescale ? (((int32_t)(v) + 3) >> 2) : (int32_t)(v))
    /* left border, spatial.c:277-286 / :3697-3706 */
    {
        int32_t s = 5 * P(x[0]) - 11 * P(x[1]) + 4 * P(x[2]) + 4 * P(x[3]) - P(x[4]) - P(x[5]) + 4;
        high[0] = sat16(s >> 3);
    }
    for (i = 0; i < m; i++) {
        const int c = 2 * i;
        const int simd = (c < post);
        /* lowpass */
        if (!prescale) {
            low[i] = sat16((int32_t)x[c] + x[c + 1]);               /* adds == clamp of exact sum */
        } else if (simd) {
            int16_t a = adds(x[c], 3), b = adds(x[c + 1], 3);           /* spatial.c:3712,:3763 */
            low[i] = sra16(subs(adds(a, b), 3), 2);                    /* :3776-3778 */
        } else {
            low[i] = sat16(((int32_t)x[c] + x[c + 1] + 3) >> 2);       /* :3962 */
        }
        /* highpass */
        if (i == 0 || i == m - 1) continue;
        if (c < post) {
            /* output i is produced by the SIMD iteration whose window starts at x[c-2] */
            int16_t t[6];
            int k;
            for (k = 0; k < 6; k++) {
                int16_t v = x[c - 2 + k];
                t[k] = prescale ? sra16(adds(v, 3), 2) : v;
            }
            int16_t s = subs(0, t[0]);
            s = subs(s, t[1]);
            int16_t sb = subs(t[2], t[3]);
            s = adds(s, t[4]);
            s = adds(s, t[5]);
            s = adds(s, 4);
            s = sra16(s, 3);
            high[i] = adds(s, sb);
        } else {
            int32_t s = -P(x[c - 2]) - P(x[c - 1]) + P(x[c + 2]) + P(x[c + 3]) + 4;
            s >>= 3;
            s += P(x[c]) - P(x[c + 1]);
            high[i] = sat16(s);
        }
    }
    /* right border, spatial.c:558-569 / :3984-3996 */
    {
        const int c = last_column;
        int32_t s = 11 * P(x[c]) - 5 * P(x[c + 1]) - 4 * P(x[c - 1]) - 4 * P(x[c - 2]) + P(x[c - 3]) + P(x[c - 4]) + 4;
        high[m - 1] = sat16(s >> 3);
    }
#undef P
}

/* ------------------------------------------------------------------------- */
/* Codec/quantize.c:1395-1516 */
void orc_quantize_row(const int16_t *in, int16_t *out, int length, int divisor, int midpoint_prequant)
{
    int mid = 0, c;
    if (midpoint_prequant >= 2 && midpoint_prequant < 9) {
        mid = divisor / midpoint_prequant;
        if (midpoint_prequant == 2 && mid) mid--;
    }
    if (divisor <= 1) { memmove(out, in, (size_t)length * sizeof(int16_t)); return; }
    const uint32_t mult = (uint32_t)(1 << 16) / (uint32_t)divisor;
    const int post = length - (length % 8);
    for (c = 0; c < length; c++) {
        if (c < post) {
            /* SSE2: 16-bit abs (wraps for -32768), wrapping add, unsigned mulhi, sign restore */
            uint16_t v = (uint16_t)in[c];
            uint16_t s = (in[c] < 0) ? 0xFFFFu : 0u;
            uint16_t a = (uint16_t)((uint16_t)(v ^ s) - s);
            a = (uint16_t)(a + (uint16_t)mid);
            uint16_t q = (uint16_t)(((uint32_t)a * (uint32_t)(uint16_t)mult) >> 16);
            out[c] = (int16_t)(uint16_t)((uint16_t)(q ^ s) - s);
        } else {
            int value = in[c];
            if (value >= 0) {
                uint32_t r = (uint32_t)(value + mid) * mult;
                out[c] = sat16((int16_t)(uint16_t)(r >> 16));
            } else {
                value = -value;
                uint32_t r = (uint32_t)(value + mid) * mult;
                out[c] = sat16(-(int32_t)(int16_t)(uint16_t)(r >> 16));
            }
        }
    }
}

/* ------------------------------------------------------------------------- */
/* Vertical 2-6 + quantisation shared by the three fused forward variants.
 * L/H: horizontal low/high results, `rows` rows of ow columns (dense).
 * Codec/spatial.c:10166-10583 (and the identical bodies at :12942, :14726). */
static void fwd_vertical(const int16_t *L, const int16_t *H, int ow, int rows, int variant,
                         const int quant[4], int g,
                         int16_t *ll, int16_t *lh, int16_t *hl, int16_t *hh, int out_pitch)
{
    const int oh = rows / 2;
    const int post = ow - (ow % 8);
    const int op = out_pitch / (int)sizeof(int16_t);
    int16_t *bll = (int16_t *)malloc((size_t)ow * 2), *blh = (int16_t *)malloc((size_t)ow * 2);
    int16_t *bhl = (int16_t *)malloc((size_t)ow * 2), *bhh = (int16_t *)malloc((size_t)ow * 2);
    int r, c, k;
    for (r = 0; r < oh; r++) {
        /* window of six horizontal rows feeding output row r */
        int base = (r == 0) ? 0 : (r == oh - 1 ? rows - 6 : 2 * r - 2);
        const int16_t *l[6], *h[6];
        for (k = 0; k < 6; k++) { l[k] = L + (size_t)(base + k) * ow; h[k] = H + (size_t)(base + k) * ow; }
        for (c = 0; c < ow; c++) {
            if (r == 0) {                                   /* spatial.c:10166-10208 */
                bll[c] = sat16((int32_t)l[0][c] + l[1][c]);
                bhl[c] = sat16((5 * l[0][c] - 11 * l[1][c] + 4 * l[2][c] + 4 * l[3][c] - l[4][c] - l[5][c] + 4) >> 3);
                blh[c] = sat16((int32_t)h[0][c] + h[1][c]);
                bhh[c] = sat16((5 * h[0][c] - 11 * h[1][c] + 4 * h[2][c] + 4 * h[3][c] - h[4][c] - h[5][c] + 4) >> 3);
            } else if (r == oh - 1) {                       /* spatial.c:10516-10558 */
                bll[c] = sat16((int32_t)l[4][c] + l[5][c]);
                bhl[c] = sat16((11 * l[4][c] - 5 * l[5][c] - 4 * l[3][c] - 4 * l[2][c] + l[1][c] + l[0][c] + 4) >> 3);
                blh[c] = sat16((int32_t)h[4][c] + h[5][c]);
                bhh[c] = sat16((11 * h[4][c] - 5 * h[5][c] - 4 * h[3][c] - 4 * h[2][c] + h[1][c] + h[0][c] + 4) >> 3);
            } else if (c < post) {                          /* SSE2 loop, spatial.c:10290-10413 */
                int16_t s, s8;
                bll[c] = adds(l[2][c], l[3][c]);
                s = subs(0, l[0][c]); s = subs(s, l[1][c]);
                s8 = adds(0, l[2][c]); s8 = subs(s8, l[3][c]);
                s = adds(s, l[4][c]); s = adds(s, l[5][c]);
                s = adds(s, 4); s = sra16(s, 3);
                bhl[c] = adds(s, s8);
                blh[c] = adds(h[2][c], h[3][c]);
                s = subs(0, h[0][c]); s = subs(s, h[1][c]);
                s8 = adds(0, h[2][c]); s8 = subs(s8, h[3][c]);
                s = adds(s, h[4][c]); s = adds(s, h[5][c]);
                s = adds(s, 4); s = sra16(s, 3);
                bhh[c] = adds(s, s8);
            } else {                                        /* scalar tail, spatial.c:10421-10462 */
                int32_t s;
                bll[c] = sat16((int32_t)l[2][c] + l[3][c]);
                s = (-(int32_t)l[0][c] - l[1][c] + l[4][c] + l[5][c] + 4) >> 3;
                s += l[2][c] - l[3][c];
                bhl[c] = (variant == ORC_FWD_YUV) ? sat16(s) : wrap16(s);   /* :15054 vs :10443 */
                s = (int32_t)h[2][c] + h[3][c];
                blh[c] = (variant == ORC_FWD_YUV) ? sat16(s) : wrap16(s);
                s = (-(int32_t)h[0][c] - h[1][c] + h[4][c] + h[5][c] + 4) >> 3;
                s += h[2][c] - h[3][c];
                bhh[c] = (variant == ORC_FWD_YUV) ? sat16(s) : wrap16(s);
            }
        }
        /* only the plain variant quantises LL (if quant[0] > 1, spatial.c:10480); the YUV variant never does
         * (spatial.c:14845) and the V210 one has it compiled out (_QUANTIZE_SPATIAL_LOWPASS 0, spatial.c:70) */
        if (variant == ORC_FWD_PLAIN && quant[0] > 1)
            orc_quantize_row(bll, ll + (size_t)r * op, ow, quant[0], g);
        else
            memcpy(ll + (size_t)r * op, bll, (size_t)ow * 2);
        orc_quantize_row(blh, lh + (size_t)r * op, ow, quant[1], g);
        orc_quantize_row(bhl, hl + (size_t)r * op, ow, quant[2], g);
        orc_quantize_row(bhh, hh + (size_t)r * op, ow, quant[3], g);
    }
    free(bll); free(blh); free(bhl); free(bhh);
}

void orc_fwd_level(const int16_t *in, int in_pitch, int width, int height, int variant,
                   const int quant[4], int midpoint_prequant,
                   int16_t *ll, int16_t *lh, int16_t *hl, int16_t *hh, int out_pitch)
{
    const int ow = width / 2;
    int16_t *L = (int16_t *)malloc((size_t)ow * height * 2);
    int16_t *H = (int16_t *)malloc((size_t)ow * height * 2);
    int r;
    for (r = 0; r < height; r++)
        orc_fwd_row((const int16_t *)((const uint8_t *)in + (size_t)r * in_pitch),
                    L + (size_t)r * ow, H + (size_t)r * ow, width, variant == ORC_FWD_V210 ? 2 : 0);
    fwd_vertical(L, H, ow, height, variant, quant, midpoint_prequant, ll, lh, hl, hh, out_pitch);
    free(L); free(H);
}

/* Codec/convert.c:4667-5287 (YUYV) / :5289-5565 (UYVY), default branch: value << shift.
 * channel 0 = Y, 1 = V, 2 = U (convert.c:4793 "channel == 2 // U channel"). */
void orc_unpack_row_422(const uint8_t *in, int16_t *out, int width, int channel, int format, int shift)
{
    int i;
    const int yoff = (format == ORC_FMT_YUYV) ? 0 : 1;
    const int coff = (format == ORC_FMT_YUYV) ? 1 : 0;
    for (i = 0; i < width; i++) {
        int v;
        if (channel == 0) v = in[2 * i + yoff];
        else if (channel == 2) v = in[4 * i + coff];        /* U */
        else v = in[4 * i + coff + 2];                       /* V */
        out[i] = (int16_t)(v << shift);
    }
}

void orc_fwd_level_422(const uint8_t *frame, int frame_pitch, int width, int height, int channel,
                       int format, int precision, const int quant[4], int midpoint_prequant,
                       int16_t *ll, int16_t *lh, int16_t *hl, int16_t *hh, int out_pitch)
{
    const int ow = width / 2;
    int16_t *row = (int16_t *)malloc((size_t)width * 2);
    int16_t *L = (int16_t *)malloc((size_t)ow * height * 2);
    int16_t *H = (int16_t *)malloc((size_t)ow * height * 2);
    int r;
    for (r = 0; r < height; r++) {
        orc_unpack_row_422(frame + (size_t)r * frame_pitch, row, width, channel, format, precision - 8);
        orc_fwd_row(row, L + (size_t)r * ow, H + (size_t)r * ow, width, 0);   /* spatial.c:4005 */
    }
    fwd_vertical(L, H, ow, height, ORC_FWD_YUV, quant, midpoint_prequant, ll, lh, hl, hh, out_pitch);
    free(row); free(L); free(H);
}

/* ------------------------------------------------------------------------- */
/* Codec/decoder.c:20551-20626: the FSM tables are pre-multiplied by quant, i.e.
 * each decoded value v becomes (int16)(v * quant). */
void orc_dequantize_band(const int16_t *in, int16_t *out, int width, int height, int pitch, int quant)
{
    int r, c;
    const int p = pitch / 2;
    for (r = 0; r < height; r++)
        for (c = 0; c < width; c++)
            out[(size_t)r * p + c] = wrap16((int32_t)in[(size_t)r * p + c] * quant);
}

/* ------------------------------------------------------------------------- */
/* Inverse horizontal 2-6 on one row of (lowpass, highpass) -> 2*width samples.
 * descale 0: Codec/InvertHorizontalStrip16s.c:459-896
 * descale 2: Codec/InvertHorizontalStrip16s.c:1700-2166                       */
static void inv_row(const int16_t *l, const int16_t *h, int width, int descale, int16_t *out)
{
    const int last = width - 1;
    int post = last - (last % 8);
    int i;
    if (post == last) post -= 8;
    const int loop_runs = (post > 0);
    /* left border */
    {
        int32_t e = ((11 * l[0] - 4 * l[1] + l[2] + 4) >> 3) + h[0];
        int32_t o = ((5 * l[0] + 4 * l[1] - l[2] + 4) >> 3) - h[0];
        if (!descale) {
            e >>= 1; o >>= 1;
            /* inserted into an xmm lane (truncation) when the SIMD loop runs, else SATURATE */
            out[0] = loop_runs ? wrap16(e) : sat16(e);
            out[1] = loop_runs ? wrap16(o) : sat16(o);
        } else {
            if (loop_runs) {        /* insert_epi16 (truncate) then adds(out,out) : :1890-1893 */
                int16_t te = wrap16(e), to = wrap16(o);
                out[0] = adds(te, te); out[1] = adds(to, to);
            } else {
                out[0] = sat16(e << 1); out[1] = sat16(o << 1);
            }
        }
    }
    for (i = 1; i < last; i++) {
        if (i <= post) {            /* SSE2 loop */
            int16_t e = subs(l[i - 1], l[i + 1]);
            e = adds(e, 4); e = sra16(e, 3); e = adds(e, l[i]); e = adds(e, h[i]);
            int16_t o = subs(l[i + 1], l[i - 1]);
            o = adds(o, 4); o = sra16(o, 3); o = adds(o, l[i]); o = subs(o, h[i]);
            if (!descale) { out[2 * i] = sra16(e, 1); out[2 * i + 1] = sra16(o, 1); }
            else { out[2 * i] = adds(e, e); out[2 * i + 1] = adds(o, o); }
        } else if (!descale) {      /* 'short' accumulators: every step wraps, :786-822 */
            int16_t e = 0, o = 0;
            e = wrap16(e + l[i - 1]); e = wrap16(e - l[i + 1]); e = wrap16(e + 4); e = wrap16(e >> 3);
            e = wrap16(e + l[i]); e = wrap16(e + h[i]); e = wrap16(e >> 1);
            o = wrap16(o - l[i - 1]); o = wrap16(o + l[i + 1]); o = wrap16(o + 4); o = wrap16(o >> 3);
            o = wrap16(o + l[i]); o = wrap16(o - h[i]); o = wrap16(o >> 1);
            out[2 * i] = e; out[2 * i + 1] = o;
        } else {                    /* int accumulators, :2072-2112 */
            int32_t e = (((int32_t)l[i - 1] - l[i + 1] + 4) >> 3) + l[i] + h[i];
            int32_t o = ((-(int32_t)l[i - 1] + l[i + 1] + 4) >> 3) + l[i] - h[i];
            out[2 * i] = sat16(e << 1); out[2 * i + 1] = sat16(o << 1);
        }
    }
    /* right border */
    {
        int32_t e = ((5 * l[last] + 4 * l[last - 1] - l[last - 2] + 4) >> 3) + h[last];
        int32_t o = ((11 * l[last] - 4 * l[last - 1] + l[last - 2] + 4) >> 3) - h[last];
        if (!descale) { out[2 * last] = sat16(e >> 1); out[2 * last + 1] = sat16(o >> 1); }
        else { out[2 * last] = sat16(e << 1); out[2 * last + 1] = sat16(o << 1); }
    }
}

/* Vertical inverse of one band pair (low band a, high band b) for band row r:
 * writes the even and odd intermediate rows.
 * descale 0: Codec/spatial.c:21980-22318 ; descale 2: Codec/spatial.c:22520-23150 */
static void inv_vertical_pair(const int16_t *a, const int16_t *b, int pitch_px, int width, int height,
                              int r, int descale, int16_t *even, int16_t *odd)
{
    const int post = width - (width % 8);
    int c;
    for (c = 0; c < width; c++) {
        const int16_t hv = b[(size_t)r * pitch_px + c];
        if (r == 0) {
            int32_t a0 = a[c], a1 = a[pitch_px + c], a2 = a[2 * pitch_px + c];
            int32_t e = (11 * a0 - 4 * a1 + a2 + 4) >> 3; e += hv; e >>= 1;
            int32_t o = (5 * a0 + 4 * a1 - a2 + 4) >> 3; o -= hv; o >>= 1;
            even[c] = sat16(e); odd[c] = sat16(o);
        } else if (r == height - 1) {
            int32_t a0 = a[(size_t)r * pitch_px + c], a1 = a[(size_t)(r - 1) * pitch_px + c], a2 = a[(size_t)(r - 2) * pitch_px + c];
            int32_t e = (5 * a0 + 4 * a1 - a2 + 4) >> 3; e += hv; e >>= 1;
            int32_t o = (11 * a0 - 4 * a1 + a2 + 4) >> 3; o -= hv; o >>= 1;
            even[c] = sat16(e); odd[c] = sat16(o);
        } else {
            const int16_t l0 = a[(size_t)(r - 1) * pitch_px + c], l1 = a[(size_t)r * pitch_px + c], l2 = a[(size_t)(r + 1) * pitch_px + c];
            if (c < post && !descale) {         /* 16-bit saturating SSE2, spatial.c:22124-22190 */
                int16_t e = subs(l0, l2); e = adds(e, 4); e = sra16(e, 3); e = adds(e, l1);
                e = adds(e, hv); e = sra16(e, 1);
                int16_t o = subs(0, l0); o = adds(o, l2); o = adds(o, 4); o = sra16(o, 3); o = adds(o, l1);
                o = subs(o, hv); o = sra16(o, 1);
                even[c] = e; odd[c] = o;
            } else {                            /* int32 (+packs) : descale SIMD and both scalar tails */
                int32_t e = (((int32_t)l0 - l2 + 4) >> 3) + l1 + hv; e >>= 1;
                int32_t o = ((-(int32_t)l0 + l2 + 4) >> 3) + l1 - hv; o >>= 1;
                even[c] = sat16(e); odd[c] = sat16(o);
            }
        }
    }
}

void orc_inv_level(const int16_t *ll, const int16_t *lh, const int16_t *hl, const int16_t *hh,
                   int band_pitch, int width, int height, int descale,
                   int16_t *out, int out_pitch)
{
    const int bp = band_pitch / 2, op = out_pitch / 2;
    int16_t *el = (int16_t *)malloc((size_t)width * 2), *ol = (int16_t *)malloc((size_t)width * 2);
    int16_t *eh = (int16_t *)malloc((size_t)width * 2), *oh = (int16_t *)malloc((size_t)width * 2);
    int r;
    for (r = 0; r < height; r++) {
        inv_vertical_pair(ll, hl, bp, width, height, r, descale, el, ol);   /* left two bands  */
        inv_vertical_pair(lh, hh, bp, width, height, r, descale, eh, oh);   /* right two bands */
        inv_row(el, eh, width, descale, out + (size_t)(2 * r) * op);
        inv_row(ol, oh, width, descale, out + (size_t)(2 * r + 1) * op);
    }
    free(el); free(ol); free(eh); free(oh);
}

/* ------------------------------------------------------------------------- */
/* Interlaced (field) transform, see cfhd_oracle.h. */
static void fields_hrow(const int32_t *x, int width, int32_t *low, int32_t *high)
{
    const int m = width / 2;
    int i;
    for (i = 0; i < m; i++) low[i] = x[2 * i] + x[2 * i + 1];
    for (i = 1; i < m - 1; i++)
        high[i] = ((-x[2 * i - 2] - x[2 * i - 1] + x[2 * i + 2] + x[2 * i + 3] + 4) >> 3) + x[2 * i] - x[2 * i + 1];
    high[0] = (5 * x[0] - 11 * x[1] + 4 * x[2] + 4 * x[3] - x[4] - x[5] + 4) >> 3;                      /* spatial.c:5371-5379 */
    high[m - 1] = (11 * x[width - 2] - 5 * x[width - 1] - 4 * x[width - 3] - 4 * x[width - 4]
                   + x[width - 5] + x[width - 6] + 4) >> 3;                                             /* spatial.c:5802-5810 */
}

void orc_fwd_fields_422(const uint8_t *frame, int frame_pitch, int width, int height, int channel,
                        int format, int precision, const int quant[4], int midpoint_prequant,
                        int16_t *ll, int16_t *lh, int16_t *hl, int16_t *hh, int out_pitch)
{
    const int m = width / 2, op = out_pitch / 2;
    int16_t *e = (int16_t *)malloc((size_t)width * 2), *o = (int16_t *)malloc((size_t)width * 2);
    int32_t *tl = (int32_t *)malloc((size_t)width * 4), *th = (int32_t *)malloc((size_t)width * 4);
    int32_t *a = (int32_t *)malloc((size_t)m * 4), *b = (int32_t *)malloc((size_t)m * 4);
    int16_t *tmp = (int16_t *)malloc((size_t)m * 2);
    int r, i;
    for (r = 0; r < height / 2; r++) {
        orc_unpack_row_422(frame + (size_t)(2 * r) * frame_pitch, e, width, channel, format, precision - 8);
        orc_unpack_row_422(frame + (size_t)(2 * r + 1) * frame_pitch, o, width, channel, format, precision - 8);
        for (i = 0; i < width; i++) { tl[i] = (int32_t)e[i] + o[i]; th[i] = (int32_t)o[i] - e[i]; }
        fields_hrow(tl, width, a, b);
        for (i = 0; i < m; i++) { ll[(size_t)r * op + i] = wrap16(a[i]); tmp[i] = wrap16(b[i]); }
        orc_quantize_row(tmp, lh + (size_t)r * op, m, quant[1], midpoint_prequant);
        fields_hrow(th, width, a, b);
        {
            const int d = quant[2];
            const int mid = (d > 1 && midpoint_prequant >= 2 && midpoint_prequant < 9) ? d / midpoint_prequant : 0;
            const int mult = (d > 1) ? 65536 / d : 0;
            int32_t prev = 0;
            for (i = 0; i < m; i++) {
                int32_t q = a[i];
                if (d > 1) { const int32_t mag = ((q < 0 ? -q : q) + mid) * mult >> 16; q = q < 0 ? -mag : mag; }
                hl[(size_t)r * op + i] = wrap16(q - prev);
                prev = q;
            }
        }
        for (i = 0; i < m; i++) tmp[i] = wrap16(b[i]);
        orc_quantize_row(tmp, hh + (size_t)r * op, m, quant[3], midpoint_prequant);
    }
    free(e); free(o); free(tl); free(th); free(a); free(b); free(tmp);
}

/* Planar form of the field transform (Codec/filter.c:273 FilterFrameQuant16s, the path of 16-bit / 10-bit 4:2:2 sources
 * after their conversion to planes): as above on one int16 plane, except that LL and LH come out of
 * spatial.c:5826 FilterHorizontalRowQuant16s:
 *   outputs of the 16-column SSE2 loop (i < (width - width % 16) / 2): 16-bit |x|, + divisor / 2 (:5856-5857, no "- 1",
 *     whatever g is), unsigned mulhi by 65536 / divisor, sign restored (:6082-6140); LL only when its divisor > 1 (:6074);
 *   outputs of the scalar tail (:6192-6225) and the last output, which is redone with the border filter (:6232-6266):
 *     highpass sign * ((|x| * (65536 / divisor)) >> 16) -- NO midpoint; lowpass (x * (65536 / divisor)) >> 16. */
static int16_t fields_quant_half(int32_t v, int d, int simd, int lowpass)
{
    if (d <= 1) return wrap16(v);
    if (!simd) {
        const int32_t mult = 65536 / d;
        if (lowpass) return sat16((int32_t)(((int64_t)v * mult) >> 16));
        return sat16(v < 0 ? -(int32_t)(((int64_t)(-v) * mult) >> 16) : (int32_t)(((int64_t)v * mult) >> 16));
    }
    {
        const int16_t x = wrap16(v);
        const uint16_t sgn = x < 0 ? 0xFFFFu : 0u;
        uint16_t a = (uint16_t)(((uint16_t)x ^ sgn) - sgn);
        a = (uint16_t)(a + (uint16_t)(d / 2));
        {
            const uint16_t q = (uint16_t)(((uint32_t)a * (uint32_t)(uint16_t)(65536 / d)) >> 16);
            return (int16_t)(uint16_t)((uint16_t)(q ^ sgn) - sgn);
        }
    }
}

void orc_fwd_fields_plane(const int16_t *plane, int plane_pitch, int width, int height,
                          const int quant[4], int midpoint_prequant,
                          int16_t *ll, int16_t *lh, int16_t *hl, int16_t *hh, int out_pitch)
{
    const int m = width / 2, op = out_pitch / 2, pp = plane_pitch / 2;
    int32_t *tl = (int32_t *)malloc((size_t)width * 4), *th = (int32_t *)malloc((size_t)width * 4);
    int32_t *a = (int32_t *)malloc((size_t)m * 4), *b = (int32_t *)malloc((size_t)m * 4);
    int16_t *tmp = (int16_t *)malloc((size_t)m * 2);
    int r, i;
    for (r = 0; r < height / 2; r++) {
        const int16_t *e = plane + (size_t)(2 * r) * pp, *o = plane + (size_t)(2 * r + 1) * pp;
        for (i = 0; i < width; i++) { tl[i] = (int32_t)e[i] + o[i]; th[i] = (int32_t)o[i] - e[i]; }
        fields_hrow(tl, width, a, b);
        for (i = 0; i < m; i++) {
            const int simd = (2 * i < width - width % 16) && (i != m - 1);
            ll[(size_t)r * op + i] = fields_quant_half(a[i], quant[0], simd, 1);
            lh[(size_t)r * op + i] = fields_quant_half(b[i], quant[1], simd, 0);
        }
        fields_hrow(th, width, a, b);
        {
            const int d = quant[2];
            const int mid = (d > 1 && midpoint_prequant >= 2 && midpoint_prequant < 9) ? d / midpoint_prequant : 0;
            const int mult = (d > 1) ? 65536 / d : 0;
            int32_t prev = 0;
            for (i = 0; i < m; i++) {
                int32_t q = a[i];
                if (d > 1) { const int32_t mag = ((q < 0 ? -q : q) + mid) * mult >> 16; q = q < 0 ? -mag : mag; }
                hl[(size_t)r * op + i] = wrap16(q - prev);
                prev = q;
            }
        }
        for (i = 0; i < m; i++) tmp[i] = wrap16(b[i]);
        orc_quantize_row(tmp, hh + (size_t)r * op, m, quant[3], midpoint_prequant);
    }
    free(tl); free(th); free(a); free(b); free(tmp);
}

static void fields_hinv(const int16_t *l, const int16_t *h, int n, int32_t *out)
{
    int i;
    for (i = 1; i < n - 1; i++) {
        out[2 * i] = ((((int32_t)l[i - 1] - l[i + 1] + 4) >> 3) + l[i] + h[i]) >> 1;
        out[2 * i + 1] = (((-(int32_t)l[i - 1] + l[i + 1] + 4) >> 3) + l[i] - h[i]) >> 1;
    }
    out[0] = (((11 * (int32_t)l[0] - 4 * l[1] + l[2] + 4) >> 3) + h[0]) >> 1;
    out[1] = (((5 * (int32_t)l[0] + 4 * l[1] - l[2] + 4) >> 3) - h[0]) >> 1;
    out[2 * n - 2] = (((5 * (int32_t)l[n - 1] + 4 * l[n - 2] - l[n - 3] + 4) >> 3) + h[n - 1]) >> 1;
    out[2 * n - 1] = (((11 * (int32_t)l[n - 1] - 4 * l[n - 2] + l[n - 3] + 4) >> 3) - h[n - 1]) >> 1;
}

void orc_inv_fields(const int16_t *ll, const int16_t *lh, const int16_t *hl, const int16_t *hh,
                    int band_pitch, int width, int height, int16_t *out, int out_pitch)
{
    const int bp = band_pitch / 2, op = out_pitch / 2;
    int32_t *tl = (int32_t *)malloc((size_t)width * 8), *th = (int32_t *)malloc((size_t)width * 8);
    int r, i;
    for (r = 0; r < height; r++) {
        fields_hinv(ll + (size_t)r * bp, lh + (size_t)r * bp, width, tl);
        fields_hinv(hl + (size_t)r * bp, hh + (size_t)r * bp, width, th);
        for (i = 0; i < 2 * width; i++) {
            out[(size_t)(2 * r) * op + i] = wrap16((tl[i] - th[i]) >> 1);
            out[(size_t)(2 * r + 1) * op + i] = wrap16((tl[i] + th[i]) >> 1);
        }
    }
    free(tl); free(th);
}

/* ------------------------------------------------------------------------- */
/* Two-frame GOP temporal Haar, see cfhd_oracle.h. */
void orc_temporal_fwd(const int16_t *a, const int16_t *b, int in_pitch, int width, int height,
                      int16_t *low, int16_t *high, int out_pitch)
{
    int r, c;
    for (r = 0; r < height; r++) {
        const int16_t *x = (const int16_t *)((const uint8_t *)a + (size_t)r * in_pitch);
        const int16_t *y = (const int16_t *)((const uint8_t *)b + (size_t)r * in_pitch);
        int16_t *l = (int16_t *)((uint8_t *)low + (size_t)r * out_pitch);
        int16_t *h = (int16_t *)((uint8_t *)high + (size_t)r * out_pitch);
        for (c = 0; c < width; c++) { l[c] = adds(x[c], y[c]); h[c] = subs(y[c], x[c]); }
    }
}

void orc_temporal_inv(const int16_t *low, const int16_t *high, int in_pitch, int width, int height, int precision,
                      int16_t *a, int16_t *b, int out_pitch)
{
    const int post = width - (width % 40);
    int r, c;
    for (r = 0; r < height; r++) {
        const int16_t *l = (const int16_t *)((const uint8_t *)low + (size_t)r * in_pitch);
        const int16_t *h = (const int16_t *)((const uint8_t *)high + (size_t)r * in_pitch);
        int16_t *x = (int16_t *)((uint8_t *)a + (size_t)r * out_pitch);
        int16_t *y = (int16_t *)((uint8_t *)b + (size_t)r * out_pitch);
        for (c = 0; c < post; c++) {
            const int16_t ht = (precision == 8) ? (int16_t)((c + r + 1) & 1) : 0;
            x[c] = sra16(subs(l[c], h[c]), 1);
            y[c] = sra16(adds(adds(l[c], h[c]), ht), 1);
        }
        for (; c < width; c++) {
            const int t = (precision == 8) ? ((c + r) & 1) : 0;
            x[c] = wrap16(((int32_t)l[c] - h[c]) >> 1);
            y[c] = wrap16(((int32_t)l[c] + h[c] + t) >> 1);
        }
    }
}
