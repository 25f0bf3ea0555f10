/* cfhd_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C, scalar) of the CineForm transform hot path used as
 * the parity checker for the CUDA kernels.  It is NOT product code: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load liboracle.so.  The product library
 * (cineform-sdk_b200/csrc) never links or calls it and has no CPU fallback.
 *
 * Parity status: PINNED -- every function below is validated bit-for-bit
 * against the unmodified reference compiled in place (oracle/_ref, see
 * Makefile and tests/test_oracle_vs_ref.py), on Qbist frames and on
 * adversarial inputs that exercise the reference's saturating-SIMD /
 * int32-scalar column split.
 *
 * Each function cites the reference file:line it restates (paths relative to
 * the reference tree, commit 11574d02).
 */
#ifndef CFHD_ORACLE_H
#define CFHD_ORACLE_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* variants of the fused forward level (differences only matter on int16 overflow) */
#define ORC_FWD_PLAIN 0  /* FilterSpatialQuant16s      Codec/spatial.c:10026 (prescale 0)        */
#define ORC_FWD_V210  1  /* FilterSpatialV210Quant16s  Codec/spatial.c:12942 (prescale 2)        */
#define ORC_FWD_YUV   2  /* FilterSpatialYUVQuant16s   Codec/spatial.c:14726 (packed 4:2:2 rows) */

#define ORC_FMT_YUYV 0
#define ORC_FMT_UYVY 1

/* Codec/spatial.c:253 FilterHorizontalRow16s (prescale==0) and
 * Codec/spatial.c:3669 FilterHorizontalRow10bit16s (prescale==2). width even, >= 18. */
void orc_fwd_row(const int16_t *in, int16_t *low, int16_t *high, int width, int prescale);

/* Codec/quantize.c:1395 QuantizeRow16sTo16s. midpoint_prequant is the value of the
 * reference's global g_midpoint_prequant (2 for the default pre-emphasis). */
void orc_quantize_row(const int16_t *in, int16_t *out, int length, int divisor, int midpoint_prequant);

/* One fused forward level on an int16 plane: horizontal 2-6, vertical 2-6, quantise.
 * Pitches are in BYTES (as in the reference). quant[0..3] = LL,LH,HL,HH divisors. */
void orc_fwd_level(const int16_t *in, int in_pitch, int width, int height, int variant,
                   const int quant[4], int midpoint_prequant,
                   int16_t *ll, int16_t *lh, int16_t *hl, int16_t *hh, int out_pitch);

/* Codec/convert.c:4667 UnpackRowYUV16s default branch (value << shift), one channel
 * (0 = Y, 1 = V, 2 = U as the encoder numbers them) of one packed row. */
void orc_unpack_row_422(const uint8_t *in, int16_t *out, int width, int channel, int format, int shift);

/* Level 1 of a packed 8-bit 4:2:2 frame for one channel (Codec/wavelet.c:2823 +
 * Codec/spatial.c:14726).  width/height are the CHANNEL's input dimensions. */
void orc_fwd_level_422(const uint8_t *frame, int frame_pitch, int width, int height, int channel,
                       int format, int precision, const int quant[4], int midpoint_prequant,
                       int16_t *ll, int16_t *lh, int16_t *hl, int16_t *hh, int out_pitch);

/* Codec/decoder.c:20551 DeQuantFSM semantics applied to a dense band: c = (int16)(v*quant). */
void orc_dequantize_band(const int16_t *in, int16_t *out, int width, int height, int pitch, int quant);

/* Inverse level (bands already dequantised): Codec/spatial.c:21877 InvertSpatialQuant16s +
 * Codec/InvertHorizontalStrip16s.c:459 (descale==0) or Codec/spatial.c:22414
 * InvertSpatialQuantDescale16s + InvertHorizontalStrip16s.c:1700 (descale==2).
 * width/height = band dimensions; output is 2*width x 2*height. Pitches in bytes. */
void orc_inv_level(const int16_t *ll, const int16_t *lh, const int16_t *hl, const int16_t *hh,
                   int band_pitch, int width, int height, int descale,
                   int16_t *out, int out_pitch);

/* Reduced-resolution decode, 4:2:2 sources, 8-bit packed output from the lowpass images of the three channels
 * (y = channel 0, v = channel 1, u = channel 2 in the encoder's numbering; bytes Y0 U Y1 V or U Y0 V Y1):
 *   half (unsigned_shift = 0): Codec/frame.c:11742 ConvertLowpass16s10bitToYUV scalar loop (:11880-11892; the MMX
 *     dither branch is compiled out on x86-64):  sat_u8(ll >> shift), shift = PRESCALE_LUMA10 = 4;
 *   quarter (unsigned_shift = 1): Codec/temporal.c:11362 CopyQuarterRowToBuffer: the coefficients are read as
 *     uint16 and `_mm_srli_epi16(x, 4)` then `_mm_packus_epi16`.  (Its SIMD loop stores the first 8 pixels of every
 *     16 in YUYV order even when UYVY is requested, :11395-11398; that quirk is NOT restated: uyvy=1 gives UYVY.)
 * width/height = luma lowpass dimensions; pitches in bytes. */
void orc_lowpass_to_422(const int16_t *y, int y_pitch, const int16_t *v, int v_pitch, const int16_t *u, int u_pitch,
                        int width, int height, int shift, int unsigned_shift, int uyvy,
                        uint8_t *out, int out_pitch);

/* ---- interlaced sources: the frame (field) transform replaces the spatial transform at level 1 ----
 * Forward, packed 8-bit 4:2:2, one channel (Codec/wavelet.c:6076 TransformForwardFrameYUV, called from
 * Codec/encoder.c:2976 when the encoder was opened with progressive = 0):
 *   per row pair: temporal low = even + odd, temporal high = odd - even on the samples << (precision - 8)
 *   (temporal.c:1568 FilterTemporalRowYUYVChannelTo16s), then the horizontal 2-6 filter on both rows
 *   (spatial.c:253 FilterHorizontalRow16s):  LL = low(t_low) unquantised,  LH = Q(high(t_low)),
 *   HH = Q(high(t_high)) with the ordinary quantiser (quantize.c:1395), and the HL band is
 *   "difference filtered" (spatial.c:5327 FilterHorizontalRowScaled16sDifferenceFiltered, DIFFERENCE_CODING 1,
 *   codec.h:161):  q[i] = sign * (((|low(t_high)[i]| + divisor / g) * (65536 / divisor)) >> 16)   (no "-1" on the
 *   midpoint, :5356-5358)  and  HL[i] = q[i] - q[i-1], HL[0] = q[0].
 * The arithmetic is restated in int32 and wrapped to int16 on store: identical to the reference's saturating SSE2
 * chains for every 8-bit source (|t| <= 2040, far inside int16).  width/height = the CHANNEL's input dimensions. */
void orc_fwd_fields_422(const uint8_t *frame, int frame_pitch, int width, int height, int channel,
                        int format, int precision, const int quant[4], int midpoint_prequant,
                        int16_t *ll, int16_t *lh, int16_t *hl, int16_t *hh, int out_pitch);

/* planar form (Codec/filter.c:273 FilterFrameQuant16s): one int16 plane; LL / LH quantised by FilterHorizontalRowQuant16s */
void orc_fwd_fields_plane(const int16_t *plane, int plane_pitch, int width, int height,
                          const int quant[4], int midpoint_prequant,
                          int16_t *ll, int16_t *lh, int16_t *hl, int16_t *hh, int out_pitch);

/* Inverse of the above on DECODER-SIDE bands, i.e. dequantised and with HL already integrated along each row
 * (Codec/decoder.c:20822-20836 `line[x] += line[x-1]` after the FSM decode):
 *   t_low = hinv(LL, LH), t_high = hinv(HL, HH)  (Codec/decoder.c:21493 TransformInverseFrameToYUV ->
 *   InvertHorizontalRow16s8sTo16sBuffered: interior ((l[i-1] - l[i+1] + 4) >> 3 + l[i] + h[i]) >> 1 etc., borders
 *   (11,-4,1)/(5,4,-1)),  even row = (t_low - t_high) >> 1, odd row = (t_low + t_high) >> 1
 *   (temporal.c:3741 InvertInterlaced16s / InvertInterlacedRow16s10bitToYUV).
 * width/height = band dimensions; out is 2*width x 2*height int16 at the codec precision. */
void orc_inv_fields(const int16_t *ll, const int16_t *lh, const int16_t *hl, const int16_t *hh,
                    int band_pitch, int width, int height, int16_t *out, int out_pitch);

/* ---- two-frame GOP: temporal Haar between two int16 planes (the level-1 lowpass images of frames A and B) ----
 * Forward, Codec/temporal.c:498 FilterTemporal16s (the 16-bit branch :603-645; the _HIGHPASS_8S branch is compiled
 * out): low = adds(a, b), high = subs(b, a), saturating, width % 16 == 0 (asserted there). */
void orc_temporal_fwd(const int16_t *a, const int16_t *b, int in_pitch, int width, int height,
                      int16_t *low, int16_t *high, int out_pitch);
/* Inverse, Codec/temporal.c:9402 InvertTemporalQuant16s (the coefficients arrive dequantised; its quantisation
 * arguments are unused): SSE2 part, columns < width - width % 40: a = subs(low, high) >> 1,
 * b = adds(adds(low, high), halftone) >> 1; scalar tail in int: a = (low - high) >> 1, b = (low + high + t) >> 1 stored
 * as int16.  precision 8 only: halftone = (column + row + 1) & 1 in the SSE2 part (set_epi16 patterns :9437-9442) but
 * t = (column + row) & 1 in the tail (:9625); 0 for precision >= 10. */
void orc_temporal_inv(const int16_t *low, const int16_t *high, int in_pitch, int width, int height, int precision,
                      int16_t *a, int16_t *b, int out_pitch);

/* 3-level pyramid helpers are composed in Python (tests/) from the calls above. */

int orc_version(void);

#ifdef __cplusplus
}
#endif
#endif
