// ref_probe.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Our own shim, compiled together with the UNMODIFIED reference sources (in
// place, from /root/reference) into oracle/_ref/libcfhd_ref.so.  It exposes the
// reference's coefficient-level functions of the transform path through a flat
// C ABI so that tests can (1) pin oracle/cfhd_oracle.c against the real thing and
// (2) use the real thing as the CPU baseline.  It contains no codec logic: each
// entry point only marshals buffers (16-byte aligned copies, as the SSE2 loads
// in the reference require) and calls the cited reference function.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

extern "C" {
#include "config.h"
#include "encoder.h"
#include "wavelet.h"
#include "spatial.h"
#include "quantize.h"
#include "convert.h"
#include "frame.h"
#include "image.h"
#include "bitstream.h"
#include "codec.h"
#include "decoder.h"
}
#include "qbist.h"
#include "CFHDTypes.h"
#include "CFHDDecoder.h"

extern "C" int g_midpoint_prequant;   // Codec/quantize.c:183
extern "C" void FilterHorizontalRow10bit16s(PIXEL *input, PIXEL *lowpass, PIXEL *highpass, int width, PIXEL *buffer);

namespace {
struct Aligned {
    void *p = nullptr;
    explicit Aligned(size_t n) { if (posix_memalign(&p, 64, n ? n : 64)) p = nullptr; else memset(p, 0, n ? n : 64); }
    ~Aligned() { free(p); }
    template <class T> T *as() { return (T *)p; }
};
inline int align16(int x) { return (x + 15) & ~15; }
inline size_t align64(size_t x) { return (x + 63) & ~(size_t)63; }

// copy a (pitch-strided) 2-D array into / out of an aligned, 16-byte-pitched scratch
void copy_in(uint8_t *dst, int dpitch, const uint8_t *src, int spitch, int rowbytes, int rows) {
    for (int r = 0; r < rows; r++) memcpy(dst + (size_t)r * dpitch, src + (size_t)r * spitch, rowbytes);
}
}  // namespace

extern "C" {

int ref_probe_version(void) { return 2; }

// Codec/spatial.c:253 / :3669
void ref_fwd_row(const int16_t *in, int16_t *low, int16_t *high, int width, int prescale)
{
    Aligned a((size_t)width * 2 + 64), l((size_t)width + 64), h((size_t)width + 64), b((size_t)width * 2 + 64);
    memcpy(a.p, in, (size_t)width * 2);
    if (prescale) FilterHorizontalRow10bit16s(a.as<PIXEL>(), l.as<PIXEL>(), h.as<PIXEL>(), width, b.as<PIXEL>());
    else FilterHorizontalRow16s(a.as<PIXEL>(), l.as<PIXEL>(), h.as<PIXEL>(), width);
    memcpy(low, l.p, (size_t)width);
    memcpy(high, h.p, (size_t)width);
}

// Codec/quantize.c:1395
void ref_quantize_row(const int16_t *in, int16_t *out, int length, int divisor, int midpoint_prequant)
{
    Aligned a((size_t)length * 2 + 64), o((size_t)length * 2 + 64);
    memcpy(a.p, in, (size_t)length * 2);
    g_midpoint_prequant = midpoint_prequant;
    QuantizeRow16sTo16s(a.as<PIXEL>(), o.as<PIXEL>(), length, divisor);
    memcpy(out, o.p, (size_t)length * 2);
}

// variant 0: FilterSpatialQuant16s (spatial.c:10026); 1: FilterSpatialV210Quant16s (:12942)
void ref_fwd_level(const int16_t *in, int in_pitch, int width, int height, int variant,
                   const int quant[4], int midpoint_prequant,
                   int16_t *ll, int16_t *lh, int16_t *hl, int16_t *hh, int out_pitch)
{
    const int ow = width / 2, oh = height / 2;
    const int ip = align16(width * 2), op = align16(ow * 2);
    Aligned ain((size_t)ip * height), b0((size_t)op * oh), b1((size_t)op * oh), b2((size_t)op * oh), b3((size_t)op * oh);
    const size_t bufsize = 32 * align64((size_t)width * 2) + 4096;
    Aligned scratch(bufsize);
    copy_in(ain.as<uint8_t>(), ip, (const uint8_t *)in, in_pitch, width * 2, height);
    int q[4] = {quant[0], quant[1], quant[2], quant[3]};
    ROI roi = {width, height};
    g_midpoint_prequant = midpoint_prequant;
    if (variant == 1)
        FilterSpatialV210Quant16s(ain.as<PIXEL>(), ip, b0.as<PIXEL>(), op, b1.as<PIXEL>(), op, b2.as<PIXEL>(), op,
                                  b3.as<PIXEL>(), op, scratch.as<PIXEL>(), bufsize, roi, q);
    else
        FilterSpatialQuant16s(ain.as<PIXEL>(), ip, b0.as<PIXEL>(), op, b1.as<PIXEL>(), op, b2.as<PIXEL>(), op,
                              b3.as<PIXEL>(), op, scratch.as<PIXEL>(), bufsize, roi, q);
    copy_in((uint8_t *)ll, out_pitch, b0.as<uint8_t>(), op, ow * 2, oh);
    copy_in((uint8_t *)lh, out_pitch, b1.as<uint8_t>(), op, ow * 2, oh);
    copy_in((uint8_t *)hl, out_pitch, b2.as<uint8_t>(), op, ow * 2, oh);
    copy_in((uint8_t *)hh, out_pitch, b3.as<uint8_t>(), op, ow * 2, oh);
}

// Codec/spatial.c:14726 FilterSpatialYUVQuant16s for one channel of a packed 4:2:2 frame.
// width = channel input width (luma: frame width; chroma: frame width / 2). format: 0 YUYV, 1 UYVY.
void ref_fwd_level_422(const uint8_t *frame, int frame_pitch, int width, int height, int channel,
                       int format, int precision, const int quant[4], int midpoint_prequant,
                       int16_t *ll, int16_t *lh, int16_t *hl, int16_t *hh, int out_pitch)
{
    const int ow = width / 2, oh = height / 2;
    const int frame_width = (channel == 0) ? width : 2 * width;
    const int ip = align16(frame_width * 2), op = align16(ow * 2);
    Aligned ain((size_t)ip * height + 64), b0((size_t)op * oh), b1((size_t)op * oh), b2((size_t)op * oh), b3((size_t)op * oh);
    const size_t bufsize = 40 * align64((size_t)frame_width * 2) + 4096;
    Aligned scratch(bufsize);
    copy_in(ain.as<uint8_t>(), ip, frame, frame_pitch, frame_width * 2, height);
    int q[4] = {quant[0], quant[1], quant[2], quant[3]};
    ROI roi = {width, height};
    FRAME_INFO info;
    memset(&info, 0, sizeof(info));
    info.width = frame_width; info.height = height;
    info.format = format ? COLOR_FORMAT_UYVY : COLOR_FORMAT_YUYV;
    g_midpoint_prequant = midpoint_prequant;
    FilterSpatialYUVQuant16s(ain.as<uint8_t>(), ip, b0.as<PIXEL>(), op, b1.as<PIXEL>(), op, b2.as<PIXEL>(), op,
                             b3.as<PIXEL>(), op, scratch.as<PIXEL>(), bufsize, roi, channel, q, &info,
                             precision, 0, 0);
    copy_in((uint8_t *)ll, out_pitch, b0.as<uint8_t>(), op, ow * 2, oh);
    copy_in((uint8_t *)lh, out_pitch, b1.as<uint8_t>(), op, ow * 2, oh);
    copy_in((uint8_t *)hl, out_pitch, b2.as<uint8_t>(), op, ow * 2, oh);
    copy_in((uint8_t *)hh, out_pitch, b3.as<uint8_t>(), op, ow * 2, oh);
}

// descale 0: InvertSpatialQuant16s (spatial.c:21877); descale 2: InvertSpatialQuantDescale16s (:22414)
void ref_inv_level(const int16_t *ll, const int16_t *lh, const int16_t *hl, const int16_t *hh,
                   int band_pitch, int width, int height, int descale, int16_t *out, int out_pitch)
{
    const int bp = align16(width * 2), op = align16(width * 4);
    // the SIMD loops read up to 8 coefficients past the row end and 2 rows below: pad generously
    const size_t bsz = (size_t)bp * (height + 4) + 256;
    Aligned b0(bsz), b1(bsz), b2(bsz), b3(bsz), o((size_t)op * height * 2 + 256);
    const size_t bufsize = 16 * (size_t)align16(width * 2) + 4096;
    Aligned scratch(bufsize);
    copy_in(b0.as<uint8_t>(), bp, (const uint8_t *)ll, band_pitch, width * 2, height);
    copy_in(b1.as<uint8_t>(), bp, (const uint8_t *)lh, band_pitch, width * 2, height);
    copy_in(b2.as<uint8_t>(), bp, (const uint8_t *)hl, band_pitch, width * 2, height);
    copy_in(b3.as<uint8_t>(), bp, (const uint8_t *)hh, band_pitch, width * 2, height);
    ROI roi = {width, height};
    int q[4] = {1, 1, 1, 1};
    if (descale)
        InvertSpatialQuantDescale16s(b0.as<PIXEL>(), bp, b1.as<PIXEL>(), bp, b2.as<PIXEL>(), bp, b3.as<PIXEL>(), bp,
                                     o.as<PIXEL>(), op, roi, scratch.as<PIXEL>(), bufsize, descale, q);
    else
        InvertSpatialQuant16s(b0.as<PIXEL>(), bp, b1.as<PIXEL>(), bp, b2.as<PIXEL>(), bp, b3.as<PIXEL>(), bp,
                              o.as<PIXEL>(), op, roi, scratch.as<PIXEL>(), bufsize, q);
    copy_in((uint8_t *)out, out_pitch, o.as<uint8_t>(), op, width * 4, height * 2);
}

// GOP-2 temporal Haar between two int16 planes (level-1 lowpass images of frame A and frame B):
// Codec/temporal.c:498 FilterTemporal16s (live branch :603-645) and Codec/temporal.c:9402 InvertTemporalQuant16s.
extern "C" void FilterTemporal16s(PIXEL *field1, int pitch1, PIXEL *field2, int pitch2,
                                  PIXEL *lowpass, int lowpass_pitch, PIXEL *highpass, int highpass_pitch, ROI roi);
extern "C" void InvertTemporalQuant16s(PIXEL *lowpass, int lowpass_quantization, int lowpass_pitch,
                                       PIXEL *highpass, int highpass_quantization, int highpass_pitch,
                                       PIXEL *field1, int pitch1, PIXEL *field2, int pitch2, ROI roi,
                                       PIXEL *buffer, size_t buffer_size, int precision);
void ref_temporal_fwd(const int16_t *a, const int16_t *b, int in_pitch, int width, int height,
                      int16_t *low, int16_t *high, int out_pitch)
{
    const int pp = align16(width * 2) + 64;
    Aligned x((size_t)pp * height + 64), y((size_t)pp * height + 64), l((size_t)pp * height + 64), h((size_t)pp * height + 64);
    copy_in(x.as<uint8_t>(), pp, (const uint8_t *)a, in_pitch, width * 2, height);
    copy_in(y.as<uint8_t>(), pp, (const uint8_t *)b, in_pitch, width * 2, height);
    ROI roi = {width, height};
    FilterTemporal16s(x.as<PIXEL>(), pp, y.as<PIXEL>(), pp, l.as<PIXEL>(), pp, h.as<PIXEL>(), pp, roi);
    copy_in((uint8_t *)low, out_pitch, l.as<uint8_t>(), pp, width * 2, height);
    copy_in((uint8_t *)high, out_pitch, h.as<uint8_t>(), pp, width * 2, height);
}

void ref_temporal_inv(const int16_t *low, const int16_t *high, int in_pitch, int width, int height, int precision,
                      int16_t *a, int16_t *b, int out_pitch)
{
    const int pp = align16(width * 2) + 128;        // the software-pipelined loop loads 3 vectors ahead
    Aligned x((size_t)pp * (height + 1) + 64), y((size_t)pp * (height + 1) + 64), l((size_t)pp * (height + 1) + 64), h((size_t)pp * (height + 1) + 64);
    copy_in(l.as<uint8_t>(), pp, (const uint8_t *)low, in_pitch, width * 2, height);
    copy_in(h.as<uint8_t>(), pp, (const uint8_t *)high, in_pitch, width * 2, height);
    ROI roi = {width, height};
    InvertTemporalQuant16s(l.as<PIXEL>(), 1, pp, h.as<PIXEL>(), 1, pp, x.as<PIXEL>(), pp, y.as<PIXEL>(), pp, roi, NULL, 0, precision);
    copy_in((uint8_t *)a, out_pitch, x.as<uint8_t>(), pp, width * 2, height);
    copy_in((uint8_t *)b, out_pitch, y.as<uint8_t>(), pp, width * 2, height);
}

// Example/qbist.cpp:252 RunQBist driven exactly like Example/TestCFHD.cpp:1149-1219:
// GetRand(seed); initBaseTransform(); then one RunQBist() per frame; returns frame number `nframes` (1-based).
void ref_qbist_frames(unsigned seed, int width, int height, int pitch, unsigned pixel_format, int nframes, uint8_t *out)
{
    Aligned buf((size_t)width * height * 8 + 64);
    GetRand(seed);
    initBaseTransform();
    // frame 1 = first RunQBist call, frame k = k-th call (RunQBist mutates its genes at the end of each call)
    for (int i = 0; i < nframes; i++)
        RunQBist(width, height, pitch, (CFHD_PixelFormat)pixel_format, 0, buf.as<unsigned char>());
    memcpy(out, buf.p, (size_t)pitch * height);
}

// Frames 1 .. nframes of the same sequence in one pass (out: nframes consecutive buffers of pitch * height bytes).
void ref_qbist_sequence(unsigned seed, int width, int height, int pitch, unsigned pixel_format, int nframes, uint8_t *out)
{
    Aligned buf((size_t)width * height * 8 + 64);
    GetRand(seed);
    initBaseTransform();
    for (int i = 0; i < nframes; i++) {
        RunQBist(width, height, pitch, (CFHD_PixelFormat)pixel_format, 0, buf.as<unsigned char>());
        memcpy(out + (size_t)i * pitch * height, buf.p, (size_t)pitch * height);
    }
}

// ---------------------------------------------------------------------------------------------
// Whole-frame probe: run the reference's own Codec/encoder.c:1897 EncodeSample on one frame and copy
// out every wavelet band plus the quantisation / prescale tables it used.
//   color_format : COLOR_FORMAT_* (Codec/color.h), e.g. 1 = YUYV, 2 = UYVY, 120 = RG48 ...
//   sampling_444 : 0 -> FRAME_SAMPLING_422, 1 -> FRAME_SAMPLING_444
// Outputs (caller-allocated):
//   dims[c*3+k][3] = {width, height, pitch_bytes};  quant[c*12 + k*4 + b];  prescale[c*3 + k]
//   bands: for c, for k (level 1..3), for b (LL,LH,HL,HH): height*width int16, dense, concatenated.
// Returns the encoded sample size in bytes (0 on failure).

static int g_probe_bayer_format = -1;
// Bayer sources only: phase (BAYER_FORMAT_*, Codec/DemoasicFrames.h:30) used by the next ref_encode_frame_bands call;
// the frame is treated as CFHD_ENCODING_FLAGS_CURVE_APPLIED (encode_curve_preset = 1, linear >> 4).  -1 disables.
void ref_set_bayer_format(int fmt) { g_probe_bayer_format = fmt; }
static int g_probe_bayer_preset = 1;
// 1 (default): the frame already carries its curve (encode_curve_preset = 1, samples >> 4); 0: the encoder builds and
// applies its default encode curve (log base 90, Codec/frame.c:5208-5245) itself.
void ref_set_bayer_curve_preset(int preset) { g_probe_bayer_preset = preset; }
static int g_probe_interlaced = 0;
// Interlaced source (CFHD_ENCODING_FLAGS_YUV_INTERLACED -> parameters.progressive = 0, EncoderSDK/SampleEncoder.cpp:210,
// :293): the next ref_encode_frame_bands calls use the frame (field) transform at level 1 (Codec/encoder.c:2949-2993).
void ref_set_interlaced(int on) { g_probe_interlaced = on; }

int ref_encode_frame_bands(const uint8_t *frame, int width, int height, int pitch, int color_format,
                           int sampling_444, int num_channels, int quality,
                           int32_t *dims, int32_t *quant, int32_t *prescale, int16_t *bands, int64_t bands_capacity,
                           uint8_t *sample_out, int64_t sample_capacity)
{
    ENCODER *enc = (ENCODER *)calloc(1, sizeof(ENCODER));
    TRANSFORM *tr[FRAME_MAX_CHANNELS];
    for (int c = 0; c < FRAME_MAX_CHANNELS; c++) { tr[c] = (TRANSFORM *)calloc(1, sizeof(TRANSFORM)); InitTransform(tr[c]); }
    ENCODING_PARAMETERS p;
    memset(&p, 0, sizeof(p));
    p.version = 1; p.gop_length = 1; p.encoded_width = width; p.encoded_height = height;
    p.fixed_quality = quality; p.progressive = g_probe_interlaced ? 0 : 1; p.format = color_format;
    p.frame_sampling = sampling_444 ? FRAME_SAMPLING_444 : FRAME_SAMPLING_422;
    p.colorspace_yuv = 2; p.colorspace_rgb = 1;
    if (!InitializeEncoderWithParameters(NULL, enc, tr, num_channels, &p)) return 0;
    if (g_probe_bayer_format >= 0) { enc->bayer.format = g_probe_bayer_format; enc->encode_curve_preset = g_probe_bayer_preset; }
    size_t scratch_size = 0;
    PIXEL *scratch = CreateEncodingBuffer(NULL, width, height, pitch, color_format, 1, true, &scratch_size);
    const size_t outcap = (size_t)width * height * 16 + 65536;
    Aligned out(outcap), fr((size_t)pitch * (height + 16) + 64);
    memcpy(fr.p, frame, (size_t)pitch * height);
    BITSTREAM bs;
    InitBitstreamBuffer(&bs, out.as<uint8_t>(), outcap, BITSTREAM_ACCESS_WRITE);
    bool ok = EncodeSample(enc, fr.as<uint8_t>(), width, height, pitch, color_format, tr, num_channels, &bs,
                           scratch, scratch_size, quality, 0, NULL, 0.0f, NULL);
    if (!ok) return 0;
    int64_t pos = 0;
    for (int c = 0; c < num_channels; c++) {
        for (int k = 0; k < 3; k++) {
            IMAGE *w = tr[c]->wavelet[k];
            dims[(c * 3 + k) * 3 + 0] = w->width; dims[(c * 3 + k) * 3 + 1] = w->height; dims[(c * 3 + k) * 3 + 2] = w->pitch;
            prescale[c * 3 + k] = tr[c]->prescale[k];
            for (int b = 0; b < 4; b++) {
                quant[c * 12 + k * 4 + b] = w->quant[b];
                if (pos + (int64_t)w->width * w->height > bands_capacity) return 0;
                for (int r = 0; r < w->height; r++)
                    memcpy(bands + pos + (int64_t)r * w->width, (uint8_t *)w->band[b] + (size_t)r * w->pitch, (size_t)w->width * 2);
                pos += (int64_t)w->width * w->height;
            }
        }
    }
    int size = (int)BitstreamSize(&bs);
    if (sample_out && size > 0 && size <= sample_capacity) memcpy(sample_out, out.p, size);
    return size;
}

// ---------------------------------------------------------------------------------------------
// Run-length / VLC stage: the reference's own code tables and its own EncodeQuantLongRuns, so that
// tests can require cfb_sparse_vlc_band's output to be bit-identical.  The tables are read from an encoder the
// reference initialised itself (InitializeEncoderWithParameters -> InitCodebooks, Codec/codebooks.c:202).
extern "C" void EncodeQuantLongRuns(ENCODER *encoder, BITSTREAM *stream, PIXEL *image, int width, int height, int pitch,
                                    int divisor, int active_codebook);
static ENCODER *probe_vlc_encoder(void)
{
    static ENCODER *enc = nullptr;
    if (enc) return enc;
    enc = (ENCODER *)calloc(1, sizeof(ENCODER));
    static TRANSFORM *tr[FRAME_MAX_CHANNELS];
    for (int c = 0; c < FRAME_MAX_CHANNELS; c++) { tr[c] = (TRANSFORM *)calloc(1, sizeof(TRANSFORM)); InitTransform(tr[c]); }
    ENCODING_PARAMETERS p;
    memset(&p, 0, sizeof(p));
    p.version = 1; p.gop_length = 1; p.encoded_width = 256; p.encoded_height = 64;
    p.fixed_quality = 4; p.progressive = 1; p.format = COLOR_FORMAT_YUYV;
    p.frame_sampling = FRAME_SAMPLING_422; p.colorspace_yuv = 2; p.colorspace_rgb = 1;
    if (!InitializeEncoderWithParameters(NULL, enc, tr, 3, &p)) { free(enc); enc = nullptr; }
    return enc;
}

// lengths of the run table and the value table of code set `codebook` (0 on failure)
int ref_vlc_table_lengths(int codebook, int32_t *run_length, int32_t *value_length)
{
    ENCODER *enc = probe_vlc_encoder();
    if (!enc || codebook < 0 || codebook >= CODEC_NUM_CODESETS || !enc->codebook_runbook[codebook] || !enc->valuebook[codebook]) return 0;
    *run_length = enc->codebook_runbook[codebook]->length;
    *value_length = VALUE_TABLE_LENGTH;
    return 1;
}

int ref_vlc_tables(int codebook, uint32_t *run_bits, uint8_t *run_size, uint32_t *run_count, uint32_t *value_bits, uint8_t *value_size)
{
    ENCODER *enc = probe_vlc_encoder();
    if (!enc || codebook < 0 || codebook >= CODEC_NUM_CODESETS) return 0;
    RLCBOOK *rb = enc->codebook_runbook[codebook];
    VALBOOK *vb = enc->valuebook[codebook];
    if (!rb || !vb) return 0;
    const RLC *rlc = (const RLC *)((const char *)rb + sizeof(RLCBOOK));
    for (int i = 0; i < rb->length; i++) { run_bits[i] = rlc[i].bits; run_size[i] = (uint8_t)rlc[i].size; run_count[i] = (uint32_t)rlc[i].count; }
    const VLE *tab = (const VLE *)((const char *)vb + sizeof(VALBOOK));
    for (int i = 0; i < VALUE_TABLE_LENGTH; i++) { value_bits[i] = tab[i].entry & VLE_CODEWORD_MASK; value_size[i] = (uint8_t)(tab[i].entry >> VLE_CODESIZE_SHIFT); }
    return 1;
}

// The reference's coder on one dense band (rows of `width` int16, `pitch` bytes apart).  The bit stream is primed with
// `lead_bits` one-bits so that the band starts in the middle of a word.  Returns the bytes written so far (whole words),
// and the state the coder left in the buffer through buffer_out / bits_free_out.
int64_t ref_vlc_encode_band(const int16_t *band, int width, int height, int pitch, int codebook, int lead_bits,
                            uint8_t *out, int64_t capacity, uint32_t *buffer_out, int32_t *bits_free_out)
{
    ENCODER *enc = probe_vlc_encoder();
    if (!enc) return -1;
    Aligned img((size_t)pitch * height + 64), buf((size_t)capacity + 64);
    memcpy(img.p, band, (size_t)pitch * height);
    BITSTREAM bs;
    InitBitstreamBuffer(&bs, buf.as<uint8_t>(), (size_t)capacity, BITSTREAM_ACCESS_WRITE);
    for (int i = 0; i < lead_bits; i++) PutBits(&bs, 1, 1);
    EncodeQuantLongRuns(enc, &bs, img.as<PIXEL>(), width, height, pitch, 1, codebook);
    *buffer_out = bs.wBuffer; *bits_free_out = bs.nBitsFree;
    const int64_t n = bs.nWordsUsed;
    if (n > capacity) return -1;
    memcpy(out, buf.p, (size_t)n);
    return n;
}

// The same band as it stands in a sample: codes, then the end-of-band code (encoder.c:6503 FinishEncodeBand), padded to a
// 32-bit boundary.  Returns the bytes written.
extern "C" void FinishEncodeBand(BITSTREAM *output, unsigned int code, int size);
int64_t ref_vlc_encode_band_finished(const int16_t *band, int width, int height, int pitch, int codebook, uint8_t *out, int64_t capacity)
{
    ENCODER *enc = probe_vlc_encoder();
    if (!enc) return -1;
    Aligned img((size_t)pitch * height + 64), buf((size_t)capacity + 64);
    memcpy(img.p, band, (size_t)pitch * height);
    BITSTREAM bs;
    InitBitstreamBuffer(&bs, buf.as<uint8_t>(), (size_t)capacity, BITSTREAM_ACCESS_WRITE);
    EncodeQuantLongRuns(enc, &bs, img.as<PIXEL>(), width, height, pitch, 1, codebook);
    FinishEncodeBand(&bs, enc->band_end_code[codebook], enc->band_end_size[codebook]);
    PadBits(&bs);
    FlushBitstream(&bs);
    const int64_t n = bs.nWordsUsed;
    if (n > capacity) return -1;
    memcpy(out, buf.p, (size_t)n);
    return n;
}

int ref_vlc_band_end(int codebook, uint32_t *bits, int32_t *size)
{
    ENCODER *enc = probe_vlc_encoder();
    if (!enc || codebook < 0 || codebook >= CODEC_NUM_CODESETS) return 0;
    *bits = enc->band_end_code[codebook]; *size = enc->band_end_size[codebook];
    return 1;
}

// The reference's FSM band decoder (Codec/decoder.c:19534 DecodeBandFSM16sNoGap) on such a stream, with its tables scaled by
// `quant` as the sample decoder does (decoder.c:20551 DeQuantFSM).  out = pitch * height bytes.  Returns 0 on success.
extern "C" bool DecodeBandFSM16sNoGap(FSM *fsm, BITSTREAM *stream, PIXEL16S *image, int width, int height, int pitch);
extern "C" void DeQuantFSM(FSM *fsm, int quant);
int ref_vlc_decode_band(const uint8_t *stream, int64_t nbytes, int width, int height, int pitch, int codebook, int quant, int16_t *out)
{
    static DECODER *dec = nullptr;
    if (!dec) {
        dec = (DECODER *)calloc(1, DecoderSize());
        if (!DecodeInit(NULL, dec, 256, 64, DECODED_FORMAT_YUYV, DECODED_RESOLUTION_FULL, NULL)) { free(dec); dec = nullptr; return 1; }
    }
    if (codebook < 0 || codebook >= CODEC_NUM_CODESETS) return 2;
    FSM *fsm = &dec->fsm[codebook];
    DeQuantFSM(fsm, quant);
    Aligned smp((size_t)nbytes + 1024), img((size_t)pitch * height + 4096);
    memcpy(smp.p, stream, (size_t)nbytes);
    BITSTREAM bs;
    InitBitstreamBuffer(&bs, smp.as<uint8_t>(), (size_t)nbytes + 512, BITSTREAM_ACCESS_READ);
    if (!DecodeBandFSM16sNoGap(fsm, &bs, img.as<PIXEL16S>(), width, height, pitch)) return 3;
    memcpy(out, img.p, (size_t)pitch * height);
    return 0;
}

// ---------------------------------------------------------------------------------------------
// Two-frame GOP (CFHD_ENCODING_FLAGS_YUV_2FRAME_GOP -> parameters.gop_length = 2, EncoderSDK/SampleEncoder.cpp:211):
// run the reference's EncodeSample on frame A then frame B and copy out all six wavelets of every channel of the
// FIELDPLUS transform (Codec/encoder.c:8431 FinishFieldPlusTransformQuant): wavelet 0/1 = level 1 of frame A/B,
// 2 = temporal (bands 0 = low, 1 = high), 3 = spatial of the temporal highpass, 4 = spatial of the temporal lowpass,
// 5 = spatial of wavelet 4's lowpass.  dims[(c*6+k)*4] = {width, height, pitch, num_bands}; quant[(c*6+k)*4+b];
// prescale[c*8+k]; bands dense, concatenated in (c, k, b) order.  Returns the number of int16 written (0 on failure).
int64_t ref_encode_gop2_bands(const uint8_t *frame_a, const uint8_t *frame_b, int width, int height, int pitch,
                              int color_format, int num_channels, int quality,
                              int32_t *dims, int32_t *quant, int32_t *prescale, int16_t *bands, int64_t bands_capacity)
{
    ENCODER *enc = (ENCODER *)calloc(1, sizeof(ENCODER));
    TRANSFORM *tr[FRAME_MAX_CHANNELS];
    for (int c = 0; c < FRAME_MAX_CHANNELS; c++) { tr[c] = (TRANSFORM *)calloc(1, sizeof(TRANSFORM)); InitTransform(tr[c]); }
    ENCODING_PARAMETERS p;
    memset(&p, 0, sizeof(p));
    p.version = 1; p.gop_length = 2; p.encoded_width = width; p.encoded_height = height;
    p.fixed_quality = quality; p.progressive = g_probe_interlaced ? 0 : 1; p.format = color_format;
    p.frame_sampling = FRAME_SAMPLING_422;
    p.colorspace_yuv = 2; p.colorspace_rgb = 1;
    if (!InitializeEncoderWithParameters(NULL, enc, tr, num_channels, &p)) return 0;
    size_t scratch_size = 0;
    PIXEL *scratch = CreateEncodingBuffer(NULL, width, height, pitch, color_format, 2, true, &scratch_size);
    const size_t outcap = (size_t)width * height * 32 + 65536;
    Aligned out(outcap), fr((size_t)pitch * (height + 16) + 64);
    BITSTREAM bs;
    InitBitstreamBuffer(&bs, out.as<uint8_t>(), outcap, BITSTREAM_ACCESS_WRITE);
    const uint8_t *src[2] = {frame_a, frame_b};
    for (int i = 0; i < 2; i++) {
        memcpy(fr.p, src[i], (size_t)pitch * height);
        if (!EncodeSample(enc, fr.as<uint8_t>(), width, height, pitch, color_format, tr, num_channels, &bs,
                          scratch, scratch_size, quality, 0, NULL, 0.0f, NULL)) return 0;
    }
    int64_t pos = 0;
    for (int c = 0; c < num_channels; c++) {
        for (int k = 0; k < 8; k++) prescale[c * 8 + k] = tr[c]->prescale[k];
        for (int k = 0; k < 6; k++) {
            IMAGE *w = tr[c]->wavelet[k];
            int32_t *d = dims + (c * 6 + k) * 4;
            if (!w) { d[0] = d[1] = d[2] = d[3] = 0; continue; }
            d[0] = w->width; d[1] = w->height; d[2] = w->pitch; d[3] = w->num_bands;
            for (int b = 0; b < w->num_bands && b < 4; b++) {
                quant[(c * 6 + k) * 4 + b] = w->quant[b];
                if (pos + (int64_t)w->width * w->height > bands_capacity) return 0;
                for (int r = 0; r < w->height; r++)
                    memcpy(bands + pos + (int64_t)r * w->width, (uint8_t *)w->band[b] + (size_t)r * w->pitch, (size_t)w->width * 2);
                pos += (int64_t)w->width * w->height;
            }
        }
    }
    return pos;
}

// ---------------------------------------------------------------------------------------------
// Decode a sample with the reference's public API (DecoderSDK/CFHDDecoder.cpp:716 CFHD_DecodeSample)
// at full resolution into `pixel_format` (FOURCC). Returns 0 on success, else the CFHD_Error.
int ref_decode_sample(const uint8_t *sample, int64_t size, int width, int height, unsigned pixel_format,
                      uint8_t *out, int out_pitch)
{
    CFHD_DecoderRef dec = NULL;
    CFHD_Error err = CFHD_OpenDecoder(&dec, NULL);
    if (err) return (int)err;
    Aligned smp((size_t)size + 64), o((size_t)out_pitch * height + 64);
    memcpy(smp.p, sample, (size_t)size);
    int aw = 0, ah = 0;
    CFHD_PixelFormat af = (CFHD_PixelFormat)0;
    err = CFHD_PrepareToDecode(dec, width, height, (CFHD_PixelFormat)pixel_format, CFHD_DECODED_RESOLUTION_FULL,
                               CFHD_DECODING_FLAGS_NONE, smp.p, (size_t)size, &aw, &ah, &af);
    if (!err) err = CFHD_DecodeSample(dec, smp.p, (size_t)size, o.p, out_pitch);
    if (!err) memcpy(out, o.p, (size_t)out_pitch * height);
    CFHD_CloseDecoder(dec);
    return (int)err;
}

// Public-API decode at a reduced resolution (CFHD_DECODED_RESOLUTION_HALF = 2, _QUARTER = 3; SDK
// CFHDTypes.h).  `out` must hold out_pitch * height bytes (full-size is always enough); the actual decoded
// dimensions are returned in dims[0..1].
int ref_decode_sample_res(const uint8_t *sample, int64_t size, int width, int height, unsigned pixel_format,
                          int resolution, uint8_t *out, int out_pitch, int32_t *dims)
{
    CFHD_DecoderRef dec = NULL;
    CFHD_Error err = CFHD_OpenDecoder(&dec, NULL);
    if (err) return (int)err;
    Aligned smp((size_t)size + 64), o((size_t)out_pitch * height + 64);
    memcpy(smp.p, sample, (size_t)size);
    int aw = 0, ah = 0;
    CFHD_PixelFormat af = (CFHD_PixelFormat)0;
    err = CFHD_PrepareToDecode(dec, 0, 0, (CFHD_PixelFormat)pixel_format, (CFHD_DecodedResolution)resolution,
                               CFHD_DECODING_FLAGS_NONE, smp.p, (size_t)size, &aw, &ah, &af);
    if (!err) err = CFHD_DecodeSample(dec, smp.p, (size_t)size, o.p, out_pitch);
    if (!err) memcpy(out, o.p, (size_t)out_pitch * height);
    dims[0] = aw; dims[1] = ah;
    CFHD_CloseDecoder(dec);
    return (int)err;
}

// ---------------------------------------------------------------------------------------------
// Codec-level decode (Codec/decoder.c:1497 DecodeInit + :10078 DecodeSample) that also copies out the
// DECODER's wavelet bands as they stand after the decode: highpass bands hold the DEQUANTISED
// coefficients produced by the FSM entropy decoder (decoder.c:20551), band[0] of level 3 the raw LL3
// and band[0] of levels 2,1 the reconstructed lowpass images.  Same output conventions as
// ref_encode_frame_bands.  decoded_format: DECODED_FORMAT_* (== COLOR_FORMAT_*).  Returns 0 on success.
static int g_decode_resolution = DECODED_RESOLUTION_FULL;
void ref_set_decode_resolution(int resolution) { g_decode_resolution = resolution ? resolution : DECODED_RESOLUTION_FULL; }
int ref_decode_sample_bands(const uint8_t *sample, int64_t size, int width, int height, int decoded_format,
                            int num_channels, uint8_t *out, int out_pitch,
                            int32_t *dims, int32_t *quant, int16_t *bands, int64_t bands_capacity)
{
    DECODER *dec = (DECODER *)calloc(1, DecoderSize());
    if (!DecodeInit(NULL, dec, width, height, decoded_format, g_decode_resolution, NULL)) return 1;
    SetDecoderColorFlags(dec, COLOR_SPACE_CG_709);
    SetDecoderFlags(dec, DECODER_FLAGS_RENDER);      // as CSampleDecoder::DecodeSample does (SampleDecoder.cpp:1507)
    Aligned smp((size_t)size + 64), o((size_t)out_pitch * (height + 16) + 64);
    memcpy(smp.p, sample, (size_t)size);
    BITSTREAM bs;
    InitBitstreamBuffer(&bs, smp.as<uint8_t>(), (size_t)size, BITSTREAM_ACCESS_READ);
    if (!DecodeSample(dec, &bs, o.as<uint8_t>(), out_pitch, NULL, NULL)) return 2;
    memcpy(out, o.p, (size_t)out_pitch * height);
    int64_t pos = 0;
    for (int c = 0; c < num_channels; c++) {
        for (int k = 0; k < 3; k++) {
            IMAGE *w = dec->transform[c]->wavelet[k];
            if (!w) { dims[(c * 3 + k) * 3 + 0] = dims[(c * 3 + k) * 3 + 1] = 0; continue; }   // reduced-resolution decode
            dims[(c * 3 + k) * 3 + 0] = w->width; dims[(c * 3 + k) * 3 + 1] = w->height; dims[(c * 3 + k) * 3 + 2] = w->pitch;
            for (int b = 0; b < 4; b++) {
                quant[c * 12 + k * 4 + b] = w->quantization[b];
                if (pos + (int64_t)w->width * w->height > bands_capacity) return 4;
                for (int r = 0; r < w->height; r++)
                    memcpy(bands + pos + (int64_t)r * w->width, (uint8_t *)w->band[b] + (size_t)r * w->pitch, (size_t)w->width * 2);
                pos += (int64_t)w->width * w->height;
            }
        }
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------
// CPU baseline: time exactly the reference calls that the CUDA path replaces, on one thread,
// for one packed 4:2:2 frame (call from several threads with separate buffers for a pool run).
//   forward  = Codec/encoder.c:3121 TransformForwardSpatialYUV + :3254 ComputeGroupTransformQuant
//   inverse  = Codec/decoder.c:11756/:11765 ReconstructWaveletBand (levels 3->2, 2->1, every channel)
//              + Codec/decoder.c:11836 ReconstructSampleFrameToBuffer (level 1 -> 8-bit YUYV)
// The entropy coder/decoder run once outside the timed loops (to build a valid DECODER state).
// Returns 0 on success; *fwd_seconds / *inv_seconds = total time of `iters` iterations.
#include <time.h>
static double now_s(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }
extern "C" void ReconstructWaveletBand(DECODER *decoder, TRANSFORM *transform, int channel, IMAGE *wavelet, int index,
                                       int precision, const SCRATCH *scratch, int allocations_only);
extern "C" void ReconstructSampleFrameToBuffer(DECODER *decoder, int frame, uint8_t *output, int pitch);

int ref_time_transform_422(const uint8_t *frame, int width, int height, int pitch, int quality, int iters, int cpu_limit,
                           double *fwd_seconds, double *inv_seconds, uint8_t *decoded_out)
{
    const int color_format = COLOR_FORMAT_YUYV, num_channels = 3;
    ENCODER *enc = (ENCODER *)calloc(1, sizeof(ENCODER));
    TRANSFORM *tr[FRAME_MAX_CHANNELS];
    for (int c = 0; c < FRAME_MAX_CHANNELS; c++) { tr[c] = (TRANSFORM *)calloc(1, sizeof(TRANSFORM)); InitTransform(tr[c]); }
    ENCODING_PARAMETERS p;
    memset(&p, 0, sizeof(p));
    p.version = 1; p.gop_length = 1; p.encoded_width = width; p.encoded_height = height;
    p.fixed_quality = quality; p.progressive = g_probe_interlaced ? 0 : 1; p.format = color_format;
    p.frame_sampling = FRAME_SAMPLING_422; p.colorspace_yuv = 2; p.colorspace_rgb = 1;
    if (!InitializeEncoderWithParameters(NULL, enc, tr, num_channels, &p)) return 1;
    size_t scratch_size = 0;
    PIXEL *scratch = CreateEncodingBuffer(NULL, width, height, pitch, color_format, 1, true, &scratch_size);
    const size_t outcap = (size_t)width * height * 8 + 65536;
    Aligned out(outcap), fr((size_t)pitch * (height + 16) + 64), dec_out((size_t)pitch * (height + 16) + 64);
    memcpy(fr.p, frame, (size_t)pitch * height);
    BITSTREAM bs;
    InitBitstreamBuffer(&bs, out.as<uint8_t>(), outcap, BITSTREAM_ACCESS_WRITE);
    if (!EncodeSample(enc, fr.as<uint8_t>(), width, height, pitch, color_format, tr, num_channels, &bs,
                      scratch, scratch_size, quality, 0, NULL, 0.0f, NULL)) return 2;
    const int sample_size = (int)BitstreamSize(&bs);
    FRAME_INFO info;
    InitFrameInfo(&info, width, height, color_format);
    double t0 = now_s();
    for (int i = 0; i < iters; i++) {
        TransformForwardSpatialYUV(fr.as<uint8_t>(), pitch, &info, tr, 0, num_channels, scratch, scratch_size,
                                   enc->codec.chroma_offset, 0, enc->codec.precision, 0, 0);
        ComputeGroupTransformQuant(enc, tr, num_channels);
    }
    *fwd_seconds = now_s() - t0;

    DECODER *dec = (DECODER *)calloc(1, DecoderSize());
    if (!DecodeInit(NULL, dec, width, height, DECODED_FORMAT_YUYV, DECODED_RESOLUTION_FULL, NULL)) return 3;
    SetDecoderColorFlags(dec, COLOR_SPACE_CG_709);
    SetDecoderFlags(dec, DECODER_FLAGS_RENDER);
    if (cpu_limit > 0) { dec->cfhddata.cpu_limit = cpu_limit; }
    BITSTREAM in;
    InitBitstreamBuffer(&in, out.as<uint8_t>(), (size_t)sample_size, BITSTREAM_ACCESS_READ);
    if (!DecodeSample(dec, &in, dec_out.as<uint8_t>(), pitch, NULL, NULL)) return 4;
    const int precision = dec->codec.precision;
    t0 = now_s();
    for (int i = 0; i < iters; i++) {
        for (int c = 0; c < num_channels; c++) {
            TRANSFORM *t = dec->transform[c];
            ReconstructWaveletBand(dec, t, c, t->wavelet[2], 2, precision, &dec->scratch, 0);
            ReconstructWaveletBand(dec, t, c, t->wavelet[1], 1, precision, &dec->scratch, 0);
        }
        ReconstructSampleFrameToBuffer(dec, 0, dec_out.as<uint8_t>(), pitch);
    }
    *inv_seconds = now_s() - t0;
    if (decoded_out) memcpy(decoded_out, dec_out.p, (size_t)pitch * height);
    return 0;
}

// CPU baseline for the planar sources (BASELINE configs 4 and 5: RG48 -> RGB 4:4:4, BYR4 -> four Bayer-derived planes):
// time the reference's own calls for the forward path of one frame on one thread --
//   Codec/encoder.c:2768 ConvertRGB48ToFrame16s (or :2638 ConvertBYR4ToFrame16s), then per channel
//   Codec/encoder.c:3193 TransformForwardSpatial (level 1) and :3254 ComputeGroupTransformQuant (levels 2, 3).
// width / height / pitch as the encoder takes them (BYR4: plane dimensions and the doubled pitch, SampleEncoder.cpp:494).
// One real EncodeSample runs first (outside the timed loop) to build the ENCODER state.  Returns 0 on success.
int ref_time_forward_planar(const uint8_t *frame, int width, int height, int pitch, int color_format, int num_channels,
                            int quality, int iters, double *fwd_seconds)
{
    ENCODER *enc = (ENCODER *)calloc(1, sizeof(ENCODER));
    TRANSFORM *tr[FRAME_MAX_CHANNELS];
    for (int c = 0; c < FRAME_MAX_CHANNELS; c++) { tr[c] = (TRANSFORM *)calloc(1, sizeof(TRANSFORM)); InitTransform(tr[c]); }
    ENCODING_PARAMETERS p;
    memset(&p, 0, sizeof(p));
    p.version = 1; p.gop_length = 1; p.encoded_width = width; p.encoded_height = height;
    p.fixed_quality = quality; p.progressive = 1; p.format = color_format;
    p.frame_sampling = FRAME_SAMPLING_444; p.colorspace_yuv = 2; p.colorspace_rgb = 1;
    if (!InitializeEncoderWithParameters(NULL, enc, tr, num_channels, &p)) return 1;
    if (g_probe_bayer_format >= 0) { enc->bayer.format = g_probe_bayer_format; enc->encode_curve_preset = g_probe_bayer_preset; }
    size_t scratch_size = 0;
    PIXEL *scratch = CreateEncodingBuffer(NULL, width, height, pitch, color_format, 1, true, &scratch_size);
    const size_t outcap = (size_t)width * height * 16 + 65536;
    Aligned out(outcap), fr((size_t)pitch * (height + 16) + 64);
    memcpy(fr.p, frame, (size_t)pitch * height);
    BITSTREAM bs;
    InitBitstreamBuffer(&bs, out.as<uint8_t>(), outcap, BITSTREAM_ACCESS_WRITE);
    if (!EncodeSample(enc, fr.as<uint8_t>(), width, height, pitch, color_format, tr, num_channels, &bs,
                      scratch, scratch_size, quality, 0, NULL, 0.0f, NULL)) return 2;
    FRAME *f = enc->frame;
    if (!f) return 3;
    const double t0 = now_s();
    for (int i = 0; i < iters; i++) {
        if (color_format == COLOR_FORMAT_BYR4)
            ConvertBYR4ToFrame16s(enc->bayer.format, enc->encode_curve, enc->encode_curve_preset, fr.as<uint8_t>(), pitch, f, enc->codec.precision);
        else
            ConvertRGB48ToFrame16s(fr.as<uint8_t>(), pitch, f, (uint8_t *)scratch, enc->codec.precision, color_format);
        for (int c = 0; c < num_channels; c++) {
            IMAGE *wavelet = tr[c]->wavelet[0];
            TransformForwardSpatial(NULL, f->channel[c], 0, wavelet, 1, scratch, scratch_size, 0, wavelet->quant, 0);
        }
        ComputeGroupTransformQuant(enc, tr, num_channels);
    }
    *fwd_seconds = now_s() - t0;
    return 0;
}

}  // extern "C"
