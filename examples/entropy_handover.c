/* entropy_handover.c -- the host side of the sparse transfer format from plain C (gcc only, runs WITHOUT a GPU: none of
 * these entry points touches the device).
 *
 * What an entropy coder / decoder does with the buffers the GPU path exchanges with the host:
 *   encoder:  sparse buffer --cfb_sparse_vlc_band--> run-length / VLC stream of every band   (replaces the walk of
 *             Codec/encoder.c:5386 EncodeQuantLongRuns over dense bands)
 *   decoder:  band streams --cfb_vlc_decode_band + cfb_sparse_writer--> sparse buffer          (replaces the dense bands
 *             Codec/decoder.c:19534 DecodeBandFSM16sNoGap fills)
 * The code tables are the caller's (the codec's live in Codec/codebooks.c); this example uses a small prefix-free set.
 * Round trip: dense coded region -> sparse -> streams -> sparse' ; sparse' must equal sparse byte for byte.
 *
 *   gcc -std=c99 -O2 -I include examples/entropy_handover.c -L cineform-sdk_b200 -lcfhd_b200 -o entropy_handover
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "cfhd_b200.h"

#define CHECK(call)                                                                                   \
    do {                                                                                              \
        cfb_error e_ = (call);                                                                        \
        if (e_ != CFB_OK) { fprintf(stderr, "%s failed: %d (%s)\n", #call, (int)e_, cfb_last_error_string()); return 2; } \
    } while (0)

/* toy code set: runs start with 1, values with 0, end of band = 1111 */
static const uint32_t run_bits[4] = {0, 0x2, 0x6, 0xE}, run_count[4] = {0, 1, 2, 3};
static const uint8_t run_size[4] = {0, 2, 3, 4};
/* value table of 8 entries: index v for v = 1..3, 8 + v for v = -1..-3 */
static const uint32_t value_bits[8] = {0x3F, 0x0, 0x6, 0x1E, 0x3F, 0x3E, 0xE, 0x2};
static const uint8_t value_size[8] = {7, 2, 4, 6, 7, 7, 5, 3};

static const uint32_t dec_bits[10] = {0x2, 0x6, 0xE, 0xF, 0x0, 0x2, 0x6, 0xE, 0x1E, 0x3E};
static const uint8_t dec_size[10] = {2, 3, 4, 4, 2, 3, 4, 5, 6, 7};
static const uint8_t dec_kind[10] = {1, 1, 1, 2, 0, 0, 0, 0, 0, 0};
static const int32_t dec_arg[10] = {1, 2, 3, 0, 1, -1, 2, -2, 3, -3};

int main(int argc, char **argv)
{
    const int width = argc > 1 ? atoi(argv[1]) : 1920, height = argc > 2 ? atoi(argv[2]) : 1080;
    cfb_frame_desc desc = {width, height, CFB_PIXEL_YUYV, 0};
    cfb_layout lay;
    CHECK(cfb_layout_compute(&desc, &lay));

    /* a synthetic coded region: sparse small values in every coded band (pitch gaps stay zero, as the kernels leave them) */
    int16_t *coded = (int16_t *)calloc((size_t)lay.coded_bytes, 1);
    uint32_t rng = 12345u;
    long nonzero = 0;
    for (int c = 0; c < lay.num_channels; c++)
        for (int k = 0; k < CFB_NUM_LEVELS; k++)
            for (int b = (k == CFB_NUM_LEVELS - 1 ? 0 : 1); b < CFB_NUM_BANDS; b++) {
                const cfb_band_layout *bl = &lay.band[c][k][b];
                for (int y = 0; y < bl->height; y++)
                    for (int x = 0; x < bl->width; x++) {
                        rng = rng * 1664525u + 1013904223u;
                        if ((rng >> 24) < (k == 2 ? 80u : 8u)) {
                            const int v = (int)((rng >> 8) % 3u) + 1;
                            coded[(bl->offset + (int64_t)y * bl->pitch) / 2 + x] = (int16_t)((rng & 0x80u) ? -v : v);
                            nonzero++;
                        }
                    }
            }

    const size_t cap = cfb_sparse_max_bytes(&lay);
    void *sparse = malloc(cap), *sparse2 = malloc(cap);
    size_t sparse_bytes = 0, sparse2_bytes = 0;
    CHECK(cfb_sparse_compact(&lay, coded, sparse, &sparse_bytes));

    cfb_vlc_codebook book = {4, 8, run_bits, run_size, run_count, value_bits, value_size};
    cfb_vlc_decodebook dbook = {10, dec_bits, dec_size, dec_kind, dec_arg};
    cfb_vlc_decoder *dec = NULL;
    cfb_sparse_writer *wr = NULL;
    CHECK(cfb_vlc_decoder_create(&dbook, &dec));
    CHECK(cfb_sparse_writer_create(&lay, &wr));
    CHECK(cfb_sparse_writer_begin(wr, sparse2, cap));

    uint8_t *stream = (uint8_t *)malloc((size_t)lay.coded_bytes + 64);
    size_t stream_total = 0;
    for (int c = 0; c < lay.num_channels; c++)
        for (int k = CFB_NUM_LEVELS - 1; k >= 0; k--)            /* coded order: LL3, level 3, level 2, level 1 */
            for (int b = (k == CFB_NUM_LEVELS - 1 ? 0 : 1); b < CFB_NUM_BANDS; b++) {
                /* encoder side */
                cfb_bitwriter bw = {stream, stream + lay.coded_bytes + 64, 0, 32, 0};
                CHECK(cfb_sparse_vlc_band(&lay, sparse, c, k, b, &book, &bw));
                /* finish the band the way the host coder does: end-of-band code, then flush the pending bits padded to a word */
                uint64_t acc = bw.bits_free < 32 ? bw.buffer : 0;
                int n = 32 - bw.bits_free;
                acc = (acc << 4) | 0xFu; n += 4;
                while (n > 0) {
                    const int take = n >= 32 ? 32 : n;
                    const uint32_t word = (uint32_t)((n >= 32 ? acc >> (n - 32) : acc << (32 - n)) & 0xFFFFFFFFu);
                    bw.cur[0] = (uint8_t)(word >> 24); bw.cur[1] = (uint8_t)(word >> 16); bw.cur[2] = (uint8_t)(word >> 8); bw.cur[3] = (uint8_t)word;
                    bw.cur += 4; bw.bytes += 4; n -= take;
                }
                stream_total += (size_t)bw.bytes;
                /* decoder side */
                size_t used = 0;
                CHECK(cfb_vlc_decode_band(dec, wr, c, k, b, stream, (size_t)bw.bytes, 1, &used));
                if (used + 4 < (size_t)bw.bytes || used > (size_t)bw.bytes) { fprintf(stderr, "band (%d,%d,%d): parser consumed %zu of %lld bytes\n", c, k, b, used, (long long)bw.bytes); return 1; }
            }
    CHECK(cfb_sparse_writer_end(wr, &sparse2_bytes));

    const int same = sparse_bytes == sparse2_bytes && memcmp(sparse, sparse2, sparse_bytes) == 0;
    int16_t *back = (int16_t *)malloc((size_t)lay.coded_bytes);
    CHECK(cfb_sparse_expand(&lay, sparse2, back));
    const int dense_same = memcmp(back, coded, (size_t)lay.coded_bytes) == 0;
    printf("{\"width\": %d, \"height\": %d, \"coded_bytes\": %lld, \"nonzero_coefficients\": %ld, \"sparse_bytes\": %zu, "
           "\"stream_bytes\": %zu, \"sparse_round_trip_identical\": %d, \"dense_round_trip_identical\": %d}\n",
           width, height, (long long)lay.coded_bytes, nonzero, sparse_bytes, stream_total, same, dense_same);
    cfb_vlc_decoder_destroy(dec);
    cfb_sparse_writer_destroy(wr);
    free(coded); free(sparse); free(sparse2); free(stream); free(back);
    return (same && dense_same) ? 0 : 1;
}
