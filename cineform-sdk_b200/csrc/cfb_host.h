// cfb_host.h -- host-side internals shared by the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <cstring>
#include <string>

#include "../../include/cfhd_b200.h"
#include "cfb_common.cuh"

namespace cfb {

void set_error(const char *fmt, ...);
cfb_error cuda_fail(cudaError_t e, const char *what);

#define CFB_CUDA(call)                                                  \
    do {                                                                \
        cudaError_t e_ = (call);                                        \
        if (e_ != cudaSuccess) return ::cfb::cuda_fail(e_, #call);      \
    } while (0)

QuantParam make_quant_param(int divisor, int midpoint_prequant, bool plain_midpoint = false);
// wait for everything queued on the context's stream without busy-waiting on a CPU core
cudaError_t stream_wait(cfb_context *ctx);

// kernel launchers (cfb_forward.cu / cfb_inverse.cu)
cudaError_t launch_fwd_plane(const FwdParams &p, int prescale, cudaStream_t stream);
cudaError_t launch_fwd_422(const FwdParams &p, cudaStream_t stream);
cudaError_t launch_fwd_rg48(const FwdParams &p, int sel, cudaStream_t stream);
cudaError_t launch_fwd_byr4(const FwdParams &p, cudaStream_t stream);
cudaError_t launch_fwd_rgb30(const FwdParams &p, cudaStream_t stream);
cudaError_t launch_inv_plane(const InvParams &p, int descale, cudaStream_t stream);
cudaError_t launch_inv_422(const InvParams &p, cudaStream_t stream);
cudaError_t launch_lowpass_422(const InvParams &p, cudaStream_t stream);
cudaError_t launch_inv_fields(const InvParams &p, const FieldsAux &a, bool planar, cudaStream_t stream);
cudaError_t launch_fwd_422_fields(const FwdParams &p, cudaStream_t stream);
cudaError_t launch_fwd_yu64(const FwdParams &p, cudaStream_t stream);
cudaError_t launch_fwd_v210(const FwdParams &p, cudaStream_t stream);

}  // namespace cfb

struct cfb_context {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr;             // blocking-sync event: host threads sleep instead of spinning
    int sm_count = 0;
    std::atomic<uint64_t> kernel_launches{0}, frames_forward{0}, frames_inverse{0}, h2d_bytes{0}, d2h_bytes{0};
};

struct cfb_codec {
    cfb_context *ctx = nullptr;
    cfb_frame_desc desc{};
    cfb_layout layout{};
    int max_batch = 0;
    unsigned char *d_frames = nullptr;      // max_batch packed frames
    unsigned char *d_pyramids = nullptr;    // max_batch pyramids
    size_t frame_stride = 0;                // bytes between device frame slots
    size_t pyramid_stride = 0;
    int bayer_phase = 0;                    // BAYER_FORMAT_* (0 RED_GRN, 1 GRN_RED, 2 GRN_BLU, 3 BLU_GRN), DemoasicFrames.h:30
    int fwd_mask = 7, inv_mask = 7;         // profiling aid: levels to run
    int interlaced = 0;                     // level 1 is the field transform (CFHD_ENCODING_FLAGS_YUV_INTERLACED)
    int *d_carry = nullptr;                 // interlaced inverse: HL row carries, kMaxBatch frames
    unsigned short *d_curve = nullptr;      // Bayer encode curve (1 << 14 entries), null = frame already curved
    unsigned char *d_gop = nullptr;         // two-frame GOP buffer (cfb_gop2_layout.total_bytes), allocated on first use
    int carry_strips = 0;
    int decode_res = 1;                     // CFB_RESOLUTION_*: 1 full, 2 half (LL1), 3 quarter (LL2)
    // sparse transfer format staging (allocated on first use)
    unsigned char *d_sparse = nullptr;      // max_batch sparse buffers
    unsigned *d_counts = nullptr;           // max_batch * (nseg + 1)
    unsigned *h_headers = nullptr;          // pinned, 4 u32 per slot
    size_t sparse_stride = 0;
    unsigned value_guess = 0;               // running estimate of non-zero words per frame (speculative single-pass D2H)
};
