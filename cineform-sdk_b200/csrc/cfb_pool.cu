// cfb_pool.cu -- asynchronous, in-order, multi-GPU frame pool (see include/cfhd_b200.h).
//
// GPU re-hosting of the reference's CEncoderPool / EncoderJobQueue (EncoderSDK/EncoderPool.cpp:239,
// EncoderQueue.h:311-352): same contract (bounded queue, round-robin assignment, strict in-order delivery,
// borrowed buffers), but a "worker" is a (GPU, stream, staging) slot instead of a CPU thread running the
// SSE2 transform.  Frames are independent, so GPUs never exchange data (no NCCL).
#include <condition_variable>
#include <deque>
#include <memory>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

#include "cfb_host.h"

using namespace cfb;

namespace {

struct Job {
    uint32_t frame_number = 0;
    bool inverse = false;
    bool sparse = false;
    const void *src = nullptr;
    void *dst = nullptr;
    int pitch = 0;
    int out_format = 0;
    cfb_quant quant{};
    int device_index = 0;
    bool taken = false, done = false;
    cfb_error error = CFB_OK;
};

struct Slot {
    cfb_context *ctx = nullptr;
    cfb_codec *codec = nullptr;
    std::thread thread;
};

}  // namespace

struct cfb_pool {
    cfb_frame_desc desc{};
    cfb_layout layout{};
    int batch = 1, queue_length = 1;
    std::vector<int> devices;
    std::vector<std::unique_ptr<Slot>> slots;       // slots_per_device * ndevices
    int slots_per_device = 1;

    std::mutex mu;
    std::condition_variable cv_work, cv_done, cv_space;
    std::deque<std::shared_ptr<Job>> queue;         // submission order; front = oldest undelivered
    uint64_t submitted = 0;
    bool stopping = false;

    void worker(int slot_index, int device_index);
};

void cfb_pool::worker(int slot_index, int device_index)
{
    Slot &s = *slots[slot_index];
    cudaSetDevice(s.ctx->device);
    cfb_bind_thread_to_device(s.ctx->device);       // copies are issued from the GPU's own NUMA node
    std::vector<std::shared_ptr<Job>> mine;
    for (;;) {
        mine.clear();
        {
            std::unique_lock<std::mutex> lk(mu);
            for (;;) {
                // take up to `batch` untaken jobs of this device, all of the same direction and quant table
                for (auto &j : queue) {
                    if (j->taken || j->device_index != device_index) continue;
                    if (!mine.empty() && (j->inverse != mine[0]->inverse || j->sparse != mine[0]->sparse || j->out_format != mine[0]->out_format ||
                                          j->pitch != mine[0]->pitch ||
                                          memcmp(&j->quant, &mine[0]->quant, sizeof(cfb_quant)) != 0))
                        break;
                    j->taken = true;
                    mine.push_back(j);
                    if ((int)mine.size() == batch) break;
                }
                if (!mine.empty() || stopping) break;
                cv_work.wait(lk);
            }
            if (mine.empty() && stopping) return;
        }
        const int n = (int)mine.size();
        const void *src[kMaxBatch];
        void *dst[kMaxBatch];
        for (int i = 0; i < n; i++) { src[i] = mine[i]->src; dst[i] = mine[i]->dst; }
        cfb_error e;
        if (!mine[0]->inverse)
            e = mine[0]->sparse ? cfb_forward_host_sparse(s.codec, n, src, mine[0]->pitch, &mine[0]->quant, dst, nullptr)
                                : cfb_forward_host(s.codec, n, src, mine[0]->pitch, &mine[0]->quant, dst);
        else
            e = mine[0]->sparse ? cfb_inverse_host_sparse(s.codec, n, src, &mine[0]->quant, mine[0]->out_format, dst, mine[0]->pitch)
                                : cfb_inverse_host(s.codec, n, src, &mine[0]->quant, mine[0]->out_format, dst, mine[0]->pitch);
        {
            std::lock_guard<std::mutex> lk(mu);
            for (auto &j : mine) { j->error = e; j->done = true; }
        }
        cv_done.notify_all();
    }
}

extern "C" {

cfb_error cfb_host_alloc(size_t bytes, void **out)
{
    if (!out) return CFB_ERROR_INVALID_ARGUMENT;
    *out = nullptr;
    CFB_CUDA(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocPortable));
    return CFB_OK;
}

void cfb_host_free(void *p) { if (p) cudaFreeHost(p); }

cfb_error cfb_pool_create(const int *devices, int ndevices, const cfb_frame_desc *desc,
                          int slots, int batch, int queue_length, cfb_pool **out)
{
    if (!devices || !desc || !out || ndevices < 1) { set_error("null/empty argument"); return CFB_ERROR_INVALID_ARGUMENT; }
    *out = nullptr;
    if (slots < 1 || slots > 32 || batch < 1 || batch > CFB_MAX_BATCH || queue_length < 1) {
        set_error("slots %d (1..16), batch %d (1..%d), queue_length %d (>=1) out of range", slots, batch, CFB_MAX_BATCH, queue_length);
        return CFB_ERROR_INVALID_ARGUMENT;
    }
    cfb_layout lay;
    cfb_error err = cfb_layout_compute(desc, &lay);
    if (err) return err;
    std::unique_ptr<cfb_pool> p(new (std::nothrow) cfb_pool());
    if (!p) return CFB_ERROR_OUTOFMEMORY;
    p->desc = *desc; p->layout = lay; p->batch = batch; p->queue_length = queue_length;
    p->devices.assign(devices, devices + ndevices);
    p->slots_per_device = slots;
    for (int d = 0; d < ndevices && !err; d++)
        for (int k = 0; k < slots && !err; k++) {
            std::unique_ptr<Slot> s(new Slot());
            err = cfb_context_create(devices[d], &s->ctx);
            if (!err) err = cfb_codec_create(s->ctx, desc, batch, &s->codec);
            if (err) { if (s->codec) cfb_codec_destroy(s->codec); if (s->ctx) cfb_context_destroy(s->ctx); break; }
            p->slots.push_back(std::move(s));
        }
    if (err) {
        for (auto &s : p->slots) { cfb_codec_destroy(s->codec); cfb_context_destroy(s->ctx); }
        return err;
    }
    cfb_pool *raw = p.release();
    for (int i = 0; i < (int)raw->slots.size(); i++)
        raw->slots[i]->thread = std::thread(&cfb_pool::worker, raw, i, i / slots);
    *out = raw;
    return CFB_OK;
}

cfb_error cfb_pool_set_interlaced(cfb_pool *pool, int interlaced)
{
    if (!pool) { set_error("null pool"); return CFB_ERROR_INVALID_ARGUMENT; }
    std::lock_guard<std::mutex> lk(pool->mu);          // applies to jobs submitted after this call returns
    for (auto &s : pool->slots) {
        cfb_error e = cfb_codec_set_interlaced(s->codec, interlaced);
        if (e) return e;
    }
    return CFB_OK;
}

cfb_error cfb_pool_set_decode_resolution(cfb_pool *pool, int resolution)
{
    if (!pool) { set_error("null pool"); return CFB_ERROR_INVALID_ARGUMENT; }
    std::lock_guard<std::mutex> lk(pool->mu);          // applies to jobs submitted after this call returns
    for (auto &s : pool->slots) {
        cfb_error e = cfb_codec_set_decode_resolution(s->codec, resolution);
        if (e) return e;
    }
    return CFB_OK;
}

void cfb_pool_destroy(cfb_pool *pool)
{
    if (!pool) return;
    {
        std::lock_guard<std::mutex> lk(pool->mu);
        pool->stopping = true;
    }
    pool->cv_work.notify_all();
    for (auto &s : pool->slots) if (s->thread.joinable()) s->thread.join();
    for (auto &s : pool->slots) { cfb_codec_destroy(s->codec); cfb_context_destroy(s->ctx); }
    delete pool;
}

static cfb_error pool_submit(cfb_pool *pool, std::shared_ptr<Job> job)
{
    {
        std::unique_lock<std::mutex> lk(pool->mu);
        pool->cv_space.wait(lk, [&] { return (int)pool->queue.size() < pool->queue_length; });
        job->device_index = (int)(pool->submitted % pool->devices.size());     // EncoderPool.cpp:284
        pool->submitted++;
        pool->queue.push_back(std::move(job));
    }
    pool->cv_work.notify_all();
    return CFB_OK;
}

cfb_error cfb_pool_submit_forward(cfb_pool *pool, uint32_t frame_number, const void *h_frame, int frame_pitch,
                                  const cfb_quant *quant, void *h_coded)
{
    if (!pool || !h_frame || !quant || !h_coded) { set_error("null argument"); return CFB_ERROR_INVALID_ARGUMENT; }
    auto j = std::make_shared<Job>();
    j->frame_number = frame_number; j->inverse = false; j->src = h_frame; j->dst = h_coded; j->pitch = frame_pitch; j->quant = *quant;
    return pool_submit(pool, std::move(j));
}

cfb_error cfb_pool_submit_inverse(cfb_pool *pool, uint32_t frame_number, const void *h_coded,
                                  const cfb_quant *quant, int out_format, void *h_frame, int frame_pitch)
{
    if (!pool || !h_frame || !quant || !h_coded) { set_error("null argument"); return CFB_ERROR_INVALID_ARGUMENT; }
    auto j = std::make_shared<Job>();
    j->frame_number = frame_number; j->inverse = true; j->src = h_coded; j->dst = h_frame; j->pitch = frame_pitch;
    j->out_format = out_format; j->quant = *quant;
    return pool_submit(pool, std::move(j));
}

cfb_error cfb_pool_submit_forward_sparse(cfb_pool *pool, uint32_t frame_number, const void *h_frame, int frame_pitch,
                                         const cfb_quant *quant, void *h_sparse)
{
    if (!pool || !h_frame || !quant || !h_sparse) { set_error("null argument"); return CFB_ERROR_INVALID_ARGUMENT; }
    auto j = std::make_shared<Job>();
    j->frame_number = frame_number; j->inverse = false; j->sparse = true; j->src = h_frame; j->dst = h_sparse; j->pitch = frame_pitch; j->quant = *quant;
    return pool_submit(pool, std::move(j));
}

cfb_error cfb_pool_submit_inverse_sparse(cfb_pool *pool, uint32_t frame_number, const void *h_sparse,
                                         const cfb_quant *quant, int out_format, void *h_frame, int frame_pitch)
{
    if (!pool || !h_frame || !quant || !h_sparse) { set_error("null argument"); return CFB_ERROR_INVALID_ARGUMENT; }
    auto j = std::make_shared<Job>();
    j->frame_number = frame_number; j->inverse = true; j->sparse = true; j->src = h_sparse; j->dst = h_frame; j->pitch = frame_pitch;
    j->out_format = out_format; j->quant = *quant;
    return pool_submit(pool, std::move(j));
}

static cfb_error pool_pop(cfb_pool *pool, bool block, uint32_t *frame_number, cfb_error *job_error)
{
    if (!pool) return CFB_ERROR_INVALID_ARGUMENT;
    std::shared_ptr<Job> j;
    {
        std::unique_lock<std::mutex> lk(pool->mu);
        if (pool->queue.empty()) { set_error("no job outstanding"); return CFB_ERROR_INVALID_ARGUMENT; }
        if (!pool->queue.front()->done) {
            if (!block) return CFB_ERROR_NOT_FINISHED;
            pool->cv_done.wait(lk, [&] { return pool->queue.front()->done; });
        }
        j = pool->queue.front();
        pool->queue.pop_front();
    }
    pool->cv_space.notify_all();
    if (frame_number) *frame_number = j->frame_number;
    if (job_error) *job_error = j->error;
    return CFB_OK;
}

cfb_error cfb_pool_wait(cfb_pool *pool, uint32_t *frame_number, cfb_error *job_error) { return pool_pop(pool, true, frame_number, job_error); }
cfb_error cfb_pool_test(cfb_pool *pool, uint32_t *frame_number, cfb_error *job_error) { return pool_pop(pool, false, frame_number, job_error); }

cfb_error cfb_pool_stats(cfb_pool *pool, cfb_stats *out)
{
    if (!pool || !out) return CFB_ERROR_INVALID_ARGUMENT;
    memset(out, 0, sizeof(*out));
    for (auto &s : pool->slots) {
        cfb_stats t;
        cfb_context_stats(s->ctx, &t);
        out->kernel_launches += t.kernel_launches; out->frames_forward += t.frames_forward; out->frames_inverse += t.frames_inverse;
        out->h2d_bytes += t.h2d_bytes; out->d2h_bytes += t.d2h_bytes;
    }
    return CFB_OK;
}

}  // extern "C"
