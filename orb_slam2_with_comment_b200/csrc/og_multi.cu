// og_multi.cu — single-process multi-GPU extraction (north_star: "batches of independent frames are partitioned
// across the GPUs of one box with plain per-device streams and host gather; no NCCL, no cross-frame reduction exists").
//
// One persistent host thread per device owns an orbgpu_extractor there.  A call cuts the batch into contiguous frame ranges
// [g*B/G, (g+1)*B/G) — output order = input order without a permutation — and every worker runs the chunked
// H2D -> kernels -> D2H pipeline of orbgpu_extract_batch on its range, reading the caller's images and writing the caller's
// key-point / descriptor / count arrays at the range's offsets: the "gather" is the D2H copies landing in one caller buffer.
// Results are byte-identical to a one-device run (each frame is processed independently by the same kernels).
#include "og_nvtx.h"
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <condition_variable>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/orbgpu.h"

int og_fail(int code, const std::string& msg);   // og_capi.cu (thread-local last error)

namespace {

struct Job {
    const uint8_t* images = nullptr;
    int width = 0, height = 0;
    size_t row_stride = 0, frame_stride = 0;
    orbgpu_keypoint* kp = nullptr;
    uint8_t* desc = nullptr;
    int kp_capacity = 0;
    int32_t* counts = nullptr;
    int f0 = 0, f1 = 0;
};

struct Worker {
    int device = 0;
    orbgpu_extractor* ex = nullptr;
    std::thread th;
    std::mutex mu;
    std::condition_variable cv;
    bool has_job = false, done = false, quit = false;
    Job job;
    int rc = 0;
    std::string err;
    int max_batch = 0;
    int launches = 0;
};

void run_job(Worker* w) {
    const Job& j = w->job;
    w->rc = ORBGPU_OK;
    w->err.clear();
    w->launches = 0;
    for (int f = j.f0; f < j.f1 && w->rc == ORBGPU_OK; f += w->max_batch) {
        const int nb = std::min(w->max_batch, j.f1 - f);
        w->rc = orbgpu_extract_batch(w->ex, j.images + (size_t)f * j.frame_stride, nb, j.width, j.height, j.row_stride, j.frame_stride,
                                     j.kp + (size_t)f * j.kp_capacity, j.desc + (size_t)f * j.kp_capacity * 32, j.kp_capacity, j.counts + f);
        if (w->rc != ORBGPU_OK) w->err = orbgpu_last_error();
        w->launches += orbgpu_extractor_last_launches(w->ex);
    }
}

void worker_main(Worker* w) {
    cudaSetDevice(w->device);
    std::unique_lock<std::mutex> lk(w->mu);
    for (;;) {
        w->cv.wait(lk, [w] { return w->has_job || w->quit; });
        if (w->quit) return;
        w->has_job = false;
        lk.unlock();
        run_job(w);
        lk.lock();
        w->done = true;
        w->cv.notify_all();
    }
}

}  // namespace

struct orbgpu_multi_extractor {
    std::vector<Worker*> workers;
    int kp_cap = 0;
    int last_launches = 0;
};

extern "C" {

int orbgpu_multi_extractor_destroy(orbgpu_multi_extractor* me) {
    if (!me) return ORBGPU_OK;
    for (Worker* w : me->workers) {
        if (w->th.joinable()) {
            {
                std::lock_guard<std::mutex> lk(w->mu);
                w->quit = true;
            }
            w->cv.notify_all();
            w->th.join();
        }
        if (w->ex) orbgpu_extractor_destroy(w->ex);
        delete w;
    }
    delete me;
    return ORBGPU_OK;
}

int orbgpu_multi_extractor_create(orbgpu_multi_extractor** out, const int* devices, int n_devices, int nfeatures, float scale_factor,
                                  int nlevels, int ini_th_fast, int min_th_fast, int max_width, int max_height, int max_batch_per_device) {
    if (!out) return og_fail(ORBGPU_ERR_ARG, "null out");
    *out = nullptr;
    if (n_devices < 1 || max_batch_per_device < 1) return og_fail(ORBGPU_ERR_ARG, "multi extractor: need at least one device and a positive batch");
    for (int a = 0; a < n_devices; ++a)
        for (int b = a + 1; b < n_devices; ++b)
            if (devices && devices[a] == devices[b]) return og_fail(ORBGPU_ERR_ARG, "multi extractor: a device is listed twice");
    orbgpu_multi_extractor* me = new orbgpu_multi_extractor();
    for (int g = 0; g < n_devices; ++g) {
        Worker* w = new Worker();
        me->workers.push_back(w);
        w->device = devices ? devices[g] : g;
        w->max_batch = max_batch_per_device;
        const int rc = orbgpu_extractor_create(&w->ex, w->device, nfeatures, scale_factor, nlevels, ini_th_fast, min_th_fast, max_width, max_height,
                                               max_batch_per_device);
        if (rc != ORBGPU_OK) {
            const std::string msg = std::string("multi extractor, device ") + std::to_string(w->device) + ": " + orbgpu_last_error();
            orbgpu_multi_extractor_destroy(me);
            return og_fail(rc, msg);
        }
    }
    me->kp_cap = orbgpu_extractor_max_keypoints(me->workers[0]->ex);
    for (Worker* w : me->workers) w->th = std::thread(worker_main, w);
    *out = me;
    return ORBGPU_OK;
}

int orbgpu_multi_extractor_device_count(const orbgpu_multi_extractor* me) { return me ? (int)me->workers.size() : 0; }
int orbgpu_multi_extractor_max_keypoints(const orbgpu_multi_extractor* me) { return me ? me->kp_cap : 0; }
int orbgpu_multi_extractor_last_launches(const orbgpu_multi_extractor* me) { return me ? me->last_launches : 0; }

int orbgpu_multi_extractor_frame_range(const orbgpu_multi_extractor* me, int batch, int g, int* first, int* last) {
    OG_NVTX("orbgpu_multi_extractor_frame_range");
    if (!me || g < 0 || g >= (int)me->workers.size() || batch < 0) return og_fail(ORBGPU_ERR_ARG, "frame_range: bad arguments");
    const long long G = (long long)me->workers.size();
    if (first) *first = (int)((long long)batch * g / G);
    if (last) *last = (int)((long long)batch * (g + 1) / G);
    return ORBGPU_OK;
}

int orbgpu_multi_extract_batch(orbgpu_multi_extractor* me, const uint8_t* images, int batch, int width, int height, size_t row_stride,
                               size_t frame_stride, orbgpu_keypoint* kp_out, uint8_t* desc_out, int kp_capacity, int32_t* counts) {
    OG_NVTX("orbgpu_multi_extract_batch");
    if (!me) return og_fail(ORBGPU_ERR_ARG, "null multi extractor");
    if (batch < 0) return og_fail(ORBGPU_ERR_ARG, "negative batch");
    if (batch == 0) return ORBGPU_OK;
    if (!images || !kp_out || !desc_out || !counts) return og_fail(ORBGPU_ERR_ARG, "null pointer");
    if (kp_capacity < me->kp_cap) return og_fail(ORBGPU_ERR_CAPACITY, "kp_capacity below orbgpu_multi_extractor_max_keypoints()");
    const int G = (int)me->workers.size();
    for (int g = 0; g < G; ++g) {
        Worker* w = me->workers[g];
        Job j;
        j.images = images; j.width = width; j.height = height; j.row_stride = row_stride; j.frame_stride = frame_stride;
        j.kp = kp_out; j.desc = desc_out; j.kp_capacity = kp_capacity; j.counts = counts;
        orbgpu_multi_extractor_frame_range(me, batch, g, &j.f0, &j.f1);
        std::lock_guard<std::mutex> lk(w->mu);
        w->job = j;
        w->done = false;
        w->has_job = true;
        w->cv.notify_all();
    }
    int rc = ORBGPU_OK;
    std::string err;
    me->last_launches = 0;
    for (int g = 0; g < G; ++g) {
        Worker* w = me->workers[g];
        std::unique_lock<std::mutex> lk(w->mu);
        w->cv.wait(lk, [w] { return w->done; });
        me->last_launches += w->launches;
        if (w->rc != ORBGPU_OK && rc == ORBGPU_OK) {
            rc = w->rc;
            err = std::string("device ") + std::to_string(w->device) + ": " + w->err;
        }
    }
    return rc == ORBGPU_OK ? ORBGPU_OK : og_fail(rc, err);
}

}  // extern "C"
