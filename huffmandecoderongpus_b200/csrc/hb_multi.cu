/*
 * hb_multi.cu -- the multi-GPU shard/merge driver behind the C ABI: ONE process, N devices.
 *
 * The reference has no multi-GPU path; SURVEY.md 8(b)(3)/8(e) ask for one under the approach
 * table.  A stream is cut into N contiguous byte ranges (16-byte aligned, 16-byte halo); every
 * device computes its shard's 32-entry transfer map (hb_shard_map), stores it straight into the
 * exchange tables of the devices to its right over NVLink, waits for the maps of the devices to
 * its left and composes them -- hb_shard_exchange, ONE kernel, the contexts connected with
 * hb_peer_connect_local: no NCCL, no copy engine, no host round trip, no ordering between the
 * device threads -- and emits its shard into its own output slice (hb_shard_emit).  Two earlier
 * forms of the exchange are kept behind HB_MULTI_PUSH: 1 = hb_push_map_kernel (peer stores) +
 * events + hb_shard_compose, 0 = the right-hand device pulls the 256 bytes with
 * cudaMemcpyPeerAsync (what is used where a pair of devices has no peer access).  Nothing here
 * decodes on the CPU.
 *
 * Two ways in:
 *   resident   hb_multi_load / hb_multi_generate, then hb_multi_decode (device-timed: CUDA
 *              events per device, the maximum is reported) and hb_multi_download / _verify
 *   host       hb_multi_decode_host: upload, decode, download -- what b200ApproachMulti does
 */
#include "huffb200.h"

#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <new>
#include <thread>

extern "C" int hb_gen_ctx_stream(hb_ctx *ctx, int *device, void **stream);

struct hb_mdev {
    int device = 0;
    hb_ctx *ctx = nullptr;
    hb_codebook *cb = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_start = nullptr, ev_map = nullptr, ev_end = nullptr;
    uint8_t *d_comp = nullptr;
    size_t comp_cap = 0;
    uint8_t *d_out = nullptr;
    size_t out_cap = 0;
    uint64_t *d_maps = nullptr;    /* HB_MULTI_MAX x 32 words: slot r = shard r's map */
    uint64_t *d_eb = nullptr;      /* 4 words: entry, base, total through this shard */
    uint64_t *h_eb = nullptr;      /* pinned copy */
    /* shard geometry */
    uint64_t a = 0, bytes = 0, readable = 0, bits_own = 0, bits_avail = 0;
    uint64_t n_symbols = 0, out_base = 0;
};

/* One host thread per device: the ~25 CUDA runtime calls a decode queues on a device take
 * ~60 us of host time, which, issued from a single thread for 8 devices one after the other,
 * was most of the wall time of a 1 GiB-symbol decode on 8 GPUs (0.55 ms against 0.25 ms with
 * one process per GPU).  The workers are parked on a condition variable between calls. */
struct hb_pool {
    std::thread th[HB_MULTI_MAX];
    std::mutex mu;
    std::condition_variable cv_job, cv_done;
    std::function<int(int)> job;
    uint64_t job_seq = 0;
    int n = 0, pending = 0;
    int rc[HB_MULTI_MAX];
    bool quit = false;
};

struct hb_peer_maps { uint64_t *p[HB_MULTI_MAX]; };

/* device i's map (32 words) into slot i of the map tables of the devices to its right: warp w of
 * the one CTA serves peer w.  The stores travel over NVLink; the event recorded behind this kernel
 * is what the receivers' streams wait for. */
__global__ void hb_push_map_kernel(const uint64_t *__restrict__ src, hb_peer_maps dst, int n_peers) {
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (w < n_peers) dst.p[w][l] = src[l];
}

struct hb_multi {
    int n = 0;
    bool push_ok = false;          /* every device can store into every other device's memory */
    bool xchg_ok = false;          /* ... and the contexts are connected for hb_shard_exchange (the default) */
    hb_pool *pool = nullptr;
    std::atomic<uint64_t> map_issued[HB_MULTI_MAX];   /* run number whose ev_map record has been queued */
    std::atomic<int> abort_run{0};
    uint64_t run_no = 0;
    hb_mdev d[HB_MULTI_MAX];
    hb_node_abi *tree = nullptr;
    int nodes = 0;
    uint32_t minlen = 1;
    int n_active = 0;              /* shards of the loaded stream */
    uint64_t bits = 0, n_symbols = 0;
    bool loaded = false;
    char err[256];
};

static int mfail(hb_multi *m, cudaError_t e, const char *what) {
    snprintf(m->err, sizeof(m->err), "%s: %s", what, cudaGetErrorString(e));
    return HB_ERR_CUDA;
}
#define MCK(call)                                                   \
    do {                                                            \
        cudaError_t e_ = (call);                                    \
        if (e_ != cudaSuccess) return mfail(m, e_, #call);          \
    } while (0)
#define MRC(call)                                                   \
    do {                                                            \
        int rc_ = (call);                                           \
        if (rc_ != HB_OK) {                                         \
            snprintf(m->err, sizeof(m->err), "%s: %s", #call, hb_strerror(rc_)); \
            return rc_;                                             \
        }                                                           \
    } while (0)

static double now_ms(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC_RAW, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

static void pool_worker(hb_pool *p, int i, int device) {
    cudaSetDevice(device);
    uint64_t seen = 0;
    for (;;) {
        std::function<int(int)> job;
        {
            std::unique_lock<std::mutex> lk(p->mu);
            p->cv_job.wait(lk, [&] { return p->quit || p->job_seq != seen; });
            if (p->quit) return;
            seen = p->job_seq;
            job = p->job;
        }
        const int rc = job(i);
        {
            std::lock_guard<std::mutex> lk(p->mu);
            p->rc[i] = rc;
            if (--p->pending == 0) p->cv_done.notify_all();
        }
    }
}

/* run fn(i) for every device i < n on its own thread; first non-zero result */
static int pool_run(hb_multi *m, int n, const std::function<int(int)> &fn) {
    hb_pool *p = m->pool;
    if (!p || n <= 1) {
        int rc = HB_OK;
        for (int i = 0; i < n && rc == HB_OK; i++) rc = fn(i);
        return rc;
    }
    {
        std::unique_lock<std::mutex> lk(p->mu);
        p->job = [fn, n](int i) { return i < n ? fn(i) : HB_OK; };
        p->pending = p->n;
        p->job_seq++;
    }
    p->cv_job.notify_all();
    std::unique_lock<std::mutex> lk(p->mu);
    p->cv_done.wait(lk, [&] { return p->pending == 0; });
    for (int i = 0; i < n; i++)
        if (p->rc[i] != HB_OK) return p->rc[i];
    return HB_OK;
}

extern "C" const char *hb_multi_last_error(const hb_multi *m) { return m ? m->err : "no multi context"; }
extern "C" int hb_multi_devices(const hb_multi *m) { return m ? m->n : 0; }

extern "C" void hb_multi_destroy(hb_multi *m) {
    if (!m) return;
    if (m->pool) {
        {
            std::lock_guard<std::mutex> lk(m->pool->mu);
            m->pool->quit = true;
        }
        m->pool->cv_job.notify_all();
        for (int i = 0; i < m->pool->n; i++) m->pool->th[i].join();
        delete m->pool;
        m->pool = nullptr;
    }
    for (int i = 0; i < m->n; i++) {
        hb_mdev &v = m->d[i];
        cudaSetDevice(v.device);
        if (v.stream) cudaStreamSynchronize(v.stream);
        if (v.cb) hb_codebook_destroy(v.cb);
        if (v.d_comp) cudaFree(v.d_comp);
        if (v.d_out) cudaFree(v.d_out);
        if (v.d_maps) cudaFree(v.d_maps);
        if (v.d_eb) cudaFree(v.d_eb);
        if (v.h_eb) cudaFreeHost(v.h_eb);
        if (v.ev_start) cudaEventDestroy(v.ev_start);
        if (v.ev_map) cudaEventDestroy(v.ev_map);
        if (v.ev_end) cudaEventDestroy(v.ev_end);
        if (v.ctx) hb_ctx_destroy(v.ctx);
    }
    free(m->tree);
    delete m;
}

extern "C" int hb_multi_create(const int *devices, int n_devices, hb_multi **out) {
    if (!out) return HB_ERR_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { cudaGetLastError(); return HB_ERR_CUDA; }
    if (n_devices <= 0) n_devices = ndev;
    if (n_devices > HB_MULTI_MAX || n_devices > ndev) return HB_ERR_ARG;
    hb_multi *m = new (std::nothrow) hb_multi();
    if (!m) return HB_ERR_NOMEM;
    m->err[0] = 0;
    for (int i = 0; i < n_devices; i++) {
        hb_mdev &v = m->d[i];
        v.device = devices ? devices[i] : i;
        for (int j = 0; j < i; j++)
            if (m->d[j].device == v.device) { hb_multi_destroy(m); return HB_ERR_ARG; }
        int rc = hb_ctx_create(v.device, nullptr, &v.ctx);
        m->n = i + 1;
        if (rc != HB_OK) { hb_multi_destroy(m); return rc; }
        void *sp = nullptr;
        int dv = 0;
        hb_gen_ctx_stream(v.ctx, &dv, &sp);
        v.stream = (cudaStream_t)sp;
        cudaSetDevice(v.device);
        cudaError_t e = cudaEventCreate(&v.ev_start);
        if (e == cudaSuccess) e = cudaEventCreate(&v.ev_map);
        if (e == cudaSuccess) e = cudaEventCreate(&v.ev_end);
        if (e == cudaSuccess) e = cudaMalloc((void **)&v.d_maps, sizeof(uint64_t) * 32 * HB_MULTI_MAX);
        if (e == cudaSuccess) e = cudaMalloc((void **)&v.d_eb, sizeof(uint64_t) * 4);
        if (e == cudaSuccess) e = cudaMallocHost((void **)&v.h_eb, sizeof(uint64_t) * 4);
        if (e != cudaSuccess) { cudaGetLastError(); hb_multi_destroy(m); return HB_ERR_CUDA; }
    }
    /* peer access where the topology allows it: maps are then pushed by a kernel (peer stores);
     * otherwise pulled by 256-byte peer copies, which work either way */
    bool all_peers = m->n > 1;
    for (int i = 0; i < m->n; i++) {
        cudaSetDevice(m->d[i].device);
        for (int j = 0; j < m->n; j++) {
            if (i == j) continue;
            int can = 0;
            bool ok = false;
            if (cudaDeviceCanAccessPeer(&can, m->d[i].device, m->d[j].device) == cudaSuccess && can) {
                const cudaError_t e = cudaDeviceEnablePeerAccess(m->d[j].device, 0);
                ok = e == cudaSuccess || e == cudaErrorPeerAccessAlreadyEnabled;
            }
            if (!ok) all_peers = false;
        }
    }
    cudaGetLastError();
    /* HB_MULTI_PUSH: 0 = peer copies pulled by the receiver, 1 = hb_push_map_kernel + events + hb_shard_compose,
     * unset / 2 = hb_shard_exchange (stores, wait and composition in one kernel, no host-side ordering) */
    const char *pm = getenv("HB_MULTI_PUSH");
    m->push_ok = all_peers && !(pm && pm[0] == '0');
    if (m->push_ok && !(pm && pm[0] == '1')) {
        hb_ctx *cs[HB_MULTI_MAX];
        for (int i = 0; i < m->n; i++) cs[i] = m->d[i].ctx;
        bool ok = true;
        for (int i = 0; i < m->n && ok; i++) ok = hb_peer_connect_local(cs[i], i, m->n, cs) == HB_OK;
        m->xchg_ok = ok;
    }
    for (int i = 0; i < HB_MULTI_MAX; i++) m->map_issued[i].store(0);
    if (m->n > 1 && !(getenv("HB_MULTI_THREADS") && getenv("HB_MULTI_THREADS")[0] == '0')) {
        m->pool = new (std::nothrow) hb_pool();
        if (m->pool) {
            m->pool->n = m->n;
            for (int i = 0; i < m->n; i++) m->pool->th[i] = std::thread(pool_worker, m->pool, i, m->d[i].device);
        }
    }
    *out = m;
    return HB_OK;
}

static int grow(hb_multi *m, hb_mdev &v, uint8_t **p, size_t *cap, size_t bytes) {
    if (bytes <= *cap) return HB_OK;
    MCK(cudaSetDevice(v.device));
    MCK(cudaStreamSynchronize(v.stream));
    if (*p) { MCK(cudaFree(*p)); *p = nullptr; *cap = 0; }
    cudaError_t e = cudaMalloc((void **)p, bytes);
    if (e != cudaSuccess) {
        cudaGetLastError();
        snprintf(m->err, sizeof(m->err), "cudaMalloc(%zu) on device %d: %s", bytes, v.device, cudaGetErrorString(e));
        *p = nullptr;
        return HB_ERR_NOMEM;
    }
    *cap = bytes;
    return HB_OK;
}

static int set_tree(hb_multi *m, const hb_node_abi *tree, int nodes) {
    bool same = m->tree && m->nodes == nodes;
    for (int i = 0; same && i < nodes; i++)
        same = tree[i].sym == m->tree[i].sym && tree[i].izero == m->tree[i].izero && tree[i].ione == m->tree[i].ione;
    if (same) return HB_OK;
    for (int i = 0; i < m->n; i++)
        if (m->d[i].cb) { hb_codebook_destroy(m->d[i].cb); m->d[i].cb = nullptr; }
    free(m->tree);
    m->tree = nullptr;
    m->nodes = 0;
    for (int i = 0; i < m->n; i++) MRC(hb_codebook_create(m->d[i].ctx, tree, nodes, &m->d[i].cb));
    m->tree = (hb_node_abi *)malloc(sizeof(hb_node_abi) * (size_t)nodes);
    if (!m->tree) return HB_ERR_NOMEM;
    for (int i = 0; i < nodes; i++) { m->tree[i].sym = tree[i].sym; m->tree[i].izero = tree[i].izero; m->tree[i].ione = tree[i].ione; }
    m->nodes = nodes;
    uint32_t mn = 1;
    MRC(hb_codebook_info(m->d[0].cb, nullptr, &mn, nullptr, nullptr));
    m->minlen = mn ? mn : 1;
    return HB_OK;
}

/* byte-range shards of a stream of `bits` bits; shards smaller than 64 KiB are not worth a device */
static void cut_shards(hb_multi *m, uint64_t bits) {
    const uint64_t nbytes = (bits + 7) / 8;
    int n = m->n;
    while (n > 1 && nbytes / (uint64_t)n < (64u << 10)) n--;
    const uint64_t per = (nbytes / (uint64_t)n) / 16 * 16;
    m->n_active = n;
    m->bits = bits;
    for (int i = 0; i < n; i++) {
        hb_mdev &v = m->d[i];
        const bool last = i == n - 1;
        v.a = (uint64_t)i * per;
        const uint64_t b = last ? nbytes : (uint64_t)(i + 1) * per;
        const uint64_t halo_end = last ? nbytes : (b + 16 < nbytes ? b + 16 : nbytes);
        v.bytes = halo_end - v.a;
        v.readable = (v.bytes + 15) / 16 * 16 + 32;
        v.bits_own = last ? bits - 8 * v.a : 8 * (b - v.a);
        v.bits_avail = last ? v.bits_own : (bits - 8 * v.a < 8 * v.bytes ? bits - 8 * v.a : 8 * v.bytes);
        v.n_symbols = v.out_base = 0;
    }
}

/* Per device (on its own host thread): map, then the maps of the left neighbours are copied
 * over and composed; optionally the emit right behind.  All asynchronous on the device's
 * stream.  A cross-device cudaStreamWaitEvent only orders against an event record that has
 * already been QUEUED, so device i first waits (on the host) until its left neighbours have
 * queued theirs for this run. */
static int dev_fail(hb_multi *m, int i, int rc, const char *what) {
    snprintf(m->err, sizeof(m->err), "device %d: %s: %s (%s)", m->d[i].device, what, hb_strerror(rc),
             hb_last_error(m->d[i].ctx));
    m->abort_run.store(1);
    return rc;
}
static int dev_cuda(hb_multi *m, int i, cudaError_t e, const char *what) {
    snprintf(m->err, sizeof(m->err), "device %d: %s: %s", m->d[i].device, what, cudaGetErrorString(e));
    m->abort_run.store(1);
    return HB_ERR_CUDA;
}
#define DCK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return dev_cuda(m, i, e_, #call); } while (0)

static int dev_map(hb_multi *m, int i, uint64_t run, bool upload, const uint8_t *data) {
    hb_mdev &v = m->d[i];
    int rc;
    cudaError_t e = cudaSetDevice(v.device);
    if (e == cudaSuccess && upload) {
        if (v.bytes) e = cudaMemcpyAsync(v.d_comp, data + v.a, v.bytes, cudaMemcpyHostToDevice, v.stream);
        if (e == cudaSuccess) e = cudaMemsetAsync(v.d_comp + v.bytes, 0, v.readable - v.bytes, v.stream);
    }
    if (e == cudaSuccess) e = cudaEventRecord(v.ev_start, v.stream);
    rc = e == cudaSuccess ? hb_ctx_set_shard_origin(v.ctx, v.a, 1) : HB_ERR_CUDA;
    if (rc == HB_OK)
        rc = hb_shard_map(v.ctx, v.cb, v.d_comp, v.readable, v.bits_own, v.bits_avail, v.d_maps + 32 * i);
    if (rc == HB_OK && m->xchg_ok) {
        /* stores into the right neighbours' tables, wait for the left neighbours, composition: one kernel */
        rc = hb_shard_exchange(v.ctx, run, v.d_eb);
        m->map_issued[i].store(run, std::memory_order_release);
        if (rc != HB_OK) return dev_fail(m, i, rc, "hb_shard_exchange");
        return HB_OK;
    }
    if (rc == HB_OK && m->push_ok && i + 1 < m->n_active) {
        hb_peer_maps pm;
        int np = 0;
        for (int r = i + 1; r < m->n_active; r++) pm.p[np++] = m->d[r].d_maps + 32 * i;
        hb_push_map_kernel<<<1, 32 * np, 0, v.stream>>>(v.d_maps + 32 * i, pm, np);
        if ((e = cudaGetLastError()) != cudaSuccess) rc = HB_ERR_CUDA;
    }
    if (rc == HB_OK && (e = cudaEventRecord(v.ev_map, v.stream)) != cudaSuccess) rc = HB_ERR_CUDA;
    if (rc != HB_OK) m->abort_run.store(1);
    m->map_issued[i].store(run, std::memory_order_release);   /* even on failure: nobody may wait for ever */
    if (e != cudaSuccess) return dev_cuda(m, i, e, "map phase");
    if (rc != HB_OK) return dev_fail(m, i, rc, "hb_shard_map");
    for (int r = 0; r < i; r++) {
        while (m->map_issued[r].load(std::memory_order_acquire) != run) std::this_thread::yield();
        if (m->abort_run.load()) return HB_ERR_STATE;
        DCK(cudaStreamWaitEvent(v.stream, m->d[r].ev_map, 0));
        if (!m->push_ok)
            DCK(cudaMemcpyPeerAsync(v.d_maps + 32 * r, v.device, m->d[r].d_maps + 32 * r, m->d[r].device,
                                    32 * sizeof(uint64_t), v.stream));
    }
    if ((rc = hb_shard_compose(v.ctx, v.d_maps, i + 1, i, v.d_eb))) return dev_fail(m, i, rc, "hb_shard_compose");
    return HB_OK;
}

static int dev_emit(hb_multi *m, int i) {
    hb_mdev &v = m->d[i];
    DCK(cudaSetDevice(v.device));
    int rc = hb_shard_emit(v.ctx, v.cb, v.d_comp, v.readable, v.bits_own, v.bits_avail, v.d_eb, v.d_out,
                           v.out_cap, nullptr);
    if (rc) return dev_fail(m, i, rc, "hb_shard_emit");
    DCK(cudaEventRecord(v.ev_end, v.stream));
    return HB_OK;
}

static int queue_maps(hb_multi *m, bool upload = false, const uint8_t *data = nullptr, bool emit_too = false) {
    const uint64_t run = ++m->run_no;
    m->abort_run.store(0);
    return pool_run(m, m->n_active, [m, run, upload, data, emit_too](int i) {
        int rc = dev_map(m, i, run, upload, data);
        if (rc == HB_OK && emit_too) rc = dev_emit(m, i);
        return rc;
    });
}

static int queue_emits(hb_multi *m) {
    return pool_run(m, m->n_active, [m](int i) { return dev_emit(m, i); });
}

/* entry/base/total of every shard back to the host (one small copy per device) */
static int fetch_bases(hb_multi *m) {
    for (int i = 0; i < m->n_active; i++) {
        hb_mdev &v = m->d[i];
        MCK(cudaSetDevice(v.device));
        MCK(cudaMemcpyAsync(v.h_eb, v.d_eb, 4 * sizeof(uint64_t), cudaMemcpyDeviceToHost, v.stream));
    }
    for (int i = 0; i < m->n_active; i++) {
        hb_mdev &v = m->d[i];
        MCK(cudaSetDevice(v.device));
        MCK(cudaStreamSynchronize(v.stream));
        v.out_base = v.h_eb[1];
        v.n_symbols = v.h_eb[2] - v.h_eb[1];
    }
    hb_mdev &l = m->d[m->n_active - 1];
    m->n_symbols = l.h_eb[2];
    /* a last codeword cut off by the end of the stream emits nothing (hb_shard_emit applies the
     * same rule): its count is only known after the emit; the upper bound is used for sizing */
    return HB_OK;
}

static int size_outputs(hb_multi *m) {
    for (int i = 0; i < m->n_active; i++) {
        hb_mdev &v = m->d[i];
        int rc = grow(m, v, &v.d_out, &v.out_cap, (size_t)v.n_symbols + 64);
        if (rc) return rc;
    }
    return HB_OK;
}

static int collect(hb_multi *m, hb_multi_result *res, double wall_ms) {
    uint64_t total = 0;
    float ms_max = 0;
    uint32_t launches = 0;
    if (res) memset(res, 0, sizeof(*res));
    for (int i = 0; i < m->n_active; i++) {
        hb_mdev &v = m->d[i];
        hb_result r;
        int rc = hb_shard_result(v.ctx, &r);
        if (rc != HB_OK) {
            snprintf(m->err, sizeof(m->err), "shard %d: %s (%s)", i, hb_strerror(rc), hb_last_error(v.ctx));
            return rc;
        }
        MCK(cudaSetDevice(v.device));
        float ms = 0;
        MCK(cudaEventElapsedTime(&ms, v.ev_start, v.ev_end));
        v.n_symbols = r.n_symbols;
        v.out_base = r.out_base;
        total += r.n_symbols;
        launches += r.launches + 1;   /* + hb_compose_kernel */
        if (ms > ms_max) ms_max = ms;
        if (res) { res->shard_symbols[i] = r.n_symbols; res->shard_ms[i] = ms; }
    }
    m->n_symbols = total;
    if (res) {
        res->n_symbols = total;
        res->n_devices = m->n_active;
        res->ms_device_max = ms_max;
        res->ms_wall = (float)wall_ms;
        res->launches = launches;
    }
    return HB_OK;
}

/* ---- resident path --------------------------------------------------------------- */

static int after_load(hb_multi *m) {
    /* one map pass sizes every device's output slice exactly */
    int rc = queue_maps(m);
    if (rc) return rc;
    if ((rc = fetch_bases(m))) return rc;
    if ((rc = size_outputs(m))) return rc;
    m->loaded = true;
    return HB_OK;
}

extern "C" int hb_multi_load(hb_multi *m, const hb_node_abi *tree, int nodes, const uint8_t *data,
                             uint64_t bits) {
    if (!m || !tree || (!data && bits)) return HB_ERR_ARG;
    m->loaded = false;
    int rc = set_tree(m, tree, nodes);
    if (rc) return rc;
    cut_shards(m, bits);
    for (int i = 0; i < m->n_active; i++) {
        hb_mdev &v = m->d[i];
        if ((rc = grow(m, v, &v.d_comp, &v.comp_cap, (size_t)v.readable))) return rc;
        MCK(cudaSetDevice(v.device));
        if (v.bytes) MCK(cudaMemcpyAsync(v.d_comp, data + v.a, v.bytes, cudaMemcpyHostToDevice, v.stream));
        MCK(cudaMemsetAsync(v.d_comp + v.bytes, 0, v.readable - v.bytes, v.stream));
    }
    return after_load(m);
}

extern "C" int hb_multi_generate(hb_multi *m, int model_kind, uint64_t seed, uint64_t n_symbols,
                                 uint64_t *bits_out) {
    if (!m) return HB_ERR_ARG;
    m->loaded = false;
    hb_model *mod = (hb_model *)malloc(sizeof(hb_model));
    if (!mod) return HB_ERR_NOMEM;
    int rc = hb_model_build(model_kind, mod);
    if (rc == HB_OK) rc = set_tree(m, mod->tree, mod->nodes);
    uint64_t bits = 0;
    if (rc == HB_OK) rc = hb_gen_count_bits_device(m->d[0].ctx, mod, seed, 0, n_symbols, &bits);
    if (rc != HB_OK) { free(mod); snprintf(m->err, sizeof(m->err), "generator setup: %s", hb_strerror(rc)); return rc; }
    cut_shards(m, bits);
    const uint64_t whole_cap = ((bits + 7) / 8 + 15) / 16 * 16 + 64;
    for (int i = 0; i < m->n_active && rc == HB_OK; i++) {
        /* every device encodes the whole stream (setup, untimed) and keeps its byte range */
        hb_mdev &v = m->d[i];
        void *whole = nullptr;
        uint64_t b2 = 0;
        rc = grow(m, v, &v.d_comp, &v.comp_cap, (size_t)v.readable);
        if (rc) break;
        cudaSetDevice(v.device);
        if (cudaMalloc(&whole, whole_cap) != cudaSuccess) { cudaGetLastError(); rc = HB_ERR_NOMEM; break; }
        rc = hb_gen_encode_device(v.ctx, mod, seed, 0, n_symbols, whole, whole_cap, &b2);
        if (rc == HB_OK && b2 != bits) rc = HB_ERR_STATE;
        if (rc == HB_OK) {
            cudaError_t e = cudaMemsetAsync(v.d_comp, 0, v.readable, v.stream);
            if (e == cudaSuccess)
                e = cudaMemcpyAsync(v.d_comp, (const uint8_t *)whole + v.a, v.bytes, cudaMemcpyDeviceToDevice, v.stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(v.stream);
            if (e != cudaSuccess) rc = mfail(m, e, "shard copy");
        }
        cudaFree(whole);
    }
    free(mod);
    if (rc != HB_OK) return rc;
    if (bits_out) *bits_out = bits;
    return after_load(m);
}

extern "C" int hb_multi_decode(hb_multi *m, hb_multi_result *res) {
    if (!m || !m->loaded) return HB_ERR_STATE;
    const double t0 = now_ms();
    int rc = queue_maps(m, false, nullptr, true);
    if (rc != HB_OK) return rc;
    for (int i = 0; i < m->n_active; i++) {
        MCK(cudaSetDevice(m->d[i].device));
        MCK(cudaStreamSynchronize(m->d[i].stream));
    }
    return collect(m, res, now_ms() - t0);
}

extern "C" int hb_multi_download(hb_multi *m, uint8_t *out, uint64_t out_capacity) {
    if (!m || !m->loaded || (!out && out_capacity)) return HB_ERR_STATE;
    if (m->n_symbols > out_capacity) return HB_ERR_OUTPUT_FULL;
    for (int i = 0; i < m->n_active; i++) {
        hb_mdev &v = m->d[i];
        MCK(cudaSetDevice(v.device));
        if (v.n_symbols)
            MCK(cudaMemcpyAsync(out + v.out_base, v.d_out, v.n_symbols, cudaMemcpyDeviceToHost, v.stream));
    }
    for (int i = 0; i < m->n_active; i++) {
        MCK(cudaSetDevice(m->d[i].device));
        MCK(cudaStreamSynchronize(m->d[i].stream));
    }
    return HB_OK;
}

extern "C" int hb_multi_verify(hb_multi *m, int model_kind, uint64_t seed, uint64_t *mismatches) {
    if (!m || !m->loaded || !mismatches) return HB_ERR_STATE;
    hb_model *mod = (hb_model *)malloc(sizeof(hb_model));
    if (!mod) return HB_ERR_NOMEM;
    int rc = hb_model_build(model_kind, mod);
    uint64_t bad_total = 0;
    for (int i = 0; i < m->n_active && rc == HB_OK; i++) {
        hb_mdev &v = m->d[i];
        uint64_t bad = 0;
        rc = hb_gen_verify_device(v.ctx, mod, seed, v.out_base, v.n_symbols, v.d_out, &bad);
        bad_total += bad;
    }
    free(mod);
    *mismatches = bad_total;
    return rc;
}

/* ---- host buffers in, host buffers out ---------------------------------------------
 * Every device: upload its shard, map; the maps travel; compose; the host reads the bases
 * (32 bytes per device: the only host synchronisation before the emit), emit, download into
 * the caller's buffer at the shard's base.  The copies of the N devices run concurrently
 * when the caller's buffers are page-locked (hb_host_pin; b200ApproachMulti does that). */
extern "C" int hb_multi_decode_host(hb_multi *m, const hb_node_abi *tree, int nodes, const uint8_t *data,
                                    uint64_t bits, uint8_t *out, uint64_t out_capacity,
                                    hb_multi_result *res) {
    if (!m || !tree || (!data && bits) || (!out && out_capacity)) return HB_ERR_ARG;
    const double t0 = now_ms();
    m->loaded = false;
    int rc = set_tree(m, tree, nodes);
    if (rc) return rc;
    cut_shards(m, bits);
    for (int i = 0; i < m->n_active; i++) {
        hb_mdev &v = m->d[i];
        if ((rc = grow(m, v, &v.d_comp, &v.comp_cap, (size_t)v.readable))) return rc;
    }
    if ((rc = queue_maps(m, true, data))) return rc;
    if ((rc = fetch_bases(m))) return rc;
    if (m->n_symbols > out_capacity + 1) {   /* + 1: a cut-off last codeword is counted until the emit */
        snprintf(m->err, sizeof(m->err), "decoded size %llu exceeds the output buffer (%llu)",
                 (unsigned long long)m->n_symbols, (unsigned long long)out_capacity);
        return HB_ERR_OUTPUT_FULL;
    }
    if ((rc = size_outputs(m))) return rc;
    if ((rc = queue_emits(m))) return rc;
    /* the emit's own count (cut-off rule applied) before the download is sized */
    if ((rc = collect(m, res, 0.0))) return rc;
    if (m->n_symbols > out_capacity) return HB_ERR_OUTPUT_FULL;
    for (int i = 0; i < m->n_active; i++) {
        hb_mdev &v = m->d[i];
        MCK(cudaSetDevice(v.device));
        if (v.n_symbols)
            MCK(cudaMemcpyAsync(out + v.out_base, v.d_out, v.n_symbols, cudaMemcpyDeviceToHost, v.stream));
    }
    for (int i = 0; i < m->n_active; i++) {
        MCK(cudaSetDevice(m->d[i].device));
        MCK(cudaStreamSynchronize(m->d[i].stream));
    }
    if (res) res->ms_wall = (float)(now_ms() - t0);
    m->loaded = true;
    return HB_OK;
}
