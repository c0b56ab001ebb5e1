// C ABI of the engine: the reference's rpg_* entry points (LogitWrapper.h:27-35)
// plus the bl_* extensions declared in include/bayeslogit_b200.h.
//
// Host-pointer entry points stream the batch through HBM in chunks over three
// CUDA streams (H2D of chunk k+1 and D2H of chunk k-1 overlap the kernel of chunk
// k; fully asynchronous when the caller's buffers are pinned).  There is no CPU
// path: any CUDA failure is reported and the outputs are left untouched, which
// is how the reference behaves on an exception (LogitWrapper.cpp:226-229).
#include <cuda_runtime.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "engine.h"

#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace bl {

namespace {

constexpr int kSlots = 3;

// Large one-shot copies between pageable and pinned memory: streaming (non-temporal) stores, so the destination
// lines are not read into the cache first -- a third less memory traffic than a cached memcpy, and the copy is
// bound by exactly that traffic.
#if defined(__x86_64__)
__attribute__((target("avx2"))) void copy_stream_avx2(char *d, const char *s, size_t n)
{
    while (n && (reinterpret_cast<uintptr_t>(d) & 31)) { *d++ = *s++; --n; }
    for (; n >= 128; n -= 128, d += 128, s += 128) {
        __m256i a = _mm256_loadu_si256((const __m256i *)s), b = _mm256_loadu_si256((const __m256i *)(s + 32));
        __m256i c = _mm256_loadu_si256((const __m256i *)(s + 64)), e = _mm256_loadu_si256((const __m256i *)(s + 96));
        _mm256_stream_si256((__m256i *)d, a);
        _mm256_stream_si256((__m256i *)(d + 32), b);
        _mm256_stream_si256((__m256i *)(d + 64), c);
        _mm256_stream_si256((__m256i *)(d + 96), e);
    }
    _mm_sfence();
    if (n) memcpy(d, s, n);
}
#endif

void copy_stream(char *d, const char *s, size_t n)
{
#if defined(__x86_64__)
    static const bool avx2 = __builtin_cpu_supports("avx2");
    if (avx2 && n >= 4096) { copy_stream_avx2(d, s, n); return; }
#endif
    memcpy(d, s, n);
}

// Host threads that copy between the caller's PAGEABLE buffers and the pipeline's pinned staging buffers.
// R's .C() interface hands over ordinary (pageable) vectors (LogitWrapper.R:29,49); a cudaMemcpyAsync from
// such memory is staged by the driver on one thread at a fraction of the link's bandwidth and blocks the
// caller meanwhile.  Here every chunk is copied by all pool threads at once into pinned memory and goes
// over the link as a true asynchronous DMA that overlaps the kernels of the neighbouring chunks.
class CopyPool {
    struct Job { char *d; const char *s; size_t n; };
    std::vector<std::thread> threads_;
    std::vector<Job> jobs_;
    std::mutex mu_;
    std::condition_variable go_, done_;
    uint64_t generation_ = 0;
    int pending_ = 0;
    bool stop_ = false;

    void worker(int k)
    {
        uint64_t seen = 0;
        for (;;) {
            Job j;
            {
                std::unique_lock<std::mutex> lk(mu_);
                go_.wait(lk, [&] { return stop_ || generation_ != seen; });
                if (stop_) return;
                seen = generation_;
                j = jobs_[k];
            }
            if (j.n) copy_stream(j.d, j.s, j.n);
            {
                std::lock_guard<std::mutex> lk(mu_);
                if (--pending_ == 0) done_.notify_one();
            }
        }
    }

public:
    int size() const { return (int)threads_.size(); }
    void start()
    {
        if (!threads_.empty()) return;
        const char *env = getenv("BAYESLOGIT_COPY_THREADS");
        int hw = (int)std::thread::hardware_concurrency();
        int n = env ? atoi(env) : (hw > 2 ? (hw - 1 < 24 ? hw - 1 : 24) : 1);
        if (n < 1) n = 1;
        jobs_.resize(n);
        for (int k = 0; k < n; ++k) threads_.emplace_back([this, k] { worker(k); });
    }
    // dst[0..bytes) = src[0..bytes), split into 64-byte-aligned pieces over the pool (the caller takes one too)
    void copy(void *dst, const void *src, size_t bytes)
    {
        const int n = size();
        if (n == 0 || bytes < (1u << 20)) { memcpy(dst, src, bytes); return; }
        size_t piece = ((bytes / (n + 1)) + 63) & ~(size_t)63;
        size_t off = 0;
        {
            std::lock_guard<std::mutex> lk(mu_);
            for (int k = 0; k < n; ++k) {
                size_t len = off < bytes ? (bytes - off < piece ? bytes - off : piece) : 0;
                jobs_[k] = Job{(char *)dst + off, (const char *)src + off, len};
                off += len;
            }
            pending_ = n;
            ++generation_;
        }
        go_.notify_all();
        if (off < bytes) copy_stream((char *)dst + off, (const char *)src + off, bytes - off);
        std::unique_lock<std::mutex> lk(mu_);
        done_.wait(lk, [&] { return pending_ == 0; });
    }
    ~CopyPool()
    {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
        }
        go_.notify_all();
        for (auto &t : threads_) t.join();
    }
};
// observations per full-size pipeline chunk.  With the ramped schedule (run_host) and 100M mixed draws:
// 1M 1.72e9, 2M 2.71e9, 4M 2.97e9, 8M 2.89e9, 16M 2.77e9, 32M 2.39e9 draws/s end to end (PCIe bound 3.1-3.3e9):
// smaller chunks shorten fill and drain, larger ones amortise the ~27 launches of a binned batch.
constexpr int64_t kChunkDefault = 1 << 22;

struct Slot {
    cudaStream_t stream = nullptr;
    void *shape = nullptr;
    double *z = nullptr, *x = nullptr;
    int *iter = nullptr;
    void *work = nullptr;      // regime-binning scratch (pg_hybrid.cu)
    int64_t cap = 0;
    // pinned staging buffers for pageable callers (allocated on the first such call)
    void *h_shape = nullptr;
    double *h_z = nullptr, *h_x = nullptr;
    int *h_iter = nullptr;
    int64_t h_cap = 0;
};

constexpr int64_t kStageMin = 1 << 18;   // smaller pageable batches go through the driver's own staging

constexpr int64_t kBinMin = 1 << 15;         // below this the per-lane dispatch kernel is used
constexpr int64_t kBinMax = 1 << 30;         // observations per binned launch (int32 index lists)

struct Context {
    std::mutex mu;
    bool ready = false;
    int device = -1;
    uint64_t seed = 0;
    uint32_t call = 0;
    Slot slot[kSlots];
    CopyPool pool;
    std::string err;
};

Context g;
std::atomic<uint64_t> g_launches{0};
thread_local std::string t_err;

int fail(const std::string &msg)
{
    t_err = msg;
    g.err = msg;
    fprintf(stderr, "Error: %s\n", msg.c_str());
    return 1;
}

#define BL_CK(expr)                                                                         \
    do {                                                                                    \
        cudaError_t e_ = (expr);                                                            \
        if (e_ != cudaSuccess)                                                              \
            return fail(std::string(#expr) + ": " + cudaGetErrorString(e_));                \
    } while (0)

int ensure_ready()
{
    if (g.ready) return 0;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(std::string("bayeslogit_b200: no usable CUDA device (") +
                    (e != cudaSuccess ? cudaGetErrorString(e) : "device count 0") +
                    "); this engine has no CPU fallback");
    int dev = g.device;
    if (dev < 0) {
        const char *env = getenv("BAYESLOGIT_DEVICE");
        if (!env) env = getenv("LOCAL_RANK");
        dev = env ? atoi(env) % count : 0;
    }
    BL_CK(cudaSetDevice(dev));
    cudaDeviceProp prop;
    BL_CK(cudaGetDeviceProperties(&prop, dev));
    if (prop.major < 10)
        return fail(std::string("bayeslogit_b200: device '") + prop.name +
                    "' is not sm_100 class; this library carries sm_100a code only");
    g.device = dev;
    for (int s = 0; s < kSlots; ++s) BL_CK(cudaStreamCreateWithFlags(&g.slot[s].stream, cudaStreamNonBlocking));
    {
        // the sweeps take their scratch from the stream-ordered pool; keep it mapped between calls
        // (a fresh 2 GB of scratch costs ~0.1-0.2 s to map, every call, with the default threshold 0)
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
            const char *env = getenv("BAYESLOGIT_POOL_KEEP_MB");
            uint64_t keep = (env ? strtoull(env, nullptr, 0) : 8192ull) << 20;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
    }
    if (g.seed == 0) {
        const char *env = getenv("BAYESLOGIT_SEED");
        g.seed = env ? strtoull(env, nullptr, 0)
                     : (uint64_t)std::chrono::high_resolution_clock::now().time_since_epoch().count() *
                           0x9E3779B97F4A7C15ull;
    }
    g.ready = true;
    return 0;
}

int slot_reserve(Slot &s, int64_t n)
{
    if (n <= s.cap) return 0;
    BL_CK(cudaStreamSynchronize(s.stream));
    if (s.shape) cudaFree(s.shape);
    if (s.z) cudaFree(s.z);
    if (s.x) cudaFree(s.x);
    if (s.iter) cudaFree(s.iter);
    if (s.work) cudaFree(s.work);
    s.shape = s.z = s.x = nullptr;
    s.iter = nullptr;
    s.work = nullptr;
    s.cap = 0;
    BL_CK(cudaMalloc(&s.shape, n * sizeof(double)));
    BL_CK(cudaMalloc((void **)&s.z, n * sizeof(double)));
    BL_CK(cudaMalloc((void **)&s.x, n * sizeof(double)));
    BL_CK(cudaMalloc((void **)&s.iter, n * sizeof(int)));
    BL_CK(cudaMalloc(&s.work, hybrid_workspace_bytes(n)));
    s.cap = n;
    return 0;
}

int stage_reserve(Slot &s, int64_t n)
{
    if (n <= s.h_cap) return 0;
    BL_CK(cudaStreamSynchronize(s.stream));
    if (s.h_shape) cudaFreeHost(s.h_shape);
    if (s.h_z) cudaFreeHost(s.h_z);
    if (s.h_x) cudaFreeHost(s.h_x);
    if (s.h_iter) cudaFreeHost(s.h_iter);
    s.h_shape = nullptr; s.h_z = s.h_x = nullptr; s.h_iter = nullptr; s.h_cap = 0;
    BL_CK(cudaHostAlloc(&s.h_shape, n * sizeof(double), cudaHostAllocDefault));
    BL_CK(cudaHostAlloc((void **)&s.h_z, n * sizeof(double), cudaHostAllocDefault));
    BL_CK(cudaHostAlloc((void **)&s.h_x, n * sizeof(double), cudaHostAllocDefault));
    BL_CK(cudaHostAlloc((void **)&s.h_iter, n * sizeof(int), cudaHostAllocDefault));
    s.h_cap = n;
    return 0;
}

// true when the driver knows nothing about the pointer: ordinary malloc'ed (pageable) host memory
bool is_pageable(const void *p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return true; }
    return a.type == cudaMemoryTypeUnregistered;
}

size_t shape_size(Method m)
{
    return (m == kDevroye || m == kDevroyePlain || m == kDevroyeLoop) ? sizeof(int) : sizeof(double);
}

// BAYESLOGIT_PIPE_CHUNK_LOG2 (measurement aid): observations per full-size chunk, 2^20 .. 2^26
int64_t pipeline_chunk()
{
    static const int64_t chunk = [] {
        const char *env = getenv("BAYESLOGIT_PIPE_CHUNK_LOG2");
        int lg = env ? atoi(env) : 0;
        return (lg >= 20 && lg <= 26) ? (int64_t)1 << lg : kChunkDefault;
    }();
    return chunk;
}

// Chunk sizes of a host-pointer batch of `num` observations, in order (see run_host).
std::vector<int64_t> pipeline_schedule(int64_t num, int64_t chunk)
{
    std::vector<int64_t> sizes, tl;
    const int64_t ramp[3] = {chunk / 8, chunk / 4, chunk / 2};
    int64_t head = 0, tail = 0;
    if (num >= 4 * chunk) {
        for (int64_t r : ramp) { sizes.push_back(r); head += r; }
        for (int64_t r : ramp) { tl.push_back(r); tail += r; }
    }
    int64_t body = num - head - tail;
    while (body > 0) {
        int64_t n = body < chunk ? body : chunk;
        sizes.push_back(n);
        body -= n;
    }
    for (auto it = tl.rbegin(); it != tl.rend(); ++it) sizes.push_back(*it);
    return sizes;
}

// Host-pointer batch through the chunk pipeline.
int run_host(Method m, double *x, const void *shape, const double *z, int64_t num, int trunc,
             int *iter, StreamId id)
{
    std::lock_guard<std::mutex> lock(g.mu);
    if (ensure_ready()) return 1;
    if (num < 0) return fail("negative batch size");
    if (num == 0) return 0;
    if (!x || !shape || !z) return fail("null argument");
    // Chunk schedule: the pipeline's fill (H2D of the first chunk, nothing to compute yet) and drain
    // (kernel + D2H of the last chunk, nothing left to copy in) are pure latency, so the batch opens
    // and closes with small chunks -- kChunk / 8, / 4, / 2, then kChunk-sized ones, then / 2, / 4, / 8 -- and the
    // link stays busy from ~0.3 ms after the call until ~0.4 ms before it returns.  Results do not
    // depend on the schedule (streams are keyed by the global observation index).
    const int64_t kChunk = pipeline_chunk();
    const std::vector<int64_t> sizes = pipeline_schedule(num, kChunk);
    const int64_t nchunks = (int64_t)sizes.size();
    for (int s = 0; s < kSlots && s < nchunks; ++s)
        if (slot_reserve(g.slot[s], num < kChunk ? num : kChunk)) return 1;
    size_t ss = shape_size(m);
    // Pageable caller buffers (what R's .C() passes): stage every chunk through pinned memory with the copy pool.
    // BAYESLOGIT_NO_STAGING=1 keeps the driver's own staging (measurement aid).
    const bool staged = num >= kStageMin && !getenv("BAYESLOGIT_NO_STAGING") &&
                        (is_pageable(x) || is_pageable(shape) || is_pageable(z));
    if (staged) {
        g.pool.start();
        for (int s = 0; s < kSlots && s < nchunks; ++s)
            if (stage_reserve(g.slot[s], num < kChunk ? num : kChunk)) return 1;
    }
    std::vector<int64_t> offs(nchunks);
    {
        int64_t o = 0;
        for (int64_t c = 0; c < nchunks; ++c) { offs[c] = o; o += sizes[c]; }
    }
    // results of chunk c leave the staging buffer once its stream has drained (its slot is reused by chunk c + kSlots)
    auto stage_out = [&](int64_t c) -> int {
        Slot &s = g.slot[c % kSlots];
        BL_CK(cudaStreamSynchronize(s.stream));
        g.pool.copy(x + offs[c], s.h_x, sizes[c] * sizeof(double));
        if (iter) g.pool.copy(iter + offs[c], s.h_iter, sizes[c] * sizeof(int));
        return 0;
    };
    for (int64_t c = 0; c < nchunks; ++c) {
        Slot &s = g.slot[c % kSlots];
        const int64_t n = sizes[c], off = offs[c];
        const void *src_shape = (const char *)shape + off * ss;
        const double *src_z = z + off;
        const int *src_iter = iter ? iter + off : nullptr;
        double *dst_x = x + off;
        int *dst_iter = iter ? iter + off : nullptr;
        if (staged) {
            if (c >= kSlots && stage_out(c - kSlots)) return 1;
            g.pool.copy(s.h_shape, src_shape, n * ss);
            g.pool.copy(s.h_z, src_z, n * sizeof(double));
            if (iter) g.pool.copy(s.h_iter, src_iter, n * sizeof(int));
            src_shape = s.h_shape; src_z = s.h_z; src_iter = s.h_iter; dst_x = s.h_x; dst_iter = s.h_iter;
        }
        BL_CK(cudaMemcpyAsync(s.shape, src_shape, n * ss, cudaMemcpyHostToDevice, s.stream));
        BL_CK(cudaMemcpyAsync(s.z, src_z, n * sizeof(double), cudaMemcpyHostToDevice, s.stream));
        if (iter) BL_CK(cudaMemcpyAsync(s.iter, src_iter, n * sizeof(int), cudaMemcpyHostToDevice, s.stream));
        StreamId cid = id;
        cid.obs0 += (uint64_t)off;
        if (m == kHybrid && n >= kBinMin)
            BL_CK(launch_hybrid_binned(s.x, (const double *)s.shape, s.z, (int)n, cid, s.work, s.stream));
        else if (m == kDevroye)
            BL_CK(launch_devroye_refill(s.x, (const int *)s.shape, s.z, n, cid, s.stream, s.work));
        else
            BL_CK(launch_rpg(m, s.x, s.shape, s.z, n, trunc, iter ? s.iter : nullptr, cid, s.stream));
        BL_CK(cudaMemcpyAsync(dst_x, s.x, n * sizeof(double), cudaMemcpyDeviceToHost, s.stream));
        if (iter) BL_CK(cudaMemcpyAsync(dst_iter, s.iter, n * sizeof(int), cudaMemcpyDeviceToHost, s.stream));
    }
    if (staged) {
        for (int64_t c = nchunks > kSlots ? nchunks - kSlots : 0; c < nchunks; ++c)
            if (stage_out(c)) return 1;
    } else {
        for (int s = 0; s < kSlots && s < nchunks; ++s) BL_CK(cudaStreamSynchronize(g.slot[s].stream));
    }
    return 0;
}

// Drop-in flavour: global seed + call counter.
int run_dropin(Method m, double *x, const void *shape, const double *z, const int *num, int trunc,
               int *iter)
{
    if (!num) return fail("null argument");
    uint64_t seed;
    uint32_t call;
    {
        std::lock_guard<std::mutex> lock(g.mu);
        if (ensure_ready()) return 1;
        seed = g.seed;
        call = g.call++;
    }
    return run_host(m, x, shape, z, (int64_t)*num, trunc, iter, StreamId{seed, 0, call});
}

int run_dev(Method m, double *x, const void *shape, const double *z, int64_t num, int trunc,
            int *iter, StreamId id, void *stream)
{
    {
        std::lock_guard<std::mutex> lock(g.mu);
        if (ensure_ready()) return 1;
    }
    if (num < 0) return fail("negative batch size");
    if ((m == kHybrid && num >= kBinMin) || (m == kDevroye && num >= (1 << 20))) {
        // index lists of the binned launches: stream-ordered scratch of THIS call on the caller's stream
        // (the pool keeps it mapped), so concurrent calls on different streams never share a buffer
        int64_t per = num < kBinMax ? num : kBinMax;
        void *work = nullptr;
        BL_CK(cudaMallocAsync(&work, hybrid_workspace_bytes(per), (cudaStream_t)stream));
        int rc = 0;
        for (int64_t off = 0; off < num && !rc; off += per) {
            int64_t n = num - off < per ? num - off : per;
            StreamId cid = id;
            cid.obs0 += (uint64_t)off;
            cudaError_t e = m == kHybrid
                ? launch_hybrid_binned(x + off, (const double *)shape + off, z + off, (int)n, cid, work, (cudaStream_t)stream)
                : launch_devroye_refill(x + off, (const int *)shape + off, z + off, n, cid, (cudaStream_t)stream, work);
            if (e != cudaSuccess) rc = fail(std::string("binned launch: ") + cudaGetErrorString(e));
        }
        cudaFreeAsync(work, (cudaStream_t)stream);
        return rc;
    }
    BL_CK(launch_rpg(m, x, shape, z, num, trunc, iter, id, (cudaStream_t)stream));
    return 0;
}

struct DevBuf {
    void *p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    int put(const void *host, size_t bytes)
    {
        if (bytes == 0) return 0;
        BL_CK(cudaMalloc(&p, bytes));
        if (host) {
            // a pageable-memory cudaMemcpy may return once the data sits in the driver's staging buffer;
            // the kernels run on non-blocking streams that do not order against the legacy stream, so
            // wait for the DMA itself before anything can read the buffer
            BL_CK(cudaMemcpy(p, host, bytes, cudaMemcpyHostToDevice));
            BL_CK(cudaStreamSynchronize(cudaStreamLegacy));
        }
        return 0;
    }
    int get(void *host, size_t bytes)
    {
        if (bytes == 0 || !host) return 0;
        BL_CK(cudaMemcpy(host, p, bytes, cudaMemcpyDeviceToHost));
        return 0;
    }
};

int run_tape(Method m, double *x, const void *shape, const double *z, int64_t num, int trunc,
             int *iter, const bl_tape *tape, int *trace)
{
    std::lock_guard<std::mutex> lock(g.mu);
    if (ensure_ready()) return 1;
    if (num <= 0) return num < 0 ? fail("negative batch size") : 0;
    if (!x || !shape || !z || !tape) return fail("null argument");
    DevBuf dx, dshape, dz, diter, dtr, du, de, dn, dg;
    size_t n = (size_t)num;
    if (dx.put(nullptr, n * 8) || dshape.put(shape, n * shape_size(m)) || dz.put(z, n * 8)) return 1;
    if (iter && diter.put(iter, n * 4)) return 1;
    if (trace && dtr.put(nullptr, n * BL_TRACE_W * 4)) return 1;
    if (tape->tu && du.put(tape->tu, n * tape->lu * 8)) return 1;
    if (tape->te && de.put(tape->te, n * tape->le * 8)) return 1;
    if (tape->tn && dn.put(tape->tn, n * tape->ln * 8)) return 1;
    if (tape->tg && dg.put(tape->tg, n * tape->lg * 8)) return 1;
    DevTape tp{(const double *)du.p, (const double *)de.p, (const double *)dn.p, (const double *)dg.p,
               tape->lu, tape->le, tape->ln, tape->lg};
    cudaStream_t st = g.slot[0].stream;
    BL_CK(launch_rpg_tape(m, (double *)dx.p, dshape.p, (const double *)dz.p, num, trunc,
                          iter ? (int *)diter.p : nullptr, tp, trace ? (int *)dtr.p : nullptr, st));
    BL_CK(cudaStreamSynchronize(st));
    if (dx.get(x, n * 8)) return 1;
    if (iter && diter.get(iter, n * 4)) return 1;
    if (trace && dtr.get(trace, n * BL_TRACE_W * 4)) return 1;
    return 0;
}

}  // namespace

void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }
bool pdl_enabled()
{
    static const bool on = getenv("BL_GIBBS_NO_PDL") == nullptr;
    return on;
}

}  // namespace bl

using namespace bl;

extern "C" {

// ---- Part 1: reference entry points -------------------------------------------------

void rpg_gamma(double *x, double *n, double *z, int *num, int *trunc)
{
    run_dropin(kGamma, x, n, z, num, trunc ? *trunc : 200, nullptr);
}

void rpg_devroye(double *x, int *n, double *z, int *num)
{
    run_dropin(kDevroye, x, n, z, num, 0, nullptr);
}

void rpg_alt(double *x, double *h, double *z, int *num)
{
    run_dropin(kAlt, x, h, z, num, 0, nullptr);
}

void rpg_sp(double *x, double *h, double *z, int *num, int *iter)
{
    run_dropin(kSP, x, h, z, num, 0, iter);
}

void rpg_hybrid(double *x, double *h, double *z, int *num)
{
    run_dropin(kHybrid, x, h, z, num, 0, nullptr);
}

// ---- Part 2: extensions ---------------------------------------------------------------

int bl_version(void) { return 100; }

const char *bl_last_error(void) { return g.err.c_str(); }

void bl_clear_error(void)
{
    std::lock_guard<std::mutex> lock(g.mu);
    g.err.clear();
}

int bl_set_device(int device)
{
    std::lock_guard<std::mutex> lock(g.mu);
    if (g.ready && device != g.device)
        return fail("bl_set_device: the engine is already bound to another device (one process per GPU)");
    g.device = device;
    return ensure_ready();
}

int bl_get_device(void) { return g.device; }

void bl_set_seed(uint64_t seed)
{
    std::lock_guard<std::mutex> lock(g.mu);
    g.seed = seed ? seed : 0x9E3779B97F4A7C15ull;
    g.call = 0;
}

uint64_t bl_get_seed(void) { return g.seed; }
uint32_t bl_get_call_counter(void) { return g.call; }

int bl_rpg_devroye_dev(double *x, const int *n, const double *z, int64_t num, uint64_t seed,
                       uint32_t call_id, uint64_t obs0, void *stream)
{
    return run_dev(kDevroye, x, n, z, num, 0, nullptr, StreamId{seed, obs0, call_id}, stream);
}
int bl_rpg_devroye_plain_dev(double *x, const int *n, const double *z, int64_t num, uint64_t seed,
                             uint32_t call_id, uint64_t obs0, void *stream)
{
    return run_dev(kDevroyePlain, x, n, z, num, 0, nullptr, StreamId{seed, obs0, call_id}, stream);
}
int bl_rpg_devroye_loop_dev(double *x, const int *n, const double *z, int64_t num, uint64_t seed,
                            uint32_t call_id, uint64_t obs0, void *stream)
{
    return run_dev(kDevroyeLoop, x, n, z, num, 0, nullptr, StreamId{seed, obs0, call_id}, stream);
}
int bl_rpg_devroye_plain_tape(double *x, const int *n, const double *z, int64_t num,
                              const bl_tape *tape, int *trace)
{
    return run_tape(kDevroyePlain, x, n, z, num, 0, nullptr, tape, trace);
}
int bl_rpg_gamma_dev(double *x, const double *n, const double *z, int64_t num, int trunc,
                     uint64_t seed, uint32_t call_id, uint64_t obs0, void *stream)
{
    return run_dev(kGamma, x, n, z, num, trunc, nullptr, StreamId{seed, obs0, call_id}, stream);
}
int bl_rpg_alt_dev(double *x, const double *h, const double *z, int64_t num, uint64_t seed,
                   uint32_t call_id, uint64_t obs0, void *stream)
{
    return run_dev(kAlt, x, h, z, num, 0, nullptr, StreamId{seed, obs0, call_id}, stream);
}
int bl_rpg_sp_dev(double *x, const double *h, const double *z, int64_t num, int *iter,
                  uint64_t seed, uint32_t call_id, uint64_t obs0, void *stream)
{
    return run_dev(kSP, x, h, z, num, 0, iter, StreamId{seed, obs0, call_id}, stream);
}
int bl_rpg_hybrid_dev(double *x, const double *h, const double *z, int64_t num, uint64_t seed,
                      uint32_t call_id, uint64_t obs0, void *stream)
{
    return run_dev(kHybrid, x, h, z, num, 0, nullptr, StreamId{seed, obs0, call_id}, stream);
}

int bl_rpg_devroye_seeded(double *x, const int *n, const double *z, int64_t num, uint64_t seed,
                          uint32_t call_id, uint64_t obs0)
{
    return run_host(kDevroye, x, n, z, num, 0, nullptr, StreamId{seed, obs0, call_id});
}
int bl_rpg_gamma_seeded(double *x, const double *n, const double *z, int64_t num, int trunc,
                        uint64_t seed, uint32_t call_id, uint64_t obs0)
{
    return run_host(kGamma, x, n, z, num, trunc, nullptr, StreamId{seed, obs0, call_id});
}
int bl_rpg_alt_seeded(double *x, const double *h, const double *z, int64_t num, uint64_t seed,
                      uint32_t call_id, uint64_t obs0)
{
    return run_host(kAlt, x, h, z, num, 0, nullptr, StreamId{seed, obs0, call_id});
}
int bl_rpg_sp_seeded(double *x, const double *h, const double *z, int64_t num, int *iter,
                     uint64_t seed, uint32_t call_id, uint64_t obs0)
{
    return run_host(kSP, x, h, z, num, 0, iter, StreamId{seed, obs0, call_id});
}
int bl_rpg_hybrid_seeded(double *x, const double *h, const double *z, int64_t num, uint64_t seed,
                         uint32_t call_id, uint64_t obs0)
{
    return run_host(kHybrid, x, h, z, num, 0, nullptr, StreamId{seed, obs0, call_id});
}

int bl_rpg_devroye_tape(double *x, const int *n, const double *z, int64_t num, const bl_tape *tape,
                        int *trace)
{
    return run_tape(kDevroye, x, n, z, num, 0, nullptr, tape, trace);
}
int bl_rpg_gamma_tape(double *x, const double *n, const double *z, int64_t num, int trunc,
                      const bl_tape *tape, int *trace)
{
    return run_tape(kGamma, x, n, z, num, trunc, nullptr, tape, trace);
}
int bl_rpg_alt_tape(double *x, const double *h, const double *z, int64_t num, const bl_tape *tape,
                    int *trace)
{
    return run_tape(kAlt, x, h, z, num, 0, nullptr, tape, trace);
}
int bl_rpg_sp_tape(double *x, const double *h, const double *z, int64_t num, int *iter,
                   const bl_tape *tape, int *trace)
{
    return run_tape(kSP, x, h, z, num, 0, iter, tape, trace);
}
int bl_rpg_hybrid_tape(double *x, const double *h, const double *z, int64_t num,
                       const bl_tape *tape, int *trace)
{
    return run_tape(kHybrid, x, h, z, num, 0, nullptr, tape, trace);
}

int bl_probe_pg_moments(double *m1, double *m2, const double *b, const double *z, int64_t num)
{
    std::lock_guard<std::mutex> lock(g.mu);
    if (ensure_ready()) return 1;
    DevBuf d1, d2, db, dz;
    size_t n = (size_t)num * 8;
    if (d1.put(nullptr, n) || d2.put(nullptr, n) || db.put(b, n) || dz.put(z, n)) return 1;
    BL_CK(launch_probe_moments((double *)d1.p, (double *)d2.p, (const double *)db.p,
                               (const double *)dz.p, num, g.slot[0].stream));
    BL_CK(cudaStreamSynchronize(g.slot[0].stream));
    return d1.get(m1, n) || d2.get(m2, n);
}

int bl_probe_v_eval(double *v, const double *y, int64_t num)
{
    std::lock_guard<std::mutex> lock(g.mu);
    if (ensure_ready()) return 1;
    DevBuf dv, dy;
    size_t n = (size_t)num * 8;
    if (dv.put(nullptr, n) || dy.put(y, n)) return 1;
    BL_CK(launch_probe_v_eval((double *)dv.p, (const double *)dy.p, num, g.slot[0].stream));
    BL_CK(cudaStreamSynchronize(g.slot[0].stream));
    return dv.get(v, n);
}

int bl_probe_specfun(double *out, int which, const double *a, const double *b, const double *c,
                     int64_t num)
{
    std::lock_guard<std::mutex> lock(g.mu);
    if (ensure_ready()) return 1;
    DevBuf dout, da, db, dc;
    size_t n = (size_t)num * 8;
    if (dout.put(nullptr, n) || da.put(a, n) || db.put(b ? b : a, n) || dc.put(c ? c : a, n)) return 1;
    BL_CK(launch_probe_specfun((double *)dout.p, which, (const double *)da.p, (const double *)db.p,
                               (const double *)dc.p, num, g.slot[0].stream));
    BL_CK(cudaStreamSynchronize(g.slot[0].stream));
    return dout.get(out, n);
}

int bl_probe_philox(uint32_t *out4, const uint32_t *ctr4, const uint32_t *key2)
{
    std::lock_guard<std::mutex> lock(g.mu);
    if (ensure_ready()) return 1;
    DevBuf dout, dc, dk;
    if (dout.put(nullptr, 16) || dc.put(ctr4, 16) || dk.put(key2, 8)) return 1;
    BL_CK(launch_probe_philox((uint32_t *)dout.p, (const uint32_t *)dc.p, (const uint32_t *)dk.p,
                              g.slot[0].stream));
    BL_CK(cudaStreamSynchronize(g.slot[0].stream));
    return dout.get(out4, 16);
}

int bl_probe_peaks(double *out6)
{
    std::lock_guard<std::mutex> lock(g.mu);
    if (ensure_ready()) return 1;
    if (!out6) return fail("null argument");
    std::string err;
    if (probe_peaks(out6, g.slot[0].stream, err)) return fail("bl_probe_peaks: " + err);
    return 0;
}

int bl_probe_dmma_scaling(double *out4)
{
    std::lock_guard<std::mutex> lock(g.mu);
    if (ensure_ready()) return 1;
    if (!out4) return fail("null argument");
    std::string err;
    if (probe_dmma_scaling(out4, g.slot[0].stream, err)) return fail("bl_probe_dmma_scaling: " + err);
    return 0;
}

uint64_t bl_kernel_launches(void) { return g_launches.load(); }

// Host logic only (no device): the chunk sizes run_host would use for a batch of `num` observations.
int bl_probe_pipeline_schedule(int64_t num, int64_t *sizes, int cap)
{
    if (num < 0) return -1;
    const std::vector<int64_t> v = bl::pipeline_schedule(num, bl::pipeline_chunk());
    for (size_t k = 0; k < v.size() && (int)k < cap; ++k) sizes[k] = v[k];
    return (int)v.size();
}

void bl_hybrid_timing(int enable) { hybrid_timing_enable(enable != 0); }

int bl_hybrid_timing_last(double *ms8, int *launches8) { return hybrid_timing_last(ms8, launches8); }

int bl_ensure_ready_internal(void)
{
    std::lock_guard<std::mutex> lock(g.mu);
    return ensure_ready();
}

void bl_set_error_internal(const char *msg) { g.err = msg ? msg : ""; }

void *bl_stream_internal(void) { return (void *)g.slot[0].stream; }

uint64_t bl_next_call_internal(void)
{
    std::lock_guard<std::mutex> lock(g.mu);
    return g.call++;
}

}  // extern "C"
