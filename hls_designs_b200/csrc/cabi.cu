// C-ABI layer of libhlsd.so (include/hlsd.h): argument checking, design -> arithmetic variant
// mapping, host-buffer staging.  No CPU compute path exists here: every entry point ends in a
// CUDA kernel launch or fails.
#include <atomic>
#include <cstdio>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include "kernels.h"

namespace hlsd {

static std::atomic<int> g_options[HLSD_OPT_COUNT];
int option(int key) { return (key >= 0 && key < HLSD_OPT_COUNT) ? g_options[key].load(std::memory_order_relaxed) : 0; }

static std::atomic<int64_t> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

int sm_count() {
    static thread_local int cached_dev = -1, cached = 0;
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev != cached_dev) {
        cudaDeviceGetAttribute(&cached, cudaDevAttrMultiProcessorCount, dev);
        cached_dev = dev;
    }
    return cached;
}

// kernel attribute cache: (kernel, device) -> largest dynamic shared-memory size granted so far
cudaError_t ensure_dynamic_smem(const void* kernel, size_t bytes) {
    static std::mutex mu;
    static std::map<std::pair<const void*, int>, size_t> granted;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lock(mu);
    size_t& have = granted[{kernel, dev}];
    if (have >= bytes && have != 0) return cudaSuccess;
    if (bytes <= 48 * 1024 && have == 0) {  // the default limit already covers it
        have = 48 * 1024;
        return cudaSuccess;
    }
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess) have = bytes;
    return e;
}

// library-owned stream-ordered memory pool per device (workspaces of the blocked paths); memory stays reserved
// between calls (release threshold = max) until hlsd_release_workspace() trims it
static std::mutex g_pool_mu;
static std::map<int, cudaMemPool_t> g_pools;
static cudaError_t device_pool(cudaMemPool_t* out) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lock(g_pool_mu);
    auto it = g_pools.find(dev);
    if (it == g_pools.end()) {
        cudaMemPoolProps props = {};
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = dev;
        cudaMemPool_t pool;
        if ((e = cudaMemPoolCreate(&pool, &props)) != cudaSuccess) return e;
        uint64_t keep = UINT64_MAX;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        it = g_pools.emplace(dev, pool).first;
    }
    *out = it->second;
    return cudaSuccess;
}
cudaError_t ws_alloc(void** p, size_t bytes, cudaStream_t st) {
    cudaMemPool_t pool;
    cudaError_t e = device_pool(&pool);
    if (e != cudaSuccess) return e;
    return cudaMallocFromPoolAsync(p, bytes ? bytes : 16, pool, st);
}
void ws_free(void* p, cudaStream_t st) {
    if (p) cudaFreeAsync(p, st);
}

// ---- per-launch timing (hlsd_profile_begin / hlsd_profile_end) ----
static std::atomic<bool> g_prof_on{false};
static std::mutex g_prof_mu;
struct ProfRec {
    int cls;
    cudaEvent_t a, b;
};
static std::vector<ProfRec> g_prof;
ProfScope::ProfScope(int c, cudaStream_t s) : cls(c), st(s) {
    if (!g_prof_on.load(std::memory_order_relaxed)) return;
    if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) {
        if (a) cudaEventDestroy(a);
        a = b = nullptr;
        return;
    }
    cudaEventRecord(a, st);
}
ProfScope::~ProfScope() {
    if (!a) return;
    cudaEventRecord(b, st);
    std::lock_guard<std::mutex> lock(g_prof_mu);
    g_prof.push_back(ProfRec{cls, a, b});
}

static thread_local std::string g_err = "";

static int fail(hlsd_status st, const std::string& msg) {
    g_err = msg;
    return (int)st;
}
static int cuda_fail(cudaError_t e, const char* what) {
    return fail(HLSD_ERR_CUDA, std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")");
}

static bool chol_variant(int design, int* variant) {
    switch (design) {
        case HLSD_CHOL_V1_3:
        case HLSD_CHOL_V3_2:
        case HLSD_CHOL_V4_0: *variant = 0; return true;
        case HLSD_CHOL_V2_2: *variant = 1; return true;
        default: return false;
    }
}
static bool lu_pivoting(int design, int* pivoting) {
    switch (design) {
        case HLSD_LU_V1_1:
        case HLSD_LU_V1_2:
        case HLSD_LU_V2_1: *pivoting = 1; return true;
        case HLSD_LU_NOPIVOT: *pivoting = 0; return true;
        default: return false;
    }
}
static bool qr_variant(int design, int* variant) {
    switch (design) {
        case HLSD_QR_V2_1: *variant = 0; return true;
        case HLSD_QR_V1_1:
        case HLSD_QR_V1_2: *variant = 1; return true;
        default: return false;
    }
}

static int check_flags(int flags, int* path, bool* exact, bool* fused) {
    if (flags & ~(HLSD_PATH_MASK | HLSD_FLAG_MASK)) return fail(HLSD_ERR_INVALID, "unknown flag bits");
    *path = flags & HLSD_PATH_MASK;
    *exact = (flags & HLSD_FLAG_EXACT) != 0;
    *fused = (flags & HLSD_FLAG_FUSED) != 0;
    if (*path > HLSD_PATH_TENSOR) return fail(HLSD_ERR_INVALID, "unknown kernel path");
    if (*exact && *fused) return fail(HLSD_ERR_INVALID, "HLSD_FLAG_EXACT and HLSD_FLAG_FUSED exclude each other");
    if (*exact && *path == HLSD_PATH_TENSOR) return fail(HLSD_ERR_INVALID, "HLSD_FLAG_EXACT excludes HLSD_PATH_TENSOR");
    return 0;
}

// ---- device entry points ----------------------------------------------------------------------
static int chol_dev(const float* A, float* L, int n, int64_t batch, int design, int flags, cudaStream_t st) {
    int variant, path;
    bool exact, fused;
    if (!A || !L) return fail(HLSD_ERR_INVALID, "null pointer");
    if (n < 1 || batch < 0) return fail(HLSD_ERR_INVALID, "n must be >= 1 and batch >= 0");
    if (!chol_variant(design, &variant)) return fail(HLSD_ERR_INVALID, "unknown Cholesky design");
    if (int rc = check_flags(flags, &path, &exact, &fused)) return rc;
    // large right-looking problems go to the blocked tensor-core path unless a path was forced or the caller
    // asked for the csim's bits (HLSD_FLAG_EXACT)
    if (path == HLSD_PATH_AUTO && n >= 256 && n % 128 == 0 && (((uintptr_t)A | (uintptr_t)L) & 15) == 0)
        path = (exact || variant != 0) ? HLSD_PATH_BLOCKED : HLSD_PATH_TENSOR;  // both blocked; BLOCKED keeps the csim's bits
    if (path == HLSD_PATH_BLOCKED || path == HLSD_PATH_TENSOR) {
        if (variant != 0 && path == HLSD_PATH_TENSOR)
            return fail(HLSD_ERR_UNSUPPORTED, "the tensor-core Cholesky implements the right-looking designs only (cholesky_v2.2: HLSD_PATH_BLOCKED)");
        cudaError_t e = cudaSuccess;
        const int64_t kSlice = 32768;  // the blocked kernels index the batch with gridDim.y
        for (int64_t b0 = 0; b0 < batch && e == cudaSuccess; b0 += kSlice) {
            const int64_t cnt = batch - b0 < kSlice ? batch - b0 : kSlice;
            e = cholesky_blocked(A + b0 * n * n, L + b0 * n * n, n, (long)cnt, path == HLSD_PATH_TENSOR, variant, st);
        }
        if (e == cudaErrorNotSupported) return fail(HLSD_ERR_UNSUPPORTED, "the blocked Cholesky paths need n % 128 == 0, n >= 256 and 16-byte aligned buffers");
        if (e != cudaSuccess) return cuda_fail(e, "cholesky_blocked");
        return 0;
    }
    cudaError_t e = cholesky_exact(A, L, n, (long)batch, variant, path, fused, st);
    if (e == cudaErrorInvalidValue) return fail(HLSD_ERR_UNSUPPORTED, "requested kernel path cannot serve this size/alignment");
    if (e != cudaSuccess) return cuda_fail(e, "cholesky");
    return 0;
}

static int lu_dev(const float* A, float* L, float* U, int32_t* P, int n, int64_t batch, int design, int flags,
                  cudaStream_t st) {
    int pivoting, path;
    bool exact, fused;
    if (!A || !L || !U || !P) return fail(HLSD_ERR_INVALID, "null pointer");
    if (n < 1 || batch < 0) return fail(HLSD_ERR_INVALID, "n must be >= 1 and batch >= 0");
    if (!lu_pivoting(design, &pivoting)) return fail(HLSD_ERR_INVALID, "unknown LU design");
    if (int rc = check_flags(flags, &path, &exact, &fused)) return rc;
    if (path == HLSD_PATH_AUTO && n >= 256 && n <= 1024 && n % 128 == 0 &&
        (((uintptr_t)A | (uintptr_t)L | (uintptr_t)U) & 15) == 0)
        path = exact ? HLSD_PATH_BLOCKED : HLSD_PATH_TENSOR;  // both blocked; BLOCKED keeps the csim's bits
    if (path == HLSD_PATH_TENSOR || path == HLSD_PATH_BLOCKED) {
        cudaError_t e = cudaSuccess;
        const int64_t kSlice = 32768;
        for (int64_t b0 = 0; b0 < batch && e == cudaSuccess; b0 += kSlice) {
            const int64_t cnt = batch - b0 < kSlice ? batch - b0 : kSlice;
            e = path == HLSD_PATH_TENSOR
                    ? lu_blocked(A + b0 * n * n, L + b0 * n * n, U + b0 * n * n, P + b0 * n, n, (long)cnt, pivoting, st)
                    : lu_blocked_exact(A + b0 * n * n, L + b0 * n * n, U + b0 * n * n, P + b0 * n, n, (long)cnt, pivoting, st);
        }
        if (e == cudaErrorNotSupported) return fail(HLSD_ERR_UNSUPPORTED, "the blocked LU paths need n % 128 == 0, 256 <= n <= 1024 and 16-byte aligned buffers");
        if (e != cudaSuccess) return cuda_fail(e, "lu_blocked");
        return 0;
    }
    cudaError_t e = lu_exact(A, L, U, P, n, (long)batch, pivoting, path, fused, st);
    if (e == cudaErrorInvalidValue) return fail(HLSD_ERR_UNSUPPORTED, "requested kernel path cannot serve this size/alignment");
    if (e != cudaSuccess) return cuda_fail(e, "lu");
    return 0;
}

static int qr_dev(const float* A, float* Q, float* R, int rows, int cols, int64_t batch, int design, int flags,
                  cudaStream_t st) {
    int variant, path;
    bool exact, fused;
    if (!A || !Q || !R) return fail(HLSD_ERR_INVALID, "null pointer");
    // the reference's top() rejects ROWS < COLS (qr_v2.1/model4x4/qr.cpp:325-332)
    if (cols < 1 || rows < cols || batch < 0) return fail(HLSD_ERR_INVALID, "need rows >= cols >= 1 and batch >= 0");
    if (!qr_variant(design, &variant)) return fail(HLSD_ERR_INVALID, "unknown QR design");
    if (int rc = check_flags(flags, &path, &exact, &fused)) return rc;
    if (path == HLSD_PATH_BLOCKED || path == HLSD_PATH_TENSOR)
        return fail(HLSD_ERR_UNSUPPORTED, "Givens QR has no blocked path");
    cudaError_t e = qr_exact(A, Q, R, rows, cols, (long)batch, variant, path, fused, st);
    if (e == cudaErrorInvalidValue) return fail(HLSD_ERR_UNSUPPORTED, "requested kernel path cannot serve this size/alignment");
    if (e != cudaSuccess) return cuda_fail(e, "qr");
    return 0;
}

// ---- host staging -------------------------------------------------------------------------------
// Streams a batch through the device in chunks on a small ring of streams so that the copy-in of
// chunk i+1, the kernels of chunk i and the copy-out of chunk i-1 overlap (PCIe is full duplex).
struct Buf {
    const void* host_in;  // nullptr for pure outputs
    void* host_out;       // nullptr for pure inputs
    size_t bytes_per_matrix;
};

// Staging workspace: three streams and, per stream, one growable device buffer per argument.  (cudaMalloc /
// cudaFree / stream creation per call cost more than the whole factorisation of a small batch.)  Workspaces live
// in a process-wide pool keyed by device: a call takes one (or creates it), gives it back when it returns, and
// hlsd_release_workspace() frees the idle ones - nothing is tied to the calling thread, so short-lived threads
// leak nothing.  Buffers above kKeepBytes are dropped when the call ends instead of being kept.
struct Workspace {
    static constexpr int kStreams = 3;
    // Staging buffers above this size are freed when a call ends (a 4096 x 4096 matrix must not pin gigabytes for ever).
    // It was 256 MB, which made every hlsd_lu_host / hlsd_cholesky_host call at n = 1024 free and re-allocate its 0.6 GB
    // chunk buffers (bench: the LU-1024 host leg ran at half the rate of a copy-only pass with small records); a chunk
    // of one matrix per SM at n = 1024 now stays.  hlsd_release_workspace() still frees everything.
    static constexpr size_t kKeepBytes = (size_t)1536 << 20;
    int device = -1;
    cudaStream_t st[kStreams] = {};
    std::vector<void*> buf[kStreams];
    std::vector<size_t> cap[kStreams];

    void release() {
        for (int i = 0; i < kStreams; i++) {
            for (void* p : buf[i]) cudaFree(p);
            buf[i].clear();
            cap[i].clear();
            if (st[i]) cudaStreamDestroy(st[i]);
            st[i] = nullptr;
        }
    }
    void shrink() {
        for (int i = 0; i < kStreams; i++)
            for (size_t j = 0; j < buf[i].size(); j++)
                if (cap[i][j] > kKeepBytes) {
                    cudaFree(buf[i][j]);
                    buf[i][j] = nullptr;
                    cap[i][j] = 0;
                }
    }
    cudaError_t prepare(int slot, const std::vector<size_t>& bytes) {
        cudaError_t e;
        if (!st[slot] && (e = cudaStreamCreateWithFlags(&st[slot], cudaStreamNonBlocking)) != cudaSuccess) return e;
        if (buf[slot].size() < bytes.size()) {
            buf[slot].resize(bytes.size(), nullptr);
            cap[slot].resize(bytes.size(), 0);
        }
        for (size_t i = 0; i < bytes.size(); i++)
            if (cap[slot][i] < bytes[i]) {
                if (buf[slot][i]) cudaFree(buf[slot][i]);
                buf[slot][i] = nullptr;
                cap[slot][i] = 0;
                if ((e = cudaMalloc(&buf[slot][i], bytes[i])) != cudaSuccess) return e;
                cap[slot][i] = bytes[i];
            }
        return cudaSuccess;
    }
};
static std::mutex g_ws_mu;
static std::multimap<int, Workspace*> g_ws_idle;  // device -> idle workspaces

static Workspace* ws_acquire(int dev) {
    {
        std::lock_guard<std::mutex> lock(g_ws_mu);
        auto it = g_ws_idle.find(dev);
        if (it != g_ws_idle.end()) {
            Workspace* w = it->second;
            g_ws_idle.erase(it);
            return w;
        }
    }
    Workspace* w = new Workspace();
    w->device = dev;
    return w;
}
static void ws_giveback(Workspace* w) {
    w->shrink();
    std::lock_guard<std::mutex> lock(g_ws_mu);
    g_ws_idle.emplace(w->device, w);
}
static void ws_release_idle() {
    std::vector<Workspace*> all;
    {
        std::lock_guard<std::mutex> lock(g_ws_mu);
        for (auto& kv : g_ws_idle) all.push_back(kv.second);
        g_ws_idle.clear();
    }
    int cur = 0;
    const bool have_cur = cudaGetDevice(&cur) == cudaSuccess;
    for (Workspace* w : all) {
        if (cudaSetDevice(w->device) == cudaSuccess) w->release();
        delete w;
    }
    {
        std::lock_guard<std::mutex> lock(g_pool_mu);
        for (auto& kv : g_pools) cudaMemPoolTrimTo(kv.second, 0);
    }
    if (have_cur) cudaSetDevice(cur);
}

// Chunk size of the staging pipeline: about kChunkBytes of arguments per chunk, but at least one matrix per SM
// (the kernels for n >= 128 give every matrix one CTA or a few: a chunk of 5 matrices of 1024 x 1024 left the GPU
// 95 % idle and the pipeline kernel-bound at half the PCIe rate) as long as that stays under kMaxChunkBytes; never
// more than the batch; and - so that the copy-in of chunk i+1, the kernels of chunk i and the copy-out of chunk
// i-1 can overlap - no more than a third of the batch when the batch is large enough to be worth splitting.
// Every chunk is copied to the BASE of its staging buffers (cudaMalloc alignment), so chunk starts need no
// alignment of their own.
static int64_t staging_chunk(int64_t batch, size_t per_matrix) {
    const size_t kChunkBytes = (size_t)64 << 20, kMaxChunkBytes = (size_t)2 << 30, kMinSplitBytes = (size_t)8 << 20;
    if (per_matrix == 0) per_matrix = 1;
    int64_t chunk = (int64_t)(kChunkBytes / per_matrix);
    const int64_t fill = sm_count() > 0 ? sm_count() : 148;
    if (chunk < fill) chunk = std::min<int64_t>(fill, (int64_t)(kMaxChunkBytes / per_matrix));
    if (chunk < 1) chunk = 1;
    if (chunk > batch) chunk = batch;
    if (chunk == batch && batch >= 3 && (size_t)batch * per_matrix >= 3 * kMinSplitBytes) chunk = (batch + 2) / 3;
    return chunk;
}

template <typename LaunchF>
static int run_host_pipeline(int64_t batch, std::vector<Buf> bufs, LaunchF launch) {
    if (batch == 0) return 0;
    size_t per_matrix = 0;
    for (auto& b : bufs) per_matrix += b.bytes_per_matrix;
    const int64_t chunk = staging_chunk(batch, per_matrix);
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice");
    Workspace* ws = ws_acquire(dev);
    int rc = 0;
    int used = 0;
    std::vector<size_t> bytes;
    for (auto& b : bufs) bytes.push_back(b.bytes_per_matrix * (size_t)chunk);
    for (; used < Workspace::kStreams && used * chunk < batch; used++)
        if ((e = ws->prepare(used, bytes)) != cudaSuccess) {
            rc = cuda_fail(e, "staging workspace");
            break;
        }
    if (!rc) {
        int slot = 0;
        for (int64_t m0 = 0; m0 < batch && !rc; m0 += chunk, slot = (slot + 1) % used) {
            const int64_t cnt = (batch - m0 < chunk) ? batch - m0 : chunk;
            cudaStream_t st = ws->st[slot];
            std::vector<void*>& dbuf = ws->buf[slot];
            for (size_t i = 0; i < bufs.size() && !rc; i++)
                if (bufs[i].host_in) {
                    e = cudaMemcpyAsync(dbuf[i], (const char*)bufs[i].host_in + bufs[i].bytes_per_matrix * (size_t)m0,
                                        bufs[i].bytes_per_matrix * (size_t)cnt, cudaMemcpyHostToDevice, st);
                    if (e != cudaSuccess) rc = cuda_fail(e, "cudaMemcpyAsync H2D");
                }
            if (!rc) rc = launch(dbuf, cnt, st);
            for (size_t i = 0; i < bufs.size() && !rc; i++)
                if (bufs[i].host_out) {
                    e = cudaMemcpyAsync((char*)bufs[i].host_out + bufs[i].bytes_per_matrix * (size_t)m0, dbuf[i],
                                        bufs[i].bytes_per_matrix * (size_t)cnt, cudaMemcpyDeviceToHost, st);
                    if (e != cudaSuccess) rc = cuda_fail(e, "cudaMemcpyAsync D2H");
                }
        }
    }
    for (int i = 0; i < Workspace::kStreams; i++)
        if (ws->st[i]) {
            e = cudaStreamSynchronize(ws->st[i]);
            if (e != cudaSuccess && !rc) rc = cuda_fail(e, "cudaStreamSynchronize");
        }
    ws_giveback(ws);
    return rc;
}

// One host batch over several devices: contiguous balanced spans, one worker thread per device, each running
// `span(first, count)` (a *_host call on its slice) after cudaSetDevice.  The first failure is reported.
template <typename SpanF>
static int run_multi(int64_t batch, const int* devices, int ndevices, SpanF span) {
    std::vector<int> devs;
    if (!devices) {
        int n = 0;
        if (cudaGetDeviceCount(&n) != cudaSuccess || n < 1) return fail(HLSD_ERR_CUDA, "no CUDA device");
        if (ndevices > 0 && ndevices < n) n = ndevices;
        for (int i = 0; i < n; i++) devs.push_back(i);
    } else {
        if (ndevices < 1) return fail(HLSD_ERR_INVALID, "ndevices must be >= 1");
        devs.assign(devices, devices + ndevices);
    }
    const int nd = (int)devs.size();
    std::vector<int> rcs(nd, 0);
    std::vector<std::string> errs(nd);
    std::vector<std::thread> workers;
    const int64_t base = batch / nd, extra = batch % nd;
    int64_t first = 0;
    for (int i = 0; i < nd; i++) {
        const int64_t cnt = base + (i < extra ? 1 : 0);
        if (cnt > 0 || i == 0)
            workers.emplace_back([&, i, first, cnt]() {
                cudaError_t e = cudaSetDevice(devs[i]);
                if (e != cudaSuccess) {
                    rcs[i] = cuda_fail(e, "cudaSetDevice");
                } else {
                    rcs[i] = span(first, cnt);
                }
                if (rcs[i]) errs[i] = g_err;  // g_err is thread-local: carry the message to the caller's thread
            });
        first += cnt;
    }
    for (auto& w : workers) w.join();
    for (int i = 0; i < nd; i++)
        if (rcs[i]) {
            g_err = "device " + std::to_string(devs[i]) + ": " + errs[i];
            return rcs[i];
        }
    return 0;
}

}  // namespace hlsd

using namespace hlsd;

extern "C" {

int hlsd_cholesky_dev(const float* dA, float* dL, int n, int64_t batch, int design, int flags, void* stream) {
    return chol_dev(dA, dL, n, batch, design, flags, (cudaStream_t)stream);
}
int hlsd_lu_dev(const float* dA, float* dL, float* dU, int32_t* dP, int n, int64_t batch, int design, int flags,
                void* stream) {
    return lu_dev(dA, dL, dU, dP, n, batch, design, flags, (cudaStream_t)stream);
}
int hlsd_qr_dev(const float* dA, float* dQ, float* dR, int rows, int cols, int64_t batch, int design, int flags,
                void* stream) {
    return qr_dev(dA, dQ, dR, rows, cols, batch, design, flags, (cudaStream_t)stream);
}

int hlsd_cholesky_host(const float* A, float* L, int n, int64_t batch, int design, int flags) {
    if (!A || !L) return fail(HLSD_ERR_INVALID, "null pointer");
    if (n < 1 || batch < 0) return fail(HLSD_ERR_INVALID, "n must be >= 1 and batch >= 0");
    const size_t mb = sizeof(float) * (size_t)n * n;
    return run_host_pipeline(batch, {{A, nullptr, mb}, {nullptr, L, mb}},
                             [&](std::vector<void*>& d, int64_t cnt, cudaStream_t st) {
                                 return chol_dev((const float*)d[0], (float*)d[1], n, cnt, design, flags, st);
                             });
}

int hlsd_lu_host(const float* A, float* L, float* U, int32_t* P, int n, int64_t batch, int design, int flags) {
    if (!A || !L || !U || !P) return fail(HLSD_ERR_INVALID, "null pointer");
    if (n < 1 || batch < 0) return fail(HLSD_ERR_INVALID, "n must be >= 1 and batch >= 0");
    const size_t mb = sizeof(float) * (size_t)n * n;
    return run_host_pipeline(batch, {{A, nullptr, mb}, {nullptr, L, mb}, {nullptr, U, mb}, {nullptr, P, sizeof(int32_t) * (size_t)n}},
                             [&](std::vector<void*>& d, int64_t cnt, cudaStream_t st) {
                                 return lu_dev((const float*)d[0], (float*)d[1], (float*)d[2], (int32_t*)d[3], n, cnt,
                                               design, flags, st);
                             });
}

int hlsd_qr_host(const float* A, float* Q, float* R, int rows, int cols, int64_t batch, int design, int flags) {
    if (!A || !Q || !R) return fail(HLSD_ERR_INVALID, "null pointer");
    if (cols < 1 || rows < cols || batch < 0) return fail(HLSD_ERR_INVALID, "need rows >= cols >= 1 and batch >= 0");
    const size_t ab = sizeof(float) * (size_t)rows * cols, qb = sizeof(float) * (size_t)rows * rows;
    return run_host_pipeline(batch, {{A, nullptr, ab}, {nullptr, Q, qb}, {nullptr, R, ab}},
                             [&](std::vector<void*>& d, int64_t cnt, cudaStream_t st) {
                                 return qr_dev((const float*)d[0], (float*)d[1], (float*)d[2], rows, cols, cnt, design,
                                               flags, st);
                             });
}

void* hlsd_host_alloc(size_t bytes) {
    void* p = nullptr;
    cudaError_t e = cudaMallocHost(&p, bytes ? bytes : 1);
    if (e != cudaSuccess) {
        cuda_fail(e, "cudaMallocHost");
        return nullptr;
    }
    return p;
}
void hlsd_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

void hlsd_release_workspace(void) { ws_release_idle(); }

int hlsd_cholesky_host_multi(const float* A, float* L, int n, int64_t batch, int design, int flags,
                             const int* devices, int ndevices) {
    if (!A || !L) return fail(HLSD_ERR_INVALID, "null pointer");
    if (n < 1 || batch < 0) return fail(HLSD_ERR_INVALID, "n must be >= 1 and batch >= 0");
    const size_t mm = (size_t)n * n;
    return run_multi(batch, devices, ndevices, [&](int64_t first, int64_t cnt) {
        return hlsd_cholesky_host(A + first * mm, L + first * mm, n, cnt, design, flags);
    });
}
int hlsd_lu_host_multi(const float* A, float* L, float* U, int32_t* P, int n, int64_t batch, int design, int flags,
                       const int* devices, int ndevices) {
    if (!A || !L || !U || !P) return fail(HLSD_ERR_INVALID, "null pointer");
    if (n < 1 || batch < 0) return fail(HLSD_ERR_INVALID, "n must be >= 1 and batch >= 0");
    const size_t mm = (size_t)n * n;
    return run_multi(batch, devices, ndevices, [&](int64_t first, int64_t cnt) {
        return hlsd_lu_host(A + first * mm, L + first * mm, U + first * mm, P + first * n, n, cnt, design, flags);
    });
}
int hlsd_qr_host_multi(const float* A, float* Q, float* R, int rows, int cols, int64_t batch, int design, int flags,
                       const int* devices, int ndevices) {
    if (!A || !Q || !R) return fail(HLSD_ERR_INVALID, "null pointer");
    if (cols < 1 || rows < cols || batch < 0) return fail(HLSD_ERR_INVALID, "need rows >= cols >= 1 and batch >= 0");
    const size_t am = (size_t)rows * cols, qm = (size_t)rows * rows;
    return run_multi(batch, devices, ndevices, [&](int64_t first, int64_t cnt) {
        return hlsd_qr_host(A + first * am, Q + first * qm, R + first * am, rows, cols, cnt, design, flags);
    });
}

int hlsd_selftest_copy_host(const void* in, void* out, size_t in_bytes, size_t out_bytes, int64_t units) {
    if (!in || !out || units < 0 || in_bytes == 0 || out_bytes == 0)
        return fail(HLSD_ERR_INVALID, "hlsd_selftest_copy_host: bad argument");
    return run_host_pipeline(units, {{in, nullptr, in_bytes}, {nullptr, out, out_bytes}},
                             [](std::vector<void*>&, int64_t, cudaStream_t) { return 0; });
}

int hlsd_set_device(int device) {
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
    return 0;
}
int hlsd_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}
int64_t hlsd_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
const char* hlsd_last_error(void) { return g_err.c_str(); }
const char* hlsd_version(void) { return "hlsd 0.2 (sm_100a)"; }

int hlsd_set_option(int opt, int value) {
    if (opt < 1 || opt >= HLSD_OPT_COUNT) return fail(HLSD_ERR_INVALID, "unknown option");
    g_options[opt].store(value);
    return 0;
}

int hlsd_profile_begin(void) {
    std::lock_guard<std::mutex> lock(g_prof_mu);
    for (auto& r : g_prof) {
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
    }
    g_prof.clear();
    g_prof_on.store(true);
    return 0;
}
int hlsd_profile_end(double* ms, int64_t* launches, int n) {
    g_prof_on.store(false);
    if (!ms || n < 1) return fail(HLSD_ERR_INVALID, "hlsd_profile_end: bad argument");
    std::lock_guard<std::mutex> lock(g_prof_mu);
    int rc = 0;
    for (auto& r : g_prof) {
        float t = 0.0f;
        cudaError_t e = cudaEventSynchronize(r.b);
        if (e == cudaSuccess) e = cudaEventElapsedTime(&t, r.a, r.b);
        if (e != cudaSuccess && !rc) rc = cuda_fail(e, "hlsd_profile_end");
        if (e == cudaSuccess && r.cls >= 0 && r.cls < n) {
            ms[r.cls] += t;
            if (launches) launches[r.cls] += 1;
        }
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
    }
    g_prof.clear();
    return rc;
}

int hlsd_selftest_division(const float* num, const float* den, int64_t count, int64_t* out3, void* stream) {
    if (!num || !den || !out3 || count < 0) return fail(HLSD_ERR_INVALID, "hlsd_selftest_division: bad argument");
    cudaError_t e = hlsd::selftest_division(num, den, (long)count, reinterpret_cast<long long*>(out3), (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "selftest_division");
    return 0;
}

}  // extern "C"
