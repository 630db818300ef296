// C-ABI layer of libhlsd.so (include/hlsd.h): argument checking, design -> arithmetic variant
// mapping, host-buffer staging.  No CPU compute path exists here: every entry point ends in a
// CUDA kernel launch or fails.
#include <atomic>
#include <cstdio>
#include <string>
#include <vector>

#include "kernels.h"

namespace hlsd {

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

static int check_path(int flags, int* path) {
    if (flags & ~HLSD_PATH_MASK) return fail(HLSD_ERR_INVALID, "unknown flag bits");
    *path = flags & HLSD_PATH_MASK;
    if (*path > HLSD_PATH_TENSOR) return fail(HLSD_ERR_INVALID, "unknown kernel path");
    return 0;
}

// ---- device entry points ----------------------------------------------------------------------
static int chol_dev(const float* A, float* L, int n, int64_t batch, int design, int flags, cudaStream_t st) {
    int variant, path;
    if (!A || !L) return fail(HLSD_ERR_INVALID, "null pointer");
    if (n < 1 || batch < 0) return fail(HLSD_ERR_INVALID, "n must be >= 1 and batch >= 0");
    if (!chol_variant(design, &variant)) return fail(HLSD_ERR_INVALID, "unknown Cholesky design");
    if (int rc = check_path(flags, &path)) return rc;
    if (path == HLSD_PATH_BLOCKED || path == HLSD_PATH_TENSOR) {
        if (variant != 0) return fail(HLSD_ERR_UNSUPPORTED, "blocked Cholesky implements the right-looking designs only");
        cudaError_t e = cholesky_blocked(A, L, n, (long)batch, path == HLSD_PATH_TENSOR, st);
        if (e == cudaErrorNotSupported) return fail(HLSD_ERR_UNSUPPORTED, "blocked Cholesky needs n to be a multiple of 64, n >= 128");
        if (e != cudaSuccess) return cuda_fail(e, "cholesky_blocked");
        return 0;
    }
    cudaError_t e = cholesky_exact(A, L, n, (long)batch, variant, path, st);
    if (e == cudaErrorInvalidValue) return fail(HLSD_ERR_UNSUPPORTED, "requested kernel path cannot serve this size/alignment");
    if (e != cudaSuccess) return cuda_fail(e, "cholesky");
    return 0;
}

static int lu_dev(const float* A, float* L, float* U, int32_t* P, int n, int64_t batch, int design, int flags,
                  cudaStream_t st) {
    int pivoting, path;
    if (!A || !L || !U || !P) return fail(HLSD_ERR_INVALID, "null pointer");
    if (n < 1 || batch < 0) return fail(HLSD_ERR_INVALID, "n must be >= 1 and batch >= 0");
    if (!lu_pivoting(design, &pivoting)) return fail(HLSD_ERR_INVALID, "unknown LU design");
    if (int rc = check_path(flags, &path)) return rc;
    if (path == HLSD_PATH_BLOCKED || path == HLSD_PATH_TENSOR)
        return fail(HLSD_ERR_UNSUPPORTED, "blocked LU is not built in this version");
    cudaError_t e = lu_exact(A, L, U, P, n, (long)batch, pivoting, path, st);
    if (e == cudaErrorInvalidValue) return fail(HLSD_ERR_UNSUPPORTED, "requested kernel path cannot serve this size/alignment");
    if (e != cudaSuccess) return cuda_fail(e, "lu");
    return 0;
}

static int qr_dev(const float* A, float* Q, float* R, int rows, int cols, int64_t batch, int design, int flags,
                  cudaStream_t st) {
    int variant, path;
    if (!A || !Q || !R) return fail(HLSD_ERR_INVALID, "null pointer");
    // the reference's top() rejects ROWS < COLS (qr_v2.1/model4x4/qr.cpp:325-332)
    if (cols < 1 || rows < cols || batch < 0) return fail(HLSD_ERR_INVALID, "need rows >= cols >= 1 and batch >= 0");
    if (!qr_variant(design, &variant)) return fail(HLSD_ERR_INVALID, "unknown QR design");
    if (int rc = check_path(flags, &path)) return rc;
    if (path == HLSD_PATH_BLOCKED || path == HLSD_PATH_TENSOR)
        return fail(HLSD_ERR_UNSUPPORTED, "Givens QR has no blocked path");
    cudaError_t e = qr_exact(A, Q, R, rows, cols, (long)batch, variant, path, st);
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

template <typename LaunchF>
static int run_host_pipeline(int64_t batch, std::vector<Buf> bufs, LaunchF launch) {
    if (batch == 0) return 0;
    size_t per_matrix = 0;
    for (auto& b : bufs) per_matrix += b.bytes_per_matrix;
    const size_t kChunkBytes = (size_t)96 << 20;
    int64_t chunk = (int64_t)(kChunkBytes / (per_matrix ? per_matrix : 1));
    if (chunk < 1) chunk = 1;
    if (chunk > batch) chunk = batch;
    chunk = (chunk + 31) / 32 * 32;  // keep chunk starts 16-byte aligned for every element size used here
    const int kStreams = 3;
    cudaStream_t st[kStreams] = {};
    std::vector<void*> dev[kStreams];
    int rc = 0;
    cudaError_t e = cudaSuccess;
    int made = 0;
    for (; made < kStreams && made * chunk < batch; made++) {
        if ((e = cudaStreamCreateWithFlags(&st[made], cudaStreamNonBlocking)) != cudaSuccess) {
            rc = cuda_fail(e, "cudaStreamCreate");
            break;
        }
        for (auto& b : bufs) {
            void* p = nullptr;
            if ((e = cudaMalloc(&p, b.bytes_per_matrix * (size_t)chunk)) != cudaSuccess) {
                rc = cuda_fail(e, "cudaMalloc");
                break;
            }
            dev[made].push_back(p);
        }
        if (rc) {
            made++;
            break;
        }
    }
    if (!rc) {
        int slot = 0;
        for (int64_t m0 = 0; m0 < batch && !rc; m0 += chunk, slot = (slot + 1) % made) {
            const int64_t cnt = (batch - m0 < chunk) ? batch - m0 : chunk;
            for (size_t i = 0; i < bufs.size() && !rc; i++)
                if (bufs[i].host_in) {
                    e = cudaMemcpyAsync(dev[slot][i], (const char*)bufs[i].host_in + bufs[i].bytes_per_matrix * (size_t)m0,
                                        bufs[i].bytes_per_matrix * (size_t)cnt, cudaMemcpyHostToDevice, st[slot]);
                    if (e != cudaSuccess) rc = cuda_fail(e, "cudaMemcpyAsync H2D");
                }
            if (!rc) rc = launch(dev[slot], cnt, st[slot]);
            for (size_t i = 0; i < bufs.size() && !rc; i++)
                if (bufs[i].host_out) {
                    e = cudaMemcpyAsync((char*)bufs[i].host_out + bufs[i].bytes_per_matrix * (size_t)m0, dev[slot][i],
                                        bufs[i].bytes_per_matrix * (size_t)cnt, cudaMemcpyDeviceToHost, st[slot]);
                    if (e != cudaSuccess) rc = cuda_fail(e, "cudaMemcpyAsync D2H");
                }
        }
    }
    for (int i = 0; i < made; i++) {
        if (st[i]) {
            e = cudaStreamSynchronize(st[i]);
            if (e != cudaSuccess && !rc) rc = cuda_fail(e, "cudaStreamSynchronize");
        }
        for (void* p : dev[i]) cudaFree(p);
        if (st[i]) cudaStreamDestroy(st[i]);
    }
    return rc;
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
const char* hlsd_version(void) { return "hlsd 0.1 (sm_100a)"; }

}  // extern "C"
