// C ABI of the B200-native NMFk hot path (include/nmfk_b200.h): context, X upload, batches of
// restarts, solve, robustness, and the one-call forms that mirror the reference's
// execute_run / execute (/root/reference/src/NMFkExecute.jl:178-233, 483-711).
// Host-side logic only; every numeric step runs in the CUDA kernels of this directory.
#include <dlfcn.h>
#include <nccl.h>  // types and prototypes only: the library is bound at run time (dlopen), see ShardNccl

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <numeric>
#include <string>
#include <vector>

#include "../../include/nmfk_b200.h"
#include "nmfk_internal.h"
#include "philox.h"

using namespace nmfk;

struct nmfk_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::vector<cudaStream_t> pool;
    std::vector<cudaEvent_t> pool_ev;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    void* Xp = nullptr;
    void* Xpt = nullptr;
    int64_t n = 0, m = 0;
    int dtype = NMFK_F64;
    double lambda = 1e-32;
    bool has_X = false;
    nmfk_xinfo info{};
    std::string err;
    int64_t launches = 0;
    double last_solve_ms = 0.0;
    double* d_partials = nullptr;  // residual partial sums
    size_t partials_cap = 0;
    // row-sharded X (nmfk_ctx_comm_init): this ctx holds rows [row0, row0 + n) of an n_global x m matrix
    ShardComm shard{nullptr, 1, 0, nullptr};
    bool sharded = false;
    int64_t row0 = 0, n_global = 0;
};

struct nmfk_batch {
    nmfk_ctx* ctx = nullptr;
    int k = 0, R = 0;
    void* W = nullptr;
    void* H = nullptr;
    UnitState* st = nullptr;
    int32_t* canon = nullptr;
    void* ximp = nullptr;
    bool inited = false;
};

static thread_local std::string g_err;
static int32_t residual(nmfk_ctx* c, int k, const void* W, const void* H, int restore, double weight, double out[2]);

static size_t esize(int dtype) { return dtype == NMFK_F64 ? 8 : 4; }

static int32_t fail(nmfk_ctx* c, int32_t code, const std::string& msg) {
    g_err = msg;
    if (c) c->err = msg;
    return code;
}

#define CU(ctx, call)                                                                                   \
    do {                                                                                                \
        cudaError_t e__ = (call);                                                                       \
        if (e__ != cudaSuccess)                                                                         \
            return fail((ctx), (int32_t)e__, std::string(#call) + ": " + cudaGetErrorString(e__));      \
    } while (0)

// NCCL is bound with dlopen at the first nmfk_comm_* call: in a process that already loaded a
// libnccl.so.2 (torch ships its own) that copy is reused, otherwise the system library is loaded;
// libnmfk_b200.so itself has no link-time dependency on NCCL.
namespace {
struct ShardNccl {
    void* lib = nullptr;
    decltype(&ncclGetUniqueId) getUniqueId = nullptr;
    decltype(&ncclCommInitRank) commInitRank = nullptr;
    decltype(&ncclAllReduce) allReduce = nullptr;
    decltype(&ncclCommDestroy) commDestroy = nullptr;
    decltype(&ncclGetErrorString) getErrorString = nullptr;
    std::string err;
    bool load() {
        if (lib) return true;
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (lib) break;
        }
        if (!lib) {
            err = std::string("cannot load NCCL: ") + dlerror();
            return false;
        }
        getUniqueId = reinterpret_cast<decltype(getUniqueId)>(dlsym(lib, "ncclGetUniqueId"));
        commInitRank = reinterpret_cast<decltype(commInitRank)>(dlsym(lib, "ncclCommInitRank"));
        allReduce = reinterpret_cast<decltype(allReduce)>(dlsym(lib, "ncclAllReduce"));
        commDestroy = reinterpret_cast<decltype(commDestroy)>(dlsym(lib, "ncclCommDestroy"));
        getErrorString = reinterpret_cast<decltype(getErrorString)>(dlsym(lib, "ncclGetErrorString"));
        if (!getUniqueId || !commInitRank || !allReduce || !commDestroy || !getErrorString) {
            err = "NCCL library lacks a required symbol";
            lib = nullptr;
            return false;
        }
        return true;
    }
};
ShardNccl g_nccl;

cudaError_t nccl_allreduce(void* comm, void* buf, size_t count, int dtype, cudaStream_t s) {
    const ncclResult_t r = g_nccl.allReduce(buf, buf, count, dtype == 1 ? ncclDouble : ncclFloat, ncclSum,
                                            static_cast<ncclComm_t>(comm), s);
    return r == ncclSuccess ? cudaSuccess : cudaErrorUnknown;
}
// single-rank stand-in (nranks == 1): the sum over one rank is the buffer itself
cudaError_t identity_allreduce(void*, void*, size_t, int, cudaStream_t) { return cudaSuccess; }
}  // namespace

// Definitions below inherit C linkage from their declarations in include/nmfk_b200.h.

int32_t nmfk_comm_unique_id(void* id128) {
    if (!id128) return fail(nullptr, NMFK_E_INVALID, "nmfk_comm_unique_id: NULL buffer");
    if (!g_nccl.load()) return fail(nullptr, NMFK_E_UNSUPPORTED, g_nccl.err);
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    const ncclResult_t r = g_nccl.getUniqueId(&id);
    if (r != ncclSuccess) return fail(nullptr, NMFK_E_UNSUPPORTED, std::string("ncclGetUniqueId: ") + g_nccl.getErrorString(r));
    std::memcpy(id128, &id, 128);
    return NMFK_OK;
}

int32_t nmfk_ctx_comm_init(nmfk_ctx* c, int32_t nranks, int32_t rank, const void* id128, int64_t row0, int64_t n_global) {
    if (!c || nranks < 1 || rank < 0 || rank >= nranks || row0 < 0 || n_global < 1)
        return fail(c, NMFK_E_INVALID, "nmfk_ctx_comm_init: bad arguments");
    if (c->sharded) return fail(c, NMFK_E_INVALID, "nmfk_ctx_comm_init: ctx already has a communicator");
    CU(c, cudaSetDevice(c->device));
    if (nranks == 1 && !id128) {
        c->shard = ShardComm{nullptr, 1, 0, identity_allreduce};
    } else {
        if (!id128) return fail(c, NMFK_E_INVALID, "nmfk_ctx_comm_init: NULL unique id");
        if (!g_nccl.load()) return fail(c, NMFK_E_UNSUPPORTED, g_nccl.err);
        ncclUniqueId id;
        std::memcpy(&id, id128, 128);
        ncclComm_t comm = nullptr;
        const ncclResult_t r = g_nccl.commInitRank(&comm, nranks, id, rank);
        if (r != ncclSuccess)
            return fail(c, NMFK_E_UNSUPPORTED, std::string("ncclCommInitRank: ") + g_nccl.getErrorString(r));
        c->shard = ShardComm{comm, nranks, rank, nccl_allreduce};
    }
    c->sharded = true;
    c->row0 = row0;
    c->n_global = n_global;
    return NMFK_OK;
}

int32_t nmfk_ctx_comm_destroy(nmfk_ctx* c) {
    if (!c) return fail(nullptr, NMFK_E_INVALID, "ctx is NULL");
    if (c->sharded && c->shard.comm) g_nccl.commDestroy(static_cast<ncclComm_t>(c->shard.comm));
    c->shard = ShardComm{nullptr, 1, 0, nullptr};
    c->sharded = false;
    return NMFK_OK;
}

int32_t nmfk_abi_version(void) { return NMFK_ABI_VERSION; }

const char* nmfk_last_error(const nmfk_ctx* ctx) { return ctx ? ctx->err.c_str() : g_err.c_str(); }

void nmfk_default_params(nmfk_params* p) {
    if (!p) return;
    std::memset(p, 0, sizeof(*p));
    p->tol = 1e-19;
    p->tolOF = 1e-3;
    p->eps_clamp = 2.220446049250313e-16;
    p->weight = 1.0;
    p->maxiter = 10000;
    p->maxbaditers = 10;
    p->maxreattempts = 2;
    p->stopconv = 1000;
    p->check_every = 10;
    p->Wfixed = 0;
    p->Hfixed = 0;
    p->normalize = 1;
    p->iter_limit = 0;
    p->engine = NMFK_ENGINE_AUTO;
}

int32_t nmfk_ctx_create(int32_t device, nmfk_ctx** out) {
    if (!out) return fail(nullptr, NMFK_E_INVALID, "nmfk_ctx_create: out is NULL");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, NMFK_E_NO_DEVICE,
                    std::string("no CUDA device available (") + cudaGetErrorString(e) +
                        "); nmfk_b200 has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(nullptr, NMFK_E_INVALID, "nmfk_ctx_create: bad device index");
    nmfk_ctx* c = new nmfk_ctx();
    c->device = device;
    CU(c, cudaSetDevice(device));
    CU(c, cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CU(c, cudaEventCreate(&c->ev0));
    CU(c, cudaEventCreate(&c->ev1));
    *out = c;
    return NMFK_OK;
}

int32_t nmfk_ctx_destroy(nmfk_ctx* c) {
    if (!c) return NMFK_OK;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    nmfk_ctx_comm_destroy(c);
    if (c->Xp) cudaFree(c->Xp);
    if (c->Xpt) cudaFree(c->Xpt);
    if (c->d_partials) cudaFree(c->d_partials);
    for (auto s : c->pool) cudaStreamDestroy(s);
    for (auto e : c->pool_ev) cudaEventDestroy(e);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return NMFK_OK;
}

int32_t nmfk_ctx_sync(nmfk_ctx* c) {
    if (!c) return fail(nullptr, NMFK_E_INVALID, "ctx is NULL");
    CU(c, cudaSetDevice(c->device));
    CU(c, cudaStreamSynchronize(c->stream));
    for (auto s : c->pool) CU(c, cudaStreamSynchronize(s));
    return NMFK_OK;
}

int32_t nmfk_set_X(nmfk_ctx* c, const void* X, int64_t n, int64_t m, int32_t dtype, double lambda,
                   const void* normalizevector, int32_t on_device) {
    if (!c) return fail(nullptr, NMFK_E_INVALID, "ctx is NULL");
    if (!X) return fail(c, NMFK_E_INVALID, "nmfk_set_X: X is NULL");
    if (dtype != NMFK_F32 && dtype != NMFK_F64) return fail(c, NMFK_E_INVALID, "nmfk_set_X: bad dtype");
    if (n <= 0 || m <= 0) return fail(c, NMFK_E_EMPTY, "Input array has a zero dimension!");
    if (n > INT32_MAX || m > INT32_MAX) return fail(c, NMFK_E_UNSUPPORTED, "nmfk_set_X: dimension exceeds int32");
    if (normalizevector)
        return fail(c, NMFK_E_UNSUPPORTED, "nmfk_set_X: normalizevector is not on the B200 path yet");
    CU(c, cudaSetDevice(c->device));
    c->has_X = false;
    if (c->Xp) cudaFree(c->Xp);
    if (c->Xpt) cudaFree(c->Xpt);
    c->Xp = c->Xpt = nullptr;
    const size_t bytes = (size_t)n * m * esize(dtype);
    CU(c, cudaMalloc(&c->Xp, bytes));
    CU(c, cudaMalloc(&c->Xpt, bytes));
    void* raw = nullptr;
    const void* src = X;
    if (!on_device) {
        CU(c, cudaMalloc(&raw, bytes));
        CU(c, cudaMemcpyAsync(raw, X, bytes, cudaMemcpyHostToDevice, c->stream));
        src = raw;
    }
    PreStats* d_stats = nullptr;
    unsigned char *d_rf = nullptr, *d_cf = nullptr;
    double* d_bm = nullptr;
    const long long nb = ((n + 31) / 32) * ((m + 31) / 32);
    CU(c, cudaMalloc(&d_stats, sizeof(PreStats)));
    CU(c, cudaMalloc(&d_rf, (size_t)n));
    CU(c, cudaMalloc(&d_cf, (size_t)m));
    CU(c, cudaMalloc(&d_bm, (size_t)nb * sizeof(double)));
    cudaError_t e = launch_preprocess(src, c->Xp, c->Xpt, n, m, dtype, lambda, d_stats, d_rf, d_cf, d_bm, (int)nb, c->stream);
    c->launches += 3;
    PreStats hs{};
    std::vector<double> bm((size_t)nb);
    if (e == cudaSuccess) e = cudaMemcpyAsync(&hs, d_stats, sizeof(hs), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(bm.data(), d_bm, (size_t)nb * sizeof(double), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(d_stats);
    cudaFree(d_rf);
    cudaFree(d_cf);
    cudaFree(d_bm);
    if (raw) cudaFree(raw);
    CU(c, e);
    double xmin = std::numeric_limits<double>::infinity();
    for (double v : bm) xmin = std::min(xmin, v);
    c->n = n;
    c->m = m;
    c->dtype = dtype;
    c->lambda = lambda;
    c->info = nmfk_xinfo{};
    c->info.n = n;
    c->info.m = m;
    c->info.nnan = (int64_t)hs.nnan;
    c->info.nzero = (int64_t)hs.nzero;
    c->info.zero_row = hs.zero_rows > 0;
    c->info.zero_col = hs.zero_cols > 0;
    c->info.xmin = xmin;
    c->info.dtype = dtype;
    // `minimum(X) < 0` is false when X holds a NaN (minimum propagates it): NMFkMultiplicative.jl:4
    if (hs.nneg > 0 && hs.nnan == 0) {
        cudaFree(c->Xp);
        cudaFree(c->Xpt);
        c->Xp = c->Xpt = nullptr;
        return fail(c, NMFK_E_NEGATIVE, "All matrix entries must be nonnegative!");
    }
    c->has_X = true;
    return NMFK_OK;
}

int32_t nmfk_get_xinfo(const nmfk_ctx* c, nmfk_xinfo* out) {
    if (!c || !out) return fail(nullptr, NMFK_E_INVALID, "nmfk_get_xinfo: NULL argument");
    if (!c->has_X) return fail(const_cast<nmfk_ctx*>(c), NMFK_E_NO_X, "nmfk_set_X has not been called");
    *out = c->info;
    return NMFK_OK;
}

int32_t nmfk_batch_create(nmfk_ctx* c, int32_t k, int32_t R, nmfk_batch** out) {
    if (!c || !out) return fail(c, NMFK_E_INVALID, "nmfk_batch_create: NULL argument");
    *out = nullptr;
    if (!c->has_X) return fail(c, NMFK_E_NO_X, "nmfk_set_X has not been called");
    if (k < 1 || R < 1) return fail(c, NMFK_E_INVALID, "nmfk_batch_create: k and R must be >= 1");
    CU(c, cudaSetDevice(c->device));
    nmfk_batch* b = new nmfk_batch();
    b->ctx = c;
    b->k = k;
    b->R = R;
    const size_t es = esize(c->dtype);
    cudaError_t e = cudaMalloc(&b->W, (size_t)c->n * k * R * es);
    if (e == cudaSuccess) e = cudaMalloc(&b->H, (size_t)k * c->m * R * es);
    if (e == cudaSuccess) e = cudaMalloc(&b->st, (size_t)R * sizeof(UnitState));
    if (e == cudaSuccess) e = cudaMalloc(&b->canon, (size_t)R * c->m * sizeof(int32_t));
    if (e == cudaSuccess && c->info.nnan > 0) e = cudaMalloc(&b->ximp, (size_t)R * c->n * c->m * es);
    if (e != cudaSuccess) {
        nmfk_batch_destroy(b);
        CU(c, e);
    }
    *out = b;
    return NMFK_OK;
}

int32_t nmfk_batch_create_hstack(nmfk_ctx* c, int32_t k, int32_t R, nmfk_batch** out) {
    if (!c || !out) return fail(c, NMFK_E_INVALID, "nmfk_batch_create_hstack: NULL argument");
    *out = nullptr;
    if (!c->has_X) return fail(c, NMFK_E_NO_X, "nmfk_set_X has not been called");
    if (k < 1 || R < 1) return fail(c, NMFK_E_INVALID, "nmfk_batch_create_hstack: k and R must be >= 1");
    CU(c, cudaSetDevice(c->device));
    nmfk_batch* b = new nmfk_batch();
    b->ctx = c;
    b->k = k;
    b->R = R;
    cudaError_t e = cudaMalloc(&b->H, (size_t)k * c->m * R * esize(c->dtype));
    if (e == cudaSuccess) e = cudaMalloc(&b->st, (size_t)R * sizeof(UnitState));
    if (e != cudaSuccess) {
        nmfk_batch_destroy(b);
        CU(c, e);
    }
    *out = b;
    return NMFK_OK;
}

int32_t nmfk_batch_device_ptrs(nmfk_batch* b, void** W, void** H) {
    if (!b) return fail(nullptr, NMFK_E_INVALID, "nmfk_batch_device_ptrs: NULL batch");
    if (W) *W = b->W;
    if (H) *H = b->H;
    return NMFK_OK;
}

int32_t nmfk_batch_import(nmfk_batch* b, const void* W, const void* H, const double* obj_norm, const int32_t* iters,
                          int32_t on_device) {
    if (!b || !H || !obj_norm) return fail(b ? b->ctx : nullptr, NMFK_E_INVALID, "nmfk_batch_import: NULL argument");
    nmfk_ctx* c = b->ctx;
    CU(c, cudaSetDevice(c->device));
    const size_t es = esize(c->dtype);
    const cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    if (W) {
        if (!b->W) return fail(c, NMFK_E_INVALID, "nmfk_batch_import: this batch has no W stack");
        CU(c, cudaMemcpyAsync(b->W, W, (size_t)c->n * b->k * b->R * es, kind, c->stream));
    }
    CU(c, cudaMemcpyAsync(b->H, H, (size_t)b->k * c->m * b->R * es, kind, c->stream));
    std::vector<UnitState> st((size_t)b->R);
    for (int r = 0; r < b->R; ++r) {
        std::memset(&st[r], 0, sizeof(UnitState));
        st[r].it = iters ? iters[r] : 0;
        st[r].stop = NMFK_STOP_MAXITER;
        st[r].done = 1;
        st[r].best = st[r].obj_chk = st[r].obj_ssq = std::numeric_limits<double>::quiet_NaN();
        st[r].obj_norm = obj_norm[r];
    }
    CU(c, cudaMemcpyAsync(b->st, st.data(), st.size() * sizeof(UnitState), cudaMemcpyHostToDevice, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    b->inited = true;
    return NMFK_OK;
}

int32_t nmfk_batch_destroy(nmfk_batch* b) {
    if (!b) return NMFK_OK;
    if (b->ctx) cudaSetDevice(b->ctx->device);
    if (b->W) cudaFree(b->W);
    if (b->H) cudaFree(b->H);
    if (b->st) cudaFree(b->st);
    if (b->canon) cudaFree(b->canon);
    if (b->ximp) cudaFree(b->ximp);
    delete b;
    return NMFK_OK;
}

static int32_t reset_state(nmfk_batch* b) {
    nmfk_ctx* c = b->ctx;
    std::vector<UnitState> st((size_t)b->R);
    for (auto& s : st) {
        std::memset(&s, 0, sizeof(s));
        s.best = std::numeric_limits<double>::infinity();
        s.obj_chk = std::numeric_limits<double>::quiet_NaN();
        s.obj_ssq = std::numeric_limits<double>::quiet_NaN();
        s.obj_norm = std::numeric_limits<double>::quiet_NaN();
    }
    CU(c, cudaMemcpyAsync(b->st, st.data(), st.size() * sizeof(UnitState), cudaMemcpyHostToDevice, c->stream));
    CU(c, cudaMemsetAsync(b->canon, 0xff, (size_t)b->R * c->m * sizeof(int32_t), c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    b->inited = true;
    return NMFK_OK;
}

template <typename T>
static bool has_nan_host(const T* p, size_t len) {
    for (size_t i = 0; i < len; ++i)
        if (p[i] != p[i]) return true;
    return false;
}

int32_t nmfk_batch_set_init(nmfk_batch* b, const void* Winit, const void* Hinit) {
    if (!b || !Winit || !Hinit) return fail(b ? b->ctx : nullptr, NMFK_E_INVALID, "nmfk_batch_set_init: NULL argument");
    nmfk_ctx* c = b->ctx;
    CU(c, cudaSetDevice(c->device));
    const size_t wl = (size_t)c->n * b->k * b->R, hl = (size_t)b->k * c->m * b->R;
    bool wn, hn;
    if (c->dtype == NMFK_F64) {
        wn = has_nan_host((const double*)Winit, wl);
        hn = has_nan_host((const double*)Hinit, hl);
    } else {
        wn = has_nan_host((const float*)Winit, wl);
        hn = has_nan_host((const float*)Hinit, hl);
    }
    if (wn) return fail(c, NMFK_E_NAN_INIT, "Initial values for the W matrix entries include NaNs!");
    if (hn) return fail(c, NMFK_E_NAN_INIT, "Initial values for the H matrix entries include NaNs!");
    CU(c, cudaMemcpyAsync(b->W, Winit, wl * esize(c->dtype), cudaMemcpyHostToDevice, c->stream));
    CU(c, cudaMemcpyAsync(b->H, Hinit, hl * esize(c->dtype), cudaMemcpyHostToDevice, c->stream));
    return reset_state(b);
}

int32_t nmfk_batch_init_random(nmfk_batch* b, uint64_t seed0) {
    if (!b) return fail(nullptr, NMFK_E_INVALID, "nmfk_batch_init_random: NULL batch");
    nmfk_ctx* c = b->ctx;
    CU(c, cudaSetDevice(c->device));
    if (c->sharded && (c->row0 + c->n > c->n_global))
        return fail(c, NMFK_E_SHAPE, "row-sharded ctx: row0 + local rows exceeds n_global");
    CU(c, launch_philox_init(b->W, b->H, c->sharded ? c->n_global : c->n, c->sharded ? c->row0 : 0, c->n, b->k, c->m, b->R,
                             seed0, c->dtype, c->stream));
    c->launches += 1;
    return reset_state(b);
}

static int32_t check_params(nmfk_ctx* c, const nmfk_params* p) {
    if (!p) return fail(c, NMFK_E_INVALID, "params is NULL");
    if (p->check_every < 1 || p->maxiter < 0 || p->maxbaditers < 1 || p->maxreattempts < 1)
        return fail(c, NMFK_E_INVALID, "params: check_every, maxbaditers, maxreattempts must be >= 1, maxiter >= 0");
    if (p->normalize < 0 || p->normalize > 2) return fail(c, NMFK_E_INVALID, "params: normalize must be 0, 1 or 2");
    return NMFK_OK;
}

static void fill_args(const nmfk_batch* b, const nmfk_params* p, SolveArgs& a) {
    const nmfk_ctx* c = b->ctx;
    a.X = c->Xp;
    a.Xt = c->Xpt;
    a.W = b->W;
    a.H = b->H;
    a.st = b->st;
    a.canon = b->canon;
    a.ximp = b->ximp;
    a.n = (int)c->n;
    a.m = (int)c->m;
    a.k = b->k;
    a.R = b->R;
    a.has_nan = c->info.nnan > 0;
    a.SH = a.SW = 1;
    a.maxiter = p->maxiter;
    a.maxbad = p->maxbaditers;
    a.maxre = p->maxreattempts;
    a.stopconv = p->stopconv;
    a.check_every = p->check_every;
    a.Wfixed = p->Wfixed;
    a.Hfixed = p->Hfixed;
    a.normalize = p->normalize;
    a.iter_limit = p->iter_limit;
    a.lambda = c->lambda;
    a.tol = p->tol;
    a.tolOF = p->tolOF;
    a.eps_clamp = p->eps_clamp;
    a.weight = p->weight;
    a.tiled_tc = p->engine != NMFK_ENGINE_TILED_SCALAR;
    a.shard = c->sharded ? &c->shard : nullptr;
}

// tiled engine (kl_tiled.cu): host-driven, for factors that do not fit in shared memory
namespace nmfk {
cudaError_t solve_tiled(const SolveArgs& a, int dtype, cudaStream_t s, int64_t* launches);
bool tiled_supported(int k);
}  // namespace nmfk

int32_t nmfk_solve(nmfk_ctx* c, nmfk_batch* const* batches, int32_t nb, const nmfk_params* p) {
    if (!c || !batches || nb < 1) return fail(c, NMFK_E_INVALID, "nmfk_solve: bad arguments");
    int32_t rc = check_params(c, p);
    if (rc) return rc;
    CU(c, cudaSetDevice(c->device));
    for (int i = 0; i < nb; ++i) {
        if (!batches[i] || batches[i]->ctx != c) return fail(c, NMFK_E_INVALID, "nmfk_solve: batch of another ctx");
        if (!batches[i]->inited) return fail(c, NMFK_E_INVALID, "nmfk_solve: batch has no initialisation");
        if (!batches[i]->W || !batches[i]->canon) return fail(c, NMFK_E_INVALID, "nmfk_solve: H-only batch");
    }
    while ((int)c->pool.size() < nb) {
        cudaStream_t s;
        cudaEvent_t ev;
        CU(c, cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
        CU(c, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        c->pool.push_back(s);
        c->pool_ev.push_back(ev);
    }
    // largest k first: its restarts are the longest units (LPT order for the block scheduler)
    std::vector<int> ord(nb);
    std::iota(ord.begin(), ord.end(), 0);
    std::stable_sort(ord.begin(), ord.end(), [&](int x, int y) { return batches[x]->k > batches[y]->k; });
    const size_t es = esize(c->dtype);
    CU(c, cudaEventRecord(c->ev0, c->stream));
    // Resident engine: one CTA runs one restart to its stop, and restarts of one sweep stop anywhere
    // between a few hundred and maxiter iterations.  To keep the tail of the sweep short, a solve is
    // enqueued as a sequence of launches that each advance every unfinished restart by at most
    // kResidentChunk iterations (the state is resumable; finished restarts exit at once): SMs freed by
    // short restarts are refilled at chunk granularity and the last wave is at most one chunk long.
    constexpr int kResidentChunk = 500;
    const int user_limit = p->iter_limit > 0 ? std::min(p->iter_limit, p->maxiter) : p->maxiter;
    std::vector<int> mode(nb, 0);  // 1 dmma, 2 scalar resident, 3 tiled
    std::vector<SolveArgs> args(nb);
    for (int q = 0; q < nb; ++q) {
        nmfk_batch* b = batches[ord[q]];
        SolveArgs& a = args[q];
        fill_args(b, p, a);
        CU(c, cudaStreamWaitEvent(c->pool[q], c->ev0, 0));
        const bool want_scalar = (p->engine == NMFK_ENGINE_RESIDENT_SCALAR) || c->dtype != NMFK_F64;
        const bool fits_dmma = !want_scalar && resident_dmma_fits(a.n, a.m, a.k);
        const bool fits_scalar = resident_fits(a.n, a.m, a.k, es);
        bool resident = fits_dmma || fits_scalar;
        if (p->engine == NMFK_ENGINE_TILED || p->engine == NMFK_ENGINE_TILED_SCALAR) resident = false;
        if (c->sharded) {  // rows of X / W live on several GPUs: only the tiled engine exchanges partials
            if (p->engine == NMFK_ENGINE_RESIDENT || p->engine == NMFK_ENGINE_RESIDENT_SCALAR)
                return fail(c, NMFK_E_UNSUPPORTED, "row-sharded ctx: only the tiled engine is available");
            if (p->normalize == 2) return fail(c, NMFK_E_UNSUPPORTED, "row-sharded ctx: clusterWmatrix normalisation is not available");
            resident = false;
        }
        if ((p->engine == NMFK_ENGINE_RESIDENT || p->engine == NMFK_ENGINE_RESIDENT_SCALAR) && !resident)
            return fail(c, NMFK_E_UNSUPPORTED, "resident engine: factors do not fit in shared memory (or k > 32)");
        if (!resident && !tiled_supported(a.k)) return fail(c, NMFK_E_UNSUPPORTED, "tiled engine: k > 32 is not supported");
        mode[q] = resident ? (fits_dmma ? 1 : 2) : 3;
    }
    for (int lim = std::min(kResidentChunk, user_limit);; lim = std::min(lim + kResidentChunk, user_limit)) {
        for (int q = 0; q < nb; ++q) {
            if (mode[q] == 3) continue;
            SolveArgs a = args[q];
            a.iter_limit = lim < p->maxiter ? lim : p->iter_limit;  // the last launch runs to the real stop
            if (mode[q] == 1)
                CU(c, launch_kl_resident_dmma(a, c->pool[q]));
            else
                CU(c, c->dtype == NMFK_F64 ? launch_kl_resident_f64(a, c->pool[q]) : launch_kl_resident_f32(a, c->pool[q]));
            c->launches += 1;
        }
        if (lim >= user_limit) break;
    }
    for (int q = 0; q < nb; ++q) {
        if (mode[q] == 3) CU(c, solve_tiled(args[q], c->dtype, c->pool[q], &c->launches));
        CU(c, cudaEventRecord(c->pool_ev[q], c->pool[q]));
        CU(c, cudaStreamWaitEvent(c->stream, c->pool_ev[q], 0));
    }
    CU(c, cudaEventRecord(c->ev1, c->stream));
    CU(c, cudaEventSynchronize(c->ev1));
    float ms = 0.f;
    CU(c, cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    c->last_solve_ms = ms;
    return NMFK_OK;
}

static int32_t fetch_state(nmfk_batch* b, std::vector<UnitState>& st) {
    nmfk_ctx* c = b->ctx;
    st.resize((size_t)b->R);
    CU(c, cudaMemcpy(st.data(), b->st, st.size() * sizeof(UnitState), cudaMemcpyDeviceToHost));
    return NMFK_OK;
}

int32_t nmfk_batch_get(nmfk_batch* b, void* W_out, void* H_out, double* obj_ssq, double* obj_norm, int32_t* iters,
                       int32_t* stop_reason) {
    if (!b) return fail(nullptr, NMFK_E_INVALID, "nmfk_batch_get: NULL batch");
    nmfk_ctx* c = b->ctx;
    CU(c, cudaSetDevice(c->device));
    const size_t es = esize(c->dtype);
    if (W_out) CU(c, cudaMemcpy(W_out, b->W, (size_t)c->n * b->k * b->R * es, cudaMemcpyDeviceToHost));
    if (H_out) CU(c, cudaMemcpy(H_out, b->H, (size_t)b->k * c->m * b->R * es, cudaMemcpyDeviceToHost));
    if (obj_ssq || obj_norm || iters || stop_reason) {
        std::vector<UnitState> st;
        int32_t rc = fetch_state(b, st);
        if (rc) return rc;
        for (int r = 0; r < b->R; ++r) {
            if (obj_ssq) obj_ssq[r] = st[r].obj_ssq;
            if (obj_norm) obj_norm[r] = st[r].obj_norm;
            if (iters) iters[r] = st[r].it;
            if (stop_reason) stop_reason[r] = st[r].stop;
        }
    }
    return NMFK_OK;
}

// residual sums of one (W,H) device pair; out[0] = weighted ssq, out[1] = plain ssq
static int32_t residual(nmfk_ctx* c, int k, const void* W, const void* H, int restore, double weight, double out[2]) {
    const int nb = residual_blocks((int)c->n);
    if (c->partials_cap < (size_t)nb * 2) {
        if (c->d_partials) cudaFree(c->d_partials);
        c->d_partials = nullptr;
        CU(c, cudaMalloc(&c->d_partials, (size_t)nb * 2 * sizeof(double)));
        c->partials_cap = (size_t)nb * 2;
    }
    CU(c, launch_residual(c->Xp, c->dtype, (int)c->n, (int)c->m, k, W, H, c->lambda, restore, weight, c->d_partials,
                          c->stream));
    c->launches += 1;
    std::vector<double> h((size_t)nb * 2);
    CU(c, cudaMemcpyAsync(h.data(), c->d_partials, h.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    out[0] = out[1] = 0.0;
    for (int i = 0; i < nb; ++i) {
        out[0] += h[2 * i];
        out[1] += h[2 * i + 1];
    }
    return NMFK_OK;
}

int32_t nmfk_batch_objective(nmfk_batch* b, double weight, double* obj_ssq) {
    if (!b || !obj_ssq) return fail(b ? b->ctx : nullptr, NMFK_E_INVALID, "nmfk_batch_objective: NULL argument");
    nmfk_ctx* c = b->ctx;
    CU(c, cudaSetDevice(c->device));
    const size_t es = esize(c->dtype);
    for (int r = 0; r < b->R; ++r) {
        double o[2];
        int32_t rc = residual(c, b->k, (const char*)b->W + (size_t)r * c->n * b->k * es,
                              (const char*)b->H + (size_t)r * b->k * c->m * es, 0, weight, o);
        if (rc) return rc;
        obj_ssq[r] = o[0];
    }
    return NMFK_OK;
}

int32_t nmfk_batch_cluster(nmfk_batch* b, int32_t clusterWmatrix, int32_t* order_out, int32_t* labels_out,
                           double* sil_out, double* clustersil_out, double* robustness_out, void* centroids_out,
                           int32_t* centroid_cols) {
    if (!b) return fail(nullptr, NMFK_E_INVALID, "nmfk_batch_cluster: NULL batch");
    nmfk_ctx* c = b->ctx;
    CU(c, cudaSetDevice(c->device));
    const int k = b->k, R = b->R;
    std::vector<UnitState> st;
    int32_t rc = fetch_state(b, st);
    if (rc) return rc;
    // idxsort = sortperm(objvalue) (NMFkExecute.jl:545): stable, ascending, NaN last; objvalue is a Vector{T}
    std::vector<double> obj((size_t)R);
    for (int r = 0; r < R; ++r) obj[r] = c->dtype == NMFK_F32 ? (double)(float)st[r].obj_norm : st[r].obj_norm;
    std::vector<int32_t> order((size_t)R);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) {
        const double a = obj[x], d = obj[y];
        if (a != a) return false;  // NaN is never less
        if (d != d) return true;
        return a < d;
    });
    if (order_out) std::copy(order.begin(), order.end(), order_out);
    if (k == 1) {  // minsilhouette = 1 when nk == 1 (NMFkExecute.jl:618)
        if (labels_out) std::fill(labels_out, labels_out + R, 1);
        if (sil_out) std::fill(sil_out, sil_out + R, 1.0);
        if (clustersil_out) clustersil_out[0] = 1.0;
        if (robustness_out) *robustness_out = 1.0;
        if (centroid_cols) *centroid_cols = 0;
        return NMFK_OK;
    }
    const size_t es = esize(c->dtype);
    const int len = clusterWmatrix ? (int)c->n : (int)c->m;
    const int N = R * k, ld = len + 1;
    if (clusterWmatrix && !b->W) return fail(c, NMFK_E_INVALID, "nmfk_batch_cluster: clusterWmatrix on an H-only batch");
    // nanaction = :zeroed (:566-580)
    if (b->W) CU(c, launch_zero_nan(b->W, (long long)c->n * k * R, c->dtype, c->stream));
    CU(c, launch_zero_nan(b->H, (long long)k * c->m * R, c->dtype, c->stream));
    c->launches += 2;
    ClusterArgs a{};
    int32_t* d_order = nullptr;
    cudaError_t e = cudaMalloc(&d_order, (size_t)R * sizeof(int32_t));
    if (e == cudaSuccess) e = cudaMalloc(&a.labels, (size_t)N * sizeof(int32_t));
    if (e == cudaSuccess) e = cudaMalloc(&a.cent, (size_t)k * ld * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&a.sil, (size_t)N * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&a.clustersil, (size_t)k * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&a.bias, sizeof(int32_t));
    if (e == cudaSuccess) e = cudaMalloc(&a.V, (size_t)N * ld * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&a.vnorm, (size_t)N * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&a.Dm, (size_t)N * N * sizeof(double));
    std::vector<int32_t> labels((size_t)N);
    std::vector<double> sil((size_t)N), csil((size_t)k), cent((size_t)k * ld);
    int32_t bias = 0;
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_order, order.data(), (size_t)R * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) {
        a.F = clusterWmatrix ? b->W : b->H;
        a.len = len;
        a.k = k;
        a.R = R;
        a.use_W = clusterWmatrix;
        a.order = d_order;
        e = launch_cluster(a, c->dtype, c->stream);
        c->launches += 6;
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(labels.data(), a.labels, labels.size() * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(sil.data(), a.sil, sil.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(csil.data(), a.clustersil, csil.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(cent.data(), a.cent, cent.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(&bias, a.bias, sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(d_order);
    cudaFree(a.labels);
    cudaFree(a.cent);
    cudaFree(a.sil);
    cudaFree(a.clustersil);
    cudaFree(a.bias);
    cudaFree(a.V);
    cudaFree(a.vnorm);
    cudaFree(a.Dm);
    CU(c, e);
    (void)es;
    if (labels_out) std::copy(labels.begin(), labels.end(), labels_out);
    if (sil_out) std::copy(sil.begin(), sil.end(), sil_out);
    if (clustersil_out) std::copy(csil.begin(), csil.end(), clustersil_out);
    if (robustness_out) {  // minimum(clustersilhouettes) (:638); Julia's minimum propagates NaN
        double mn = std::numeric_limits<double>::infinity();
        bool anynan = false;
        for (double v : csil) {
            if (v != v) anynan = true;
            mn = std::min(mn, v);
        }
        *robustness_out = anynan ? std::numeric_limits<double>::quiet_NaN() : mn;
        if (c->dtype == NMFK_F32) *robustness_out = (double)(float)*robustness_out;
    }
    const int cols = bias ? ld : len;
    if (centroid_cols) *centroid_cols = cols;
    if (centroids_out) {  // permutedims(newClusterCenters): k x cols, column-major, element type T
        for (int cc = 0; cc < k; ++cc)
            for (int j = 0; j < cols; ++j) {
                const double v = cent[(size_t)cc * ld + j];
                if (c->dtype == NMFK_F64)
                    ((double*)centroids_out)[(size_t)cc + (size_t)j * k] = v;
                else
                    ((float*)centroids_out)[(size_t)cc + (size_t)j * k] = (float)v;
            }
    }
    return NMFK_OK;
}

int32_t nmfk_fit(nmfk_ctx* c, int32_t k, const void* W, const void* H, double* phi) {
    if (!c || !W || !H || !phi || k < 1) return fail(c, NMFK_E_INVALID, "nmfk_fit: bad arguments");
    if (!c->has_X) return fail(c, NMFK_E_NO_X, "nmfk_set_X has not been called");
    CU(c, cudaSetDevice(c->device));
    const size_t es = esize(c->dtype);
    const size_t wb = (size_t)c->n * k * es, hb = (size_t)k * c->m * es;
    void *dW = nullptr, *dH = nullptr;
    CU(c, cudaMalloc(&dW, wb));
    cudaError_t e = cudaMalloc(&dH, hb);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dW, W, wb, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dH, H, hb, cudaMemcpyHostToDevice, c->stream);
    double o[2] = {0, 0};
    int32_t rc = NMFK_OK;
    if (e == cudaSuccess) rc = residual(c, k, dW, dH, 1, 1.0, o);
    cudaFree(dW);
    if (dH) cudaFree(dH);
    CU(c, e);
    if (rc) return rc;
    *phi = std::sqrt(o[1]);
    if (c->dtype == NMFK_F32) *phi = (double)(float)*phi;
    return NMFK_OK;
}

// Wmean, Hmean, Wvar, Hvar of finalize (NMFkFinalize.jl:68-74): what execute_run returns with best=false (:655-658) and what
// the "-all" result file stores (:650-654).  order / labels as returned by nmfk_batch_cluster (labels k x R, 1-based, column t =
// trial t in sorted order).  Any output may be NULL.
int32_t nmfk_batch_cluster_means(nmfk_batch* b, const int32_t* order, const int32_t* labels, void* Wmean, void* Hmean,
                                 void* Wvar, void* Hvar) {
    if (!b || !order || !labels) return fail(b ? b->ctx : nullptr, NMFK_E_INVALID, "nmfk_batch_cluster_means: NULL argument");
    nmfk_ctx* c = b->ctx;
    CU(c, cudaSetDevice(c->device));
    const int k = b->k, R = b->R;
    const long long n = c->n, m = c->m;
    if ((Wmean || Wvar) && !b->W) return fail(c, NMFK_E_INVALID, "nmfk_batch_cluster_means: W statistics on an H-only batch");
    std::vector<int32_t> amap((size_t)k * R, -1);
    for (int t = 0; t < R; ++t) {
        if (order[t] < 0 || order[t] >= R) return fail(c, NMFK_E_INVALID, "nmfk_batch_cluster_means: order out of range");
        for (int a = 0; a < k; ++a) {
            const int lab = labels[(size_t)t * k + a];
            if (lab >= 1 && lab <= k && amap[(size_t)(lab - 1) * R + t] < 0) amap[(size_t)(lab - 1) * R + t] = a;
        }
    }
    const size_t es = esize(c->dtype);
    int32_t *d_order = nullptr, *d_amap = nullptr;
    unsigned char* d_out = nullptr;
    const size_t wbytes = (size_t)n * k * es, hbytes = (size_t)k * m * es;
    cudaError_t e = cudaMalloc(&d_order, (size_t)R * sizeof(int32_t));
    if (e == cudaSuccess) e = cudaMalloc(&d_amap, amap.size() * sizeof(int32_t));
    if (e == cudaSuccess) e = cudaMalloc(&d_out, 2 * (wbytes + hbytes));
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_order, order, (size_t)R * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_amap, amap.data(), amap.size() * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream);
    unsigned char *dWm = d_out, *dWv = d_out + wbytes, *dHm = d_out + 2 * wbytes, *dHv = dHm + hbytes;
    if (e == cudaSuccess && (Wmean || Wvar)) {
        e = launch_cluster_means(b->W, (int)n, k, R, 1, d_order, d_amap, dWm, dWv, c->dtype, c->stream);
        c->launches += 1;
    }
    if (e == cudaSuccess && (Hmean || Hvar)) {
        e = launch_cluster_means(b->H, (int)m, k, R, 0, d_order, d_amap, dHm, dHv, c->dtype, c->stream);
        c->launches += 1;
    }
    if (e == cudaSuccess && Wmean) e = cudaMemcpyAsync(Wmean, dWm, wbytes, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess && Wvar) e = cudaMemcpyAsync(Wvar, dWv, wbytes, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess && Hmean) e = cudaMemcpyAsync(Hmean, dHm, hbytes, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess && Hvar) e = cudaMemcpyAsync(Hvar, dHv, hbytes, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(d_order);
    cudaFree(d_amap);
    cudaFree(d_out);
    CU(c, e);
    return NMFK_OK;
}

int32_t nmfk_run_batch(nmfk_ctx* c, int32_t k, int32_t R, const void* Winit, const void* Hinit, const nmfk_params* p,
                       void* W_out, void* H_out, double* obj_ssq, double* obj_norm, int32_t* iters,
                       int32_t* stop_reason) {
    nmfk_batch* b = nullptr;
    int32_t rc = nmfk_batch_create(c, k, R, &b);
    if (rc) return rc;
    rc = nmfk_batch_set_init(b, Winit, Hinit);
    if (!rc) rc = nmfk_solve(c, &b, 1, p);
    if (!rc) rc = nmfk_batch_get(b, W_out, H_out, obj_ssq, obj_norm, iters, stop_reason);
    nmfk_batch_destroy(b);
    return rc;
}

int32_t nmfk_trace(nmfk_ctx* c, int32_t k, const void* Winit, const void* Hinit, const nmfk_params* p, int32_t niter,
                   void* W_t, void* H_t, double* obj_t) {
    if (!c || !p || niter < 1) return fail(c, NMFK_E_INVALID, "nmfk_trace: bad arguments");
    nmfk_batch* b = nullptr;
    int32_t rc = nmfk_batch_create(c, k, 1, &b);
    if (rc) return rc;
    rc = nmfk_batch_set_init(b, Winit, Hinit);
    const size_t es = esize(c->dtype);
    const size_t wb = (size_t)c->n * k * es, hb = (size_t)k * c->m * es;
    nmfk_params q = *p;
    for (int t = 1; t <= niter && !rc; ++t) {
        q.iter_limit = t;
        rc = nmfk_solve(c, &b, 1, &q);
        if (rc) break;
        rc = nmfk_batch_get(b, W_t ? (char*)W_t + (size_t)(t - 1) * wb : nullptr,
                            H_t ? (char*)H_t + (size_t)(t - 1) * hb : nullptr, nullptr, nullptr, nullptr, nullptr);
        if (!rc && obj_t) rc = nmfk_batch_objective(b, p->weight, obj_t + (t - 1));
    }
    nmfk_batch_destroy(b);
    return rc;
}

// host copy of one restart's factors
static int32_t fetch_factors(nmfk_batch* b, int r, std::vector<char>& W, std::vector<char>& H) {
    nmfk_ctx* c = b->ctx;
    const size_t es = esize(c->dtype);
    const size_t wb = (size_t)c->n * b->k * es, hb = (size_t)b->k * c->m * es;
    W.resize(wb);
    H.resize(hb);
    CU(c, cudaMemcpy(W.data(), (const char*)b->W + (size_t)r * wb, wb, cudaMemcpyDeviceToHost));
    CU(c, cudaMemcpy(H.data(), (const char*)b->H + (size_t)r * hb, hb, cudaMemcpyDeviceToHost));
    return NMFK_OK;
}

template <typename T>
static void signal_sums(const T* W, const T* H, int64_t n, int k, int64_t m, std::vector<double>& s) {
    // sum(W[:,i:i] * H[i:i,:]) == sum(W[:,i]) * sum(H[i,:]) up to rounding (NMFkPostprocess.jl:151-153)
    s.assign((size_t)k, 0.0);
    for (int a = 0; a < k; ++a) {
        double sw = 0.0, sh = 0.0;
        for (int64_t i = 0; i < n; ++i) sw += (double)W[i + (int64_t)a * n];
        for (int64_t j = 0; j < m; ++j) sh += (double)H[a + j * k];
        s[a] = sw * sh;
    }
}

int32_t nmfk_signalorder(const void* W, const void* H, int64_t n, int32_t k, int64_t m, int32_t dtype,
                         int32_t* order) {
    if (!W || !H || !order || k < 1) return fail(nullptr, NMFK_E_INVALID, "nmfk_signalorder: bad arguments");
    std::vector<double> s;
    if (dtype == NMFK_F64)
        signal_sums((const double*)W, (const double*)H, n, k, m, s);
    else
        signal_sums((const float*)W, (const float*)H, n, k, m, s);
    std::iota(order, order + k, 0);
    // sortperm(signal_sum; rev=true): stable, descending
    std::stable_sort(order, order + k, [&](int x, int y) {
        const double a = s[x], d = s[y];
        if (a != a) return d == d;  // rev=true puts NaN first (isless ordering reversed)
        if (d != d) return false;
        return a > d;
    });
    return NMFK_OK;
}

int32_t nmfk_getk(const int32_t* ks, const double* rob, int32_t nks, double cutoff, int32_t strict) {
    if (!ks || !rob || nks < 1) return -1;
    bool allnan = true;
    for (int i = 0; i < nks; ++i)
        if (rob[i] == rob[i]) allnan = false;
    if (allnan) return 0;  // NMFkPostprocess.jl:11-13
    if (nks == 1) {        // :14-23
        if (strict) return rob[0] > cutoff ? ks[0] : -1;
        return ks[0];
    }
    int kn = -1;
    for (int i = 0; i < nks; ++i)
        if (rob[i] > cutoff) kn = i;  // findlast (:25)
    if (kn >= 0) return ks[kn];
    if (strict) return -1;
    int best = -1;  // findmax with NaN -> -Inf (:30-34): first maximum
    double bv = -std::numeric_limits<double>::infinity();
    for (int i = 0; i < nks; ++i) {
        const double v = (rob[i] != rob[i]) ? -std::numeric_limits<double>::infinity() : rob[i];
        if (best < 0 || v > bv) {
            best = i;
            bv = v;
        }
    }
    return ks[best];
}

// post-solve part of execute_run for one batch (NMFkExecute.jl:545-711 with the default keywords)
static int32_t finish_run(nmfk_batch* b, int32_t clusterW, std::vector<char>& Wb, std::vector<char>& Hb, double* phi,
                          double* robustness, double* aic, int64_t* total_iters) {
    nmfk_ctx* c = b->ctx;
    const int k = b->k, R = b->R;
    std::vector<int32_t> order((size_t)R), labels((size_t)k * R);
    double rob = 1.0;
    int32_t rc = nmfk_batch_cluster(b, clusterW, order.data(), labels.data(), nullptr, nullptr, &rob, nullptr, nullptr);
    if (rc) return rc;
    const int best = order[0];
    std::vector<char> W0, H0;
    rc = fetch_factors(b, best, W0, H0);
    if (rc) return rc;
    const size_t es = esize(c->dtype);
    Wb = W0;
    Hb = H0;
    if (k > 1) {  // Wbest[:, i] = WBig[bestIdx][:, ci[i]]; Hbest[i, :] = HBig[bestIdx][ci[i], :] (:631-635)
        for (int i = 0; i < k; ++i) {
            const int ci = labels[(size_t)i] - 1;
            std::memcpy(Wb.data() + (size_t)i * c->n * es, W0.data() + (size_t)ci * c->n * es, (size_t)c->n * es);
            for (int64_t j = 0; j < c->m; ++j)
                std::memcpy(Hb.data() + ((size_t)i + (size_t)j * k) * es, H0.data() + ((size_t)ci + (size_t)j * k) * es, es);
        }
    }
    // phi_final = normnan(X - Wa*Ha) with NaN residuals zeroed (:664-668): computed on the device
    void *dW = nullptr, *dH = nullptr;
    CU(c, cudaMalloc(&dW, Wb.size()));
    cudaError_t e = cudaMalloc(&dH, Hb.size());
    if (e == cudaSuccess) e = cudaMemcpyAsync(dW, Wb.data(), Wb.size(), cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dH, Hb.data(), Hb.size(), cudaMemcpyHostToDevice, c->stream);
    double o[2] = {0, 0};
    if (e == cudaSuccess) rc = residual(c, k, dW, dH, 1, 1.0, o);
    cudaFree(dW);
    if (dH) cudaFree(dH);
    CU(c, e);
    if (rc) return rc;
    double ph = std::sqrt(o[1]);
    if (c->dtype == NMFK_F32) ph = (double)(float)ph;
    const double nobs = (double)(c->n * c->m - c->info.nnan);     // sum(.!isnan.(X)) (:697)
    const double nparam = (double)(c->n * k) + (double)(k * c->m);  // (:698)
    if (phi) *phi = ph;
    if (robustness) *robustness = rob;
    if (aic) *aic = 2.0 * nparam + nobs * std::log(ph / nobs);  // (:708)
    if (total_iters) {
        std::vector<UnitState> st;
        rc = fetch_state(b, st);
        if (rc) return rc;
        int64_t s = 0;
        for (auto& u : st) s += u.it;
        *total_iters = s;
    }
    return NMFK_OK;
}

int32_t nmfk_execute_run(nmfk_ctx* c, int32_t k, int32_t R, const void* Winit, const void* Hinit, uint64_t seed0,
                         const nmfk_params* p, void* W_best, void* H_best, double* phi, double* robustness, double* aic,
                         int64_t* total_iters) {
    nmfk_batch* b = nullptr;
    int32_t rc = nmfk_batch_create(c, k, R, &b);
    if (rc) return rc;
    if (Winit && Hinit)
        rc = nmfk_batch_set_init(b, Winit, Hinit);
    else
        rc = nmfk_batch_init_random(b, seed0);
    if (!rc) rc = nmfk_solve(c, &b, 1, p);
    std::vector<char> Wb, Hb;
    if (!rc) rc = finish_run(b, p->normalize == 2, Wb, Hb, phi, robustness, aic, total_iters);
    if (!rc) {
        if (W_best) std::memcpy(W_best, Wb.data(), Wb.size());
        if (H_best) std::memcpy(H_best, Hb.data(), Hb.size());
    }
    nmfk_batch_destroy(b);
    return rc;
}

int32_t nmfk_execute(nmfk_ctx* c, const int32_t* ks, int32_t nks, int32_t R, const void* const* Winit,
                     const void* const* Hinit, uint64_t seed0, const nmfk_params* p, double cutoff, void* const* W_out,
                     void* const* H_out, double* fitquality, double* robustness, double* aic, int32_t* kopt,
                     int64_t* total_iters) {
    if (!c || !ks || nks < 1 || R < 1) return fail(c, NMFK_E_INVALID, "nmfk_execute: bad arguments");
    std::vector<nmfk_batch*> bs((size_t)nks, nullptr);
    int32_t rc = NMFK_OK;
    for (int i = 0; i < nks && !rc; ++i) {
        rc = nmfk_batch_create(c, ks[i], R, &bs[i]);
        if (rc) break;
        if (Winit && Hinit && Winit[i] && Hinit[i])
            rc = nmfk_batch_set_init(bs[i], Winit[i], Hinit[i]);
        else
            rc = nmfk_batch_init_random(bs[i], seed0);
    }
    if (!rc) rc = nmfk_solve(c, bs.data(), nks, p);
    int64_t tot = 0;
    std::vector<double> rob((size_t)nks, 0.0), fit((size_t)nks, 0.0);
    const size_t es = esize(c ? c->dtype : NMFK_F64);
    for (int i = 0; i < nks && !rc; ++i) {
        std::vector<char> Wb, Hb;
        double ph = 0, rb = 0, ai = 0;
        int64_t it = 0;
        rc = finish_run(bs[i], p->normalize == 2, Wb, Hb, &ph, &rb, &ai, &it);
        if (rc) break;
        tot += it;
        const int k = ks[i];
        // signal ordering of execute(X, nk) (:311-318) unless Wfixed/Hfixed (:305-307)
        std::vector<int32_t> so((size_t)k);
        std::iota(so.begin(), so.end(), 0);
        if (!p->Wfixed && !p->Hfixed) nmfk_signalorder(Wb.data(), Hb.data(), c->n, k, c->m, c->dtype, so.data());
        std::vector<char> Wo(Wb.size()), Ho(Hb.size());
        for (int a = 0; a < k; ++a) {
            std::memcpy(Wo.data() + (size_t)a * c->n * es, Wb.data() + (size_t)so[a] * c->n * es, (size_t)c->n * es);
            for (int64_t j = 0; j < c->m; ++j)
                std::memcpy(Ho.data() + ((size_t)a + (size_t)j * k) * es, Hb.data() + ((size_t)so[a] + (size_t)j * k) * es, es);
        }
        if (W_out && W_out[i]) std::memcpy(W_out[i], Wo.data(), Wo.size());
        if (H_out && H_out[i]) std::memcpy(H_out[i], Ho.data(), Ho.size());
        rob[i] = rb;
        fit[i] = ph;  // execute re-derives fit = normnan(X - W*H) from the returned factors (:212-222)
        if (fitquality) fitquality[i] = ph;
        if (robustness) robustness[i] = rb;
        if (aic) aic[i] = ai;
    }
    for (auto b : bs) nmfk_batch_destroy(b);
    if (rc) return rc;
    if (kopt) {
        bool allinf = true;
        for (double f : fit)
            if (!std::isinf(f)) allinf = false;
        *kopt = allinf ? 0 : nmfk_getk(ks, rob.data(), nks, cutoff, 1);  // (:206-208, :225)
    }
    if (total_iters) *total_iters = tot;
    return NMFK_OK;
}

int64_t nmfk_launch_count(const nmfk_ctx* c) { return c ? c->launches : 0; }

double nmfk_last_solve_ms(const nmfk_ctx* c) { return c ? c->last_solve_ms : 0.0; }

int32_t nmfk_measure_peak(nmfk_ctx* c, int32_t which, double* value) {
    if (!c || !value) return fail(c, NMFK_E_INVALID, "nmfk_measure_peak: NULL argument");
    CU(c, cudaSetDevice(c->device));
    CU(c, measure_peak(which, value, c->stream));
    return NMFK_OK;
}

// tcgen05 / tensor-memory building blocks of the Float32 tiled engine, checked on the device (tc_selftest.cu)
int32_t nmfk_umma_selftest(nmfk_ctx* c, const float* U, const float* V, int32_t mode, float* Pss, float* Pts, float* ACCa,
                           float* ACCb, int32_t* err) {
    if (!c || !U || !V || !Pss || !Pts || !ACCa || !ACCb || !err) return fail(c, NMFK_E_INVALID, "nmfk_umma_selftest: NULL argument");
    CU(c, cudaSetDevice(c->device));
    int e = 0;
    CU(c, umma_selftest(U, V, mode, Pss, Pts, ACCa, ACCb, &e, c->stream));
    *err = e;
    return NMFK_OK;
}

int32_t nmfk_umma_timing(nmfk_ctx* c, const float* U, const float* V, int32_t reps, int64_t* cycles8, float* acc) {
    if (!c || !U || !V || !cycles8 || !acc) return fail(c, NMFK_E_INVALID, "nmfk_umma_timing: NULL argument");
    CU(c, cudaSetDevice(c->device));
    CU(c, umma_timing(U, V, reps, reinterpret_cast<long long*>(cycles8), acc, c->stream));
    return NMFK_OK;
}

// host-side Philox stream (tests: bit-compatibility with numpy without a GPU)
int32_t nmfk_philox_host(uint64_t seed, int64_t count, double* out) {
    if (!out || count < 0) return NMFK_E_INVALID;
    for (int64_t b = 0; b * 4 < count; ++b) {
        uint64_t o[4];
        philox4x64_10((uint64_t)b + 1, seed, o);
        for (int q = 0; q < 4 && b * 4 + q < count; ++q) out[b * 4 + q] = philox_to_double(o[q]);
    }
    return NMFK_OK;
}

