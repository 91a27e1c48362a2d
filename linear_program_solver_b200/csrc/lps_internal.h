// Internal declarations shared by the translation units of liblps_b200.so (not part of the C-ABI).
#pragma once
#include <cuda_runtime.h>

#include <vector>

#include "../../include/lps_b200.h"
#include "lps_f64.cuh"
#include "lps_f64_blk.cuh"

int lps_fail(int code, const char* fmt, ...);
#define fail lps_fail
#define CU(call)                                                                                      \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess) return fail(LPS_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)
#define TRY(x)                       \
    do {                             \
        int rc_ = (x);               \
        if (rc_ < 0) return rc_;     \
    } while (0)

struct ExactData;   // lps_exact_api.cu

using namespace lps;

constexpr int kWideChunk = 1024;     // doubles per column chunk for large grids
constexpr int kNarrowChunk = 512;    // ... for small ones (more CTAs)
constexpr size_t kL2Bytes = 120ull << 20;

struct lps_handle {
    int mode = 0;                    // 0 = fp64, 1 = exact (multi-modular integers)
    ExactData* x = nullptr;          // exact-mode state
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int m = 0, n = 0;
    size_t ld = 0;
    long long row0 = 0, m_total = 0;
    int rank = 0, nranks = 1;

    double* T = nullptr;
    bool own_T = false;
    double* ucol = nullptr;
    unsigned char* cand_send = nullptr;
    unsigned char* cand_all = nullptr;
    bool own_cand = false;
    size_t cand_stride = 0;
    // peer-store exchange (lps_f64.cuh XchgBlock): this rank's block, every rank's block (device array), IPC mappings
    XchgBlock* xchg = nullptr;
    XchgBlock** d_peers = nullptr;
    std::vector<void*> ipc_open;
    bool use_xchg = false;   // set by lps_p2p_attach with more than one rank

    int chunk = kWideChunk, nchunks = 0, rows_per_cta = 64, sel_blocks = 1, sel_threads = 256;
    bool stream_hint = false;
    double* pd_val = nullptr;
    int* pd_idx = nullptr;
    int* pb_idx = nullptr;
    ArgMinF64* sel_blk = nullptr;
    SolveState* st = nullptr;
    SolveState* st_host = nullptr;   // pinned mirror
    int* basis = nullptr;            // m_total ints
    int* marks = nullptr;            // n ints
    bool use_marks = false;
    int* trace = nullptr;
    long long trace_cap = 0;
    int* scratch_i = nullptr;        // n + 8 ints (isCanonical)
    void* slab = nullptr;            // the one allocation the small device buffers are carved from
    bool solving = false;

    // pivots in blocks (lps_f64_blk.cuh): the look-ahead panel and the fused update
    int blk_k = kBlkMax;             // pivots decided per panel launch (1..kBlkMax); 0: rank-1 path only
    int sms = 148;                   // SMs of the device: size of the persistent grids
    int pan_blocks = 1, pan_threads = 256;
    int pan_cluster = 0;             // > 0: the panel is ONE cluster of that many CTAs, state in distributed shared memory
    int pan_rpc = 0, pan_cpc = 0;    // rows / columns per CTA of the cluster panel
    BlkPanel blk{};
    bool blk_valid = false;          // O / B / pricing partials of the panel describe the grid
    bool pp_valid = false;           // pd_val / pd_idx / pb_idx (rank-1 path) describe the grid's objective row
    size_t xchg_bytes = 0;

    // a burst of (select, update) pairs captured once as a CUDA graph: small tableaux are launch bound on the host
    cudaGraphExec_t gexec = nullptr;
    unsigned long long gsig = 0;     // the launch arguments the graph was captured with
    long long gburst = 0;

    // profiling of the update kernel
    bool prof = false;
    std::vector<cudaEvent_t> ev;
    size_t ev_used = 0;
    long long launches_update = 0, launches_all = 0;
};


// scratch device buffer of one call: freed on every return path
struct DevBuf {
    void* p = nullptr;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 1); }
    template <typename T> T* as() const { return static_cast<T*>(p); }
};

// One device allocation carved into a handle's many small buffers: cudaMalloc / cudaFree cost a driver round trip
// each (and a device synchronisation on free), which dominated the life of a small tableau (50 x 100: 7.6 ms of a
// 13.6 ms solve).  add() records a pointer and its size, commit() allocates once and assigns the pointers.
struct Slab {
    struct Item { void** p; size_t off; };
    std::vector<Item> items;
    size_t total = 0;
    template <typename T> void add(T** p, size_t count) {
        items.push_back(Item{reinterpret_cast<void**>(p), total});
        total += (count * sizeof(T) + 255) / 256 * 256;
    }
    cudaError_t commit(void** base) {
        cudaError_t e = cudaMalloc(base, total ? total : 256);
        if (e != cudaSuccess) return e;
        for (const Item& it : items) *it.p = static_cast<char*>(*base) + it.off;
        return cudaSuccess;
    }
};
// Small pinned host blocks (the state mirrors): recycled process-wide instead of cudaMallocHost / cudaFreeHost per handle.
void* lps_pinned_get(size_t bytes);          // bytes <= 256; NULL when the pinned allocation fails
void lps_pinned_put(void* p);

int lps_use(const lps_handle* h);
#define USE(h)                         \
    do {                               \
        int rc_ = lps_use(h);          \
        if (rc_ != LPS_OK) return rc_; \
    } while (0)
#define NEED_MODE(h, want)                                                                         \
    do {                                                                                           \
        if ((h)->mode != (want)) return fail(LPS_E_STATE, "%s: wrong arithmetic mode for this handle", __func__); \
    } while (0)

// exact-mode hooks used by the mode-independent entry points (lps_exact_api.cu)
int lpsx_destroy(lps_handle* h);
int lpsx_pivot(lps_handle* h, int r, int c);
int lpsx_find_pivot(lps_handle* h, int rule, int* r, int* c);
int lpsx_is_canonical(lps_handle* h, int* bcols, int* canonical);
int lpsx_is_infeasible(lps_handle* h, int* infeasible);
int lpsx_solve(lps_handle* h, long long max_pivots, int* basis, int* marks, int* trace, long long trace_cap, long long* npiv);
