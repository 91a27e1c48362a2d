// Batch half of the C-ABI (include/lps_b200.h, "batched node LPs"): sizes the cluster / shared-memory
// plan, precomputes the prime tables and launches k_bnb_nodes.  No arithmetic of the path runs here.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "lps_internal.h"
#include "lps_batch.cuh"

struct lps_batch {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int count = 0, m = 0, n = 0, art_cap = 0, P = 0, C = 1, mloc = 0, cw = 0, tb = 0, ld = 0, trace_cap = 0;
    size_t smem = 0;
    int ctas_per_sm = 0;
    PrimeQ* pr = nullptr;
    uint32_t *invtab = nullptr, *half = nullptr;
    double* wmant = nullptr;
    int* wexp = nullptr;
    int32_t* grids = nullptr;
    int *status = nullptr, *npiv = nullptr, *basis = nullptr, *trace = nullptr;
    signed char* sign = nullptr;
    uint32_t* digits = nullptr;
    long long* timing = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool timed = false;
    long long launches = 0;
    long long stuck_limit = -1;
    int threads = 256;
    int wcheck = 0;
};

static const size_t kBatchSmemMax = 227 * 1024 - 64;       // opt-in maximum (the kernel has no static shared memory)

static int batch_ld(int m, int n, int art_cap) { return (n + art_cap + 1 + 3) / 4 * 4; }
static int batch_mloc(int m, int C) { return (m + C - 1) / C; }
static int batch_cw(int n, int art_cap, int C) { return std::max(1, (n + art_cap + C - 1) / C); }

static size_t batch_smem(int m, int n, int art_cap, int P, int C) {
    const int mloc = batch_mloc(m, C);
    return batch_ctx_bytes() + batch_layout(m, P, C, mloc, batch_cw(n, art_cap, C), batch_ld(m, n, art_cap), std::max(mloc, C)).total;
}

// Cluster width.  A pivot is a chain of short latency-bound phases (reconstruct, decide, exchange) around one
// update, so what counts is how many LPs are resident (shared memory decides: one LP needs P (m + C) ld words) and
// how many threads walk each LP's rows.  Measured on B200 (tools/bench_batch.py, 64 x 128 LPs, 12 primes): the
// narrowest cluster that fits wins (C = 2, 74 resident LPs, 512 threads per CTA: 117 k LPs/s; C = 4: 61 k; C = 8: 48 k).
static int batch_pick_cluster(int m, int n, int art_cap, int P) {
    for (int c = 1; c <= kBMaxCluster; c *= 2) {
        if (c > 1 && batch_mloc(m, c) * (c - 1) >= m) break;   // a CTA without rows
        if (batch_smem(m, n, art_cap, P, c) <= kBatchSmemMax) return c;
    }
    return 0;
}

/* Shared memory one CTA needs for this shape with `cluster` CTAs per LP (0: the automatic choice);
 * 0 is returned when the shape does not fit. */
extern "C" size_t lps_batch_smem_bytes(int m, int n, int art_cap, int nprimes, int cluster) {
    if (m <= 0 || n <= 0 || art_cap < 0 || nprimes <= 0) return 0;
    if (cluster == 0) cluster = batch_pick_cluster(m, n, art_cap, nprimes);
    if (cluster <= 0) return 0;
    const size_t s = batch_smem(m, n, art_cap, nprimes, cluster);
    return s <= kBatchSmemMax ? s : 0;
}

static void batch_free(lps_batch* b) {
    if (!b) return;
    cudaSetDevice(b->device);
    if (b->stream) cudaStreamSynchronize(b->stream);
    cudaFree(b->pr); cudaFree(b->invtab); cudaFree(b->half); cudaFree(b->wmant); cudaFree(b->wexp);
    cudaFree(b->grids); cudaFree(b->status); cudaFree(b->npiv); cudaFree(b->basis); cudaFree(b->trace);
    cudaFree(b->sign); cudaFree(b->digits); cudaFree(b->timing);
    if (b->ev0) cudaEventDestroy(b->ev0);
    if (b->ev1) cudaEventDestroy(b->ev1);
    if (b->own_stream && b->stream) cudaStreamDestroy(b->stream);
    delete b;
}

#define CUB(call)                                                                                   \
    do {                                                                                            \
        cudaError_t e_ = (call);                                                                    \
        if (e_ != cudaSuccess) {                                                                    \
            batch_free(b);                                                                          \
            return fail(e_ == cudaErrorMemoryAllocation ? LPS_E_NOMEM : LPS_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
        }                                                                                           \
    } while (0)

// one instantiation per (even) number of primes: the Garner loops are fully unrolled over P
#define BATCH_DISPATCH(P_, EXPR)                     \
    switch (P_) {                                    \
        case 2: { constexpr int PP = 2; EXPR; } break;   \
        case 4: { constexpr int PP = 4; EXPR; } break;   \
        case 6: { constexpr int PP = 6; EXPR; } break;   \
        case 8: { constexpr int PP = 8; EXPR; } break;   \
        case 10: { constexpr int PP = 10; EXPR; } break; \
        case 12: { constexpr int PP = 12; EXPR; } break; \
        case 14: { constexpr int PP = 14; EXPR; } break; \
        default: { constexpr int PP = 16; EXPR; } break; \
    }

extern "C" int lps_batch_create(int count, int m, int n, int art_cap, int nprimes, const uint32_t* primes, int cluster,
                                int trace_cap, int device, void* stream, lps_batch** out) {
    if (!out) return fail(LPS_E_INVALID, "null out");
    *out = nullptr;
    if (count <= 0 || m <= 0 || n <= 0 || art_cap < 0 || art_cap > m) return fail(LPS_E_INVALID, "bad batch shape (count %d, m %d, n %d, artificial columns %d)", count, m, n, art_cap);
    if (nprimes < 2 || nprimes > kBMaxPrimes || (nprimes & 1) || !primes) return fail(LPS_E_INVALID, "the batched kernel takes an even number of primes, 2..%d (got %d): solve larger LPs one at a time", kBMaxPrimes, nprimes);
    for (int k = 0; k < nprimes; ++k) {
        if (!(primes[k] & 1u) || primes[k] < 3u || primes[k] >= (1u << 31)) return fail(LPS_E_INVALID, "prime %u is not an odd number below 2^31", primes[k]);
        if (k > 0 && primes[k] <= primes[k - 1]) return fail(LPS_E_INVALID, "the batched kernel takes its primes in ascending order");
    }
    if (cluster != 0 && cluster != 1 && cluster != 2 && cluster != 4 && cluster != 8) return fail(LPS_E_INVALID, "cluster must be 0 (auto), 1, 2, 4 or 8");
    const int C = cluster ? cluster : batch_pick_cluster(m, n, art_cap, nprimes);
    if (C <= 0 || batch_smem(m, n, art_cap, nprimes, C) > kBatchSmemMax)
        return fail(LPS_E_INVALID, "a %d x %d LP with %d primes does not fit the shared memory of %d CTAs", m, n, nprimes, cluster ? cluster : kBMaxCluster);
    if (trace_cap < 0) trace_cap = 0;
    int ndev = 0;
    CU(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(LPS_E_INVALID, "device %d out of range (%d visible)", device, ndev);
    CU(cudaSetDevice(device));
    lps_batch* b = new (std::nothrow) lps_batch();
    if (!b) return fail(LPS_E_NOMEM, "out of host memory");
    b->device = device;
    if (stream) b->stream = (cudaStream_t)stream;
    else { CUB(cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking)); b->own_stream = true; }
    const int P = nprimes;
    b->count = count; b->m = m; b->n = n; b->art_cap = art_cap; b->P = P; b->C = C;
    b->mloc = batch_mloc(m, C); b->cw = batch_cw(n, art_cap, C); b->tb = std::max(b->mloc, C);
    b->ld = batch_ld(m, n, art_cap); b->trace_cap = trace_cap;
    b->smem = batch_smem(m, n, art_cap, P, C);
    // one CTA per SM by shared memory: 16 warps; several CTAs per SM: 8 warps each (128 registers per thread)
    b->threads = (b->smem + 1024 > (227 * 1024) / 2) ? 512 : 256;
    if (const char* e = getenv("LPS_BATCH_THREADS")) { const int t = atoi(e); if (t >= 64 && t <= 512 && t % 32 == 0) b->threads = t; }

    std::vector<PrimeQ> pr(P);
    for (int k = 0; k < P; ++k) pr[k] = make_prime(primes[k]);
    std::vector<uint32_t> half(P);
    {
        uint64_t carry = 0;
        for (int k = P - 1; k >= 0; --k) {
            const uint64_t cur = carry * primes[k] + (primes[k] - 1u);
            half[k] = (uint32_t)(cur / 2);
            carry = cur & 1u;
        }
    }
    std::vector<double> wm(P + 1);
    std::vector<int> we(P + 1);
    {
        long double mant = 0.5L;
        int ex = 1;
        wm[0] = (double)mant; we[0] = ex;
        for (int t = 1; t <= P; ++t) {
            int e2;
            mant = frexpl(mant * (long double)primes[t - 1], &e2);
            ex += e2;
            wm[t] = (double)mant; we[t] = ex;
        }
    }
    CUB(cudaMalloc(&b->pr, P * sizeof(PrimeQ)));
    CUB(cudaMemcpyAsync(b->pr, pr.data(), P * sizeof(PrimeQ), cudaMemcpyHostToDevice, b->stream));
    CUB(cudaMalloc(&b->half, P * sizeof(uint32_t)));
    CUB(cudaMemcpyAsync(b->half, half.data(), P * sizeof(uint32_t), cudaMemcpyHostToDevice, b->stream));
    CUB(cudaMalloc(&b->wmant, (P + 1) * sizeof(double)));
    CUB(cudaMemcpyAsync(b->wmant, wm.data(), (P + 1) * sizeof(double), cudaMemcpyHostToDevice, b->stream));
    CUB(cudaMalloc(&b->wexp, (P + 1) * sizeof(int)));
    CUB(cudaMemcpyAsync(b->wexp, we.data(), (P + 1) * sizeof(int), cudaMemcpyHostToDevice, b->stream));
    CUB(cudaStreamSynchronize(b->stream));
    CUB(cudaMalloc(&b->invtab, (size_t)P * P * sizeof(uint32_t)));
    k_xinvtab<<<dim3((P + 127) / 128, P), 128, 0, b->stream>>>(b->pr, P, b->invtab);
    CUB(cudaGetLastError());

    const size_t cells = (size_t)(m + 1) * (n + 1);
    CUB(cudaMalloc(&b->grids, (size_t)count * cells * sizeof(int32_t)));
    CUB(cudaMalloc(&b->status, (size_t)count * sizeof(int)));
    CUB(cudaMalloc(&b->npiv, (size_t)count * 4 * sizeof(int)));
    CUB(cudaMalloc(&b->basis, (size_t)count * m * sizeof(int)));
    CUB(cudaMalloc(&b->sign, (size_t)count * (m + 2)));
    CUB(cudaMalloc(&b->digits, (size_t)count * P * (m + 2) * sizeof(uint32_t)));
    if (trace_cap > 0) CUB(cudaMalloc(&b->trace, (size_t)count * trace_cap * 3 * sizeof(int)));
    CUB(cudaEventCreate(&b->ev0));
    CUB(cudaEventCreate(&b->ev1));

    cudaError_t e = cudaSuccess;
    // The attribute belongs to the kernel instantiation, not to this batch: batches of the same prime count but other
    // shapes live next to each other (the host pools them), so it is raised to the most any of them can ask for --
    // setting it to THIS batch's size let a later, smaller batch lower it under an earlier one ("invalid argument"
    // at that one's next launch).
    BATCH_DISPATCH(P, e = cudaFuncSetAttribute(k_bnb_nodes<PP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBatchSmemMax));
    CUB(e);
    CUB(cudaStreamSynchronize(b->stream));
    *out = b;
    return LPS_OK;
}

/* Sizing of a batch on the device (bookkeeping only, see k_batch_plan): the largest number of artificial columns
 * any LP needs and the bits of a Hadamard bound that covers every LP, from host grids [count][(m+1)][(n+1)]. */
extern "C" int lps_batch_plan(int count, int m, int n, const int32_t* grids, int device, int* art_cap, double* bound_bits) {
    if (count <= 0 || m <= 0 || n <= 0 || !grids || !art_cap || !bound_bits) return fail(LPS_E_INVALID, "bad arguments");
    int ndev = 0;
    CU(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(LPS_E_INVALID, "device %d out of range (%d visible)", device, ndev);
    CU(cudaSetDevice(device));
    const size_t bytes = (size_t)count * (m + 1) * (n + 1) * sizeof(int32_t);
    const size_t smem = (size_t)(2 * m + 2) * sizeof(int) + (size_t)(n + m + 4) * sizeof(double);
    if (smem > 200 * 1024) return fail(LPS_E_INVALID, "LP too large for the batch planner");
    int32_t* d_g = nullptr;
    int* d_k = nullptr;
    double* d_b = nullptr;
    cudaError_t e = cudaMalloc(&d_g, bytes);
    if (e == cudaSuccess) e = cudaMalloc(&d_k, (size_t)count * sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc(&d_b, (size_t)count * sizeof(double));
    std::vector<int> hk(count);
    std::vector<double> hb(count);
    if (e == cudaSuccess) e = cudaMemcpy(d_g, grids, bytes, cudaMemcpyDefault);   // host or device memory (UVA)
    if (e == cudaSuccess && smem > 48 * 1024) e = cudaFuncSetAttribute(k_batch_plan, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) {
        k_batch_plan<<<count, 256, smem>>>(d_g, m, n, d_k, d_b);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(hk.data(), d_k, (size_t)count * sizeof(int), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(hb.data(), d_b, (size_t)count * sizeof(double), cudaMemcpyDeviceToHost);
    cudaFree(d_g); cudaFree(d_k); cudaFree(d_b);
    if (e != cudaSuccess) return fail(LPS_E_CUDA, "batch planning failed: %s", cudaGetErrorString(e));
    *art_cap = *std::max_element(hk.begin(), hk.end());
    *bound_bits = *std::max_element(hb.begin(), hb.end());
    return LPS_OK;
}

extern "C" int lps_batch_destroy(lps_batch* b) {
    batch_free(b);
    return LPS_OK;
}

#define USEB(b)                                              \
    do {                                                     \
        if (!(b)) return fail(LPS_E_INVALID, "null batch");  \
        CU(cudaSetDevice((b)->device));                      \
    } while (0)

extern "C" void* lps_batch_grids_ptr(lps_batch* b) { return b ? (void*)b->grids : nullptr; }

extern "C" int lps_batch_upload(lps_batch* b, const int32_t* grids) {
    USEB(b);
    if (!grids) return fail(LPS_E_INVALID, "null grids");
    const size_t bytes = (size_t)b->count * (b->m + 1) * (b->n + 1) * sizeof(int32_t);
    CU(cudaMemcpyAsync(b->grids, grids, bytes, cudaMemcpyDefault, b->stream));    // host or device memory (UVA)
    return LPS_OK;
}

extern "C" int lps_batch_run(lps_batch* b, long long max_pivots) {
    USEB(b);
    BatchParams a;
    a.count = b->count; a.m = b->m; a.n = b->n; a.art_cap = b->art_cap; a.C = b->C; a.mloc = b->mloc; a.cw = b->cw;
    a.ld = b->ld; a.tb = b->tb; a.max_pivots = max_pivots; a.stuck_limit = b->stuck_limit;
    a.pr = b->pr; a.invtab = b->invtab; a.half = b->half; a.wmant = b->wmant; a.wexp = b->wexp;
    a.grids = b->grids; a.status = b->status; a.npiv = b->npiv; a.basis = b->basis; a.sign = b->sign; a.digits = b->digits;
    a.trace = b->trace; a.trace_cap = b->trace_cap; a.timing = b->timing; a.wcheck = b->wcheck;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)b->count * b->C, 1, 1);
    cfg.blockDim = dim3(b->threads, 1, 1);
    cfg.dynamicSmemBytes = b->smem;
    cfg.stream = b->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = b->C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    CU(cudaEventRecord(b->ev0, b->stream));
    cudaError_t e = cudaSuccess;
    BATCH_DISPATCH(b->P, e = cudaLaunchKernelEx(&cfg, k_bnb_nodes<PP>, a));
    if (e != cudaSuccess) return fail(LPS_E_CUDA, "k_bnb_nodes launch failed: %s", cudaGetErrorString(e));
    CU(cudaEventRecord(b->ev1, b->stream));
    b->timed = true;
    b->launches++;
    return LPS_OK;
}

/* Results of the last run (host arrays, any may be NULL):
 *   status [count]; npiv [count][4] = phase-1, drive-out, phase-2 pivots, artificial columns; basis [count][m];
 *   sign [count][m+2] and digits [count][P][m+2]: b_0..b_{m-1}, the z cell, D as sign + mixed-radix digits;
 *   trace [count][trace_cap][3]. */
extern "C" int lps_batch_download(lps_batch* b, int* status, int* npiv, int* basis, signed char* sign, uint32_t* digits, int* trace) {
    USEB(b);
    const size_t c = (size_t)b->count;
    if (status) CU(cudaMemcpyAsync(status, b->status, c * sizeof(int), cudaMemcpyDeviceToHost, b->stream));
    if (npiv) CU(cudaMemcpyAsync(npiv, b->npiv, c * 4 * sizeof(int), cudaMemcpyDeviceToHost, b->stream));
    if (basis) CU(cudaMemcpyAsync(basis, b->basis, c * b->m * sizeof(int), cudaMemcpyDeviceToHost, b->stream));
    if (sign) CU(cudaMemcpyAsync(sign, b->sign, c * (b->m + 2), cudaMemcpyDeviceToHost, b->stream));
    if (digits) CU(cudaMemcpyAsync(digits, b->digits, c * b->P * (b->m + 2) * sizeof(uint32_t), cudaMemcpyDeviceToHost, b->stream));
    if (trace && b->trace) CU(cudaMemcpyAsync(trace, b->trace, c * b->trace_cap * 3 * sizeof(int), cudaMemcpyDeviceToHost, b->stream));
    CU(cudaStreamSynchronize(b->stream));
    return LPS_OK;
}

/* Degenerate-pivot count after which Simplex.solve switches from Dantzig to Bland pricing; negative (the
 * default) is the reference's m + n (simplex.py:123).  A lower limit lets tests reach the Bland loop. */
extern "C" int lps_batch_set_capacity_check(lps_batch* b, int on) {
    USEB(b);
    b->wcheck = on ? 1 : 0;
    return LPS_OK;
}
extern "C" int lps_batch_set_stuck_limit(lps_batch* b, long long limit) {
    USEB(b);
    b->stuck_limit = limit;
    return LPS_OK;
}

/* Profiling aid: per-LP cycle counts of the kernel's phases (CTA 0 of each cluster, 16 slots, see b_tick in
 * lps_batch.cuh).  out == NULL switches the collection on for the following runs; otherwise the counts of the
 * last run are copied to out[count][16]. */
extern "C" int lps_batch_timing(lps_batch* b, long long* out) {
    USEB(b);
    if (!b->timing) {
        CU(cudaMalloc(&b->timing, (size_t)b->count * 16 * sizeof(long long)));
        CU(cudaMemsetAsync(b->timing, 0, (size_t)b->count * 16 * sizeof(long long), b->stream));
    }
    if (out) {
        CU(cudaMemcpyAsync(out, b->timing, (size_t)b->count * 16 * sizeof(long long), cudaMemcpyDeviceToHost, b->stream));
        CU(cudaStreamSynchronize(b->stream));
    }
    return LPS_OK;
}

extern "C" int lps_batch_sync(lps_batch* b) {
    USEB(b);
    CU(cudaStreamSynchronize(b->stream));
    return LPS_OK;
}

/* Plan of this batch: CTAs per cluster, primes per CTA, dynamic shared memory per CTA, resident clusters
 * per device (occupancy query), CUDA-event time of the last run in ms (after a sync), launches so far. */
extern "C" int lps_batch_info(lps_batch* b, int* cluster, int* rows_per_cta, size_t* smem_bytes, int* resident_clusters,
                              double* last_run_ms, long long* launches) {
    USEB(b);
    if (cluster) *cluster = b->C;
    if (rows_per_cta) *rows_per_cta = b->mloc;
    if (smem_bytes) *smem_bytes = b->smem;
    if (resident_clusters) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)b->count * b->C, 1, 1);
        cfg.blockDim = dim3(b->threads, 1, 1);
        cfg.dynamicSmemBytes = b->smem;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = b->C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        int nc = 0;
        cudaError_t e = cudaSuccess;
        BATCH_DISPATCH(b->P, e = cudaOccupancyMaxActiveClusters(&nc, k_bnb_nodes<PP>, &cfg));
        if (e != cudaSuccess) { cudaGetLastError(); nc = -1; }
        *resident_clusters = nc;
    }
    if (last_run_ms) {
        *last_run_ms = 0.0;
        if (b->timed) {
            CU(cudaEventSynchronize(b->ev1));
            float ms = 0.f;
            CU(cudaEventElapsedTime(&ms, b->ev0, b->ev1));
            *last_run_ms = ms;
        }
    }
    if (launches) *launches = b->launches;
    return LPS_OK;
}
