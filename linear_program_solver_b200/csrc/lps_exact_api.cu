// Exact-mode half of the C-ABI (include/lps_b200.h): host-side orchestration of the kernels in
// lps_exact.cuh.  No arithmetic of the path runs here: the host only sizes buffers, precomputes the
// prime-dependent tables and enqueues launches.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <new>
#include <vector>

#include "lps_internal.h"
#include "lps_exact.cuh"

struct ExactData {
    int P = 0, spt = 1, lane_threads = 1024;   // P: primes of the working set (all kernels run on these)
    int Pcap = 0;                               // primes of the handle (sized by the caller's static bound): tables cover all
    int Palloc = 0;                             // residue planes actually allocated (grown with the working set)
    std::vector<uint32_t> primes_h;             // host copies for the working-set arithmetic
    std::vector<int> wexp_h;
    int* wcol = nullptr;                        // device [n+1] column weights (NULL until lps_exact_set_weights)
    int* xbasis = nullptr;                      // device [m] last pivot column per row since the upload
    bool adaptive = false;
    bool static_ok = true;                      // the numbers are still minors of the uploaded matrix (no row operation since)
    long long wmax256 = 0, guard256 = 0;
    int grows = 0;
    PrimeQ* pr = nullptr;
    uint32_t* invtab = nullptr;
    uint32_t* half = nullptr;
    double* wmant = nullptr;
    int* wexp = nullptr;
    uint32_t* R = nullptr;
    size_t ld = 0, plane = 0;
    uint32_t *Dres = nullptr, *dinv_m = nullptr, *alpha_m = nullptr;
    uint32_t* nbeta = nullptr;
    size_t nbeta_ld = 0;
    uint32_t *sigma = nullptr, *one = nullptr;
    size_t sstride = 0;
    int* alive = nullptr;
    signed char *obj_sign = nullptr, *cr_sign = nullptr, *rd_sign = nullptr;
    int *obj_top = nullptr, *cr_top = nullptr, *rd_top = nullptr;
    double *obj_mant = nullptr, *cr_mant = nullptr, *rd_mant = nullptr;
    uint32_t *obj_dig = nullptr, *rd_dig = nullptr;
    size_t obj_dstride = 0, rd_stride = 0;
    int* cand = nullptr;
    XState* st = nullptr;
    XState* st_host = nullptr;
    int chunk = 2048, rows_per_cta = 64;
    bool stream_hint = false;
    bool resume = false;             // next lps_solve continues the loop state set by lps_exact_loop_state
    long long graph_launches = 6;    // kernel launches per pivot inside the captured burst graph
};

static XGrid xgrid_of(const lps_handle* h) {
    const ExactData* x = h->x;
    XGrid g;
    g.R = x->R; g.ld = x->ld; g.plane = x->plane; g.m = h->m; g.n = h->n; g.P = x->P; g.Pld = x->Pcap;
    g.pr = x->pr; g.invtab = x->invtab; g.half = x->half; g.wmant = x->wmant; g.wexp = x->wexp;
    return g;
}
static XPivotConst pconst_of(const lps_handle* h) { return XPivotConst{h->x->Dres, h->x->dinv_m, h->x->alpha_m}; }

extern "C" size_t lps_exact_ld(int n) { return ((size_t)n + 1 + 31) / 32 * 32; }

static PrimeQ make_prime(uint32_t q) {
    PrimeQ p;
    p.q = q;
    uint32_t inv = q;                                   // Newton: q * inv = 1 mod 2^32
    for (int i = 0; i < 5; ++i) inv *= 2u - q * inv;
    p.qinv = 0u - inv;
    p.r1 = (uint32_t)((1ull << 32) % q);
    p.r2 = (uint32_t)(((unsigned __int128)1 << 64) % q);
    return p;
}

#include <chrono>
#include <cstdio>
#include <cstdlib>
static double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

#define CUX(call)                                                                                   \
    do {                                                                                            \
        cudaError_t e_ = (call);                                                                    \
        if (e_ != cudaSuccess) {                                                                    \
            lpsx_destroy(h);                                                                        \
            return fail(e_ == cudaErrorMemoryAllocation ? LPS_E_NOMEM : LPS_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
        }                                                                                           \
    } while (0)

extern "C" int lps_create_exact(int m, int n, int nprimes, const uint32_t* primes, int device, void* stream, lps_handle** out) {
    if (!out) return fail(LPS_E_INVALID, "null out");
    *out = nullptr;
    if (m <= 0) return fail(LPS_E_INVALID, "need m > 0, provided m = %d", m);
    if (n <= 0) return fail(LPS_E_INVALID, "need n > 0, provided n = %d", n);
    if (nprimes <= 0 || nprimes > 1024 * kMaxSPT || !primes) return fail(LPS_E_INVALID, "need 1..%d primes", 1024 * kMaxSPT);
    for (int k = 0; k < nprimes; ++k) {
        if (!(primes[k] & 1u) || primes[k] < 3u || primes[k] >= (1u << 31)) return fail(LPS_E_INVALID, "prime %u is not an odd number below 2^31", primes[k]);
        for (int j = 0; j < k; ++j)
            if (primes[j] == primes[k]) return fail(LPS_E_INVALID, "prime %u listed twice", primes[k]);
    }
    int ndev = 0;
    CU(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(LPS_E_INVALID, "device %d out of range (%d visible)", device, ndev);
    CU(cudaSetDevice(device));
    lps_handle* h = new (std::nothrow) lps_handle();
    ExactData* x = new (std::nothrow) ExactData();
    if (!h || !x) { delete h; delete x; return fail(LPS_E_NOMEM, "out of host memory"); }
    h->mode = 1; h->x = x; h->device = device; h->m = m; h->n = n; h->m_total = m;
    static const bool trace_alloc = getenv("LPS_TRACE_ALLOC") != nullptr;
    const double tc0 = now_ms();
    if (stream) h->stream = (cudaStream_t)stream;
    else { CUX(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)); h->own_stream = true; }
    const double tc1 = now_ms();
    const int P = nprimes;
    x->P = P; x->Pcap = P;
    x->primes_h.assign(primes, primes + P);
    x->spt = P <= 1024 ? 1 : kMaxSPT;
    x->lane_threads = P <= 1024 ? std::max(32, (P + 31) / 32 * 32) : 1024;
    x->ld = lps_exact_ld(n);
    h->ld = x->ld;
    x->plane = (size_t)(m + 1) * x->ld;

    // prime-dependent tables (host side, exact integer / long double arithmetic)
    std::vector<PrimeQ> pr(P);
    for (int k = 0; k < P; ++k) pr[k] = make_prime(primes[k]);
    std::vector<uint32_t> half(P);
    {   // (M-1)/2 in the mixed radix (p_0, p_1, ...): long division by 2 from the top digit of M-1
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
        int ex = 1;                                       // W_0 = 1 = 0.5 * 2^1
        wm[0] = (double)mant; we[0] = ex;
        for (int t = 1; t <= P; ++t) {
            int e2;
            mant = frexpl(mant * (long double)primes[t - 1], &e2);
            ex += e2;
            wm[t] = (double)mant; we[t] = ex;
        }
    }
    x->wexp_h = we;
    // every small device buffer of the handle: one allocation (lps_internal.h, Slab)
    x->nbeta_ld = ((size_t)m + 1 + kXMaxRowBlock - 1) / kXMaxRowBlock * kXMaxRowBlock + kXMaxRowBlock;
    x->sstride = ((size_t)m + 1 + 3) / 4 * 4;
    const size_t nobj = (size_t)n + 1, ncr = 2 * ((size_t)m + 1), nrd = std::max(nobj, (size_t)m + 1);
    x->obj_dstride = (nobj + 7) / 8 * 8;
    x->rd_stride = (nrd + 7) / 8 * 8;
    {
        Slab sl;
        sl.add(&x->pr, (size_t)P); sl.add(&x->half, (size_t)P); sl.add(&x->wmant, (size_t)P + 1); sl.add(&x->wexp, (size_t)P + 1);
        sl.add(&x->invtab, (size_t)P * P);
        sl.add(&x->nbeta, (size_t)P * x->nbeta_ld);
        sl.add(&x->Dres, (size_t)P); sl.add(&x->dinv_m, (size_t)P); sl.add(&x->alpha_m, (size_t)P);
        sl.add(&x->sigma, (size_t)P * x->sstride); sl.add(&x->one, (size_t)P * x->sstride);
        sl.add(&x->alive, (size_t)m + 1);
        sl.add(&x->obj_sign, nobj); sl.add(&x->obj_top, nobj); sl.add(&x->obj_mant, nobj);
        sl.add(&x->cr_sign, ncr); sl.add(&x->cr_top, ncr); sl.add(&x->cr_mant, ncr);
        sl.add(&x->rd_sign, nrd); sl.add(&x->rd_top, nrd); sl.add(&x->rd_mant, nrd);
        sl.add(&x->obj_dig, (size_t)P * x->obj_dstride); sl.add(&x->rd_dig, (size_t)P * x->rd_stride);
        sl.add(&x->cand, (size_t)m); sl.add(&x->wcol, (size_t)n + 1); sl.add(&x->xbasis, (size_t)m);
        sl.add(&x->st, (size_t)1);
        sl.add(&h->basis, (size_t)m); sl.add(&h->marks, (size_t)n); sl.add(&h->scratch_i, (size_t)n + 8);
        CUX(sl.commit(&h->slab));
    }
    const double tc2 = now_ms();
    CUX(cudaMemcpyAsync(x->pr, pr.data(), P * sizeof(PrimeQ), cudaMemcpyHostToDevice, h->stream));
    CUX(cudaMemcpyAsync(x->half, half.data(), P * sizeof(uint32_t), cudaMemcpyHostToDevice, h->stream));
    CUX(cudaMemcpyAsync(x->wmant, wm.data(), (P + 1) * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CUX(cudaMemcpyAsync(x->wexp, we.data(), (P + 1) * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    k_xinvtab<<<dim3((P + 127) / 128, P), 128, 0, h->stream>>>(x->pr, P, x->invtab);
    CUX(cudaGetLastError());
    // the residue planes are allocated by the upload: all of them, or -- with column weights -- the working set only
    x->R = nullptr; x->Palloc = 0;
    CUX(cudaMemsetAsync(x->nbeta, 0, (size_t)P * x->nbeta_ld * sizeof(uint32_t), h->stream));
    static_assert(sizeof(XState) <= 256, "pinned state blocks are 256 bytes");
    x->st_host = (XState*)lps_pinned_get(sizeof(XState));
    if (!x->st_host) { lpsx_destroy(h); return fail(LPS_E_NOMEM, "no pinned host memory for the state mirror"); }
    CUX(cudaMemsetAsync(h->basis, 0xff, (size_t)m * sizeof(int), h->stream));
    CUX(cudaMemsetAsync(h->marks, 0, (size_t)n * sizeof(int), h->stream));
    CUX(cudaStreamSynchronize(h->stream));                // the host vectors go out of scope below
    if (trace_alloc) fprintf(stderr, "[lps] create: stream %.3f  tables+slab %.3f  copies+invtab+sync %.3f ms (P = %d)\n", tc1 - tc0, tc2 - tc1, now_ms() - tc2, P);
    *out = h;
    return LPS_OK;
}

int lpsx_destroy(lps_handle* h) {
    if (!h) return LPS_OK;
    static const bool trace = getenv("LPS_TRACE_ALLOC") != nullptr;
    const double t0 = now_ms();
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    const double t1 = now_ms();
    ExactData* x = h->x;
    if (x) {
        cudaFree(x->R);
        lps_pinned_put(x->st_host);
        delete x;
    }
    const double t2 = now_ms();
    if (h->gexec) cudaGraphExecDestroy(h->gexec);
    const double t3 = now_ms();
    for (cudaEvent_t e : h->ev) cudaEventDestroy(e);
    cudaFree(h->slab);
    if (h->trace) cudaFree(h->trace);
    const double t4 = now_ms();
    if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
    const double t5 = now_ms();
    if (trace) fprintf(stderr, "[lps] destroy: sync %.3f  free R %.3f  graph %.3f  free slab+trace %.3f  stream %.3f ms\n", t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4);
    delete h;
    return LPS_OK;
}

// ------------------------------------------------------------------------------------------------
// state helpers
// ------------------------------------------------------------------------------------------------
static int xread_state(lps_handle* h) {
    CU(cudaMemcpyAsync(h->x->st_host, h->x->st, sizeof(XState), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return LPS_OK;
}
static int xwrite_state(lps_handle* h) {
    CU(cudaMemcpyAsync(h->x->st, h->x->st_host, sizeof(XState), cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return LPS_OK;
}

// ------------------------------------------------------------------------------------------------
// working prime set (XState::wlimit): the first P of the Pcap allocated primes
// ------------------------------------------------------------------------------------------------
// bits (x 256) a number may have under the first P primes: |x| <= (M - 1) / 2, log2 M >= wexp[P] - 1
static long long cap256(const ExactData* x, int P) { return 256LL * ((long long)x->wexp_h[P] - 2); }

// at least `want` residue planes; the first `keep` planes survive a reallocation, new planes are zero
static int ensure_planes(lps_handle* h, int want, int keep) {
    ExactData* x = h->x;
    if (want <= x->Palloc) return LPS_OK;
    int np = std::min(x->Pcap, std::max(want, 2 * x->Palloc));
    uint32_t* nr = nullptr;
    cudaError_t e = cudaMalloc(&nr, (size_t)np * x->plane * sizeof(uint32_t));
    if (e != cudaSuccess && np > want) { cudaGetLastError(); np = want; e = cudaMalloc(&nr, (size_t)np * x->plane * sizeof(uint32_t)); }
    if (e != cudaSuccess) return fail(LPS_E_NOMEM, "cudaMalloc of %zu bytes (%d residue planes) failed: %s", (size_t)np * x->plane * sizeof(uint32_t), np, cudaGetErrorString(e));
    keep = (x->R && keep > 0) ? std::min(keep, x->Palloc) : 0;
    if (keep > 0) CU(cudaMemcpyAsync(nr, x->R, (size_t)keep * x->plane * sizeof(uint32_t), cudaMemcpyDeviceToDevice, h->stream));
    CU(cudaMemsetAsync(nr + (size_t)keep * x->plane, 0, (size_t)(np - keep) * x->plane * sizeof(uint32_t), h->stream));
    CU(cudaStreamSynchronize(h->stream));
    if (x->R) CU(cudaFree(x->R));
    x->R = nr; x->Palloc = np;
    return LPS_OK;
}

static int set_working(lps_handle* h, int P) {
    ExactData* x = h->x;
    x->P = P;
    x->spt = P <= 1024 ? 1 : kMaxSPT;
    x->lane_threads = P <= 1024 ? std::max(32, (P + 31) / 32 * 32) : 1024;
    std::vector<uint32_t> half(P);
    uint64_t carry = 0;                                   // (M-1)/2 in the mixed radix (p_0, ..., p_(P-1))
    for (int k = P - 1; k >= 0; --k) {
        const uint64_t cur = carry * x->primes_h[k] + (x->primes_h[k] - 1u);
        half[k] = (uint32_t)(cur / 2);
        carry = cur & 1u;
    }
    CU(cudaMemcpyAsync(x->half, half.data(), P * sizeof(uint32_t), cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    const size_t bytes = (size_t)P * x->plane * sizeof(uint32_t), cells = x->plane * P;
    x->stream_hint = bytes > kL2Bytes;
    if (cells >= (size_t)148 * 8 * 2048 * 64) { x->chunk = 2048; x->rows_per_cta = 64; }
    else if (cells >= (size_t)148 * 4 * 1024 * 16) { x->chunk = 1024; x->rows_per_cta = 16; }
    else { x->chunk = 1024; x->rows_per_cta = 8; }
    return LPS_OK;
}

// smallest working set that leaves `head` (x 256 bits) of room above `need`
static int pick_working(const ExactData* x, long long need, int at_least) {
    const long long target = need + std::max(need / 4, 256LL * 96);
    for (int P = std::max(1, at_least); P <= x->Pcap; ++P)
        if (cap256(x, P) >= target) return P;
    return x->Pcap;
}
static long long wlimit_of(const ExactData* x, int P) {
    return (P >= x->Pcap && x->static_ok) ? (1LL << 62)     // all primes: the static bound of the uploaded matrix holds
                                          : cap256(x, P) - 2 * x->wmax256 - x->guard256;
}

template <int NB>
static int launch_extend(lps_handle* h, int Pnew, uint32_t* vec, long long count) {
    ExactData* x = h->x;
    const size_t smem = (size_t)NB * x->P * sizeof(uint32_t);
    const unsigned blocks = (unsigned)((count + NB - 1) / NB);
    if (x->spt == 1) {
        if (smem > 48 * 1024) CU(cudaFuncSetAttribute(k_xextend<NB, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_xextend<NB, 1><<<blocks, x->lane_threads, smem, h->stream>>>(xgrid_of(h), Pnew, vec, count);
    } else {
        if (smem > 48 * 1024) CU(cudaFuncSetAttribute(k_xextend<NB, kMaxSPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_xextend<NB, kMaxSPT><<<blocks, x->lane_threads, smem, h->stream>>>(xgrid_of(h), Pnew, vec, count);
    }
    h->launches_all++;
    CU(cudaGetLastError());
    return LPS_OK;
}

// LPS_NEED_PRIMES was reported (nothing modified): extend the working set so that the pending pivot fits, with
// head-room, and clear the status.  The residues modulo the new primes are computed on the device.
// The working prime set grows from x->P to Pnew primes: the common denominator D first -- a new prime that divides D
// has no D^-1 and cannot join (the pivot checks only see the primes that were in the set when the pivot was made) --
// then the grid.  LPS_BAD_PRIME with *bad = index of such a prime, nothing modified inside the working set.
static int extend_working(lps_handle* h, int Pnew, int* bad) {
    ExactData* x = h->x;
    const int Pold = x->P;
    *bad = -1;
    TRY(ensure_planes(h, Pnew, Pold));
    TRY(launch_extend<4>(h, Pnew, x->Dres, 1));
    std::vector<uint32_t> dres((size_t)(Pnew - Pold));
    CU(cudaMemcpyAsync(dres.data(), x->Dres + Pold, dres.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    for (int k = 0; k < Pnew - Pold; ++k)
        if (dres[(size_t)k] == 0) { *bad = Pold + k; return LPS_BAD_PRIME; }
    TRY(launch_extend<4>(h, Pnew, nullptr, (long long)(h->m + 1) * (h->n + 1)));
    TRY(set_working(h, Pnew));
    x->grows++;
    return LPS_OK;
}

static int grow_working(lps_handle* h) {
    ExactData* x = h->x;
    TRY(xread_state(h));
    XState& s = *x->st_host;
    const long long need = s.wnext + 2 * x->wmax256 + x->guard256;
    const int Pnew = pick_working(x, need, x->P + 1);
    if (Pnew > x->P) {
        int bad = -1;
        const int rc = extend_working(h, Pnew, &bad);
        if (rc < 0) return rc;
        if (rc == LPS_BAD_PRIME) {                       // reported like a prime dividing the next pivot: the host
            s.status = LPS_BAD_PRIME;                    // re-encodes without it and resumes
            s.bad_prime = bad;
            TRY(xwrite_state(h));
            return LPS_BAD_PRIME;
        }
    }
    s.wlimit = wlimit_of(x, x->P);
    if (s.wnext > s.wlimit) return LPS_NEED_PRIMES;      // every prime of the handle is in use and the bound (grown by row
                                                         // operations) still does not fit: the caller re-encodes the tableau
    s.status = LPS_OK;
    TRY(xwrite_state(h));
    return LPS_OK;
}

// Column weights of the uploaded integer matrix (wcol[j] >= 256 log2 of the 2-norm of column j over all m+1 rows,
// j = 0..n) and the guard (bits x 256: row scales sigma_i and rounding slack).  Switches the handle to the smallest
// working set that holds the uploaded matrix with head-room; pivots then extend it on demand.  Call right after
// lps_upload_exact; without it every kernel runs on all the primes of lps_create_exact.
extern "C" int lps_exact_set_weights(lps_handle* h, const int* wcol, int guard_bits) {
    USE(h);
    NEED_MODE(h, 1);
    ExactData* x = h->x;
    if (!wcol || guard_bits < 0) return fail(LPS_E_INVALID, "bad weights");
    if (!x->R) return fail(LPS_E_STATE, "upload first");
    long long wmax = 0;
    for (int j = 0; j <= h->n; ++j) {
        if (wcol[j] < 0) return fail(LPS_E_INVALID, "negative weight");
        wmax = std::max<long long>(wmax, wcol[j]);
    }
    x->wmax256 = wmax; x->guard256 = 256LL * guard_bits; x->adaptive = true; x->grows = 0;
    CU(cudaMemcpyAsync(x->wcol, wcol, ((size_t)h->n + 1) * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemsetAsync(x->xbasis, 0xff, (size_t)h->m * sizeof(int), h->stream));
    TRY(set_working(h, pick_working(x, 2 * wmax + x->guard256, 1)));
    TRY(xread_state(h));
    x->st_host->wsum = 0; x->st_host->wnext = 0; x->st_host->wlimit = wlimit_of(x, x->P);
    TRY(xwrite_state(h));
    return LPS_OK;
}

// ------------------------------------------------------------------------------------------------
// upload / read-out
// ------------------------------------------------------------------------------------------------
static int upload_common(lps_handle* h, const uint32_t* limbs, int nlimbs, const unsigned char* neg, size_t ld_host,
                         const uint32_t* sigma_limbs, int sigma_nlimbs, const int* wcol, int guard_bits) {
    ExactData* x = h->x;
    if (!limbs || !neg || nlimbs <= 0 || ld_host < (size_t)h->n + 1) return fail(LPS_E_INVALID, "bad host grid");
    if (!sigma_limbs || sigma_nlimbs <= 0) return fail(LPS_E_INVALID, "bad row scales");
    // which primes carry the tableau from the start: all of them, or the working set the weights ask for
    int P0 = x->Pcap;
    x->adaptive = false;
    x->grows = 0; x->resume = false; x->static_ok = true;   // a handle may serve several tableaux in a row (the host pools them)
    if (wcol) {
        if (guard_bits < 0) return fail(LPS_E_INVALID, "bad guard");
        long long wmax = 0;
        for (int j = 0; j <= h->n; ++j) {
            if (wcol[j] < 0) return fail(LPS_E_INVALID, "negative weight");
            wmax = std::max<long long>(wmax, wcol[j]);
        }
        x->wmax256 = wmax; x->guard256 = 256LL * guard_bits; x->adaptive = true; x->grows = 0;
        P0 = pick_working(x, 2 * wmax + x->guard256, 1);
    }
    if (!x->adaptive) TRY(ensure_planes(h, x->Pcap, 0));
    else if (x->Palloc < P0) TRY(ensure_planes(h, std::min(x->Pcap, std::max(2 * P0, 8)), 0));
    TRY(set_working(h, P0));
    const size_t entries = (size_t)(h->m + 1) * ld_host;
    DevBuf b_limbs, b_neg, b_sig;
    CU(b_limbs.alloc(entries * nlimbs * sizeof(uint32_t)));
    CU(b_neg.alloc(entries));
    CU(b_sig.alloc(((size_t)h->m + 1) * sigma_nlimbs * sizeof(uint32_t)));
    uint32_t* d_limbs = b_limbs.as<uint32_t>();
    unsigned char* d_neg = b_neg.as<unsigned char>();
    uint32_t* d_sig = b_sig.as<uint32_t>();
    CU(cudaMemcpyAsync(d_limbs, limbs, entries * nlimbs * sizeof(uint32_t), cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(d_neg, neg, entries, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(d_sig, sigma_limbs, ((size_t)h->m + 1) * sigma_nlimbs * sizeof(uint32_t), cudaMemcpyHostToDevice, h->stream));
    const XGrid g = xgrid_of(h);
    const unsigned bx = (unsigned)std::min<size_t>((x->plane + 255) / 256, 148 * 8);
    if (x->Palloc > x->P)                                    // planes beyond the working set: zero until an extension fills them
        CU(cudaMemsetAsync(x->R + (size_t)x->P * x->plane, 0, (size_t)(x->Palloc - x->P) * x->plane * sizeof(uint32_t), h->stream));
    k_xencode<<<dim3(bx, x->P), 256, 0, h->stream>>>(g, d_limbs, nlimbs, d_neg, ld_host);
    k_xencode_vec<<<dim3((h->m + 256) / 256, x->Pcap), 256, 0, h->stream>>>(g, d_sig, sigma_nlimbs, h->m + 1, x->sigma, x->sstride);   // row scales: every prime
    h->launches_all += 2;
    CU(cudaGetLastError());
    std::vector<uint32_t> ones(x->Pcap, 1u);
    CU(cudaMemcpyAsync(x->Dres, ones.data(), x->Pcap * sizeof(uint32_t), cudaMemcpyHostToDevice, h->stream));
    std::vector<int> alive(h->m + 1, 1);
    CU(cudaMemcpyAsync(x->alive, alive.data(), alive.size() * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    if (x->adaptive) {
        CU(cudaMemcpyAsync(x->wcol, wcol, ((size_t)h->n + 1) * sizeof(int), cudaMemcpyHostToDevice, h->stream));
        CU(cudaMemsetAsync(x->xbasis, 0xff, (size_t)h->m * sizeof(int), h->stream));
    }
    XState s{};
    s.status = LPS_OK; s.max_pivots = -1; s.sgnD = 1; s.c = s.r = s.last_r = s.last_c = -1; s.bad_prime = -1;
    s.wsum = 0; s.wnext = 0; s.wlimit = x->adaptive ? wlimit_of(x, x->P) : (1LL << 62);
    *x->st_host = s;
    CU(cudaMemcpyAsync(x->st, x->st_host, sizeof(XState), cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return LPS_OK;
}

extern "C" int lps_upload_exact(lps_handle* h, const uint32_t* limbs, int nlimbs, const unsigned char* neg, size_t ld_host,
                                const uint32_t* sigma_limbs, int sigma_nlimbs) {
    USE(h);
    NEED_MODE(h, 1);
    return upload_common(h, limbs, nlimbs, neg, ld_host, sigma_limbs, sigma_nlimbs, nullptr, 0);
}

// Upload with the column weights of lps_exact_set_weights in the same call: only the working set of residue planes
// is allocated and encoded (C3: 5 planes of 8 MB instead of 362); the rest appears when an extension needs it.
extern "C" int lps_upload_exact_weighted(lps_handle* h, const uint32_t* limbs, int nlimbs, const unsigned char* neg, size_t ld_host,
                                         const uint32_t* sigma_limbs, int sigma_nlimbs, const int* wcol, int guard_bits) {
    USE(h);
    NEED_MODE(h, 1);
    if (!wcol) return fail(LPS_E_INVALID, "null weights");
    return upload_common(h, limbs, nlimbs, neg, ld_host, sigma_limbs, sigma_nlimbs, wcol, guard_bits);
}

// Row operations outside a pivot on the residues (rowMult / rowDiv tableau.py:254-267, rowAdd / rowSub :269-280,
// rowAddToObj / rowSubFromObj :282-293):   N_rd <- a N_rd + b N_rs   (rs < 0: N_rd <- a N_rd),   rd == m: the
// objective row.  a, b: signed multi-limb integers; sigma_new: row rd's new scale (the row carries a scale again).
// The numbers stop being minors of the uploaded matrix: guard_bits is the caller's new allowance on top of the
// column-weight bound (it grows by log2(|a| + |b|) per operation) and the static bound no longer waives the capacity
// test.  LPS_NEED_PRIMES, with nothing modified, when even all the handle's primes cannot hold that.
extern "C" int lps_row_combine_exact(lps_handle* h, int rd, int rs, const uint32_t* a_limbs, int a_nlimbs, int a_neg,
                                     const uint32_t* b_limbs, int b_nlimbs, int b_neg, const uint32_t* sigma_limbs, int sigma_nlimbs,
                                     int guard_bits) {
    USE(h);
    NEED_MODE(h, 1);
    ExactData* x = h->x;
    if (!x->R) return fail(LPS_E_STATE, "upload first");
    if (!x->adaptive) return fail(LPS_E_STATE, "row operations on the residues need the working prime set (column weights)");
    if (rd < 0 || rd > h->m || rs >= h->m) return fail(LPS_E_INVALID, "rows (%d,%d) out of range", rd, rs);
    if (!a_limbs || a_nlimbs <= 0 || (rs >= 0 && (!b_limbs || b_nlimbs <= 0)) || !sigma_limbs || sigma_nlimbs <= 0 || guard_bits < 0)
        return fail(LPS_E_INVALID, "bad arguments");
    TRY(xread_state(h));
    XState& s = *x->st_host;
    if (s.status == LPS_NEED_PRIMES || s.status == LPS_BAD_PRIME) return fail(LPS_E_STATE, "a pivot is pending");
    const long long guard_new = 256LL * guard_bits;
    const long long need = s.wsum + 2 * x->wmax256 + guard_new;
    if (cap256(x, x->Pcap) < need) return LPS_NEED_PRIMES;
    x->guard256 = guard_new;
    x->static_ok = false;
    if (s.wsum > wlimit_of(x, x->P)) {
        const int Pnew = pick_working(x, need, x->P + 1);
        if (Pnew > x->P) {
            int bad = -1;
            const int rc = extend_working(h, Pnew, &bad);
            if (rc < 0) return rc;
            if (rc == LPS_BAD_PRIME) return LPS_NEED_PRIMES;   // a new prime divides D: refused, nothing modified (host copy)
        }
    }
    s.wlimit = wlimit_of(x, x->P);
    s.wnext = s.wsum;
    TRY(xwrite_state(h));
    const XGrid g = xgrid_of(h);
    DevBuf buf_limbs, b_res;
    const size_t nl = (size_t)a_nlimbs + (rs >= 0 ? b_nlimbs : 0) + sigma_nlimbs;
    CU(buf_limbs.alloc(nl * sizeof(uint32_t) + 2));
    CU(b_res.alloc(2 * (size_t)x->P * sizeof(uint32_t)));
    uint32_t* d_a = buf_limbs.as<uint32_t>();
    uint32_t* d_b = d_a + a_nlimbs;
    uint32_t* d_s = d_b + (rs >= 0 ? b_nlimbs : 0);
    unsigned char negs[2] = {(unsigned char)(a_neg ? 1 : 0), (unsigned char)(b_neg ? 1 : 0)};
    unsigned char* d_neg = reinterpret_cast<unsigned char*>(d_s + sigma_nlimbs);
    CU(cudaMemcpyAsync(d_a, a_limbs, (size_t)a_nlimbs * sizeof(uint32_t), cudaMemcpyHostToDevice, h->stream));
    if (rs >= 0) CU(cudaMemcpyAsync(d_b, b_limbs, (size_t)b_nlimbs * sizeof(uint32_t), cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(d_s, sigma_limbs, (size_t)sigma_nlimbs * sizeof(uint32_t), cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(d_neg, negs, 2, cudaMemcpyHostToDevice, h->stream));
    uint32_t* ra = b_res.as<uint32_t>();
    uint32_t* rb = ra + x->P;
    k_xencode_vec<<<dim3(1, x->P), 32, 0, h->stream>>>(g, d_a, a_nlimbs, 1, ra, 1, d_neg);
    if (rs >= 0) k_xencode_vec<<<dim3(1, x->P), 32, 0, h->stream>>>(g, d_b, b_nlimbs, 1, rb, 1, d_neg + 1);
    k_xencode_vec<<<dim3(1, x->Pcap), 32, 0, h->stream>>>(g, d_s, sigma_nlimbs, 1, x->sigma + rd, x->sstride);
    k_xrow_combine<<<dim3(std::max(1, std::min((h->n + 256) / 256, 64)), x->P), 256, 0, h->stream>>>(g, rd, rs, ra, rb);
    h->launches_all += 4;
    CU(cudaGetLastError());
    const int one = 1;
    CU(cudaMemcpyAsync(x->alive + rd, &one, sizeof(int), cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return LPS_OK;
}

// Leaving phase 1 without a host round trip of the tableau (Simplex._find_bfs simplex.py:95-105 when no row is
// dropped).  The artificial columns n_new..n-1 are cut and the objective row becomes the original objective priced
// out against the basis (k_xrestore_obj).  cost: n_new + 1 signed integers C_j = sigma_m c_j (C_n: the _z cell);
// mult: m signed integers G_i = C_{basis[i]} / s_i; sigma_m: the new scale of the objective row.  From here on the
// numbers of the tableau are minors of the matrix [constraint rows as uploaded ; C] over the basis columns of ALL
// rows: wcol holds that matrix's column weights (lps_exact_set_weights), every row's basis column counts, and the
// working prime set is extended first if it must.  The caller sized the handle's prime set for that matrix too.
extern "C" int lps_phase2_exact(lps_handle* h, int n_new, const uint32_t* cost_limbs, int cost_nlimbs, const unsigned char* cost_neg,
                                const uint32_t* mult_limbs, int mult_nlimbs, const unsigned char* mult_neg,
                                const uint32_t* sigma_m_limbs, int sigma_m_nlimbs, const int* basis, const int* wcol, int guard_bits) {
    USE(h);
    NEED_MODE(h, 1);
    ExactData* x = h->x;
    if (!x->R) return fail(LPS_E_STATE, "upload first");
    if (n_new <= 0 || n_new > h->n) return fail(LPS_E_INVALID, "n_new = %d outside 1..%d", n_new, h->n);
    if (!cost_limbs || !cost_neg || cost_nlimbs <= 0 || !mult_limbs || !mult_neg || mult_nlimbs <= 0 || !sigma_m_limbs ||
        sigma_m_nlimbs <= 0 || !basis)
        return fail(LPS_E_INVALID, "null argument");
    for (int i = 0; i < h->m; ++i)
        if (basis[i] < 0 || basis[i] >= n_new) return fail(LPS_E_INVALID, "basis[%d] = %d is not a column below %d", i, basis[i], n_new);
    if (x->adaptive && (!wcol || guard_bits < 0)) return fail(LPS_E_INVALID, "the working prime set needs the new column weights");
    static const bool trace = getenv("LPS_TRACE_ALLOC") != nullptr;
    const double tp0 = now_ms();
    TRY(xread_state(h));
    if (x->st_host->status == LPS_NEED_PRIMES || x->st_host->status == LPS_BAD_PRIME) return fail(LPS_E_STATE, "a pivot is pending");

    // 1. cut the artificial columns (nbeta is free between pivots: scratch for the right-hand side)
    if (n_new < h->n) {
        const XGrid g = xgrid_of(h);
        k_xgather_rhs<<<dim3(std::max(1, std::min((h->m + 256) / 256, 64)), x->P), 256, 0, h->stream>>>(g, x->nbeta, x->nbeta_ld);
        const long long cells = (long long)(h->m + 1) * (h->n - n_new + 1);
        k_xcut_cols<<<dim3((unsigned)std::max<long long>(1, std::min<long long>((cells + 255) / 256, 148 * 4)), x->P), 256, 0, h->stream>>>(g, n_new, x->nbeta, x->nbeta_ld);
        h->launches_all += 2;
        CU(cudaGetLastError());
        h->n = n_new;                                    // ld, planes and buffers keep their (larger) sizes
        if (h->gexec) { cudaGraphExecDestroy(h->gexec); h->gexec = nullptr; }   // the captured launches carry the old n
    }

    if (trace) { cudaStreamSynchronize(h->stream); fprintf(stderr, "[lps] phase2: cut %.3f ms\n", now_ms() - tp0); }
    const double tp1 = now_ms();
    // 2. weights of the new matrix; every row's basis column counts from now on
    if (x->adaptive) {
        long long wmax = 0, wsum = 0;
        for (int j = 0; j <= n_new; ++j) {
            if (wcol[j] < 0) return fail(LPS_E_INVALID, "negative weight");
            wmax = std::max<long long>(wmax, wcol[j]);
        }
        for (int i = 0; i < h->m; ++i) wsum += wcol[basis[i]];
        x->wmax256 = wmax; x->guard256 = 256LL * guard_bits;
        CU(cudaMemcpyAsync(x->wcol, wcol, ((size_t)n_new + 1) * sizeof(int), cudaMemcpyHostToDevice, h->stream));
        CU(cudaMemcpyAsync(x->xbasis, basis, (size_t)h->m * sizeof(int), cudaMemcpyHostToDevice, h->stream));
        if (wsum > wlimit_of(x, x->P)) {                 // the restored objective row may not fit the working set
            const int Pnew = pick_working(x, wsum + 2 * wmax + x->guard256, x->P + 1);
            if (Pnew > x->P) {
                int bad = -1;
                const int rc = extend_working(h, Pnew, &bad);
                if (rc < 0) return rc;
                // a new prime divides D (about 2^-31 per prime): the columns are already cut, so this call cannot be
                // refused cleanly -- fail loudly rather than carry a prime without D^-1
                if (rc == LPS_BAD_PRIME)
                    return fail(LPS_E_STATE, "prime %d of the handle divides the common denominator: re-encode the tableau without it", bad);
            }
        }
        XState& s = *x->st_host;
        s.wsum = wsum; s.wnext = wsum; s.wlimit = wlimit_of(x, x->P);
    }
    x->st_host->status = LPS_OK;
    TRY(xwrite_state(h));

    if (trace) { cudaStreamSynchronize(h->stream); fprintf(stderr, "[lps] phase2: weights + extension %.3f ms (P = %d)\n", now_ms() - tp1, x->P); }
    const double tp2 = now_ms();
    // 3. residues of C and G on the working planes, of sigma_m on every prime; the new objective row
    const XGrid g = xgrid_of(h);
    const size_t cs = ((size_t)n_new + 1 + 3) / 4 * 4, gs = ((size_t)h->m + 3) / 4 * 4;
    DevBuf b_cl, b_cn, b_ml, b_mn, b_sl, b_c, b_g;
    CU(b_cl.alloc(((size_t)n_new + 1) * cost_nlimbs * sizeof(uint32_t)));
    CU(b_cn.alloc((size_t)n_new + 1));
    CU(b_ml.alloc((size_t)h->m * mult_nlimbs * sizeof(uint32_t)));
    CU(b_mn.alloc((size_t)h->m));
    CU(b_sl.alloc((size_t)sigma_m_nlimbs * sizeof(uint32_t)));
    CU(b_c.alloc((size_t)x->P * cs * sizeof(uint32_t)));
    CU(b_g.alloc((size_t)x->P * gs * sizeof(uint32_t)));
    CU(cudaMemcpyAsync(b_cl.p, cost_limbs, ((size_t)n_new + 1) * cost_nlimbs * sizeof(uint32_t), cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(b_cn.p, cost_neg, (size_t)n_new + 1, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(b_ml.p, mult_limbs, (size_t)h->m * mult_nlimbs * sizeof(uint32_t), cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(b_mn.p, mult_neg, (size_t)h->m, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(b_sl.p, sigma_m_limbs, (size_t)sigma_m_nlimbs * sizeof(uint32_t), cudaMemcpyHostToDevice, h->stream));
    k_xencode_vec<<<dim3((n_new + 256) / 256, x->P), 256, 0, h->stream>>>(g, b_cl.as<uint32_t>(), cost_nlimbs, n_new + 1, b_c.as<uint32_t>(), cs, b_cn.as<unsigned char>());
    k_xencode_vec<<<dim3((h->m + 255) / 256, x->P), 256, 0, h->stream>>>(g, b_ml.as<uint32_t>(), mult_nlimbs, h->m, b_g.as<uint32_t>(), gs, b_mn.as<unsigned char>());
    k_xencode_vec<<<dim3(1, x->Pcap), 256, 0, h->stream>>>(g, b_sl.as<uint32_t>(), sigma_m_nlimbs, 1, x->sigma + h->m, x->sstride);
    k_xrestore_obj<<<dim3((n_new + 1 + 127) / 128, x->P), 128, 0, h->stream>>>(g, x->Dres, b_c.as<uint32_t>(), cs, b_g.as<uint32_t>(), gs);
    h->launches_all += 4;
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(h->stream));                // the scratch buffers go out of scope
    if (trace) fprintf(stderr, "[lps] phase2: objective row %.3f ms\n", now_ms() - tp2);
    return LPS_OK;
}

template <int NB>
static int launch_garner(lps_handle* h, const GarnerJob& job) {
    const ExactData* x = h->x;
    static const bool no_warp = getenv("LPS_GARNER_BLOCK") != nullptr;   // A/B switch for measurements
    if (x->P <= 128 && !no_warp) {                       // one number per warp, digits by shuffle
        const int blocks = (job.count + kGarnerWarps - 1) / kGarnerWarps;
        if (x->P <= 32) k_garner_w<1><<<blocks, 32 * kGarnerWarps, 0, h->stream>>>(xgrid_of(h), x->st, job);
        else if (x->P <= 64) k_garner_w<2><<<blocks, 32 * kGarnerWarps, 0, h->stream>>>(xgrid_of(h), x->st, job);
        else k_garner_w<4><<<blocks, 32 * kGarnerWarps, 0, h->stream>>>(xgrid_of(h), x->st, job);
        h->launches_all++;
        CU(cudaGetLastError());
        return LPS_OK;
    }
    const int blocks = (job.count + NB - 1) / NB;
    if (x->spt == 1) k_garner<NB, 1><<<blocks, x->lane_threads, 0, h->stream>>>(xgrid_of(h), x->st, job);
    else k_garner<NB, kMaxSPT><<<blocks, x->lane_threads, 0, h->stream>>>(xgrid_of(h), x->st, job);
    h->launches_all++;
    CU(cudaGetLastError());
    return LPS_OK;
}

// what: 0 = row `index` (n+1 numbers), 1 = column `index` (m+1 numbers), 2 = the denominator D (1 number).
// sign_out: count bytes (-1/0/+1); digits_out: P x count uint32, digit k of |x_e| at [k*count + e].
extern "C" int lps_read_exact(lps_handle* h, int what, int index, signed char* sign_out, uint32_t* digits_out) {
    USE(h);
    NEED_MODE(h, 1);
    ExactData* x = h->x;
    if (!sign_out || !digits_out) return fail(LPS_E_INVALID, "null output");
    GarnerJob job{};
    job.check_status = 0;
    job.meta = XMeta{x->rd_sign, x->rd_top, x->rd_mant};
    job.dig = x->rd_dig; job.dstride = x->rd_stride;
    if (what == 0) {
        if (index < 0 || index > h->m) return fail(LPS_E_INVALID, "row %d out of range", index);
        job.src = XSRC_ROW; job.index = index; job.count = h->n + 1;
    } else if (what == 1) {
        if (index < 0 || index > h->n) return fail(LPS_E_INVALID, "column %d out of range", index);
        job.src = XSRC_COL; job.index = index; job.count = h->m + 1;
    } else if (what == 2) {
        job.src = XSRC_CUSTOM; job.custom = x->Dres; job.cstride = 1; job.count = 1; job.index = 0;
    } else return fail(LPS_E_INVALID, "bad selector %d", what);
    TRY(launch_garner<8>(h, job));
    CU(cudaMemcpyAsync(sign_out, x->rd_sign, job.count, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaMemcpy2DAsync(digits_out, (size_t)job.count * sizeof(uint32_t), x->rd_dig, x->rd_stride * sizeof(uint32_t),
                         (size_t)job.count * sizeof(uint32_t), x->P, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return LPS_OK;
}

// Scattered cells (rows[e], cols[e]), e < count <= max(n, m) + 1, reconstructed like lps_read_exact does rows.
extern "C" int lps_read_exact_cells(lps_handle* h, int count, const int* rows, const int* cols, signed char* sign_out, uint32_t* digits_out) {
    USE(h);
    NEED_MODE(h, 1);
    ExactData* x = h->x;
    if (count <= 0 || (size_t)count > x->rd_stride || !rows || !cols || !sign_out || !digits_out) return fail(LPS_E_INVALID, "bad arguments");
    for (int e = 0; e < count; ++e)
        if (rows[e] < 0 || rows[e] > h->m || cols[e] < 0 || cols[e] > h->n) return fail(LPS_E_INVALID, "cell (%d,%d) out of range", rows[e], cols[e]);
    DevBuf b_idx, b_res;
    CU(b_idx.alloc(2 * (size_t)count * sizeof(int)));
    CU(b_res.alloc((size_t)x->P * count * sizeof(uint32_t)));
    int* d_rows = b_idx.as<int>();
    int* d_cols = d_rows + count;
    CU(cudaMemcpyAsync(d_rows, rows, (size_t)count * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(d_cols, cols, (size_t)count * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    k_xgather_cells<<<dim3((count + 255) / 256, x->P), 256, 0, h->stream>>>(xgrid_of(h), d_rows, d_cols, count, b_res.as<uint32_t>(), (size_t)count);
    h->launches_all++;
    GarnerJob job{};
    job.check_status = 0;
    job.meta = XMeta{x->rd_sign, x->rd_top, x->rd_mant};
    job.dig = x->rd_dig; job.dstride = x->rd_stride;
    job.src = XSRC_CUSTOM; job.custom = b_res.as<uint32_t>(); job.cstride = (size_t)count; job.count = count; job.index = 0;
    TRY(launch_garner<8>(h, job));
    CU(cudaMemcpyAsync(sign_out, x->rd_sign, count, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaMemcpy2DAsync(digits_out, (size_t)count * sizeof(uint32_t), x->rd_dig, x->rd_stride * sizeof(uint32_t),
                         (size_t)count * sizeof(uint32_t), x->P, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return LPS_OK;
}

// Simplex.findPivotAll (simplex.py:330-360) and the per-column half of Simplex.findPivotMaxIncrease (:286-328) in
// exact mode: the ratio test of every column on the device, a batch of columns per launch pair (reconstruction of
// the (column, rhs) pairs, then k_xratio_cols: one block per column).  minrow[j] (n ints): the lowest row of minimum
// ratio b_i / a_ij over a_ij > 0; -1: no positive entry; -2: column not tested (only_negative and c_j >= 0).  With
// tie_off (n + 1 offsets) and ties: every row of minimum ratio, ascending, column after column; at most ties_cap
// entries are written, tie_off[n] is the total.
extern "C" int lps_exact_col_ratios(lps_handle* h, int only_negative, int* minrow, long long* tie_off, int* ties, long long ties_cap) {
    USE(h);
    NEED_MODE(h, 1);
    ExactData* x = h->x;
    if (!minrow || (tie_off && ties_cap > 0 && !ties) || ties_cap < 0) return fail(LPS_E_INVALID, "bad arguments");
    if (!x->R) return fail(LPS_E_STATE, "upload first");
    const int m = h->m, n = h->n;
    const long long per = 2LL * (m + 1);
    const int batch = (int)std::max<long long>(1, std::min<long long>(n, (1LL << 22) / per));
    const bool want_ties = tie_off != nullptr;
    DevBuf b_sign, b_top, b_mant, b_cand, b_min, b_nt, b_ties;
    CU(b_sign.alloc((size_t)batch * per));
    CU(b_top.alloc((size_t)batch * per * sizeof(int)));
    CU(b_mant.alloc((size_t)batch * per * sizeof(double)));
    CU(b_cand.alloc((size_t)batch * m * sizeof(int)));
    CU(b_min.alloc((size_t)batch * sizeof(int)));
    if (want_ties) { CU(b_nt.alloc((size_t)batch * sizeof(int))); CU(b_ties.alloc((size_t)batch * m * sizeof(int))); }
    if (only_negative) {                                  // signs of the reduced costs
        GarnerJob job{};
        job.src = XSRC_ROW; job.index = m; job.count = n + 1; job.check_status = 0;
        job.meta = XMeta{x->obj_sign, x->obj_top, x->obj_mant};
        TRY(launch_garner<8>(h, job));
    }
    k_xdinv<<<1, 1024, 0, h->stream>>>(xgrid_of(h), pconst_of(h));
    h->launches_all++;
    std::vector<int> h_nt(batch), h_ties;
    if (want_ties) h_ties.resize((size_t)batch * m);
    long long total = 0;
    for (int j0 = 0; j0 < n; j0 += batch) {
        const int nc = std::min(batch, n - j0);
        GarnerJob cj{};
        cj.src = XSRC_COLPAIRS; cj.index = j0; cj.count = (int)(nc * per); cj.check_status = 0;
        cj.meta = XMeta{b_sign.as<signed char>(), b_top.as<int>(), b_mant.as<double>()};
        TRY(launch_garner<8>(h, cj));
        XColsArgs a{xgrid_of(h), x->st, cj.meta, x->dinv_m, j0, nc, only_negative ? x->obj_sign : nullptr,
                    b_cand.as<int>(), b_min.as<int>(), want_ties ? b_nt.as<int>() : nullptr, want_ties ? b_ties.as<int>() : nullptr};
        if (x->spt == 1) k_xratio_cols<1><<<nc, kLaneThreads, 0, h->stream>>>(a);
        else k_xratio_cols<kMaxSPT><<<nc, kLaneThreads, 0, h->stream>>>(a);
        h->launches_all++;
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(minrow + j0, b_min.p, (size_t)nc * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        if (want_ties) {
            CU(cudaMemcpyAsync(h_nt.data(), b_nt.p, (size_t)nc * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
            CU(cudaMemcpyAsync(h_ties.data(), b_ties.p, (size_t)nc * m * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        }
        CU(cudaStreamSynchronize(h->stream));
        if (want_ties) {
            for (int jj = 0; jj < nc; ++jj) {
                tie_off[j0 + jj] = total;
                int* src = h_ties.data() + (size_t)jj * m;
                std::sort(src, src + h_nt[jj]);
                for (int t = 0; t < h_nt[jj]; ++t, ++total)
                    if (total < ties_cap) ties[total] = src[t];
            }
        }
    }
    if (want_ties) tie_off[n] = total;
    return LPS_OK;
}

extern "C" int lps_read_alive(lps_handle* h, int* alive) {
    USE(h);
    NEED_MODE(h, 1);
    if (!alive) return fail(LPS_E_INVALID, "null output");
    CU(cudaMemcpyAsync(alive, h->x->alive, ((size_t)h->m + 1) * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return LPS_OK;
}

extern "C" int lps_exact_info(lps_handle* h, int* nprimes, int* bad_prime, size_t* residue_bytes) {
    USE(h);
    NEED_MODE(h, 1);
    TRY(xread_state(h));
    if (nprimes) *nprimes = h->x->P;
    if (bad_prime) *bad_prime = h->x->st_host->bad_prime;
    if (residue_bytes) *residue_bytes = (size_t)h->x->P * h->x->plane * sizeof(uint32_t);
    return LPS_OK;
}

extern "C" int lps_exact_working(lps_handle* h, int* working, int* allocated, int* extensions) {
    USE(h);
    NEED_MODE(h, 1);
    if (working) *working = h->x->P;
    if (allocated) *allocated = h->x->Pcap;
    if (extensions) *extensions = h->x->grows;
    return LPS_OK;
}

// Loop state of Simplex.solve (rule, stuck counter, "objective moved", pending objective check):
// set == 0 reads it; set != 0 writes it and makes the next lps_solve continue from it instead of
// starting over (used when a solve is resumed on a re-encoded tableau after LPS_BAD_PRIME).
extern "C" int lps_exact_loop_state(lps_handle* h, int set, int* rule, long long* stuck, int* moved, int* pending) {
    USE(h);
    NEED_MODE(h, 1);
    if (!rule || !stuck || !moved || !pending) return fail(LPS_E_INVALID, "null argument");
    TRY(xread_state(h));
    XState& s = *h->x->st_host;
    if (!set) { *rule = s.rule; *stuck = s.stuck; *moved = s.moved; *pending = s.pending_zcheck; return LPS_OK; }
    s.rule = *rule; s.stuck = *stuck; s.moved = *moved; s.pending_zcheck = *pending;
    TRY(xwrite_state(h));
    h->x->resume = true;
    return LPS_OK;
}

// ------------------------------------------------------------------------------------------------
// launches of one pivot
// ------------------------------------------------------------------------------------------------
// Small tableaux are launch bound (a pivot's six kernels take a few microseconds each): the single-block decision
// kernels then reconstruct their numbers themselves and the ratio kernel stages the entering column, three launches
// per pivot instead of six.  *staged tells the caller whether k_xstage is still due.
static bool fuse_small(const ExactData* x) {
    static const bool off = getenv("LPS_NO_FUSE") != nullptr;     // A/B switch for measurements
    return !off && x->P <= 64;
}
static int xlaunch_select(lps_handle* h, bool* staged = nullptr) {
    ExactData* x = h->x;
    const bool small = fuse_small(x);
    GarnerJob job{};
    job.src = XSRC_ROW; job.index = h->m; job.count = h->n + 1; job.check_status = 1;
    job.meta = XMeta{x->obj_sign, x->obj_top, x->obj_mant};
    job.dig = x->obj_dig; job.dstride = x->obj_dstride;
    XPriceArgs pa{xgrid_of(h), x->st, job.meta, x->obj_dig, x->obj_dstride};
    if (small && job.count <= 128) {                     // at most four numbers per warp of the block: beyond that
                                                         // a grid of warps (k_garner_w) wins (profiles/r01f_ab_fuse.txt)
        if (x->P <= 32) k_xprice_fused<1><<<1, kLaneThreads, 0, h->stream>>>(pa, job);
        else k_xprice_fused<2><<<1, kLaneThreads, 0, h->stream>>>(pa, job);
        h->launches_all += 1;
    } else {
        TRY(launch_garner<8>(h, job));
        k_xprice<<<1, kLaneThreads, 0, h->stream>>>(pa);
        h->launches_all += 1;
    }
    GarnerJob cj{};
    cj.src = XSRC_COLPAIR; cj.index = -1; cj.count = 2 * (h->m + 1); cj.check_status = 1;
    cj.meta = XMeta{x->cr_sign, x->cr_top, x->cr_mant};
    XRatioArgs ra{xgrid_of(h), x->st, cj.meta, pconst_of(h), x->cand, x->adaptive ? x->wcol : nullptr, x->xbasis, x->nbeta, x->nbeta_ld};
    const bool fuse_col = small && staged && cj.count <= 128 && (long long)x->P * (h->m + 1) <= 8192;
    if (fuse_col) {
        if (x->P <= 32) k_xratio_fused<1><<<1, kLaneThreads, 0, h->stream>>>(ra, cj);
        else k_xratio_fused<2><<<1, kLaneThreads, 0, h->stream>>>(ra, cj);
        h->launches_all += 1;
    } else {
        TRY(launch_garner<8>(h, cj));
        if (x->spt == 1) k_xratio<1><<<1, kLaneThreads, 0, h->stream>>>(ra);
        else k_xratio<kMaxSPT><<<1, kLaneThreads, 0, h->stream>>>(ra);
        h->launches_all += 1;
    }
    if (staged) *staged = fuse_col;
    CU(cudaGetLastError());
    return LPS_OK;
}

static int xlaunch_apply(lps_handle* h, int commit, bool staged = false) {
    ExactData* x = h->x;
    const XGrid g = xgrid_of(h);
    if (!staged) {
        k_xstage<<<dim3(std::max(1, std::min((h->m + 256) / 256, 64)), x->P), 256, 0, h->stream>>>(g, x->st, pconst_of(h), x->nbeta, x->nbeta_ld);
        h->launches_all++;
    }
    XUpdateArgs a;
    a.g = g; a.st = x->st; a.alpha_m = x->alpha_m; a.nbeta = x->nbeta; a.nbeta_ld = x->nbeta_ld;
    a.rows_per_cta = x->rows_per_cta; a.basis = h->basis; a.marks = h->use_marks ? h->marks : nullptr;
    a.trace = h->trace; a.trace_cap = h->trace_cap; a.alive = x->alive; a.xbasis = x->adaptive ? x->xbasis : nullptr; a.commit = commit;
    dim3 grid((unsigned)((x->ld + x->chunk - 1) / x->chunk), (unsigned)((h->m + 1 + x->rows_per_cta - 1) / x->rows_per_cta), (unsigned)x->P);
    const bool timed = h->prof && h->ev_used + 2 <= h->ev.size();
    if (timed) CU(cudaEventRecord(h->ev[h->ev_used++], h->stream));
    if (x->chunk == 2048) {
        if (x->stream_hint) k_xupdate<2048, true><<<grid, kXThreads, 0, h->stream>>>(a);
        else k_xupdate<2048, false><<<grid, kXThreads, 0, h->stream>>>(a);
    } else {
        if (x->stream_hint) k_xupdate<1024, true><<<grid, kXThreads, 0, h->stream>>>(a);
        else k_xupdate<1024, false><<<grid, kXThreads, 0, h->stream>>>(a);
    }
    if (timed) CU(cudaEventRecord(h->ev[h->ev_used++], h->stream));
    h->launches_update++;
    h->launches_all += 1;
    CU(cudaGetLastError());
    return LPS_OK;
}

// ------------------------------------------------------------------------------------------------
// hooks behind the mode-independent entry points
// ------------------------------------------------------------------------------------------------
int lpsx_pivot(lps_handle* h, int r, int c) {
    ExactData* x = h->x;
    if (r < 0 || r >= h->m || c < 0 || c >= h->n) return fail(LPS_E_INVALID, "pivot (%d,%d) out of range", r, c);
    TRY(xread_state(h));
    x->st_host->status = LPS_OK; x->st_host->dry_run = 0; x->st_host->bad_prime = -1;
    TRY(xwrite_state(h));
    for (int attempt = 0;; ++attempt) {
        const int* wc = x->adaptive ? x->wcol : nullptr;
        if (x->spt == 1) k_xforced<1><<<1, kLaneThreads, 0, h->stream>>>(xgrid_of(h), x->st, pconst_of(h), r, c, wc, x->xbasis);
        else k_xforced<kMaxSPT><<<1, kLaneThreads, 0, h->stream>>>(xgrid_of(h), x->st, pconst_of(h), r, c, wc, x->xbasis);
        h->launches_all++;
        CU(cudaGetLastError());
        TRY(xlaunch_apply(h, 0));
        TRY(xread_state(h));
        if (x->st_host->status != LPS_NEED_PRIMES || attempt > 0) break;
        const int grc = grow_working(h);                 // nothing was modified: extend the working set, try again
        if (grc < 0) return grc;
        if (grc == LPS_NEED_PRIMES || grc == LPS_BAD_PRIME) break;
    }
    const int st = x->st_host->status;
    return (st == LPS_ZERO_PIVOT || st == LPS_BAD_PRIME || st == LPS_NEED_PRIMES) ? st : LPS_OK;
}

int lpsx_find_pivot(lps_handle* h, int rule, int* r, int* c) {
    ExactData* x = h->x;
    TRY(xread_state(h));
    XState keep = *x->st_host;
    x->st_host->status = LPS_OK; x->st_host->dry_run = 1; x->st_host->rule = rule;
    x->st_host->last_r = x->st_host->last_c = -1;
    TRY(xwrite_state(h));
    TRY(xlaunch_select(h));
    TRY(xread_state(h));
    const XState o = *x->st_host;
    keep.status = LPS_OK;
    *x->st_host = keep;
    TRY(xwrite_state(h));
    *r = o.last_r; *c = o.last_c;
    if (o.status == LPS_OPTIMAL) return LPS_OPTIMAL;
    if (o.status == LPS_UNBOUNDED) return LPS_UNBOUNDED;
    return LPS_OK;
}

int lpsx_is_canonical(lps_handle* h, int* bcols, int* canonical) {
    ExactData* x = h->x;
    const XGrid g = xgrid_of(h);
    GarnerJob job{};
    job.src = XSRC_COL; job.index = h->n; job.count = h->m + 1; job.check_status = 0;
    job.meta = XMeta{x->rd_sign, x->rd_top, x->rd_mant};
    TRY(launch_garner<8>(h, job));
    int* flag = h->scratch_i + h->n;
    CU(cudaMemsetAsync(flag, 0, sizeof(int), h->stream));
    k_xany_neg<<<std::max(1, std::min((h->m + 255) / 256, 592)), 256, 0, h->stream>>>(x->rd_sign, h->m, x->st, flag);
    int neg = 0;
    CU(cudaMemcpyAsync(&neg, flag, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    h->launches_all++;
    if (neg) { *canonical = 0; return LPS_OK; }                  // bcols untouched (tableau.py:476-477)
    k_xone<<<dim3(std::max(1, std::min((h->m + 256) / 256, 64)), x->P), 256, 0, h->stream>>>(g, x->Dres, x->sigma, x->sstride, x->alive, x->one);
    k_xcanon_a<<<(h->n + 127) / 128, 128, 0, h->stream>>>(g, x->one, h->scratch_i);
    if (x->P > 1) k_xcanon_b<<<dim3((h->n + 127) / 128, x->P - 1), 128, 0, h->stream>>>(g, x->one, x->sstride, h->scratch_i);
    DevBuf b_bcols;
    CU(b_bcols.alloc((size_t)h->m * sizeof(int)));
    int* d_bcols = b_bcols.as<int>();
    CU(cudaMemsetAsync(d_bcols, 0x7f, (size_t)h->m * sizeof(int), h->stream));
    k_canon_rows<<<std::max(1, std::min((h->n + 255) / 256, 592)), 256, 0, h->stream>>>(h->scratch_i, nullptr, h->n, d_bcols);
    h->launches_all += 4;
    std::vector<int> tmp(h->m);
    CU(cudaMemcpyAsync(tmp.data(), d_bcols, (size_t)h->m * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    int all = 1;
    for (int i = 0; i < h->m; ++i) {
        const int j = (tmp[i] == 0x7f7f7f7f) ? -1 : tmp[i];
        if (j < 0) all = 0;
        if (bcols) bcols[i] = j;
    }
    *canonical = all;
    return LPS_OK;
}

// Tableau.isInfeasible (tableau.py:510-514): batches of rows reconstructed on the device, one warp per row on the signs.
int lpsx_is_infeasible(lps_handle* h, int* infeasible) {
    ExactData* x = h->x;
    if (!x->R) return fail(LPS_E_STATE, "upload first");
    const int m = h->m, n = h->n;
    const long long per = (long long)n + 1;
    const int batch = (int)std::max<long long>(1, std::min<long long>(m, (1LL << 22) / per));
    DevBuf b_sign, b_top, b_mant, b_flag;
    CU(b_sign.alloc((size_t)batch * per));
    CU(b_top.alloc((size_t)batch * per * sizeof(int)));
    CU(b_mant.alloc((size_t)batch * per * sizeof(double)));
    CU(b_flag.alloc(sizeof(int)));
    CU(cudaMemsetAsync(b_flag.p, 0, sizeof(int), h->stream));
    for (int i0 = 0; i0 < m; i0 += batch) {
        const int nr = std::min(batch, m - i0);
        GarnerJob job{};
        job.src = XSRC_ROWS; job.index = i0; job.count = (int)(nr * per); job.check_status = 0;
        job.meta = XMeta{b_sign.as<signed char>(), b_top.as<int>(), b_mant.as<double>()};
        TRY(launch_garner<8>(h, job));
        k_xrow_infeasible<<<std::max(1, std::min((nr + 7) / 8, 148 * 8)), 256, 0, h->stream>>>(b_sign.as<signed char>(), nr, n, x->st, b_flag.as<int>());
        h->launches_all++;
        CU(cudaGetLastError());
    }
    CU(cudaMemcpyAsync(infeasible, b_flag.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return LPS_OK;
}

static int xensure_burst_graph(lps_handle* h, long long burst) {
    ExactData* x = h->x;
    unsigned long long sig = 1469598103934665603ull;
    auto mix = [&](unsigned long long v) { sig = (sig ^ v) * 1099511628211ull; };
    mix((unsigned long long)(uintptr_t)h->trace); mix((unsigned long long)h->trace_cap); mix(h->use_marks ? 1 : 0);
    mix((unsigned long long)(uintptr_t)x->R); mix((unsigned long long)x->P); mix((unsigned long long)x->lane_threads); mix((unsigned long long)x->spt);
    mix((unsigned long long)x->chunk); mix((unsigned long long)x->rows_per_cta); mix(x->stream_hint ? 1 : 0);
    mix(x->adaptive ? 1 : 0); mix((unsigned long long)burst);
    if (h->gexec && h->gsig == sig && h->gburst == burst) return LPS_OK;
    if (h->gexec) { cudaGraphExecDestroy(h->gexec); h->gexec = nullptr; }
    const long long la = h->launches_all, lu = h->launches_update;
    cudaGraph_t graph = nullptr;
    CU(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
    int rc = LPS_OK;
    for (long long k = 0; k < burst && rc == LPS_OK; ++k) {
        bool staged = false;
        rc = xlaunch_select(h, &staged);
        if (rc == LPS_OK) rc = xlaunch_apply(h, 1, staged);
    }
    cudaError_t e = cudaStreamEndCapture(h->stream, &graph);
    x->graph_launches = std::max<long long>(1, (h->launches_all - la) / std::max<long long>(1, burst));
    h->launches_all = la; h->launches_update = lu;               // counted when the graph is launched
    if (rc != LPS_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (e != cudaSuccess) return fail(LPS_E_CUDA, "stream capture failed: %s", cudaGetErrorString(e));
    e = cudaGraphInstantiate(&h->gexec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) { h->gexec = nullptr; return fail(LPS_E_CUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(e)); }
    h->gsig = sig; h->gburst = burst;
    return LPS_OK;
}

int lpsx_solve(lps_handle* h, long long max_pivots, int* basis, int* marks, int* trace, long long trace_cap, long long* npiv) {
    ExactData* x = h->x;
    CU(cudaMemcpyAsync(h->basis, basis, (size_t)h->m * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    h->use_marks = marks != nullptr;
    if (marks) CU(cudaMemcpyAsync(h->marks, marks, (size_t)h->n * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    if (!trace) trace_cap = 0;
    if (trace_cap > h->trace_cap) {
        if (h->trace) CU(cudaFree(h->trace));
        h->trace = nullptr; h->trace_cap = 0;
        CU(cudaMalloc(&h->trace, (size_t)trace_cap * 3 * sizeof(int)));
        h->trace_cap = trace_cap;
    }
    TRY(xread_state(h));
    XState& s = *x->st_host;
    s.status = LPS_OK; s.npiv = 0; s.max_pivots = max_pivots; s.dry_run = 0; s.obj_increased = 0; s.bad_prime = -1;
    if (!x->resume) { s.rule = LPS_RULE_DANTZIG; s.stuck = 0; s.moved = 0; s.pending_zcheck = 0; }
    x->resume = false;
    TRY(xwrite_state(h));
    const double est_us = std::max(20.0, (double)x->plane * x->P * 8.0 / 5.0e6);
    const long long burst = (long long)std::max(2.0, std::min(128.0, 2000.0 / est_us));
    long long enqueued = 0;
    // Six launches per pivot: below a few hundred microseconds of device work per pivot the host is the limit, so
    // the burst is captured once as a CUDA graph (again after every change of the working prime set) and replayed;
    // pairs behind optimality, the pivot prefix or LPS_NEED_PRIMES are no-ops on the device.
    const bool graphed = !h->prof && est_us < 400.0;
    const long long gburst = std::min<long long>(burst, 16);   // short graphs: capture stays cheap next to a 30-pivot solve
    for (;;) {
        if (graphed) {
            TRY(xensure_burst_graph(h, gburst));
            for (long long k = 0; k < burst; k += gburst) CU(cudaGraphLaunch(h->gexec, h->stream));
            const long long done = (burst + gburst - 1) / gburst * gburst;
            h->launches_all += x->graph_launches * done;
            h->launches_update += done;
            enqueued += done;
        } else {
            const long long room = max_pivots >= 0 ? max_pivots - enqueued : burst;
            const long long now = std::min(burst, room);
            for (long long k = 0; k < now; ++k) {
                bool staged = false;
                TRY(xlaunch_select(h, &staged));
                TRY(xlaunch_apply(h, 1, staged));
            }
            enqueued += now;
            if (now == 0) TRY(xlaunch_select(h));
        }
        TRY(xread_state(h));
        if (x->st_host->status == LPS_NEED_PRIMES) {     // the pending pivot needs a larger working prime set
            enqueued = x->st_host->npiv;                 // the launches behind it were no-ops
            const int grc = grow_working(h);
            if (grc < 0) return grc;
            if (grc == LPS_NEED_PRIMES || grc == LPS_BAD_PRIME) break;   // reported: the host re-encodes and resumes
            continue;
        }
        if (x->st_host->status != LPS_OK) break;
    }
    const long long k = x->st_host->npiv;
    if (npiv) *npiv = k;
    CU(cudaMemcpyAsync(basis, h->basis, (size_t)h->m * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    if (marks) CU(cudaMemcpyAsync(marks, h->marks, (size_t)h->n * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    if (trace && trace_cap > 0) {
        const long long cnt = std::min(k, std::min(trace_cap, h->trace_cap));
        if (cnt > 0) CU(cudaMemcpyAsync(trace, h->trace, (size_t)cnt * 3 * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    }
    CU(cudaStreamSynchronize(h->stream));
    return x->st_host->status;
}
