// Batched exact-mode node LPs for branch and bound: one LP per thread-block CLUSTER, the whole
// multi-modular tableau resident in (distributed) shared memory, the complete `Simplex(tab); solve()`
// of the reference -- phase 1 included -- inside ONE kernel launch.
//
// The LP's fraction-free integer tableau N (lps_exact.cuh) is held as residues modulo P 31-bit primes.
// A cluster of C CTAs owns one LP and shards it BY ROWS, like the multi-GPU path shards a large tableau:
// CTA q keeps constraint rows [q*mloc, (q+1)*mloc) plus its own replica of the objective row, for ALL
// primes, in its shared memory (C = 1 when the whole LP fits one CTA: "one node LP per CTA").  So every
// number a decision needs has all its residues in ONE CTA: reconstruction (Garner, one THREAD per number,
// the digit vector in registers) is local and is done once per cluster, not once per CTA.  Per pivot the
// CTAs exchange, through distributed shared memory,
//   1. their best entering candidate of a slice of the objective row (index + magnitude digits),
//   2. their best leaving row (approximate ratio key + the residues an exact comparison needs),
//   3. the pivot row (owner CTA -> every peer),
// each followed by one cluster barrier; the rank-1 update then touches local rows only.
//
// Reference semantics reproduced (file:line under /root/reference/lpsol):
//   simplex.py:36-108   Simplex._find_bfs: negate rows with b_i < 0, isCanonical (tableau.py:466-496),
//                       artificial columns for the rows without a basic column, solve, infeasibility
//                       test, driving basic artificials out (pivot may be negative), objective restore
//   simplex.py:110-148  Simplex.solve: Dantzig pricing with the stale-objective stuck counter, then Bland
//   simplex.py:262-267 / :228-232 / :272-279  pricing and ratio test tie-breaks (lowest index)
//   tableau.py:295-308  Tableau.pivot, here N_ij <- (p N_ij - N_ic N_rj) / D, D <- p  (SURVEY 8a note)
#pragma once
#include <cooperative_groups.h>

#include "lps_exact.cuh"

namespace lps {
namespace cg = cooperative_groups;

constexpr int kBMaxThreads = 512;   // 512 threads when shared memory leaves one CTA per SM, else 256
constexpr int kBMaxCluster = 8;
constexpr int kBMaxPrimes = 16;
#ifndef LPS_BATCH_MINB
#define LPS_BATCH_MINB 1          // x 512 threads: the register budget (128 per thread) also fits 2 CTAs of 256
#endif

struct BatchParams {
    int count, m, n, art_cap;        // LPs, constraint rows, variables, artificial columns reserved per LP
    int C, mloc, cw;                 // CTAs per cluster, constraint rows per CTA, objective columns priced per CTA
    int ld;                          // plane row stride (multiple of 4, >= n + art_cap + 1)
    int tb;                          // slots of the exact tie-break scratch: max(mloc, C)
    long long max_pivots;            // safety cap per LP (< 0: none)
    long long stuck_limit;           // < 0: the reference's m + n (simplex.py:123); tests lower it to reach the Bland loop
    const PrimeQ* pr;                // [P] ascending primes
    const uint32_t* invtab;          // [P][P], Montgomery form of p_j^-1 mod p_k (j < k)
    const uint32_t* half;            // [P]
    const double* wmant;             // [P+1]
    const int* wexp;                 // [P+1]
    const int32_t* grids;            // [count][(m+1)][(n+1)] device layout (row m = costs, column n = rhs)
    int* status;                     // [count]
    int* npiv;                       // [count][4]: phase-1 pivots, drive-out pivots, phase-2 pivots, artificial columns
    int* basis;                      // [count][m]
    signed char* sign;               // [count][m+2]: rhs rows 0..m-1, the z cell, D
    uint32_t* digits;                // [count][P][m+2] mixed-radix digits of the magnitudes
    int* trace;                      // [count][trace_cap][3] or NULL
    int trace_cap;
    long long* timing;               // NULL, or [count][16] cycles of CTA 0 per phase (profiling aid, see b_tick)
    int wcheck;                      // 1: the prime set is shorter than the static bound asks for -- refuse (LPS_NEED_PRIMES)
                                     // any pivot after which the Hadamard product of the basis columns might not fit
};

// Shared-memory carve-up, identical on host (sizing) and device.
struct BatchLayout {
    size_t planes, prow, objdig, nbeta, tbc, tbb, inv, half, dres, dinv, alpha, piv, cand_price, cand_ratio, obj_top,
        col_top, rhs_top, tb_row, cand, basis, rowsign, wcol, wexp, wmant, rcp, col_mant, rhs_mant, pr, obj_sign, col_sign, rhs_sign,
        total;
};
__host__ __device__ inline int batch_price_words(int P) { return P + 2; }          // idx, top, digits[P]
__host__ __device__ inline int batch_ratio_words(int P) { return 2 * P + 8; }      // row, col, -, -, ratio(2), sa, sb, N_rc[P], N_rn[P]
__host__ __device__ inline BatchLayout batch_layout(int m, int P, int C, int mloc, int cw, int ld, int tb) {
    BatchLayout L;
    size_t o = 0;
    auto take = [&](size_t bytes, size_t align) { o = (o + align - 1) / align * align; const size_t at = o; o += bytes; return at; };
    const int mr = (mloc + 1 + 3) / 4 * 4;
    L.planes = take((size_t)P * (mloc + 1) * ld * 4, 16);
    L.prow = take((size_t)P * ld * 4, 16);
    L.objdig = take((size_t)P * cw * 4, 16);
    L.nbeta = take((size_t)P * mr * 4, 16);
    L.tbc = take((size_t)P * tb * 4, 16);
    L.tbb = take((size_t)P * tb * 4, 16);
    L.inv = take((size_t)P * P * 4, 16);
    L.half = take((size_t)P * 4, 4);
    L.dres = take((size_t)P * 4, 4);
    L.dinv = take((size_t)P * 4, 4);
    L.alpha = take((size_t)P * 4, 4);
    L.piv = take((size_t)P * 4, 4);
    L.cand_price = take((size_t)C * batch_price_words(P) * 4, 8);
    L.cand_ratio = take((size_t)C * batch_ratio_words(P) * 4, 8);
    L.obj_top = take((size_t)cw * 4, 4);
    L.col_top = take((size_t)mr * 4, 4);
    L.rhs_top = take((size_t)mr * 4, 4);
    L.tb_row = take((size_t)tb * 4, 4);
    L.cand = take((size_t)(mr > tb ? mr : tb) * 4, 4);
    L.basis = take((size_t)m * 4, 4);
    L.rowsign = take((size_t)m * 4, 4);
    L.wcol = take((size_t)ld * 4, 4);
    L.wexp = take((size_t)(P + 1) * 4, 4);
    L.wmant = take((size_t)(P + 1) * 8, 8);
    L.rcp = take((size_t)(P + 1) * 8, 8);
    L.col_mant = take((size_t)mr * 8, 8);
    L.rhs_mant = take((size_t)mr * 8, 8);
    L.pr = take((size_t)P * sizeof(PrimeQ), 16);
    L.obj_sign = take((size_t)cw, 4);
    L.col_sign = take((size_t)mr, 4);
    L.rhs_sign = take((size_t)mr, 4);
    L.total = (o + 15) / 16 * 16;
    return L;
}

// ------------------------------------------------------------------------------------------------
// Garner, lane-parallel: a group of LP consecutive lanes of a warp holds one number, lane g*LP + k its
// residue modulo prime k (lanes with k >= P idle), so a warp reconstructs 32/LP numbers at once and the
// dependent chain is P-1 steps of (shuffle digit j from lane j, one modular multiply).  The primes ascend,
// so digit j (< p_j) is already reduced modulo every later prime.  Must be called by whole warps.
// On return t = digit k of |x| in the mixed radix (p_0, p_1, ...); sign in {-1,0,1} and top (index of the
// leading nonzero digit, -1 for 0) are the same in every lane of the group.
// ------------------------------------------------------------------------------------------------
template <int P>
struct LaneCfg {
    static constexpr int LP = P <= 2 ? 2 : (P <= 4 ? 4 : (P <= 8 ? 8 : 16));   // lanes per number
    static constexpr int G = 32 / LP;                                           // numbers per warp
};
template <int P>
__device__ __forceinline__ void garner_lane(uint32_t& t, int k, uint32_t q, uint32_t qinv, const uint32_t* sInv,
                                            uint32_t half_k, int& sign, int& top) {
    constexpr int LP = LaneCfg<P>::LP;
    const int g0 = (threadIdx.x & 31) & ~(LP - 1);
    const bool live = k < P;
    const uint32_t* invcol = sInv + (live ? k : 0);
#pragma unroll
    for (int j = 0; j < P - 1; ++j) {
        const uint32_t dj = __shfl_sync(0xffffffffu, t, g0 + j);
        if (live && k > j) t = mmul(msub(t, dj, q), invcol[j * P], q, qinv);
    }
    const unsigned gm = (1u << LP) - 1u;
    const unsigned neq = (__ballot_sync(0xffffffffu, live && t != half_k) >> g0) & gm;
    const unsigned gt = (__ballot_sync(0xffffffffu, live && t > half_k) >> g0) & gm;
    const unsigned nz = (__ballot_sync(0xffffffffu, live && t != 0u) >> g0) & gm;
    const int hi = neq ? 31 - __clz(neq) : -1;
    const bool neg = hi >= 0 && ((gt >> hi) & 1u);             // x > (M-1)/2: a negative number
    const int lo = nz ? __ffs(nz) - 1 : -1;
    uint32_t a = live ? t : 0u;
    if (live && neg) a = k < lo ? 0u : (k == lo ? q - t : q - 1u - t);
    t = a;
    const unsigned nza = (__ballot_sync(0xffffffffu, a != 0u) >> g0) & gm;
    top = nza ? 31 - __clz(nza) : -1;
    sign = nz == 0u ? 0 : (neg ? -1 : 1);
}
// |x| as a double from its three leading digits times W_top (relative error below 2^-50); all lanes of the group
// take part, every lane returns the value.  wval[t] = p_0 ... p_(t-1) and rcp[t] = 1 / p_t as doubles.
template <int P>
__device__ __forceinline__ double value_lane(uint32_t t, int top, const double* rcp, const double* wval) {
    constexpr int LP = LaneCfg<P>::LP;
    const int g0 = (threadIdx.x & 31) & ~(LP - 1);
    const int tp = max(top, 0);
    const uint32_t d0 = __shfl_sync(0xffffffffu, t, g0 + tp);
    const uint32_t d1 = __shfl_sync(0xffffffffu, t, g0 + max(tp - 1, 0));
    const uint32_t d2 = __shfl_sync(0xffffffffu, t, g0 + max(tp - 2, 0));
    if (top < 0) return 0.0;
    double xv = (double)d0;
    if (top >= 1) {
        double y = (double)d1;
        if (top >= 2) y += (double)d2 * rcp[top - 2];
        xv += y * rcp[top - 1];
    }
    return xv * wval[top];
}

// Montgomery reduction with a branch-free final step: min(t, t - q) as unsigned numbers
__device__ __forceinline__ uint32_t redc_min(uint64_t T, uint32_t q, uint32_t qinv) {
    const uint32_t mlo = (uint32_t)T * qinv;
    const uint32_t t = (uint32_t)((T + (uint64_t)mlo * q) >> 32);
    return min(t, t - q);
}

extern __shared__ __align__(16) unsigned char lps_bsmem[];   // BCtx at offset 0, then the carve-up of batch_layout

struct BState {                  // one copy per CTA, kept identical across the cluster
    int status;
    int phase;                   // 0 phase 1, 1 driving artificials out, 2 phase 2
    int ncur, rc;                // variable columns in play, index of the rhs column
    int K;                       // artificial columns of this LP
    int rule, moved, pending;
    long long stuck;
    long long npiv_total;
    int npiv[3];
    int sgnD;
    int flag, count, low, low2;  // scratch
    // capacity of a prime set shorter than the static bound (BatchParams::wcheck): every entry is, up to sign, a minor
    // of the LP's extended matrix over the basis columns plus at most two more, so log2|entry| <= wsum + 2 wmax in
    // 1/256 bits, wsum = sum of wcol[basis[i]], wcol[j] >= 256 log2 of column j's 2-norm (cost rows of both phases)
    long long wsum, wlimit;
};

// Everything a device function of the kernel needs, built once per CTA in shared memory.
struct BCtx {
    BatchParams a;
    int qr, lp, r0, mq;          // rank in the cluster, LP index, first local row, local constraint rows
    int mr;                      // mloc + 1 rounded up to 4
    size_t plane_sz;             // (mloc + 1) * ld
    // byte offsets into the dynamic shared memory; the accessors rebuild pointers the compiler knows to be
    // in the shared window (LDS/STS with 32-bit addresses instead of generic loads)
    unsigned o_planes, o_prow, o_objdig, o_nbeta, o_tbc, o_tbb, o_inv, o_half, o_Dres, o_dinv_m, o_alpha_m, o_piv;
    unsigned o_cand_price, o_cand_ratio, o_obj_top, o_col_top, o_rhs_top, o_tb_row, o_cand, o_basis, o_rowsign, o_wcol, o_wexp;
    unsigned o_wmant, o_rcp, o_col_mant, o_rhs_mant;
    unsigned o_pr;
    unsigned o_obj_sign, o_col_sign, o_rhs_sign;
    __device__ __forceinline__ uint32_t* planes() const { return reinterpret_cast<uint32_t*>(lps_bsmem + o_planes); }
    __device__ __forceinline__ uint32_t* prow() const { return reinterpret_cast<uint32_t*>(lps_bsmem + o_prow); }
    __device__ __forceinline__ uint32_t* objdig() const { return reinterpret_cast<uint32_t*>(lps_bsmem + o_objdig); }
    __device__ __forceinline__ uint32_t* nbeta() const { return reinterpret_cast<uint32_t*>(lps_bsmem + o_nbeta); }
    __device__ __forceinline__ uint32_t* tbc() const { return reinterpret_cast<uint32_t*>(lps_bsmem + o_tbc); }
    __device__ __forceinline__ uint32_t* tbb() const { return reinterpret_cast<uint32_t*>(lps_bsmem + o_tbb); }
    __device__ __forceinline__ uint32_t* inv() const { return reinterpret_cast<uint32_t*>(lps_bsmem + o_inv); }
    __device__ __forceinline__ uint32_t* half() const { return reinterpret_cast<uint32_t*>(lps_bsmem + o_half); }
    __device__ __forceinline__ uint32_t* Dres() const { return reinterpret_cast<uint32_t*>(lps_bsmem + o_Dres); }
    __device__ __forceinline__ uint32_t* dinv_m() const { return reinterpret_cast<uint32_t*>(lps_bsmem + o_dinv_m); }
    __device__ __forceinline__ uint32_t* alpha_m() const { return reinterpret_cast<uint32_t*>(lps_bsmem + o_alpha_m); }
    __device__ __forceinline__ uint32_t* piv() const { return reinterpret_cast<uint32_t*>(lps_bsmem + o_piv); }
    __device__ __forceinline__ int* cand_price() const { return reinterpret_cast<int*>(lps_bsmem + o_cand_price); }
    __device__ __forceinline__ int* cand_ratio() const { return reinterpret_cast<int*>(lps_bsmem + o_cand_ratio); }
    __device__ __forceinline__ int* obj_top() const { return reinterpret_cast<int*>(lps_bsmem + o_obj_top); }
    __device__ __forceinline__ int* col_top() const { return reinterpret_cast<int*>(lps_bsmem + o_col_top); }
    __device__ __forceinline__ int* rhs_top() const { return reinterpret_cast<int*>(lps_bsmem + o_rhs_top); }
    __device__ __forceinline__ int* tb_row() const { return reinterpret_cast<int*>(lps_bsmem + o_tb_row); }
    __device__ __forceinline__ int* cand() const { return reinterpret_cast<int*>(lps_bsmem + o_cand); }
    __device__ __forceinline__ int* basis() const { return reinterpret_cast<int*>(lps_bsmem + o_basis); }
    __device__ __forceinline__ int* rowsign() const { return reinterpret_cast<int*>(lps_bsmem + o_rowsign); }
    __device__ __forceinline__ int* wcol() const { return reinterpret_cast<int*>(lps_bsmem + o_wcol); }
    __device__ __forceinline__ int* wexp() const { return reinterpret_cast<int*>(lps_bsmem + o_wexp); }
    __device__ __forceinline__ double* wmant() const { return reinterpret_cast<double*>(lps_bsmem + o_wmant); }
    __device__ __forceinline__ double* rcp() const { return reinterpret_cast<double*>(lps_bsmem + o_rcp); }
    __device__ __forceinline__ double* col_mant() const { return reinterpret_cast<double*>(lps_bsmem + o_col_mant); }
    __device__ __forceinline__ double* rhs_mant() const { return reinterpret_cast<double*>(lps_bsmem + o_rhs_mant); }
    __device__ __forceinline__ PrimeQ* pr() const { return reinterpret_cast<PrimeQ*>(lps_bsmem + o_pr); }
    __device__ __forceinline__ signed char* obj_sign() const { return reinterpret_cast<signed char*>(lps_bsmem + o_obj_sign); }
    __device__ __forceinline__ signed char* col_sign() const { return reinterpret_cast<signed char*>(lps_bsmem + o_col_sign); }
    __device__ __forceinline__ signed char* rhs_sign() const { return reinterpret_cast<signed char*>(lps_bsmem + o_rhs_sign); }
    const int32_t* G;
    BState S;
    int s_red[32];
    int s_cls[32], s_e[32];
    double s_f[32], s_g[32];
    long long tacc[16], tlast;   // phase clocks (thread 0), only when a.timing != NULL
};

__device__ __forceinline__ BCtx& b_ctx() { return *reinterpret_cast<BCtx*>(lps_bsmem); }
__host__ __device__ inline size_t batch_ctx_bytes() { return (sizeof(BCtx) + 15) / 16 * 16; }

// phase clock: 0 price local, 1 barrier 1, 2 price fold, 3 ratio Garner, 4 ratio local pick + publish, 5 barrier 2,
// 6 ratio global pick, 7 pivot constants + row broadcast, 8 barrier 3, 9 update, 10 inverse of the new D,
// 11 loop bookkeeping, 12 set-up + encode, 13 phase-1 transitions, 14 results
__device__ __noinline__ void b_tick_record(int slot) {
    BCtx& x = b_ctx();
    const long long t = clock64();
    x.tacc[slot] += t - x.tlast;
    x.tlast = t;
}
__device__ __forceinline__ void b_tick(int slot) {
    if (threadIdx.x == 0 && b_ctx().a.timing) b_tick_record(slot);
}
__device__ __forceinline__ void b_csync() {
    const BCtx& x = b_ctx();
    if (x.a.C > 1) cg::this_cluster().sync(); else __syncthreads();
}
template <typename T>
__device__ __forceinline__ T* b_peer(T* p, int rank) {
    const BCtx& x = b_ctx();
    return (x.a.C > 1 && rank != x.qr) ? cg::this_cluster().map_shared_rank(p, rank) : p;
}
// ------------------------------------------------------------------------------------------------
// exact minimum of N_xn / N_xc over `count` candidates whose residues sit in tbc / tbb (slot-major per
// prime) and whose global rows are tb_row[]: rounds of "every candidate against the current best", each
// comparison the sign of (N_kn N_rc - N_rn N_kc) / D -- a tableau entry one pivot away, so it fits the
// prime set.  Ties go to the lowest global row (simplex.py:277).  Returns the winning slot.
// ------------------------------------------------------------------------------------------------
template <int P>
__device__ __noinline__ int b_exact_min(int count) {
    BCtx& x = b_ctx();
    const int tb = x.a.tb, tid = threadIdx.x, sgnD = x.S.sgnD;
    if (count <= 1) return 0;
    constexpr int LP = LaneCfg<P>::LP, G = LaneCfg<P>::G;
    const int lane = tid & 31, warp = tid >> 5, nwarps = (int)blockDim.x >> 5;
    const int k = lane & (LP - 1), sub = lane / LP;
    const bool live = k < P;
    const PrimeQ pk = x.pr()[live ? k : 0];
    const uint32_t half_k = x.half()[live ? k : 0], dinv_k = x.dinv_m()[live ? k : 0];
    const uint32_t* tbc = x.tbc();
    const uint32_t* tbb = x.tbb();
    const int* tb_row = x.tb_row();
    __syncthreads();
    if (tid == 0) x.S.low = INT_MAX;
    __syncthreads();
    for (int e = tid; e < count; e += (int)blockDim.x) atomicMin(&x.S.low, tb_row[e]);     // deterministic start: lowest row
    __syncthreads();
    int rrow = x.S.low;
    for (;;) {
        __syncthreads();
        if (tid == 0) { x.S.low = INT_MAX; x.S.low2 = INT_MAX; }
        __syncthreads();
        int rs = 0;
        for (int e = 0; e < count; ++e)
            if (tb_row[e] == rrow) rs = e;
#pragma unroll 1
        for (int base = warp * G; base < count; base += nwarps * G) {      // whole warps: garner_lane is warp-collective
            const int e = base + sub;
            const bool valid = e < count && e != rs;
            uint32_t t = 0u;
            if (valid && live) {
                const uint32_t xx = mmul(tbb[k * tb + e], tbc[k * tb + rs], pk.q, pk.qinv);
                const uint32_t yy = mmul(tbb[k * tb + rs], tbc[k * tb + e], pk.q, pk.qinv);
                t = mmul(mmul(msub(xx, yy, pk.q), dinv_k, pk.q, pk.qinv), pk.r2, pk.q, pk.qinv);
            }
            int sg, tp;
            garner_lane<P>(t, k, pk.q, pk.qinv, x.inv(), half_k, sg, tp);
            sg *= sgnD;                                        // sign of ratio_e - ratio_r
            if (valid && k == 0) {
                if (sg < 0) atomicMin(&x.S.low, tb_row[e]);    // strictly smaller ratio
                else if (sg == 0) atomicMin(&x.S.low2, tb_row[e]);
            }
        }
        __syncthreads();
        const int lt = x.S.low, eq = x.S.low2;
        if (lt != INT_MAX) { rrow = lt; continue; }            // restart from a strictly smaller row
        rrow = min(rrow, eq);
        break;
    }
    int rs = 0;
    for (int e = 0; e < count; ++e)
        if (tb_row[e] == rrow) rs = e;
    __syncthreads();
    return rs;
}

// Dantzig tie on the leading digit: the candidate (local index) with the larger magnitude, the lower index when
// the magnitudes are equal (simplex.py:266 keeps the first of equal minima)
__device__ __noinline__ int b_deeper(int u, int v) {
    const BCtx& x = b_ctx();
    const int cmp = cmp_mag(x.obj_top(), x.objdig(), (size_t)x.a.cw, u, v);
    return cmp > 0 ? u : (cmp < 0 ? v : min(u, v));
}

// ------------------------------------------------------------------------------------------------
// pricing (simplex.py:262-267 / :228-232): this CTA reconstructs its slice of the objective row and
// publishes its best candidate to every peer; after the barrier all pick the same entering column.
// Returns -1 when no reduced cost is negative.
// ------------------------------------------------------------------------------------------------
template <int P>
__device__ __noinline__ int b_price(int rule) {
    BCtx& x = b_ctx();
    const int tid = threadIdx.x, T = (int)blockDim.x, lane = tid & 31, warp = tid >> 5, cw = x.a.cw, ld = x.a.ld, C = x.a.C;
    const int sgnD = x.S.sgnD, ncur = x.S.ncur;
    const int j0 = x.qr * cw, cnt = max(0, min(ncur - j0, cw));
    uint32_t* const objdig = x.objdig();
    int* const obj_top = x.obj_top();
    // Dantzig key of a candidate: (top + 1, leading digit) decides almost every comparison in registers; equal keys
    // fall back to the digit vectors in shared memory.  Bland key: lower index = larger key.
    unsigned long long key = 0ull;
    int best = -1;                                             // local index of this thread's candidate
    {
        constexpr int LP = LaneCfg<P>::LP, G = LaneCfg<P>::G;
        const int k = lane & (LP - 1), sub = lane / LP, nwarps = T >> 5;
        const bool live = k < P;
        const PrimeQ pk = x.pr()[live ? k : 0];
        const uint32_t half_k = x.half()[live ? k : 0];
        const uint32_t* obj = x.planes() + (size_t)(live ? k : 0) * x.plane_sz + (size_t)x.a.mloc * ld + j0;
        const uint32_t* inv = x.inv();
#pragma unroll 1
        for (int base = warp * G; base < cnt; base += nwarps * G) {   // whole warps: garner_lane is warp-collective
            const int e = base + sub;
            const bool valid = e < cnt;
            uint32_t t = (valid && live) ? obj[e] : 0u;
            int sg, tp;
            garner_lane<P>(t, k, pk.q, pk.qinv, inv, half_k, sg, tp);
            const uint32_t dtop = __shfl_sync(0xffffffffu, t, (lane & ~(LP - 1)) + max(tp, 0));
            const bool negc = valid && sg * sgnD < 0;          // a negative reduced cost: a candidate
            if (negc && live) objdig[k * cw + e] = t;
            if (negc && k == 0) obj_top[e] = tp;
            __syncwarp();
            if (negc && k == 0) {
                const unsigned long long ke = rule == LPS_RULE_BLAND ? (unsigned long long)(INT_MAX - e)
                                                                     : (((unsigned long long)(tp + 1) << 32) | dtop);
                if (best < 0 || ke > key) { key = ke; best = e; }      // later e only on a strictly larger key
                else if (ke == key && b_deeper(e, best) == e && e != best) best = e;
            }
        }
    }
    __syncwarp();
    auto pick = [&](unsigned long long ku, int u, unsigned long long kv, int v, unsigned long long& ko, int& o) {
        if (v < 0) { ko = ku; o = u; return; }
        if (u < 0) { ko = kv; o = v; return; }
        if (ku != kv) { if (ku > kv) { ko = ku; o = u; } else { ko = kv; o = v; } return; }
        ko = ku;
        o = b_deeper(u, v);                                    // same (top, leading digit): compare the digit vectors
    };
#pragma unroll 1
    for (int off = 16; off > 0; off >>= 1) {
        const unsigned long long ko = __shfl_xor_sync(0xffffffffu, key, off);
        const int bo = __shfl_xor_sync(0xffffffffu, best, off);
        pick(key, best, ko, bo, key, best);
    }
    unsigned long long* s_key = reinterpret_cast<unsigned long long*>(x.s_f);
    if (lane == 0) { s_key[warp] = key; x.s_red[warp] = best; }
    __syncthreads();
    if (warp == 0) {
        const int nw = T >> 5;
        unsigned long long k2 = lane < nw ? s_key[lane] : 0ull;
        int b2 = lane < nw ? x.s_red[lane] : -1;
#pragma unroll 1
        for (int off = 16; off > 0; off >>= 1) {
            const unsigned long long ko = __shfl_xor_sync(0xffffffffu, k2, off);
            const int bo = __shfl_xor_sync(0xffffffffu, b2, off);
            pick(k2, b2, ko, bo, k2, b2);
        }
        // publish {global index, top, digits} into slot qr of every CTA of the cluster
        const int words = batch_price_words(P);
#pragma unroll 1
        for (int peer = 0; peer < C; ++peer) {
            int* dst = b_peer(x.cand_price(), peer) + x.qr * words;
#pragma unroll 1
            for (int w = lane; w < words; w += 32) {
                int v;
                if (w == 0) v = b2 < 0 ? -1 : j0 + b2;
                else if (b2 < 0) v = 0;
                else if (w == 1) v = obj_top[b2];
                else v = (int)objdig[(w - 2) * cw + b2];
                dst[w] = v;
            }
        }
    }
    b_tick(0);
    b_csync();                                                 // barrier 1: candidates visible everywhere
    b_tick(1);
    const int words = batch_price_words(P);
    const int* cp = x.cand_price();
    int win = -1;
#pragma unroll 1
    for (int s = 0; s < C; ++s) {                              // identical fold on every thread of every CTA
        const int* cs = cp + s * words;
        if (cs[0] < 0) continue;
        if (win < 0) { win = s; continue; }
        const int* cwn = cp + win * words;
        if (rule == LPS_RULE_BLAND) { if (cs[0] < cwn[0]) win = s; continue; }
        int cmp = 0;
        if (cs[1] != cwn[1]) cmp = cs[1] > cwn[1] ? 1 : -1;
        else
            for (int k = cs[1]; k >= 0 && cmp == 0; --k) {
                const uint32_t da = (uint32_t)cs[2 + k], db = (uint32_t)cwn[2 + k];
                if (da != db) cmp = da > db ? 1 : -1;
            }
        if (cmp > 0 || (cmp == 0 && cs[0] < cwn[0])) win = s;
    }
    const int c = win < 0 ? -1 : cp[win * words];
    b_tick(2);
    return c;
}

// ------------------------------------------------------------------------------------------------
// ratio test (simplex.py:272-279) for entering column c.  Each CTA reconstructs N_ic and N_in of its rows; the
// ratio N_in / N_ic is first taken in floating point from the leading digits (relative error below 2^-44: three
// 31-bit digits of each number, products of primes held as doubles -- at most 2^496, inside the double range);
// every row within 2^-38 of the smallest approximation is a candidate and the candidates are compared EXACTLY.
// The local winner is published with the residues an exact comparison needs; after the barrier every CTA picks
// the same global leaving row the same way.  Returns the winning SLOT (= rank of the CTA that owns the row) or
// -1: unbounded.  The winner's record stays in cand_ratio[slot].
// ------------------------------------------------------------------------------------------------
// Approximate ratios are kept as (num, den) pairs of doubles, den > 0, and compared by cross-multiplication
// (no division on the path; |num|, den <= 2^496, so the products stay inside the double range).
struct Frac2 {
    double num, den;                                           // den == 0: no candidate
};
__device__ __forceinline__ bool frac_less(const Frac2& a, const Frac2& b) {   // a < b; "no candidate" is the largest
    if (a.den == 0.0) return false;
    if (b.den == 0.0) return true;
    return a.num * b.den < b.num * a.den;
}
__device__ __forceinline__ bool frac_close(const Frac2& v, const Frac2& best) {   // |v / best - 1| <= 2^-38, best.num != 0
    return fabs(v.num * best.den - best.num * v.den) <= 3.637978807091713e-12 * fabs(best.num) * v.den;
}

template <int P>
__device__ __noinline__ int b_ratio(int c) {
    BCtx& x = b_ctx();
    const int tid = threadIdx.x, T = (int)blockDim.x, lane = tid & 31, warp = tid >> 5, ld = x.a.ld, C = x.a.C, tb = x.a.tb;
    const int mq = x.mq, sgnD = x.S.sgnD, rc = x.S.rc, r0 = x.r0;
    const int words = batch_ratio_words(P);
    const size_t plane_sz = x.plane_sz;
    const uint32_t* const planes = x.planes();
    signed char* const col_sign = x.col_sign();
    signed char* const rhs_sign = x.rhs_sign();
    double* const col_val = x.col_mant();                      // |N_ic| as a double
    double* const rhs_val = x.rhs_mant();                      // |N_in| as a double
    int* const cand = x.cand();
    {
        constexpr int LP = LaneCfg<P>::LP, G = LaneCfg<P>::G;
        const int k = lane & (LP - 1), sub = lane / LP, nwarps = T >> 5;
        const bool live = k < P;
        const PrimeQ pk = x.pr()[live ? k : 0];
        const uint32_t half_k = x.half()[live ? k : 0];
        const uint32_t* pl = planes + (size_t)(live ? k : 0) * plane_sz;
        const uint32_t* inv = x.inv();
        const double* wval = x.wmant();
        const double* rcp = x.rcp();
#pragma unroll 1
        for (int base = warp * G; base < 2 * mq; base += nwarps * G) {   // whole warps: garner_lane is warp-collective
            const int e = base + sub;
            const bool valid = e < 2 * mq;
            const bool is_col = e < mq;
            const int il = valid ? (is_col ? e : e - mq) : 0;
            uint32_t t = (valid && live) ? pl[(size_t)il * ld + (is_col ? c : rc)] : 0u;
            int sg, tp;
            garner_lane<P>(t, k, pk.q, pk.qinv, inv, half_k, sg, tp);
            const double val = value_lane<P>(t, tp, rcp, wval);
            if (valid && k == 0) {
                (is_col ? col_sign : rhs_sign)[il] = (signed char)sg;
                (is_col ? col_val : rhs_val)[il] = val;
            }
        }
    }
    __syncthreads();
    b_tick(3);
    // approximate ratio of this thread's rows; rows with a_ic <= 0 are not eligible (simplex.py:273-275)
    auto frac_of = [&](int i) {
        Frac2 f{0.0, 0.0};
        if ((int)col_sign[i] * sgnD > 0) { f.num = (double)((int)rhs_sign[i] * (int)col_sign[i]) * rhs_val[i]; f.den = col_val[i]; }
        return f;
    };
    Frac2 best{0.0, 0.0};
#pragma unroll 1
    for (int i = tid; i < mq; i += T) {
        const Frac2 f = frac_of(i);
        if (frac_less(f, best)) best = f;
    }
#pragma unroll 1
    for (int off = 16; off > 0; off >>= 1) {
        Frac2 o;
        o.num = __shfl_xor_sync(0xffffffffu, best.num, off);
        o.den = __shfl_xor_sync(0xffffffffu, best.den, off);
        if (frac_less(o, best)) best = o;
    }
    double* s_num = x.s_f;
    if (lane == 0) { s_num[warp] = best.num; x.s_g[warp] = best.den; }
    if (tid == 0) { x.S.count = 0; x.S.low = INT_MAX; }
    __syncthreads();
    best = Frac2{s_num[0], x.s_g[0]};
#pragma unroll 1
    for (int w = 1; w < (T >> 5); ++w) {
        const Frac2 o{s_num[w], x.s_g[w]};
        if (frac_less(o, best)) best = o;
    }
    int il_win = -1;                                           // local winner (local row index)
    if (best.den != 0.0) {
        if (best.num == 0.0) {                                 // ratio exactly 0 (b_i == 0): lowest such row
#pragma unroll 1
            for (int i = tid; i < mq; i += T)
                if ((int)col_sign[i] * sgnD > 0 && rhs_sign[i] == 0) { atomicMin(&x.S.low, i); break; }
            __syncthreads();
            il_win = x.S.low;
        } else {
#pragma unroll 1
            for (int i = tid; i < mq; i += T) {
                const Frac2 f = frac_of(i);
                if (f.den != 0.0 && f.num != 0.0 && frac_close(f, best)) cand[atomicAdd(&x.S.count, 1)] = i;
            }
            __syncthreads();
            const int count = x.S.count;
            il_win = cand[0];
            if (count > 1) {                                   // near-ties: stage their residues, decide exactly
                uint32_t* tbc = x.tbc();
                uint32_t* tbb = x.tbb();
                int* tb_row = x.tb_row();
#pragma unroll 1
                for (int e = tid; e < count * P; e += T) {
                    const int s = e / P, k = e - s * P, il = cand[s];
                    const uint32_t* row = planes + (size_t)k * plane_sz + (size_t)il * ld;
                    tbc[k * tb + s] = row[c];
                    tbb[k * tb + s] = row[rc];
                    if (k == 0) tb_row[s] = r0 + il;
                }
                __syncthreads();
                il_win = cand[b_exact_min<P>(count)];
            }
        }
    }
    __syncthreads();
    // publish the local winner: {row, col, num (2 words), den (2 words), sign(N_rc), sign(N_rn), N_rc[P], N_rn[P]}
    if (warp == 0) {
        const Frac2 fw = il_win >= 0 ? frac_of(il_win) : Frac2{0.0, 0.0};
        const long long nb = __double_as_longlong(fw.num), db = __double_as_longlong(fw.den);
#pragma unroll 1
        for (int peer = 0; peer < C; ++peer) {
            int* dst = b_peer(x.cand_ratio(), peer) + x.qr * words;
#pragma unroll 1
            for (int w = lane; w < words; w += 32) {
                int v = 0;
                if (w == 0) v = il_win >= 0 ? r0 + il_win : -1;
                else if (il_win < 0) v = 0;
                else if (w == 1) v = c;
                else if (w == 2) v = (int)(nb & 0xffffffffll);
                else if (w == 3) v = (int)(nb >> 32);
                else if (w == 4) v = (int)(db & 0xffffffffll);
                else if (w == 5) v = (int)(db >> 32);
                else if (w == 6) v = (int)col_sign[il_win];
                else if (w == 7) v = (int)rhs_sign[il_win];
                else {
                    const int k = w - 8 < P ? w - 8 : w - 8 - P;
                    const uint32_t* row = planes + (size_t)k * plane_sz + (size_t)il_win * ld;
                    v = (int)(w - 8 < P ? row[c] : row[rc]);
                }
                dst[w] = v;
            }
        }
    }
    b_tick(4);
    b_csync();                                                 // barrier 2: leaving candidates visible everywhere
    b_tick(5);
    // global pick among the C slots, same rule: approximate minimum, exact among the slots close to it
    const int* cr = x.cand_ratio();
    auto slot_frac = [&](int s) {
        const int* cs = cr + s * words;
        Frac2 f{0.0, 0.0};
        if (cs[0] >= 0) {
            f.num = __longlong_as_double(((long long)cs[3] << 32) | (unsigned int)cs[2]);
            f.den = __longlong_as_double(((long long)cs[5] << 32) | (unsigned int)cs[4]);
        }
        return f;
    };
    Frac2 gbest{0.0, 0.0};
#pragma unroll 1
    for (int s = 0; s < C; ++s) {
        const Frac2 f = slot_frac(s);
        if (frac_less(f, gbest)) gbest = f;
    }
    if (gbest.den == 0.0) { b_tick(6); return -1; }
    int cnt = 0, first = -1;
    unsigned inmask = 0u;
#pragma unroll 1
    for (int s = 0; s < C; ++s) {
        const Frac2 f = slot_frac(s);
        const bool in = f.den != 0.0 && (gbest.num == 0.0 ? f.num == 0.0 : (f.num != 0.0 && frac_close(f, gbest)));
        if (in) { if (first < 0) first = s; ++cnt; inmask |= 1u << s; }
    }
    // exact zeros: the slots are in row order, so the first one holds the lowest row
    if (cnt == 1 || gbest.num == 0.0) { b_tick(6); return first; }
    __syncthreads();
    if (tid < C && ((inmask >> tid) & 1u)) {
        const int pos = __popc(inmask & ((1u << tid) - 1u));
        const int* cs = cr + tid * words;
        x.tb_row()[pos] = cs[0];
        cand[pos] = tid;
#pragma unroll 1
        for (int k2 = 0; k2 < P; ++k2) { x.tbc()[k2 * tb + pos] = (uint32_t)cs[8 + k2]; x.tbb()[k2 * tb + pos] = (uint32_t)cs[8 + P + k2]; }
    }
    __syncthreads();
    const int w = b_exact_min<P>(cnt);
    const int slot = cand[w];
    __syncthreads();
    b_tick(6);
    return slot;
}

// ------------------------------------------------------------------------------------------------
// Tableau.pivot(r, c) (tableau.py:295-308 in fraction-free form) with the winner's record in cand_ratio[slot]:
// per-prime constants, multipliers of the local rows, pivot row from its owner to every peer, update of the
// local rows (objective replica included), D <- p and its inverse for the next pivot.
// Returns false (nothing modified) when a prime divides the pivot element.
// ------------------------------------------------------------------------------------------------
template <int P>
__device__ __noinline__ bool b_apply(int slot) {
    BCtx& x = b_ctx();
    const int tid = threadIdx.x, T = (int)blockDim.x, ld = x.a.ld, C = x.a.C, mloc = x.a.mloc, mr = x.mr, mq = x.mq;
    const int words = batch_ratio_words(P);
    const int* rec = x.cand_ratio() + slot * words;
    const int r = rec[0], c = rec[1];
    const int owner = r / mloc, il_r = r - owner * mloc;
    const bool mine = owner == x.qr;
    const int sgn_p = rec[6];                                  // the record may be overwritten once barrier 3 is passed
    uint32_t* const planes = x.planes();
    const size_t plane_sz = x.plane_sz;
    const PrimeQ* const pr = x.pr();
    uint32_t* const nbeta = x.nbeta();
    uint32_t* const alpha_m = x.alpha_m();
    uint32_t* const dinv_m = x.dinv_m();
    if (tid == 0) x.S.flag = -1;
    __syncthreads();
    if (tid < P) {
        const PrimeQ p = pr[tid];
        const uint32_t piv = (uint32_t)rec[8 + tid];
        if (piv == 0u) atomicMax(&x.S.flag, tid);
        x.piv()[tid] = piv;
        alpha_m[tid] = mmul(mmul(piv, p.r2, p.q, p.qinv), dinv_m[tid], p.q, p.qinv);
    }
    // multipliers of the local rows: nbeta[k][il] = Montgomery form of -(N_ic / D); slot mloc = objective replica
#pragma unroll 1
    for (int e = tid; e < P * mr; e += T) {
        const int k = e / mr, il = e - k * mr;
        uint32_t nb = 0u;
        if ((il < mq || il == mloc) && !(mine && il == il_r)) {
            const PrimeQ p = pr[k];
            const uint32_t v = planes[(size_t)k * plane_sz + (size_t)il * ld + c];
            const uint32_t beta = mmul(mmul(v, p.r2, p.q, p.qinv), dinv_m[k], p.q, p.qinv);
            nb = beta ? p.q - beta : 0u;
        }
        nbeta[e] = nb;
    }
    const int vrow = (x.S.rc + 4) >> 2;                        // 128-bit vectors covering columns 0..rc
    if (mine) {                                                // pivot row to every CTA's prow (own copy included)
#pragma unroll 1
        for (int peer = 0; peer < C; ++peer) {
            uint32_t* dst = b_peer(x.prow(), peer);
#pragma unroll 1
            for (int e = tid; e < P * vrow; e += T) {
                const int k = e / vrow, jv = e - k * vrow;
                reinterpret_cast<uint4*>(dst + (size_t)k * ld)[jv] =
                    reinterpret_cast<const uint4*>(planes + (size_t)k * plane_sz + (size_t)il_r * ld)[jv];
            }
        }
    }
    __syncthreads();
    if (x.S.flag >= 0) return false;                           // a prime divides the pivot: nothing was modified
    b_tick(7);
    b_csync();                                                 // barrier 3: pivot row visible everywhere
    b_tick(8);
    // The last warp computes D <- p and its inverse for the next pivot (a 31-step chain) while the other warps
    // update.  Update: a thread owns (prime, 4 columns): the pivot-row vector and the prime's constants stay in
    // registers while it walks down the local rows (one 128-bit load + one 128-bit store per row and vector).
    const int Tu = T - 32;
    if (tid >= Tu) {
        const int k = tid - Tu;
        if (k < P) {
            const PrimeQ p = pr[k];
            const uint32_t piv = x.piv()[k];
            x.Dres()[k] = piv;
            dinv_m[k] = minv_mont(mmul(piv, p.r2, p.q, p.qinv), p);
        }
        if (k == 0) x.S.sgnD = sgn_p;
    } else {
        const uint32_t* const prow_all = x.prow();
        const int skip = mine ? il_r : -1;
#pragma unroll 1
        for (int it = tid; it < P * vrow; it += Tu) {
            const int k = it / vrow, jv = it - k * vrow;
            const uint32_t qq = pr[k].q, qinv = pr[k].qinv, al = alpha_m[k];
            const uint4 v4 = reinterpret_cast<const uint4*>(prow_all + (size_t)k * ld)[jv];
            const uint32_t* nbp = nbeta + k * mr;
            uint4* col = reinterpret_cast<uint4*>(planes + (size_t)k * plane_sz) + jv;
            const int ldv = ld >> 2;
            auto one_row = [&](int i) {
                uint4 c4 = col[i * ldv];
                const uint32_t nb = nbp[i];
                c4.x = redc_min((uint64_t)al * c4.x + (uint64_t)nb * v4.x, qq, qinv);
                c4.y = redc_min((uint64_t)al * c4.y + (uint64_t)nb * v4.y, qq, qinv);
                c4.z = redc_min((uint64_t)al * c4.z + (uint64_t)nb * v4.z, qq, qinv);
                c4.w = redc_min((uint64_t)al * c4.w + (uint64_t)nb * v4.w, qq, qinv);
                col[i * ldv] = c4;
            };
#pragma unroll 2
            for (int i = 0; i < mq; ++i)
                if (i != skip) one_row(i);
            one_row(mloc);                                     // objective replica
        }
        b_tick(9);
    }
    __syncthreads();
    b_tick(10);
    return true;
}

// would the tableau after pivot (r, c) still fit the prime set?  (block- and cluster-uniform: replicated state)
__device__ __forceinline__ bool b_fits(int r, int c) {
    const BCtx& x = b_ctx();
    if (!x.a.wcheck) return true;
    const int old = x.basis()[r];
    return x.S.wsum + x.wcol()[c] - x.wcol()[old] <= x.S.wlimit;
}

__device__ __forceinline__ void b_commit(int r, int c, int rule) {
    BCtx& x = b_ctx();   // Simplex._pivot simplex.py:192-197
    if (threadIdx.x == 0) {
        BState& S = x.S;
        S.wsum += x.wcol()[c] - x.wcol()[x.basis()[r]];
        if (x.qr == 0 && x.a.trace && S.npiv_total < x.a.trace_cap) {
            int* tr = x.a.trace + ((size_t)x.lp * x.a.trace_cap + S.npiv_total) * 3;
            tr[0] = r; tr[1] = c; tr[2] = rule;
        }
        x.basis()[r] = c;
        S.npiv_total++;
        S.npiv[S.phase]++;
    }
}

// Simplex.solve (simplex.py:110-148) on the current phase's tableau
template <int P>
__device__ __noinline__ void b_solve() {
    BCtx& x = b_ctx();
    BState& S = x.S;
    const int tid = threadIdx.x, m = x.a.m;
    const int words = batch_ratio_words(P);
    if (tid == 0) { S.rule = LPS_RULE_DANTZIG; S.stuck = 0; S.moved = 0; S.pending = 0; }
    __syncthreads();
    for (;;) {
        if (S.status != LPS_OK) return;
        if (x.a.max_pivots >= 0 && S.npiv_total >= x.a.max_pivots) {
            __syncthreads();
            if (tid == 0) S.status = LPS_MAX_PIVOTS;
            __syncthreads();
            return;
        }
        // stuck counter and rule switch: in exact arithmetic the objective is monotone, so it equals its
        // value at entry exactly until the first pivot with b_r > 0 (see k_xprice)
        if (tid == 0) {
            if (S.pending) { S.stuck = S.moved ? 0 : S.stuck + 1; S.pending = 0; }
            const long long limit = x.a.stuck_limit >= 0 ? x.a.stuck_limit : (long long)m + S.ncur;
            if (S.rule == LPS_RULE_DANTZIG && !(S.stuck < limit)) S.rule = LPS_RULE_BLAND;
        }
        __syncthreads();
        b_tick(11);
        const int rule = S.rule;
        const int c = b_price<P>(rule);
        if (c < 0) return;                                     // optimal for this phase
        const int slot = b_ratio<P>(c);
        if (slot < 0) {
            __syncthreads();
            if (tid == 0) S.status = LPS_UNBOUNDED;
            __syncthreads();
            return;
        }
        const int* rec = x.cand_ratio() + slot * words;
        const int r = rec[0];
        const int sgn_br = rec[7] * S.sgnD;
        __syncthreads();
        if (!b_fits(r, c)) {                                   // nothing modified: the host re-runs this LP on the full prime set
            __syncthreads();
            if (tid == 0) S.status = LPS_NEED_PRIMES;
            __syncthreads();
            return;
        }
        if (!b_apply<P>(slot)) {
            if (tid == 0) S.status = LPS_BAD_PRIME;
            __syncthreads();
            return;
        }
        b_commit(r, c, rule);
        if (tid == 0 && rule == LPS_RULE_DANTZIG) {
            S.pending = 1;
            if (sgn_br > 0) S.moved = 1;
            else if (sgn_br < 0 && !S.moved) S.status = LPS_OBJ_INCREASED;
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
template <int P>
__global__ void __launch_bounds__(kBMaxThreads, LPS_BATCH_MINB) k_bnb_nodes(BatchParams a) {
    BCtx& x = b_ctx();

    const int C = a.C, m = a.m, n = a.n, ld = a.ld, mloc = a.mloc;
    const int tid = threadIdx.x;
    if (tid == 0) {
        const BatchLayout L = batch_layout(m, P, C, mloc, a.cw, ld, a.tb);
        const size_t base = batch_ctx_bytes();
        x.a = a;
        x.qr = (C > 1) ? (int)cg::this_cluster().block_rank() : 0;
        x.lp = blockIdx.x / C;
        x.r0 = x.qr * mloc;
        x.mq = max(0, min(mloc, m - x.r0));
        x.mr = (mloc + 1 + 3) / 4 * 4;
        x.plane_sz = (size_t)(mloc + 1) * ld;
        x.o_planes = (unsigned)(base + L.planes);
        x.o_prow = (unsigned)(base + L.prow);
        x.o_objdig = (unsigned)(base + L.objdig);
        x.o_nbeta = (unsigned)(base + L.nbeta);
        x.o_tbc = (unsigned)(base + L.tbc);
        x.o_tbb = (unsigned)(base + L.tbb);
        x.o_inv = (unsigned)(base + L.inv);
        x.o_half = (unsigned)(base + L.half);
        x.o_Dres = (unsigned)(base + L.dres);
        x.o_dinv_m = (unsigned)(base + L.dinv);
        x.o_alpha_m = (unsigned)(base + L.alpha);
        x.o_piv = (unsigned)(base + L.piv);
        x.o_cand_price = (unsigned)(base + L.cand_price);
        x.o_cand_ratio = (unsigned)(base + L.cand_ratio);
        x.o_obj_top = (unsigned)(base + L.obj_top);
        x.o_col_top = (unsigned)(base + L.col_top);
        x.o_rhs_top = (unsigned)(base + L.rhs_top);
        x.o_tb_row = (unsigned)(base + L.tb_row);
        x.o_cand = (unsigned)(base + L.cand);
        x.o_basis = (unsigned)(base + L.basis);
        x.o_rowsign = (unsigned)(base + L.rowsign);
        x.o_wcol = (unsigned)(base + L.wcol);
        x.o_wexp = (unsigned)(base + L.wexp);
        x.o_wmant = (unsigned)(base + L.wmant);
        x.o_rcp = (unsigned)(base + L.rcp);
        x.o_col_mant = (unsigned)(base + L.col_mant);
        x.o_rhs_mant = (unsigned)(base + L.rhs_mant);
        x.o_pr = (unsigned)(base + L.pr);
        x.o_obj_sign = (unsigned)(base + L.obj_sign);
        x.o_col_sign = (unsigned)(base + L.col_sign);
        x.o_rhs_sign = (unsigned)(base + L.rhs_sign);
        x.G = a.grids + (size_t)x.lp * (m + 1) * (n + 1);
        BState& S0 = x.S;
        S0.status = LPS_OK; S0.phase = 0; S0.K = 0; S0.rule = LPS_RULE_DANTZIG; S0.moved = 0; S0.pending = 0; S0.stuck = 0;
        S0.npiv_total = 0; S0.npiv[0] = S0.npiv[1] = S0.npiv[2] = 0; S0.sgnD = 1; S0.flag = -1; S0.ncur = n; S0.rc = n;
        for (int s = 0; s < 16; ++s) x.tacc[s] = 0;
        x.tlast = clock64();
    }
    __syncthreads();
    BState& S = x.S;
    const int32_t* G = x.G;
    int* basis = x.basis();
    int* rowsign = x.rowsign();
    const int qr = x.qr, lp = x.lp, r0 = x.r0, mq = x.mq;
    const size_t plane_sz = x.plane_sz;
    uint32_t* planes = x.planes();
    const int words = batch_ratio_words(P);

    // ---- tables, zeroed planes -------------------------------------------------------------------
    for (int k = tid; k < P; k += (int)blockDim.x) { x.pr()[k] = a.pr[k]; x.half()[k] = a.half[k]; x.Dres()[k] = 1u; x.dinv_m()[k] = a.pr[k].r1; }
    for (int k = tid; k <= P; k += (int)blockDim.x) {          // W_k = p_0 ... p_(k-1) (< 2^496) and 1 / p_k as doubles
        x.wmant()[k] = ldexp(a.wmant[k], a.wexp[k]);
        x.wexp()[k] = a.wexp[k];
        x.rcp()[k] = k < P ? 1.0 / (double)a.pr[k].q : 0.0;
    }
    for (int e = tid; e < P * P; e += (int)blockDim.x) x.inv()[e] = a.invtab[e];
    {
        uint4* z = reinterpret_cast<uint4*>(planes);
        const int nz = (int)((size_t)P * plane_sz / 4);
        for (int e = tid; e < nz; e += (int)blockDim.x) z[e] = make_uint4(0u, 0u, 0u, 0u);
    }

    // ---- Simplex._find_bfs, part 1 (simplex.py:42-60) on the integer input, replicated on every CTA --------
    // rows with b_i < 0 are negated; isCanonical: column j is basic for row i when c_j == 0 and its
    // only nonzero is an exact 1 in row i (the sign flip applied); bcols[i] = lowest such j.
    for (int i = tid; i < m; i += (int)blockDim.x) { basis[i] = INT_MAX; rowsign[i] = (G[(size_t)i * (n + 1) + n] < 0) ? -1 : 1; }
    __syncthreads();
    for (int j = tid; j < n; j += (int)blockDim.x) {
        if (G[(size_t)m * (n + 1) + j] != 0) continue;
        int hit = -1;
        bool clean = true;
        for (int i = 0; clean && i < m; ++i) {
            const int v = G[(size_t)i * (n + 1) + j] * rowsign[i];
            if (v == 0) continue;
            if (v == 1 && hit == -1) hit = i; else clean = false;
        }
        if (clean && hit != -1) atomicMin(&basis[hit], j);
    }
    __syncthreads();
    if (tid == 0) {                                            // artificial columns in row order (simplex.py:54-60)
        int K = 0;
        for (int i = 0; i < m; ++i)
            if (basis[i] == INT_MAX) { basis[i] = n + K; ++K; }
        S.K = K;
        S.ncur = n + K;
        S.rc = n + K;
        S.phase = K > 0 ? 0 : 2;
        if (K > a.art_cap) S.status = LPS_NEED_ARTIFICIALS;
    }
    __syncthreads();
    const int K = S.K;
    if (a.wcheck) {
        // column weights of the extended matrix [A | artificial units | b ; c | 1 | 0] (the cost row of either phase)
        int* wcol = x.wcol();
        for (int j = tid; j <= n + K; j += (int)blockDim.x) {
            double sq = 0.0;
            if (j < n || j == n + K) {
                const int js = j < n ? j : n;
                for (int i = 0; i <= m; ++i) { const double v = (double)G[(size_t)i * (n + 1) + js]; sq += v * v; }
            } else sq = 2.0;                                  // an artificial column: one 1 and cost 1
            wcol[j] = sq <= 1.0 ? 0 : (int)ceil(128.0 * log2(sq * (1.0 + 1e-9))) + 1;
        }
        __syncthreads();
        if (tid == 0) {
            long long wmax = 0, wsum = 0;
            for (int j = 0; j <= n + K; ++j) wmax = max(wmax, (long long)wcol[j]);
            for (int i = 0; i < m; ++i) wsum += wcol[basis[i]];
            S.wsum = wsum;
            S.wlimit = 256LL * (a.wexp[P] - 2) - 2 * wmax - 256LL * 3;   // bits of the prime set, two extra columns, guard
        }
        __syncthreads();
    }

    // ---- encode the local rows and the objective replica into residues --------------------------------
    // phase-1 tableau when K > 0: costs 0, z 0, artificial unit columns with cost 1, then
    // rowSubFromObj(i, 1) for every artificial row: objective row = -(sum of those rows) on the original
    // columns and the rhs, 0 on the artificial columns (simplex.py:49-60)
    if (S.status == LPS_OK) {
        const int W = n + K + 1;
        for (int e = tid; e < (mq + 1) * W; e += (int)blockDim.x) {
            const int il0 = e / W, j = e - il0 * W;
            const bool is_obj = il0 == mq;
            const int il = is_obj ? mloc : il0, i = r0 + il0;
            long long v = 0;
            if (!is_obj) {
                if (j < n) v = (long long)G[(size_t)i * (n + 1) + j] * rowsign[i];
                else if (j < n + K) v = (basis[i] == j) ? 1 : 0;
                else v = (long long)G[(size_t)i * (n + 1) + n] * rowsign[i];
            } else if (K == 0) {
                v = G[(size_t)m * (n + 1) + (j < n ? j : n)];
            } else if (j < n || j == n + K) {
                const int js = j < n ? j : n;
                for (int i2 = 0; i2 < m; ++i2)
                    if (basis[i2] >= n) v -= (long long)G[(size_t)i2 * (n + 1) + js] * rowsign[i2];
            }
            const unsigned long long mag = v < 0 ? (unsigned long long)(-v) : (unsigned long long)v;
#pragma unroll 1
            for (int k = 0; k < P; ++k) {
                const uint32_t qq = x.pr()[k].q;
                uint32_t res = mag < qq ? (uint32_t)mag : (uint32_t)(mag % qq);   // coefficients are usually far below the primes
                if (v < 0 && res) res = qq - res;
                planes[(size_t)k * plane_sz + (size_t)il * ld + j] = res;
            }
        }
    }
    b_csync();                                                // every CTA of the cluster is set up
    b_tick(12);

    // ---- the constructor's phase 1 -----------------------------------------------------------------
    if (S.status == LPS_OK && K > 0) {
        b_solve<P>();                                         // simplex.py:62
        if (S.status == LPS_OK) {                              // artificial optimum != 0 ?  (simplex.py:63-67)
            if (tid == 0) {
                bool nz = false;
                for (int k = 0; k < P; ++k) nz |= planes[(size_t)k * plane_sz + (size_t)mloc * ld + S.rc] != 0u;
                if (nz) S.status = LPS_INFEASIBLE;
            }
            __syncthreads();
        }
        if (S.status == LPS_OK) {                              // drive basic artificials out (simplex.py:68-84)
            if (tid == 0) S.phase = 1;
            __syncthreads();
            bool redundant = false;
            for (int i = 0; i < m && S.status == LPS_OK; ++i) {
                if (basis[i] < n) continue;
                const int owner = i / mloc, il = i - owner * mloc;
                if (owner == qr) {
                    // first original column with a nonzero entry in row i: nonzero modulo some prime
                    if (tid == 0) S.low = INT_MAX;
                    __syncthreads();
                    int first = INT_MAX;
                    for (int j = tid; j < n; j += (int)blockDim.x) {
                        bool nz = false;
                        for (int k = 0; k < P; ++k) nz |= planes[(size_t)k * plane_sz + (size_t)il * ld + j] != 0u;
                        if (nz) { first = j; break; }
                    }
                    first = __reduce_min_sync(0xffffffffu, first);
                    if ((tid & 31) == 0 && first != INT_MAX) atomicMin(&S.low, first);
                    __syncthreads();
                    const int j1 = S.low;
                    // publish the forced pivot in this CTA's ratio slot of every peer: sign of the pivot from one Garner
                    if (j1 != INT_MAX && tid < 32) {           // warp 0 (garner_lane is warp-collective), group 0 holds the number
                        constexpr int LP = LaneCfg<P>::LP;
                        const int k = tid & (LP - 1);
                        const bool live = k < P;
                        const PrimeQ pk = x.pr()[live ? k : 0];
                        uint32_t t = (live && tid < LP) ? planes[(size_t)k * plane_sz + (size_t)il * ld + j1] : 0u;
                        int sg, tp;
                        garner_lane<P>(t, k, pk.q, pk.qinv, x.inv(), x.half()[live ? k : 0], sg, tp);
                        if (tid == 0) S.low2 = sg;
                    }
                    __syncthreads();
                    const int sg = S.low2;
                    for (int e = tid; e < C * words; e += (int)blockDim.x) {
                        const int peer = e / words, w = e - peer * words;
                        int v = 0;
                        if (w == 0) v = j1 == INT_MAX ? -1 : i;
                        else if (j1 == INT_MAX) v = 0;
                        else if (w == 1) v = j1;
                        else if (w == 6) v = sg;
                        else if (w >= 8 && w < 8 + P) v = (int)planes[(size_t)(w - 8) * plane_sz + (size_t)il * ld + j1];
                        b_peer(x.cand_ratio(), peer)[qr * words + w] = v;
                    }
                }
                b_csync();
                const int* rec = x.cand_ratio() + owner * words;
                if (rec[0] < 0) { redundant = true; b_csync(); continue; }   // dependent row: marked for deletion (simplex.py:82-84)
                const int j1 = rec[1];
                if (!b_fits(i, j1)) {
                    __syncthreads();
                    if (tid == 0) S.status = LPS_NEED_PRIMES;
                    __syncthreads();
                    break;
                }
                if (!b_apply<P>(owner)) {
                    if (tid == 0) S.status = LPS_BAD_PRIME;
                    __syncthreads();
                    break;
                }
                b_commit(i, j1, 2);
                __syncthreads();
            }
            if (redundant && S.status == LPS_OK) {             // the reference fails after dropping the row
                __syncthreads();
                if (tid == 0) S.status = LPS_REDUNDANT_ROW;
                __syncthreads();
            }
        }
        if (S.status == LPS_OK) {
            // cut the artificial columns, restore the objective and price it out against the basis
            // (simplex.py:96-105): N_mj = D c_j - sum_i c_B(i) N_ij, the rhs column included.  Each CTA sums
            // its own rows into prow, then every CTA adds up the partial sums of the whole cluster.
            const int rc_old = S.rc;
            for (int e = tid; e < P * (mloc + 1); e += (int)blockDim.x) {
                const int k = e / (mloc + 1), il = e - k * (mloc + 1);
                uint32_t* row = planes + (size_t)k * plane_sz + (size_t)il * ld;
                row[n] = row[rc_old];
            }
            __syncthreads();
            auto to_mont = [&](long long v, const PrimeQ& p) -> uint32_t {
                const unsigned long long mag = v < 0 ? (unsigned long long)(-v) : (unsigned long long)v;
                uint32_t res = (uint32_t)(mag % p.q);
                if (v < 0 && res) res = p.q - res;
                return mmul(res, p.r2, p.q, p.qinv);
            };
            for (int e = tid; e < P * (n + 1); e += (int)blockDim.x) {
                const int k = e / (n + 1), j = e - k * (n + 1);
                const PrimeQ p = x.pr()[k];
                const uint32_t* pl = planes + (size_t)k * plane_sz;
                uint32_t acc = 0u;
                for (int il = 0; il < mq; ++il) {
                    const int cb = G[(size_t)m * (n + 1) + basis[r0 + il]];
                    if (cb == 0) continue;
                    acc += mmul(to_mont(cb, p), pl[(size_t)il * ld + j], p.q, p.qinv);
                    if (acc >= p.q) acc -= p.q;
                }
                x.prow()[(size_t)k * ld + j] = acc;
            }
            b_csync();
            for (int e = tid; e < P * (n + 1); e += (int)blockDim.x) {
                const int k = e / (n + 1), j = e - k * (n + 1);
                const PrimeQ p = x.pr()[k];
                uint32_t acc = mmul(to_mont(G[(size_t)m * (n + 1) + j], p), x.Dres()[k], p.q, p.qinv);
                for (int s = 0; s < C; ++s) acc = msub(acc, b_peer(x.prow(), s)[(size_t)k * ld + j], p.q);
                planes[(size_t)k * plane_sz + (size_t)mloc * ld + j] = acc;
            }
            if (tid == 0) { S.ncur = n; S.rc = n; S.phase = 2; }
            b_csync();
        }
    }
    // ---- the optimisation proper -------------------------------------------------------------------
    b_tick(13);
    if (S.status == LPS_OK) {
        b_solve<P>();
        if (S.status == LPS_OK) {
            __syncthreads();
            if (tid == 0) S.status = LPS_OPTIMAL;
            __syncthreads();
        }
    }

    // ---- results: rhs column (b_0..b_{m-1}, z cell) and D as sign + mixed-radix digits -------------
    const int st = S.status;
    if (st != LPS_NEED_ARTIFICIALS) {
        constexpr int LP = LaneCfg<P>::LP, G = LaneCfg<P>::G;
        const int lane = tid & 31, warp = tid >> 5, nwarps = (int)blockDim.x >> 5;
        const int k = lane & (LP - 1), sub = lane / LP;
        const bool live = k < P;
        const PrimeQ pk = x.pr()[live ? k : 0];
        const uint32_t half_k = x.half()[live ? k : 0];
        signed char* so = a.sign + (size_t)lp * (m + 2);
        uint32_t* dg = a.digits + (size_t)lp * P * (m + 2);
        const int extra = qr == 0 ? 2 : 0;                     // CTA 0 also writes the z cell and D
        const uint32_t* pl = planes + (size_t)(live ? k : 0) * plane_sz + S.rc;
        for (int base = warp * G; base < mq + extra; base += nwarps * G) {   // whole warps: garner_lane is warp-collective
            const int e = base + sub;
            const bool valid = e < mq + extra;
            uint32_t t = 0u;
            int out = 0;
            if (valid && live) {
                if (e < mq) t = pl[(size_t)e * ld];
                else if (e == mq) t = pl[(size_t)mloc * ld];
                else t = x.Dres()[k];
            }
            if (valid) out = e < mq ? r0 + e : (e == mq ? m : m + 1);
            int sg, tp;
            garner_lane<P>(t, k, pk.q, pk.qinv, x.inv(), half_k, sg, tp);
            if (valid && live) dg[(size_t)k * (m + 2) + out] = t;
            if (valid && k == 0) so[out] = (signed char)sg;
        }
    }
    if (qr == 0) {
        for (int i = tid; i < m; i += (int)blockDim.x) a.basis[(size_t)lp * m + i] = basis[i];
        if (tid == 0) {
            a.status[lp] = st;
            a.npiv[lp * 4 + 0] = S.npiv[0]; a.npiv[lp * 4 + 1] = S.npiv[1]; a.npiv[lp * 4 + 2] = S.npiv[2];
            a.npiv[lp * 4 + 3] = S.K;
        }
    }
    b_tick(14);
    if (a.timing && qr == 0 && tid == 0)
        for (int s = 0; s < 16; ++s) a.timing[(size_t)lp * 16 + s] = x.tacc[s];
    b_csync();                                                // nobody leaves while a peer may still write into it
}

// ------------------------------------------------------------------------------------------------
// sizing of a batch on the device (one CTA per LP): how many rows get an artificial column (same rule as the
// set-up of k_bnb_nodes) and log2 of a Hadamard bound on every minor of the LP's extended matrix -- the
// smaller of the product of all row norms and the product of the min(m+1, n+1) largest column norms, each norm
// at least 1, the objective row taken as max(|costs|, sum of |rows that get an artificial|) so that it covers
// the phase-1 objective too.  Pure bookkeeping: no pivot arithmetic.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_batch_plan(const int32_t* grids, int m, int n, int* out_k, double* out_bits) {
    extern __shared__ __align__(16) unsigned char plan_smem[];
    int* basis = reinterpret_cast<int*>(plan_smem);                  // m
    int* rowsign = basis + m;                                        // m
    double* colsq = reinterpret_cast<double*>(rowsign + m + ((2 * m) & 1));   // n + 1
    double* rowlog = colsq + (n + 1);                                // m + 1 partial sums
    __shared__ double s_sum[2];
    const int tid = threadIdx.x, T = blockDim.x;
    const int32_t* G = grids + (size_t)blockIdx.x * (m + 1) * (n + 1);
    for (int i = tid; i < m; i += T) { basis[i] = INT_MAX; rowsign[i] = (G[(size_t)i * (n + 1) + n] < 0) ? -1 : 1; }
    if (tid < 2) s_sum[tid] = 0.0;
    __syncthreads();
    for (int j = tid; j < n; j += T) {
        if (G[(size_t)m * (n + 1) + j] != 0) continue;
        int hit = -1;
        bool clean = true;
        for (int i = 0; clean && i < m; ++i) {
            const int v = G[(size_t)i * (n + 1) + j] * rowsign[i];
            if (v == 0) continue;
            if (v == 1 && hit == -1) hit = i; else clean = false;
        }
        if (clean && hit != -1) atomicMin(&basis[hit], j);
    }
    __syncthreads();
    // column norms (objective entry: max of the cost and the phase-1 cost bound) and row norms
    for (int j = tid; j <= n; j += T) {
        double sq = 0.0, art = 0.0;
        for (int i = 0; i < m; ++i) {
            const double v = fabs((double)G[(size_t)i * (n + 1) + j]);
            sq += v * v;
            if (basis[i] == INT_MAX) art += v;
        }
        const double o = fmax(fabs((double)G[(size_t)m * (n + 1) + j]), art);
        colsq[j] = fmax(sq + o * o, 1.0);
        rowlog[0] = 0.0;
        atomicAdd(&s_sum[1], o * o);                               // squared norm of the objective row
    }
    __syncthreads();
    double part = 0.0;
    int k_art = 0;
    for (int i = tid; i < m; i += T) {
        double sq = 0.0;
        for (int j = 0; j <= n; ++j) { const double v = (double)G[(size_t)i * (n + 1) + j]; sq += v * v; }
        part += log2(fmax(sq, 1.0));
        k_art += basis[i] == INT_MAX;
    }
    // rows: all m + 1 of them when m + 1 <= n + 1, else the n + 1 largest -- the second case (more rows than
    // columns) is bounded by the column product alone, so the row product is only used in the first
    atomicAdd(&s_sum[0], part);
    __syncthreads();
    const int order = min(m + 1, n + 1);
    double by_cols = 0.0;
    for (int j = tid; j <= n; j += T) {                              // the `order` largest column norms (rank by counting)
        const double v = colsq[j];
        int rank = 0;
        for (int i = 0; i <= n; ++i) rank += (colsq[i] > v) || (colsq[i] == v && i < j);
        if (rank < order) by_cols += log2(v);
    }
    __shared__ double s_cols;
    __shared__ int s_k;
    if (tid == 0) { s_cols = 0.0; s_k = 0; }
    __syncthreads();
    atomicAdd(&s_cols, by_cols);
    atomicAdd(&s_k, k_art);
    __syncthreads();
    if (tid == 0) {
        const double by_rows = (s_sum[0] + log2(fmax(s_sum[1], 1.0))) / 2;
        double bits = s_cols / 2;
        if (m + 1 <= n + 1) bits = fmin(bits, by_rows);
        out_bits[blockIdx.x] = bits;
        out_k[blockIdx.x] = s_k;
    }
}

}  // namespace lps
