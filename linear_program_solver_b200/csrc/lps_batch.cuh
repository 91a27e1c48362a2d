// Batched exact-mode node LPs for branch and bound: one LP per thread-block CLUSTER, the whole
// multi-modular tableau resident in (distributed) shared memory, the complete `Simplex(tab); solve()`
// of the reference -- phase 1 included -- inside ONE kernel launch.
//
// The LP's fraction-free integer tableau N (lps_exact.cuh) is held as residues modulo P 31-bit primes.
// A cluster of C CTAs owns one LP; CTA q keeps the residue planes of primes [q*PL, (q+1)*PL) in its own
// shared memory (C = 1 when all planes fit one CTA: "one node LP per CTA").  Every decision needs all
// residues of a few numbers (objective row, entering column, right-hand side): each CTA gathers them
// from its peers through distributed shared memory, reconstructs them (Garner, one THREAD per number,
// the digit vector in registers) and takes the decision redundantly, so the only cluster-wide
// synchronisation is two barriers per pivot (planes stable -> read; all reads done -> update).
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

constexpr int kBThreads = 256;
constexpr int kBMaxCluster = 8;

struct BatchParams {
    int count, m, n, art_cap;        // LPs, constraint rows, variables, artificial columns reserved per LP
    int P, PL, C;                    // primes, primes per CTA, CTAs per cluster
    int ld, mr;                      // plane row stride (>= n + art_cap + 1, multiple of 4); m+1 rounded up to 4
    long long max_pivots;            // safety cap per LP (< 0: none)
    long long stuck_limit;           // < 0: the reference's m + n (simplex.py:123); tests lower it to reach the Bland loop
    const PrimeQ* pr;
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
};

// Shared-memory carve-up, identical on host (sizing) and device.
struct BatchLayout {
    size_t planes, objdig, colres, rhsres, nbeta, inv, half, dres, dinv, alpha, obj_top, col_top, rhs_top, basis, cand,
        wexp, slots, wmant, col_mant, rhs_mant, pr, obj_sign, col_sign, rhs_sign, total;
};
__host__ __device__ inline BatchLayout batch_layout(int m, int P, int PL, int ld, int mr) {
    BatchLayout L;
    size_t o = 0;
    auto take = [&](size_t bytes, size_t align) { o = (o + align - 1) / align * align; const size_t at = o; o += bytes; return at; };
    L.planes = take((size_t)PL * (m + 1) * ld * 4, 16);
    L.objdig = take((size_t)P * ld * 4, 16);
    L.colres = take((size_t)P * mr * 4, 16);
    L.rhsres = take((size_t)P * mr * 4, 16);
    L.nbeta = take((size_t)PL * mr * 4, 16);
    L.inv = take((size_t)P * P * 4, 16);
    L.half = take((size_t)P * 4, 4);
    L.dres = take((size_t)P * 4, 4);
    L.dinv = take((size_t)P * 4, 4);
    L.alpha = take((size_t)P * 4, 4);
    L.obj_top = take((size_t)ld * 4, 4);
    L.col_top = take((size_t)mr * 4, 4);
    L.rhs_top = take((size_t)mr * 4, 4);
    L.basis = take((size_t)mr * 4, 4);
    L.cand = take((size_t)mr * 4, 4);
    L.wexp = take((size_t)(P + 1) * 4, 4);
    L.slots = take((size_t)2 * kBMaxCluster * 4, 4);
    L.wmant = take((size_t)(P + 1) * 8, 8);
    L.col_mant = take((size_t)mr * 8, 8);
    L.rhs_mant = take((size_t)mr * 8, 8);
    L.pr = take((size_t)P * sizeof(PrimeQ), 16);
    L.obj_sign = take((size_t)ld, 4);
    L.col_sign = take((size_t)mr, 4);
    L.rhs_sign = take((size_t)mr, 4);
    L.total = (o + 15) / 16 * 16;
    return L;
}

// ------------------------------------------------------------------------------------------------
// Garner, one thread per number: t[k] = residue modulo prime k on entry; on return t[] holds the
// mixed-radix digits of |x|, sign in {-1,0,1}, top = index of the leading nonzero digit (-1 for 0).
// The j-outer order keeps the dependent chain at P steps with P/2 independent multiplies per step.
// ------------------------------------------------------------------------------------------------
template <int PMAX>
__device__ __forceinline__ void garner_num(uint32_t (&t)[PMAX], int P, const uint32_t (&q)[PMAX], const uint32_t (&qi)[PMAX],
                                           const uint32_t* sInv, const uint32_t* sHalf, int& sign, int& top) {
#pragma unroll
    for (int j = 0; j < PMAX - 1; ++j) {
        if (j < P) {
            const uint32_t dj = t[j];
            const uint32_t* invrow = sInv + j * P;
#pragma unroll
            for (int k = j + 1; k < PMAX; ++k) {
                if (k < P) {
                    uint32_t v = dj;
                    if (v >= q[k]) { v -= q[k]; if (v >= q[k]) v %= q[k]; }
                    t[k] = mmul(msub(t[k], v, q[k]), invrow[k], q[k], qi[k]);
                }
            }
        }
    }
    bool decided = false, neg = false;
    int lo = -1;
#pragma unroll
    for (int k = PMAX - 1; k >= 0; --k) {
        if (k < P) {
            const uint32_t h = sHalf[k];
            if (!decided && t[k] != h) { decided = true; neg = t[k] > h; }
            if (t[k] != 0u) lo = k;
        }
    }
    top = -1;
    if (lo < 0) { sign = 0; return; }
    sign = neg ? -1 : 1;
#pragma unroll
    for (int k = 0; k < PMAX; ++k) {
        if (k < P) {
            uint32_t a = t[k];
            if (neg) a = k < lo ? 0u : (k == lo ? q[k] - a : q[k] - 1u - a);
            t[k] = a;
            if (a) top = k;
        }
    }
}
// |x| ~ mant * W_top from the three leading digits (same formula as garner_mant in lps_exact.cuh)
template <int PMAX>
__device__ __forceinline__ double mant_of(const uint32_t (&t)[PMAX], int top, const uint32_t (&q)[PMAX]) {
    if (top < 0) return 0.0;
    uint32_t d0 = 0, d1 = 0, d2 = 0, q1 = 1, q2 = 1;
#pragma unroll
    for (int k = 0; k < PMAX; ++k) {
        if (k == top) d0 = t[k];
        if (k == top - 1) { d1 = t[k]; q1 = q[k]; }
        if (k == top - 2) { d2 = t[k]; q2 = q[k]; }
    }
    double x = (double)d0;
    if (top >= 1) {
        double y = (double)d1;
        if (top >= 2) y += (double)d2 / (double)q2;
        x += y / (double)q1;
    }
    return x;
}

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
};

// Everything a device function of the kernel needs, built once per CTA in shared memory.
struct BCtx {
    BatchParams a;
    int qr, lp, k0, nloc;        // rank in the cluster, LP index, first local prime, local primes
    size_t plane_sz;
    uint32_t *planes, *objdig, *colres, *rhsres, *nbeta, *inv, *half, *Dres, *dinv_m, *alpha_m;
    int *obj_top, *col_top, *rhs_top, *basis, *cand, *wexp, *slots;
    double *wmant, *col_mant, *rhs_mant;
    PrimeQ* pr;
    signed char *obj_sign, *col_sign, *rhs_sign;
    const int32_t* G;
    BState S;
    int s_red[32];
    int s_cls[8], s_e[8];
    double s_f[8];
};

__device__ __forceinline__ void b_csync(const BCtx& x) {
    if (x.a.C > 1) cg::this_cluster().sync(); else __syncthreads();
}
// residue of cell (i, j) modulo prime k, wherever its plane lives in the cluster
__device__ __forceinline__ uint32_t b_cell(const BCtx& x, int k, int i, int j) {
    const int owner = k / x.a.PL;
    const uint32_t* base = x.planes;
    if (owner != x.qr) base = cg::this_cluster().map_shared_rank(x.planes, owner);
    return base[(size_t)(k - owner * x.a.PL) * x.plane_sz + (size_t)i * x.a.ld + j];
}

// ------------------------------------------------------------------------------------------------
// stage: gather the residues of a set of numbers from the cluster, reconstruct them, leave what the
// decisions need in shared memory.
//   mode 0  objective row, columns 0..ncur-1: sign / top, digits of the negative ones (pricing)
//   mode 1  column c rows 0..m and the rhs column rows 0..m: residues of all primes (colres / rhsres),
//           sign / top / leading-digit mantissa (ratio test, multipliers of the update)
//   mode 2  rhs column only, digits into objdig (results; infeasibility test)
//   mode 3  the denominator D from Dres, digits into objdig[.][m+1], sign into rhs_sign slot m+1 (mr > m+1
//           is not guaranteed, so the sign goes to S.flag)
// ------------------------------------------------------------------------------------------------
template <int PMAX>
__device__ __noinline__ void b_stage(BCtx& x, int mode, int c) {
    const int P = x.a.P, m = x.a.m, ld = x.a.ld, mr = x.a.mr, tid = threadIdx.x;
    uint32_t pq[PMAX], pqi[PMAX];
#pragma unroll
    for (int k = 0; k < PMAX; ++k) { pq[k] = k < P ? x.pr[k].q : 3u; pqi[k] = k < P ? x.pr[k].qinv : 0u; }
    const int sgnD = x.S.sgnD, rc = x.S.rc;
    const int count = mode == 0 ? x.S.ncur : (mode == 1 ? 2 * (m + 1) : (mode == 2 ? m + 1 : 1));
    for (int e = tid; e < count; e += kBThreads) {
        const bool is_col = (mode == 1 && e <= m);
        int i, j;
        if (mode == 0) { i = m; j = e; }
        else if (mode == 1) { i = is_col ? e : e - (m + 1); j = is_col ? c : rc; }
        else { i = e; j = rc; }
        uint32_t t[PMAX];
#pragma unroll
        for (int k = 0; k < PMAX; ++k) t[k] = (k < P) ? (mode == 3 ? x.Dres[k] : b_cell(x, k, i, j)) : 0u;
        if (mode == 1) {
            uint32_t* res = is_col ? x.colres : x.rhsres;
#pragma unroll
            for (int k = 0; k < PMAX; ++k)
                if (k < P) res[k * mr + i] = t[k];
        }
        int sg, tp;
        garner_num<PMAX>(t, P, pq, pqi, x.inv, x.half, sg, tp);
        if (mode == 0) {
            x.obj_sign[e] = (signed char)sg;
            x.obj_top[e] = tp;
            if (sg * sgnD < 0) {
#pragma unroll
                for (int k = 0; k < PMAX; ++k)
                    if (k < P) x.objdig[k * ld + e] = t[k];
            }
        } else if (mode == 1) {
            (is_col ? x.col_sign : x.rhs_sign)[i] = (signed char)sg;
            (is_col ? x.col_top : x.rhs_top)[i] = tp;
            (is_col ? x.col_mant : x.rhs_mant)[i] = mant_of<PMAX>(t, tp, pq);
        } else {
            const int slot = mode == 2 ? i : m + 1;
            if (mode == 2) x.rhs_sign[i] = (signed char)sg; else x.S.flag = sg;
#pragma unroll
            for (int k = 0; k < PMAX; ++k)
                if (k < P) x.objdig[k * ld + slot] = t[k];
        }
    }
    __syncthreads();
}

// pricing (simplex.py:262-267 / :228-232) on the staged objective row; block-uniform result, -1: optimal
__device__ __noinline__ int b_price(BCtx& x, int rule) {
    const int ncur = x.S.ncur, sgnD = x.S.sgnD, ld = x.a.ld, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    auto pick = [&](int u, int v) -> int {
        if (u < 0) return v;
        if (v < 0) return u;
        if (rule == LPS_RULE_BLAND) return min(u, v);
        const int cmp = cmp_mag(x.obj_top, x.objdig, (size_t)ld, u, v);
        return cmp > 0 ? u : (cmp < 0 ? v : min(u, v));
    };
    int best = -1;
    for (int j = tid; j < ncur; j += kBThreads)
        if ((int)x.obj_sign[j] * sgnD < 0) best = pick(best, j);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) best = pick(best, __shfl_xor_sync(0xffffffffu, best, off));
    if (lane == 0) x.s_red[warp] = best;
    __syncthreads();
    if (warp == 0) {
        int b2 = lane < kBThreads / 32 ? x.s_red[lane] : -1;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) b2 = pick(b2, __shfl_xor_sync(0xffffffffu, b2, off));
        if (lane == 0) x.s_red[0] = b2;
    }
    __syncthreads();
    const int c = x.s_red[0];
    __syncthreads();
    return c;
}

// ratio test (simplex.py:272-279) on the staged column / rhs: floating filter on the leading digits,
// exact decision among the near-ties by the sign of (N_kn N_rc - N_rn N_kc) / D.  Returns -1: unbounded.
template <int PMAX>
__device__ __noinline__ int b_ratio(BCtx& x) {
    const int P = x.a.P, m = x.a.m, mr = x.a.mr, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int sgnD = x.S.sgnD;
    auto key_of = [&](int i) {
        RKey k;
        k.cls = (int)x.rhs_sign[i] * (int)x.col_sign[i];
        k.e = 0; k.f = 0.0;
        if (k.cls != 0) {
            int ex;
            k.f = frexp((x.rhs_mant[i] * x.wmant[x.rhs_top[i]]) / (x.col_mant[i] * x.wmant[x.col_top[i]]), &ex);
            k.e = ex + x.wexp[x.rhs_top[i]] - x.wexp[x.col_top[i]];
        }
        return k;
    };
    RKey best{2, 0, 0.0};
    for (int i = tid; i < m; i += kBThreads) {
        if ((int)x.col_sign[i] * sgnD <= 0) continue;
        const RKey k = key_of(i);
        if (rkey_less(k, best)) best = k;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        RKey o;
        o.cls = __shfl_xor_sync(0xffffffffu, best.cls, off);
        o.e = __shfl_xor_sync(0xffffffffu, best.e, off);
        o.f = __shfl_xor_sync(0xffffffffu, best.f, off);
        if (rkey_less(o, best)) best = o;
    }
    if (lane == 0) { x.s_cls[warp] = best.cls; x.s_e[warp] = best.e; x.s_f[warp] = best.f; }
    if (tid == 0) { x.S.count = 0; x.S.low = INT_MAX; }
    __syncthreads();
    best = RKey{x.s_cls[0], x.s_e[0], x.s_f[0]};
    for (int w = 1; w < kBThreads / 32; ++w) {
        const RKey o{x.s_cls[w], x.s_e[w], x.s_f[w]};
        if (rkey_less(o, best)) best = o;
    }
    if (best.cls == 2) { __syncthreads(); return -1; }
    if (best.cls == 0) {                                       // ratio exactly 0: lowest such row
        int low = INT_MAX;
        for (int i = tid; i < m; i += kBThreads)
            if ((int)x.col_sign[i] * sgnD > 0 && x.rhs_sign[i] == 0) low = min(low, i);
        low = __reduce_min_sync(0xffffffffu, low);
        if (lane == 0 && low != INT_MAX) atomicMin(&x.S.low, low);
        __syncthreads();
        const int r = x.S.low;
        __syncthreads();
        return r;
    }
    for (int i = tid; i < m; i += kBThreads) {
        if ((int)x.col_sign[i] * sgnD <= 0) continue;
        if (rkey_close(key_of(i), best)) x.cand[atomicAdd(&x.S.count, 1)] = i;
    }
    __syncthreads();
    const int count = x.S.count;
    int r = x.cand[0];
    if (count > 1) {
        uint32_t pq[PMAX], pqi[PMAX];
#pragma unroll
        for (int k = 0; k < PMAX; ++k) { pq[k] = k < P ? x.pr[k].q : 3u; pqi[k] = k < P ? x.pr[k].qinv : 0u; }
        for (int i = tid; i < count; i += kBThreads) r = min(r, x.cand[i]);   // deterministic start: lowest row
        r = __reduce_min_sync(0xffffffffu, r);
        if (lane == 0) atomicMin(&x.S.low, r);
        __syncthreads();
        r = x.S.low;
        for (;;) {                                             // rounds: every candidate against the current best
            __syncthreads();
            if (tid == 0) { x.S.low = INT_MAX; x.S.low2 = INT_MAX; }
            __syncthreads();
            for (int e = tid; e < count; e += kBThreads) {
                const int k = x.cand[e];
                if (k == r) continue;
                uint32_t t[PMAX];
#pragma unroll
                for (int kk = 0; kk < PMAX; ++kk) {
                    uint32_t v = 0u;
                    if (kk < P) {
                        const uint32_t xx = mmul(x.rhsres[kk * mr + k], x.colres[kk * mr + r], pq[kk], pqi[kk]);
                        const uint32_t yy = mmul(x.rhsres[kk * mr + r], x.colres[kk * mr + k], pq[kk], pqi[kk]);
                        v = mmul(mmul(msub(xx, yy, pq[kk]), x.dinv_m[kk], pq[kk], pqi[kk]), x.pr[kk].r2, pq[kk], pqi[kk]);
                    }
                    t[kk] = v;
                }
                int sg, tp;
                garner_num<PMAX>(t, P, pq, pqi, x.inv, x.half, sg, tp);
                sg *= sgnD;                                    // sign of ratio_k - ratio_r
                if (sg < 0) atomicMin(&x.S.low, k);            // strictly smaller ratio
                else if (sg == 0) atomicMin(&x.S.low2, k);     // equal ratio
            }
            __syncthreads();
            const int lt = x.S.low, eq = x.S.low2;
            if (lt != INT_MAX) { r = lt; continue; }           // restart from a strictly smaller row
            r = min(r, eq);                                    // lowest row among the equal minima (simplex.py:277)
            break;
        }
    }
    __syncthreads();
    return r;
}

// Tableau.pivot(r, c) on the staged column (tableau.py:295-308 in fraction-free form): per-prime
// constants, multiplier snapshot, update of the local planes, D <- p and its inverse for the next pivot.
// Returns false (nothing modified) when a prime divides the pivot element.
__device__ __noinline__ bool b_apply(BCtx& x, int r) {
    const int P = x.a.P, m = x.a.m, ld = x.a.ld, mr = x.a.mr, tid = threadIdx.x;
    if (tid == 0) x.S.flag = -1;
    __syncthreads();
    for (int k = tid; k < P; k += kBThreads) {
        const PrimeQ p = x.pr[k];
        const uint32_t piv = x.colres[k * mr + r];
        if (piv == 0u) atomicMax(&x.S.flag, k);
        x.alpha_m[k] = mmul(mmul(piv, p.r2, p.q, p.qinv), x.dinv_m[k], p.q, p.qinv);
    }
    __syncthreads();
    if (x.S.flag >= 0) return false;
    for (int e = tid; e < x.nloc * mr; e += kBThreads) {
        const int kl = e / mr, i = e - kl * mr, k = x.k0 + kl;
        uint32_t nb = 0u;
        if (i <= m && i != r) {
            const PrimeQ p = x.pr[k];
            const uint32_t beta = mmul(mmul(x.colres[k * mr + i], p.r2, p.q, p.qinv), x.dinv_m[k], p.q, p.qinv);
            nb = beta ? p.q - beta : 0u;
        }
        x.nbeta[e] = nb;
    }
    b_csync(x);                                                // every CTA has finished reading remote planes
    // update: items = (local prime, group of 4 rows, column); a row is written by its own items only and the
    // pivot row is never written, so the items are independent
    const int W = x.S.rc + 1, groups = mr / 4, items = x.nloc * groups * W;
    for (int it = tid; it < items; it += kBThreads) {
        const int gg = it / W, j = it - gg * W;
        const int kl = gg / groups, i0 = (gg - kl * groups) * 4, k = x.k0 + kl;
        const uint32_t qq = x.pr[k].q, qinv = x.pr[k].qinv, al = x.alpha_m[k];
        uint32_t* pl = x.planes + (size_t)kl * x.plane_sz;
        const uint32_t v = pl[(size_t)r * ld + j];
        const uint4 nb4 = *reinterpret_cast<const uint4*>(&x.nbeta[kl * mr + i0]);
        const uint32_t nbv[4] = {nb4.x, nb4.y, nb4.z, nb4.w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + u;
            if (i <= m && i != r) {
                uint32_t* cellp = pl + (size_t)i * ld + j;
                *cellp = redc((uint64_t)al * *cellp + (uint64_t)nbv[u] * v, qq, qinv);
            }
        }
    }
    for (int k = tid; k < P; k += kBThreads) {                 // D <- p; inverse for the next pivot
        const PrimeQ p = x.pr[k];
        const uint32_t piv = x.colres[k * mr + r];
        x.Dres[k] = piv;
        x.dinv_m[k] = minv_mont(mmul(piv, p.r2, p.q, p.qinv), p);
    }
    if (tid == 0) x.S.sgnD = (int)x.col_sign[r];
    b_csync(x);                                                // planes stable again
    return true;
}

__device__ __forceinline__ void b_commit(BCtx& x, int r, int c, int rule) {   // Simplex._pivot simplex.py:192-197
    if (threadIdx.x == 0) {
        BState& S = x.S;
        if (x.qr == 0 && x.a.trace && S.npiv_total < x.a.trace_cap) {
            int* tr = x.a.trace + ((size_t)x.lp * x.a.trace_cap + S.npiv_total) * 3;
            tr[0] = r; tr[1] = c; tr[2] = rule;
        }
        x.basis[r] = c;
        S.npiv_total++;
        S.npiv[S.phase]++;
    }
}

// Simplex.solve (simplex.py:110-148) on the current phase's tableau
template <int PMAX>
__device__ __noinline__ void b_solve(BCtx& x) {
    BState& S = x.S;
    const int tid = threadIdx.x, m = x.a.m;
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
        const int rule = S.rule;
        b_stage<PMAX>(x, 0, -1);
        const int c = b_price(x, rule);
        if (c < 0) return;                                     // optimal for this phase
        b_stage<PMAX>(x, 1, c);
        const int r = b_ratio<PMAX>(x);
        if (r < 0) {
            if (tid == 0) S.status = LPS_UNBOUNDED;
            __syncthreads();
            return;
        }
        const int sgn_br = (int)x.rhs_sign[r] * S.sgnD;
        __syncthreads();
        if (!b_apply(x, r)) {
            if (tid == 0) S.status = LPS_BAD_PRIME;
            __syncthreads();
            return;
        }
        b_commit(x, r, c, rule);
        if (tid == 0 && rule == LPS_RULE_DANTZIG) {
            S.pending = 1;
            if (sgn_br > 0) S.moved = 1;
            else if (sgn_br < 0 && !S.moved) S.status = LPS_OBJ_INCREASED;
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
template <int PMAX>
__global__ void __launch_bounds__(kBThreads) k_bnb_nodes(BatchParams a) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ BCtx x;

    const int C = a.C, P = a.P, PL = a.PL, m = a.m, n = a.n, ld = a.ld, mr = a.mr;
    const int tid = threadIdx.x, lane = tid & 31;
    (void)lane;
    if (tid == 0) {
        const BatchLayout L = batch_layout(m, P, PL, ld, mr);
        x.a = a;
        x.qr = (C > 1) ? (int)cg::this_cluster().block_rank() : 0;
        x.lp = blockIdx.x / C;
        x.k0 = x.qr * PL;
        x.nloc = max(0, min(PL, P - x.k0));
        x.plane_sz = (size_t)(m + 1) * ld;
        x.planes = reinterpret_cast<uint32_t*>(smem + L.planes);
        x.objdig = reinterpret_cast<uint32_t*>(smem + L.objdig);
        x.colres = reinterpret_cast<uint32_t*>(smem + L.colres);
        x.rhsres = reinterpret_cast<uint32_t*>(smem + L.rhsres);
        x.nbeta = reinterpret_cast<uint32_t*>(smem + L.nbeta);
        x.inv = reinterpret_cast<uint32_t*>(smem + L.inv);
        x.half = reinterpret_cast<uint32_t*>(smem + L.half);
        x.Dres = reinterpret_cast<uint32_t*>(smem + L.dres);
        x.dinv_m = reinterpret_cast<uint32_t*>(smem + L.dinv);
        x.alpha_m = reinterpret_cast<uint32_t*>(smem + L.alpha);
        x.obj_top = reinterpret_cast<int*>(smem + L.obj_top);
        x.col_top = reinterpret_cast<int*>(smem + L.col_top);
        x.rhs_top = reinterpret_cast<int*>(smem + L.rhs_top);
        x.basis = reinterpret_cast<int*>(smem + L.basis);
        x.cand = reinterpret_cast<int*>(smem + L.cand);
        x.wexp = reinterpret_cast<int*>(smem + L.wexp);
        x.slots = reinterpret_cast<int*>(smem + L.slots);
        x.wmant = reinterpret_cast<double*>(smem + L.wmant);
        x.col_mant = reinterpret_cast<double*>(smem + L.col_mant);
        x.rhs_mant = reinterpret_cast<double*>(smem + L.rhs_mant);
        x.pr = reinterpret_cast<PrimeQ*>(smem + L.pr);
        x.obj_sign = reinterpret_cast<signed char*>(smem + L.obj_sign);
        x.col_sign = reinterpret_cast<signed char*>(smem + L.col_sign);
        x.rhs_sign = reinterpret_cast<signed char*>(smem + L.rhs_sign);
        x.G = a.grids + (size_t)x.lp * (m + 1) * (n + 1);
        BState& S0 = x.S;
        S0.status = LPS_OK; S0.phase = 0; S0.K = 0; S0.rule = LPS_RULE_DANTZIG; S0.moved = 0; S0.pending = 0; S0.stuck = 0;
        S0.npiv_total = 0; S0.npiv[0] = S0.npiv[1] = S0.npiv[2] = 0; S0.sgnD = 1; S0.flag = -1; S0.ncur = n; S0.rc = n;
    }
    __syncthreads();
    BState& S = x.S;
    const int32_t* G = x.G;
    int* basis = x.basis;
    int* rowsign = x.cand;
    const int k0 = x.k0, nloc = x.nloc, qr = x.qr, lp = x.lp;
    const size_t plane_sz = x.plane_sz;
    uint32_t* planes = x.planes;

    // ---- tables ---------------------------------------------------------------------------------
    for (int k = tid; k < P; k += kBThreads) { x.pr[k] = a.pr[k]; x.half[k] = a.half[k]; x.Dres[k] = 1u; x.dinv_m[k] = a.pr[k].r1; }
    for (int k = tid; k <= P; k += kBThreads) { x.wmant[k] = a.wmant[k]; x.wexp[k] = a.wexp[k]; }
    for (int e = tid; e < P * P; e += kBThreads) x.inv[e] = a.invtab[e];

    // ---- Simplex._find_bfs, part 1 (simplex.py:42-60) on the integer input --------------------------
    // rows with b_i < 0 are negated; isCanonical: column j is basic for row i when c_j == 0 and its
    // only nonzero is an exact 1 in row i (the sign flip applied); bcols[i] = lowest such j.
    for (int i = tid; i < m; i += kBThreads) { basis[i] = INT_MAX; rowsign[i] = (G[(size_t)i * (n + 1) + n] < 0) ? -1 : 1; }
    __syncthreads();
    for (int j = tid; j < n; j += kBThreads) {
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

    // ---- encode into residues (local primes only) ------------------------------------------------
    // phase-1 tableau when K > 0: costs 0, z 0, artificial unit columns with cost 1, then
    // rowSubFromObj(i, 1) for every artificial row: objective row = -(sum of those rows) on the original
    // columns and the rhs, 0 on the artificial columns (simplex.py:49-60)
    if (S.status == LPS_OK) {
        const int W = n + K + 1;
        for (int e = tid; e < (m + 1) * W; e += kBThreads) {
            const int i = e / W, j = e - i * W;
            long long v = 0;
            if (i < m) {
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
            for (int kl = 0; kl < nloc; ++kl) {
                const uint32_t qq = x.pr[k0 + kl].q;
                uint32_t res = (uint32_t)(mag % qq);
                if (v < 0 && res) res = qq - res;
                planes[(size_t)kl * plane_sz + (size_t)i * ld + j] = res;
            }
        }
    }
    b_csync(x);                                                // all planes of the cluster are encoded

    // ---- the constructor's phase 1 -----------------------------------------------------------------
    if (S.status == LPS_OK && K > 0) {
        b_solve<PMAX>(x);                                      // simplex.py:62
        if (S.status == LPS_OK) {
            b_stage<PMAX>(x, 2, -1);
            if (x.rhs_sign[m] != 0) {                          // artificial optimum != 0 (simplex.py:63-67)
                __syncthreads();
                if (tid == 0) S.status = LPS_INFEASIBLE;
                __syncthreads();
            }
        }
        if (S.status == LPS_OK) {                              // drive basic artificials out (simplex.py:68-84)
            if (tid == 0) S.phase = 1;
            __syncthreads();
            int round = 0;
            bool redundant = false;
            for (int i = 0; i < m && S.status == LPS_OK; ++i) {
                if (basis[i] < n) continue;
                // first original column with a nonzero entry in row i: nonzero modulo some prime
                int first = INT_MAX;
                for (int j = tid; j < n; j += kBThreads) {
                    bool nz = false;
                    for (int kl = 0; kl < nloc; ++kl) nz |= planes[(size_t)kl * plane_sz + (size_t)i * ld + j] != 0u;
                    if (nz) { first = j; break; }
                }
                first = __reduce_min_sync(0xffffffffu, first);
                if (tid == 0) S.low = INT_MAX;
                __syncthreads();
                if ((tid & 31) == 0 && first != INT_MAX) atomicMin(&S.low, first);
                __syncthreads();
                int j1 = S.low;
                if (C > 1) {                                   // minimum over the cluster through the peers' slots
                    cg::cluster_group cluster = cg::this_cluster();
                    int* myslots = x.slots + (round & 1) * kBMaxCluster;
                    if (tid < C) cluster.map_shared_rank(myslots, tid)[qr] = j1;
                    cluster.sync();
                    for (int k = 0; k < C; ++k) j1 = min(j1, myslots[k]);
                    ++round;
                }
                __syncthreads();
                if (j1 == INT_MAX) { redundant = true; continue; }   // dependent row: marked for deletion (simplex.py:82-84)
                b_stage<PMAX>(x, 1, j1);
                if (!b_apply(x, i)) {
                    if (tid == 0) S.status = LPS_BAD_PRIME;
                    __syncthreads();
                    break;
                }
                b_commit(x, i, j1, 2);
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
            // (simplex.py:96-105): N_mj = D c_j - sum_i c_B(i) N_ij, the rhs column included
            const int rc_old = S.rc;
            for (int e = tid; e < nloc * (m + 1); e += kBThreads) {
                const int kl = e / (m + 1), i = e - kl * (m + 1);
                uint32_t* row = planes + (size_t)kl * plane_sz + (size_t)i * ld;
                row[n] = row[rc_old];
            }
            __syncthreads();
            auto to_mont = [&](long long v, const PrimeQ& p) -> uint32_t {
                const unsigned long long mag = v < 0 ? (unsigned long long)(-v) : (unsigned long long)v;
                uint32_t res = (uint32_t)(mag % p.q);
                if (v < 0 && res) res = p.q - res;
                return mmul(res, p.r2, p.q, p.qinv);
            };
            for (int e = tid; e < nloc * (n + 1); e += kBThreads) {
                const int kl = e / (n + 1), j = e - kl * (n + 1), k = k0 + kl;
                const PrimeQ p = x.pr[k];
                uint32_t* pl = planes + (size_t)kl * plane_sz;
                uint32_t acc = mmul(to_mont(G[(size_t)m * (n + 1) + j], p), x.Dres[k], p.q, p.qinv);
                for (int i = 0; i < m; ++i) {
                    const int cb = G[(size_t)m * (n + 1) + basis[i]];
                    if (cb == 0) continue;
                    acc = msub(acc, mmul(to_mont(cb, p), pl[(size_t)i * ld + j], p.q, p.qinv), p.q);
                }
                pl[(size_t)m * ld + j] = acc;
            }
            if (tid == 0) { S.ncur = n; S.rc = n; S.phase = 2; }
            b_csync(x);
        }
    }
    // ---- the optimisation proper -------------------------------------------------------------------
    if (S.status == LPS_OK) {
        b_solve<PMAX>(x);
        if (S.status == LPS_OK) {
            __syncthreads();
            if (tid == 0) S.status = LPS_OPTIMAL;
            __syncthreads();
        }
    }

    // ---- results: rhs column (b_0..b_{m-1}, z cell) and D as sign + mixed-radix digits -------------
    const int st = S.status;
    if (st != LPS_NEED_ARTIFICIALS) {
        b_stage<PMAX>(x, 2, -1);
        b_stage<PMAX>(x, 3, -1);
        if (qr == 0) {
            signed char* so = a.sign + (size_t)lp * (m + 2);
            uint32_t* dg = a.digits + (size_t)lp * P * (m + 2);
            for (int i = tid; i <= m + 1; i += kBThreads) {
                so[i] = i <= m ? x.rhs_sign[i] : (signed char)S.flag;
                for (int k = 0; k < P; ++k) dg[(size_t)k * (m + 2) + i] = x.objdig[k * ld + i];
            }
        }
    }
    if (qr == 0) {
        for (int i = tid; i < m; i += kBThreads) a.basis[(size_t)lp * m + i] = basis[i];
        if (tid == 0) {
            a.status[lp] = st;
            a.npiv[lp * 4 + 0] = S.npiv[0]; a.npiv[lp * 4 + 1] = S.npiv[1]; a.npiv[lp * 4 + 2] = S.npiv[2];
            a.npiv[lp * 4 + 3] = S.K;
        }
    }
    b_csync(x);                                                // nobody leaves while a peer may still read its planes
}

}  // namespace lps
