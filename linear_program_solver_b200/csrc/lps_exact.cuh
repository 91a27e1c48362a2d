// Exact-mode simplex pivot kernels for sm_100a: the tableau is a fraction-free (Bareiss) integer
// matrix N with one common denominator D, held as residues modulo P 31-bit primes
// (R[k][i][j] = N_ij mod p_k, one prime per grid slice), so that the reference's Fraction entry is
// N_ij / (D * sigma_i)  (sigma_i: a positive per-row scale that only rows never pivoted since the
// upload keep; 1 for integer inputs).
//
//   pivot (r,c), p = N_rc:   N_ij <- (p N_ij - N_ic N_rj) / D   for i != r,  row r unchanged,  D <- p
//   modulo p_k the division is a multiplication by D^-1 (SURVEY.md section 8a note).
//
// Exact comparisons come from mixed-radix (Garner) reconstruction of only the objective row, the
// entering column and the right-hand side each iteration:
//   pricing     sign and magnitude of N_mj (common denominator D sigma_m): digits compared exactly
//   ratio test  N_in / N_ic: a rigorous floating filter on the leading digits picks the candidates,
//               ties / near-ties are decided exactly by the sign of (N_kn N_bc - N_bn N_kc) / D,
//               which is itself a tableau entry of a neighbouring basis and therefore fits the
//               prime set
//
// Reference semantics reproduced (file:line under /root/reference/lpsol): tableau.py:295-308 (pivot),
// simplex.py:262-267 / :228-232 (pricing, lowest index among equals), :272-279 (ratio test, lowest row
// among equals), :118-137 (stuck counter against the objective at entry, rule switch).
#pragma once
#include <math.h>

#include "lps_common.cuh"
#include "../../include/lps_b200.h"

namespace lps {

constexpr int kMaxSPT = 4;            // primes per thread in the lane-per-prime kernels (P <= 4096)
constexpr int kLaneThreads = 1024;    // block size of the single-block decision kernels
constexpr int kXThreads = 256;        // block size of the update kernel
constexpr int kXMaxRowBlock = 128;

struct PrimeQ {
    uint32_t q;      // the prime, odd, < 2^31
    uint32_t qinv;   // -q^-1 mod 2^32
    uint32_t r1;     // 2^32 mod q  (Montgomery form of 1)
    uint32_t r2;     // 2^64 mod q  (to Montgomery form: mmul(a, r2) = a 2^32)
};

// Montgomery reduction: T < q 2^32  ->  T 2^-32 mod q
__host__ __device__ __forceinline__ uint32_t redc(uint64_t T, uint32_t q, uint32_t qinv) {
    const uint32_t mlo = (uint32_t)T * qinv;
    const uint32_t t = (uint32_t)((T + (uint64_t)mlo * q) >> 32);
    return t >= q ? t - q : t;
}
__host__ __device__ __forceinline__ uint32_t mmul(uint32_t a, uint32_t b, uint32_t q, uint32_t qinv) {
    return redc((uint64_t)a * b, q, qinv);
}
__host__ __device__ __forceinline__ uint32_t msub(uint32_t a, uint32_t b, uint32_t q) { return a >= b ? a - b : a + q - b; }
// a^-1 mod q for a in Montgomery form, result in Montgomery form (Fermat); a != 0
__host__ __device__ inline uint32_t minv_mont(uint32_t a_m, const PrimeQ& p) {
    uint32_t result = p.r1, base = a_m;
    uint32_t e = p.q - 2;
    while (e) {
        if (e & 1u) result = mmul(result, base, p.q, p.qinv);
        base = mmul(base, base, p.q, p.qinv);
        e >>= 1;
    }
    return result;
}

struct XGrid {
    uint32_t* R;             // [P][(m+1)][ld]
    size_t ld, plane;        // plane = (m+1) * ld
    int m, n, P;             // P = primes of the WORKING set (a prefix of the allocated set, see XState::wlimit)
    int Pld;                 // allocated primes = row stride of invtab
    const PrimeQ* pr;        // [Pld]
    const uint32_t* invtab;  // [Pld][Pld]: invtab[j*Pld+k] = Montgomery form of p_j^-1 mod p_k  (j < k)
    const uint32_t* half;    // [P] mixed-radix digits of (M-1)/2
    const double* wmant;     // [P+1] W_t = prod_{k<t} p_k = wmant[t] * 2^wexp[t], wmant in [0.5,1)
    const int* wexp;
};

struct XState {
    int status;              // LPS_OK while running
    int rule;                // Dantzig until the stuck counter trips
    long long stuck;
    long long npiv, max_pivots;
    int moved;               // some committed pivot lowered the objective (b_r > 0)
    int pending_zcheck;      // last committed pivot was Dantzig-priced
    int dry_run;             // find_pivot: select only
    int c, r;                // current entering column / leaving row
    int sgnD;                // sign of the common denominator
    int sgn_p, sgn_br;       // sign of N_rc and of b_r for the pivot being applied
    int sel_rule;            // rule that priced the current pivot (2 = explicit)
    int last_r, last_c;
    int bad_prime;           // index of a prime dividing the pivot, or -1
    int obj_increased;       // a Dantzig pivot raised the objective above its value at entry (simplex.py:133)
    // Working prime set.  Every entry of the tableau (and D) is, up to sign, a minor of the uploaded integer matrix
    // whose columns are the last pivot columns of the rows pivoted since the upload plus at most one more column
    // (two for the tie-break quantity, an entry one pivot away), so |entry| <= product of those columns' 2-norms
    // (Hadamard).  wsum = sum over those rows of w[column], w = ceil(256 log2 norm); a pivot is applied only while
    // wnext <= wlimit = 256 (bits of the working set) - 2 max w - guard; otherwise LPS_NEED_PRIMES is reported with
    // nothing modified and the host extends the working set (k_xextend) before the loop continues.
    long long wsum, wnext, wlimit;
};

struct XMeta {               // per reconstructed number
    signed char* sign;       // -1 / 0 / +1
    int* top;                // index of the leading nonzero mixed-radix digit of |x|, -1 for 0
    double* mant;            // |x| ~ mant * W_top  (three leading digits)
};

// ------------------------------------------------------------------------------------------------
// Garner mixed-radix conversion, lane-per-prime inside one block
// ------------------------------------------------------------------------------------------------
template <int NB>
struct GarnerShared {
    uint32_t v[2][NB];
    uint32_t done[3];
    int hi[NB], lo[NB], top[NB], neg[NB];
    uint32_t d3[NB][3];
};

// t[s][b] holds the residue of number b modulo prime k = threadIdx.x + s*blockDim.x (anything for
// k >= P).  On return t[s][b] is digit k of |x_b| in the mixed radix (p_0, p_1, ...), and S holds per
// number: lo (INT_MAX for zero), neg, top, d3 (three leading digits of |x|).
template <int NB, int SPT>
__device__ __forceinline__ void block_garner(const XGrid& g, uint32_t (&t)[SPT][NB], GarnerShared<NB>& S) {
    const int bd = blockDim.x, tid = threadIdx.x, P = g.P;
    uint32_t q[SPT], qi[SPT];
#pragma unroll
    for (int s = 0; s < SPT; ++s) {
        const int k = tid + s * bd;
        q[s] = k < P ? g.pr[k].q : 3u;
        qi[s] = k < P ? g.pr[k].qinv : 0u;
    }
    // Early exit: after round j-1 lane k >= j holds (x - X_j) / (p_0 ... p_{j-1}) mod p_k, X_j the number made of the
    // digits found so far.  When every such lane holds 0 the quotient is 0 (it is smaller than the product of the
    // remaining primes), when every lane holds p_k - 1 it is -1 (x negative, |x| below the product of the first j
    // primes); in both cases the lanes already hold their final digits and the remaining rounds would not change
    // them.  Tableau entries are usually far below the Hadamard bound that sized the prime set, so the conversion
    // costs O(P * digits of |x|) instead of O(P^2).  Checked every fourth round (one warp vote + one shared atomic).
    if (tid == 0) S.done[0] = S.done[1] = S.done[2] = 0xffffffffu;
    __syncthreads();
    for (int j = 0; j < P; ++j) {
        const int os = j / bd, ot = j - os * bd;
        const bool vote = (j & 3) == 0;
        if (vote) {
            uint32_t mask = 0xffffffffu;                       // bit b: my lanes of number b are all 0; bit NB+b: all p_k - 1
#pragma unroll
            for (int s = 0; s < SPT; ++s) {
                const int k = tid + s * bd;
                if (k >= j && k < P) {
#pragma unroll
                    for (int b = 0; b < NB; ++b) {
                        if (t[s][b] != 0u) mask &= ~(1u << b);
                        if (t[s][b] != q[s] - 1u) mask &= ~(1u << (NB + b));
                    }
                }
            }
            mask = __reduce_and_sync(0xffffffffu, mask);
            if ((tid & 31) == 0 && mask != 0xffffffffu) atomicAnd(&S.done[(j >> 2) % 3], mask);
        }
        if (tid == ot) {
#pragma unroll
            for (int s = 0; s < SPT; ++s)
                if (s == os) {
#pragma unroll
                    for (int b = 0; b < NB; ++b) S.v[j & 1][b] = t[s][b];
                }
        }
        __syncthreads();
        if (vote) {
            const uint32_t d = S.done[(j >> 2) % 3];
            if ((((d | (d >> NB)) & ((1u << NB) - 1u)) == ((1u << NB) - 1u))) break;
            if (tid == 0) S.done[((j >> 2) + 2) % 3] = 0xffffffffu;   // slot of the vote after next; last read before this round's barrier
        }
        const uint32_t* invrow = g.invtab + (size_t)j * g.Pld;
#pragma unroll
        for (int s = 0; s < SPT; ++s) {
            const int k = tid + s * bd;
            if (k > j && k < P) {
                const uint32_t inv = __ldg(invrow + k);
#pragma unroll
                for (int b = 0; b < NB; ++b) {
                    uint32_t v = S.v[j & 1][b];
                    if (v >= q[s]) v %= q[s];
                    t[s][b] = mmul(msub(t[s][b], v, q[s]), inv, q[s], qi[s]);
                }
            }
        }
    }
    __syncthreads();
    if (tid < NB) {
        S.hi[tid] = -1; S.lo[tid] = INT_MAX; S.top[tid] = -1; S.neg[tid] = 0;
        S.d3[tid][0] = S.d3[tid][1] = S.d3[tid][2] = 0;
    }
    __syncthreads();
    const int lane = tid & 31;
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        int hi = -1, lo = INT_MAX;
#pragma unroll
        for (int s = 0; s < SPT; ++s) {
            const int k = tid + s * bd;
            if (k < P) {
                if (t[s][b] != g.half[k]) hi = max(hi, k);
                if (t[s][b] != 0u) lo = min(lo, k);
            }
        }
        hi = __reduce_max_sync(0xffffffffu, hi);
        lo = __reduce_min_sync(0xffffffffu, lo);
        if (lane == 0) {
            if (hi >= 0) atomicMax(&S.hi[b], hi);
            if (lo != INT_MAX) atomicMin(&S.lo[b], lo);
        }
    }
    __syncthreads();
#pragma unroll
    for (int b = 0; b < NB; ++b) {
#pragma unroll
        for (int s = 0; s < SPT; ++s) {
            const int k = tid + s * bd;
            if (k < P && k == S.hi[b]) S.neg[b] = t[s][b] > g.half[k];
        }
    }
    __syncthreads();
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        const int neg = S.neg[b], lo = S.lo[b];
        int top = -1;
#pragma unroll
        for (int s = 0; s < SPT; ++s) {
            const int k = tid + s * bd;
            if (k < P) {
                uint32_t a = t[s][b];
                if (lo == INT_MAX) a = 0u;
                else if (neg) a = k < lo ? 0u : (k == lo ? q[s] - a : q[s] - 1u - a);
                t[s][b] = a;
                if (a) top = max(top, k);
            }
        }
        top = __reduce_max_sync(0xffffffffu, top);
        if (lane == 0 && top >= 0) atomicMax(&S.top[b], top);
    }
    __syncthreads();
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        const int tp = S.top[b];
#pragma unroll
        for (int s = 0; s < SPT; ++s) {
            const int k = tid + s * bd;
            if (k < P && tp >= 0 && k <= tp && k >= tp - 2) S.d3[b][tp - k] = t[s][b];
        }
    }
    __syncthreads();
}

template <int NB>
__device__ __forceinline__ int garner_sign(const GarnerShared<NB>& S, int b) {
    return S.lo[b] == INT_MAX ? 0 : (S.neg[b] ? -1 : 1);
}
template <int NB>
__device__ __forceinline__ double garner_mant(const XGrid& g, const GarnerShared<NB>& S, int b) {
    const int tp = S.top[b];
    if (tp < 0) return 0.0;
    double x = (double)S.d3[b][0];
    if (tp >= 1) {
        double y = (double)S.d3[b][1];
        if (tp >= 2) y += (double)S.d3[b][2] / (double)g.pr[tp - 2].q;
        x += y / (double)g.pr[tp - 1].q;
    }
    return x;
}

// ---- batch reconstruction of a row, a column, the (column c, rhs) pair, or custom residues -----------
enum { XSRC_ROW = 0, XSRC_COL = 1, XSRC_COLPAIR = 2, XSRC_CUSTOM = 3, XSRC_COLPAIRS = 4, XSRC_ROWS = 5 };
// 4: (column, rhs) pairs of columns index, index+1, ...;  5: rows index, index+1, ... (n+1 numbers each)
struct GarnerJob {
    int src;                 // XSRC_*
    int index;               // row / column; column -1 means "the entering column st->c"
    int count;               // numbers to reconstruct
    int check_status;        // skip when st->status != LPS_OK (solve loop)
    const uint32_t* custom;  // XSRC_CUSTOM: residue of number e modulo prime k at custom[k*cstride + e]
    size_t cstride;
    XMeta meta;
    uint32_t* dig;           // optional: digit k of |x_e| at dig[k*dstride + e]
    size_t dstride;
};

template <int NB, int SPT>
__global__ void __launch_bounds__(kLaneThreads) k_garner(XGrid g, const XState* st, GarnerJob job) {
    __shared__ GarnerShared<NB> S;
    if (job.check_status && st->status != LPS_OK) return;
    const int e0 = blockIdx.x * NB, bd = blockDim.x, tid = threadIdx.x;
    uint32_t t[SPT][NB];
    const int col = (job.index < 0) ? st->c : job.index;
#pragma unroll
    for (int s = 0; s < SPT; ++s) {
        const int k = tid + s * bd;
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            const int e = e0 + b;
            uint32_t v = 0u;
            if (k < g.P && e < job.count) {
                const uint32_t* plane = g.R + (size_t)k * g.plane;
                if (job.src == XSRC_ROW) v = plane[(size_t)job.index * g.ld + e];
                else if (job.src == XSRC_COL) v = plane[(size_t)e * g.ld + col];
                else if (job.src == XSRC_COLPAIR) v = (e <= g.m) ? plane[(size_t)e * g.ld + col] : plane[(size_t)(e - g.m - 1) * g.ld + g.n];
                else if (job.src == XSRC_COLPAIRS) {
                    const int per = 2 * (g.m + 1), jj = e / per, rem = e - jj * per;
                    v = (rem <= g.m) ? plane[(size_t)rem * g.ld + job.index + jj] : plane[(size_t)(rem - g.m - 1) * g.ld + g.n];
                }
                else if (job.src == XSRC_ROWS) { const int per = g.n + 1, ii = e / per; v = plane[(size_t)(job.index + ii) * g.ld + (e - ii * per)]; }
                else v = job.custom[(size_t)k * job.cstride + e];
            }
            t[s][b] = v;
        }
    }
    block_garner<NB, SPT>(g, t, S);
    if (tid < NB && e0 + tid < job.count) {
        const int e = e0 + tid;
        job.meta.sign[e] = (signed char)garner_sign<NB>(S, tid);
        job.meta.top[e] = S.top[tid];
        job.meta.mant[e] = garner_mant<NB>(g, S, tid);
    }
    if (job.dig) {
#pragma unroll
        for (int s = 0; s < SPT; ++s) {
            const int k = tid + s * bd;
            if (k < g.P) {
#pragma unroll
                for (int b = 0; b < NB; ++b)
                    if (e0 + b < job.count) job.dig[(size_t)k * job.dstride + e0 + b] = t[s][b];
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// The same conversion with ONE NUMBER PER WARP (P <= 32 * SPT): lane l holds the residues modulo primes l, l + 32, ...;
// the digit of a round travels by shuffle instead of through shared memory and a block barrier, so a round costs tens
// of cycles instead of a barrier round trip.  This is what the solve loop uses while the working prime set is short
// (k_garner_w); results are identical to block_garner's (digits of |x|, sign, leading digit, 3-digit mantissa).
// ------------------------------------------------------------------------------------------------
template <int SPT>
__device__ __forceinline__ void warp_garner(const XGrid& g, uint32_t (&t)[SPT], int& sign_out, int& top_out, double& mant_out) {
    const int lane = threadIdx.x & 31, P = g.P;
    uint32_t q[SPT], qi[SPT], hf[SPT];
#pragma unroll
    for (int s = 0; s < SPT; ++s) {
        const int k = lane + 32 * s;
        q[s] = k < P ? g.pr[k].q : 3u;
        qi[s] = k < P ? g.pr[k].qinv : 0u;
        hf[s] = k < P ? g.half[k] : 0u;
    }
    for (int j = 0; j < P; ++j) {
        if ((j & 3) == 0) {                                      // early exit, see block_garner
            bool zero = true, ones = true;
#pragma unroll
            for (int s = 0; s < SPT; ++s) {
                const int k = lane + 32 * s;
                if (k >= j && k < P) {
                    zero &= (t[s] == 0u);
                    ones &= (t[s] == q[s] - 1u);
                }
            }
            if (__all_sync(0xffffffffu, zero) || __all_sync(0xffffffffu, ones)) break;
        }
        const int os = j >> 5, ol = j & 31;
        uint32_t mine = 0u;
#pragma unroll
        for (int s = 0; s < SPT; ++s)
            if (s == os) mine = t[s];
        const uint32_t v0 = __shfl_sync(0xffffffffu, mine, ol);
        const uint32_t* invrow = g.invtab + (size_t)j * g.Pld;
#pragma unroll
        for (int s = 0; s < SPT; ++s) {
            const int k = lane + 32 * s;
            if (k > j && k < P) {
                uint32_t v = v0;
                if (v >= q[s]) v %= q[s];
                t[s] = mmul(msub(t[s], v, q[s]), __ldg(invrow + k), q[s], qi[s]);
            }
        }
    }
    // sign: compare with (M-1)/2 digit by digit from the top; lowest nonzero digit for the negation
    int hi = -1, lo = INT_MAX;
#pragma unroll
    for (int s = 0; s < SPT; ++s) {
        const int k = lane + 32 * s;
        if (k < P) {
            if (t[s] != hf[s]) hi = max(hi, k);
            if (t[s] != 0u) lo = min(lo, k);
        }
    }
    hi = __reduce_max_sync(0xffffffffu, hi);
    lo = __reduce_min_sync(0xffffffffu, lo);
    int negv = 0;
#pragma unroll
    for (int s = 0; s < SPT; ++s) {
        const int k = lane + 32 * s;
        if (k < P && k == hi) negv = t[s] > hf[s];
    }
    const int neg = __reduce_max_sync(0xffffffffu, negv);
    int top = -1;
#pragma unroll
    for (int s = 0; s < SPT; ++s) {
        const int k = lane + 32 * s;
        if (k < P) {
            uint32_t a = t[s];
            if (lo == INT_MAX) a = 0u;
            else if (neg) a = k < lo ? 0u : (k == lo ? q[s] - a : q[s] - 1u - a);
            t[s] = a;
            if (a) top = max(top, k);
        }
    }
    top = __reduce_max_sync(0xffffffffu, top);
    // the three leading digits, fetched from their lanes
    uint32_t d3[3] = {0u, 0u, 0u};
#pragma unroll
    for (int u = 0; u < 3; ++u) {
        const int kd = top - u;
        uint32_t mine = 0u;
        if (kd >= 0) {
#pragma unroll
            for (int s = 0; s < SPT; ++s)
                if (s == (kd >> 5)) mine = t[s];
        }
        d3[u] = kd >= 0 ? __shfl_sync(0xffffffffu, mine, kd & 31) : 0u;
    }
    double x = 0.0;
    if (top >= 0) {
        x = (double)d3[0];
        if (top >= 1) {
            double y = (double)d3[1];
            if (top >= 2) y += (double)d3[2] / (double)g.pr[top - 2].q;
            x += y / (double)g.pr[top - 1].q;
        }
    }
    sign_out = lo == INT_MAX ? 0 : (neg ? -1 : 1);
    top_out = top;
    mant_out = x;
}

constexpr int kGarnerWarps = 8;       // numbers per block of k_garner_w

// number e of a job, by the calling warp
template <int SPT>
__device__ __forceinline__ void garner_number(const XGrid& g, const XState* st, const GarnerJob& job, int e) {
    const int lane = threadIdx.x & 31;
    const int col = (job.index < 0) ? st->c : job.index;
    uint32_t t[SPT];
#pragma unroll
    for (int s = 0; s < SPT; ++s) {
        const int k = lane + 32 * s;
        uint32_t v = 0u;
        if (k < g.P) {
            const uint32_t* plane = g.R + (size_t)k * g.plane;
            if (job.src == XSRC_ROW) v = plane[(size_t)job.index * g.ld + e];
            else if (job.src == XSRC_COL) v = plane[(size_t)e * g.ld + col];
            else if (job.src == XSRC_COLPAIR) v = (e <= g.m) ? plane[(size_t)e * g.ld + col] : plane[(size_t)(e - g.m - 1) * g.ld + g.n];
            else if (job.src == XSRC_COLPAIRS) {
                const int per = 2 * (g.m + 1), jj = e / per, rem = e - jj * per;
                v = (rem <= g.m) ? plane[(size_t)rem * g.ld + job.index + jj] : plane[(size_t)(rem - g.m - 1) * g.ld + g.n];
            } else if (job.src == XSRC_ROWS) {
                const int per = g.n + 1, ii = e / per;
                v = plane[(size_t)(job.index + ii) * g.ld + (e - ii * per)];
            } else v = job.custom[(size_t)k * job.cstride + e];
        }
        t[s] = v;
    }
    int sg, top;
    double mant;
    warp_garner<SPT>(g, t, sg, top, mant);
    if (lane == 0) {
        job.meta.sign[e] = (signed char)sg;
        job.meta.top[e] = top;
        job.meta.mant[e] = mant;
    }
    if (job.dig) {
#pragma unroll
        for (int s = 0; s < SPT; ++s) {
            const int k = lane + 32 * s;
            if (k < g.P) job.dig[(size_t)k * job.dstride + e] = t[s];
        }
    }
}

template <int SPT>
__global__ void __launch_bounds__(32 * kGarnerWarps) k_garner_w(XGrid g, const XState* st, GarnerJob job) {
    if (job.check_status && st->status != LPS_OK) return;
    const int e = blockIdx.x * kGarnerWarps + (threadIdx.x >> 5);
    if (e >= job.count) return;                                  // whole warps leave
    garner_number<SPT>(g, st, job, e);
}

// the same inside a single block that goes on to use the numbers (small tableaux: one launch instead of two)
template <int SPT>
__device__ __forceinline__ void garner_job_in_block(const XGrid& g, const XState* st, const GarnerJob& job) {
    for (int e = threadIdx.x >> 5; e < job.count; e += blockDim.x >> 5) garner_number<SPT>(g, st, job, e);
    __syncthreads();                                             // the block reads what its warps wrote
}

// ------------------------------------------------------------------------------------------------
// Extension of the working prime set from g.P to Pnew primes: every number is reconstructed from its residues
// modulo the first g.P primes (Garner, with the early exit) and evaluated modulo each new prime by Horner's rule
// on its mixed-radix digits, x = d_0 + p_0 (d_1 + p_1 (d_2 + ...)).  Source and destination are the residue
// planes (vec == NULL: entry e = i (n+1) + j of the grid, rows 0..m, columns 0..n) or a vector with one number
// per prime (vec[k], e.g. the denominator D).  Dynamic shared memory: NB * g.P words.
// ------------------------------------------------------------------------------------------------
template <int NB, int SPT>
__global__ void __launch_bounds__(kLaneThreads) k_xextend(XGrid g, int Pnew, uint32_t* vec, long long count) {
    extern __shared__ uint32_t s_dig[];
    __shared__ GarnerShared<NB> S;
    const int bd = blockDim.x, tid = threadIdx.x, P = g.P, ncol = g.n + 1;
    const long long e0 = (long long)blockIdx.x * NB;
    uint32_t t[SPT][NB];
#pragma unroll
    for (int s = 0; s < SPT; ++s) {
        const int k = tid + s * bd;
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            const long long e = e0 + b;
            uint32_t v = 0u;
            if (k < P && e < count) {
                if (vec) v = vec[k];
                else {
                    const long long i = e / ncol, j = e - i * ncol;
                    v = g.R[(size_t)k * g.plane + (size_t)i * g.ld + j];
                }
            }
            t[s][b] = v;
        }
    }
    block_garner<NB, SPT>(g, t, S);
#pragma unroll
    for (int s = 0; s < SPT; ++s) {
        const int k = tid + s * bd;
        if (k < P) {
#pragma unroll
            for (int b = 0; b < NB; ++b) s_dig[b * P + k] = t[s][b];
        }
    }
    __syncthreads();
    const int dP = Pnew - P;
    for (int idx = tid; idx < NB * dP; idx += bd) {
        const int b = idx / dP, kk = P + (idx - b * dP);
        const long long e = e0 + b;
        if (e >= count) continue;
        const PrimeQ p = g.pr[kk];
        const int top = S.top[b];
        uint32_t acc = 0u;
        if (top >= 0) {
            const uint32_t* dg = s_dig + b * P;
            acc = dg[top] % p.q;
            for (int u = top - 1; u >= 0; --u) {
                const uint32_t ptm = mmul(g.pr[u].q % p.q, p.r2, p.q, p.qinv);   // Montgomery form of p_u mod q
                acc = mmul(acc, ptm, p.q, p.qinv) + dg[u] % p.q;                 // acc p_u + d_u
                if (acc >= p.q) acc -= p.q;
            }
            if (S.neg[b] && acc) acc = p.q - acc;
        }
        if (vec) vec[kk] = acc;
        else {
            const long long i = e / ncol, j = e - i * ncol;
            g.R[(size_t)kk * g.plane + (size_t)i * g.ld + j] = acc;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// pricing on the reconstructed objective row (single block)
// ------------------------------------------------------------------------------------------------
// exact comparison of |x_a| and |x_b| from their mixed-radix digits: >0 if |x_a| > |x_b|
__device__ __forceinline__ int cmp_mag(const int* top, const uint32_t* dig, size_t dstride, int a, int b) {
    const int ta = top[a], tb = top[b];
    if (ta != tb) return ta > tb ? 1 : -1;
    for (int k = ta; k >= 0; --k) {
        const uint32_t da = dig[(size_t)k * dstride + a], db = dig[(size_t)k * dstride + b];
        if (da != db) return da > db ? 1 : -1;
    }
    return 0;
}

struct XPriceArgs {
    XGrid g;
    XState* st;
    XMeta obj;               // objective row, n+1 numbers
    const uint32_t* dig;     // its digits
    size_t dstride;
};

__device__ __forceinline__ void xprice_body(const XPriceArgs& a) {
    __shared__ int s_best[kLaneThreads];
    XState* st = a.st;
    const XGrid& g = a.g;
    if (st->status != LPS_OK) return;
    const int tid = threadIdx.x;
    if (!st->dry_run && st->obj_increased) {
        __syncthreads();
        if (tid == 0) st->status = LPS_OBJ_INCREASED;
        return;
    }
    if (!st->dry_run && st->max_pivots >= 0 && st->npiv >= st->max_pivots) {
        __syncthreads();
        if (tid == 0) st->status = LPS_MAX_PIVOTS;
        return;
    }
    // stuck counter and rule (simplex.py:123,132-137).  In exact arithmetic the objective is
    // monotone, so "z equals its value at entry" holds exactly until the first pivot with b_r > 0.
    int rule = st->rule;
    long long stuck = st->stuck;
    if (!st->dry_run) {
        if (st->pending_zcheck) stuck = st->moved ? 0 : stuck + 1;
        if (rule == LPS_RULE_DANTZIG && !(stuck < (long long)g.m + g.n)) rule = LPS_RULE_BLAND;
    }
    const int sgnD = st->sgnD;
    int best = -1;
    for (int j = tid; j < g.n; j += blockDim.x) {
        if ((int)a.obj.sign[j] * sgnD >= 0) continue;            // c_j >= 0
        if (best < 0) best = j;
        else if (rule == LPS_RULE_DANTZIG && cmp_mag(a.obj.top, a.dig, a.dstride, j, best) > 0) best = j;
    }
    s_best[tid] = best;
    __syncthreads();
    for (int off = blockDim.x >> 1; off > 0; off >>= 1) {
        if (tid < off) {
            const int x = s_best[tid], y = s_best[tid + off];
            int pick;
            if (x < 0) pick = y;
            else if (y < 0) pick = x;
            else if (rule == LPS_RULE_BLAND) pick = min(x, y);
            else {
                const int c = cmp_mag(a.obj.top, a.dig, a.dstride, x, y);
                pick = c > 0 ? x : (c < 0 ? y : min(x, y));
            }
            s_best[tid] = pick;
        }
        __syncthreads();
    }
    if (tid == 0) {
        const int c = s_best[0];
        if (!st->dry_run) { st->rule = rule; st->stuck = stuck; st->pending_zcheck = 0; }
        st->sel_rule = rule;
        st->c = c;
        st->r = -1;
        if (c < 0) {
            st->status = LPS_OPTIMAL;
            st->last_r = -1; st->last_c = -1;
        }
    }
}
__global__ void __launch_bounds__(kLaneThreads) k_xprice(XPriceArgs a) { xprice_body(a); }
// small tableaux: the objective row is reconstructed by this block's warps first (one launch instead of two)
template <int SPTW>
__global__ void __launch_bounds__(kLaneThreads) k_xprice_fused(XPriceArgs a, GarnerJob job) {
    if (a.st->status != LPS_OK) return;
    garner_job_in_block<SPTW>(a.g, a.st, job);
    xprice_body(a);
}

// ------------------------------------------------------------------------------------------------
// per-prime constants of a pivot: D^-1, alpha = p D^-1 (Montgomery forms), divisibility check
// ------------------------------------------------------------------------------------------------
struct XPivotConst {
    uint32_t* Dres;          // [P] residues of D (plain)
    uint32_t* dinv_m;        // [P] Montgomery form of D^-1
    uint32_t* alpha_m;       // [P] Montgomery form of p D^-1
};

__device__ __forceinline__ void compute_dinv(const XGrid& g, const XPivotConst& pc) {
    for (int k = threadIdx.x; k < g.P; k += blockDim.x) {
        const PrimeQ p = g.pr[k];
        pc.dinv_m[k] = minv_mont(mmul(pc.Dres[k], p.r2, p.q, p.qinv), p);
    }
}
// after r and c are known; returns (block-uniform) the index of a prime dividing the pivot or -1
__device__ __forceinline__ int compute_alpha(const XGrid& g, const XPivotConst& pc, int r, int c, int* s_flag) {
    if (threadIdx.x == 0) *s_flag = -1;
    __syncthreads();
    for (int k = threadIdx.x; k < g.P; k += blockDim.x) {
        const PrimeQ p = g.pr[k];
        const uint32_t piv = g.R[(size_t)k * g.plane + (size_t)r * g.ld + c];
        if (piv == 0u) atomicMax(s_flag, k);
        pc.alpha_m[k] = mmul(mmul(piv, p.r2, p.q, p.qinv), pc.dinv_m[k], p.q, p.qinv);
    }
    __syncthreads();
    return *s_flag;
}

// ------------------------------------------------------------------------------------------------
// ratio test (single block): floating filter, exact tie-break, pivot constants
// ------------------------------------------------------------------------------------------------
struct RKey {                // value = cls * f * 2^e with f in [0.5,1); cls 2 = no candidate
    int cls, e;
    double f;
};
__device__ __forceinline__ bool rkey_less(const RKey& a, const RKey& b) {
    if (a.cls != b.cls) return a.cls < b.cls;
    if (a.cls == 0 || a.cls == 2) return false;
    const bool mag_less = (a.e != b.e) ? (a.e < b.e) : (a.f < b.f);
    const bool mag_more = (a.e != b.e) ? (a.e > b.e) : (a.f > b.f);
    return a.cls > 0 ? mag_less : mag_more;
}
// |a / b - 1| <= 2^-38 for keys of the same nonzero class
__device__ __forceinline__ bool rkey_close(const RKey& a, const RKey& b) {
    if (a.cls != b.cls) return false;
    const int de = a.e - b.e;
    if (de > 1 || de < -1) return false;
    const double d = ldexp(a.f / b.f, de);
    return fabs(d - 1.0) <= 3.637978807091713e-12;
}

struct XRatioArgs {
    XGrid g;
    XState* st;
    XMeta cr;                // 2(m+1) numbers: column c rows 0..m, then rhs rows 0..m
    XPivotConst pc;
    int* cand;               // m ints of scratch
    const int* wcol;         // [n+1] column weights, NULL: fixed prime set
    const int* xbasis;       // [m] last pivot column of each row since the upload, -1: never pivoted
    uint32_t* nbeta;         // fused variant: the stage (k_xstage) runs at the end of this block
    size_t nbeta_ld;
};

template <int SPT>
__device__ __forceinline__ void xratio_body(const XRatioArgs& a) {
    __shared__ GarnerShared<1> S;
    __shared__ int s_cls[kLaneThreads], s_e[kLaneThreads];
    __shared__ double s_f[kLaneThreads];
    __shared__ int s_count, s_int, s_flag;
    XState* st = a.st;
    const XGrid& g = a.g;
    if (st->status != LPS_OK) return;
    const int tid = threadIdx.x, c = st->c, sgnD = st->sgnD, m = g.m;
    const signed char* sa = a.cr.sign;
    const signed char* sb = a.cr.sign + (m + 1);
    const int* ta = a.cr.top;
    const int* tb = a.cr.top + (m + 1);
    const double* ma = a.cr.mant;
    const double* mb = a.cr.mant + (m + 1);

    compute_dinv(g, a.pc);                                       // needed by the tie-break and the update

    auto key_of = [&](int i) {
        RKey k;
        k.cls = (int)sb[i] * (int)sa[i];                         // sign of N_in / N_ic
        k.e = 0; k.f = 0.0;
        if (k.cls != 0) {
            int ex;
            const double fr = frexp((mb[i] * g.wmant[tb[i]]) / (ma[i] * g.wmant[ta[i]]), &ex);
            k.f = fr;
            k.e = ex + g.wexp[tb[i]] - g.wexp[ta[i]];
        }
        return k;
    };
    // pass 1: smallest approximate ratio over the eligible rows (a_ic > 0, simplex.py:273-275)
    RKey best{2, 0, 0.0};
    for (int i = tid; i < m; i += blockDim.x) {
        if ((int)sa[i] * sgnD <= 0) continue;
        const RKey k = key_of(i);
        if (rkey_less(k, best)) best = k;
    }
    s_cls[tid] = best.cls; s_e[tid] = best.e; s_f[tid] = best.f;
    if (tid == 0) { s_count = 0; s_int = INT_MAX; }
    __syncthreads();
    for (int off = blockDim.x >> 1; off > 0; off >>= 1) {
        if (tid < off) {
            const RKey x{s_cls[tid], s_e[tid], s_f[tid]}, y{s_cls[tid + off], s_e[tid + off], s_f[tid + off]};
            if (rkey_less(y, x)) { s_cls[tid] = y.cls; s_e[tid] = y.e; s_f[tid] = y.f; }
        }
        __syncthreads();
    }
    best = RKey{s_cls[0], s_e[0], s_f[0]};
    __syncthreads();
    if (best.cls == 2) {                                         // no positive entry: unbounded
        if (tid == 0) {
            if (st->dry_run) { st->last_r = -1; st->last_c = c; }
            st->status = LPS_UNBOUNDED;
        }
        return;
    }
    // pass 2: rows that can still be the exact minimum
    int r;
    if (best.cls == 0) {                                         // ratio exactly 0: lowest such row
        int low = INT_MAX;
        for (int i = tid; i < m; i += blockDim.x)
            if ((int)sa[i] * sgnD > 0 && sb[i] == 0) low = min(low, i);
        low = __reduce_min_sync(0xffffffffu, low);
        if ((tid & 31) == 0 && low != INT_MAX) atomicMin(&s_int, low);
        __syncthreads();
        r = s_int;
    } else {
        for (int i = tid; i < m; i += blockDim.x) {
            if ((int)sa[i] * sgnD <= 0) continue;
            const RKey k = key_of(i);
            if (rkey_close(k, best)) a.cand[atomicAdd(&s_count, 1)] = i;
        }
        __syncthreads();
        const int count = s_count;
        r = a.cand[0];
        for (int qi = 1; qi < count; ++qi) {                     // exact fold; order-independent result
            const int k = a.cand[qi];
            // G = (N_kn N_rc - N_rn N_kc) / D: an entry of the tableau one pivot away, so |G| fits
            uint32_t t[SPT][1];
#pragma unroll
            for (int s = 0; s < SPT; ++s) {
                const int kk = tid + s * blockDim.x;
                uint32_t v = 0u;
                if (kk < g.P) {
                    const PrimeQ p = g.pr[kk];
                    const uint32_t* pl = g.R + (size_t)kk * g.plane;
                    const uint32_t x = mmul(pl[(size_t)k * g.ld + g.n], pl[(size_t)r * g.ld + c], p.q, p.qinv);
                    const uint32_t y = mmul(pl[(size_t)r * g.ld + g.n], pl[(size_t)k * g.ld + c], p.q, p.qinv);
                    v = mmul(mmul(msub(x, y, p.q), a.pc.dinv_m[kk], p.q, p.qinv), p.r2, p.q, p.qinv);
                }
                t[s][0] = v;
            }
            __syncthreads();
            block_garner<1, SPT>(g, t, S);
            const int sg = garner_sign<1>(S, 0) * sgnD;          // sign of N_kn N_rc - N_rn N_kc
            if (sg < 0 || (sg == 0 && k < r)) r = k;             // ratio_k < ratio_r, or equal and lower row
            __syncthreads();
        }
    }
    if (st->dry_run) {
        if (tid == 0) { st->last_r = r; st->last_c = c; st->r = r; }
        return;
    }
    if (a.wcol) {                                                // room in the working prime set for this pivot?
        const int old = a.xbasis[r];
        const long long wn = st->wsum + a.wcol[c] - (old >= 0 ? a.wcol[old] : 0);
        if (wn > st->wlimit) {                                   // block-uniform: nothing has been modified yet
            __syncthreads();
            if (tid == 0) { st->r = r; st->wnext = wn; st->status = LPS_NEED_PRIMES; }
            return;
        }
        if (tid == 0) st->wnext = wn;
    }
    const int bad = compute_alpha(g, a.pc, r, c, &s_flag);
    if (tid == 0) {
        st->r = r;
        st->sgn_p = (int)sa[r];
        st->sgn_br = (int)sb[r] * sgnD;
        if (bad >= 0) { st->bad_prime = bad; st->status = LPS_BAD_PRIME; }
    }
}
template <int SPT>
__global__ void __launch_bounds__(kLaneThreads) k_xratio(XRatioArgs a) { xratio_body<SPT>(a); }

// small tableaux: this block reconstructs the (column, rhs) pair itself, runs the ratio test, and -- when a pivot
// was chosen -- stages the entering column (what k_xstage does): one launch instead of three
template <int SPTW>
__global__ void __launch_bounds__(kLaneThreads) k_xratio_fused(XRatioArgs a, GarnerJob job) {
    XState* st = a.st;
    const XGrid& g = a.g;
    if (st->status != LPS_OK) return;
    garner_job_in_block<SPTW>(g, st, job);
    xratio_body<1>(a);
    __syncthreads();
    if (st->status != LPS_OK || st->dry_run) return;             // block-uniform: written by thread 0 before the barrier
    const int c = st->c, r = st->r, rows = g.m + 1;
    for (int idx = threadIdx.x; idx < g.P * rows; idx += blockDim.x) {
        const int k = idx / rows, i = idx - k * rows;
        const PrimeQ p = g.pr[k];
        const uint32_t beta = mmul(mmul(g.R[(size_t)k * g.plane + (size_t)i * g.ld + c], p.r2, p.q, p.qinv), a.pc.dinv_m[k], p.q, p.qinv);
        a.nbeta[(size_t)k * a.nbeta_ld + i] = beta ? p.q - beta : 0u;
    }
    for (int k = threadIdx.x; k < g.P; k += blockDim.x) a.pc.Dres[k] = g.R[(size_t)k * g.plane + (size_t)r * g.ld + c];
}

// Explicit pivot (Tableau.pivot(r,c)): sign / zero test of N_rc, pivot constants.
template <int SPT>
__global__ void __launch_bounds__(kLaneThreads) k_xforced(XGrid g, XState* st, XPivotConst pc, int r, int c, const int* wcol,
                                                          const int* xbasis) {
    __shared__ GarnerShared<1> S;
    __shared__ int s_flag;
    const int tid = threadIdx.x;
    uint32_t t[SPT][1];
#pragma unroll
    for (int s = 0; s < SPT; ++s) {
        const int k = tid + s * blockDim.x;
        t[s][0] = k < g.P ? g.R[(size_t)k * g.plane + (size_t)r * g.ld + c] : 0u;
    }
    block_garner<1, SPT>(g, t, S);
    const int sg = garner_sign<1>(S, 0);
    if (sg == 0) {
        if (tid == 0) st->status = LPS_ZERO_PIVOT;
        return;
    }
    if (wcol) {
        const int old = xbasis[r];
        const long long wn = st->wsum + wcol[c] - (old >= 0 ? wcol[old] : 0);
        if (wn > st->wlimit) {
            __syncthreads();
            if (tid == 0) { st->r = r; st->c = c; st->wnext = wn; st->status = LPS_NEED_PRIMES; }
            return;
        }
        if (tid == 0) st->wnext = wn;
    }
    compute_dinv(g, pc);
    __syncthreads();
    const int bad = compute_alpha(g, pc, r, c, &s_flag);
    if (tid == 0) {
        st->r = r; st->c = c; st->sgn_p = sg; st->sgn_br = 0; st->sel_rule = 2;
        if (bad >= 0) { st->bad_prime = bad; st->status = LPS_BAD_PRIME; }
    }
}

// ------------------------------------------------------------------------------------------------
// whole-tableau rules in exact mode (Simplex.findPivotAll simplex.py:330-360, the per-column half of
// Simplex.findPivotMaxIncrease :286-328): the ratio test of MANY columns in one launch, one block per column, on
// the numbers a XSRC_COLPAIRS reconstruction left in `cr` (2(m+1) per column: the column, then the right-hand side).
// Same filter and exact tie-break as k_xratio; nothing of the solver state is touched.
// ------------------------------------------------------------------------------------------------
__global__ void k_xdinv(XGrid g, XPivotConst pc) { compute_dinv(g, pc); }

struct XColsArgs {
    XGrid g;
    const XState* st;
    XMeta cr;                 // ncols * 2(m+1) numbers
    const uint32_t* dinv_m;   // [P] Montgomery form of D^-1 (k_xdinv)
    int j0, ncols;            // columns j0 .. j0+ncols-1
    const signed char* obj_sign;   // objective row signs: only columns with c_j < 0 are tested; NULL: every column
    int* cand;                // ncols * m ints of scratch
    int* minrow;              // [ncols] lowest row of minimum ratio; -1: no positive entry; -2: column not tested
    int* nties;               // [ncols] rows of minimum ratio (NULL: not wanted)
    int* ties;                // [ncols * m] those rows, in no particular order
};

template <int SPT>
__global__ void __launch_bounds__(kLaneThreads) k_xratio_cols(XColsArgs a) {
    __shared__ GarnerShared<1> S;
    __shared__ int s_cls[kLaneThreads], s_e[kLaneThreads];
    __shared__ double s_f[kLaneThreads];
    __shared__ int s_count, s_int, s_nt;
    const XGrid& g = a.g;
    const int tid = threadIdx.x, jj = blockIdx.x, c = a.j0 + jj, m = g.m, sgnD = a.st->sgnD;
    if (a.obj_sign && (int)a.obj_sign[c] * sgnD >= 0) {         // c_j >= 0: not a candidate of findPivotMaxIncrease
        if (tid == 0) { a.minrow[jj] = -2; if (a.nties) a.nties[jj] = 0; }
        return;
    }
    const size_t off = (size_t)jj * 2 * (m + 1);
    const signed char* sa = a.cr.sign + off;
    const signed char* sb = sa + (m + 1);
    const int* ta = a.cr.top + off;
    const int* tb = ta + (m + 1);
    const double* ma = a.cr.mant + off;
    const double* mb = ma + (m + 1);
    int* cand = a.cand + (size_t)jj * m;
    int* ties = a.ties ? a.ties + (size_t)jj * m : nullptr;

    auto key_of = [&](int i) {
        RKey k;
        k.cls = (int)sb[i] * (int)sa[i];
        k.e = 0; k.f = 0.0;
        if (k.cls != 0) {
            int ex;
            const double fr = frexp((mb[i] * g.wmant[tb[i]]) / (ma[i] * g.wmant[ta[i]]), &ex);
            k.f = fr;
            k.e = ex + g.wexp[tb[i]] - g.wexp[ta[i]];
        }
        return k;
    };
    // sign of N_kn N_rc - N_rn N_kc, exactly (block-uniform result)
    auto cmp_rows = [&](int k, int r) {
        uint32_t t[SPT][1];
#pragma unroll
        for (int s = 0; s < SPT; ++s) {
            const int kk = tid + s * blockDim.x;
            uint32_t v = 0u;
            if (kk < g.P) {
                const PrimeQ p = g.pr[kk];
                const uint32_t* pl = g.R + (size_t)kk * g.plane;
                const uint32_t x = mmul(pl[(size_t)k * g.ld + g.n], pl[(size_t)r * g.ld + c], p.q, p.qinv);
                const uint32_t y = mmul(pl[(size_t)r * g.ld + g.n], pl[(size_t)k * g.ld + c], p.q, p.qinv);
                v = mmul(mmul(msub(x, y, p.q), a.dinv_m[kk], p.q, p.qinv), p.r2, p.q, p.qinv);
            }
            t[s][0] = v;
        }
        __syncthreads();
        block_garner<1, SPT>(g, t, S);
        const int sg = garner_sign<1>(S, 0) * sgnD;
        __syncthreads();
        return sg;
    };

    RKey best{2, 0, 0.0};
    for (int i = tid; i < m; i += blockDim.x) {
        if ((int)sa[i] * sgnD <= 0) continue;
        const RKey k = key_of(i);
        if (rkey_less(k, best)) best = k;
    }
    s_cls[tid] = best.cls; s_e[tid] = best.e; s_f[tid] = best.f;
    if (tid == 0) { s_count = 0; s_int = INT_MAX; s_nt = 0; }
    __syncthreads();
    for (int o = blockDim.x >> 1; o > 0; o >>= 1) {
        if (tid < o) {
            const RKey x{s_cls[tid], s_e[tid], s_f[tid]}, y{s_cls[tid + o], s_e[tid + o], s_f[tid + o]};
            if (rkey_less(y, x)) { s_cls[tid] = y.cls; s_e[tid] = y.e; s_f[tid] = y.f; }
        }
        __syncthreads();
    }
    best = RKey{s_cls[0], s_e[0], s_f[0]};
    __syncthreads();
    if (best.cls == 2) {                                         // no positive entry in this column
        if (tid == 0) { a.minrow[jj] = -1; if (a.nties) a.nties[jj] = 0; }
        return;
    }
    int r;
    if (best.cls == 0) {                                         // ratio exactly 0: every such row ties
        int low = INT_MAX;
        for (int i = tid; i < m; i += blockDim.x)
            if ((int)sa[i] * sgnD > 0 && sb[i] == 0) {
                low = min(low, i);
                if (ties) ties[atomicAdd(&s_nt, 1)] = i;
            }
        low = __reduce_min_sync(0xffffffffu, low);
        if ((tid & 31) == 0 && low != INT_MAX) atomicMin(&s_int, low);
        __syncthreads();
        r = s_int;
    } else {
        for (int i = tid; i < m; i += blockDim.x) {
            if ((int)sa[i] * sgnD <= 0) continue;
            const RKey k = key_of(i);
            if (rkey_close(k, best)) cand[atomicAdd(&s_count, 1)] = i;
        }
        __syncthreads();
        const int count = s_count;
        r = cand[0];
        for (int qi = 1; qi < count; ++qi) {
            const int k = cand[qi];
            const int sg = cmp_rows(k, r);
            if (sg < 0 || (sg == 0 && k < r)) r = k;
        }
        if (ties) {                                              // rows whose ratio equals the minimum exactly
            for (int qi = 0; qi < count; ++qi) {
                const int k = cand[qi];
                const int sg = (k == r) ? 0 : cmp_rows(k, r);
                if (sg == 0 && tid == 0) ties[s_nt++] = k;
            }
            __syncthreads();
        }
    }
    if (tid == 0) { a.minrow[jj] = r; if (a.nties) a.nties[jj] = s_nt; }
}

// Tableau.isInfeasible (tableau.py:510-514) on the signs of a batch of reconstructed rows (XSRC_ROWS): some row with
// b_i > 0 has no positive entry.  One warp per row.
__global__ void k_xrow_infeasible(const signed char* sign, int nrows, int n, const XState* st, int* flag) {
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5, sgnD = st->sgnD;
    for (int i = blockIdx.x * wpb + (threadIdx.x >> 5); i < nrows; i += gridDim.x * wpb) {
        const signed char* row = sign + (size_t)i * (n + 1);
        if ((int)row[n] * sgnD <= 0) continue;
        bool pos = false;
        for (int j = lane; j < n && !pos; j += 32) pos = (int)row[j] * sgnD > 0;
        if (!__any_sync(0xffffffffu, pos) && lane == 0) atomicOr(flag, 1);
    }
}

// residues of scattered cells (rows[e], cols[e]) as a XSRC_CUSTOM source: out[k*ostride + e]
__global__ void __launch_bounds__(256) k_xgather_cells(XGrid g, const int* rows, const int* cols, int count, uint32_t* out, size_t ostride) {
    const int k = blockIdx.y;
    const uint32_t* pl = g.R + (size_t)k * g.plane;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < count; e += gridDim.x * blockDim.x)
        out[(size_t)k * ostride + e] = pl[(size_t)rows[e] * g.ld + cols[e]];
}

// ------------------------------------------------------------------------------------------------
// stage: snapshot of the entering column as negated multipliers, new denominator
// ------------------------------------------------------------------------------------------------
// nbeta[k][i] = Montgomery form of -(N_ic D^-1) mod p_k  (the column is overwritten by the update, and
// the update's CTAs need its old values).  Also D <- p.
__global__ void __launch_bounds__(256) k_xstage(XGrid g, const XState* st, XPivotConst pc, uint32_t* nbeta, size_t nbeta_ld) {
    if (st->status != LPS_OK) return;
    const int k = blockIdx.y, c = st->c, r = st->r;
    const PrimeQ p = g.pr[k];
    const uint32_t dinv = pc.dinv_m[k];
    const uint32_t* pl = g.R + (size_t)k * g.plane;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i <= g.m; i += gridDim.x * blockDim.x) {
        const uint32_t beta = mmul(mmul(pl[(size_t)i * g.ld + c], p.r2, p.q, p.qinv), dinv, p.q, p.qinv);
        nbeta[(size_t)k * nbeta_ld + i] = beta ? p.q - beta : 0u;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) pc.Dres[k] = pl[(size_t)r * g.ld + c];
}

// ------------------------------------------------------------------------------------------------
// update: N_ij <- alpha N_ij - beta_i N_rj (mod p_k), one prime per grid slice
// ------------------------------------------------------------------------------------------------
struct XUpdateArgs {
    XGrid g;
    XState* st;
    const uint32_t* alpha_m;
    const uint32_t* nbeta;
    size_t nbeta_ld;
    int rows_per_cta;
    int* basis;
    int* marks;
    int* trace;
    long long trace_cap;
    int* alive;              // per row: sigma_i still applies (row not pivoted since the upload)
    int* xbasis;             // [m] last pivot column per row since the upload (working prime set), or NULL
    int commit;
};

template <int CW, bool STREAM>
__global__ void __launch_bounds__(kXThreads) k_xupdate(XUpdateArgs a) {
    constexpr int VPT = CW / (4 * kXThreads);      // uint4 vectors per thread per row
    constexpr int RU = 4;
    static_assert(VPT >= 1, "chunk too narrow");
    __shared__ alignas(128) uint32_t s_v[CW];
    __shared__ alignas(128) uint32_t s_nb[kXMaxRowBlock];
    __shared__ alignas(8) uint64_t s_bar;
    XState* st = a.st;
    const XGrid& g = a.g;
    if (st->status != LPS_OK) return;
    const int k = blockIdx.z, r = st->r;
    const PrimeQ p = g.pr[k];
    const uint32_t alpha = a.alpha_m[k];
    uint32_t* plane = g.R + (size_t)k * g.plane;
    const int col0 = blockIdx.x * CW, row0 = blockIdx.y * a.rows_per_cta;
    const int ncols = min(CW, (int)g.ld - col0);
    const int nrows = min(a.rows_per_cta, g.m + 1 - row0);

    if (threadIdx.x == 0) {
        mbar_init(&s_bar, 1);
        mbar_fence_init();
        const uint32_t vb = (uint32_t)ncols * 4u, ub = (uint32_t)a.rows_per_cta * 4u;
        mbar_expect_tx(&s_bar, vb + ub);
        tma_load_1d(s_v, plane + (size_t)r * g.ld + col0, vb, &s_bar);
        tma_load_1d(s_nb, a.nbeta + (size_t)k * a.nbeta_ld + row0, ub, &s_bar);
    }
    __syncthreads();
    mbar_wait(&s_bar, 0);

    uint4 v[VPT];
    bool vok[VPT];
#pragma unroll
    for (int q = 0; q < VPT; ++q) {
        const int jl = 4 * threadIdx.x + q * 4 * kXThreads;
        vok[q] = jl < ncols;
        v[q] = vok[q] ? *reinterpret_cast<const uint4*>(&s_v[jl]) : make_uint4(0, 0, 0, 0);
    }
    uint32_t* base = plane + (size_t)row0 * g.ld + col0 + 4 * threadIdx.x;
    for (int rg = 0; rg < nrows; rg += RU) {
        uint4 t[RU][VPT];
        uint32_t nb[RU];
        bool act[RU];
#pragma unroll
        for (int u = 0; u < RU; ++u) {
            const int il = rg + u;
            act[u] = il < nrows && row0 + il != r;
            nb[u] = act[u] ? s_nb[il] : 0u;
            if (act[u]) {
#pragma unroll
                for (int q = 0; q < VPT; ++q)
                    if (vok[q]) t[u][q] = ld_u32x4<STREAM>(base + (size_t)il * g.ld + q * 4 * kXThreads);
            }
        }
#pragma unroll
        for (int u = 0; u < RU; ++u) {
            if (!act[u]) continue;
            const int il = rg + u;
#pragma unroll
            for (int q = 0; q < VPT; ++q) {
                if (!vok[q]) continue;
                uint4 x = t[u][q];
                x.x = redc((uint64_t)alpha * x.x + (uint64_t)nb[u] * v[q].x, p.q, p.qinv);
                x.y = redc((uint64_t)alpha * x.y + (uint64_t)nb[u] * v[q].y, p.q, p.qinv);
                x.z = redc((uint64_t)alpha * x.z + (uint64_t)nb[u] * v[q].z, p.q, p.qinv);
                x.w = redc((uint64_t)alpha * x.w + (uint64_t)nb[u] * v[q].w, p.q, p.qinv);
                st_u32x4<STREAM>(base + (size_t)il * g.ld + q * 4 * kXThreads, x);
            }
        }
    }

    if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 0) {
        const int c = st->c;
        st->last_r = r; st->last_c = c;
        st->sgnD = st->sgn_p;
        a.alive[r] = 0;
        if (a.xbasis) { a.xbasis[r] = c; st->wsum = st->wnext; }
        if (a.commit) {                                          // Simplex._pivot simplex.py:192-197
            const long long kk = st->npiv;
            if (a.trace && kk < a.trace_cap) { a.trace[3 * kk] = r; a.trace[3 * kk + 1] = c; a.trace[3 * kk + 2] = st->sel_rule; }
            if (a.marks) {
                const int old = a.basis[r];
                if (old >= 0 && old < g.n) a.marks[old] = 0;
                a.marks[c] = 1;
            }
            a.basis[r] = c;
            st->npiv = kk + 1;
            if (st->sel_rule == LPS_RULE_DANTZIG) {
                st->pending_zcheck = 1;
                if (st->sgn_br > 0) st->moved = 1;
                else if (st->sgn_br < 0 && !st->moved) st->obj_increased = 1;   // reported by the next k_xprice (simplex.py:133)
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// encode: signed multi-limb integers -> residues
// ------------------------------------------------------------------------------------------------
// limbs: entry e has nlimbs little-endian 32-bit limbs at limbs[e*nlimbs ..]; neg[e] != 0: negative.
// Entry e = i*ld_host + j of the host grid goes to R[k][i][j]; padding columns are zeroed.
__global__ void __launch_bounds__(256) k_xencode(XGrid g, const uint32_t* limbs, int nlimbs, const unsigned char* neg, size_t ld_host) {
    const int k = blockIdx.y;
    const PrimeQ p = g.pr[k];
    uint32_t* plane = g.R + (size_t)k * g.plane;
    const size_t total = (size_t)(g.m + 1) * g.ld;
    for (size_t x = (size_t)blockIdx.x * blockDim.x + threadIdx.x; x < total; x += (size_t)gridDim.x * blockDim.x) {
        const size_t i = x / g.ld, j = x - i * g.ld;
        uint32_t res = 0u;
        if (j <= (size_t)g.n) {
            const size_t e = i * ld_host + j;
            const uint32_t* L = limbs + e * nlimbs;
            for (int l = nlimbs - 1; l >= 0; --l) {
                res = mmul(res, p.r2, p.q, p.qinv);              // res * 2^32
                uint32_t w = L[l];
                if (w >= p.q) w %= p.q;
                res += w;
                if (res >= p.q) res -= p.q;
            }
            if (neg[e] && res) res = p.q - res;
        }
        plane[x] = res;
    }
}
// same for a vector (the row scales; signed when `neg` is given): out[k*ostride + e]
__global__ void __launch_bounds__(256) k_xencode_vec(XGrid g, const uint32_t* limbs, int nlimbs, int count, uint32_t* out, size_t ostride,
                                                     const unsigned char* neg = nullptr) {
    const int k = blockIdx.y;
    const PrimeQ p = g.pr[k];
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < count; e += gridDim.x * blockDim.x) {
        uint32_t res = 0u;
        const uint32_t* L = limbs + (size_t)e * nlimbs;
        for (int l = nlimbs - 1; l >= 0; --l) {
            res = mmul(res, p.r2, p.q, p.qinv);
            uint32_t w = L[l];
            if (w >= p.q) w %= p.q;
            res += w;
            if (res >= p.q) res -= p.q;
        }
        if (neg && neg[e] && res) res = p.q - res;
        out[(size_t)k * ostride + e] = res;
    }
}

// row operations outside a pivot (tableau.py:254-293): N_rd <- a N_rd + b N_rs (mod p_k), rs < 0: N_rd <- a N_rd
__global__ void __launch_bounds__(256) k_xrow_combine(XGrid g, int rd, int rs, const uint32_t* a_res, const uint32_t* b_res) {
    const int k = blockIdx.y;
    const PrimeQ p = g.pr[k];
    uint32_t* dst = g.R + (size_t)k * g.plane + (size_t)rd * g.ld;
    const uint32_t* src = rs >= 0 ? g.R + (size_t)k * g.plane + (size_t)rs * g.ld : nullptr;
    const uint32_t am = mmul(a_res[k], p.r2, p.q, p.qinv);        // Montgomery forms: mmul(am, x) = a x
    const uint32_t bm = src ? mmul(b_res[k], p.r2, p.q, p.qinv) : 0u;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j <= g.n; j += gridDim.x * blockDim.x) {
        uint32_t v = mmul(am, dst[j], p.q, p.qinv);
        if (src) {
            v += mmul(bm, src[j], p.q, p.qinv);
            if (v >= p.q) v -= p.q;
        }
        dst[j] = v;
    }
}

// ------------------------------------------------------------------------------------------------
// leaving phase 1 on the device (Simplex._find_bfs simplex.py:95-105, every row kept)
// ------------------------------------------------------------------------------------------------
// rhs[k][i] <- N_i,n (the right-hand side column of every plane, before the columns are cut)
__global__ void __launch_bounds__(256) k_xgather_rhs(XGrid g, uint32_t* rhs, size_t rstride) {
    const int k = blockIdx.y;
    const uint32_t* pl = g.R + (size_t)k * g.plane;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i <= g.m; i += gridDim.x * blockDim.x) rhs[(size_t)k * rstride + i] = pl[(size_t)i * g.ld + g.n];
}
// cut the trailing columns n_new..g.n-1 (the artificial variables): the right-hand side lands in column n_new, the
// columns behind it become padding (zero)
__global__ void __launch_bounds__(256) k_xcut_cols(XGrid g, int n_new, const uint32_t* rhs, size_t rstride) {
    const int k = blockIdx.y, width = g.n - n_new + 1;
    uint32_t* pl = g.R + (size_t)k * g.plane;
    const long long total = (long long)(g.m + 1) * width;
    for (long long x = (long long)blockIdx.x * blockDim.x + threadIdx.x; x < total; x += (long long)gridDim.x * blockDim.x) {
        const long long i = x / width;
        const int jj = (int)(x - i * width);
        pl[(size_t)i * g.ld + n_new + jj] = jj == 0 ? rhs[(size_t)k * rstride + i] : 0u;
    }
}
// The original objective priced out against the current basis (setZ / setC, then rowSubFromObj(i, c_{B_i}) for every
// row: simplex.py:101-105) in the fraction-free form: with C_j = sigma_m c_j (integers; C_n the _z cell) and
// G_i = C_{B_i} / s_i (s_i: the scale row i still carries, or 1)
//     N_mj <- D C_j - sum_i G_i N_ij      so that  N_mj / (D sigma_m) = c_j - sum_i c_{B_i} T_ij.
// One thread per column and plane walks the rows (coalesced across the warp); rows with G_i = 0 are skipped.
__global__ void __launch_bounds__(128) k_xrestore_obj(XGrid g, const uint32_t* Dres, const uint32_t* C, size_t cstride,
                                                      const uint32_t* G, size_t gstride) {
    const int k = blockIdx.y, j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j > g.n) return;
    const PrimeQ p = g.pr[k];
    uint32_t* pl = g.R + (size_t)k * g.plane;
    const uint32_t* Gk = G + (size_t)k * gstride;
    uint64_t sum = 0;                                            // plain residues: m < 2^31 terms below 2^31 each
    for (int i = 0; i < g.m; ++i) {
        const uint32_t gi = Gk[i];
        if (gi == 0u) continue;
        sum += mmul(mmul(gi, p.r2, p.q, p.qinv), pl[(size_t)i * g.ld + j], p.q, p.qinv);
    }
    const uint32_t dc = mmul(mmul(Dres[k], p.r2, p.q, p.qinv), C[(size_t)k * cstride + j], p.q, p.qinv);
    pl[(size_t)g.m * g.ld + j] = msub(dc, (uint32_t)(sum % p.q), p.q);
}

// invtab[j*P+k] = Montgomery form of p_j^-1 mod p_k for j < k (0 elsewhere)
__global__ void k_xinvtab(const PrimeQ* pr, int P, uint32_t* invtab) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
    if (k >= P) return;
    uint32_t out = 0u;
    if (j < k) {
        const PrimeQ p = pr[k];
        const uint32_t pj = pr[j].q % p.q;
        out = minv_mont(mmul(pj, p.r2, p.q, p.qinv), p);
    }
    invtab[(size_t)j * P + k] = out;
}

// ------------------------------------------------------------------------------------------------
// Tableau.isCanonical (tableau.py:466-496) on residues
// ------------------------------------------------------------------------------------------------
// one[k][i] = residue of the integer that represents the value 1 in row i: D * sigma_i (or D)
__global__ void k_xone(XGrid g, const uint32_t* Dres, const uint32_t* sigma, size_t sstride, const int* alive, uint32_t* one) {
    const int k = blockIdx.y;
    const PrimeQ p = g.pr[k];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i <= g.m; i += gridDim.x * blockDim.x) {
        uint32_t v = Dres[k];
        if (alive[i]) v = mmul(mmul(v, p.r2, p.q, p.qinv), sigma[(size_t)k * sstride + i], p.q, p.qinv);
        one[(size_t)k * sstride + i] = v;
    }
}
// pass A on plane 0: columns that look basic modulo p_0 (a column rejected here is rejected for good)
__global__ void k_xcanon_a(XGrid g, const uint32_t* one, int* colrow) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= g.n) return;
    const uint32_t* pl = g.R;
    int hit = -1;
    bool clean = (pl[(size_t)g.m * g.ld + j] == 0u);
    for (int i = 0; clean && i < g.m; ++i) {
        const uint32_t v = pl[(size_t)i * g.ld + j];
        if (v == 0u) continue;
        if (v == one[i] && hit == -1) hit = i; else clean = false;
    }
    colrow[j] = (clean && hit != -1) ? hit : -1;
}
// pass B: confirm the surviving columns modulo every other prime
__global__ void k_xcanon_b(XGrid g, const uint32_t* one, size_t sstride, int* colrow) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, k = blockIdx.y + 1;
    if (j >= g.n || k >= g.P) return;
    const int hit = colrow[j];
    if (hit < 0) return;
    const uint32_t* pl = g.R + (size_t)k * g.plane;
    bool ok = (pl[(size_t)g.m * g.ld + j] == 0u);
    for (int i = 0; ok && i < g.m; ++i) {
        const uint32_t v = pl[(size_t)i * g.ld + j];
        ok = (i == hit) ? (v == one[(size_t)k * sstride + i]) : (v == 0u);
    }
    if (!ok) colrow[j] = -1;
}
// any b_i < 0 ?  from the reconstructed right-hand side
__global__ void k_xany_neg(const signed char* sign, int m, const XState* st, int* flag) {
    int neg = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) neg |= ((int)sign[i] * st->sgnD < 0);
    if (__any_sync(0xffffffffu, neg) && (threadIdx.x & 31) == 0) atomicOr(flag, 1);
}

}  // namespace lps
