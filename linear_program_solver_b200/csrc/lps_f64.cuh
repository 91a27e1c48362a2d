// fp64 simplex pivot kernels for sm_100a.
//
// One pivot = k_select_f64 (pricing on the objective row + ratio test on the local rows, produces
// this rank's candidate: key + normalised pivot row) followed -- after the caller's all-gather of
// candidates when the tableau is row-sharded -- by k_update_f64 (picks the global leaving row,
// applies T <- T - u (x) v to the local rows, prices the new objective row for the next pivot and
// commits the pivot).  No host round-trip inside the loop.
//
// Reference semantics reproduced (file:line under /root/reference/lpsol):
//   tableau.py:295-308  pivot = row r times 1/p, objective += (-c_c) row r, row i += (-a_ic) row r
//                       (product then sum, two roundings -- never an FMA; rows with a zero
//                       multiplier are skipped, tableau.py:272/:285)
//   simplex.py:262-267  Dantzig: most negative cost, lowest index among equals
//   simplex.py:228-232  Bland: first negative cost
//   simplex.py:272-279  ratio test: a_ic > 0 only, b_i / a_ic, lowest row among equals
//   simplex.py:118-137  stuck counter against the objective at entry, then the rule switch
#pragma once
#include "lps_common.cuh"
#include "../../include/lps_b200.h"

namespace lps {

constexpr int kThreads = 256;
constexpr int kSelMaxThreads = 1024;  // k_select_f64 runs as one CTA of 1024 threads on small tableaux
constexpr int kMaxRowBlock = 128;     // rows of u staged per CTA
constexpr int kMaxRanks = 64;

// Candidate slot exchanged between ranks: 64-byte header followed by ld doubles (the normalised row).
struct alignas(16) CandHeader {
    double ratio;        // b_i / a_ic of this rank's best row
    long long row;       // GLOBAL row index of that row, -1: no row with a_ic > 0
    int c;               // entering column (identical on every rank: the objective row is replicated)
    int rule;            // rule that priced it
    long long stuck;     // stuck counter after the objective check of this step
    int pad[8];
};
static_assert(sizeof(CandHeader) == 64, "candidate header must stay 64 bytes");

// Exchange block: one per rank, in memory every peer GPU can write (cudaMalloc + CUDA IPC over NVLink).  Per pivot
//   1. every rank writes the 64-byte header of its best leaving row into slot `rank` of EVERY rank's block and
//      then releases hdr_seq[rank] = pivot sequence number;
//   2. the rank that owns the winning row normalises it and writes it into EVERY rank's vrow, then releases
//      row_seq = sequence number;
// so the argmin "all-reduce" and the pivot-row "broadcast" are plain peer stores issued by the pivot kernels
// themselves -- no collective call, no host step.  Headers are double buffered by sequence parity (a rank can be at
// most one pivot ahead of a peer), vrow is single buffered (it is rewritten only after every rank has published
// the next header, i.e. finished the previous update).
constexpr int kMaxPanelCtas = 256;
struct XchgBlock {
    CandHeader hdr[2][kMaxRanks];
    unsigned long long hdr_seq[kMaxRanks];
    unsigned long long row_seq;
    unsigned long long pad[15];
    // blocked loop (lps_f64_blk.cuh): ONE exchange per pivot.  Every rank stores a 32-byte self-validating candidate
    // record (ratio bits, row | sequence number << 32, check word) into slot `rank` of every rank's hdr2 -- no fence,
    // no flag: a reader takes a record only when its sequence number and check word agree -- and its own normalised
    // candidate row into every rank's row area, slice by slice, each slice followed by a record of the same kind in
    // rowrec[rank][cta]: a position-weighted checksum of the slice, the sequence number, the row's right-hand side.
    // No fence there either: the receiver reads the slice (it needs it anyway), recomputes the checksum and reads
    // again until it agrees -- a record that overtakes its data is noticed, not trusted.
    unsigned long long hdr2[2][kMaxRanks][4];
    unsigned long long rowrec[2][kMaxRanks][kMaxPanelCtas][4];   // by sequence parity, like hdr2: a rank can be one pivot ahead
    // followed by the row area (128-byte aligned): the rank-1 loop uses its first ld doubles (vrow); the blocked loop
    // [set 2][rank nranks][pivot kBlkMax][ld]
};
static_assert(sizeof(XchgBlock) % 128 == 0, "vrow must stay 128-byte aligned");
__host__ __device__ __forceinline__ double* xchg_vrow(XchgBlock* x) { return reinterpret_cast<double*>(x + 1); }
__host__ __device__ __forceinline__ const double* xchg_vrow(const XchgBlock* x) { return reinterpret_cast<const double*>(x + 1); }

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// fire-and-forget arrival: the release orders this thread's (and, through the preceding block barrier, its block's)
// earlier stores before the increment becomes visible
__device__ __forceinline__ void red_release_gpu(unsigned long long* p) {
    asm volatile("red.release.gpu.global.add.u64 [%0], 1;" ::"l"(p) : "memory");
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ unsigned long long ld_acquire_gpu(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ld_relaxed_gpu(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// spin until *p >= want; false after ten seconds (a peer died or never enqueued this pivot: do not hang the GPU).
// The polls are relaxed loads (an acquire load invalidates L1 on every poll); one acquire fence follows the hit.
__device__ __forceinline__ bool wait_seq(const unsigned long long* p, unsigned long long want) {
    if (ld_acquire_sys(p) >= want) return true;
    const unsigned long long t0 = global_ns();
    unsigned int polls = 0;
    while (ld_relaxed_sys(p) < want) {
        if ((++polls & 1023u) == 0 && global_ns() - t0 > 10000000000ull) return false;
    }
    asm volatile("fence.acq_rel.sys;" ::: "memory");
    return true;
}
// same on a flag raised by another CTA of this kernel (twenty seconds: outlasts wait_seq)
__device__ __forceinline__ bool wait_seq_gpu(const unsigned long long* p, unsigned long long want) {
    if (ld_acquire_gpu(p) >= want) return true;
    const unsigned long long t0 = global_ns();
    unsigned int polls = 0;
    while (ld_relaxed_gpu(p) < want) {
        if ((++polls & 1023u) == 0 && global_ns() - t0 > 20000000000ull) return false;
    }
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
    return true;
}

struct SolveState {
    int status;              // LPS_OK while running
    int rule;                // committed rule (Dantzig until the stuck counter trips)
    long long stuck;
    double z_entry_cell;     // corner cell when solve() was entered (simplex.py:118)
    long long npiv;
    long long max_pivots;    // < 0: unlimited
    int pending_zcheck;      // last committed pivot was Dantzig-priced: compare the objective next
    unsigned int ticket;     // last-block election of k_select
    int dry_run;             // find_pivot: select only, nothing is committed
    int last_r, last_c;      // last selected / committed pivot (global row)
    unsigned int ticket2;    // last-block election of the pivot-row broadcast
    unsigned long long xseq; // select rounds decided so far; raised inside k_select as its "decision published" flag
    int sel_w;               // decision of the current round: winning rank, -1 nothing to normalise, -2 peer time-out
    int pad;
    long long sel_row;       // global row of the winner
};

struct F64Grid {
    double* T;               // (m+1) x ld, row m = objective row
    size_t ld;
    int m, n;                // local constraint rows, variables
    long long row0, m_total; // global placement of the local rows
};

struct PricePartials {       // one entry per column chunk of the objective row
    double* d_val;           // Dantzig: most negative value in the chunk (0 if none)
    int* d_idx;              //          its lowest index, INT_MAX if none
    int* b_idx;              // Bland:   lowest index with a negative value, INT_MAX if none
};

// ------------------------------------------------------------------------------------------------
// pricing of one chunk of the objective row (used standalone and fused in the update)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void price_consider(double v, int j, int n, ArgMinF64& best, int& first_neg) {
    if (j < n && v < 0.0) {
        if (best.idx == INT_MAX || v < best.key || (v == best.key && j < best.idx)) {
            best.key = v;
            best.idx = j;
        }
        first_neg = min(first_neg, j);
    }
}

template <int CW>
__global__ void __launch_bounds__(kThreads) k_price_f64(F64Grid g, PricePartials pp) {
    __shared__ ArgMinF64 s_arg[32];
    __shared__ int s_int[32];
    const double* obj = g.T + (size_t)g.m * g.ld;
    const int col0 = blockIdx.x * CW;
    ArgMinF64 best{0.0, INT_MAX};
    int first_neg = INT_MAX;
    for (int j = col0 + threadIdx.x; j < min(col0 + CW, g.n); j += kThreads) price_consider(obj[j], j, g.n, best, first_neg);
    best = block_argmin(best, s_arg);
    first_neg = block_min_int(first_neg, s_int);
    if (threadIdx.x == 0) {
        pp.d_val[blockIdx.x] = best.key;
        pp.d_idx[blockIdx.x] = best.idx;
        pp.b_idx[blockIdx.x] = first_neg;
    }
}

// ------------------------------------------------------------------------------------------------
// select: entering column from the partials, ratio test over the local rows, candidate slot
// ------------------------------------------------------------------------------------------------
struct SelectArgs {
    F64Grid g;
    PricePartials pp;
    int nchunks;
    SolveState* st;
    double* ucol;            // m+1 (+pad): column c of the local grid, gathered here for the update
    ArgMinF64* blk;          // gridDim.x per-block ratio partials
    CandHeader* cand;        // this rank's candidate slot (header + ld doubles): one rank, or the all-gather exchange
    XchgBlock* xchg;         // peer-store exchange: this rank's exchange block (NULL: candidate slots)
    XchgBlock* const* peers; //                      every rank's exchange block
    int rank, nranks;
};

// header of this rank's candidate into slot `rank` of every rank's exchange block, then the sequence flag
__device__ __forceinline__ void publish_header(XchgBlock* const* peers, int rank, int nranks, const CandHeader& h,
                                               unsigned long long seq) {
    if ((int)threadIdx.x < nranks) {
        XchgBlock* pb = peers[threadIdx.x];
        const int4* src = reinterpret_cast<const int4*>(&h);
        int4* dst = reinterpret_cast<int4*>(&pb->hdr[seq & 1][rank]);
#pragma unroll
        for (int q = 0; q < 4; ++q) dst[q] = src[q];
        __threadfence_system();
        st_release_sys(&pb->hdr_seq[rank], seq);
    }
}

__device__ __forceinline__ int pick_winner(const XchgBlock* x, int par, int nranks) {
    int w = -1;
    double w_ratio = 0.0;
    long long w_row = -1;
    for (int k = 0; k < nranks; ++k) {
        const CandHeader* h = &x->hdr[par][k];
        const long long row = h->row;
        if (row < 0) continue;
        const double ratio = h->ratio;
        if (w < 0 || ratio < w_ratio || (ratio == w_ratio && row < w_row)) { w = k; w_ratio = ratio; w_row = row; }
    }
    return w;
}

// One launch per pivot decision, three phases separated by an in-kernel flag (the grid is at most one CTA per SM,
// so every CTA is resident and spinning on the flag cannot starve the CTA that raises it):
//   A  every CTA: rule / entering column from the pricing partials, ratio test over its rows, column c -> ucol
//   B  the CTA that finishes last: reduces the partials to this rank's best leaving row; with the peer-store
//      exchange it writes the 64-byte header into every peer GPU and waits for theirs (the argmin all-reduce),
//      then publishes the decision (winning rank, row) and raises the flag
//   C  every CTA of the rank that owns the winning row: normalises its slice of that row (row r times the
//      reciprocal of the pivot, tableau.py:262-267) and stores it where the update reads it -- the local candidate
//      slot, or straight into EVERY rank's exchange block over NVLink (the pivot-row broadcast); the other ranks
//      wait for the row's sequence flag.  When the kernel ends on a rank, that rank holds all its update needs.
__global__ void __launch_bounds__(kSelMaxThreads, 1) k_select_f64(SelectArgs a) {
    __shared__ ArgMinF64 s_arg[32];
    __shared__ int s_int[32];
    __shared__ int s_last, s_w;
    __shared__ long long s_row;
    SolveState* st = a.st;
    const F64Grid& g = a.g;
    const int nt = (int)blockDim.x;
    if (st->status != LPS_OK) return;
    if (st->max_pivots >= 0 && st->npiv >= st->max_pivots) {
        if (blockIdx.x == 0 && threadIdx.x == 0 && !st->dry_run) st->status = LPS_MAX_PIVOTS;
        return;
    }
    // this round's sequence number: read before the ticket below, i.e. before the last CTA can advance it
    const unsigned long long seq = *reinterpret_cast<volatile unsigned long long*>(&st->xseq) + 1;
    const int dry = st->dry_run;
    // stuck counter and rule (simplex.py:123,132-137), recomputed identically by every block
    int rule = st->rule;
    long long stuck = st->stuck;
    if (st->pending_zcheck) {
        const double corner = g.T[(size_t)g.m * g.ld + g.n];
        // simplex.py:133 `assert z <= obj_val`: the objective (minus the corner cell) may not rise above its value at
        // entry; a float run can trip this on rounding noise, and the reference then stops with the pivot applied
        if (!(corner >= st->z_entry_cell)) {
            if (blockIdx.x == 0 && threadIdx.x == 0 && !dry) st->status = LPS_OBJ_INCREASED;
            if (!dry) return;
        }
        stuck = (corner == st->z_entry_cell) ? stuck + 1 : 0;
    }
    if (rule == LPS_RULE_DANTZIG && !(stuck < g.m_total + (long long)g.n)) rule = LPS_RULE_BLAND;

    // entering column
    int c;
    if (rule == LPS_RULE_DANTZIG) {
        ArgMinF64 best{0.0, INT_MAX};
        for (int k = threadIdx.x; k < a.nchunks; k += nt) best = argmin_pick(best, ArgMinF64{a.pp.d_val[k], a.pp.d_idx[k]});
        best = block_argmin(best, s_arg);
        c = best.idx;
    } else {
        int first = INT_MAX;
        for (int k = threadIdx.x; k < a.nchunks; k += nt) first = min(first, a.pp.b_idx[k]);
        c = block_min_int(first, s_int);
    }
    if (c == INT_MAX) {
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            if (dry) { st->last_r = -1; st->last_c = -1; }
            st->status = LPS_OPTIMAL;
        }
        return;
    }

    // ---- A: ratio test over this block's rows; column c is gathered into ucol on the way ----
    ArgMinF64 best{0.0, INT_MAX};
    for (int i = blockIdx.x * nt + threadIdx.x; i < g.m; i += gridDim.x * nt) {
        const double aic = g.T[(size_t)i * g.ld + c];
        a.ucol[i] = aic;
        if (aic > 0.0) {
            const double ratio = g.T[(size_t)i * g.ld + g.n] / aic;
            if (best.idx == INT_MAX || ratio < best.key) { best.key = ratio; best.idx = i; }
        }
    }
    best = block_argmin(best, s_arg);
    const bool solo = (gridDim.x == 1);                        // small tableaux: one CTA, no cross-CTA handshake at all
    if (threadIdx.x == 0) {
        if (solo) s_last = 1;
        else {
            a.blk[blockIdx.x] = best;
            __threadfence();
            const unsigned int t = atomicAdd(&st->ticket, 1u);
            s_last = (t == gridDim.x - 1);
        }
    }
    __syncthreads();

    if (s_last) {
        // ---- B: reduce the per-block partials, exchange, publish the decision ----
        if (!solo) {
            __threadfence();
            best = ArgMinF64{0.0, INT_MAX};
            for (int k = threadIdx.x; k < (int)gridDim.x; k += nt) {
                ArgMinF64 o;
                o.key = __ldcg(&a.blk[k].key);
                o.idx = __ldcg(&a.blk[k].idx);
                best = argmin_pick(best, o);
            }
            best = block_argmin(best, s_arg);
        }
        const int r = best.idx;
        CandHeader h{};
        h.ratio = best.key; h.row = (r == INT_MAX) ? -1 : g.row0 + r; h.c = c; h.rule = rule; h.stuck = stuck;
        int w;
        long long w_row;
        if (a.xchg && !dry) {
            publish_header(a.peers, a.rank, a.nranks, h, seq);
            int ok = 1;
            if ((int)threadIdx.x < a.nranks) ok = wait_seq(&a.xchg->hdr_seq[threadIdx.x], seq) ? 1 : 0;
            ok = __syncthreads_and(ok);
            w = ok ? pick_winner(a.xchg, (int)(seq & 1), a.nranks) : -2;
            w_row = (w >= 0) ? a.xchg->hdr[seq & 1][w].row : -1;
        } else {
            if (threadIdx.x == 0) *a.cand = h;
            w = (r == INT_MAX || dry) ? -1 : a.rank;
            w_row = h.row;
        }
        if (threadIdx.x == 0) {
            st->ticket = 0;
            a.ucol[g.m] = g.T[(size_t)g.m * g.ld + c];
            if (dry) { st->last_r = (r == INT_MAX) ? -1 : (int)(g.row0 + r); st->last_c = c; }
            if (w == -2) st->status = LPS_E_PEER;
            st->sel_w = w;
            st->sel_row = w_row;
            if (solo) { s_w = w; s_row = w_row; st->xseq = seq; }
            else {
                __threadfence();
                st_release_gpu(&st->xseq, seq);
            }
        }
    }

    // ---- C: the decision, then the normalised pivot row ----
    if (threadIdx.x == 0 && !solo) {
        int w = -3;
        if (wait_seq_gpu(&st->xseq, seq)) { w = st->sel_w; s_row = st->sel_row; }
        s_w = w;
    }
    __syncthreads();
    const int w = s_w;
    if (w == -3) {                                             // the deciding CTA never showed up: do not hang
        if (threadIdx.x == 0) st->status = LPS_E_PEER;
        return;
    }
    if (w < 0) return;                                         // dry run, peer time-out, or no row with a_ic > 0 (k_update reports unbounded)
    if (w != a.rank) {
        if (blockIdx.x == 0 && threadIdx.x == 0 && !wait_seq(&a.xchg->row_seq, seq)) st->status = LPS_E_PEER;
        return;
    }
    const double* src = g.T + (size_t)(s_row - g.row0) * g.ld;
    const double inv = 1.0 / src[c];
    const bool spread = (a.xchg != nullptr);
    double* local = spread ? nullptr : reinterpret_cast<double*>(a.cand + 1);
    for (int j = 2 * (blockIdx.x * nt + threadIdx.x); j < (int)g.ld; j += 2 * gridDim.x * nt) {
        double2 x = *reinterpret_cast<const double2*>(src + j);
        if (inv != 1.0) { x.x = x.x * inv; x.y = x.y * inv; }
        if (spread) {
            for (int k = 0; k < a.nranks; ++k) *reinterpret_cast<double2*>(xchg_vrow(a.peers[k]) + j) = x;
        } else {
            *reinterpret_cast<double2*>(local + j) = x;
        }
    }
    if (!spread) return;
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(&st->ticket2, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!s_last) return;
    __threadfence_system();
    if ((int)threadIdx.x < a.nranks) st_release_sys(&a.peers[threadIdx.x]->row_seq, seq);
    if (threadIdx.x == 0) st->ticket2 = 0;
}

// Explicit pivot (Tableau.pivot(r,c)): gather column c and build the candidate for a caller-chosen (r,c).
__global__ void __launch_bounds__(kThreads) k_forced_f64(F64Grid g, SolveState* st, double* ucol, CandHeader* cand, int r, int c) {
    const double p = g.T[(size_t)r * g.ld + c];
    if (p == 0.0) {
        if (blockIdx.x == 0 && threadIdx.x == 0) st->status = LPS_ZERO_PIVOT;
        return;
    }
    for (int i = blockIdx.x * kThreads + threadIdx.x; i <= g.m; i += gridDim.x * kThreads) ucol[i] = g.T[(size_t)i * g.ld + c];
    double* vrow = reinterpret_cast<double*>(cand + 1);
    const double* src = g.T + (size_t)r * g.ld;
    const double inv = 1.0 / p;
    for (int j = blockIdx.x * kThreads + threadIdx.x; j < (int)g.ld; j += gridDim.x * kThreads)
        vrow[j] = (inv == 1.0) ? src[j] : src[j] * inv;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        cand->ratio = 0.0;
        cand->row = g.row0 + r;
        cand->c = c;
        cand->rule = 2;
        cand->stuck = st->stuck;
    }
}

// ------------------------------------------------------------------------------------------------
// update: T <- T - u (x) v on the local rows, row r <- v, fused pricing of the new objective row
// ------------------------------------------------------------------------------------------------
struct UpdateArgs {
    F64Grid g;
    PricePartials pp;
    SolveState* st;
    const double* ucol;
    const unsigned char* cand_all;   // legacy exchange: nranks slots of cand_stride bytes
    size_t cand_stride;
    const XchgBlock* xchg;           // peer-store exchange: this rank's block (NULL: legacy)
    int nranks;
    int rows_per_cta;                // <= kMaxRowBlock, multiple of RU
    int* basis;                      // m_total entries, replicated
    int* marks;                      // n entries (Tableau._cm as 0/1), replicated; may be NULL
    int* trace;                      // 3 * trace_cap ints or NULL
    long long trace_cap;
    int commit;                      // 0 for an explicit Tableau.pivot (no basis / trace / counter)
};

template <int CW, bool STREAM>
__global__ void __launch_bounds__(kThreads) k_update_f64(UpdateArgs a) {
    constexpr int VPT = CW / (2 * kThreads);   // double2 vectors per thread per row
    constexpr int RU = 4;                      // rows in flight per thread
    static_assert(VPT >= 1, "chunk too narrow");
    __shared__ alignas(128) double s_v[CW];
    __shared__ alignas(128) double s_u[kMaxRowBlock];
    __shared__ alignas(8) uint64_t s_bar;
    __shared__ ArgMinF64 s_arg[32];
    __shared__ int s_int[32];

    SolveState* st = a.st;
    const F64Grid& g = a.g;
    if (st->status != LPS_OK) return;

    // global leaving row: lowest ratio, then lowest global row (simplex.py:277), over all ranks' candidates
    int w = -1;
    long long w_row = -1;
    const CandHeader* hw = nullptr;
    const double* vrow = nullptr;
    if (a.xchg) {
        const int par = (int)(st->xseq & 1);
        w = pick_winner(a.xchg, par, a.nranks);
        if (w >= 0) { hw = &a.xchg->hdr[par][w]; w_row = hw->row; vrow = xchg_vrow(a.xchg); }
    } else {
        double w_ratio = 0.0;
        for (int k = 0; k < a.nranks; ++k) {
            const CandHeader* h = reinterpret_cast<const CandHeader*>(a.cand_all + (size_t)k * a.cand_stride);
            const long long row = h->row;
            if (row < 0) continue;
            const double ratio = h->ratio;
            if (w < 0 || ratio < w_ratio || (ratio == w_ratio && row < w_row)) { w = k; w_ratio = ratio; w_row = row; }
        }
        if (w >= 0) {
            hw = reinterpret_cast<const CandHeader*>(a.cand_all + (size_t)w * a.cand_stride);
            vrow = reinterpret_cast<const double*>(hw + 1);
        }
    }
    const bool lead = (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0);
    if (w < 0) {
        if (lead) st->status = LPS_UNBOUNDED;
        return;
    }
    const int c = hw->c;
    const long long r_local_ll = w_row - g.row0;
    const int r_local = (r_local_ll >= 0 && r_local_ll < g.m) ? (int)r_local_ll : -1;

    const int col0 = blockIdx.x * CW;
    const int row0 = blockIdx.y * a.rows_per_cta;
    const int mrows = g.m + 1;
    const int ncols = min(CW, (int)g.ld - col0);
    const int nrows = min(a.rows_per_cta, mrows - row0);

    // stage v (pivot row chunk) and u (pivot column block) in shared memory with TMA bulk copies
    if (threadIdx.x == 0) {
        mbar_init(&s_bar, 1);
        mbar_fence_init();
        const uint32_t vb = (uint32_t)ncols * 8u, ub = (uint32_t)a.rows_per_cta * 8u;
        mbar_expect_tx(&s_bar, vb + ub);
        tma_load_1d(s_v, vrow + col0, vb, &s_bar);
        tma_load_1d(s_u, a.ucol + row0, ub, &s_bar);
    }
    __syncthreads();
    mbar_wait(&s_bar, 0);

    double2 v[VPT];
    bool vok[VPT];
#pragma unroll
    for (int k = 0; k < VPT; ++k) {
        const int jl = 2 * threadIdx.x + k * 2 * kThreads;
        vok[k] = jl < ncols;
        v[k] = vok[k] ? *reinterpret_cast<const double2*>(&s_v[jl]) : make_double2(0.0, 0.0);
    }
    double* base = g.T + (size_t)row0 * g.ld + col0 + 2 * threadIdx.x;
    const bool has_obj = (row0 + nrows == mrows);
    double2 objv[VPT];

    for (int rg = 0; rg < nrows; rg += RU) {
        double2 t[RU][VPT];
        double nu[RU];
        bool act[RU];
#pragma unroll
        for (int q = 0; q < RU; ++q) {
            const int il = rg + q, i = row0 + il;
            const double u = (il < nrows) ? s_u[il] : 0.0;
            nu[q] = -u;
            // a zero multiplier leaves the row untouched (tableau.py:272); the objective row is
            // always loaded because its new values feed the pricing partials
            act[q] = (il < nrows) && (i != r_local) && (u != 0.0 || i == g.m);
            if (act[q]) {
#pragma unroll
                for (int k = 0; k < VPT; ++k)
                    if (vok[k]) t[q][k] = ld_f64x2<STREAM>(base + (size_t)il * g.ld + k * 2 * kThreads);
            }
        }
#pragma unroll
        for (int q = 0; q < RU; ++q) {
            const int il = rg + q, i = row0 + il;
            if (act[q]) {
                const bool write = (nu[q] != 0.0);
#pragma unroll
                for (int k = 0; k < VPT; ++k) {
                    if (!vok[k]) continue;
                    if (write) {
                        t[q][k].x = __dadd_rn(t[q][k].x, __dmul_rn(nu[q], v[k].x));
                        t[q][k].y = __dadd_rn(t[q][k].y, __dmul_rn(nu[q], v[k].y));
                        st_f64x2<STREAM>(base + (size_t)il * g.ld + k * 2 * kThreads, t[q][k]);
                    }
                    if (i == g.m) objv[k] = t[q][k];
                }
            } else if (il < nrows && i == r_local) {
#pragma unroll
                for (int k = 0; k < VPT; ++k)
                    if (vok[k]) st_f64x2<STREAM>(base + (size_t)il * g.ld + k * 2 * kThreads, v[k]);
            }
        }
    }

    if (has_obj) {   // price the new objective row chunk for the next select
        ArgMinF64 best{0.0, INT_MAX};
        int first_neg = INT_MAX;
#pragma unroll
        for (int k = 0; k < VPT; ++k) {
            if (!vok[k]) continue;
            const int j = col0 + 2 * threadIdx.x + k * 2 * kThreads;
            price_consider(objv[k].x, j, g.n, best, first_neg);
            price_consider(objv[k].y, j + 1, g.n, best, first_neg);
        }
        best = block_argmin(best, s_arg);
        first_neg = block_min_int(first_neg, s_int);
        if (threadIdx.x == 0) {
            a.pp.d_val[blockIdx.x] = best.key;
            a.pp.d_idx[blockIdx.x] = best.idx;
            a.pp.b_idx[blockIdx.x] = first_neg;
        }
    }

    if (lead) {   // commit (Simplex._pivot simplex.py:192-197: basis bookkeeping; marks live on the host)
        st->last_r = (int)w_row;
        st->last_c = c;
        if (a.commit) {
            const long long k = st->npiv;
            if (a.trace && k < a.trace_cap) {
                a.trace[3 * k] = (int)w_row;
                a.trace[3 * k + 1] = c;
                a.trace[3 * k + 2] = hw->rule;
            }
            if (a.marks) {   // simplex.py:195-197: unmark the leaving variable, mark the entering one
                const int old = a.basis[w_row];
                if (old >= 0 && old < g.n) a.marks[old] = 0;
                a.marks[c] = 1;
            }
            a.basis[w_row] = c;
            st->npiv = k + 1;
            st->rule = hw->rule;
            st->stuck = hw->stuck;
            st->pending_zcheck = (hw->rule == LPS_RULE_DANTZIG);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// small kernels outside the loop
// ------------------------------------------------------------------------------------------------
__global__ void k_row_scale_f64(F64Grid g, int r, double mult) {   // tableau.py:254-260
    double* row = g.T + (size_t)r * g.ld;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j <= g.n; j += gridDim.x * blockDim.x) row[j] = row[j] * mult;
}
__global__ void k_row_axpy_f64(F64Grid g, int rd, int rs, double mult) {   // tableau.py:269-289
    double* dst = g.T + (size_t)rd * g.ld;
    const double* src = g.T + (size_t)rs * g.ld;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j <= g.n; j += gridDim.x * blockDim.x)
        dst[j] = __dadd_rn(dst[j], __dmul_rn(mult, src[j]));
}
__global__ void k_gather_col_f64(F64Grid g, int j, double* out) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i <= g.m; i += gridDim.x * blockDim.x) out[i] = g.T[(size_t)i * g.ld + j];
}

// Tableau.isInfeasible (tableau.py:510-514): some row with b_i > 0 has no positive entry.  One warp per row.
__global__ void k_row_infeasible_f64(F64Grid g, int* flag) {
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    for (int i = blockIdx.x * wpb + (threadIdx.x >> 5); i < g.m; i += gridDim.x * wpb) {
        const double* row = g.T + (size_t)i * g.ld;
        if (!(row[g.n] > 0.0)) continue;
        bool pos = false;
        for (int j = lane; j < g.n && !pos; j += 32) pos = row[j] > 0.0;
        if (!__any_sync(0xffffffffu, pos) && lane == 0) atomicOr(flag, 1);
    }
}

// ------------------------------------------------------------------------------------------------
// leaving phase 1 on the device (Simplex._find_bfs simplex.py:95-105, every row kept)
// ------------------------------------------------------------------------------------------------
// cut the trailing columns n_new..g.n-1 (the artificial variables): the right-hand side `rhs` (column g.n, gathered
// before this launch) lands in column n_new, the columns behind it become padding (zero)
__global__ void k_cut_cols_f64(F64Grid g, int n_new, const double* rhs) {
    const int width = g.n - n_new + 1;
    const long long total = (long long)(g.m + 1) * width;
    for (long long x = (long long)blockIdx.x * blockDim.x + threadIdx.x; x < total; x += (long long)gridDim.x * blockDim.x) {
        const long long i = x / width;
        const int jj = (int)(x - i * width);
        g.T[(size_t)i * g.ld + n_new + jj] = jj == 0 ? rhs[i] : 0.0;
    }
}
// One step of simplex.py:104-105, `tab.rowSubFromObj(i, tab.getCj(j))` (-> tableau.py:282-293): the multiplier is the
// CURRENT reduced cost of row i's basis column -- in floating point the basis columns are not exact unit vectors
// (rowDiv multiplies by 1/a_rc, tableau.py:262-267), so earlier steps may have moved it by an ulp -- and every column
// adds mult * a_ij with product and sum rounded separately; a zero multiplier leaves the row as it is (tableau.py:285).
// src and dst are two buffers (the multiplier's own cell is rewritten by this step), swapped by the caller.
__global__ void k_obj_step_f64(F64Grid g, const double* src, double* dst, int row, int bcol) {
    const double ml = -src[bcol];
    const double* a = g.T + (size_t)row * g.ld;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j <= g.n; j += gridDim.x * blockDim.x)
        dst[j] = (ml == 0.0) ? src[j] : __dadd_rn(src[j], __dmul_rn(ml, a[j]));
}

// ------------------------------------------------------------------------------------------------
// whole-tableau pivot rules: Simplex.findPivotMaxIncrease simplex.py:286-328 and
// Simplex.findPivotAll simplex.py:330-360.  Both run the ratio test of EVERY column: one thread per column walks
// the rows (coalesced across the warp, the right-hand side is a broadcast load).
// ------------------------------------------------------------------------------------------------
// per column j: lowest ratio b_i / a_ij over a_ij > 0, its first row (strict "<": simplex.py:308), and how many
// rows attain it (findPivotAll lists them all); minrow = -1 when no entry is positive
__global__ void k_col_ratio_f64(F64Grid g, double* minratio, int* minrow, int* ties) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= g.n) return;
    double best = 0.0;
    int row = -1, cnt = 0;
    for (int i = 0; i < g.m; ++i) {
        const double a = g.T[(size_t)i * g.ld + j];
        if (!(a > 0.0)) continue;
        const double r = g.T[(size_t)i * g.ld + g.n] / a;
        if (row < 0 || r < best) { best = r; row = i; cnt = 1; }
        else if (r == best) ++cnt;
    }
    minratio[j] = best; minrow[j] = row; ties[j] = cnt;
}
// findPivotMaxIncrease: over the columns with c_j < 0, the pivot with the largest objective decrease -c_j * ratio,
// first column among equals (strict ">": simplex.py:321); a negative column without a positive entry: unbounded.
// out[0] = status (LPS_OK / LPS_OPTIMAL / LPS_UNBOUNDED), out[1] = row, out[2] = column.  One block.
__global__ void __launch_bounds__(1024) k_max_increase_f64(F64Grid g, const double* minratio, const int* minrow, int* out) {
    __shared__ ArgMinF64 s_arg[32];
    __shared__ int s_int[32];
    const double* cost = g.T + (size_t)g.m * g.ld;
    ArgMinF64 best{0.0, INT_MAX};                                 // key = -(decrease): arg-min with lowest index on ties
    int unb = INT_MAX;
    for (int j = threadIdx.x; j < g.n; j += blockDim.x) {
        const double cj = cost[j];
        if (cj >= 0.0 || cj != cj) continue;                      // "if cj >= ZERO: continue" (a NaN cost is never < 0)
        if (minrow[j] < 0) { unb = min(unb, j); continue; }
        const double dec = -cj * minratio[j];
        if (best.idx == INT_MAX || -dec < best.key) { best.key = -dec; best.idx = j; }
    }
    // any negative column at all?  encode "seen" through unb / best
    int seen = (best.idx != INT_MAX || unb != INT_MAX) ? 1 : 0;
    best = block_argmin(best, s_arg);
    unb = block_min_int(unb, s_int);
    seen = __syncthreads_or(seen);
    if (threadIdx.x == 0) {
        if (unb != INT_MAX) { out[0] = LPS_UNBOUNDED; out[1] = -1; out[2] = unb; }
        else if (!seen) { out[0] = LPS_OPTIMAL; out[1] = -1; out[2] = -1; }
        else { out[0] = LPS_OK; out[1] = minrow[best.idx]; out[2] = best.idx; }
    }
}
// findPivotAll, second pass: the rows of column j whose ratio equals the column's minimum, in row order, written at
// offset[j] (exclusive prefix sum of `ties` over the columns with minrow >= 0)
__global__ void k_col_ties_f64(F64Grid g, const double* minratio, const int* minrow, const long long* offset, int* rows, int* cols) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= g.n || minrow[j] < 0) return;
    const double best = minratio[j];
    long long at = offset[j];
    for (int i = minrow[j]; i < g.m; ++i) {
        const double a = g.T[(size_t)i * g.ld + j];
        if (!(a > 0.0)) continue;
        if (g.T[(size_t)i * g.ld + g.n] / a == best) { rows[at] = i; cols[at] = j; ++at; }
    }
}

// Tableau.isCanonical (tableau.py:466-496), column scan over a (column block) x (row slab) grid.  A column is basic
// when its cost is exactly 0 and it holds one exact 1 and zeros elsewhere.  colrow[j] = row of the 1 (claimed with a
// compare-and-swap, a second claim spoils the column), bad[j] != 0 once the column is known not to be basic -- the
// other slabs of a dense column then stop after one probe, so the scan reads little more than the basic columns.
__global__ void k_canon_init_f64(F64Grid g, int* colrow, int* bad) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= g.n) return;
    colrow[j] = -1;
    bad[j] = !(g.T[(size_t)g.m * g.ld + j] == 0.0);
}
__global__ void k_canon_cols_f64(F64Grid g, int rows_per_slab, int* colrow, int* bad) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= g.n) return;
    const int i0 = blockIdx.y * rows_per_slab, i1 = min(g.m, i0 + rows_per_slab);
    int hit = -1;
    for (int i = i0; i < i1; ++i) {
        if (((i - i0) & 15) == 0 && __ldcg(&bad[j])) return;
        const double v = g.T[(size_t)i * g.ld + j];
        if (v == 0.0) continue;
        if (v == 1.0 && hit == -1) hit = i;
        else { bad[j] = 1; return; }
    }
    if (hit != -1 && atomicCAS(&colrow[j], -1, hit) != -1) bad[j] = 1;
}
// any b_i < 0 ?  (tableau.py:476-477)
__global__ void k_any_neg_rhs_f64(F64Grid g, int* flag) {
    int neg = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < g.m; i += gridDim.x * blockDim.x)
        neg |= (g.T[(size_t)i * g.ld + g.n] < 0.0);
    if (__any_sync(0xffffffffu, neg) && (threadIdx.x & 31) == 0) atomicOr(flag, 1);
}
// bcols[i] = lowest basic column of row i (tableau.py:490-495)
__global__ void k_canon_rows(const int* colrow, const int* bad, int n, int* bcols) {
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
        const int i = (bad && bad[j]) ? -1 : colrow[j];
        if (i >= 0) atomicMin(&bcols[i], j);
    }
}


// ------------------------------------------------------------------------------------------------
// synthetic input: the seeded dense family of SURVEY.md section 8(d) generated in place with a
// counter-based hash (same function in oracle/pivot_oracle.c and problems.hashed_dense_grid), so any
// shard of the tableau can be produced independently on its own GPU
// ------------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint64_t fill_mix(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
__host__ __device__ __forceinline__ uint64_t fill_draw(uint64_t seed, uint64_t i, uint64_t j, uint64_t span) {
    return (fill_mix(fill_mix(seed) + i * 0x100000001B3ull + j) >> 33) % span;
}
constexpr uint64_t kFillSpecial = 0xFFFFFFFFull;   // pseudo row / column index of the cost row / rhs column

__global__ void k_fill_dense_f64(F64Grid g, uint64_t seed, int lo, int hi) {
    const long long mt = g.m_total;
    const uint64_t span = (uint64_t)(hi - lo + 1);
    const long long blo = (long long)hi * mt / 4, bhi = (long long)hi * mt / 2;
    for (int il = blockIdx.y; il <= g.m; il += gridDim.y) {
        double* row = g.T + (size_t)il * g.ld;
        const bool obj = (il == g.m);
        const uint64_t i = obj ? kFillSpecial : (uint64_t)(g.row0 + il);
        for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < (int)g.ld; j += gridDim.x * blockDim.x) {
            double v = 0.0;
            if (j < mt) {
                const double a = (double)(lo + (long long)fill_draw(seed, i, (uint64_t)j, span));
                v = obj ? -a : a;
            } else if (j < g.n) {
                v = (!obj && (long long)j - mt == (long long)i) ? 1.0 : 0.0;
            } else if (j == g.n && !obj) {
                v = (double)(blo + (long long)fill_draw(seed, i, kFillSpecial, (uint64_t)(bhi - blo + 1)));
            }
            row[j] = v;
        }
    }
}

}  // namespace lps
