// fp64 simplex pivots in BLOCKS: a look-ahead panel decides up to K consecutive pivots, then ONE pass over the
// tableau applies all of them.
//
// Why: one pivot of the reference (tableau.py:295-308) is a rank-1 update T <- T - u (x) v that reads and writes
// every cell -- 16 (m+1)(n+1) bytes of HBM traffic for 2 flops per cell.  k_update_f64 already moves those bytes at
// the copy peak, so the only way to more pivots per second is fewer bytes per pivot.  The decisions of the NEXT
// pivots do not need the updated tableau, only O(m+n) numbers of it:
//
//   * the objective row and the right-hand side column after each pivot (kept current in the panel: O, B),
//   * column c_l (the entering column) and row r_l (the leaving row) of the tableau as it WOULD be after the
//     first l pivots of the block: gathered from the stale tableau and corrected with the l rank-1 terms decided
//     so far (U[lp] = column snapshot, V[lp] = normalised pivot row of pivot lp).
//
// Every corrected number goes through exactly the operations the reference applies to that cell, in the same order
// (product rounded, then sum rounded; a zero multiplier leaves the cell alone, tableau.py:272/:285; the pivot row is
// REPLACED by its normalised copy, tableau.py:262-267), so panel and tableau stay bit-identical to a run that
// updates the whole tableau after every pivot.  k_blk_update then applies the K pivots to every cell with the same
// per-cell sequence: one read and one write of the tableau per K pivots.
//
//   k_blk_init      O <- objective row, B <- right-hand side, pricing partials (first block of a solve)
//   k_blk_panel     persistent grid (<= one CTA per SM), two grid-wide synchronisations per pivot:
//                     A  pricing from the partials (Dantzig simplex.py:262-267 / Bland :228-232), stuck counter, rule
//                        switch and objective assert (simplex.py:118-137), redundantly in every CTA
//                     B  column c gathered + corrected -> the rows' records, ratio test (simplex.py:272-279) over the
//                        CTA's rows; every CTA then reduces all partials itself
//                     C  the candidate row gathered + corrected + normalised -> V[l]; row-sharded tableaux: every rank
//                        sends a self-validating candidate record and its candidate row to every rank (ONE exchange
//                        per pivot, peer stores over NVLink) and every CTA picks the winner itself; objective row and
//                        right-hand side move on, pricing partials, basis / marks / trace are committed
//   k_blk_panel_cl  the same decisions for tableaux whose panel fits the shared memory of ONE thread-block cluster:
//                   state in distributed shared memory, partials through DSMEM, hardware cluster barriers
//   k_blk_update    T <- T with the block's pivots applied: persistent, warp-specialised, TMA ring in shared memory
#pragma once
#include <cooperative_groups.h>

#include "lps_f64.cuh"

namespace lps {

constexpr int kBlkMax = 16;   // pivots per block (upper bound; the fp64 pipe saturates near 22 at the HBM rate)
constexpr int kBlkCW = 512;   // columns per update tile: 256 threads x one double2
constexpr int kBlkRU = 8;     // rows per stage of the fused update
constexpr int kBlkRec = kBlkMax + 2;   // doubles per row record of the panel's U (16-byte multiple: TMA source)
constexpr int kBlkStages = 6;          // stages of the update's shared-memory ring
constexpr int kBlkConsumers = 256;     // consumer threads of the fused update (one double2 column each)

struct alignas(128) BlkSync {        // zeroed when the panel state is (re)read from the grid; counters only ever grow
    unsigned long long arrive1;      // ratio-test partials published (grid-wide synchronisation 1 of a pivot)
    unsigned long long pad1[15];
    unsigned long long arrive2;      // end-of-pivot synchronisation
    unsigned long long pad2[15];
    unsigned long long exit_ticket;  // the CTA that leaves a launch last writes the solver state back
    unsigned long long base1, base2; // arrive1 / arrive2 when the running launch started (written by that last CTA): a
                                     // CTA knows every synchronisation's target without reading back its own arrival,
                                     // so arrivals are fire-and-forget release reductions
    unsigned long long pad3[13];
};

struct BlkPanel {
    double* U;               // upitch records of kBlkRec doubles, one per local row (record m: the objective row):
                             //   [l] for l < kBlkMax = cell (row, c_l) just before pivot l, the row's multiplier in pivot l
                             //   [kBlkMax] = two 32-bit masks: low, bit l = that multiplier is zero (the reference leaves
                             //   the row alone, tableau.py:272); high, bit l = the row IS pivot l's row (replaced)
    size_t upitch;
    double* V;               // the exchange block's row area [set 2][rank nr][pivot kBlkMax][ld]: normalised candidate rows.
                             // Pivot l's row is the one of the rank that won its ratio test, W[l]; blocks alternate
                             // between the two sets (a peer may already store its first candidate of the next block
                             // while this rank's update still reads the current one)
    int nr;                  // rank slots of the row area
    int* W;                  // kBlkMax: winning rank of pivot l
    double* O[2];            // objective row, ping-pong by pivot parity (ld each)
    double* B[2];            // right-hand side of the local rows, ping-pong (upitch each)
    int* R;                  // kBlkMax: GLOBAL leaving row of pivot l
    int* C;                  // kBlkMax: entering column
    int* hdr;                // [0] pivots decided by the last panel launch, [1] current parity of O / B,
                             // [2] non-empty blocks since lps_solve_begin, [3] row-area set of the last panel launch
    double* p_val;           // pricing partials, one per panel CTA
    int* p_idx;
    int* p_bidx;
    ArgMinF64* rblk;         // ratio-test partials, one per panel CTA
    BlkSync* sync;
};

__device__ __forceinline__ double2 ldcg2(const double* p) { return __ldcg(reinterpret_cast<const double2*>(p)); }
__device__ __forceinline__ void stcg2(double* p, double2 v) { __stcg(reinterpret_cast<double2*>(p), v); }
// one cell through one pivot of the block: u = the cell of its row in the pivot column, v = the cell of its column
// in the normalised pivot row (tableau.py:269-280: row += (-u) * pivot row, skipped when u == 0)
__device__ __forceinline__ double blk_apply(double x, double u, double v) {
    return (u == 0.0) ? x : __dadd_rn(x, __dmul_rn(-u, v));
}

// ------------------------------------------------------------------------------------------------
// init: panel state from the tableau
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kSelMaxThreads) k_blk_init(F64Grid g, BlkPanel P) {
    __shared__ ArgMinF64 s_arg[32];
    __shared__ int s_int[32];
    const int nt = (int)blockDim.x, G = (int)gridDim.x;
    const double* obj = g.T + (size_t)g.m * g.ld;
    ArgMinF64 best{0.0, INT_MAX};
    int first_neg = INT_MAX;
    for (int j = blockIdx.x * nt + threadIdx.x; j < (int)g.ld; j += G * nt) {
        const double v = obj[j];
        P.O[0][j] = v;
        price_consider(v, j, g.n, best, first_neg);
    }
    for (int i = blockIdx.x * nt + threadIdx.x; i < g.m; i += G * nt) P.B[0][i] = g.T[(size_t)i * g.ld + g.n];
    best = block_argmin(best, s_arg);
    first_neg = block_min_int(first_neg, s_int);
    if (threadIdx.x == 0) {
        P.p_val[blockIdx.x] = best.key;
        P.p_idx[blockIdx.x] = best.idx;
        P.p_bidx[blockIdx.x] = first_neg;
        if (blockIdx.x == 0) { P.hdr[0] = 0; P.hdr[1] = 0; P.hdr[2] = 0; P.hdr[3] = 0; }
    }
}

// ------------------------------------------------------------------------------------------------
// panel: up to kmax pivot decisions in one launch
// ------------------------------------------------------------------------------------------------
struct BlkPanelArgs {
    F64Grid g;
    BlkPanel p;
    SolveState* st;
    int kmax;
    XchgBlock* xchg;             // this rank's exchange block (headers, flags; V follows it)
    XchgBlock* const* peers;     // every rank's exchange block
    int rank, nranks;
    int* basis;
    int* marks;
    int* trace;
    long long trace_cap;
};

// ---- candidate records of the one-exchange protocol (XchgBlock::hdr2) ----
constexpr unsigned long long kHdrMagic = 0x5bd1e995a5a5a5a5ull;
__device__ __forceinline__ void hdr2_store(unsigned long long* slot, double ratio, long long row, unsigned long long seq) {
    const unsigned long long w0 = (unsigned long long)__double_as_longlong(ratio);
    const unsigned long long w1 = (unsigned long long)(unsigned int)(row < 0 ? 0xffffffffu : (unsigned int)row) | (seq << 32);
    asm volatile("st.volatile.global.v2.u64 [%0], {%1,%2};" ::"l"(slot), "l"(w0), "l"(w1) : "memory");
    asm volatile("st.volatile.global.v2.u64 [%0], {%1,%2};" ::"l"(slot + 2), "l"(w0 ^ w1 ^ kHdrMagic), "l"(0ull) : "memory");
}
// true once the record of sequence number seq has landed whole; ten-second time-out like wait_seq
__device__ __forceinline__ bool hdr2_wait(const unsigned long long* slot, unsigned long long seq, double* ratio, long long* row) {
    const unsigned long long t0 = global_ns();
    unsigned int polls = 0;
    for (;;) {
        const unsigned long long w1 = ld_relaxed_sys(slot + 1);
        if ((w1 >> 32) == (seq & 0xffffffffull)) {
            const unsigned long long w0 = ld_relaxed_sys(slot), w2 = ld_relaxed_sys(slot + 2);
            if ((w0 ^ w1 ^ kHdrMagic) == w2) {
                *ratio = __longlong_as_double((long long)w0);
                const unsigned int r32 = (unsigned int)(w1 & 0xffffffffull);
                *row = (r32 == 0xffffffffu) ? -1 : (long long)r32;
                return true;
            }
        }
        if ((++polls & 1023u) == 0 && global_ns() - t0 > 10000000000ull) return false;
    }
}

// ---- slice records of the one-exchange protocol (XchgBlock::rowrec) ----
// checksum of a row slice: sum over its cells of (bit pattern) x (odd weight from the column index), modulo 2^64
__device__ __forceinline__ unsigned long long slice_term(double2 y, int j) {
    return (unsigned long long)__double_as_longlong(y.x) * (2ull * (unsigned long long)j + 1ull) +
           (unsigned long long)__double_as_longlong(y.y) * (2ull * (unsigned long long)j + 3ull);
}
__device__ __forceinline__ unsigned long long block_sum_u64(unsigned long long v, unsigned long long* scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    unsigned long long r = (lane < nwarps) ? scratch[lane] : 0ull;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) r += __shfl_xor_sync(0xffffffffu, r, off);
    return r;
}
__device__ __forceinline__ void rowrec_store(unsigned long long* slot, unsigned long long sum, unsigned long long seq, double vn) {
    const unsigned long long w2 = (unsigned long long)__double_as_longlong(vn);
    asm volatile("st.volatile.global.v2.u64 [%0], {%1,%2};" ::"l"(slot), "l"(sum), "l"(seq) : "memory");
    asm volatile("st.volatile.global.v2.u64 [%0], {%1,%2};" ::"l"(slot + 2), "l"(w2), "l"(sum ^ seq ^ w2 ^ kHdrMagic) : "memory");
}
__device__ __forceinline__ bool rowrec_wait(const unsigned long long* slot, unsigned long long seq, unsigned long long* sum, double* vn) {
    const unsigned long long t0 = global_ns();
    unsigned int polls = 0;
    for (;;) {
        const unsigned long long w1 = ld_relaxed_sys(slot + 1);
        if (w1 == seq) {
            const unsigned long long w0 = ld_relaxed_sys(slot), w2 = ld_relaxed_sys(slot + 2), w3 = ld_relaxed_sys(slot + 3);
            if ((w0 ^ w1 ^ w2 ^ kHdrMagic) == w3) {
                *sum = w0;
                *vn = __longlong_as_double((long long)w2);
                return true;
            }
        }
        if ((++polls & 1023u) == 0 && global_ns() - t0 > 10000000000ull) return false;
    }
}

template <int NT>
__global__ void __launch_bounds__(NT, 1) k_blk_panel(BlkPanelArgs a) {
    constexpr int kBlkBatch = (NT > 512) ? 4 : 16;   // loads of earlier pivots' multipliers issued together (register budget)
    __shared__ ArgMinF64 s_arg[32];
    __shared__ int s_int[32];
    __shared__ double s_vc[kBlkMax];       // V[lp][c]: the entering column's cell in the earlier pivot rows
    __shared__ double s_ur[kBlkMax];       // U[lp][r]: the leaving row's cell in the earlier pivot columns
    __shared__ long long s_R[kBlkMax];     // global leaving rows decided so far in this block
    __shared__ int s_W[kBlkMax];           // ranks that owned them
    __shared__ double s_hratio[kMaxRanks]; // this pivot's candidates of all ranks
    __shared__ long long s_hrow[kMaxRanks];
    __shared__ int s_flag;
    __shared__ long long s_row;
    __shared__ unsigned long long s_sum[32], s_want;
    __shared__ double s_vn;

    const int nt = (int)blockDim.x, G = (int)gridDim.x, tid = (int)threadIdx.x, cta = (int)blockIdx.x;
    const bool solo = (G == 1);
    SolveState* st = a.st;
    const F64Grid& g = a.g;
    const BlkPanel& P = a.p;
    BlkSync* sy = P.sync;

    // The state is read here and written back by the CTA that leaves last: no CTA can see a half-updated state,
    // and every CTA (every rank) walks through identical decisions.
    int status = st->status;
    if (status != LPS_OK) {
        if (cta == 0 && tid == 0) P.hdr[0] = 0;
        return;
    }
    int rule = st->rule;
    long long stuck = st->stuck, npiv = st->npiv;
    const long long maxp = st->max_pivots;
    int pending = st->pending_zcheck;
    const double z_entry = st->z_entry_cell;
    unsigned long long seq = st->xseq;
    int last_r = st->last_r, last_c = st->last_c;
    int cur = P.hdr[1];
    const unsigned long long base1 = solo ? 0ull : sy->base1, base2 = solo ? 0ull : sy->base2;
    unsigned long long nb1 = 0, nb2 = 0;                       // grid-wide synchronisations passed in this launch
    const int set = P.hdr[2] & 1;                              // row-area set of this block
    const size_t vset = (size_t)set * P.nr;
    auto Vrow = [&](int lp, int w) { return P.V + ((vset + w) * kBlkMax + lp) * g.ld; };   // rank w's candidate row of pivot lp

    int l = 0;
    for (; l < a.kmax; ++l) {
        if (maxp >= 0 && npiv >= maxp) { status = LPS_MAX_PIVOTS; break; }
        const double* Oc = P.O[cur];
        double* On = P.O[cur ^ 1];
        const double* Bc = P.B[cur];
        double* Bn = P.B[cur ^ 1];
        // ---- A: stuck counter and rule (simplex.py:123,132-137), entering column ----
        long long stuck_now = stuck;
        if (pending) {
            const double corner = __ldcg(&Oc[g.n]);
            // simplex.py:133 `assert z <= obj_val`: the objective (minus the corner cell) may not rise above its
            // value at entry
            if (!(corner >= z_entry)) { status = LPS_OBJ_INCREASED; break; }
            stuck_now = (corner == z_entry) ? stuck + 1 : 0;
        }
        int rule_now = rule;
        if (rule_now == LPS_RULE_DANTZIG && !(stuck_now < g.m_total + (long long)g.n)) rule_now = LPS_RULE_BLAND;
        int c;
        if (rule_now == LPS_RULE_DANTZIG) {
            ArgMinF64 best{0.0, INT_MAX};
            for (int k = tid; k < G; k += nt) best = argmin_pick(best, ArgMinF64{__ldcg(&P.p_val[k]), __ldcg(&P.p_idx[k])});
            best = block_argmin(best, s_arg);
            c = best.idx;
        } else {
            int first = INT_MAX;
            for (int k = tid; k < G; k += nt) first = min(first, __ldcg(&P.p_bidx[k]));
            c = block_min_int(first, s_int);
        }
        if (c == INT_MAX) { status = LPS_OPTIMAL; break; }

        // ---- B: column c as it stands after the block's earlier pivots -> U[l]; ratio test ----
        // (the first row's loads are issued before the block waits for the column's cells of the earlier pivot rows)
        const int i_first = cta * nt + tid;
        double x_first = 0.0, b_first = 0.0;
        if (i_first < g.m) { x_first = g.T[(size_t)i_first * g.ld + c]; b_first = __ldcg(&Bc[i_first]); }
        if (tid < l) s_vc[tid] = __ldcg(Vrow(tid, s_W[tid]) + c);
        const double uobj = __ldcg(&Oc[c]);                    // the objective row's cell of the entering column
        __syncthreads();
        ArgMinF64 best{0.0, INT_MAX};
        for (int i = i_first; i < g.m; i += G * nt) {
            double x = (i == i_first) ? x_first : g.T[(size_t)i * g.ld + c];
            const double bi = (i == i_first) ? b_first : __ldcg(&Bc[i]);
            const long long gi = g.row0 + i;
            // the multipliers of the earlier pivots are independent loads: issued in batches, applied in order
            for (int l0 = 0; l0 < l; l0 += kBlkBatch) {
                double u[kBlkBatch];
#pragma unroll
                for (int q = 0; q < kBlkBatch; q += 2)
                    if (l0 + q < l) {
                        const double2 uu = ldcg2(P.U + (size_t)i * kBlkRec + l0 + q);
                        u[q] = uu.x;
                        u[q + 1] = uu.y;
                    }
#pragma unroll
                for (int q = 0; q < kBlkBatch; ++q)
                    if (l0 + q < l) x = (s_R[l0 + q] == gi) ? s_vc[l0 + q] : blk_apply(x, u[q], s_vc[l0 + q]);
            }
            {   // the row's record: multiplier of pivot l, zero-multiplier mask (a new block starts new masks)
                double* rec = P.U + (size_t)i * kBlkRec;
                __stcg(&rec[l], x);
                uint2 mk = (l == 0) ? make_uint2(0u, 0u) : __ldcg(reinterpret_cast<const uint2*>(rec + kBlkMax));
                if (x == 0.0) mk.x |= 1u << l;
                if (l == 0 || x == 0.0) __stcg(reinterpret_cast<uint2*>(rec + kBlkMax), mk);
            }
            if (x > 0.0) {
                const double ratio = bi / x;
                if (best.idx == INT_MAX || ratio < best.key) { best.key = ratio; best.idx = i; }
            }
        }
        if (cta == 0 && tid == 0) {
            double* rec = P.U + (size_t)g.m * kBlkRec;
            __stcg(&rec[l], uobj);
            uint2 mk = (l == 0) ? make_uint2(0u, 0u) : __ldcg(reinterpret_cast<const uint2*>(rec + kBlkMax));
            if (uobj == 0.0) mk.x |= 1u << l;
            if (l == 0 || uobj == 0.0) __stcg(reinterpret_cast<uint2*>(rec + kBlkMax), mk);
        }
        best = block_argmin(best, s_arg);
        seq += 1;
        // this rank's best row: every CTA publishes its partial, waits for all of them and reduces them itself (one
        // hop; a deciding CTA that reduces and then publishes the decision costs two)
        if (!solo) {
            if (tid == 0) {
                __stcg(&P.rblk[cta].key, best.key);
                __stcg(&P.rblk[cta].idx, best.idx);
                red_release_gpu(&sy->arrive1);
                s_flag = wait_seq_gpu(&sy->arrive1, base1 + (nb1 + 1) * (unsigned long long)G) ? 1 : 0;
            }
            __syncthreads();
            if (!s_flag) { status = LPS_E_PEER; last_r = -101; break; }   // (which wait gave up: reported by lps_solve_end)
            ++nb1;
            best = ArgMinF64{0.0, INT_MAX};
            for (int k = tid; k < G; k += nt) best = argmin_pick(best, ArgMinF64{__ldcg(&P.rblk[k].key), __ldcg(&P.rblk[k].idx)});
            best = block_argmin(best, s_arg);
        }
        {
            const int r = best.idx;
            const long long cand = (r == INT_MAX) ? -1 : g.row0 + r;
            // row-sharded: the rank's candidate record goes to every rank (one CTA sends it)
            if (a.nranks > 1 && cta == 0 && tid < a.nranks) hdr2_store(&a.peers[tid]->hdr2[seq & 1][a.rank][0], best.key, cand, seq);
            if (tid == 0) s_row = cand;
        }
        __syncthreads();
        const long long cand_g = s_row;                         // this rank's candidate (global row), -1: none

        // ---- C: normalised candidate row -> the row area (every rank's copy); winner; objective row, right-hand
        //         side, pricing ----
        ArgMinF64 pb{0.0, INT_MAX};
        int first_neg = INT_MAX;
        double vn;                                              // the pivot row's right-hand side, normalised
        int w;
        long long r_g;
        if (a.nranks == 1) {
            if (cand_g < 0) { status = LPS_UNBOUNDED; break; }
            w = 0;
            r_g = cand_g;
        } else { w = -1; r_g = -1; }
        if (cand_g >= 0) {
            const int r_loc = (int)(cand_g - g.row0);
            const double* src = g.T + (size_t)r_loc * g.ld;
            const int j_first = 2 * (cta * nt + tid);
            double2 y_first = make_double2(0.0, 0.0), o_first = make_double2(0.0, 0.0);
            if (j_first < (int)g.ld) { y_first = *reinterpret_cast<const double2*>(src + j_first); o_first = ldcg2(Oc + j_first); }
            if (tid < l) s_ur[tid] = __ldcg(&P.U[(size_t)r_loc * kBlkRec + tid]);
            const double pv = __ldcg(&P.U[(size_t)r_loc * kBlkRec + l]);
            const double br = __ldcg(&Bc[r_loc]);
            __syncthreads();
            const double inv = 1.0 / pv;                        // rowDiv: ONE / d, then rowMult (tableau.py:262-267)
            vn = (inv != 1.0) ? __dmul_rn(br, inv) : br;
            double* Vl = Vrow(l, a.rank);
            const size_t peer_off = (size_t)(Vl - P.V);
            unsigned long long csum = 0ull;
            for (int j = j_first; j < (int)g.ld; j += 2 * G * nt) {
                double2 y = (j == j_first) ? y_first : *reinterpret_cast<const double2*>(src + j);
                for (int l0 = 0; l0 < l; l0 += kBlkBatch / 2) {
                    double2 v[kBlkBatch / 2];
#pragma unroll
                    for (int q = 0; q < kBlkBatch / 2; ++q)
                        if (l0 + q < l) v[q] = ldcg2(Vrow(l0 + q, s_W[l0 + q]) + j);
#pragma unroll
                    for (int q = 0; q < kBlkBatch / 2; ++q) {
                        if (l0 + q >= l) continue;
                        if (s_R[l0 + q] == cand_g) y = v[q];
                        else { y.x = blk_apply(y.x, s_ur[l0 + q], v[q].x); y.y = blk_apply(y.y, s_ur[l0 + q], v[q].y); }
                    }
                }
                if (inv != 1.0) { y.x = __dmul_rn(y.x, inv); y.y = __dmul_rn(y.y, inv); }   // rowMult skips a multiplier of 1
                stcg2(Vl + j, y);
                if (a.nranks == 1) {                            // one rank: the candidate is the pivot row
                    double2 o = (j == j_first) ? o_first : ldcg2(Oc + j);
                    o.x = blk_apply(o.x, uobj, y.x);
                    o.y = blk_apply(o.y, uobj, y.y);
                    stcg2(On + j, o);
                    price_consider(o.x, j, g.n, pb, first_neg);
                    price_consider(o.y, j + 1, g.n, pb, first_neg);
                } else {
                    for (int k = 0; k < a.nranks; ++k)
                        if (k != a.rank) *reinterpret_cast<double2*>(xchg_vrow(a.peers[k]) + peer_off + j) = y;
                    csum += slice_term(y, j);
                }
            }
            if (a.nranks > 1) {
                // this CTA's slice is out: its record (checksum, sequence number, the row's right-hand side) follows it to
                // every rank -- no fence; a receiver whose copy of the slice does not add up yet reads it again
                csum = block_sum_u64(csum, s_sum);
                if (tid < a.nranks) rowrec_store(&a.peers[tid]->rowrec[seq & 1][a.rank][cta][0], csum, seq, vn);
            }
        }
        if (a.nranks > 1) {
            // every CTA reads all candidates and picks the same winner: lowest ratio, then lowest global row
            // (simplex.py:277); then it needs the winner's row only where it works -- its own slice -- and the row's
            // right-hand side, which travels in every slice record
            if (tid < a.nranks) {
                double hr = 0.0;
                long long hrow = -1;
                const bool ok = hdr2_wait(&a.xchg->hdr2[seq & 1][tid][0], seq, &hr, &hrow);
                s_hratio[tid] = hr;
                s_hrow[tid] = ok ? hrow : -2;
            }
            __syncthreads();
            bool dead = false;
            for (int k = 0; k < a.nranks; ++k) {
                const long long row = s_hrow[k];
                if (row == -2) dead = true;
                if (row < 0) continue;
                if (w < 0 || s_hratio[k] < s_hratio[w] || (s_hratio[k] == s_hratio[w] && row < r_g)) { w = k; r_g = row; }
            }
            if (dead) { status = LPS_E_PEER; last_r = -102; break; }
            if (w < 0) { status = LPS_UNBOUNDED; break; }
            if (tid == 0) {
                unsigned long long want = 0ull;
                double vnw = 0.0;
                s_flag = rowrec_wait(&a.xchg->rowrec[seq & 1][w][cta][0], seq, &want, &vnw) ? 1 : 0;
                s_want = want;
                s_vn = vnw;
            }
            __syncthreads();
            if (!s_flag) { status = LPS_E_PEER; last_r = -103; break; }
            vn = s_vn;
            const unsigned long long want = s_want;
            const double* Vw = Vrow(l, w);
            for (unsigned int tries = 0;; ++tries) {
                // objective row over my slice of the winner's row; the slice's checksum on the way (the record may
                // have overtaken the data: then the sums differ and the slice is read again)
                unsigned long long got = 0ull;
                pb = ArgMinF64{0.0, INT_MAX};
                first_neg = INT_MAX;
                for (int j = 2 * (cta * nt + tid); j < (int)g.ld; j += 2 * G * nt) {
                    const double2 y = ldcg2(Vw + j);
                    got += slice_term(y, j);
                    double2 o = ldcg2(Oc + j);
                    o.x = blk_apply(o.x, uobj, y.x);
                    o.y = blk_apply(o.y, uobj, y.y);
                    stcg2(On + j, o);
                    price_consider(o.x, j, g.n, pb, first_neg);
                    price_consider(o.y, j + 1, g.n, pb, first_neg);
                }
                got = block_sum_u64(got, s_sum);
                if (got == want) break;
                if (tries > (1u << 22)) { status = LPS_E_PEER; last_r = -104; break; }   // (uniform across the block: several seconds)
            }
            if (status != LPS_OK) break;
        }
        for (int i = cta * nt + tid; i < g.m; i += G * nt) {
            double b;
            double* rec = P.U + (size_t)i * kBlkRec;
            if (g.row0 + i == r_g) {
                b = vn;
                uint2 mk = __ldcg(reinterpret_cast<const uint2*>(rec + kBlkMax));   // this row is pivot l's row
                mk.y |= 1u << l;
                __stcg(reinterpret_cast<uint2*>(rec + kBlkMax), mk);
            } else b = blk_apply(__ldcg(&Bc[i]), __ldcg(&rec[l]), vn);
            __stcg(&Bn[i], b);
        }
        pb = block_argmin(pb, s_arg);
        first_neg = block_min_int(first_neg, s_int);
        if (tid == 0) {
            __stcg(&P.p_val[cta], pb.key);
            __stcg(&P.p_idx[cta], pb.idx);
            __stcg(&P.p_bidx[cta], first_neg);
            s_R[l] = r_g;
            s_W[l] = w;
            if (cta == 0) {   // commit (Simplex._pivot simplex.py:192-197)
                P.R[l] = (int)r_g;
                P.C[l] = c;
                P.W[l] = w;
                if (a.trace && npiv < a.trace_cap) {
                    a.trace[3 * npiv] = (int)r_g;
                    a.trace[3 * npiv + 1] = c;
                    a.trace[3 * npiv + 2] = rule_now;
                }
                if (a.marks) {
                    const int old = a.basis[r_g];
                    if (old >= 0 && old < g.n) a.marks[old] = 0;
                    a.marks[c] = 1;
                }
                a.basis[r_g] = c;
            }
        }
        npiv += 1;
        rule = rule_now;
        stuck = stuck_now;
        pending = (rule_now == LPS_RULE_DANTZIG);
        last_r = (int)r_g;
        last_c = c;
        cur ^= 1;
        // ---- end of pivot: everything this pivot wrote is visible before the next one reads it ----
        if (solo) __syncthreads();
        else {
            __syncthreads();
            if (tid == 0) {
                red_release_gpu(&sy->arrive2);
                s_flag = wait_seq_gpu(&sy->arrive2, base2 + (nb2 + 1) * (unsigned long long)G) ? 1 : 0;
            }
            __syncthreads();
            if (!s_flag) { status = LPS_E_PEER; last_r = -105; break; }
            ++nb2;
        }
    }
    if (status == LPS_OK && maxp >= 0 && npiv >= maxp) status = LPS_MAX_PIVOTS;

    __syncthreads();
    if (tid == 0) {
        if (cta == 0) {
            P.hdr[0] = l;
            P.hdr[1] = cur;
            P.hdr[3] = set;
            if (l > 0) P.hdr[2] = P.hdr[2] + 1;
        }
        bool last = true;
        if (!solo) last = ((atomicAdd(&sy->exit_ticket, 1ull) + 1) % (unsigned long long)G == 0);
        if (last) {
            if (!solo) { sy->base1 = base1 + nb1 * (unsigned long long)G; sy->base2 = base2 + nb2 * (unsigned long long)G; }
            st->status = status;
            st->rule = rule;
            st->stuck = stuck;
            st->npiv = npiv;
            st->pending_zcheck = pending;
            st->xseq = seq;
            st->last_r = last_r;
            st->last_c = last_c;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// panel for medium tableaux: one thread-block cluster, the whole panel in distributed shared memory
// ------------------------------------------------------------------------------------------------
// k_blk_panel spends most of a pivot on round trips to L2 and on three grid-wide synchronisations through global memory
// (profiles/r02e_panel_phases.txt).  When the panel fits the shared memory of ONE cluster -- K (m + ld) doubles, up to
// about 7000 rows at n = 2m with 16 CTAs -- none of that is needed: CTA k of the cluster owns rows [k rpc, (k+1) rpc)
// (their multipliers U, right-hand side) and columns [k cpc, (k+1) cpc) (their pivot-row cells V, objective row), all
// in its shared memory; the only global reads of a pivot are the two gathers from the grid (entering column, leaving
// row), partial arg-mins are written into every CTA's shared memory (DSMEM) and the two synchronisations per pivot
// are hardware cluster barriers.  Same decisions, same records for k_blk_update (written through to global memory).
// One rank only.
namespace cg = cooperative_groups;
constexpr int kClThreads = 512;
constexpr int kClMax = 16;

struct BlkClusterArgs {
    F64Grid g;
    BlkPanel p;
    SolveState* st;
    int kmax;
    int rpc, cpc;        // rows / columns per CTA (cpc even)
    int* basis;
    int* marks;
    int* trace;
    long long trace_cap;
};
__host__ __device__ inline size_t blk_cluster_smem(int rpc, int cpc) {
    return ((size_t)(kBlkMax + 2) * (rpc + cpc)) * sizeof(double) + (size_t)rpc * sizeof(uint2) + 256;
}

__global__ void __launch_bounds__(kClThreads, 1) k_blk_panel_cl(BlkClusterArgs a) {
    cg::cluster_group cl = cg::this_cluster();
    const int C = (int)cl.num_blocks(), me = (int)cl.block_rank();
    const int tid = (int)threadIdx.x, nt = kClThreads;
    extern __shared__ __align__(16) unsigned char cl_smem[];
    const int rpc = a.rpc, cpc = a.cpc;
    double* sU = reinterpret_cast<double*>(cl_smem);            // [kBlkMax][rpc]  multipliers of my rows
    double* sV = sU + (size_t)kBlkMax * rpc;                    // [kBlkMax][cpc]  pivot-row cells of my columns
    double* sO = sV + (size_t)kBlkMax * cpc;                    // [2][cpc]        objective row, my columns
    double* sB = sO + 2 * (size_t)cpc;                          // [2][rpc]        right-hand side, my rows
    uint2* sM = reinterpret_cast<uint2*>(sB + 2 * (size_t)rpc); // [rpc]           zero / pivot-row masks of my rows
    __shared__ ArgMinF64 s_arg[32];
    __shared__ int s_int[32];
    __shared__ ArgMinF64 s_rat[kClMax], s_prc[kClMax];          // partials of every CTA (written remotely)
    __shared__ int s_bld[kClMax];
    __shared__ double s_vc[kBlkMax], s_ur[kBlkMax];
    __shared__ long long s_R[kBlkMax];

    SolveState* st = a.st;
    const F64Grid& g = a.g;
    const BlkPanel& P = a.p;
    int status = st->status;
    if (status != LPS_OK) {
        if (me == 0 && tid == 0) P.hdr[0] = 0;
        return;
    }
    int rule = st->rule;
    long long stuck = st->stuck, npiv = st->npiv;
    const long long maxp = st->max_pivots;
    int pending = st->pending_zcheck;
    const double z_entry = st->z_entry_cell;
    int last_r = st->last_r, last_c = st->last_c;
    int cur = P.hdr[1];
    const int set = P.hdr[2] & 1;
    const size_t vset = (size_t)set * P.nr;
    const int row_lo = me * rpc, nrows = max(0, min(rpc, g.m - row_lo));
    const int col_lo = me * cpc, ncols = max(0, min(cpc, (int)g.ld - col_lo));
    auto remote = [&](auto* p, int rank) { return cl.map_shared_rank(p, rank); };

    // panel state in: my slices of the objective row and the right-hand side; pricing partials of my slice
    {
        ArgMinF64 pb{0.0, INT_MAX};
        int first_neg = INT_MAX;
        for (int jl = tid; jl < ncols; jl += nt) {
            const double v = P.O[cur][col_lo + jl];
            sO[(size_t)cur * cpc + jl] = v;
            price_consider(v, col_lo + jl, g.n, pb, first_neg);
        }
        for (int il = tid; il < nrows; il += nt) sB[(size_t)cur * rpc + il] = P.B[cur][row_lo + il];
        pb = block_argmin(pb, s_arg);
        first_neg = block_min_int(first_neg, s_int);
        if (tid < C) { *remote(&s_prc[me], tid) = pb; *remote(&s_bld[me], tid) = first_neg; }
    }
    cl.sync();

    int l = 0;
    for (; l < a.kmax; ++l) {
        if (maxp >= 0 && npiv >= maxp) { status = LPS_MAX_PIVOTS; break; }
        const double* Oc = sO + (size_t)cur * cpc;
        double* On = sO + (size_t)(cur ^ 1) * cpc;
        const double* Bc = sB + (size_t)cur * rpc;
        double* Bn = sB + (size_t)(cur ^ 1) * rpc;
        const int cta_n = g.n / cpc;                              // owner of the corner cell / right-hand side column

        // ---- A: stuck counter and rule, entering column (from the partials every CTA holds) ----
        long long stuck_now = stuck;
        if (pending) {
            const double corner = *remote(sO + (size_t)cur * cpc + (g.n - cta_n * cpc), cta_n);
            if (!(corner >= z_entry)) { status = LPS_OBJ_INCREASED; break; }
            stuck_now = (corner == z_entry) ? stuck + 1 : 0;
        }
        int rule_now = rule;
        if (rule_now == LPS_RULE_DANTZIG && !(stuck_now < g.m_total + (long long)g.n)) rule_now = LPS_RULE_BLAND;
        int c = INT_MAX;
        if (rule_now == LPS_RULE_DANTZIG) {
            ArgMinF64 best{0.0, INT_MAX};
            for (int k = 0; k < C; ++k) best = argmin_pick(best, s_prc[k]);
            c = best.idx;
        } else {
            for (int k = 0; k < C; ++k) c = min(c, s_bld[k]);
        }
        if (c == INT_MAX) { status = LPS_OPTIMAL; break; }
        const int cta_c = c / cpc, c_loc = c - cta_c * cpc;

        // ---- B: column c through the block's earlier pivots -> my rows' multipliers; ratio test ----
        double x_first = 0.0;                                     // (issued before the block waits for the remote cells)
        if (tid < nrows) x_first = g.T[(size_t)(row_lo + tid) * g.ld + c];
        if (tid < l) s_vc[tid] = *remote(sV + (size_t)tid * cpc + c_loc, cta_c);
        const double uobj = *remote(sO + (size_t)cur * cpc + c_loc, cta_c);
        __syncthreads();
        ArgMinF64 best{0.0, INT_MAX};
        for (int il = tid; il < nrows; il += nt) {
            const int i = row_lo + il;
            double x = (il == tid) ? x_first : g.T[(size_t)i * g.ld + c];
            for (int lp = 0; lp < l; ++lp)
                x = (s_R[lp] == (long long)i) ? s_vc[lp] : blk_apply(x, sU[(size_t)lp * rpc + il], s_vc[lp]);
            sU[(size_t)l * rpc + il] = x;
            double* rec = P.U + (size_t)i * kBlkRec;
            __stcg(&rec[l], x);
            uint2 mk = (l == 0) ? make_uint2(0u, 0u) : sM[il];
            if (x == 0.0) mk.x |= 1u << l;
            sM[il] = mk;
            if (l == 0 || x == 0.0) __stcg(reinterpret_cast<uint2*>(rec + kBlkMax), mk);
            if (x > 0.0) {
                const double ratio = Bc[il] / x;
                if (best.idx == INT_MAX || ratio < best.key) { best.key = ratio; best.idx = i; }
            }
        }
        if (me == 0 && tid == 0) {
            double* rec = P.U + (size_t)g.m * kBlkRec;
            __stcg(&rec[l], uobj);
            uint2 mk = (l == 0) ? make_uint2(0u, 0u) : __ldcg(reinterpret_cast<const uint2*>(rec + kBlkMax));
            if (uobj == 0.0) mk.x |= 1u << l;
            if (l == 0 || uobj == 0.0) __stcg(reinterpret_cast<uint2*>(rec + kBlkMax), mk);
        }
        best = block_argmin(best, s_arg);
        if (tid < C) *remote(&s_rat[me], tid) = best;
        cl.sync();
        best = ArgMinF64{0.0, INT_MAX};
        for (int k = 0; k < C; ++k) best = argmin_pick(best, s_rat[k]);
        if (best.idx == INT_MAX) { status = LPS_UNBOUNDED; break; }
        const int r = best.idx;                                 // leaving row (global = local: one rank)
        const int cta_r = r / rpc, r_loc = r - cta_r * rpc;

        // ---- C: row r through the earlier pivots, normalised -> my columns' cells; objective row; right-hand side ----
        double2 y_first = make_double2(0.0, 0.0);
        if (2 * tid < ncols) y_first = *reinterpret_cast<const double2*>(g.T + (size_t)r * g.ld + col_lo + 2 * tid);
        if (tid < l) s_ur[tid] = *remote(sU + (size_t)tid * rpc + r_loc, cta_r);
        const double pv = *remote(sU + (size_t)l * rpc + r_loc, cta_r);
        const double br = *remote(sB + (size_t)cur * rpc + r_loc, cta_r);
        __syncthreads();
        const double inv = 1.0 / pv;                            // rowDiv: ONE / d, then rowMult (tableau.py:262-267)
        const double vn = (inv != 1.0) ? __dmul_rn(br, inv) : br;
        ArgMinF64 pb{0.0, INT_MAX};
        int first_neg = INT_MAX;
        {
            const double* src = g.T + (size_t)r * g.ld + col_lo;
            double* Vg = P.V + ((vset + 0) * kBlkMax + l) * g.ld + col_lo;
            for (int jl = 2 * tid; jl < ncols; jl += 2 * nt) {
                double2 y = (jl == 2 * tid) ? y_first : *reinterpret_cast<const double2*>(src + jl);
                for (int lp = 0; lp < l; ++lp) {
                    const double2 v = *reinterpret_cast<const double2*>(sV + (size_t)lp * cpc + jl);
                    if (s_R[lp] == (long long)r) y = v;
                    else { y.x = blk_apply(y.x, s_ur[lp], v.x); y.y = blk_apply(y.y, s_ur[lp], v.y); }
                }
                if (inv != 1.0) { y.x = __dmul_rn(y.x, inv); y.y = __dmul_rn(y.y, inv); }
                *reinterpret_cast<double2*>(sV + (size_t)l * cpc + jl) = y;
                stcg2(Vg + jl, y);
                double2 o = *reinterpret_cast<const double2*>(Oc + jl);
                o.x = blk_apply(o.x, uobj, y.x);
                o.y = blk_apply(o.y, uobj, y.y);
                *reinterpret_cast<double2*>(On + jl) = o;
                price_consider(o.x, col_lo + jl, g.n, pb, first_neg);
                price_consider(o.y, col_lo + jl + 1, g.n, pb, first_neg);
            }
        }
        for (int il = tid; il < nrows; il += nt) {
            const int i = row_lo + il;
            double b;
            if (i == r) {
                b = vn;
                uint2 mk = sM[il];
                mk.y |= 1u << l;
                sM[il] = mk;
                __stcg(reinterpret_cast<uint2*>(P.U + (size_t)i * kBlkRec + kBlkMax), mk);
            } else b = blk_apply(Bc[il], sU[(size_t)l * rpc + il], vn);
            Bn[il] = b;
        }
        pb = block_argmin(pb, s_arg);
        first_neg = block_min_int(first_neg, s_int);
        if (tid < C) { *remote(&s_prc[me], tid) = pb; *remote(&s_bld[me], tid) = first_neg; }
        if (tid == 0) {
            s_R[l] = r;
            if (me == 0) {   // commit (Simplex._pivot simplex.py:192-197)
                P.R[l] = r;
                P.C[l] = c;
                P.W[l] = 0;
                if (a.trace && npiv < a.trace_cap) {
                    a.trace[3 * npiv] = r;
                    a.trace[3 * npiv + 1] = c;
                    a.trace[3 * npiv + 2] = rule_now;
                }
                if (a.marks) {
                    const int old = a.basis[r];
                    if (old >= 0 && old < g.n) a.marks[old] = 0;
                    a.marks[c] = 1;
                }
                a.basis[r] = c;
            }
        }
        npiv += 1;
        rule = rule_now;
        stuck = stuck_now;
        pending = (rule_now == LPS_RULE_DANTZIG);
        last_r = r;
        last_c = c;
        cur ^= 1;
        cl.sync();
    }
    if (status == LPS_OK && maxp >= 0 && npiv >= maxp) status = LPS_MAX_PIVOTS;

    // panel state out
    for (int jl = tid; jl < ncols; jl += nt) P.O[cur][col_lo + jl] = sO[(size_t)cur * cpc + jl];
    for (int il = tid; il < nrows; il += nt) P.B[cur][row_lo + il] = sB[(size_t)cur * rpc + il];
    cl.sync();                                                  // nobody reads a CTA's shared memory after it has left
    if (me == 0 && tid == 0) {
        P.hdr[0] = l;
        P.hdr[1] = cur;
        P.hdr[3] = set;
        if (l > 0) P.hdr[2] = P.hdr[2] + 1;
        st->status = status;
        st->rule = rule;
        st->stuck = stuck;
        st->npiv = npiv;
        st->pending_zcheck = pending;
        st->last_r = last_r;
        st->last_c = last_c;
    }
}

// ------------------------------------------------------------------------------------------------
// update: the block's pivots applied to every cell of the local rows in one pass
// ------------------------------------------------------------------------------------------------
// Persistent, warp-specialised, one CTA per SM.  Work item = (512-column chunk, stage of 8 rows); every CTA owns a
// contiguous range of the chunk-major item list.  One producer thread moves the items through a shared-memory ring
// with TMA bulk copies: 8 row segments of the grid + the 8 rows' panel records in, 8 row segments out; 256 consumer
// threads (one double2 column each, the chunk's 16 normalised pivot-row cells in registers) apply the block's pivots
// to the stage in place.  HBM reads and writes never wait for the arithmetic and the fp64 pipe never waits for HBM
// as long as the ring is deep enough: the grid is read once and written once per block.
struct BlkUpdateArgs {
    F64Grid g;
    BlkPanel p;
};

struct BlkRing {                                  // dynamic shared memory of k_blk_update
    double t[kBlkStages][kBlkRU][kBlkCW];         // ring: stage rows of the grid
    double u[kBlkStages][kBlkRU * kBlkRec];       // ring: the rows' panel records
    uint64_t full[kBlkStages];                    // producer -> consumers: stage landed (TMA transaction bytes)
    uint64_t done[kBlkStages];                    // consumers -> producer: stage updated in place (one arrival per warp)
};

__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_spin(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {}
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// shared -> global bulk copy, tracked by the issuing thread's bulk async-group
__device__ __forceinline__ void tma_store_1d(void* gmem_dst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N>
__device__ __forceinline__ void tma_wait_all() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__global__ void __launch_bounds__(kBlkConsumers + 32, 1) k_blk_update(BlkUpdateArgs a) {
    extern __shared__ __align__(128) unsigned char blk_smem[];
    BlkRing& S = *reinterpret_cast<BlkRing*>(blk_smem);
    const BlkPanel& P = a.p;
    const F64Grid& g = a.g;
    const int kk = P.hdr[0];
    if (kk <= 0) return;
    const int tid = (int)threadIdx.x;
    const int mrows = g.m + 1;
    const long long nst = (mrows + kBlkRU - 1) / kBlkRU;                    // stages per chunk
    const long long nchunk = ((long long)g.ld + kBlkCW - 1) / kBlkCW;
    const long long W = nst * nchunk;
    const long long it0 = W * blockIdx.x / gridDim.x, it1 = W * (blockIdx.x + 1) / gridDim.x;
    if (it0 >= it1) return;

    if (tid == 0) {
        for (int s = 0; s < kBlkStages; ++s) { mbar_init(&S.full[s], 1); mbar_init(&S.done[s], kBlkConsumers / 32); }
        mbar_fence_init();
    }
    __syncthreads();

    if (tid >= kBlkConsumers) {
        // ---------------- producer: one thread ----------------
        if (tid != kBlkConsumers) return;
        auto load_item = [&](long long it, long long q) {
            const long long chunk = it / nst, st = it - chunk * nst;
            const int slot = (int)(q % kBlkStages);
            const int col0 = (int)(chunk * kBlkCW);
            const int ncols = min(kBlkCW, (int)g.ld - col0);
            const int r0 = (int)(st * kBlkRU);
            const int nr = min(kBlkRU, mrows - r0);
            mbar_expect_tx(&S.full[slot], (uint32_t)nr * (uint32_t)ncols * 8u + (uint32_t)(kBlkRU * kBlkRec * 8));
            tma_load_1d(&S.u[slot][0], P.U + (size_t)r0 * kBlkRec, (uint32_t)(kBlkRU * kBlkRec * 8), &S.full[slot]);
            for (int r = 0; r < nr; ++r)
                tma_load_1d(&S.t[slot][r][0], g.T + (size_t)(r0 + r) * g.ld + col0, (uint32_t)ncols * 8u, &S.full[slot]);
        };
        auto store_item = [&](long long it, long long q) {
            const long long chunk = it / nst, st = it - chunk * nst;
            const int slot = (int)(q % kBlkStages);
            const int col0 = (int)(chunk * kBlkCW);
            const int ncols = min(kBlkCW, (int)g.ld - col0);
            const int r0 = (int)(st * kBlkRU);
            const int nr = min(kBlkRU, mrows - r0);
            for (int r = 0; r < nr; ++r)
                tma_store_1d(g.T + (size_t)(r0 + r) * g.ld + col0, &S.t[slot][r][0], (uint32_t)ncols * 8u);
            tma_commit();
        };
        const long long n = it1 - it0;
        for (long long q = 0; q < n && q < kBlkStages; ++q) load_item(it0 + q, q);
        for (long long q = 0; q < n; ++q) {
            const int slot = (int)(q % kBlkStages);
            mbar_spin(&S.done[slot], (uint32_t)((q / kBlkStages) & 1));
            store_item(it0 + q, q);
            // refill the slot of the PREVIOUS item: its store (one bulk group older) has finished reading shared memory
            if (q >= 1 && q - 1 + kBlkStages < n) {
                tma_wait_read<1>();
                load_item(it0 + q - 1 + kBlkStages, q - 1 + kBlkStages);
            }
        }
        tma_wait_all<0>();
        return;
    }

    // ---------------- consumers: 8 warps, thread = one double2 column of the chunk ----------------
    const int jl = 2 * tid;
    double2 v[kBlkMax];                       // this column's cells of the block's normalised pivot rows
    long long chunk_have = -1;
    const long long n = it1 - it0;
    const unsigned kmask = (1u << kk) - 1u;   // kk <= 16
    for (long long q = 0; q < n; ++q) {
        const long long it = it0 + q;
        const long long chunk = it / nst;
        const int slot = (int)(q % kBlkStages);
        if (chunk != chunk_have) {
            const int col = (int)(chunk * kBlkCW) + jl;
#pragma unroll
            for (int l = 0; l < kBlkMax; ++l)
                v[l] = (l < kk && col < (int)g.ld)
                           ? ldcg2(P.V + (((size_t)P.hdr[3] * P.nr + P.W[l]) * kBlkMax + l) * g.ld + col)
                           : make_double2(0.0, 0.0);
            chunk_have = chunk;
        }
        mbar_spin(&S.full[slot], (uint32_t)((q / kBlkStages) & 1));
        const double* rec = &S.u[slot][0];
        unsigned any = 0;
#pragma unroll
        for (int r = 0; r < kBlkRU; ++r) {
            const uint2 mk = *reinterpret_cast<const uint2*>(rec + r * kBlkRec + kBlkMax);
            any |= mk.x | mk.y;
        }
        double2 t[kBlkRU];
#pragma unroll
        for (int r = 0; r < kBlkRU; ++r) t[r] = *reinterpret_cast<const double2*>(&S.t[slot][r][jl]);
        if (kk == kBlkMax && (any & kmask) == 0) {
            // the common case: a full block, and every pivot of it touches every row of the stage.  Branch free; the
            // multipliers of the next two pivots are fetched while the current two are applied.
            double2 ua[kBlkRU], ub[kBlkRU];
#pragma unroll
            for (int r = 0; r < kBlkRU; ++r) ua[r] = *reinterpret_cast<const double2*>(rec + r * kBlkRec);
#pragma unroll
            for (int l = 0; l < kBlkMax; l += 2) {
                if (l + 2 < kBlkMax) {
#pragma unroll
                    for (int r = 0; r < kBlkRU; ++r) ub[r] = *reinterpret_cast<const double2*>(rec + r * kBlkRec + l + 2);
                }
#pragma unroll
                for (int r = 0; r < kBlkRU; ++r) {
                    t[r].x = __dadd_rn(t[r].x, __dmul_rn(-ua[r].x, v[l].x));
                    t[r].y = __dadd_rn(t[r].y, __dmul_rn(-ua[r].x, v[l].y));
                }
#pragma unroll
                for (int r = 0; r < kBlkRU; ++r) {
                    t[r].x = __dadd_rn(t[r].x, __dmul_rn(-ua[r].y, v[l + 1].x));
                    t[r].y = __dadd_rn(t[r].y, __dmul_rn(-ua[r].y, v[l + 1].y));
                }
#pragma unroll
                for (int r = 0; r < kBlkRU; ++r) ua[r] = ub[r];
            }
        } else if ((any & kmask) == 0) {
            // a short block (the end of a solve or of a pivot prefix): the uniform test on the block length only
#pragma unroll
            for (int l = 0; l < kBlkMax; ++l) {
                if (l < kk) {
#pragma unroll
                    for (int r = 0; r < kBlkRU; ++r) {
                        const double u = rec[r * kBlkRec + l];
                        t[r].x = __dadd_rn(t[r].x, __dmul_rn(-u, v[l].x));
                        t[r].y = __dadd_rn(t[r].y, __dmul_rn(-u, v[l].y));
                    }
                }
            }
        } else {
            // rows with a zero multiplier (left alone, tableau.py:272) / pivot rows (replaced by the normalised row): the
            // same arithmetic, its result selected per row and pivot from the masks -- no divergent control flow
            unsigned mz[kBlkRU], ms[kBlkRU];
#pragma unroll
            for (int r = 0; r < kBlkRU; ++r) {
                const uint2 mk = *reinterpret_cast<const uint2*>(rec + r * kBlkRec + kBlkMax);
                mz[r] = mk.x;
                ms[r] = mk.y;
            }
#pragma unroll
            for (int l = 0; l < kBlkMax; ++l) {
                if (l < kk) {
#pragma unroll
                    for (int r = 0; r < kBlkRU; ++r) {
                        const double u = rec[r * kBlkRec + l];
                        const double nx = __dadd_rn(t[r].x, __dmul_rn(-u, v[l].x));
                        const double ny = __dadd_rn(t[r].y, __dmul_rn(-u, v[l].y));
                        const bool spec = (ms[r] >> l) & 1u, zero = (mz[r] >> l) & 1u;
                        t[r].x = spec ? v[l].x : (zero ? t[r].x : nx);
                        t[r].y = spec ? v[l].y : (zero ? t[r].y : ny);
                    }
                }
            }
        }
#pragma unroll
        for (int r = 0; r < kBlkRU; ++r) *reinterpret_cast<double2*>(&S.t[slot][r][jl]) = t[r];
        fence_proxy_async();                  // the TMA store (async proxy) must see these writes
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(&S.done[slot]);
    }
}

}  // namespace lps
