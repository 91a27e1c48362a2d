/*
 * lps_b200.h -- C-ABI of the B200-native simplex pivot engine.
 *
 * The reference (tkoz0/linear-program-solver, `lpsol`) is pure Python and has no
 * FFI: the drop-in boundary is the class surface between `Simplex` (caller) and
 * `Tableau` (callee).  Each entry point below names the reference interface it
 * replaces (file:line under /root/reference/lpsol).  The host package
 * `linear_program_solver_b200` binds these with ctypes; INTEGRATION.md shows the
 * stub a reference maintainer would add.
 *
 * Conventions
 *   - opaque handle, int status return, no exceptions across the boundary;
 *   - the caller owns host buffers; device buffers are either owned by the
 *     library or lent by the caller (a torch tensor's data_ptr) and must then
 *     outlive the handle;
 *   - one handle = one device + one CUDA stream; calls on a handle are not
 *     thread-safe; every call is asynchronous on the stream unless it returns
 *     data to the host;
 *   - device layout of a tableau ("grid"): row-major (m+1) x ld, rows 0..m-1
 *     the constraints, row m the objective row (reduced costs), columns 0..n-1
 *     the variables, column n the right-hand side; cell (m,n) is the
 *     reference's `_z` field, i.e. MINUS the objective value
 *     (tableau.py:46,82-84,128-130).  ld = lps_grid_ld(n).
 */
#ifndef LPS_B200_H
#define LPS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes -------------------------------------------------------- */
#define LPS_OK            0   /* call succeeded / solver still running                     */
#define LPS_OPTIMAL       1   /* simplex.py:269 'optimal'                                   */
#define LPS_UNBOUNDED     2   /* simplex.py:281 'unbounded' (solve(): AssertionError :125)  */
#define LPS_ZERO_PIVOT    3   /* tableau.py:300-301 ZeroDivisionError('zero pivot r,c')     */
#define LPS_MAX_PIVOTS    4   /* stopped after the requested pivot prefix (not in reference) */
#define LPS_BAD_PRIME     5   /* exact mode: next pivot is 0 modulo a prime; re-encode      */
#define LPS_OBJ_INCREASED  6   /* simplex.py:133 'objective value increased (internal error)' */
#define LPS_INFEASIBLE       7   /* simplex.py:64-67 ValueError('infeasible problem, artificial opt = z') */
#define LPS_REDUNDANT_ROW    8   /* simplex.py:84-93 dependent row dropped: the reference then fails with IndexError */
#define LPS_NEED_ARTIFICIALS 9   /* batched solve: more artificial columns needed than the batch reserved */
#define LPS_NEED_PRIMES     10   /* exact mode, internal: the pending pivot needs a larger working prime set */
#define LPS_E_INVALID    -1   /* bad argument                                              */
#define LPS_E_CUDA       -2   /* CUDA runtime error, see lps_last_error()                  */
#define LPS_E_NOMEM      -3
#define LPS_E_STATE      -4   /* call not valid in the handle's current state              */
#define LPS_E_PEER       -5   /* a peer GPU did not answer within the exchange time-out    */

/* pricing rules */
#define LPS_RULE_DANTZIG  0   /* Simplex.findPivotStandard  simplex.py:251-284 */
#define LPS_RULE_BLAND    1   /* Simplex.findPivotMinIndex  simplex.py:218-249 */

typedef struct lps_handle lps_handle;

/* ---- library ------------------------------------------------------------- */
int         lps_version(void);
const char *lps_last_error(void);              /* thread-local, valid until the next failing call */
int         lps_device_count(int *count);
size_t      lps_grid_ld(int n);                /* leading dimension (elements) of a grid with n variables */

/* ---- fp64 tableau: storage (replaces Tableau.__init__ tableau.py:36-52) --- */
/*
 * m, n        local constraint rows and variables of this handle's grid.
 * dev_grid    NULL: the library allocates (m+1)*ld doubles; else a device
 *             pointer to that many doubles lent by the caller.
 * stream      a cudaStream_t (as void*) to enqueue on, or NULL for a private one.
 */
int lps_create_f64(int m, int n, int device, void *dev_grid, void *stream, lps_handle **out);
int lps_destroy(lps_handle *h);
int lps_sync(lps_handle *h);

/* Row-sharding: this handle holds constraint rows
 * [row0, row0+m) of an m_total-row tableau on rank `rank` of `nranks`, plus its
 * own replica of the objective row.  cand_send / cand_all are device buffers of
 * lps_cand_bytes(n) and nranks*lps_cand_bytes(n) bytes lent by the caller (the
 * all-gather send and receive buffers).  For nranks == 1 this call is optional. */
size_t lps_cand_bytes(int n);
int lps_set_shard(lps_handle *h, long long row0, long long m_total, int rank, int nranks,
                  void *cand_send, void *cand_all);

/* Host <-> device transfer of the whole grid (setA/setB/setC/setZ and the
 * getters, tableau.py:70-180).  `host` is row-major (m+1) x ld_host doubles in
 * the device layout described above. */
int lps_upload_f64(lps_handle *h, const double *host, size_t ld_host);
int lps_download_f64(lps_handle *h, double *host, size_t ld_host);
/* getCj/getBi/getAij/getZ (tableau.py:82-108): one row / one column / one cell. */
int lps_read_row_f64(lps_handle *h, int i, double *host);   /* n+1 doubles */
int lps_read_col_f64(lps_handle *h, int j, double *host);   /* m+1 doubles */
void *lps_grid_ptr(lps_handle *h);                          /* device pointer of the grid */
/* Synthetic input (bench.py, tests): fill this handle's rows of the seeded dense
 * family of SURVEY.md section 8(d) (n = 2 m_total, A in [lo,hi], slack identity,
 * b in [hi*m/4, hi*m/2], costs -[lo,hi]) in place with a counter-based hash, so
 * every shard can be generated on its own GPU.  Same values as
 * problems.hashed_dense_grid() and the oracle's orc_fill_dense_f64(). */
int lps_fill_dense_f64(lps_handle *h, unsigned long long seed, int lo, int hi);
/* A handle created on a lent grid (dev_grid != NULL) moves to another device grid of the same shape; the handle's
 * small buffers, stream and state are kept (the host pools handles of large tableaux without pooling the grids). */
int lps_rebind_grid_f64(lps_handle *h, void *dev_grid);

/* ---- the hot path --------------------------------------------------------- */
/* Tableau.pivot(r,c) tableau.py:295-308 (with rowDiv :262, rowAddToObj :282,
 * rowSub :278 fused in one rank-1 update).  LPS_ZERO_PIVOT if a_rc == 0. */
int lps_pivot(lps_handle *h, int r, int c);

/* Row operations outside a pivot: rowMult tableau.py:254, rowAdd :269 (rd == m
 * addresses the objective row, i.e. rowAddToObj :282). */
int lps_row_scale_f64(lps_handle *h, int r, double mult);
int lps_row_axpy_f64(lps_handle *h, int rd, int rs, double mult);

/* Leaving phase 1 on the device: Simplex._find_bfs simplex.py:95-105 when no constraint row is dropped
 * (private-field surgery `_c = _c[:n]`, `_a[i] = _a[i][:n]` :95-100, then setZ / setC :101-102 and
 * rowSubFromObj(i, c_{B_i}) for every row :104-105).  The trailing columns n_new..n-1 (the artificial
 * variables) are cut -- the right-hand side moves to column n_new, the handle's n becomes n_new, ld and
 * all buffers keep their sizes -- and the objective row becomes `cost` (n_new reduced costs followed by
 * the _z cell) priced out against `basis` (m columns below n_new) in the reference's row order and
 * rounding.  No host round trip of the tableau. */
int lps_phase2_f64(lps_handle *h, int n_new, const double *cost, const int *basis);

/* Simplex.findPivotStandard / findPivotMinIndex with do_pivot=False
 * (simplex.py:251-284, :218-249).  Returns LPS_OK with (*r,*c) set, or
 * LPS_OPTIMAL / LPS_UNBOUNDED. */
int lps_find_pivot(lps_handle *h, int rule, int *r, int *c);
/* Whole-tableau pivot rules, fp64 mode (exact mode evaluates them on host snapshots): the ratio test of every
 * column in one pass.  lps_find_pivot_max_increase = Simplex.findPivotMaxIncrease simplex.py:286-328 (largest
 * objective decrease -c_j * ratio, first column among equals; LPS_OPTIMAL / LPS_UNBOUNDED like the literals).
 * lps_find_pivot_all = Simplex.findPivotAll simplex.py:330-360: every (row, column) of minimum ratio, column by
 * column, rows ascending; *count = how many exist, at most cap are written. */
int lps_find_pivot_max_increase(lps_handle *h, int *r, int *c);
int lps_find_pivot_all(lps_handle *h, int *rows, int *cols, long long cap, long long *count);

/* Tableau.isCanonical(bcols) tableau.py:466-496.  bcols (m ints, may be NULL)
 * receives the lowest basic column per row or -1; *canonical is 0/1. */
int lps_is_canonical(lps_handle *h, int *bcols, int *canonical);
/* Tableau.isOptimal tableau.py:500-502 */
int lps_is_optimal(lps_handle *h, int *optimal);
/* Tableau.isInfeasible tableau.py:510-514: some row with b_i > 0 has no positive entry (either mode; exact: the
 * rows are reconstructed on the device in batches and only their signs are looked at). */
int lps_is_infeasible(lps_handle *h, int *infeasible);

/* Simplex.solve() simplex.py:110-148, the whole loop on the device: Dantzig
 * pricing while the stuck counter is below m+n (the counter compares with the
 * objective at entry, simplex.py:118/:134), then Bland.
 *   max_pivots  < 0: run to the end; else stop after that many pivots
 *   basis       in/out, m ints (Simplex._bfs): basis[r] = c after each pivot
 *   marks       in/out, n ints 0/1 or NULL (Tableau._cm): Simplex._pivot's
 *               bookkeeping simplex.py:195-197 (unmark leaving, mark entering)
 *   trace       NULL or 3*trace_cap ints: (row, column, rule) per pivot
 *   npiv        number of pivots done by this call
 * Returns LPS_OPTIMAL, LPS_UNBOUNDED or LPS_MAX_PIVOTS. */
int lps_solve(lps_handle *h, long long max_pivots, int *basis, int *marks, int *trace,
              long long trace_cap, long long *npiv);

/* Sharded stepping (one pivot = select, all-gather of candidates by the
 * caller, update).  lps_solve_begin arms the loop state (stuck counter,
 * objective at entry); lps_step_select prices on the replicated objective
 * row, ratio-tests the local rows and fills cand_send; lps_step_update picks
 * the global leaving row out of cand_all (lowest ratio, then lowest global row,
 * simplex.py:277), applies the rank-1 update to the local rows and commits the
 * pivot.  lps_solve_poll copies the state back: status, pivots so far. */
int lps_solve_begin(lps_handle *h, long long max_pivots, const int *basis, const int *marks,
                    long long trace_cap);
int lps_step_select(lps_handle *h);
/* Peer-store exchange (no collective call): after lps_p2p_attach, lps_step_select itself writes this rank's 64-byte
 * candidate header straight into every peer GPU's exchange block over NVLink, waits for the peers' headers, and the
 * rank that owns the winning row normalises it and stores it into every peer; the launch ends on a rank once that
 * rank holds the row.  It replaces the caller's all-gather between lps_step_select and lps_step_update: the argmin
 * all-reduce + pivot-row broadcast of SURVEY.md section 8e fused into the select kernel.  lps_p2p_export returns the
 * 64-byte CUDA IPC handle of this handle's exchange block; the caller gathers the handles of all ranks (any
 * transport) and passes them, in rank order, to lps_p2p_attach on every rank. */
/* lps_p2p_prepare sizes the exchange block for nranks ranks (in the blocked loop every rank stores its candidate rows
 * into every rank's block); call it before lps_p2p_export / lps_p2p_block -- it may move the block. */
int lps_p2p_prepare(lps_handle *h, int nranks);
int lps_p2p_export(lps_handle *h, unsigned char *ipc_handle64);
int lps_p2p_attach(lps_handle *h, long long row0, long long m_total, int rank, int nranks,
                   const unsigned char *ipc_handles);
/* The same exchange for ranks living in one process (one thread driving several handles, on one or several
 * GPUs): the blocks are passed as device pointers (lps_p2p_block of every rank, in rank order). */
int lps_p2p_block(lps_handle *h, void **block);
int lps_p2p_attach_local(lps_handle *h, long long row0, long long m_total, int rank, int nranks,
                         void *const *blocks);
int lps_step_update(lps_handle *h);
/* Pivots in blocks (the loop lps_solve runs in fp64 mode; csrc/lps_f64_blk.cuh).  One lps_step_block = a look-ahead
 * panel that DECIDES up to lps_block_size(h) consecutive pivots of Simplex.solve (simplex.py:110-148: pricing, ratio
 * test, stuck counter, rule switch, Simplex._pivot's bookkeeping :192-197) from O(m+n) numbers -- the objective row
 * and right-hand side kept current, the entering column and leaving row gathered from the stale grid and corrected
 * with the block's earlier pivots -- followed by ONE pass over the grid that applies all of them to every cell in the
 * reference's order and rounding (tableau.py:295-308).  Results are bit-identical to pivot-by-pivot updates; the grid
 * is read and written once per block instead of once per pivot.  On a row-sharded handle (after lps_p2p_attach) the
 * panel does the per-pivot argmin all-reduce and pivot-row broadcast with peer stores, like lps_step_select.
 * lps_set_block_size: 1..16 pivots per block, 0 = lps_solve runs the rank-1 loop (default: 16 wherever the cluster panel fits -- up to about 7000 rows at n = 2m -- and from 1.5 M cells on, else 0; or the
 * environment variable LPS_F64_BLOCK). */
int lps_step_block(lps_handle *h);
int lps_block_size(lps_handle *h);
int lps_set_block_size(lps_handle *h, int k);
/* Launch shape of the panel (tuning / tests): 1 = one CTA of 1024 threads, >= 2 = that many CTAs of 256 threads
 * (at most one per SM), -8 / -16 = ONE thread-block cluster of 8 / 16 CTAs with the whole panel in distributed shared
 * memory (medium tableaux, one rank; LPS_E_INVALID when it does not fit).  The default follows the size of the tableau:
 * the cluster panel wherever it fits. */
int lps_set_panel_ctas(lps_handle *h, int ctas);
int lps_solve_poll(lps_handle *h, int *status, long long *npiv);
int lps_solve_end(lps_handle *h, int *basis, int *marks, int *trace, long long trace_cap,
                  long long *npiv);


/* ---- exact mode: fraction-free integer tableau as residues modulo 31-bit primes ----
 * The reference's arithmetic is fractions.Fraction (tableau.py:2,130-158).  The device holds an
 * integer grid N and one common denominator D as residues modulo `nprimes` primes (one prime per
 * grid slice), such that the reference's entry (i,j) is N_ij / (D * sigma_i); sigma_i is a positive
 * per-row scale (1 for integer inputs) that a row keeps until it is first pivoted.  The caller
 * sizes the prime set from a Hadamard bound on the minors of the uploaded integer matrix.
 * lps_pivot / lps_find_pivot / lps_is_canonical / lps_is_optimal / lps_solve work on either mode.
 * LPS_BAD_PRIME (a prime divides the next pivot element; nothing was changed) asks the caller to
 * re-encode with a different prime set. */
size_t lps_exact_ld(int n);
int lps_create_exact(int m, int n, int nprimes, const uint32_t *primes, int device, void *stream,
                     lps_handle **out);
/* Upload: entry (i,j) of the (m+1) x ld_host host grid is a signed integer of `nlimbs`
 * little-endian 32-bit limbs (magnitude) at limbs[(i*ld_host+j)*nlimbs] with sign neg[i*ld_host+j];
 * sigma_limbs holds the m+1 row scales the same way.  Resets D to 1 (setA/setB/setC/setZ,
 * tableau.py:128-166). */
int lps_upload_exact(lps_handle *h, const uint32_t *limbs, int nlimbs, const unsigned char *neg,
                     size_t ld_host, const uint32_t *sigma_limbs, int sigma_nlimbs);
/* Read-out (getters, tableau.py:82-108) by mixed-radix reconstruction on the device.
 * what 0: row `index` (n+1 numbers), 1: column `index` (m+1 numbers), 2: the denominator D.
 * sign_out[e] in {-1,0,1}; digits_out[k*count+e] is digit k of |x_e| in the mixed radix
 * (p_0, p_1, ...): |x| = d_0 + d_1 p_0 + d_2 p_0 p_1 + ...   */
int lps_read_exact(lps_handle *h, int what, int index, signed char *sign_out, uint32_t *digits_out);
int lps_read_alive(lps_handle *h, int *alive);   /* m+1 flags: row still carries its sigma_i */
/* Scattered cells (rows[e], cols[e]), e < count <= max(n, m) + 1 (getAij, tableau.py:98-100, for a list of cells):
 * same outputs as lps_read_exact. */
int lps_read_exact_cells(lps_handle *h, int count, const int *rows, const int *cols, signed char *sign_out,
                         uint32_t *digits_out);
/* Simplex.findPivotAll simplex.py:330-360 and the per-column half of Simplex.findPivotMaxIncrease :286-328 in exact
 * mode: the ratio test of EVERY column on the device (a batch of columns per launch, one block per column, the
 * floating filter and exact tie-break of the solve loop).  minrow[j], j < n: lowest row of minimum ratio
 * b_i / a_ij over a_ij > 0 (strict "<" of simplex.py:308); -1: no positive entry; -2: column not tested
 * (only_negative != 0 and c_j >= 0).  With tie_off (n+1 offsets) and ties: all rows of minimum ratio, ascending,
 * column after column; at most ties_cap entries are written, tie_off[n] is their number. */
int lps_exact_col_ratios(lps_handle *h, int only_negative, int *minrow, long long *tie_off, int *ties,
                         long long ties_cap);
int lps_exact_info(lps_handle *h, int *nprimes, int *bad_prime, size_t *residue_bytes);
/* Working prime set.  lps_create_exact allocates the primes a static Hadamard bound of the uploaded matrix asks for;
 * the numbers of a tableau are minors over the CURRENT basis columns only, usually far smaller.  After
 * lps_exact_set_weights (right after lps_upload_exact; wcol[j] >= 256 log2 of the 2-norm of column j of the uploaded
 * integer matrix over all m+1 rows, j = 0..n; guard_bits covers the row scales) every kernel runs on a prefix of
 * the primes sized by the product of the current basis columns' norms, and the library extends the prefix on the
 * device (Garner + Horner, no host round trip of the tableau) before a pivot that would not fit.  Results are
 * unchanged -- the residues are those of the same integers.  lps_exact_working reports the prefix length, the
 * allocated primes and the number of extensions so far.  lps_exact_info's nprimes / residue_bytes describe the
 * working set. */
int lps_exact_set_weights(lps_handle *h, const int *wcol, int guard_bits);
/* lps_upload_exact and lps_exact_set_weights in one call: only the residue planes of the working set are allocated
 * and encoded (the handle's other planes appear when an extension needs them). */
int lps_upload_exact_weighted(lps_handle *h, const uint32_t *limbs, int nlimbs, const unsigned char *neg,
                              size_t ld_host, const uint32_t *sigma_limbs, int sigma_nlimbs, const int *wcol,
                              int guard_bits);
int lps_exact_working(lps_handle *h, int *working, int *allocated, int *extensions);
/* Row operations outside a pivot in exact mode (rowMult / rowDiv tableau.py:254-267, rowAdd / rowSub :269-280,
 * rowAddToObj / rowSubFromObj :282-293) on the residues:  N_rd <- a N_rd + b N_rs  (rs < 0: N_rd <- a N_rd; rd == m:
 * the objective row), a and b signed multi-limb integers, sigma the new scale of row rd.  Needs a handle with a
 * working prime set.  The tableau's numbers stop being minors of the uploaded matrix: guard_bits is the caller's new
 * allowance on top of the column-weight bound.  Returns LPS_NEED_PRIMES with nothing modified when all the handle's
 * primes cannot hold it (the caller re-encodes: download, edit, upload); lps_pivot / lps_solve may then report
 * LPS_NEED_PRIMES as well. */
int lps_row_combine_exact(lps_handle *h, int rd, int rs, const uint32_t *a_limbs, int a_nlimbs, int a_neg,
                          const uint32_t *b_limbs, int b_nlimbs, int b_neg, const uint32_t *sigma_limbs,
                          int sigma_nlimbs, int guard_bits);
/* lps_phase2_f64 for the exact mode.  cost: n_new + 1 signed integers C_j = sigma_m c_j (limbs as in
 * lps_upload_exact; C_n: the _z cell), sigma_m the new scale of the objective row (a multiple of every
 * row scale and of the cost denominators); mult: m signed integers G_i = C_{basis[i]} / s_i with s_i
 * the scale row i still carries (lps_read_alive) or 1.  The device computes N_mj = D C_j - sum_i G_i N_ij.
 * From here on the tableau's numbers are minors of [constraint rows as uploaded ; C] over the basis
 * columns of every row: wcol / guard_bits are that matrix's column weights (as for
 * lps_exact_set_weights; ignored on a handle without a working prime set), and the prime set given to
 * lps_create_exact must cover its Hadamard bound as well. */
int lps_phase2_exact(lps_handle *h, int n_new, const uint32_t *cost_limbs, int cost_nlimbs,
                     const unsigned char *cost_neg, const uint32_t *mult_limbs, int mult_nlimbs,
                     const unsigned char *mult_neg, const uint32_t *sigma_m_limbs, int sigma_m_nlimbs,
                     const int *basis, const int *wcol, int guard_bits);
/* Loop state of Simplex.solve (simplex.py:117-137: rule in force, stuck counter, whether the
 * objective has moved since entry, pending objective check).  set == 0 reads it; set != 0 writes it
 * and makes the next lps_solve continue from it (resuming after LPS_BAD_PRIME on a re-encoded
 * tableau without disturbing the reference's rule-switch sequence). */
int lps_exact_loop_state(lps_handle *h, int set, int *rule, long long *stuck, int *moved, int *pending);

/* ---- batched node LPs (MILP branch and bound, BASELINE.json config 5) ---------------------------
 * The reference's LinProg is an empty stub (linprog.py:383-393), so there is no branch-and-bound to
 * replace; what a B&B driver calls per node is `Simplex(tab); tab... solve()` (simplex.py:24-148) on a
 * small integer tableau.  A batch runs that WHOLE sequence -- phase 1 included (simplex.py:36-108) --
 * for `count` independent LPs of one shape in one kernel launch: one LP per thread-block cluster, the
 * exact-mode residue planes resident in (distributed) shared memory, row-sharded over `cluster` CTAs
 * per LP (0 = automatic; 1 = the whole LP in one CTA).
 *   grids      count x (m+1) x (n+1) int32 in the device layout above (integer coefficients)
 *   art_cap    artificial columns reserved per LP (rows without a basic column, simplex.py:54)
 *   primes     an even number (2..16) of distinct odd primes below 2^31 in ASCENDING order whose product
 *              exceeds twice the Hadamard bound of every LP's extended matrix (phase-1 objective included)
 * Per-LP status: LPS_OPTIMAL, LPS_UNBOUNDED, LPS_INFEASIBLE, LPS_REDUNDANT_ROW, LPS_MAX_PIVOTS,
 * LPS_BAD_PRIME (solve that LP alone with another prime set), LPS_NEED_ARTIFICIALS. */
typedef struct lps_batch lps_batch;
size_t lps_batch_smem_bytes(int m, int n, int art_cap, int nprimes, int cluster);
/* Sizing on the device (bookkeeping, no pivot arithmetic): the largest number of artificial columns any LP of
 * the batch needs and log2 of a Hadamard bound on every minor of every LP (phase-1 objective included).
 * `grids` of lps_batch_plan and lps_batch_upload may be host or device memory (unified addressing). */
int lps_batch_plan(int count, int m, int n, const int32_t *grids, int device, int *art_cap, double *bound_bits);
int lps_batch_create(int count, int m, int n, int art_cap, int nprimes, const uint32_t *primes, int cluster,
                     int trace_cap, int device, void *stream, lps_batch **out);
int lps_batch_destroy(lps_batch *b);
int lps_batch_upload(lps_batch *b, const int32_t *grids);          /* host -> device, asynchronous */
void *lps_batch_grids_ptr(lps_batch *b);                           /* device copy of the grids */
int lps_batch_run(lps_batch *b, long long max_pivots);             /* Simplex(tab); solve() per LP, asynchronous */
int lps_batch_sync(lps_batch *b);
/* Number of consecutive Dantzig pivots at the entry objective after which solve() switches to Bland's rule.
 * Negative (default): the reference's m + n (simplex.py:123).  Parity tests lower it to reach loop B. */
int lps_batch_set_stuck_limit(lps_batch *b, long long limit);
/* A batch may be created with FEWER primes than the static bound of lps_batch_plan asks for (half of them holds a
 * 64 x 128 LP in one CTA instead of a cluster of two).  With the capacity check on, every LP tracks the Hadamard
 * product of its current basis columns' norms and ends with LPS_NEED_PRIMES -- before the pivot that might not fit is
 * applied -- when the short set could overflow; the caller re-runs those LPs in a batch with the full set. */
int lps_batch_set_capacity_check(lps_batch *b, int on);
/* Results (any pointer may be NULL): status[count]; npiv[count][4] = phase-1, drive-out, phase-2 pivots and
 * the number of artificial columns; basis[count][m] (Simplex._bfs); sign[count][m+2] / digits[count][P][m+2]:
 * b_0..b_{m-1}, the `_z` cell and the common denominator D as sign + mixed-radix digits (see
 * lps_read_exact), so that b_i = N_i / D; trace[count][trace_cap][3] = (row, column, rule) per pivot,
 * rule 2 = phase-1 drive-out pivot. */
int lps_batch_download(lps_batch *b, int *status, int *npiv, int *basis, signed char *sign, uint32_t *digits,
                       int *trace);
/* Profiling aid: cycles per kernel phase and LP (out[count][16]); out == NULL arms the collection. */
int lps_batch_timing(lps_batch *b, long long *out);
int lps_batch_info(lps_batch *b, int *cluster, int *rows_per_cta, size_t *smem_bytes, int *resident_clusters,
                   double *last_run_ms, long long *launches);

/* Timing aid for bench.py: CUDA-event time (ms) and launch count of the
 * rank-1 update kernels enqueued since the last reset, measured on the
 * handle's own stream. */
int lps_prof_reset(lps_handle *h);
int lps_prof_read(lps_handle *h, double *update_ms, long long *update_launches, long long *all_launches);

#ifdef __cplusplus
}
#endif
#endif /* LPS_B200_H */
