/*
 * CPU oracle, C restatement of the fp64 simplex pivot path -- TEST INFRASTRUCTURE ONLY.
 *
 * Restates, on one dense row-major (m+1) x ld array of doubles (rows 0..m-1 the
 * constraints, row m the objective row, column n the right-hand side, cell
 * (m,n) the reference's `_z` = minus the objective), what the reference does
 * when its arithmetic type is float:
 *
 *   orc_pivot_f64   lpsol/tableau.py:295-308  (rowDiv :262, rowMult :254,
 *                   rowAddToObj :282, rowSub/rowAdd :269-280)
 *   orc_price_f64   lpsol/simplex.py:262-267 (Dantzig) and :228-232 (Bland)
 *   orc_ratio_f64   lpsol/simplex.py:272-279 (same at :237-244)
 *   orc_solve_f64   lpsol/simplex.py:110-148
 *
 * Every element update is a rounded product followed by a rounded sum (build
 * with -ffp-contract=off), in the reference's order, so results are
 * bit-identical to oracle/lpsol_oracle.py with num=float, which tests pin to
 * the float-rebound reference run (tests/golden/dense_float.json).
 *
 * Parity status: pinned through oracle/lpsol_oracle.py (tests/test_oracle_c.py).
 * Rows are independent in the rank-1 update, so the row loop is a pthread
 * fork-join over row ranges: that is the "all host cores" CPU baseline bench.py reports.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.
 */
#include <pthread.h>
#include <stddef.h>
#include <stdint.h>
#include <unistd.h>

#define ORC_OK 0
#define ORC_OPTIMAL 1
#define ORC_UNBOUNDED 2
#define ORC_ZERO_PIVOT 3
#define ORC_MAX_PIVOTS 4
#define ORC_OBJ_INCREASED 6   /* simplex.py:133 assert z <= obj_val */

/* ---- a minimal fork-join over row ranges (libgomp's spec file is not in the image) ---- */
static int g_threads = 0;

int orc_num_threads(void) {
    if (g_threads <= 0) {
        long k = sysconf(_SC_NPROCESSORS_ONLN);
        g_threads = k > 0 ? (int)k : 1;
    }
    return g_threads;
}

void orc_set_threads(int k) { g_threads = k > 0 ? k : 0; }

typedef void (*row_fn)(void *ctx, int lo, int hi);
typedef struct { row_fn fn; void *ctx; int lo, hi; } span_t;

static void *span_main(void *arg) {
    span_t *s = (span_t *)arg;
    s->fn(s->ctx, s->lo, s->hi);
    return NULL;
}

static void for_rows(row_fn fn, void *ctx, int nrows) {
    int k = orc_num_threads();
    if (k > 256) k = 256;
    if (k > nrows) k = nrows;
    if (k <= 1) { fn(ctx, 0, nrows); return; }
    pthread_t tid[256];
    span_t span[256];
    for (int t = 0; t < k; ++t) {
        span[t].fn = fn; span[t].ctx = ctx;
        span[t].lo = (int)((long)nrows * t / k);
        span[t].hi = (int)((long)nrows * (t + 1) / k);
        if (t > 0) pthread_create(&tid[t], NULL, span_main, &span[t]);
    }
    span_main(&span[0]);
    for (int t = 1; t < k; ++t) pthread_join(tid[t], NULL);
}

typedef struct { double *T; const double *prow; size_t ld; int n, r, c; } f64_ctx;

static void f64_rows(void *vctx, int lo, int hi) {
    f64_ctx *x = (f64_ctx *)vctx;
    for (int i = lo; i < hi; ++i) {
        if (i == x->r) continue;
        double *row = x->T + (size_t)i * x->ld;
        double mult = -row[x->c];         /* read before the row changes */
        if (mult == 0.0) continue;        /* tableau.py:272 / :285 */
        for (int j = 0; j <= x->n; ++j) row[j] = row[j] + mult * x->prow[j];
    }
}

/* tableau.py:295-308 */
int orc_pivot_f64(double *T, int m, int n, size_t ld, int r, int c) {
    double *prow = T + (size_t)r * ld;
    double p = prow[c];
    if (p == 0.0) return ORC_ZERO_PIVOT;
    double inv = 1.0 / p;                 /* rowDiv: ONE/d, then rowMult */
    if (inv != 1.0) {                     /* rowMult skips a multiplier of exactly 1 */
        for (int j = 0; j <= n; ++j) prow[j] = prow[j] * inv;
    }
    f64_ctx ctx = { T, prow, ld, n, r, c };
    for_rows(f64_rows, &ctx, m + 1);      /* row m is rowAddToObj(r, -c_c) */
    return ORC_OK;
}

/* simplex.py:262-267 (rule 0) / :228-232 (rule 1); -1 when optimal */
int orc_price_f64(const double *T, int m, int n, size_t ld, int rule) {
    const double *cost = T + (size_t)m * ld;
    int best = -1;
    for (int j = 0; j < n; ++j) {
        if (cost[j] >= 0.0) continue;
        if (rule == 1) return j;
        if (best == -1 || cost[j] < cost[best]) best = j;
    }
    return best;
}

/* simplex.py:272-279; -1 when unbounded */
int orc_ratio_f64(const double *T, int m, int n, size_t ld, int c) {
    int best = -1;
    double best_ratio = 0.0;
    for (int i = 0; i < m; ++i) {
        double a = T[(size_t)i * ld + c];
        if (a <= 0.0) continue;
        double ratio = T[(size_t)i * ld + n] / a;
        if (best == -1 || ratio < best_ratio) { best = i; best_ratio = ratio; }
    }
    return best;
}

/*
 * simplex.py:110-148.  basis[m] is updated in place; trace (may be NULL) gets
 * 3 ints per pivot (row, column, rule).  Returns ORC_OPTIMAL, ORC_UNBOUNDED,
 * ORC_OBJ_INCREASED (the reference's `assert z <= obj_val` after a Dantzig pivot,
 * simplex.py:133: a float run can trip it on rounding noise; the pivot stays applied)
 * or ORC_MAX_PIVOTS (prefix stop, not in the reference).
 */
int orc_solve_f64(double *T, int m, int n, size_t ld, int *basis, int *trace,
                  long max_pivots, long *npiv_out) {
    const double z_entry = -T[(size_t)m * ld + n];
    long stuck = 0, done = 0;
    int rule = 0, status = ORC_OPTIMAL;
    if (orc_price_f64(T, m, n, ld, 0) < 0) { *npiv_out = 0; return ORC_OPTIMAL; }
    for (;;) {
        if (rule == 0 && !(stuck < (long)m + n)) rule = 1;
        if (max_pivots >= 0 && done >= max_pivots) { status = ORC_MAX_PIVOTS; break; }
        int c = orc_price_f64(T, m, n, ld, rule);
        if (c < 0) { status = ORC_OPTIMAL; break; }
        int r = orc_ratio_f64(T, m, n, ld, c);
        if (r < 0) { status = ORC_UNBOUNDED; break; }
        orc_pivot_f64(T, m, n, ld, r, c);
        basis[r] = c;
        if (trace) { trace[3 * done] = r; trace[3 * done + 1] = c; trace[3 * done + 2] = rule; }
        ++done;
        if (rule == 0) {
            double z = -T[(size_t)m * ld + n];
            if (!(z <= z_entry)) { status = ORC_OBJ_INCREASED; break; }
            stuck = (z == z_entry) ? stuck + 1 : 0;
        }
    }
    *npiv_out = done;
    return status;
}

/*
 * Exact-mode restatement of one multi-modular fraction-free pivot for ONE
 * prime (SURVEY.md section 8a note): for i != r,
 *   N_ij <- (p*N_ij - N_ic*N_rj) * Dinv  (mod q),  row r unchanged.
 * Plain 64-bit `%` arithmetic: the slow, obviously-correct form the GPU's
 * Montgomery kernel is compared against.
 */
typedef struct { uint32_t *N; const uint32_t *prow; size_t ld; int n, r, c; uint32_t q, dinv; } modp_ctx;

static void modp_rows(void *vctx, int lo, int hi) {
    modp_ctx *x = (modp_ctx *)vctx;
    const uint64_t q = x->q, p = x->prow[x->c];
    for (int i = lo; i < hi; ++i) {
        if (i == x->r) continue;
        uint32_t *row = x->N + (size_t)i * x->ld;
        uint64_t f = row[x->c];
        for (int j = 0; j <= x->n; ++j) {
            uint64_t a = (p * row[j]) % q;
            uint64_t b = (f * x->prow[j]) % q;
            uint64_t d = (a + q - b) % q;
            row[j] = (uint32_t)((d * x->dinv) % q);
        }
    }
}

void orc_pivot_modp(uint32_t *N, int m, int n, size_t ld, int r, int c,
                    uint32_t q, uint32_t dinv) {
    modp_ctx ctx = { N, N + (size_t)r * ld, ld, n, r, c, q, dinv };
    for_rows(modp_rows, &ctx, m + 1);
}

/*
 * Synthetic input: the seeded dense family of SURVEY.md section 8(d) from a
 * counter-based hash (the same function as problems.hashed_dense_grid and the
 * CUDA library's input generator), rows [row0, row0 + m_local) of an
 * m_total-row tableau plus the objective row, written into T ((m_local+1) x ld).
 */
static uint64_t fill_mix(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
static uint64_t fill_draw(uint64_t seed, uint64_t i, uint64_t j, uint64_t span) {
    return (fill_mix(fill_mix(seed) + i * 0x100000001B3ull + j) >> 33) % span;
}
#define FILL_SPECIAL 0xFFFFFFFFull

typedef struct { double *T; size_t ld; long mt, row0; int m_local, n; uint64_t seed; int lo, hi; } fill_ctx;

static void fill_rows(void *vctx, int lo_row, int hi_row) {
    fill_ctx *x = (fill_ctx *)vctx;
    const uint64_t span = (uint64_t)(x->hi - x->lo + 1);
    const long blo = (long)x->hi * x->mt / 4, bhi = (long)x->hi * x->mt / 2;
    for (int il = lo_row; il < hi_row; ++il) {
        double *row = x->T + (size_t)il * x->ld;
        const int obj = (il == x->m_local);
        const uint64_t i = obj ? FILL_SPECIAL : (uint64_t)(x->row0 + il);
        for (long j = 0; j < (long)x->ld; ++j) {
            double v = 0.0;
            if (j < x->mt) {
                double a = (double)(x->lo + (long)fill_draw(x->seed, i, (uint64_t)j, span));
                v = obj ? -a : a;
            } else if (j < x->n) {
                v = (!obj && j - x->mt == (long)i) ? 1.0 : 0.0;
            } else if (j == x->n && !obj) {
                v = (double)(blo + (long)fill_draw(x->seed, i, FILL_SPECIAL, (uint64_t)(bhi - blo + 1)));
            }
            row[j] = v;
        }
    }
}

void orc_fill_dense_f64(double *T, long m_total, long row0, int m_local, int n, size_t ld,
                        unsigned long long seed, int lo, int hi) {
    fill_ctx ctx = { T, ld, m_total, row0, m_local, n, (uint64_t)seed, lo, hi };
    for_rows(fill_rows, &ctx, m_local + 1);
}
