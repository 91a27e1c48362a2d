// C-ABI of the B200 simplex pivot engine (see include/lps_b200.h for the contract and the
// reference interfaces each entry point replaces).  Host-side orchestration only: every
// arithmetic step of the path runs in the kernels of lps_f64.cuh / lps_exact.cuh.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "lps_internal.h"

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

int lps_fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

// pinned state mirrors, recycled (see lps_internal.h)
#include <mutex>
static std::mutex g_pin_mu;
static std::vector<void*> g_pin_free;
void* lps_pinned_get(size_t bytes) {
    if (bytes > 256) return nullptr;
    {
        std::lock_guard<std::mutex> lk(g_pin_mu);
        if (!g_pin_free.empty()) { void* p = g_pin_free.back(); g_pin_free.pop_back(); return p; }
    }
    void* p = nullptr;
    if (cudaMallocHost(&p, 256) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
void lps_pinned_put(void* p) {
    if (!p) return;
    std::lock_guard<std::mutex> lk(g_pin_mu);
    g_pin_free.push_back(p);
}

static F64Grid grid_of(const lps_handle* h) {
    F64Grid g;
    g.T = h->T; g.ld = h->ld; g.m = h->m; g.n = h->n; g.row0 = h->row0; g.m_total = h->m_total;
    return g;
}
static PricePartials partials_of(const lps_handle* h) { return PricePartials{h->pd_val, h->pd_idx, h->pb_idx}; }

int lps_use(const lps_handle* h) {
    if (!h) return fail(LPS_E_INVALID, "null handle");
    CU(cudaSetDevice(h->device));
    return LPS_OK;
}

// ------------------------------------------------------------------------------------------------
// library
// ------------------------------------------------------------------------------------------------
extern "C" int lps_version(void) { return 100; }
extern "C" const char* lps_last_error(void) { return g_err; }
extern "C" int lps_device_count(int* count) {
    if (!count) return fail(LPS_E_INVALID, "null count");
    CU(cudaGetDeviceCount(count));
    return LPS_OK;
}
extern "C" size_t lps_grid_ld(int n) { return ((size_t)n + 1 + 15) / 16 * 16; }
extern "C" size_t lps_cand_bytes(int n) { return sizeof(CandHeader) + lps_grid_ld(n) * sizeof(double); }

static void plan(lps_handle* h) {
    const size_t bytes = (size_t)(h->m + 1) * h->ld * sizeof(double);
    h->stream_hint = bytes > kL2Bytes;
    // wide chunks + tall row blocks once there is enough work for several waves of CTAs
    const size_t tiles_wide = ((h->ld + kWideChunk - 1) / kWideChunk) * (((size_t)h->m + 1 + 63) / 64);
    // (row blocks of 32 on shards of up to 16k rows: more waves of CTAs, a shorter tail -- +1.4 % at 4096 x 65536)
    if (tiles_wide >= 148 * 8) { h->chunk = kWideChunk; h->rows_per_cta = (h->m + 1 > 16384) ? 64 : 32; }
    else { h->chunk = kNarrowChunk; h->rows_per_cta = ((size_t)(h->m + 1) * h->ld >= (1u << 20)) ? 16 : 8; }
    // tuning knobs for sweeps (tools/sweep_update.sh); not part of the ABI
    if (const char* e = getenv("LPS_F64_CHUNK")) { const int v = atoi(e); if (v == kWideChunk || v == kNarrowChunk) h->chunk = v; }
    if (const char* e = getenv("LPS_F64_ROWS")) { const int v = atoi(e); if (v >= 4 && v <= kMaxRowBlock && v % 4 == 0) h->rows_per_cta = v; }
    h->nchunks = (int)((h->ld + h->chunk - 1) / h->chunk);
    // k_select_f64 spins on an in-kernel flag: never more CTAs than SMs, so all of them are resident.  Small
    // tableaux take ONE CTA of 1024 threads instead: no ticket, no flag, no fences -- the launch is a plain
    // dependent chain of a few memory round trips.
    if (h->m <= 8 * kSelMaxThreads && h->ld <= 16 * kSelMaxThreads) { h->sel_blocks = 1; h->sel_threads = kSelMaxThreads; }
    else {
        h->sel_threads = kThreads;
        h->sel_blocks = std::max(1, std::min(std::max((h->m + kThreads - 1) / kThreads, (int)(h->ld / (8 * kThreads))), 148));
    }
    // Pivots in blocks (lps_f64_blk.cuh).  The panel is a persistent grid that synchronises through memory: never
    // more CTAs than SMs.  Small tableaux: one CTA of 1024 threads (barriers are __syncthreads).
    int sms = 148;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device) != cudaSuccess || sms <= 0) { cudaGetLastError(); sms = 148; }
    // (measured, profiles/r02d_panel_shapes.txt: at 2000 x 4000 eight CTAs of 256 threads decide a pivot in 15 us, the
    // one-CTA panel in 25 us; at 500 x 1000 and below the one-CTA panel is the faster one)
    if (h->m < 1024 && h->ld < 2048) { h->pan_blocks = 1; h->pan_threads = kSelMaxThreads; }
    else {
        h->pan_threads = kThreads;
        h->pan_blocks = std::max(2, std::min(std::max((h->m + kThreads - 1) / kThreads, (int)((h->ld / 2 + kThreads - 1) / kThreads)), std::min(sms, 256)));
    }
    h->sel_blocks = std::min(h->sel_blocks, sms);
    // Medium tableaux: the whole panel fits the shared memory of one thread-block cluster (k_blk_panel_cl) -- no L2
    // round trips and no global-memory barriers inside a pivot.  8 CTAs when that is enough, else 16.
    h->pan_cluster = 0;
    for (int C : {8, 16}) {
        const int rpc = (h->m + C - 1) / C, cpc = (int)(((h->ld + C - 1) / C + 1) / 2 * 2);
        if (blk_cluster_smem(rpc, cpc) <= (size_t)200 * 1024) { h->pan_cluster = C; h->pan_rpc = rpc; h->pan_cpc = cpc; break; }
    }
    if (const char* e = getenv("LPS_F64_PANEL_CLUSTER")) {
        const int v = atoi(e);
        if (v == 0) h->pan_cluster = 0;
        else if (v == 8 || v == 16) {
            const int rpc = (h->m + v - 1) / v, cpc = (int)(((h->ld + v - 1) / v + 1) / 2 * 2);
            if (blk_cluster_smem(rpc, cpc) <= (size_t)200 * 1024) { h->pan_cluster = v; h->pan_rpc = rpc; h->pan_cpc = cpc; }
        }
    }
    h->sms = sms;
    // Blocks pay wherever the panel decides a pivot faster than a (select, update) pair of the rank-1 loop runs: with
    // the cluster panel at every size it fits (50 x 100: 81 k pivots/s against 36 k, 2000 x 4000: 95 k against 41 k,
    // profiles/r02g_cluster_panel.txt), with the grid-wide panel once a pass over the grid costs more than its ~15 us
    // per pivot -- from about 1.5 M cells on.
    h->blk_k = (h->pan_cluster > 0 || (size_t)(h->m + 1) * h->ld >= (size_t)1500000) ? kBlkMax : 0;
    if (const char* e = getenv("LPS_F64_BLOCK")) { const int v = atoi(e); if (v >= 0 && v <= kBlkMax) h->blk_k = v; }
    if (const char* e = getenv("LPS_F64_PANEL_CTAS")) {   // tuning: 1 = one CTA of 1024 threads, >= 2 = that many CTAs of 256
        const int v = atoi(e);
        if (v == 1) { h->pan_blocks = 1; h->pan_threads = kSelMaxThreads; }
        else if (v >= 2 && v <= std::min(sms, 256)) { h->pan_blocks = v; h->pan_threads = kThreads; }
    }
}

// Exchange block: headers and flags, then the row area -- two sets of kBlkMax candidate rows per rank (the rank-1 loop
// uses the first row as its vrow).  Sized for the number of ranks that will store into it.
static int alloc_xchg(lps_handle* h, int nranks) {
    if (h->xchg && h->blk.nr == nranks) return LPS_OK;
    if (h->xchg) { CU(cudaStreamSynchronize(h->stream)); CU(cudaFree(h->xchg)); h->xchg = nullptr; }
    const size_t xb = sizeof(XchgBlock) + 2 * (size_t)nranks * kBlkMax * h->ld * sizeof(double);
    cudaError_t e = cudaMalloc(&h->xchg, xb);
    if (e != cudaSuccess) return fail(LPS_E_NOMEM, "cudaMalloc of the %zu-byte exchange block failed: %s", xb, cudaGetErrorString(e));
    CU(cudaMemsetAsync(h->xchg, 0, xb, h->stream));
    CU(cudaMemcpyAsync(h->d_peers, &h->xchg, sizeof(XchgBlock*), cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));           // (&h->xchg is read by the copy)
    h->xchg_bytes = xb;
    h->blk.V = xchg_vrow(h->xchg);
    h->blk.nr = nranks;
    if (h->gexec) { cudaGraphExecDestroy(h->gexec); h->gexec = nullptr; }
    return LPS_OK;
}

extern "C" int lps_create_f64(int m, int n, int device, void* dev_grid, void* stream, lps_handle** out) {
    if (!out) return fail(LPS_E_INVALID, "null out");
    *out = nullptr;
    if (m <= 0) return fail(LPS_E_INVALID, "need m > 0, provided m = %d", m);
    if (n <= 0) return fail(LPS_E_INVALID, "need n > 0, provided n = %d", n);
    int ndev = 0;
    CU(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(LPS_E_INVALID, "device %d out of range (%d visible)", device, ndev);
    CU(cudaSetDevice(device));
    lps_handle* h = new (std::nothrow) lps_handle();
    if (!h) return fail(LPS_E_NOMEM, "out of host memory");
    h->device = device;
    h->m = m; h->n = n; h->ld = lps_grid_ld(n);
    h->row0 = 0; h->m_total = m;
    if (stream) h->stream = (cudaStream_t)stream;
    else { CU(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)); h->own_stream = true; }
    const size_t grid_bytes = (size_t)(m + 1) * h->ld * sizeof(double);
    if (dev_grid) h->T = (double*)dev_grid;
    else {
        cudaError_t e = cudaMalloc(&h->T, grid_bytes);
        if (e != cudaSuccess) { delete h; return fail(LPS_E_NOMEM, "cudaMalloc of %zu bytes failed: %s", grid_bytes, cudaGetErrorString(e)); }
        h->own_T = true;
        CU(cudaMemsetAsync(h->T, 0, grid_bytes, h->stream));
    }
    plan(h);
    // the fused update's shared-memory ring (stages of the grid, pivot rows): far beyond the 48 KB default
    CU(cudaFuncSetAttribute(k_blk_update, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BlkRing)));
    CU(cudaFuncSetAttribute(k_blk_panel_cl, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CU(cudaFuncSetAttribute(k_blk_panel_cl, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    const size_t ucol_n = ((size_t)m + 1 + kMaxRowBlock - 1) / kMaxRowBlock * kMaxRowBlock + kMaxRowBlock;
    const int maxchunks = (int)((h->ld + kNarrowChunk - 1) / kNarrowChunk);
    {   // the small buffers: one allocation (lps_internal.h, Slab)
        Slab sl;
        sl.add(&h->ucol, ucol_n);
        sl.add(&h->d_peers, (size_t)kMaxRanks);
        sl.add(&h->pd_val, (size_t)maxchunks);
        sl.add(&h->pd_idx, (size_t)maxchunks);
        sl.add(&h->pb_idx, (size_t)maxchunks);
        sl.add(&h->sel_blk, (size_t)148 * 4);
        sl.add(&h->st, (size_t)1);
        sl.add(&h->scratch_i, 2 * (size_t)n + 16);
        sl.add(&h->marks, (size_t)n);
        h->cand_stride = lps_cand_bytes(n);
        sl.add(&h->cand_send, h->cand_stride);           // replaced, never freed, when lps_set_shard passes the caller's slots
        // panel of the blocked pivot loop (lps_f64_blk.cuh)
        h->blk.upitch = ucol_n;
        sl.add(&h->blk.U, (size_t)kBlkRec * ucol_n);
        sl.add(&h->blk.O[0], h->ld);
        sl.add(&h->blk.O[1], h->ld);
        sl.add(&h->blk.B[0], ucol_n);
        sl.add(&h->blk.B[1], ucol_n);
        sl.add(&h->blk.R, (size_t)kBlkMax);
        sl.add(&h->blk.C, (size_t)kBlkMax);
        sl.add(&h->blk.W, (size_t)kBlkMax);
        sl.add(&h->blk.hdr, (size_t)4);
        sl.add(&h->blk.p_val, (size_t)256);
        sl.add(&h->blk.p_idx, (size_t)256);
        sl.add(&h->blk.p_bidx, (size_t)256);
        sl.add(&h->blk.rblk, (size_t)256);
        sl.add(&h->blk.sync, (size_t)1);
        CU(sl.commit(&h->slab));
    }
    CU(cudaMemsetAsync(h->blk.U, 0, (size_t)kBlkRec * ucol_n * sizeof(double), h->stream));
    CU(cudaMemsetAsync(h->blk.hdr, 0, 4 * sizeof(int), h->stream));
    CU(cudaMemsetAsync(h->blk.sync, 0, sizeof(BlkSync), h->stream));
    CU(cudaMemsetAsync(h->ucol, 0, ucol_n * sizeof(double), h->stream));
    CU(cudaMemsetAsync(h->cand_send, 0, h->cand_stride, h->stream));
    h->cand_all = h->cand_send;
    h->own_cand = false;
    TRY(alloc_xchg(h, 1));
    CU(cudaMemsetAsync(h->st, 0, sizeof(SolveState), h->stream));
    static_assert(sizeof(SolveState) <= 256, "pinned state blocks are 256 bytes");
    h->st_host = (SolveState*)lps_pinned_get(sizeof(SolveState));
    if (!h->st_host) return fail(LPS_E_NOMEM, "no pinned host memory for the state mirror");
    CU(cudaMalloc(&h->basis, (size_t)m * sizeof(int)));        // its own allocation: re-sized by lps_set_shard / lps_p2p_attach
    CU(cudaMemsetAsync(h->basis, 0xff, (size_t)m * sizeof(int), h->stream));
    CU(cudaMemsetAsync(h->marks, 0, (size_t)n * sizeof(int), h->stream));
    *out = h;
    return LPS_OK;
}

// A handle created on a lent grid (dev_grid) moves to another one of the same shape: lets the host keep handles of
// large tableaux (their small buffers, stream, state) across tableaux without keeping the grids themselves.
extern "C" int lps_rebind_grid_f64(lps_handle* h, void* dev_grid) {
    USE(h);
    NEED_MODE(h, 0);
    if (!dev_grid) return fail(LPS_E_INVALID, "null grid");
    if (h->own_T) return fail(LPS_E_STATE, "the handle owns its grid");
    if (h->solving) return fail(LPS_E_STATE, "solve in progress");
    CU(cudaStreamSynchronize(h->stream));
    h->T = (double*)dev_grid;
    if (h->gexec) { cudaGraphExecDestroy(h->gexec); h->gexec = nullptr; }   // the captured launches carry the old pointer
    return LPS_OK;
}

extern "C" int lps_set_shard(lps_handle* h, long long row0, long long m_total, int rank, int nranks, void* cand_send,
                             void* cand_all) {
    USE(h);
    NEED_MODE(h, 0);
    if (nranks < 1 || nranks > kMaxRanks || rank < 0 || rank >= nranks) return fail(LPS_E_INVALID, "bad rank %d of %d", rank, nranks);
    if (row0 < 0 || row0 + h->m > m_total) return fail(LPS_E_INVALID, "rows [%lld,%lld) outside 0..%lld", row0, row0 + h->m, m_total);
    if (nranks > 1 && (!cand_send || !cand_all)) return fail(LPS_E_INVALID, "sharded handles need candidate buffers");
    h->row0 = row0; h->m_total = m_total; h->rank = rank; h->nranks = nranks;
    if (cand_send && cand_all) {
        if (h->own_cand) { CU(cudaFree(h->cand_send)); h->own_cand = false; }
        h->cand_send = (unsigned char*)cand_send;
        h->cand_all = (unsigned char*)cand_all;
    }
    h->use_xchg = false;                // candidate slots + the caller's all-gather
    CU(cudaFree(h->basis));
    CU(cudaMalloc(&h->basis, (size_t)m_total * sizeof(int)));
    CU(cudaMemsetAsync(h->basis, 0xff, (size_t)m_total * sizeof(int), h->stream));
    return LPS_OK;
}

extern "C" int lps_destroy(lps_handle* h) {
    if (!h) return LPS_OK;
    if (h->mode == 1) return lpsx_destroy(h);
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    for (cudaEvent_t e : h->ev) cudaEventDestroy(e);
    if (h->own_T) cudaFree(h->T);
    if (h->own_cand) cudaFree(h->cand_send);
    if (h->gexec) cudaGraphExecDestroy(h->gexec);
    for (void* p : h->ipc_open) cudaIpcCloseMemHandle(p);
    cudaFree(h->xchg);
    cudaFree(h->slab);
    lps_pinned_put(h->st_host);
    cudaFree(h->basis);
    if (h->trace) cudaFree(h->trace);
    if (h->own_stream) cudaStreamDestroy(h->stream);
    delete h;
    return LPS_OK;
}

extern "C" int lps_sync(lps_handle* h) {
    USE(h);
    CU(cudaStreamSynchronize(h->stream));
    return LPS_OK;
}
extern "C" void* lps_grid_ptr(lps_handle* h) { return (h && h->mode == 0) ? (void*)h->T : nullptr; }

// ------------------------------------------------------------------------------------------------
// transfers
// ------------------------------------------------------------------------------------------------
extern "C" int lps_upload_f64(lps_handle* h, const double* host, size_t ld_host) {
    USE(h);
    NEED_MODE(h, 0);
    if (!host || ld_host < (size_t)h->n + 1) return fail(LPS_E_INVALID, "bad host grid");
    CU(cudaMemcpy2DAsync(h->T, h->ld * sizeof(double), host, ld_host * sizeof(double), ((size_t)h->n + 1) * sizeof(double),
                         (size_t)h->m + 1, cudaMemcpyHostToDevice, h->stream));
    if (h->ld > (size_t)h->n + 1)   // keep the padding columns at zero
        CU(cudaMemset2DAsync(h->T + h->n + 1, h->ld * sizeof(double), 0, (h->ld - h->n - 1) * sizeof(double), (size_t)h->m + 1, h->stream));
    return LPS_OK;
}
extern "C" int lps_download_f64(lps_handle* h, double* host, size_t ld_host) {
    USE(h);
    NEED_MODE(h, 0);
    if (!host || ld_host < (size_t)h->n + 1) return fail(LPS_E_INVALID, "bad host grid");
    CU(cudaMemcpy2DAsync(host, ld_host * sizeof(double), h->T, h->ld * sizeof(double), ((size_t)h->n + 1) * sizeof(double),
                         (size_t)h->m + 1, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return LPS_OK;
}
extern "C" int lps_read_row_f64(lps_handle* h, int i, double* host) {
    USE(h);
    NEED_MODE(h, 0);
    if (!host || i < 0 || i > h->m) return fail(LPS_E_INVALID, "row %d out of range", i);
    CU(cudaMemcpyAsync(host, h->T + (size_t)i * h->ld, ((size_t)h->n + 1) * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return LPS_OK;
}
extern "C" int lps_read_col_f64(lps_handle* h, int j, double* host) {
    USE(h);
    NEED_MODE(h, 0);
    if (!host || j < 0 || j > h->n) return fail(LPS_E_INVALID, "column %d out of range", j);
    // gathered on the device, then ONE contiguous copy (a strided 2-D copy of 8-byte rows costs a DMA descriptor per row)
    DevBuf b_col;
    CU(b_col.alloc(((size_t)h->m + 1) * sizeof(double)));
    k_gather_col_f64<<<std::max(1, std::min((h->m + 256) / 256, 592)), 256, 0, h->stream>>>(grid_of(h), j, b_col.as<double>());
    h->launches_all++;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(host, b_col.p, ((size_t)h->m + 1) * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return LPS_OK;
}


extern "C" int lps_fill_dense_f64(lps_handle* h, unsigned long long seed, int lo, int hi) {
    USE(h);
    NEED_MODE(h, 0);
    if (lo > hi) return fail(LPS_E_INVALID, "need lo <= hi");
    if ((long long)h->n != 2 * h->m_total) return fail(LPS_E_INVALID, "the dense family has n = 2 m (n = %d, m = %lld)", h->n, h->m_total);
    dim3 grid((unsigned)std::min<size_t>((h->ld + 255) / 256, 64), (unsigned)std::min(h->m + 1, 148 * 16));
    k_fill_dense_f64<<<grid, 256, 0, h->stream>>>(grid_of(h), (uint64_t)seed, lo, hi);
    h->launches_all++;
    CU(cudaGetLastError());
    return LPS_OK;
}

// ------------------------------------------------------------------------------------------------
// launches
// ------------------------------------------------------------------------------------------------
static int launch_price(lps_handle* h) {
    if (h->chunk == kWideChunk) k_price_f64<kWideChunk><<<h->nchunks, kThreads, 0, h->stream>>>(grid_of(h), partials_of(h));
    else k_price_f64<kNarrowChunk><<<h->nchunks, kThreads, 0, h->stream>>>(grid_of(h), partials_of(h));
    h->launches_all++;
    CU(cudaGetLastError());
    h->pp_valid = true;
    return LPS_OK;
}
static int launch_select(lps_handle* h) {
    if (!h->pp_valid) TRY(launch_price(h));   // the blocked loop does not maintain the rank-1 path's pricing partials
    SelectArgs a;
    a.g = grid_of(h); a.pp = partials_of(h); a.nchunks = h->nchunks; a.st = h->st; a.ucol = h->ucol;
    a.blk = h->sel_blk; a.cand = (CandHeader*)h->cand_send;
    a.xchg = h->use_xchg ? h->xchg : nullptr; a.peers = h->d_peers; a.rank = h->rank; a.nranks = h->nranks;
    k_select_f64<<<h->sel_blocks, h->sel_threads, 0, h->stream>>>(a);
    h->launches_all++;
    CU(cudaGetLastError());
    return LPS_OK;
}
static int launch_update(lps_handle* h, int commit) {
    UpdateArgs a;
    a.g = grid_of(h); a.pp = partials_of(h); a.st = h->st; a.ucol = h->ucol;
    a.cand_all = h->cand_all; a.cand_stride = h->cand_stride; a.nranks = h->nranks;
    a.xchg = h->use_xchg ? h->xchg : nullptr;
    a.rows_per_cta = h->rows_per_cta; a.basis = h->basis; a.marks = h->use_marks ? h->marks : nullptr; a.trace = h->trace; a.trace_cap = h->trace_cap;
    a.commit = commit;
    dim3 grid(h->nchunks, (h->m + 1 + h->rows_per_cta - 1) / h->rows_per_cta);
    const bool timed = h->prof && h->ev_used + 2 <= h->ev.size();
    if (timed) CU(cudaEventRecord(h->ev[h->ev_used++], h->stream));
    if (h->chunk == kWideChunk) {
        if (h->stream_hint) k_update_f64<kWideChunk, true><<<grid, kThreads, 0, h->stream>>>(a);
        else k_update_f64<kWideChunk, false><<<grid, kThreads, 0, h->stream>>>(a);
    } else {
        if (h->stream_hint) k_update_f64<kNarrowChunk, true><<<grid, kThreads, 0, h->stream>>>(a);
        else k_update_f64<kNarrowChunk, false><<<grid, kThreads, 0, h->stream>>>(a);
    }
    if (timed) CU(cudaEventRecord(h->ev[h->ev_used++], h->stream));
    h->launches_update++;
    h->launches_all++;
    h->blk_valid = false;             // the panel's copy of the objective row / right-hand side is stale now
    CU(cudaGetLastError());
    return LPS_OK;
}
// ---- pivots in blocks (lps_f64_blk.cuh) ----------------------------------------------------------
static int blk_sync_panel(lps_handle* h) {
    // O / B / pricing partials of the panel from the grid (after anything but the blocked loop touched it)
    if (h->blk_valid) return LPS_OK;
    CU(cudaMemsetAsync(h->blk.sync, 0, sizeof(BlkSync), h->stream));
    k_blk_init<<<h->pan_blocks, h->pan_threads, 0, h->stream>>>(grid_of(h), h->blk);
    h->launches_all++;
    CU(cudaGetLastError());
    h->blk_valid = true;
    return LPS_OK;
}
static int launch_panel_cluster(lps_handle* h) {
    BlkClusterArgs a;
    a.g = grid_of(h); a.p = h->blk; a.st = h->st; a.kmax = h->blk_k; a.rpc = h->pan_rpc; a.cpc = h->pan_cpc;
    a.basis = h->basis; a.marks = h->use_marks ? h->marks : nullptr; a.trace = h->trace; a.trace_cap = h->trace_cap;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)h->pan_cluster, 1, 1);
    cfg.blockDim = dim3(kClThreads, 1, 1);
    cfg.dynamicSmemBytes = blk_cluster_smem(h->pan_rpc, h->pan_cpc);
    cfg.stream = h->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)h->pan_cluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, k_blk_panel_cl, a);
    if (e != cudaSuccess) return fail(LPS_E_CUDA, "k_blk_panel_cl launch failed: %s", cudaGetErrorString(e));
    h->launches_all++;
    return LPS_OK;
}
static int launch_panel(lps_handle* h) {
    if (h->pan_cluster > 0 && !(h->use_xchg && h->nranks > 1)) return launch_panel_cluster(h);
    BlkPanelArgs a;
    a.g = grid_of(h); a.p = h->blk; a.st = h->st; a.kmax = h->blk_k;
    a.xchg = h->xchg; a.peers = h->d_peers; a.rank = h->rank; a.nranks = h->use_xchg ? h->nranks : 1;
    a.basis = h->basis; a.marks = h->use_marks ? h->marks : nullptr; a.trace = h->trace; a.trace_cap = h->trace_cap;
    if (h->pan_threads == kSelMaxThreads) k_blk_panel<kSelMaxThreads><<<h->pan_blocks, kSelMaxThreads, 0, h->stream>>>(a);
    else k_blk_panel<kThreads><<<h->pan_blocks, kThreads, 0, h->stream>>>(a);
    h->launches_all++;
    CU(cudaGetLastError());
    return LPS_OK;
}
static int launch_blk_update(lps_handle* h) {
    BlkUpdateArgs a;
    a.g = grid_of(h); a.p = h->blk;
    // persistent: one CTA per SM, each walks its share of the (512-column chunk, 8-row stage) items
    const long long items = (long long)((h->ld + kBlkCW - 1) / kBlkCW) * (((long long)h->m + 1 + kBlkRU - 1) / kBlkRU);
    const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>(h->sms, items));
    const bool timed = h->prof && h->ev_used + 2 <= h->ev.size();
    if (timed) CU(cudaEventRecord(h->ev[h->ev_used++], h->stream));
    k_blk_update<<<grid, kBlkConsumers + 32, sizeof(BlkRing), h->stream>>>(a);
    if (timed) CU(cudaEventRecord(h->ev[h->ev_used++], h->stream));
    h->launches_update++;
    h->launches_all++;
    h->pp_valid = false;
    CU(cudaGetLastError());
    return LPS_OK;
}
static int launch_block(lps_handle* h) {
    TRY(blk_sync_panel(h));
    TRY(launch_panel(h));
    return launch_blk_update(h);
}

static int read_state(lps_handle* h) {
    CU(cudaMemcpyAsync(h->st_host, h->st, sizeof(SolveState), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return LPS_OK;
}
static int write_state(lps_handle* h, const SolveState& s) {
    *h->st_host = s;
    CU(cudaMemcpyAsync(h->st, h->st_host, sizeof(SolveState), cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));   // st_host is reused
    return LPS_OK;
}
// ------------------------------------------------------------------------------------------------
// hot path entry points
// ------------------------------------------------------------------------------------------------
extern "C" int lps_pivot(lps_handle* h, int r, int c) {
    USE(h);
    if (h->mode == 1) return lpsx_pivot(h, r, c);
    if (h->nranks != 1) return fail(LPS_E_STATE, "explicit pivots are not supported on a sharded handle");
    if (r < 0 || r >= h->m || c < 0 || c >= h->n) return fail(LPS_E_INVALID, "pivot (%d,%d) out of range", r, c);
    TRY(read_state(h));
    SolveState s{};
    s.status = LPS_OK; s.max_pivots = -1; s.xseq = h->st_host->xseq;
    TRY(write_state(h, s));
    const int blocks = std::max(1, std::min((int)((std::max((size_t)h->m + 1, h->ld) + kThreads - 1) / kThreads), 148 * 2));
    k_forced_f64<<<blocks, kThreads, 0, h->stream>>>(grid_of(h), h->st, h->ucol, (CandHeader*)h->cand_send, r, c);
    h->launches_all++;
    CU(cudaGetLastError());
    TRY(launch_update(h, 0));
    TRY(read_state(h));
    const int st = h->st_host->status;
    return st == LPS_ZERO_PIVOT ? LPS_ZERO_PIVOT : LPS_OK;
}

extern "C" int lps_row_scale_f64(lps_handle* h, int r, double mult) {
    USE(h);
    NEED_MODE(h, 0);
    if (r < 0 || r >= h->m) return fail(LPS_E_INVALID, "row %d out of range", r);
    if (mult == 1.0) return LPS_OK;   // tableau.py:257
    k_row_scale_f64<<<std::min((h->n + 256) / 256, 592), 256, 0, h->stream>>>(grid_of(h), r, mult);
    h->launches_all++;
    CU(cudaGetLastError());
    return LPS_OK;
}
extern "C" int lps_row_axpy_f64(lps_handle* h, int rd, int rs, double mult) {
    USE(h);
    NEED_MODE(h, 0);
    if (rd < 0 || rd > h->m || rs < 0 || rs >= h->m) return fail(LPS_E_INVALID, "rows (%d,%d) out of range", rd, rs);
    if (mult == 0.0) return LPS_OK;   // tableau.py:272, :285
    k_row_axpy_f64<<<std::min((h->n + 256) / 256, 592), 256, 0, h->stream>>>(grid_of(h), rd, rs, mult);
    h->launches_all++;
    CU(cudaGetLastError());
    return LPS_OK;
}

// Leaving phase 1 without a host round trip of the tableau (Simplex._find_bfs simplex.py:95-105 when no row is
// dropped): the artificial columns n_new..n-1 are cut, the objective row becomes `cost` (n_new reduced costs, then
// the _z cell) and is priced out against the basis by the reference's m rowSubFromObj steps -- one small launch per
// row, in row order, each reading its multiplier from the objective row the previous steps left (k_obj_step_f64).
extern "C" int lps_phase2_f64(lps_handle* h, int n_new, const double* cost, const int* basis) {
    USE(h);
    NEED_MODE(h, 0);
    if (h->nranks != 1) return fail(LPS_E_STATE, "lps_phase2_f64 on a sharded handle");
    if (h->solving) return fail(LPS_E_STATE, "solve in progress");
    if (!cost || !basis || n_new <= 0 || n_new > h->n) return fail(LPS_E_INVALID, "bad arguments (n_new = %d, n = %d)", n_new, h->n);
    for (int i = 0; i < h->m; ++i)
        if (basis[i] < 0 || basis[i] >= n_new) return fail(LPS_E_INVALID, "basis[%d] = %d is not a column below %d", i, basis[i], n_new);
    DevBuf b_vec;
    const size_t row_n = ((size_t)n_new + 1 + 15) / 16 * 16;
    CU(b_vec.alloc(((size_t)h->m + 1 + 2 * row_n) * sizeof(double)));
    double* d_rhs = b_vec.as<double>();
    double* d_a = d_rhs + h->m + 1;
    double* d_b = d_a + row_n;
    CU(cudaMemcpyAsync(d_a, cost, ((size_t)n_new + 1) * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    if (n_new < h->n) {
        k_gather_col_f64<<<std::max(1, std::min((h->m + 256) / 256, 592)), 256, 0, h->stream>>>(grid_of(h), h->n, d_rhs);
        const long long cells = (long long)(h->m + 1) * (h->n - n_new + 1);
        k_cut_cols_f64<<<(unsigned)std::max<long long>(1, std::min<long long>((cells + 255) / 256, 148 * 16)), 256, 0, h->stream>>>(grid_of(h), n_new, d_rhs);
        h->launches_all += 2;
        h->n = n_new;                                   // ld, launch plan and buffers keep their (larger) sizes
        if (h->gexec) { cudaGraphExecDestroy(h->gexec); h->gexec = nullptr; }   // the captured launches carry the old n
    }
    const F64Grid g = grid_of(h);
    const int blocks = std::max(1, std::min((n_new + 1 + 255) / 256, 296));
    for (int i = 0; i < h->m; ++i) {
        k_obj_step_f64<<<blocks, 256, 0, h->stream>>>(g, d_a, d_b, i, basis[i]);
        std::swap(d_a, d_b);
    }
    h->launches_all += h->m;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(h->T + (size_t)h->m * h->ld, d_a, ((size_t)n_new + 1) * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));               // the scratch buffer goes out of scope
    return LPS_OK;
}

extern "C" int lps_find_pivot(lps_handle* h, int rule, int* r, int* c) {
    USE(h);
    if (!r || !c || (rule != LPS_RULE_DANTZIG && rule != LPS_RULE_BLAND)) return fail(LPS_E_INVALID, "bad arguments");
    if (h->mode == 1) return lpsx_find_pivot(h, rule, r, c);
    if (h->nranks != 1) return fail(LPS_E_STATE, "find_pivot on a sharded handle: use the stepping calls");
    TRY(read_state(h));
    SolveState s{};
    s.status = LPS_OK; s.rule = rule; s.max_pivots = -1; s.dry_run = 1;
    s.stuck = 0; s.last_r = -1; s.last_c = -1; s.xseq = h->st_host->xseq;
    TRY(write_state(h, s));
    TRY(launch_price(h));
    TRY(launch_select(h));
    TRY(read_state(h));
    const SolveState& o = *h->st_host;
    *r = o.last_r; *c = o.last_c;
    if (o.status == LPS_OPTIMAL) return LPS_OPTIMAL;
    if (o.last_r < 0) return LPS_UNBOUNDED;
    return LPS_OK;
}

extern "C" int lps_is_optimal(lps_handle* h, int* optimal) {
    int r, c;
    if (!optimal) return fail(LPS_E_INVALID, "null result");
    int rc = lps_find_pivot(h, LPS_RULE_BLAND, &r, &c);
    if (rc < 0) return rc;
    *optimal = (rc == LPS_OPTIMAL);
    return LPS_OK;
}

// Tableau.isInfeasible (tableau.py:510-514): some row with b_i > 0 and no positive entry.
extern "C" int lps_is_infeasible(lps_handle* h, int* infeasible) {
    USE(h);
    if (!infeasible) return fail(LPS_E_INVALID, "null result");
    if (h->mode == 1) return lpsx_is_infeasible(h, infeasible);
    if (h->nranks != 1) return fail(LPS_E_STATE, "is_infeasible on a sharded handle");
    DevBuf b_flag;
    CU(b_flag.alloc(sizeof(int)));
    CU(cudaMemsetAsync(b_flag.p, 0, sizeof(int), h->stream));
    k_row_infeasible_f64<<<std::max(1, std::min((h->m + 7) / 8, 148 * 8)), 256, 0, h->stream>>>(grid_of(h), b_flag.as<int>());
    h->launches_all++;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(infeasible, b_flag.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return LPS_OK;
}

// Simplex.findPivotMaxIncrease (simplex.py:286-328) on the device, fp64 mode.
extern "C" int lps_find_pivot_max_increase(lps_handle* h, int* r, int* c) {
    USE(h);
    NEED_MODE(h, 0);
    if (!r || !c) return fail(LPS_E_INVALID, "null output");
    if (h->nranks != 1) return fail(LPS_E_STATE, "whole-tableau rules are not available on a sharded handle");
    DevBuf b_ratio, b_int;
    CU(b_ratio.alloc((size_t)h->n * sizeof(double)));
    CU(b_int.alloc((2 * (size_t)h->n + 4) * sizeof(int)));
    double* ratio = b_ratio.as<double>();
    int* irow = b_int.as<int>();
    int* ties = irow + h->n;
    int* out = ties + h->n;
    k_col_ratio_f64<<<(h->n + 127) / 128, 128, 0, h->stream>>>(grid_of(h), ratio, irow, ties);
    k_max_increase_f64<<<1, 1024, 0, h->stream>>>(grid_of(h), ratio, irow, out);
    h->launches_all += 2;
    int res[3] = {0, -1, -1};
    CU(cudaMemcpyAsync(res, out, sizeof(res), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    *r = res[1]; *c = res[2];
    return res[0];
}

// Simplex.findPivotAll (simplex.py:330-360) on the device, fp64 mode: every (row, column) of minimum ratio, column by
// column, rows ascending.  *count receives the number of pivots; at most `cap` of them are written.
extern "C" int lps_find_pivot_all(lps_handle* h, int* rows, int* cols, long long cap, long long* count) {
    USE(h);
    NEED_MODE(h, 0);
    if (!count || cap < 0 || (cap > 0 && (!rows || !cols))) return fail(LPS_E_INVALID, "bad arguments");
    if (h->nranks != 1) return fail(LPS_E_STATE, "whole-tableau rules are not available on a sharded handle");
    const int n = h->n;
    DevBuf b_ratio, b_int, b_off, b_out;
    CU(b_ratio.alloc((size_t)n * sizeof(double)));
    CU(b_int.alloc(2 * (size_t)n * sizeof(int)));
    double* ratio = b_ratio.as<double>();
    int* irow = b_int.as<int>();
    int* ties = irow + n;
    k_col_ratio_f64<<<(n + 127) / 128, 128, 0, h->stream>>>(grid_of(h), ratio, irow, ties);
    h->launches_all++;
    std::vector<int> hrow(n), hties(n);
    CU(cudaMemcpyAsync(hrow.data(), irow, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaMemcpyAsync(hties.data(), ties, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    std::vector<long long> off(n);
    long long total = 0;
    for (int j = 0; j < n; ++j) { off[j] = total; if (hrow[j] >= 0) total += hties[j]; }
    *count = total;
    int rc = LPS_OK;
    if (total > 0 && total <= cap) {
        CU(b_off.alloc((size_t)n * sizeof(long long)));
        CU(b_out.alloc(2 * (size_t)total * sizeof(int)));
        long long* d_off = b_off.as<long long>();
        int* d_out = b_out.as<int>();
        CU(cudaMemcpyAsync(d_off, off.data(), (size_t)n * sizeof(long long), cudaMemcpyHostToDevice, h->stream));
        k_col_ties_f64<<<(n + 127) / 128, 128, 0, h->stream>>>(grid_of(h), ratio, irow, d_off, d_out, d_out + total);
        h->launches_all++;
        CU(cudaMemcpyAsync(rows, d_out, (size_t)total * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        CU(cudaMemcpyAsync(cols, d_out + total, (size_t)total * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
    }
    return rc;
}

extern "C" int lps_is_canonical(lps_handle* h, int* bcols, int* canonical) {
    USE(h);
    if (!canonical) return fail(LPS_E_INVALID, "null result");
    if (h->mode == 1) return lpsx_is_canonical(h, bcols, canonical);
    if (h->nranks != 1) return fail(LPS_E_STATE, "is_canonical on a sharded handle");
    int* flag = h->scratch_i + h->n;
    CU(cudaMemsetAsync(flag, 0, sizeof(int), h->stream));
    k_any_neg_rhs_f64<<<std::max(1, std::min((h->m + 255) / 256, 592)), 256, 0, h->stream>>>(grid_of(h), flag);
    int neg = 0;
    CU(cudaMemcpyAsync(&neg, flag, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    h->launches_all++;
    if (neg) { *canonical = 0; return LPS_OK; }   // bcols untouched, like tableau.py:476-477
    int* bad = h->scratch_i + h->n + 8;
    const int slabs = std::max(1, std::min((h->m + 63) / 64, 64));
    const int rows_per_slab = (h->m + slabs - 1) / slabs;
    k_canon_init_f64<<<(h->n + 255) / 256, 256, 0, h->stream>>>(grid_of(h), h->scratch_i, bad);
    k_canon_cols_f64<<<dim3((h->n + 127) / 128, slabs), 128, 0, h->stream>>>(grid_of(h), rows_per_slab, h->scratch_i, bad);
    DevBuf b_bcols;
    CU(b_bcols.alloc((size_t)h->m * sizeof(int)));
    int* d_bcols = b_bcols.as<int>();
    CU(cudaMemsetAsync(d_bcols, 0x7f, (size_t)h->m * sizeof(int), h->stream));
    k_canon_rows<<<std::max(1, std::min((h->n + 255) / 256, 592)), 256, 0, h->stream>>>(h->scratch_i, bad, h->n, d_bcols);
    h->launches_all += 3;
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

// ---- the solve loop ------------------------------------------------------------------------------
extern "C" int lps_solve_begin(lps_handle* h, long long max_pivots, const int* basis, const int* marks, long long trace_cap) {
    USE(h);
    NEED_MODE(h, 0);
    if (h->solving) return fail(LPS_E_STATE, "solve already in progress");
    if (basis) CU(cudaMemcpyAsync(h->basis, basis, (size_t)h->m_total * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    h->use_marks = marks != nullptr;
    if (marks) CU(cudaMemcpyAsync(h->marks, marks, (size_t)h->n * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    if (trace_cap > h->trace_cap) {
        if (h->trace) CU(cudaFree(h->trace));
        h->trace = nullptr;
        h->trace_cap = 0;
        CU(cudaMalloc(&h->trace, (size_t)trace_cap * 3 * sizeof(int)));
        h->trace_cap = trace_cap;
    }
    double zcell = 0.0;
    CU(cudaMemcpyAsync(&zcell, h->T + (size_t)h->m * h->ld + h->n, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    TRY(read_state(h));
    SolveState s{};
    s.status = LPS_OK; s.rule = LPS_RULE_DANTZIG; s.stuck = 0; s.z_entry_cell = zcell; s.npiv = 0;
    s.max_pivots = max_pivots; s.pending_zcheck = 0; s.ticket = 0; s.dry_run = 0; s.last_r = -1; s.last_c = -1;
    s.xseq = h->st_host->xseq;
    TRY(write_state(h, s));
    h->pp_valid = false;              // both loops re-read the objective row on their first step
    h->blk_valid = false;
    h->solving = true;
    return LPS_OK;
}
extern "C" int lps_step_select(lps_handle* h) {
    USE(h);
    NEED_MODE(h, 0);
    if (!h->solving) return fail(LPS_E_STATE, "no solve in progress");
    return launch_select(h);
}
extern "C" int lps_step_update(lps_handle* h) {
    USE(h);
    NEED_MODE(h, 0);
    if (!h->solving) return fail(LPS_E_STATE, "no solve in progress");
    return launch_update(h, 1);
}
// One block of the blocked loop: the panel decides up to lps_block_size(h) pivots (fewer at optimality, at the end of
// a pivot prefix, ...), the fused update applies them.  With the peer-store exchange attached the panel does the
// per-pivot argmin all-reduce and pivot-row broadcast itself.
extern "C" int lps_step_block(lps_handle* h) {
    USE(h);
    NEED_MODE(h, 0);
    if (!h->solving) return fail(LPS_E_STATE, "no solve in progress");
    if (h->blk_k < 1) return fail(LPS_E_STATE, "the blocked loop is switched off (LPS_F64_BLOCK=0)");
    if (h->nranks > 1 && !h->use_xchg) return fail(LPS_E_STATE, "blocks on a sharded handle need the peer-store exchange (lps_p2p_attach)");
    return launch_block(h);
}
extern "C" int lps_block_size(lps_handle* h) { return (h && h->mode == 0) ? h->blk_k : 0; }
extern "C" int lps_set_block_size(lps_handle* h, int k) {
    USE(h);
    NEED_MODE(h, 0);
    if (k < 0 || k > kBlkMax) return fail(LPS_E_INVALID, "block size %d outside 0..%d", k, kBlkMax);
    if (h->solving) return fail(LPS_E_STATE, "solve in progress");
    h->blk_k = k;
    if (h->gexec) { cudaGraphExecDestroy(h->gexec); h->gexec = nullptr; }
    return LPS_OK;
}
// Launch shape of the look-ahead panel: 1 = one CTA of 1024 threads (barriers are __syncthreads), >= 2 = that many
// CTAs of 256 threads synchronising through memory (never more than SMs: they spin on each other).
extern "C" int lps_set_panel_ctas(lps_handle* h, int ctas) {
    USE(h);
    NEED_MODE(h, 0);
    if (h->solving) return fail(LPS_E_STATE, "solve in progress");
    int sms = 148;
    CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device));
    if (ctas == -8 || ctas == -16) {   // one cluster of 8 / 16 CTAs with the panel in distributed shared memory
        const int C = -ctas, rpc = (h->m + C - 1) / C, cpc = (int)(((h->ld + C - 1) / C + 1) / 2 * 2);
        if (blk_cluster_smem(rpc, cpc) > (size_t)200 * 1024) return fail(LPS_E_INVALID, "the panel of a %d x %d tableau does not fit the shared memory of %d CTAs", h->m, h->n, C);
        h->pan_cluster = C; h->pan_rpc = rpc; h->pan_cpc = cpc;
        h->blk_valid = false;
        if (h->gexec) { cudaGraphExecDestroy(h->gexec); h->gexec = nullptr; }
        return LPS_OK;
    }
    if (ctas < 1 || ctas > std::min(sms, 256)) return fail(LPS_E_INVALID, "panel CTAs %d outside 1..%d", ctas, std::min(sms, 256));
    h->pan_cluster = 0;
    h->pan_blocks = ctas;
    h->pan_threads = (ctas == 1) ? kSelMaxThreads : kThreads;
    h->blk_valid = false;
    if (h->gexec) { cudaGraphExecDestroy(h->gexec); h->gexec = nullptr; }
    return LPS_OK;
}
extern "C" int lps_solve_poll(lps_handle* h, int* status, long long* npiv) {
    USE(h);
    NEED_MODE(h, 0);
    TRY(read_state(h));
    if (status) *status = h->st_host->status;
    if (npiv) *npiv = h->st_host->npiv;
    return LPS_OK;
}
extern "C" int lps_solve_end(lps_handle* h, int* basis, int* marks, int* trace, long long trace_cap, long long* npiv) {
    USE(h);
    NEED_MODE(h, 0);
    if (!h->solving) return fail(LPS_E_STATE, "no solve in progress");
    TRY(read_state(h));
    h->solving = false;
    const long long k = h->st_host->npiv;
    if (npiv) *npiv = k;
    if (basis) CU(cudaMemcpyAsync(basis, h->basis, (size_t)h->m_total * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    if (marks && h->use_marks) CU(cudaMemcpyAsync(marks, h->marks, (size_t)h->n * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    if (trace && trace_cap > 0) {
        const long long cnt = std::min(std::min(k, trace_cap), h->trace_cap);
        if (cnt > 0) CU(cudaMemcpyAsync(trace, h->trace, (size_t)cnt * 3 * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    }
    CU(cudaStreamSynchronize(h->stream));
    if (h->st_host->status == LPS_E_PEER) {
        static const char* what[] = {"?", "the panel's first grid-wide synchronisation", "a peer's candidate record", "the winner's slice record",
                                     "the winner's row slice (its checksum never agreed)", "the panel's end-of-pivot synchronisation"};
        const int k = -100 - h->st_host->last_r;
        return fail(LPS_E_PEER, "the blocked loop gave up after %lld pivots waiting for %s", (long long)h->st_host->npiv, what[(k >= 1 && k <= 5) ? k : 0]);
    }
    return h->st_host->status;
}

// (select, update) x burst on the handle's stream as one executable graph; re-captured when an argument baked into
// the nodes changed (trace buffer, marks, exchange mode, launch plan)
static int ensure_burst_graph(lps_handle* h, long long burst) {
    unsigned long long sig = 1469598103934665603ull;
    auto mix = [&](unsigned long long v) { sig = (sig ^ v) * 1099511628211ull; };
    mix((unsigned long long)(uintptr_t)h->trace); mix((unsigned long long)h->trace_cap); mix(h->use_marks ? 1 : 0);
    mix((unsigned long long)(uintptr_t)h->basis); mix(h->use_xchg ? 1 : 0); mix((unsigned long long)h->chunk);
    mix((unsigned long long)h->rows_per_cta); mix((unsigned long long)h->sel_blocks); mix((unsigned long long)h->sel_threads);
    mix((unsigned long long)(uintptr_t)h->cand_send); mix((unsigned long long)(uintptr_t)h->cand_all); mix((unsigned long long)burst);
    mix((unsigned long long)h->blk_k); mix((unsigned long long)h->pan_blocks); mix((unsigned long long)h->pan_cluster);
    mix((unsigned long long)h->n); mix((unsigned long long)(uintptr_t)h->T);
    if (h->gexec && h->gsig == sig && h->gburst == burst) return LPS_OK;
    if (h->gexec) { cudaGraphExecDestroy(h->gexec); h->gexec = nullptr; }
    const long long la = h->launches_all, lu = h->launches_update;
    cudaGraph_t graph = nullptr;
    CU(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
    int rc = LPS_OK;
    for (long long k = 0; k < burst && rc == LPS_OK; ++k) {
        if (h->blk_k >= 1) rc = launch_block(h);
        else {
            rc = launch_select(h);
            if (rc == LPS_OK) rc = launch_update(h, 1);
        }
    }
    cudaError_t e = cudaStreamEndCapture(h->stream, &graph);
    h->launches_all = la; h->launches_update = lu;               // counted when the graph is launched
    if (rc != LPS_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (e != cudaSuccess) return fail(LPS_E_CUDA, "stream capture failed: %s", cudaGetErrorString(e));
    e = cudaGraphInstantiate(&h->gexec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) { h->gexec = nullptr; return fail(LPS_E_CUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(e)); }
    h->gsig = sig; h->gburst = burst;
    return LPS_OK;
}

extern "C" int lps_solve(lps_handle* h, long long max_pivots, int* basis, int* marks, int* trace, long long trace_cap, long long* npiv) {
    USE(h);
    if (h->nranks != 1) return fail(LPS_E_STATE, "lps_solve on a sharded handle: drive it with the stepping calls");
    if (!basis) return fail(LPS_E_INVALID, "null basis");
    if (h->mode == 1) return lpsx_solve(h, max_pivots, basis, marks, trace, trace_cap, npiv);
    TRY(lps_solve_begin(h, max_pivots, basis, marks, trace ? trace_cap : 0));
    const double est_us = std::max(4.0, (double)(h->m + 1) * (double)h->ld * 16.0 / 5.0e6);   // one pass over the grid
    // Launch-bound sizes: a burst is captured once as a CUDA graph and replayed.  Launches enqueued past optimality or
    // past the pivot prefix are no-ops on the device (the kernels test the state first), so whole bursts are always
    // safe.  Per-launch profiling events need the plain launches.
    const bool graphed = !h->prof && est_us < 200.0;
    if (h->blk_k >= 1) {
        // blocked loop: one panel launch decides up to blk_k pivots, one pass over the grid applies them; about 2 ms of
        // device work between two looks at the state
        const double blk_us = est_us + 8.0 * h->blk_k;
        const long long burst = (long long)std::max(1.0, std::min(64.0, 2000.0 / blk_us));
        const long long gburst = std::min<long long>(burst, 8);
        TRY(blk_sync_panel(h));           // the lazy re-read of the objective row stays outside the captured burst
        if (graphed) TRY(ensure_burst_graph(h, gburst));
        long long enq_blocks = 0;
        for (;;) {
            if (graphed) {
                for (long long k = 0; k < burst; k += gburst) CU(cudaGraphLaunch(h->gexec, h->stream));
                const long long done = (burst + gburst - 1) / gburst * gburst;
                h->launches_all += 2 * done;
                h->launches_update += done;
                enq_blocks += done;
            } else {
                long long now = burst;
                if (max_pivots >= 0) now = std::max<long long>(1, std::min(burst, (max_pivots - enq_blocks * h->blk_k + h->blk_k - 1) / h->blk_k));
                for (long long k = 0; k < now; ++k) TRY(launch_block(h));
                enq_blocks += now;
            }
            TRY(read_state(h));
            if (h->st_host->status != LPS_OK) break;
        }
        return lps_solve_end(h, basis, marks, trace, trace_cap, npiv);
    }
    // rank-1 loop (LPS_F64_BLOCK=0): pivots enqueued per host check: about 2 ms of device work, between 4 and 256; a
    // pivot prefix never enqueues more (select, update) pairs than it may commit
    const long long burst = (long long)std::max(4.0, std::min(256.0, 2000.0 / est_us));
    long long enqueued = 0;
    const long long gburst = std::min<long long>(burst, 32);   // short graphs: capture stays cheap next to a short solve
    if (!h->pp_valid) TRY(launch_price(h));   // (outside the captured burst)
    if (graphed) TRY(ensure_burst_graph(h, gburst));
    for (;;) {
        if (graphed) {
            for (long long k = 0; k < burst; k += gburst) CU(cudaGraphLaunch(h->gexec, h->stream));
            const long long done = (burst + gburst - 1) / gburst * gburst;
            h->launches_all += 2 * done;
            h->launches_update += done;
            enqueued += done;
        } else {
            const long long room = max_pivots >= 0 ? max_pivots - enqueued : burst;
            const long long now = std::min(burst, room);
            for (long long k = 0; k < now; ++k) {
                TRY(launch_select(h));
                TRY(launch_update(h, 1));
            }
            enqueued += now;
            if (now == 0) TRY(launch_select(h));   // lets the device report LPS_MAX_PIVOTS (or optimality)
        }
        TRY(read_state(h));
        if (h->st_host->status != LPS_OK) break;
    }
    return lps_solve_end(h, basis, marks, trace, trace_cap, npiv);
}

// ------------------------------------------------------------------------------------------------
// peer-store exchange over CUDA IPC
// ------------------------------------------------------------------------------------------------
// Sizes this handle's exchange block for `nranks` ranks (every rank stores its candidate rows into every block).
// Call it before lps_p2p_export / lps_p2p_block: it may move the block.
extern "C" int lps_p2p_prepare(lps_handle* h, int nranks) {
    USE(h);
    NEED_MODE(h, 0);
    if (nranks < 1 || nranks > kMaxRanks) return fail(LPS_E_INVALID, "bad rank count %d", nranks);
    if (h->solving) return fail(LPS_E_STATE, "solve in progress");
    if (!h->ipc_open.empty()) return fail(LPS_E_STATE, "peers are already attached to this handle");
    return alloc_xchg(h, nranks);
}

extern "C" int lps_p2p_export(lps_handle* h, unsigned char* ipc_handle64) {
    USE(h);
    NEED_MODE(h, 0);
    if (!ipc_handle64) return fail(LPS_E_INVALID, "null handle buffer");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
    cudaIpcMemHandle_t ih;
    CU(cudaIpcGetMemHandle(&ih, h->xchg));
    memcpy(ipc_handle64, &ih, 64);
    return LPS_OK;
}

static int attach_blocks(lps_handle* h, long long row0, long long m_total, int rank, int nranks, const std::vector<XchgBlock*>& peers) {
    CU(cudaMemcpy(h->d_peers, peers.data(), (size_t)nranks * sizeof(XchgBlock*), cudaMemcpyHostToDevice));
    // The peer-store protocol counts select rounds from the same number on every rank, over zeroed headers: a
    // (pooled) handle that already ran pivots or dry-run selects starts over.  The callers put a cross-rank barrier
    // between attaching and the first pivot (sharded.py: the all-reduce of the "mapped" flags), so no peer can
    // have stored into this block yet.
    TRY(read_state(h));
    SolveState s = *h->st_host;
    s.xseq = 0; s.ticket = 0; s.ticket2 = 0;
    TRY(write_state(h, s));
    CU(cudaMemsetAsync(h->xchg, 0, sizeof(XchgBlock), h->stream));
    CU(cudaMemsetAsync(h->blk.sync, 0, sizeof(BlkSync), h->stream));
    CU(cudaStreamSynchronize(h->stream));
    if (h->gexec) { cudaGraphExecDestroy(h->gexec); h->gexec = nullptr; }
    h->row0 = row0; h->m_total = m_total; h->rank = rank; h->nranks = nranks;
    h->use_xchg = (nranks > 1);          // one rank: k_select writes the normalised row itself, no exchange
    if (nranks > 1) {
        // the panel's slices of a candidate row are exchanged CTA by CTA: every rank must cut the row the same way, so
        // the launch shape follows the LARGEST shard (shards differ by a row), not this handle's own rows
        const int mm = (int)((m_total + nranks - 1) / nranks);
        h->pan_threads = kThreads;
        h->pan_blocks = std::max(2, std::min(std::max((mm + kThreads - 1) / kThreads, (int)((h->ld / 2 + kThreads - 1) / kThreads)), std::min(h->sms, 256)));
        h->pan_cluster = 0;
    }
    CU(cudaFree(h->basis));
    CU(cudaMalloc(&h->basis, (size_t)m_total * sizeof(int)));
    CU(cudaMemsetAsync(h->basis, 0xff, (size_t)m_total * sizeof(int), h->stream));
    return LPS_OK;
}

static int check_shard(lps_handle* h, long long row0, long long m_total, int rank, int nranks) {
    if (nranks < 1 || nranks > kMaxRanks || rank < 0 || rank >= nranks) return fail(LPS_E_INVALID, "bad rank %d of %d", rank, nranks);
    if (h->blk.nr != nranks) return fail(LPS_E_STATE, "exchange block sized for %d ranks, not %d: call lps_p2p_prepare first", h->blk.nr, nranks);
    if (row0 < 0 || row0 + h->m > m_total) return fail(LPS_E_INVALID, "rows [%lld,%lld) outside 0..%lld", row0, row0 + h->m, m_total);
    return LPS_OK;
}

extern "C" int lps_p2p_attach(lps_handle* h, long long row0, long long m_total, int rank, int nranks, const unsigned char* ipc_handles) {
    USE(h);
    NEED_MODE(h, 0);
    TRY(check_shard(h, row0, m_total, rank, nranks));
    if (nranks > 1 && !ipc_handles) return fail(LPS_E_INVALID, "null IPC handles");
    CU(cudaStreamSynchronize(h->stream));
    std::vector<XchgBlock*> peers(nranks, nullptr);
    for (int k = 0; k < nranks; ++k) {
        if (k == rank) { peers[k] = h->xchg; continue; }
        cudaIpcMemHandle_t ih;
        memcpy(&ih, ipc_handles + (size_t)k * 64, 64);
        void* p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, ih, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) return fail(LPS_E_CUDA, "cudaIpcOpenMemHandle for rank %d failed: %s", k, cudaGetErrorString(e));
        h->ipc_open.push_back(p);
        peers[k] = (XchgBlock*)p;
    }
    return attach_blocks(h, row0, m_total, rank, nranks, peers);
}

// Same exchange for ranks that live in ONE process (one thread driving several handles, on one or several GPUs):
// the exchange blocks are passed as device pointers (lps_p2p_block of every rank, in rank order) instead of IPC
// handles; peer access between the devices involved is enabled here.
extern "C" int lps_p2p_block(lps_handle* h, void** block) {
    USE(h);
    NEED_MODE(h, 0);
    if (!block) return fail(LPS_E_INVALID, "null output");
    *block = h->xchg;
    return LPS_OK;
}
extern "C" int lps_p2p_attach_local(lps_handle* h, long long row0, long long m_total, int rank, int nranks, void* const* blocks) {
    USE(h);
    NEED_MODE(h, 0);
    TRY(check_shard(h, row0, m_total, rank, nranks));
    if (!blocks) return fail(LPS_E_INVALID, "null blocks");
    CU(cudaStreamSynchronize(h->stream));
    std::vector<XchgBlock*> peers(nranks, nullptr);
    for (int k = 0; k < nranks; ++k) {
        if (!blocks[k]) return fail(LPS_E_INVALID, "null block for rank %d", k);
        cudaPointerAttributes at;
        CU(cudaPointerGetAttributes(&at, blocks[k]));
        if (at.device != h->device) {
            cudaError_t e = cudaDeviceEnablePeerAccess(at.device, 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
            else if (e != cudaSuccess) return fail(LPS_E_CUDA, "no peer access from device %d to %d: %s", h->device, at.device, cudaGetErrorString(e));
        }
        peers[k] = (XchgBlock*)blocks[k];
    }
    if (peers[rank] != h->xchg) return fail(LPS_E_INVALID, "blocks[%d] is not this handle's block", rank);
    return attach_blocks(h, row0, m_total, rank, nranks, peers);
}

// ------------------------------------------------------------------------------------------------
// profiling
// ------------------------------------------------------------------------------------------------
extern "C" int lps_prof_reset(lps_handle* h) {
    USE(h);
    if (h->ev.empty()) {
        h->ev.resize(2 * 8192);
        for (auto& e : h->ev) CU(cudaEventCreate(&e));
    }
    h->ev_used = 0;
    h->launches_update = 0;
    h->launches_all = 0;
    h->prof = true;
    return LPS_OK;
}
extern "C" int lps_prof_read(lps_handle* h, double* update_ms, long long* update_launches, long long* all_launches) {
    USE(h);
    CU(cudaStreamSynchronize(h->stream));
    double total = 0.0;
    for (size_t k = 0; k + 1 < h->ev_used; k += 2) {
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, h->ev[k], h->ev[k + 1]));
        total += ms;
    }
    if (update_ms) *update_ms = total;
    if (update_launches) *update_launches = (long long)(h->ev_used / 2);
    if (all_launches) *all_launches = h->launches_all;
    return LPS_OK;
}
