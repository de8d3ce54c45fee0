// host_state.inl -- handle structs (index, lanes, prepared filters, exchange), error / allocation helpers, capacity growth.
// Part of the single translation unit mlv_index.cu (included there, in order).
#pragma once

namespace {

struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
};
struct HostBuf {
    void* p = nullptr;
    size_t bytes = 0;
};

// Scratch of one in-flight search.  A handle keeps MLV_LANES of them, picked by the stream a search
// is enqueued on, so searches on different streams overlap on the GPU (the next query's scan fills
// the SMs the previous query's straggling / exchanging CTAs have left).
constexpr int MLV_LANES = 4;
struct Lane {
    cudaStream_t stream = nullptr;
    bool used = false;
    uint64_t last_use = 0;
    DevBuf d_q, d_keys0, d_keys1, d_sched;
    DevBuf d_flist, d_fscratch;  // gather list built from a per-call filter bitmap
    DevBuf d_cert;               // shadow scan: the word its conditional fp32 launch reads (scan_kernel.cuh, ScanParams::cert)
    uint64_t seen_maint = 0;     // mlv_index::maint_gen this stream has been ordered behind
};

// One asynchronous host-buffer search in flight (mlv_index_submit / mlv_index_collect): its own
// stream (hence its own scratch lane), pinned staging, device query / result blocks and a done event.
constexpr int MLV_ASYNC_SLOTS = 4;
struct AsyncSlot {
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr;
    HostBuf stage;
    DevBuf d_q, d_out;
    uint32_t nq = 0, k = 0;
    bool busy = false;
    bool exchange = false;   // an exchange search: at most XCHG_MAX_IN_FLIGHT of those may be in flight (exchange.cuh slot reuse)
};

struct ScanCfg {
    int R, NQ, CW, PW;
    uint32_t T, S, stage_f4;
    size_t smem;
    int grid, threads;
    int evict_first;
};

}  // namespace

struct mlv_exchange {
    int device = 0;
    uint32_t world = 1, rank = 0;
    uint64_t* bufs[XCHG_MAX_WORLD] = {nullptr};  // bufs[rank] = local allocation, others IPC-opened
    int* d_error = nullptr;
    bool connected = false;
    unsigned long long timeout_ns = XCHG_TIMEOUT_NS;
};

struct mlv_filter {
    mlv_index* owner = nullptr;
    DevBuf d_bitmap, d_list, d_scratch;
    uint64_t bitmap_words = 0;   // words the caller supplied (rows appended later do not pass)
    uint64_t passing = 0;        // live AND passing rows when the list was built
    uint64_t epoch = ~0ull;      // owner->epoch the list was built at
    uint64_t compact_gen = 0;    // owner->compact_gen at creation: compaction renumbers rows, the bitmap is void after it
    bool counted = false;        // `passing` has been read back
};

// device buffers of a destroyed prepared filter, kept for the next one: cudaMalloc / cudaFree of the 4-byte-per-row
// list cost 2 - 30 ms each beside a 30 GB matrix, the kernels that fill it well under 1 ms
struct FilterBufs {
    DevBuf bitmap, list, scratch;
};
constexpr size_t MLV_FILTER_POOL = 4;

struct mlv_index {
    int device = 0;
    uint32_t dim = 0, ld = 0;
    int metric = MLV_L2;
    uint64_t rows = 0, capacity = 0, n_deleted = 0, row_base = 0;
    float* d_rows = nullptr;
    uint32_t* d_live = nullptr;
    uint64_t live_words = 0;
    cudaStream_t stream = nullptr;
    int sm_count = 0;
    size_t smem_optin = 0;
    DevBuf d_qraw, d_filter, d_outd, d_outr, d_outc, d_misc, d_range, d_timeline;
    Lane lanes[MLV_LANES];
    uint64_t lane_clock = 0;
    uint64_t epoch = 0;          // bumped by every add / delete / compact / clear: prepared filters rebuild their row list
    mlv_filter* bound_filter = nullptr;  // mlv_index_set_filter
    uint64_t compact_gen = 0;
    int tune_gather = -1;        // -1 auto, 0 never (stream + mask), 1 always when a filter is given
    HostBuf h_stage;
    HostBuf h_range;             // mapped pinned result block of the sharded range search (written by the kernel)
    unsigned int flag_seq = 0;   // completion-flag values of the batch-1 latency path (mlv_index_search)
    int tune_fast_host = 1;      // 0 = mlv_index_search always stages (H2D, preparation launch, D2H, synchronise)
    HostBuf h_upload;            // two pinned chunks for bulk row uploads (upload_rows_staged)
    int tune_staged_upload = 1;  // 0 = plain cudaMemcpy from pageable memory
    AsyncSlot slots[MLV_ASYNC_SLOTS];
    uint32_t next_slot = 0;
    std::string err;
    bool timing = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending;
    std::vector<cudaEvent_t> event_pool;
    uint64_t launches = 0;
    // tuning (mlv_index_set_tuning / MLV_SCAN_* environment)
    int tune_cw = 0, tune_stage_kb = 0, tune_evict_first = -1, tune_r = 0, tune_max_stages = 8, tune_ctas = 0, tune_pw = 0;
    int tune_timeline = 0;
    int last_grid = 0;
    // dynamic tile scheduler + fused final select (scan_kernel.cuh tail); counters live in the lanes
    int tune_dynamic = 1, tune_tile_batch = 4, tune_fused = 1;
    // fused multi-GPU exchange (exchange.cuh)
    mlv_exchange* xchg = nullptr;
    uint64_t xchg_row_bases[XCHG_MAX_WORLD] = {0};
    uint64_t xseq = 0;
    // tensor-core batch path (gemm_kernel.cuh)
    DevBuf d_norms, d_gq, d_cand, d_maxn2, d_sub, d_gx;
    uint64_t norms_valid = 0;  // rows [0, norms_valid) of d_norms are current
    DevBuf d_rows16, d_f16st, d_gx16;  // fp16 shadow of the rows (HALF tier), its frozen scale / overflow flag, halves of a gathered copy
    uint64_t f16_valid = 0;    // rows [0, f16_valid) of the shadow are current
    bool f16_overflowed = false;  // the last fp16-tier batch hit the overflow flag
    uint64_t gemm_half_queries = 0;  // queries certified by the fp16 tier
    // shadow scan (single queries over the fp16 shadow, scan_kernel_half): -1 auto, 0 never, 1 whenever the shape allows
    int tune_scan_half = -1;
    int tune_scan_half_gather = 1;   // gathered (selectively filtered) single queries may read the shadow too (0 = fp32 rows)
    int tune_scan_half_mma = 1;      // tensor-core consumers when the shadow's rows are whole 128-byte chunks (0 = FMA consumers)
    DevBuf d_half_stats;             // device counters {queries, not certified}
    HostBuf h_half_stats;            // their pinned mirror {queries, not certified, overflow flag}, written by the kernel
    uint32_t half_seen_q = 0, half_seen_u = 0;   // mirror values at the last policy check
    uint32_t half_skip = 0, half_backoff = 0;    // searches the shadow scan sits out after certifying too little
    uint64_t half_scan_launches = 0;
    // row norms and the fp16 shadow are (re)built on whichever stream first needs them; a search on ANOTHER stream that
    // finds them "valid" must still be ordered behind that work
    cudaEvent_t maint_event = nullptr;
    uint64_t maint_gen = 0;
    cudaStream_t maint_stream = nullptr;
    uint64_t range_hits_hint = 0;                // largest hit count of the last host range search
    bool half_stepped = false;                   // the latency-path probe of this search already stepped the policy
    int tune_gemm = -1;        // -1 auto, 0 never, 1 whenever the shape allows it
    int tune_gemm_min_nq = 0;  // 0 = auto (gemm_min_nq: 5 with the one-pass tier on a >= 1 GB matrix, else 9)
    int tune_gemm_bn = 0;      // queries per GEMM tile: 0 auto, or 64 / 128 / 256
    int tune_gemm_debug = 0;   // profiling only: bit 0 = epilogue drains nothing (wrong results)
    int tune_gemm_wide = 3;    // one-pass tiers of wide batches: 0 = single-tile kernel, anything else = CTA pairs (cta_group::2; fastest on every shape measured)
    int tune_gemm_passes = 0;  // 0 auto (fp16 shadow tier, then 3xTF32 for what it cannot certify), 1 one-pass TF32 tier then scan, 2 fp16 tier then scan, 3 3xTF32 only
    uint64_t gemm_fast_queries = 0;  // queries certified by the one-pass tier
    uint64_t gemm_gathered_searches = 0;  // filtered batches that multiplied a compacted copy of the passing rows
    uint32_t gemm_fast_skip = 0;     // batches the one-pass tier sits out (it certified too little last time)
    int tune_gemm_predict = 1;       // one-pass tiers: thresholds predicted from the rows seen so far (refine_kernel); 0 = the plain k'-th-best rule
    uint32_t gemm_predict_skip = 0, gemm_predict_backoff = 0;   // batches predictions sit out after failing (rows stored in an order that correlates with the queries); doubles per failure
    uint64_t gemm_mispredicted_queries = 0;   // queries whose predicted thresholds failed the final check (answered by the next tier)
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> gemm_pending;
    uint64_t gemm_searches = 0, gemm_queries = 0, gemm_fallback_queries = 0, gemm_rounds = 0, gemm_launches = 0;
    // columnar metadata (column_kernels.cuh): int32 code columns, allocated on first use
    std::vector<FilterBufs> filter_pool;
    DevBuf d_cols[MLV_MAX_COLUMNS];
    uint64_t col_rows[MLV_MAX_COLUMNS] = {0};  // rows each allocation covers (<= capacity; grown lazily)
};

namespace {

struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) ok = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard() {
        int cur = -1;
        if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
    }
};

int fail(mlv_index* h, int status, const std::string& msg) {
    if (h) h->err = msg;
    return status;
}
int fail_cuda(mlv_index* h, cudaError_t e, const char* what) {
    std::string m = std::string(what) + ": " + cudaGetErrorString(e);
    cudaGetLastError();  // clear sticky-less error state
    return fail(h, e == cudaErrorMemoryAllocation ? MLV_E_NOMEM : MLV_E_CUDA, m);
}
#define CK(h, call)                                        \
    do {                                                   \
        cudaError_t e__ = (call);                          \
        if (e__ != cudaSuccess) return fail_cuda(h, e__, #call); \
    } while (0)

int ensure_dev(mlv_index* h, DevBuf& b, size_t bytes) {
    if (b.bytes >= bytes) return MLV_OK;
    if (b.p) CK(h, cudaFree(b.p));
    b.p = nullptr;
    b.bytes = 0;
    size_t want = std::max(bytes, (size_t)4096);
    CK(h, cudaMalloc(&b.p, want));
    b.bytes = want;
    return MLV_OK;
}
int ensure_host(mlv_index* h, HostBuf& b, size_t bytes) {
    if (b.bytes >= bytes) return MLV_OK;
    if (b.p) CK(h, cudaFreeHost(b.p));
    b.p = nullptr;
    b.bytes = 0;
    size_t want = std::max(bytes, (size_t)4096);
    CK(h, cudaMallocHost(&b.p, want));
    b.bytes = want;
    return MLV_OK;
}
void free_dev(DevBuf& b) {
    if (b.p) cudaFree(b.p);
    b.p = nullptr;
    b.bytes = 0;
}

uint32_t pow2_ceil(uint32_t v) {
    uint32_t p = 1;
    while (p < v) p <<= 1;
    return p;
}

int env_int(const char* name, int dflt) {
    const char* s = getenv(name);
    return (s && *s) ? atoi(s) : dflt;
}

// ---- capacity ----------------------------------------------------------------------------------
int reserve_rows(mlv_index* h, uint64_t need) {
    if (need <= h->capacity) return MLV_OK;
    if (need >= 0xFFFFFFFEull) return fail(h, MLV_E_UNSUPPORTED, "more than 2^32-2 rows in one index shard");
    uint64_t cap = std::max<uint64_t>({need, h->capacity * 2, 1024});
    cap = (cap + 31) & ~31ull;
    float* nrows = nullptr;
    uint32_t* nlive = nullptr;
    const size_t row_bytes = (size_t)h->ld * 4;
    cudaError_t e = cudaMalloc(&nrows, cap * row_bytes);
    if (e != cudaSuccess && cap > need) {  // doubling did not fit: take exactly what is needed
        cudaGetLastError();
        cap = (need + 31) & ~31ull;
        e = cudaMalloc(&nrows, cap * row_bytes);
    }
    if (e != cudaSuccess) return fail_cuda(h, e, "cudaMalloc(row matrix)");
    const uint64_t words = cap / 32;
    e = cudaMalloc(&nlive, words * 4);
    if (e != cudaSuccess) {
        cudaFree(nrows);
        return fail_cuda(h, e, "cudaMalloc(live bitmap)");
    }
    // zero: padding columns must read as 0 forever, unused rows' bits as "not live"
    e = cudaMemsetAsync(nrows, 0, cap * row_bytes, h->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(nlive, 0, words * 4, h->stream);
    if (e == cudaSuccess && h->rows) {
        e = cudaMemcpyAsync(nrows, h->d_rows, h->rows * row_bytes, cudaMemcpyDeviceToDevice, h->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(nlive, h->d_live, ((h->rows + 31) / 32) * 4, cudaMemcpyDeviceToDevice, h->stream);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    if (e != cudaSuccess) {   // the old matrix stays in place; the new buffers go back
        cudaFree(nrows);
        cudaFree(nlive);
        return fail_cuda(h, e, "growing the row matrix");
    }
    if (h->d_rows) cudaFree(h->d_rows);
    if (h->d_live) cudaFree(h->d_live);
    h->d_rows = nrows;
    h->d_live = nlive;
    h->capacity = cap;
    h->live_words = words;
    return MLV_OK;
}

// normalize = false: the rows are already in stored form (snapshot import)
int finish_append(mlv_index* h, uint64_t n, uint64_t* first_row, bool normalize = true) {
    const uint64_t first = h->rows;
    if (h->metric == MLV_COSINE && normalize) {
        const int wpb = 8;
        normalize_rows_kernel<<<(unsigned)((n + wpb - 1) / wpb), wpb * 32, 0, h->stream>>>(h->d_rows, first, n, h->ld);
        h->launches++;
    }
    {
        const uint64_t words = ((first + n - 1) >> 5) - (first >> 5) + 1;
        set_live_range_kernel<<<(unsigned)std::min<uint64_t>((words + 255) / 256, 4096), 256, 0, h->stream>>>(h->d_live, first, n);
        h->launches++;
    }
    CK(h, cudaGetLastError());
    CK(h, cudaStreamSynchronize(h->stream));
    h->rows += n;
    h->epoch++;
    if (first_row) *first_row = first;
    return MLV_OK;
}


}  // namespace
