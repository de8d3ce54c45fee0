// mlv_index.cu -- C ABI (include/mlv_index.h) and host-side orchestration of the B200 exact-search
// index.  Everything the reference does through hnswlib inside
// src/mlvectordb/implementations/index.py (add_items :65, mark_deleted :80, knn_query :111) lands
// here.  No CPU fallback: without a CUDA device mlv_index_create fails with MLV_E_NO_DEVICE.
#include "../../include/mlv_index.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <utility>
#include <vector>

#include "common.cuh"
#include "gemm_kernel.cuh"
#include "maint_kernels.cuh"
#include "scan_kernel.cuh"
#include "select_kernel.cuh"

using namespace mlv;

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
    DevBuf d_norms, d_gq, d_cand, d_maxn2;
    uint64_t norms_valid = 0;  // rows [0, norms_valid) of d_norms are current
    int tune_gemm = -1;        // -1 auto, 0 never, 1 whenever the shape allows it
    int tune_gemm_min_nq = 32;
    int tune_gemm_bn = 256;    // queries per GEMM tile: 256 (2-stage ring) or 128 (3-stage ring)
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> gemm_pending;
    uint64_t gemm_searches = 0, gemm_queries = 0, gemm_fallback_queries = 0, gemm_rounds = 0, gemm_launches = 0;
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
    CK(h, cudaMemsetAsync(nrows, 0, cap * row_bytes, h->stream));
    CK(h, cudaMemsetAsync(nlive, 0, words * 4, h->stream));
    if (h->rows) {
        CK(h, cudaMemcpyAsync(nrows, h->d_rows, h->rows * row_bytes, cudaMemcpyDeviceToDevice, h->stream));
        CK(h, cudaMemcpyAsync(nlive, h->d_live, ((h->rows + 31) / 32) * 4, cudaMemcpyDeviceToDevice, h->stream));
    }
    CK(h, cudaStreamSynchronize(h->stream));
    if (h->d_rows) cudaFree(h->d_rows);
    if (h->d_live) cudaFree(h->d_live);
    h->d_rows = nrows;
    h->d_live = nlive;
    h->capacity = cap;
    h->live_words = words;
    return MLV_OK;
}

int finish_append(mlv_index* h, uint64_t n, uint64_t* first_row) {
    const uint64_t first = h->rows;
    if (h->metric == MLV_COSINE) {
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

// ---- scan configuration ------------------------------------------------------------------------
// Shape of one scan launch.  Defaults come from B200 sweeps (profiles/r01_sweep_*.jsonl): short rows
// are latency-bound in the consumers, so they get more warps, more rows per warp step and larger
// stages; a gathering producer is bound by its copy issue rate (~80 cycles per row copy per warp),
// so short rows get more producer warps.
int choose_cfg(mlv_index* h, uint32_t nq, uint32_t k, bool range, ScanCfg* c, bool gather = false) {
    const uint32_t ld4 = h->ld / 4;
    const size_t rowbytes = (size_t)h->ld * 4;
    int CW = h->tune_cw > 0 ? std::min(h->tune_cw, SCAN_MAX_CW) : (ld4 <= 64 ? 16 : 8);
    int R = h->tune_r ? h->tune_r : (ld4 <= 256 ? 4 : (ld4 <= 512 ? 2 : 1));
    if (R != 1 && R != 2 && R != 4) R = 1;
    int NQ = 1;
    if (!range) {
        while (NQ < 8 && (uint32_t)NQ < nq) NQ <<= 1;
        while (NQ > 1 && ((size_t)NQ * k * CW * 8 > 65536 || R * NQ > 32)) NQ >>= 1;
    }
    const int max_stages = std::min(std::max(h->tune_max_stages, 2), 16);
    const size_t fixed = (size_t)NQ * rowbytes + (range ? 0 : (size_t)CW * NQ * k * 8) + (size_t)max_stages * 24 + 256;
    if (fixed + 2 * rowbytes > h->smem_optin)
        return fail(h, MLV_E_UNSUPPORTED, "dimension too large for the shared-memory ring of this build");
    const size_t avail = h->smem_optin - fixed;
    const size_t group_bytes = (size_t)R * CW * rowbytes;
    const size_t target = (size_t)(h->tune_stage_kb > 0 ? h->tune_stage_kb : (ld4 <= 32 ? 64 : 32)) * 1024;
    uint64_t m = std::max<uint64_t>(1, target / group_bytes);
    uint64_t T = (uint64_t)R * CW * m;
    if (T * rowbytes * 2 > avail) {
        T = (avail / 2 / rowbytes) / R * R;
        if (T == 0) {
            R = 1;
            T = avail / 2 / rowbytes;
        }
        if (T == 0) return fail(h, MLV_E_UNSUPPORTED, "dimension too large for the shared-memory ring of this build");
    }
    const size_t stage = T * rowbytes;
    if (stage >= (1u << 20)) return fail(h, MLV_E_UNSUPPORTED, "ring stage exceeds the mbarrier tx-count range");
    uint32_t S = (uint32_t)std::min<size_t>(max_stages, avail / stage);
    int PW = 1;
    if (gather) PW = h->tune_pw > 0 ? std::min(h->tune_pw, SCAN_MAX_PW) : (rowbytes >= 2048 ? 1 : (rowbytes >= 1024 ? 2 : 4));
    while (PW > 1 && (uint32_t)PW > S) PW >>= 1;
    if (PW == 3) PW = 2;
    S = S / PW * PW;
    c->PW = PW;
    c->R = R;
    c->NQ = NQ;
    c->CW = CW;
    c->T = (uint32_t)T;
    c->S = S;
    c->stage_f4 = (uint32_t)(stage / 16);
    c->smem = (size_t)S * stage + (size_t)NQ * rowbytes + (range ? 0 : (size_t)CW * NQ * k * 8) + (size_t)S * 24;
    const uint64_t n_tiles = (h->rows + T - 1) / T;
    const int ctas = h->tune_ctas > 0 ? h->tune_ctas : h->sm_count;
    c->grid = (int)std::min<uint64_t>(n_tiles, (uint64_t)ctas);
    if (c->grid < 1) c->grid = 1;
    c->threads = (CW + PW) * 32;
    const size_t bytes = h->rows * rowbytes;
    c->evict_first = h->tune_evict_first >= 0 ? h->tune_evict_first : (bytes > ((size_t)96 << 20) ? 1 : 0);
    return MLV_OK;
}

template <int METRIC, int NQ, int R, bool RANGE>
cudaError_t launch_scan_t(const ScanParams& p, const ScanCfg& c, cudaStream_t st) {
    auto kern = scan_kernel<METRIC, NQ, R, RANGE>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c.smem);
    if (e != cudaSuccess) return e;
    kern<<<c.grid, c.threads, c.smem, st>>>(p);
    return cudaGetLastError();
}

template <int METRIC, bool RANGE>
cudaError_t launch_scan_m(const ScanParams& p, const ScanCfg& c, cudaStream_t st) {
#define MLV_CASE(NQv, Rv) \
    if (c.NQ == NQv && c.R == Rv) return launch_scan_t<METRIC, NQv, Rv, RANGE>(p, c, st);
    MLV_CASE(1, 1) MLV_CASE(1, 2) MLV_CASE(1, 4)
    if (!RANGE) {
        MLV_CASE(2, 1) MLV_CASE(2, 2) MLV_CASE(2, 4)
        MLV_CASE(4, 1) MLV_CASE(4, 2) MLV_CASE(4, 4)
        MLV_CASE(8, 1) MLV_CASE(8, 2) MLV_CASE(8, 4)
    }
#undef MLV_CASE
    return cudaErrorInvalidValue;
}

cudaError_t launch_scan(mlv_index* h, const ScanParams& p, const ScanCfg& c, bool range, cudaStream_t st) {
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (h->timing) {
        for (cudaEvent_t* ev : {&e0, &e1}) {
            if (!h->event_pool.empty()) {
                *ev = h->event_pool.back();
                h->event_pool.pop_back();
            } else {
                cudaError_t e = cudaEventCreate(ev);
                if (e != cudaSuccess) return e;
            }
        }
        cudaEventRecord(e0, st);
    }
    cudaError_t e;
    const bool l2 = h->metric == MLV_L2;
    if (range)
        e = l2 ? launch_scan_m<METRIC_L2, true>(p, c, st) : launch_scan_m<METRIC_IP, true>(p, c, st);
    else
        e = l2 ? launch_scan_m<METRIC_L2, false>(p, c, st) : launch_scan_m<METRIC_IP, false>(p, c, st);
    h->launches++;
    if (h->timing) {
        cudaEventRecord(e1, st);
        h->pending.emplace_back(e0, e1);
    }
    return e;
}

__global__ void fill_empty_kernel(float* d, int64_t* r, int32_t* c, uint32_t nq, uint32_t k) {
    const uint32_t total = nq * k;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        d[i] = __int_as_float(0x7f800000);
        r[i] = -1;
    }
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nq; i += gridDim.x * blockDim.x) c[i] = 0;
}

bool g_select_attr_set[64] = {false};
cudaError_t ensure_select_attrs(int device) {
    if (device >= 0 && device < 64 && g_select_attr_set[device]) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SELECT_MAX_P * 8));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(merge_pairs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SELECT_MAX_P * 8));
    if (e != cudaSuccess) return e;
    if (device >= 0 && device < 64) g_select_attr_set[device] = true;
    return cudaSuccess;
}

// The lane of the stream a search runs on; a new stream takes the least recently used lane
// (after making sure that lane's previous stream is done with the buffers).
Lane* lane_for(mlv_index* h, cudaStream_t st) {
    Lane* lru = &h->lanes[0];
    for (Lane& l : h->lanes) {
        if (l.used && l.stream == st) {
            l.last_use = ++h->lane_clock;
            return &l;
        }
        if (!l.used) {
            if (lru->used) lru = &l;
        } else if (lru->used && l.last_use < lru->last_use) {
            lru = &l;
        }
    }
    if (lru->used) cudaStreamSynchronize(lru->stream);
    lru->used = true;
    lru->stream = st;
    lru->last_use = ++h->lane_clock;
    return lru;
}

// Build the ascending list of rows that are live AND pass `bm` (device bitmap, `words` words; rows
// beyond it do not pass) into list/scratch, on `st`.  scratch[0] (u64) receives the list length.
int build_gather_list(mlv_index* h, DevBuf& list, DevBuf& scratch, const uint32_t* bm, uint64_t words, cudaStream_t st) {
    const uint64_t n = h->rows, n_words = (n + 31) / 32;
    int rc;
    if ((rc = ensure_dev(h, scratch, n_words * 8 + 8)) != MLV_OK) return rc;
    if ((rc = ensure_dev(h, list, std::max<uint64_t>(n, 1) * 4)) != MLV_OK) return rc;
    uint64_t* d_total = (uint64_t*)scratch.p;
    live_prefix_kernel<<<1, 1024, 0, st>>>(h->d_live, n, d_total + 1, d_total, bm, words);
    scatter_passing_rows_kernel<<<(unsigned)std::min<uint64_t>((n_words + 255) / 256, 2048), 256, 0, st>>>(
        h->d_live, bm, words, n, d_total + 1, (uint32_t*)list.p);
    h->launches += 2;
    CK(h, cudaGetLastError());
    return MLV_OK;
}

// What a search reads its rows through: nothing special, a bitmap checked per row, or a gather list.
struct FilterPlan {
    const uint32_t* bitmap = nullptr;      // stream + mask
    const uint32_t* gather = nullptr;      // row list (live AND passing)
    const uint32_t* n_rows_dev = nullptr;  // its length (device)
};

// filter_dev: per-call bitmap (ceil(rows/32) words) or null; a bound prepared filter applies when it is null.
int plan_filter(mlv_index* h, Lane* ln, const uint32_t* filter_dev, cudaStream_t st, FilterPlan* out) {
    *out = FilterPlan{};
    mlv_filter* f = filter_dev ? nullptr : h->bound_filter;
    if (!filter_dev && !f) return MLV_OK;
    int rc;
    if (f) {
        if (f->compact_gen != h->compact_gen)
            return fail(h, MLV_E_INVALID, "prepared filter predates a compaction / clear of the index (rows were renumbered); create it again");
        if (f->epoch != h->epoch) {  // rows were added / deleted since: rebuild (device only; count re-read lazily)
            if ((rc = build_gather_list(h, f->d_list, f->d_scratch, (const uint32_t*)f->d_bitmap.p, f->bitmap_words, st)) != MLV_OK) return rc;
            f->epoch = h->epoch;
            f->counted = false;
        }
        const bool covers = f->bitmap_words >= (h->rows + 31) / 32;  // only then can the bitmap mask a full stream
        const bool dense = f->counted && f->passing * 4 >= (h->rows - h->n_deleted) * 3;
        const bool want_stream = h->tune_gather == 0 || (h->tune_gather < 0 && dense);
        if (want_stream && covers) {
            out->bitmap = (const uint32_t*)f->d_bitmap.p;  // stream every row, mask in the epilogue
            return MLV_OK;
        }
        out->gather = (const uint32_t*)f->d_list.p;
        out->n_rows_dev = (const uint32_t*)f->d_scratch.p;  // low word of the u64 total
        return MLV_OK;
    }
    if (h->tune_gather == 0) {
        out->bitmap = filter_dev;
        return MLV_OK;
    }
    if ((rc = build_gather_list(h, ln->d_flist, ln->d_fscratch, filter_dev, (h->rows + 31) / 32, st)) != MLV_OK) return rc;
    out->gather = (const uint32_t*)ln->d_flist.p;
    out->n_rows_dev = (const uint32_t*)ln->d_fscratch.p;
    return MLV_OK;
}

int ensure_sched(mlv_index* h, Lane* ln) {
    if (ln->d_sched.p) return MLV_OK;
    int rc = ensure_dev(h, ln->d_sched, 8);
    if (rc != MLV_OK) return rc;
    CK(h, cudaMemset(ln->d_sched.p, 0, ln->d_sched.bytes));
    return MLV_OK;
}

void fill_sched(mlv_index* h, Lane* ln, ScanParams& p) {
    p.sched = h->tune_dynamic ? (uint32_t*)ln->d_sched.p : nullptr;
    p.tile_batch = (uint32_t)std::max(h->tune_tile_batch, 1);
}

// can the last CTA fold the whole grid's lists (and hold its scratch in the idle ring)?
bool fused_ok(const mlv_index* h, const ScanCfg& c, uint32_t k) {
    if (!h->tune_dynamic || !h->tune_fused) return false;
    if ((uint64_t)c.grid * k > SCAN_FUSED_MAX_KEYS || (uint64_t)c.CW * k > 1024) return false;
    return (size_t)c.S * c.stage_f4 * 16 >= ((size_t)SCAN_FUSED_MAX_KEYS + (size_t)c.NQ * k) * 8;
}

// the exchange path must take the same decision on every rank, whatever its shard's grid is
bool exchange_ok(const mlv_index* h, uint32_t k) {
    return h->xchg && h->xchg->connected && h->tune_dynamic && h->tune_fused && k <= XCHG_MAX_K &&
           (uint64_t)h->sm_count * k <= SCAN_FUSED_MAX_KEYS;
}

void fill_exchange(mlv_index* h, ExchangeView& x) {
    const mlv_exchange* e = h->xchg;
    x.world = e->world;
    x.rank = e->rank;
    for (uint32_t i = 0; i < e->world; i++) {
        x.bufs[i] = e->bufs[i];
        x.row_bases[i] = h->xchg_row_bases[i];
    }
    x.error = e->d_error;
}

// qprep: prepared queries [nq, ld] in device memory; all output pointers in device memory.
// exchange: merge with the other ranks' results over peer memory (caller checked exchange_ok).
int search_prepared(mlv_index* h, const float* qprep, uint32_t nq, uint32_t k, const uint32_t* filter_dev, float* out_d,
                    int64_t* out_r, int32_t* out_c, cudaStream_t st, bool exchange = false) {
    Lane* ln = lane_for(h, st);
    FilterPlan fp;
    int rc = plan_filter(h, ln, filter_dev, st, &fp);
    if (rc != MLV_OK) return rc;
    ScanCfg c;
    if ((rc = choose_cfg(h, nq, k, false, &c, fp.gather != nullptr)) != MLV_OK) return rc;
    if ((rc = ensure_sched(h, ln)) != MLV_OK) return rc;
    const bool fused = fused_ok(h, c, k);
    if (exchange && !fused) return fail(h, MLV_E_UNSUPPORTED, "exchange search needs the fused final select");
    const uint32_t F = SELECT_MAX_P / k;  // lists one select CTA can fold (>= 8)
    // bound the candidate scratch: chunk * grid * k keys
    uint32_t chunk = (uint32_t)std::max<size_t>(1, ((size_t)64 << 20) / ((size_t)c.grid * k * 8));
    chunk = std::max<uint32_t>(chunk / c.NQ * c.NQ, c.NQ);
    chunk = std::min(chunk, nq);
    rc = ensure_dev(h, ln->d_keys0, (size_t)chunk * c.grid * k * 8);
    if (rc != MLV_OK) return rc;
    const uint32_t lists1 = ((uint32_t)c.grid + F - 1) / F;
    if (lists1 > 1) {
        rc = ensure_dev(h, ln->d_keys1, (size_t)chunk * lists1 * k * 8);
        if (rc != MLV_OK) return rc;
    }
    CK(h, ensure_select_attrs(h->device));

    ScanParams p{};
    p.rows = reinterpret_cast<const float4*>(h->d_rows);
    p.n_rows = (uint32_t)h->rows;
    p.ld4 = h->ld / 4;
    p.tile_rows = c.T;
    p.n_tiles = (uint32_t)((h->rows + c.T - 1) / c.T);
    p.stages = c.S;
    p.producer_warps = (uint32_t)c.PW;
    p.stage_f4 = c.stage_f4;
    p.k = k;
    p.live = h->n_deleted ? h->d_live : nullptr;
    p.filter = fp.bitmap;
    p.gather = fp.gather;
    p.n_rows_dev = fp.n_rows_dev;
    p.evict_first = c.evict_first;
    fill_sched(h, ln, p);
    p.fused = fused ? 1 : 0;
    p.row_base = h->row_base;
    if (exchange) fill_exchange(h, p.xchg);
    if (h->tune_timeline) {
        rc = ensure_dev(h, h->d_timeline, (size_t)c.grid * 4 * 8);
        if (rc != MLV_OK) return rc;
        p.timeline = (unsigned long long*)h->d_timeline.p;
        h->last_grid = c.grid;
    }

    for (uint32_t q0 = 0; q0 < nq; q0 += chunk) {
        const uint32_t nchunk = std::min(chunk, nq - q0);
        for (uint32_t g0 = 0; g0 < nchunk; g0 += c.NQ) {
            p.queries = reinterpret_cast<const float4*>(qprep + (size_t)(q0 + g0) * h->ld);
            p.nq_valid = std::min<uint32_t>(c.NQ, nchunk - g0);
            p.out_keys = (uint64_t*)ln->d_keys0.p + (size_t)g0 * c.grid * k;
            if (fused) {
                p.out_dists = out_d + (size_t)(q0 + g0) * k;
                p.out_rows = out_r + (size_t)(q0 + g0) * k;
                p.out_counts = out_c + (q0 + g0);
                if (exchange) p.xchg.seq = ++h->xseq;
            }
            CK(h, launch_scan(h, p, c, false, st));
        }
        if (fused) continue;  // the last CTA of every launch already wrote the final top-k
        // fold the grid's lists into one per query
        const uint64_t* in = (const uint64_t*)ln->d_keys0.p;
        uint64_t* bufs[2] = {(uint64_t*)ln->d_keys1.p, (uint64_t*)ln->d_keys0.p};
        uint32_t n_lists = (uint32_t)c.grid;
        int flip = 0;
        for (;;) {
            SelectParams sp{};
            sp.in_keys = in;
            sp.n_lists = n_lists;
            sp.k = k;
            sp.lists_per_block = std::min(F, n_lists);
            sp.n_out_lists = (n_lists + sp.lists_per_block - 1) / sp.lists_per_block;
            sp.P = pow2_ceil(std::max<uint32_t>(sp.lists_per_block * k, 2));
            sp.final_pass = sp.n_out_lists == 1;
            sp.out_keys = bufs[flip];
            sp.out_dists = out_d + (size_t)q0 * k;
            sp.out_rows = out_r + (size_t)q0 * k;
            sp.out_counts = out_c + q0;
            sp.row_base = h->row_base;
            const int threads = (int)std::min<uint32_t>(SELECT_THREADS, std::max<uint32_t>(sp.P / 2, 32));
            select_kernel<<<dim3(sp.n_out_lists, nchunk), threads, (size_t)sp.P * 8, st>>>(sp);
            h->launches++;
            CK(h, cudaGetLastError());
            if (sp.final_pass) break;
            in = sp.out_keys;
            n_lists = sp.n_out_lists;
            flip ^= 1;
        }
    }
    return MLV_OK;
}

int prep_queries(mlv_index* h, const float* q_dev_raw, uint32_t nq, cudaStream_t st) {
    Lane* ln = lane_for(h, st);
    int rc = ensure_dev(h, ln->d_q, (size_t)nq * h->ld * 4);
    if (rc != MLV_OK) return rc;
    const int wpb = 4;
    prep_queries_kernel<<<(nq + wpb - 1) / wpb, wpb * 32, 0, st>>>(q_dev_raw, (float*)ln->d_q.p, nq, h->dim, h->ld,
                                                                 h->metric == MLV_COSINE);
    h->launches++;
    CK(h, cudaGetLastError());
    return MLV_OK;
}

// Order n > SELECT_MAX_P keys (device, in a scratch copy padded to a power of two) and decode them.
int sort_big_device(mlv_index* h, const uint64_t* keys, uint64_t n, float* out_d, int64_t* out_r, cudaStream_t st) {
    uint64_t P = SELECT_MAX_P;
    while (P < n) P <<= 1;
    if (P > (1ull << 31)) return fail(h, MLV_E_UNSUPPORTED, "range hit list too long to order");
    int rc = ensure_dev(h, h->d_misc, P * 8);
    if (rc != MLV_OK) return rc;
    uint64_t* a = (uint64_t*)h->d_misc.p;
    CK(h, cudaMemcpyAsync(a, keys, n * 8, cudaMemcpyDeviceToDevice, st));
    if (P > n) fill_sentinel_kernel<<<(unsigned)std::min<uint64_t>((P - n + 255) / 256, 1024), 256, 0, st>>>(a, n, P);
    CK(h, cudaFuncSetAttribute(bitonic_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SELECT_MAX_P * 8)));
    const unsigned blocks = (unsigned)(P / SELECT_MAX_P);
    bitonic_block_kernel<<<blocks, SELECT_THREADS, (size_t)SELECT_MAX_P * 8, st>>>(a, 2, SELECT_MAX_P);
    for (uint64_t size = 2ull * SELECT_MAX_P; size <= P; size <<= 1) {
        for (uint64_t stride = size >> 1; stride >= SELECT_MAX_P; stride >>= 1)
            bitonic_global_kernel<<<(unsigned)std::min<uint64_t>((P / 2 + 255) / 256, 4096), 256, 0, st>>>(a, (uint32_t)P, (uint32_t)size,
                                                                                                         (uint32_t)stride);
        bitonic_block_kernel<<<blocks, SELECT_THREADS, (size_t)SELECT_MAX_P * 8, st>>>(a, (uint32_t)size, (uint32_t)size);
    }
    decode_keys_kernel<<<(unsigned)std::min<uint64_t>((n + 255) / 256, 1024), 256, 0, st>>>(a, n, h->row_base, out_d, out_r);
    h->launches += 3;
    CK(h, cudaGetLastError());
    return MLV_OK;
}

// ---- tensor-core batch path (gemm_kernel.cuh) ----------------------------------------------------
typedef CUresult (*encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// cuTensorMapEncodeTiled through the runtime's driver entry point: libcuda is not linked, so the
// library still loads (and exports its symbols) on a box without a driver.
encode_tiled_fn get_encode_tiled() {
    static encode_tiled_fn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (encode_tiled_fn)p;
        else
            cudaGetLastError();
    }
    return fn;
}

// fp32 matrix [n_rows, ld] row-major -> boxes of {GEMM_BK floats, box_rows rows}, 128-byte swizzle,
// out-of-range elements read as zero (ragged last row tile, ld not a multiple of 32)
int make_tile_map(mlv_index* h, CUtensorMap* map, const float* base, uint64_t n_rows, uint32_t box_rows) {
    encode_tiled_fn enc = get_encode_tiled();
    if (!enc) return fail(h, MLV_E_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t gdim[2] = {h->ld, n_rows};
    const cuuint64_t gstride[1] = {(cuuint64_t)h->ld * 4};
    const cuuint32_t box[2] = {(cuuint32_t)GEMM_BK, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(h, MLV_E_CUDA, "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
    return MLV_OK;
}

uint32_t gemm_kprime(uint32_t k) {
    const uint32_t slack = std::max<uint32_t>(16, k / 4);
    return (k + slack + 31) & ~31u;
}

bool gemm_eligible(const mlv_index* h, uint32_t nq, uint32_t k) {
    if (h->tune_gemm == 0) return false;
    if (h->ld < (uint32_t)GEMM_BK) return false;
    if (gemm_kprime(k) * 4 > SELECT_MAX_P) return false;
    if (h->tune_gemm == 1) return true;
    return nq >= (uint32_t)std::max(h->tune_gemm_min_nq, 1) && h->rows >= 16384;
}

int ensure_row_norms(mlv_index* h, cudaStream_t st) {
    int rc;
    if (!h->d_maxn2.p) {
        if ((rc = ensure_dev(h, h->d_maxn2, 4)) != MLV_OK) return rc;
        CK(h, cudaMemsetAsync(h->d_maxn2.p, 0, 4, st));
        h->norms_valid = 0;
    }
    if (h->d_norms.bytes < h->rows * 4) {
        // growing reallocates: recompute everything (rows rarely grow between large batches)
        if ((rc = ensure_dev(h, h->d_norms, std::max<uint64_t>(h->capacity, h->rows) * 4)) != MLV_OK) return rc;
        h->norms_valid = 0;
    }
    if (h->norms_valid == 0) CK(h, cudaMemsetAsync(h->d_maxn2.p, 0, 4, st));
    if (h->norms_valid < h->rows) {
        const uint64_t n = h->rows - h->norms_valid;
        const int wpb = 8;
        row_norms_kernel<<<(unsigned)((n + wpb - 1) / wpb), wpb * 32, 0, st>>>(h->d_rows, h->norms_valid, n, h->ld,
                                                                             (float*)h->d_norms.p, (uint32_t*)h->d_maxn2.p);
        h->launches++;
        CK(h, cudaGetLastError());
        h->norms_valid = h->rows;
    }
    return MLV_OK;
}

template <int METRIC, int BN>
cudaError_t launch_gemm_tt(const CUtensorMap& mx, const CUtensorMap& mqh, const CUtensorMap& mql, const GemmParams& gp, int grid,
                           cudaStream_t st) {
    auto kern = gemm_topk_kernel<METRIC, BN>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GemmShape<BN>::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    kern<<<grid, GEMM_THREADS, GemmShape<BN>::SMEM_BYTES, st>>>(mx, mqh, mql, gp);
    return cudaGetLastError();
}
template <int METRIC>
cudaError_t launch_gemm_t(const CUtensorMap& mx, const CUtensorMap& mqh, const CUtensorMap& mql, const GemmParams& gp, int grid,
                          cudaStream_t st, int bn) {
    return bn == 128 ? launch_gemm_tt<METRIC, 128>(mx, mqh, mql, gp, grid, st) : launch_gemm_tt<METRIC, 256>(mx, mqh, mql, gp, grid, st);
}

// Large batches: tcgen05 GEMM selects k' candidates per query in geometrically growing rounds,
// rerank_kernel scores them in the reference's arithmetic and certifies; uncertified queries are
// re-run by the exact scan.  Synchronises `st` once (to read the per-query flags).
int search_gemm(mlv_index* h, const float* qprep, uint32_t nq, uint32_t k, const uint32_t* filter_dev, float* out_d,
                int64_t* out_r, int32_t* out_c, cudaStream_t st) {
    int rc;
    const uint32_t ld = h->ld;
    const uint32_t GEMM_BN = h->tune_gemm_bn == 128 ? 128 : 256;
    const uint32_t nq_pad = (nq + GEMM_BN - 1) / GEMM_BN * GEMM_BN;
    const uint32_t kprime = gemm_kprime(k);
    const uint32_t cap = std::min<uint32_t>(SELECT_MAX_P, pow2_ceil(8 * kprime));
    const uint32_t P = cap;  // power of two
    const bool l2 = h->metric == MLV_L2;
    if (h->metric != MLV_COSINE) {
        if ((rc = ensure_row_norms(h, st)) != MLV_OK) return rc;
    }
    // scratch: Qhi | Qlo | qn | thr | cnt | flags
    const size_t qmat = (size_t)nq_pad * ld * 4;
    if ((rc = ensure_dev(h, h->d_gq, 2 * qmat + (size_t)nq_pad * 16)) != MLV_OK) return rc;
    if ((rc = ensure_dev(h, h->d_cand, (size_t)nq_pad * cap * 8)) != MLV_OK) return rc;
    float* qhi = (float*)h->d_gq.p;
    float* qlo = qhi + (size_t)nq_pad * ld;
    float* qn = qlo + (size_t)nq_pad * ld;
    float* thr = qn + nq_pad;
    uint32_t* cnt = (uint32_t*)(thr + nq_pad);
    uint32_t* flags = cnt + nq_pad;
    uint64_t* cand = (uint64_t*)h->d_cand.p;
    {
        const int wpb = 8;
        split_queries_kernel<<<(nq_pad + wpb - 1) / wpb, wpb * 32, 0, st>>>(qprep, qhi, qlo, qn, thr, cnt, flags, nq, nq_pad, ld);
        h->launches++;
        CK(h, cudaGetLastError());
    }
    CUtensorMap mx, mqh, mql;
    if ((rc = make_tile_map(h, &mx, h->d_rows, h->rows, GEMM_BM)) != MLV_OK) return rc;
    if ((rc = make_tile_map(h, &mqh, qhi, nq_pad, GEMM_BN)) != MLV_OK) return rc;
    if ((rc = make_tile_map(h, &mql, qlo, nq_pad, GEMM_BN)) != MLV_OK) return rc;
    CK(h, cudaFuncSetAttribute(refine_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SELECT_MAX_P * 8)));

    GemmParams gp{};
    gp.n_rows = (uint32_t)h->rows;
    gp.nq = nq;
    gp.n_qtiles = nq_pad / GEMM_BN;
    gp.n_kchunks = (ld + GEMM_BK - 1) / GEMM_BK;
    gp.row_norms = l2 ? (const float*)h->d_norms.p : nullptr;
    gp.q_norms = qn;
    gp.thr = thr;
    gp.live = h->n_deleted ? h->d_live : nullptr;
    gp.filter = filter_dev;
    if (!filter_dev && h->bound_filter) {
        if (h->bound_filter->compact_gen != h->compact_gen)
            return fail(h, MLV_E_INVALID, "prepared filter predates a compaction / clear of the index (rows were renumbered); create it again");
        if (h->bound_filter->bitmap_words < (h->rows + 31) / 32) return fail(h, MLV_E_INVALID, "prepared filter is shorter than the index; re-create it");
        gp.filter = (const uint32_t*)h->bound_filter->d_bitmap.p;
    }
    gp.cand = cand;
    gp.cand_cnt = cnt;
    gp.cap = cap;

    // rounds: the first takes as many rows as a candidate buffer holds (no threshold yet), each
    // later one (cap - k') / (4 k') times the rows seen so far, so a buffer is expected to stay
    // at most a quarter full however the thresholds started
    const uint32_t total_tiles = (uint32_t)((h->rows + GEMM_BM - 1) / GEMM_BM);
    const double growth = (double)(cap - kprime) / (4.0 * kprime);
    uint32_t seen = 0;
    const int refine_threads = (int)std::min<uint32_t>(SELECT_THREADS, std::max<uint32_t>(P / 2, 32));
    while (seen < total_tiles) {
        uint32_t take = seen == 0 ? std::max<uint32_t>(1, cap / GEMM_BM) : std::max<uint32_t>(1, (uint32_t)(seen * growth));
        take = std::min(take, total_tiles - seen);
        gp.row_tile0 = seen;
        gp.row_tile1 = seen + take;
        const uint64_t items = (uint64_t)take * gp.n_qtiles;
        const int grid = (int)std::min<uint64_t>(items, (uint64_t)h->sm_count);
        cudaEvent_t e0 = nullptr, e1 = nullptr;
        if (h->timing) {
            for (cudaEvent_t* ev : {&e0, &e1}) {
                if (!h->event_pool.empty()) {
                    *ev = h->event_pool.back();
                    h->event_pool.pop_back();
                } else {
                    CK(h, cudaEventCreate(ev));
                }
            }
            cudaEventRecord(e0, st);
        }
        CK(h, l2 ? launch_gemm_t<METRIC_L2>(mx, mqh, mql, gp, grid, st, (int)GEMM_BN)
                 : launch_gemm_t<METRIC_IP>(mx, mqh, mql, gp, grid, st, (int)GEMM_BN));
        if (h->timing) {
            cudaEventRecord(e1, st);
            h->gemm_pending.emplace_back(e0, e1);
        }
        refine_kernel<<<nq, refine_threads, (size_t)P * 8, st>>>(cand, cnt, thr, flags, cap, P, kprime);
        CK(h, cudaGetLastError());
        h->launches += 2;
        h->gemm_launches++;
        h->gemm_rounds++;
        seen += take;
    }

    RerankParams rp{};
    rp.rows = reinterpret_cast<const float4*>(h->d_rows);
    rp.ld4 = ld / 4;
    rp.queries = reinterpret_cast<const float4*>(qprep);
    rp.q_norms = qn;
    rp.max_norm2_bits = h->metric == MLV_COSINE ? nullptr : (const uint32_t*)h->d_maxn2.p;
    rp.cand = cand;
    rp.cnt = cnt;
    rp.flags = flags;
    rp.cap = cap;
    rp.kprime = kprime;
    rp.k = k;
    rp.P = pow2_ceil(std::max<uint32_t>(kprime, 2));
    rp.out_dists = out_d;
    rp.out_rows = out_r;
    rp.out_counts = out_c;
    rp.row_base = h->row_base;
    rp.metric = h->metric;
    if (l2)
        rerank_kernel<METRIC_L2><<<nq, 256, (size_t)rp.P * 8, st>>>(rp);
    else
        rerank_kernel<METRIC_IP><<<nq, 256, (size_t)rp.P * 8, st>>>(rp);
    h->launches++;
    CK(h, cudaGetLastError());

    // certificate check: the one synchronisation of this path
    std::vector<uint32_t> hflags(nq);
    CK(h, cudaMemcpyAsync(hflags.data(), flags, (size_t)nq * 4, cudaMemcpyDeviceToHost, st));
    CK(h, cudaStreamSynchronize(st));
    h->gemm_searches++;
    h->gemm_queries += nq;
    for (uint32_t q = 0; q < nq; q++) {
        if (!hflags[q]) continue;
        h->gemm_fallback_queries++;
        rc = search_prepared(h, qprep + (size_t)q * ld, 1, k, filter_dev, out_d + (size_t)q * k, out_r + (size_t)q * k, out_c + q, st);
        if (rc != MLV_OK) return rc;
    }
    return MLV_OK;
}

}  // namespace

// =================================================================================== C ABI
extern "C" {

int mlv_abi_version(void) { return MLV_ABI_VERSION; }

int mlv_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

const char* mlv_status_string(int s) {
    switch (s) {
        case MLV_OK: return "ok";
        case MLV_E_INVALID: return "invalid argument";
        case MLV_E_CUDA: return "CUDA error";
        case MLV_E_NOMEM: return "out of memory";
        case MLV_E_UNSUPPORTED: return "unsupported by this build";
        case MLV_E_NO_DEVICE: return "no CUDA device (there is no CPU fallback)";
        default: return "unknown status";
    }
}

const char* mlv_last_error(mlv_index_t h) { return h ? h->err.c_str() : "null handle"; }

int mlv_index_create(uint32_t dim, int metric, uint64_t capacity_hint, int device, mlv_index_t* out) {
    if (!out) return MLV_E_INVALID;
    *out = nullptr;
    if (dim == 0 || metric < MLV_L2 || metric > MLV_COSINE) return MLV_E_INVALID;
    int ndev = mlv_device_count();
    if (ndev <= 0) return MLV_E_NO_DEVICE;
    if (device < 0 || device >= ndev) return MLV_E_INVALID;
    mlv_index* h = new (std::nothrow) mlv_index();
    if (!h) return MLV_E_NOMEM;
    h->device = device;
    h->dim = dim;
    h->ld = (dim + 3u) & ~3u;
    h->metric = metric;
    h->tune_cw = env_int("MLV_SCAN_CW", h->tune_cw);
    h->tune_stage_kb = env_int("MLV_SCAN_STAGE_KB", h->tune_stage_kb);
    h->tune_evict_first = env_int("MLV_SCAN_EVICT_FIRST", h->tune_evict_first);
    h->tune_r = env_int("MLV_SCAN_R", h->tune_r);
    h->tune_max_stages = env_int("MLV_SCAN_MAX_STAGES", h->tune_max_stages);
    h->tune_ctas = env_int("MLV_SCAN_CTAS", h->tune_ctas);
    h->tune_pw = env_int("MLV_SCAN_PW", h->tune_pw);
    h->tune_dynamic = env_int("MLV_SCAN_DYNAMIC", h->tune_dynamic);
    h->tune_tile_batch = env_int("MLV_SCAN_TILE_BATCH", h->tune_tile_batch);
    h->tune_fused = env_int("MLV_SCAN_FUSED", h->tune_fused);
    h->tune_gemm = env_int("MLV_GEMM", h->tune_gemm);
    h->tune_gemm_min_nq = env_int("MLV_GEMM_MIN_NQ", h->tune_gemm_min_nq);
    h->tune_gemm_bn = env_int("MLV_GEMM_BN", h->tune_gemm_bn);
    DeviceGuard g(device);
    cudaDeviceProp prop;
    cudaError_t e = g.ok ? cudaGetDeviceProperties(&prop, device) : cudaErrorInvalidDevice;
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        cudaGetLastError();
        delete h;
        return MLV_E_CUDA;
    }
    h->sm_count = prop.multiProcessorCount;
    h->smem_optin = prop.sharedMemPerBlockOptin;
    if (prop.major != 10) {
        // sm_100a cubin only: fail loudly instead of at the first launch
        cudaStreamDestroy(h->stream);
        delete h;
        return MLV_E_NO_DEVICE;
    }
    if (capacity_hint) {
        int rc = reserve_rows(h, capacity_hint);
        if (rc != MLV_OK) {
            cudaStreamDestroy(h->stream);
            delete h;
            return rc;
        }
    }
    *out = h;
    return MLV_OK;
}

int mlv_index_destroy(mlv_index_t h) {
    if (!h) return MLV_E_INVALID;
    DeviceGuard g(h->device);
    cudaStreamSynchronize(h->stream);
    if (h->d_rows) cudaFree(h->d_rows);
    if (h->d_live) cudaFree(h->d_live);
    for (DevBuf* b : {&h->d_qraw, &h->d_filter, &h->d_outd, &h->d_outr, &h->d_outc, &h->d_misc, &h->d_range, &h->d_timeline,
                      &h->d_norms, &h->d_gq, &h->d_cand, &h->d_maxn2})
        free_dev(*b);
    for (Lane& l : h->lanes)
        for (DevBuf* b : {&l.d_q, &l.d_keys0, &l.d_keys1, &l.d_sched, &l.d_flist, &l.d_fscratch}) free_dev(*b);
    if (h->h_stage.p) cudaFreeHost(h->h_stage.p);
    for (auto* vec : {&h->pending, &h->gemm_pending})
        for (auto& pr : *vec) {
            cudaEventDestroy(pr.first);
            cudaEventDestroy(pr.second);
        }
    for (auto ev : h->event_pool) cudaEventDestroy(ev);
    cudaStreamDestroy(h->stream);
    delete h;
    return MLV_OK;
}

int mlv_index_set_row_base(mlv_index_t h, uint64_t row_base) {
    if (!h) return MLV_E_INVALID;
    h->row_base = row_base;
    return MLV_OK;
}

int mlv_index_set_tuning(mlv_index_t h, const char* key, int value) {
    if (!h || !key) return MLV_E_INVALID;
    std::string k(key);
    if (k == "cw") h->tune_cw = value;
    else if (k == "stage_kb") h->tune_stage_kb = value;
    else if (k == "evict_first") h->tune_evict_first = value;
    else if (k == "r") h->tune_r = value;
    else if (k == "max_stages") h->tune_max_stages = value;
    else if (k == "ctas") h->tune_ctas = value;
    else if (k == "pw") h->tune_pw = value;
    else if (k == "timeline") h->tune_timeline = value;
    else if (k == "dynamic") h->tune_dynamic = value;
    else if (k == "tile_batch") h->tune_tile_batch = value;
    else if (k == "fused") h->tune_fused = value;
    else if (k == "gather") h->tune_gather = value;
    else if (k == "gemm") h->tune_gemm = value;
    else if (k == "gemm_min_nq") h->tune_gemm_min_nq = value;
    else if (k == "gemm_bn") h->tune_gemm_bn = value;
    else return fail(h, MLV_E_INVALID, "unknown tuning key " + k);
    return MLV_OK;
}

static int add_common(mlv_index_t h, const float* rows, uint64_t n, uint64_t* first_row, cudaMemcpyKind kind) {
    if (!h || (!rows && n)) return MLV_E_INVALID;
    if (n == 0) {
        if (first_row) *first_row = h->rows;
        return MLV_OK;
    }
    DeviceGuard g(h->device);
    int rc = reserve_rows(h, h->rows + n);
    if (rc != MLV_OK) return rc;
    float* dst = h->d_rows + h->rows * h->ld;
    // pitched copy straight into the matrix; padding columns were zeroed at allocation
    const uint64_t max_rows_per_copy = 1u << 20;  // cudaMemcpy2D height limits
    for (uint64_t r0 = 0; r0 < n; r0 += max_rows_per_copy) {
        const uint64_t nr = std::min(max_rows_per_copy, n - r0);
        CK(h, cudaMemcpy2DAsync(dst + r0 * h->ld, (size_t)h->ld * 4, rows + r0 * h->dim, (size_t)h->dim * 4,
                                (size_t)h->dim * 4, nr, kind, h->stream));
    }
    return finish_append(h, n, first_row);
}

int mlv_index_add(mlv_index_t h, const float* rows, uint64_t n, uint64_t* first_row) {
    return add_common(h, rows, n, first_row, cudaMemcpyHostToDevice);
}
int mlv_index_add_device(mlv_index_t h, const float* rows_dev, uint64_t n, uint64_t* first_row) {
    return add_common(h, rows_dev, n, first_row, cudaMemcpyDeviceToDevice);
}

int mlv_index_add_synthetic(mlv_index_t h, uint64_t seed, uint64_t first_gen_row, uint64_t n, int scaled,
                            uint64_t* first_row) {
    if (!h) return MLV_E_INVALID;
    if (n == 0) {
        if (first_row) *first_row = h->rows;
        return MLV_OK;
    }
    DeviceGuard g(h->device);
    int rc = reserve_rows(h, h->rows + n);
    if (rc != MLV_OK) return rc;
    const uint64_t key = splitmix64(seed), key2 = splitmix64(seed ^ 0xA5A5A5A5A5A5A5A5ull);
    const uint64_t total = n * h->ld;
    const unsigned blocks = (unsigned)std::min<uint64_t>((total + 255) / 256, (uint64_t)h->sm_count * 32);
    fill_synthetic_kernel<<<blocks, 256, 0, h->stream>>>(h->d_rows, h->rows, n, h->dim, h->ld, key, key2, first_gen_row,
                                                         scaled);
    h->launches++;
    CK(h, cudaGetLastError());
    return finish_append(h, n, first_row);
}

int mlv_index_mark_deleted(mlv_index_t h, const uint64_t* rows, uint64_t n, uint64_t* newly_deleted) {
    if (!h || (!rows && n)) return MLV_E_INVALID;
    if (newly_deleted) *newly_deleted = 0;
    if (n == 0 || h->rows == 0) return MLV_OK;
    DeviceGuard g(h->device);
    int rc = ensure_dev(h, h->d_misc, n * 8 + 8);
    if (rc != MLV_OK) return rc;
    unsigned long long* d_changed = (unsigned long long*)h->d_misc.p;
    uint64_t* d_ids = (uint64_t*)h->d_misc.p + 1;
    CK(h, cudaMemsetAsync(d_changed, 0, 8, h->stream));
    CK(h, cudaMemcpyAsync(d_ids, rows, n * 8, cudaMemcpyHostToDevice, h->stream));
    mark_deleted_kernel<<<(unsigned)std::min<uint64_t>((n + 255) / 256, 1024), 256, 0, h->stream>>>(h->d_live, d_ids, n,
                                                                                                h->rows, d_changed);
    h->launches++;
    CK(h, cudaGetLastError());
    unsigned long long changed = 0;
    CK(h, cudaMemcpyAsync(&changed, d_changed, 8, cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    h->n_deleted += changed;
    if (changed) h->epoch++;
    if (newly_deleted) *newly_deleted = changed;
    return MLV_OK;
}

int mlv_index_compact(mlv_index_t h, int64_t* old_to_new, uint64_t* new_rows) {
    if (!h) return MLV_E_INVALID;
    DeviceGuard g(h->device);
    const uint64_t n = h->rows;
    if (n == 0 || h->n_deleted == 0) {
        if (old_to_new)
            for (uint64_t i = 0; i < n; i++) old_to_new[i] = (int64_t)i;
        if (new_rows) *new_rows = n;
        return MLV_OK;
    }
    const uint64_t n_words = (n + 31) / 32;
    const size_t row_bytes = (size_t)h->ld * 4;
    int rc = ensure_dev(h, h->d_misc, n_words * 8 + 8 + (old_to_new ? n * 8 : 0));
    if (rc != MLV_OK) return rc;
    uint64_t* d_total = (uint64_t*)h->d_misc.p;
    uint64_t* d_wbase = d_total + 1;
    int64_t* d_map = old_to_new ? (int64_t*)(d_wbase + n_words) : nullptr;
    float* nrows = nullptr;
    CK(h, cudaMalloc(&nrows, h->capacity * row_bytes));
    cudaError_t e = cudaMemsetAsync(nrows, 0, h->capacity * row_bytes, h->stream);
    if (e == cudaSuccess) {
        live_prefix_kernel<<<1, 1024, 0, h->stream>>>(h->d_live, n, d_wbase, d_total);
        const int wpb = 8;
        compact_rows_kernel<<<(unsigned)((n + wpb - 1) / wpb), wpb * 32, 0, h->stream>>>(
            reinterpret_cast<const float4*>(h->d_rows), reinterpret_cast<float4*>(nrows), h->d_live, d_wbase, n, h->ld / 4,
            d_map);
        h->launches += 2;
        e = cudaGetLastError();
    }
    uint64_t total = 0;
    if (e == cudaSuccess) e = cudaMemcpyAsync(&total, d_total, 8, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess && old_to_new) e = cudaMemcpyAsync(old_to_new, d_map, n * 8, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    if (e != cudaSuccess) {
        cudaFree(nrows);
        return fail_cuda(h, e, "compact");
    }
    cudaFree(h->d_rows);
    h->d_rows = nrows;
    h->rows = total;
    h->n_deleted = 0;
    h->norms_valid = 0;
    h->epoch++;
    h->compact_gen++;
    CK(h, cudaMemsetAsync(h->d_live, 0, h->live_words * 4, h->stream));
    if (total) {
        const uint64_t words = (total + 31) / 32;
        set_live_range_kernel<<<(unsigned)std::min<uint64_t>((words + 255) / 256, 4096), 256, 0, h->stream>>>(h->d_live, 0, total);
        h->launches++;
        CK(h, cudaGetLastError());
    }
    CK(h, cudaStreamSynchronize(h->stream));
    if (new_rows) *new_rows = total;
    return MLV_OK;
}

int mlv_index_clear(mlv_index_t h) {
    if (!h) return MLV_E_INVALID;
    DeviceGuard g(h->device);
    if (h->d_live) {
        CK(h, cudaMemsetAsync(h->d_live, 0, h->live_words * 4, h->stream));
        CK(h, cudaStreamSynchronize(h->stream));
    }
    h->rows = 0;
    h->n_deleted = 0;
    h->norms_valid = 0;
    h->epoch++;
    h->compact_gen++;
    return MLV_OK;
}

int mlv_index_search_device(mlv_index_t h, const float* queries_dev, uint32_t nq, uint32_t k,
                            const uint32_t* filter_bitmap_dev, float* out_dists_dev, int64_t* out_rows_dev,
                            int32_t* out_counts_dev, void* stream) {
    if (!h || !queries_dev || !out_dists_dev || !out_rows_dev || !out_counts_dev || nq == 0 || k == 0) return fail(h, MLV_E_INVALID, "bad argument");
    if (k > MLV_MAX_K) return fail(h, MLV_E_UNSUPPORTED, "k exceeds MLV_MAX_K");
    DeviceGuard g(h->device);
    cudaStream_t st = (cudaStream_t)stream;  // NULL = the legacy default stream, as everywhere in CUDA
    if (h->rows == h->n_deleted) {  // nothing live (reference index.py:99-104 returns [])
        fill_empty_kernel<<<32, 256, 0, st>>>(out_dists_dev, out_rows_dev, out_counts_dev, nq, k);
        h->launches++;
        CK(h, cudaGetLastError());
        return MLV_OK;
    }
    int rc = prep_queries(h, queries_dev, nq, st);
    if (rc != MLV_OK) return rc;
    if (gemm_eligible(h, nq, k))
        return search_gemm(h, (const float*)lane_for(h, st)->d_q.p, nq, k, filter_bitmap_dev, out_dists_dev, out_rows_dev, out_counts_dev, st);
    return search_prepared(h, (const float*)lane_for(h, st)->d_q.p, nq, k, filter_bitmap_dev, out_dists_dev, out_rows_dev, out_counts_dev, st);
}

int mlv_filter_create(mlv_index_t h, const uint32_t* bitmap, uint64_t n_words, mlv_filter_t* out) {
    if (!h || !out || (!bitmap && n_words)) return MLV_E_INVALID;
    *out = nullptr;
    DeviceGuard g(h->device);
    mlv_filter* f = new (std::nothrow) mlv_filter();
    if (!f) return MLV_E_NOMEM;
    f->owner = h;
    f->compact_gen = h->compact_gen;
    f->bitmap_words = n_words;
    int rc = ensure_dev(h, f->d_bitmap, std::max<uint64_t>(n_words, 1) * 4);
    if (rc == MLV_OK && n_words) {
        cudaError_t e = cudaMemcpyAsync(f->d_bitmap.p, bitmap, n_words * 4, cudaMemcpyHostToDevice, h->stream);
        if (e != cudaSuccess) rc = fail_cuda(h, e, "filter upload");
    }
    if (rc == MLV_OK && h->rows) rc = build_gather_list(h, f->d_list, f->d_scratch, (const uint32_t*)f->d_bitmap.p, n_words, h->stream);
    uint64_t total = 0;
    if (rc == MLV_OK && h->rows) {
        cudaError_t e = cudaMemcpyAsync(&total, f->d_scratch.p, 8, cudaMemcpyDeviceToHost, h->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
        if (e != cudaSuccess) rc = fail_cuda(h, e, "filter build");
    }
    if (rc != MLV_OK) {
        for (DevBuf* b : {&f->d_bitmap, &f->d_list, &f->d_scratch}) free_dev(*b);
        delete f;
        return rc;
    }
    f->passing = total;
    f->counted = true;
    f->epoch = h->rows ? h->epoch : ~0ull;
    *out = f;
    return MLV_OK;
}

int mlv_filter_passing(mlv_filter_t f, uint64_t* passing) {
    if (!f || !passing) return MLV_E_INVALID;
    mlv_index* h = f->owner;
    DeviceGuard g(h->device);
    if (f->epoch != h->epoch && h->rows) {
        int rc = build_gather_list(h, f->d_list, f->d_scratch, (const uint32_t*)f->d_bitmap.p, f->bitmap_words, h->stream);
        if (rc != MLV_OK) return rc;
        f->epoch = h->epoch;
        f->counted = false;
    }
    if (!f->counted && h->rows) {
        uint64_t total = 0;
        CK(h, cudaMemcpyAsync(&total, f->d_scratch.p, 8, cudaMemcpyDeviceToHost, h->stream));
        CK(h, cudaStreamSynchronize(h->stream));
        f->passing = total;
        f->counted = true;
    }
    *passing = h->rows ? f->passing : 0;
    return MLV_OK;
}

int mlv_filter_destroy(mlv_filter_t f) {
    if (!f) return MLV_E_INVALID;
    mlv_index* h = f->owner;
    DeviceGuard g(h->device);
    if (h->bound_filter == f) h->bound_filter = nullptr;
    cudaDeviceSynchronize();
    for (DevBuf* b : {&f->d_bitmap, &f->d_list, &f->d_scratch}) free_dev(*b);
    delete f;
    return MLV_OK;
}

int mlv_index_set_filter(mlv_index_t h, mlv_filter_t f) {
    if (!h || (f && f->owner != h)) return MLV_E_INVALID;
    h->bound_filter = f;
    return MLV_OK;
}

int mlv_exchange_create(int device, uint32_t world, uint32_t rank, mlv_exchange_t* out, unsigned char* handle_out) {
    if (!out || !handle_out || world == 0 || world > XCHG_MAX_WORLD || rank >= world) return MLV_E_INVALID;
    *out = nullptr;
    int ndev = mlv_device_count();
    if (ndev <= 0) return MLV_E_NO_DEVICE;
    if (device < 0 || device >= ndev) return MLV_E_INVALID;
    DeviceGuard g(device);
    mlv_exchange* x = new (std::nothrow) mlv_exchange();
    if (!x) return MLV_E_NOMEM;
    x->device = device;
    x->world = world;
    x->rank = rank;
    static_assert(sizeof(cudaIpcMemHandle_t) == MLV_EXCHANGE_HANDLE_BYTES, "IPC handle size");
    cudaIpcMemHandle_t hd;
    void* buf = nullptr;
    cudaError_t e = cudaMalloc(&buf, (size_t)XCHG_WORDS * 8 + 64);
    if (e == cudaSuccess) e = cudaMemset(buf, 0, (size_t)XCHG_WORDS * 8 + 64);
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&hd, buf);
    if (e != cudaSuccess) {
        cudaGetLastError();
        if (buf) cudaFree(buf);
        delete x;
        return MLV_E_CUDA;
    }
    x->bufs[rank] = (uint64_t*)buf;
    x->d_error = (int*)((uint64_t*)buf + XCHG_WORDS);  // behind the slots, in the same allocation
    memcpy(handle_out, &hd, sizeof(hd));
    x->connected = world == 1;
    *out = x;
    return MLV_OK;
}

int mlv_exchange_connect(mlv_exchange_t x, const unsigned char* all_handles) {
    if (!x || !all_handles) return MLV_E_INVALID;
    DeviceGuard g(x->device);
    for (uint32_t r = 0; r < x->world; r++) {
        if (r == x->rank || x->bufs[r]) continue;
        cudaIpcMemHandle_t hd;
        memcpy(&hd, all_handles + (size_t)r * sizeof(hd), sizeof(hd));
        void* ptr = nullptr;
        if (cudaIpcOpenMemHandle(&ptr, hd, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
            cudaGetLastError();
            return MLV_E_CUDA;
        }
        x->bufs[r] = (uint64_t*)ptr;
    }
    x->connected = true;
    return MLV_OK;
}

int mlv_exchange_check(mlv_exchange_t x) {
    if (!x) return MLV_E_INVALID;
    DeviceGuard g(x->device);
    int err = 0;
    if (cudaMemcpy(&err, x->d_error, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) {
        cudaGetLastError();
        return MLV_E_CUDA;
    }
    return err ? MLV_E_CUDA : MLV_OK;
}

int mlv_exchange_destroy(mlv_exchange_t x) {
    if (!x) return MLV_E_INVALID;
    DeviceGuard g(x->device);
    cudaDeviceSynchronize();
    for (uint32_t r = 0; r < x->world; r++) {
        if (!x->bufs[r]) continue;
        if (r == x->rank)
            cudaFree(x->bufs[r]);
        else
            cudaIpcCloseMemHandle(x->bufs[r]);
    }
    cudaGetLastError();
    delete x;
    return MLV_OK;
}

int mlv_index_attach_exchange(mlv_index_t h, mlv_exchange_t x, const uint64_t* row_bases) {
    if (!h) return MLV_E_INVALID;
    if (!x) {
        h->xchg = nullptr;
        return MLV_OK;
    }
    if (!row_bases || x->device != h->device) return fail(h, MLV_E_INVALID, "exchange lives on another device");
    h->xchg = x;
    for (uint32_t i = 0; i < x->world; i++) h->xchg_row_bases[i] = row_bases[i];
    return MLV_OK;
}

int mlv_index_exchange_supported(mlv_index_t h, uint32_t k) {
    if (!h) return 0;
    return exchange_ok(h, k) ? 1 : 0;
}

int mlv_index_search_exchange_device(mlv_index_t h, const float* queries_dev, uint32_t nq, uint32_t k,
                                     const uint32_t* filter_bitmap_dev, float* out_dists_dev, int64_t* out_rows_dev,
                                     int32_t* out_counts_dev, void* stream) {
    if (!h || !queries_dev || !out_dists_dev || !out_rows_dev || !out_counts_dev || nq == 0 || k == 0) return fail(h, MLV_E_INVALID, "bad argument");
    if (!exchange_ok(h, k)) return fail(h, MLV_E_UNSUPPORTED, "no connected exchange, or k too large for the fused exchange");
    DeviceGuard g(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    if (h->rows == h->n_deleted) {
        // nothing to scan here, but the peers wait for this rank's (empty) lists: same launch grouping
        // as a scanning rank (groups of up to XCHG_MAX_NQ queries, one sequence number each)
        ExchangeView x{};
        fill_exchange(h, x);
        ScanCfg c;
        int rc0 = choose_cfg(h, nq, k, false, &c);  // same query grouping as a scanning rank
        if (rc0 != MLV_OK) return rc0;
        for (uint32_t g0 = 0; g0 < nq; g0 += c.NQ) {
            const uint32_t n = std::min<uint32_t>(c.NQ, nq - g0);
            x.seq = ++h->xseq;
            exchange_only_kernel<<<1, 256, 0, st>>>(x, n, k, out_dists_dev + (size_t)g0 * k, out_rows_dev + (size_t)g0 * k,
                                                   out_counts_dev + g0);
            h->launches++;
        }
        CK(h, cudaGetLastError());
        return MLV_OK;
    }
    int rc = prep_queries(h, queries_dev, nq, st);
    if (rc != MLV_OK) return rc;
    return search_prepared(h, (const float*)lane_for(h, st)->d_q.p, nq, k, filter_bitmap_dev, out_dists_dev, out_rows_dev, out_counts_dev, st, true);
}

static int search_host_common(mlv_index_t h, const float* queries, uint32_t nq, uint32_t k, const uint32_t* filter_bitmap,
                              float* out_dists, int64_t* out_rows, int32_t* out_counts, bool exchange) {
    if (!h || !queries || !out_dists || !out_rows || !out_counts || nq == 0 || k == 0) return fail(h, MLV_E_INVALID, "bad argument");
    if (k > MLV_MAX_K) return fail(h, MLV_E_UNSUPPORTED, "k exceeds MLV_MAX_K");
    DeviceGuard g(h->device);
    const size_t qbytes = (size_t)nq * h->dim * 4;
    const size_t nk = (size_t)nq * k;
    const size_t out_bytes = nk * 4 + nk * 8 + (size_t)nq * 4;
    int rc;
    if ((rc = ensure_host(h, h->h_stage, std::max(qbytes, out_bytes))) != MLV_OK) return rc;
    if ((rc = ensure_dev(h, h->d_qraw, qbytes)) != MLV_OK) return rc;
    // one device block for the results (rows | dists | counts): a single copy brings them back
    if ((rc = ensure_dev(h, h->d_outr, out_bytes)) != MLV_OK) return rc;
    int64_t* d_rows_out = (int64_t*)h->d_outr.p;
    float* d_dists_out = (float*)((char*)h->d_outr.p + nk * 8);
    int32_t* d_counts_out = (int32_t*)((char*)h->d_outr.p + nk * 12);
    const uint32_t* filter_dev = nullptr;
    if (filter_bitmap && h->rows) {
        const size_t fb = ((h->rows + 31) / 32) * 4;
        if ((rc = ensure_dev(h, h->d_filter, fb)) != MLV_OK) return rc;
        CK(h, cudaMemcpyAsync(h->d_filter.p, filter_bitmap, fb, cudaMemcpyHostToDevice, h->stream));
        filter_dev = (const uint32_t*)h->d_filter.p;
    }
    memcpy(h->h_stage.p, queries, qbytes);
    CK(h, cudaMemcpyAsync(h->d_qraw.p, h->h_stage.p, qbytes, cudaMemcpyHostToDevice, h->stream));
    if (exchange)
        rc = mlv_index_search_exchange_device(h, (const float*)h->d_qraw.p, nq, k, filter_dev, d_dists_out, d_rows_out,
                                              d_counts_out, h->stream);
    else
        rc = mlv_index_search_device(h, (const float*)h->d_qraw.p, nq, k, filter_dev, d_dists_out, d_rows_out, d_counts_out,
                                     h->stream);
    if (rc != MLV_OK) return rc;
    char* hs = (char*)h->h_stage.p;
    CK(h, cudaMemcpyAsync(hs, h->d_outr.p, out_bytes, cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    memcpy(out_rows, hs, nk * 8);
    memcpy(out_dists, hs + nk * 8, nk * 4);
    memcpy(out_counts, hs + nk * 12, (size_t)nq * 4);
    return MLV_OK;
}

int mlv_index_search(mlv_index_t h, const float* queries, uint32_t nq, uint32_t k, const uint32_t* filter_bitmap,
                     float* out_dists, int64_t* out_rows, int32_t* out_counts) {
    return search_host_common(h, queries, nq, k, filter_bitmap, out_dists, out_rows, out_counts, false);
}

int mlv_index_search_exchange(mlv_index_t h, const float* queries, uint32_t nq, uint32_t k, const uint32_t* filter_bitmap,
                              float* out_dists, int64_t* out_rows, int32_t* out_counts) {
    return search_host_common(h, queries, nq, k, filter_bitmap, out_dists, out_rows, out_counts, true);
}

int mlv_index_range_search_device(mlv_index_t h, const float* queries_dev, uint32_t nq, float radius,
                                  const uint32_t* filter_bitmap_dev, uint64_t max_hits, float* out_dists_dev,
                                  int64_t* out_rows_dev, uint64_t* out_counts_dev, void* stream) {
    if (!h || !queries_dev || !out_counts_dev || nq == 0 || (max_hits && (!out_dists_dev || !out_rows_dev))) return fail(h, MLV_E_INVALID, "bad argument");
    DeviceGuard g(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    if (h->rows == h->n_deleted) {
        CK(h, cudaMemsetAsync(out_counts_dev, 0, (size_t)nq * 8, st));
        return MLV_OK;
    }
    const uint64_t slots = std::max<uint64_t>(max_hits, 1);
    int rc;
    // the hit keys of one call: handle-level scratch (one range search at a time per handle)
    if ((rc = ensure_dev(h, h->d_range, (size_t)nq * slots * 8 + (size_t)nq * 8)) != MLV_OK) return rc;
    unsigned long long* d_counts = (unsigned long long*)h->d_range.p;
    uint64_t* d_keys = (uint64_t*)h->d_range.p + nq;
    CK(h, cudaMemsetAsync(d_counts, 0, (size_t)nq * 8, st));
    if ((rc = prep_queries(h, queries_dev, nq, st)) != MLV_OK) return rc;
    Lane* ln = lane_for(h, st);
    FilterPlan fp;
    if ((rc = plan_filter(h, ln, filter_bitmap_dev, st, &fp)) != MLV_OK) return rc;
    ScanCfg c;
    if ((rc = choose_cfg(h, 1, 1, true, &c, fp.gather != nullptr)) != MLV_OK) return rc;
    ScanParams p{};
    p.rows = reinterpret_cast<const float4*>(h->d_rows);
    p.n_rows = (uint32_t)h->rows;
    p.ld4 = h->ld / 4;
    p.tile_rows = c.T;
    p.n_tiles = (uint32_t)((h->rows + c.T - 1) / c.T);
    p.stages = c.S;
    p.producer_warps = (uint32_t)c.PW;
    p.stage_f4 = c.stage_f4;
    p.k = 1;
    p.live = h->n_deleted ? h->d_live : nullptr;
    p.filter = fp.bitmap;
    p.gather = fp.gather;
    p.n_rows_dev = fp.n_rows_dev;
    p.evict_first = c.evict_first;
    p.radius = radius;
    p.max_hits = max_hits;
    if ((rc = ensure_sched(h, ln)) != MLV_OK) return rc;
    fill_sched(h, ln, p);
    for (uint32_t q = 0; q < nq; q++) {
        p.queries = reinterpret_cast<const float4*>((const float*)ln->d_q.p + (size_t)q * h->ld);
        p.nq_valid = 1;
        p.range_counts = d_counts + q;
        p.range_keys = d_keys + (size_t)q * slots;
        CK(h, launch_scan(h, p, c, true, st));
    }
    CK(h, cudaFuncSetAttribute(range_finish_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SELECT_MAX_P * 8)));
    range_finish_kernel<<<nq, SELECT_THREADS, (size_t)SELECT_MAX_P * 8, st>>>(d_keys, d_counts, slots, max_hits, h->row_base, out_dists_dev,
                                                                            out_rows_dev, (unsigned long long*)out_counts_dev);
    h->launches++;
    CK(h, cudaGetLastError());
    return MLV_OK;
}

int mlv_index_range_search(mlv_index_t h, const float* queries, uint32_t nq, float radius, const uint32_t* filter_bitmap,
                           uint64_t max_hits, float* out_dists, int64_t* out_rows, uint64_t* out_counts) {
    if (!h || !queries || !out_counts || nq == 0 || (max_hits && (!out_dists || !out_rows))) return fail(h, MLV_E_INVALID, "bad argument");
    DeviceGuard g(h->device);
    for (uint32_t q = 0; q < nq; q++) out_counts[q] = 0;
    if (h->rows == h->n_deleted) return MLV_OK;
    const size_t qbytes = (size_t)nq * h->dim * 4;
    const size_t nh = (size_t)nq * max_hits;
    int rc;
    if ((rc = ensure_dev(h, h->d_qraw, qbytes)) != MLV_OK) return rc;
    if ((rc = ensure_dev(h, h->d_outd, std::max<size_t>(nh, 1) * 4)) != MLV_OK) return rc;
    if ((rc = ensure_dev(h, h->d_outr, std::max<size_t>(nh, 1) * 8)) != MLV_OK) return rc;
    if ((rc = ensure_dev(h, h->d_outc, (size_t)nq * 8)) != MLV_OK) return rc;
    const uint32_t* filter_dev = nullptr;
    if (filter_bitmap) {
        const size_t fb = ((h->rows + 31) / 32) * 4;
        if ((rc = ensure_dev(h, h->d_filter, fb)) != MLV_OK) return rc;
        CK(h, cudaMemcpyAsync(h->d_filter.p, filter_bitmap, fb, cudaMemcpyHostToDevice, h->stream));
        filter_dev = (const uint32_t*)h->d_filter.p;
    }
    CK(h, cudaMemcpyAsync(h->d_qraw.p, queries, qbytes, cudaMemcpyHostToDevice, h->stream));
    rc = mlv_index_range_search_device(h, (const float*)h->d_qraw.p, nq, radius, filter_dev, max_hits, (float*)h->d_outd.p,
                                       (int64_t*)h->d_outr.p, (uint64_t*)h->d_outc.p, h->stream);
    if (rc != MLV_OK) return rc;
    CK(h, cudaMemcpyAsync(out_counts, h->d_outc.p, (size_t)nq * 8, cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    const uint64_t slots = std::max<uint64_t>(max_hits, 1);
    for (uint32_t q = 0; q < nq; q++) {
        const uint64_t got = std::min<uint64_t>(out_counts[q], max_hits);
        if (!got) continue;
        float* dd = (float*)h->d_outd.p + (size_t)q * max_hits;
        int64_t* dr = (int64_t*)h->d_outr.p + (size_t)q * max_hits;
        if (got > SELECT_MAX_P) {  // more hits than one CTA sorts: global bitonic network over the query's keys
            uint64_t* keys = (uint64_t*)h->d_range.p + nq + (size_t)q * slots;
            if ((rc = sort_big_device(h, keys, got, dd, dr, h->stream)) != MLV_OK) return rc;
        }
        CK(h, cudaMemcpyAsync(out_dists + (size_t)q * max_hits, dd, got * 4, cudaMemcpyDeviceToHost, h->stream));
        CK(h, cudaMemcpyAsync(out_rows + (size_t)q * max_hits, dr, got * 8, cudaMemcpyDeviceToHost, h->stream));
    }
    CK(h, cudaStreamSynchronize(h->stream));
    return MLV_OK;
}

int mlv_index_get_rows(mlv_index_t h, const uint64_t* rows, uint64_t n, float* out) {
    if (!h || (n && (!rows || !out))) return MLV_E_INVALID;
    DeviceGuard g(h->device);
    for (uint64_t i = 0; i < n; i++) {
        if (rows[i] >= h->rows) return fail(h, MLV_E_INVALID, "row out of range");
        CK(h, cudaMemcpyAsync(out + i * h->dim, h->d_rows + rows[i] * h->ld, (size_t)h->dim * 4, cudaMemcpyDeviceToHost, h->stream));
    }
    CK(h, cudaStreamSynchronize(h->stream));
    return MLV_OK;
}

int mlv_index_info(mlv_index_t h, mlv_index_info_t* info) {
    if (!h || !info) return MLV_E_INVALID;
    info->rows = h->rows;
    info->live = h->rows - h->n_deleted;
    info->capacity = h->capacity;
    info->row_base = h->row_base;
    size_t b = (size_t)h->capacity * h->ld * 4 + (size_t)h->live_words * 4;
    for (const DevBuf* d : {&h->d_qraw, &h->d_filter, &h->d_outd, &h->d_outr, &h->d_outc, &h->d_misc, &h->d_range, &h->d_norms,
                            &h->d_gq, &h->d_cand, &h->d_maxn2})
        b += d->bytes;
    for (const Lane& l : h->lanes)
        for (const DevBuf* d : {&l.d_q, &l.d_keys0, &l.d_keys1, &l.d_sched, &l.d_flist, &l.d_fscratch}) b += d->bytes;
    info->device_bytes = b;
    info->dim = h->dim;
    info->ld = h->ld;
    info->metric = h->metric;
    info->device = h->device;
    return MLV_OK;
}

int mlv_merge_topk(int device, const float* dists_dev, const int64_t* rows_dev, uint32_t n_lists, uint32_t nq, uint32_t k,
                   float* out_dists_dev, int64_t* out_rows_dev, int32_t* out_counts_dev, void* stream) {
    if (!dists_dev || !rows_dev || !out_dists_dev || !out_rows_dev || !out_counts_dev || !n_lists || !nq || !k) return MLV_E_INVALID;
    if ((uint64_t)n_lists * k > SELECT_MAX_P) return MLV_E_UNSUPPORTED;
    DeviceGuard g(device);
    if (!g.ok) return MLV_E_CUDA;
    if (ensure_select_attrs(device) != cudaSuccess) return MLV_E_CUDA;
    MergePairsParams p{};
    p.dists = dists_dev;
    p.rows = rows_dev;
    p.n_lists = n_lists;
    p.nq = nq;
    p.k = k;
    p.P = pow2_ceil(std::max<uint32_t>(n_lists * k, 2));
    p.out_dists = out_dists_dev;
    p.out_rows = out_rows_dev;
    p.out_counts = out_counts_dev;
    const int threads = (int)std::min<uint32_t>(SELECT_THREADS, std::max<uint32_t>(p.P / 2, 32));
    merge_pairs_kernel<<<nq, threads, (size_t)p.P * 8, (cudaStream_t)stream>>>(p);
    return cudaGetLastError() == cudaSuccess ? MLV_OK : MLV_E_CUDA;
}

int mlv_index_set_timing(mlv_index_t h, int enabled) {
    if (!h) return MLV_E_INVALID;
    h->timing = enabled != 0;
    return MLV_OK;
}

int mlv_index_scan_time_ms(mlv_index_t h, double* total_ms, uint64_t* launches) {
    if (!h) return MLV_E_INVALID;
    DeviceGuard g(h->device);
    double total = 0;
    uint64_t n = 0;
    for (auto& pr : h->pending) {
        CK(h, cudaEventSynchronize(pr.second));
        float ms = 0;
        CK(h, cudaEventElapsedTime(&ms, pr.first, pr.second));
        total += ms;
        n++;
        h->event_pool.push_back(pr.first);
        h->event_pool.push_back(pr.second);
    }
    h->pending.clear();
    if (total_ms) *total_ms = total;
    if (launches) *launches = n;
    return MLV_OK;
}

int mlv_index_gemm_stats(mlv_index_t h, mlv_gemm_stats_t* out) {
    if (!h || !out) return MLV_E_INVALID;
    DeviceGuard g(h->device);
    double total = 0;
    uint64_t n = 0;
    for (auto& pr : h->gemm_pending) {
        CK(h, cudaEventSynchronize(pr.second));
        float ms = 0;
        CK(h, cudaEventElapsedTime(&ms, pr.first, pr.second));
        total += ms;
        n++;
        h->event_pool.push_back(pr.first);
        h->event_pool.push_back(pr.second);
    }
    h->gemm_pending.clear();
    out->gemm_ms = total;
    out->gemm_launches_timed = n;
    out->searches = h->gemm_searches;
    out->queries = h->gemm_queries;
    out->fallback_queries = h->gemm_fallback_queries;
    out->rounds = h->gemm_rounds;
    return MLV_OK;
}

int mlv_index_debug_gemm(mlv_index_t h, const float* queries, uint32_t nq, float* out_approx) {
    if (!h || !queries || !out_approx || nq == 0) return fail(h, MLV_E_INVALID, "bad argument");
    if (h->rows == 0 || h->rows > SELECT_MAX_P || h->ld < (uint32_t)GEMM_BK) return fail(h, MLV_E_UNSUPPORTED, "debug_gemm needs 1..8192 rows and dim >= 32");
    DeviceGuard g(h->device);
    cudaStream_t st = h->stream;
    int rc;
    const uint32_t ld = h->ld;
    const size_t qbytes = (size_t)nq * h->dim * 4;
    if ((rc = ensure_dev(h, h->d_qraw, qbytes)) != MLV_OK) return rc;
    CK(h, cudaMemcpyAsync(h->d_qraw.p, queries, qbytes, cudaMemcpyHostToDevice, st));
    if ((rc = prep_queries(h, (const float*)h->d_qraw.p, nq, st)) != MLV_OK) return rc;
    if (h->metric != MLV_COSINE && (rc = ensure_row_norms(h, st)) != MLV_OK) return rc;
    const uint32_t GEMM_BN = h->tune_gemm_bn == 128 ? 128 : 256;
    const uint32_t nq_pad = (nq + GEMM_BN - 1) / GEMM_BN * GEMM_BN;
    const uint32_t cap = pow2_ceil((uint32_t)h->rows);
    const size_t qmat = (size_t)nq_pad * ld * 4;
    if ((rc = ensure_dev(h, h->d_gq, 2 * qmat + (size_t)nq_pad * 16)) != MLV_OK) return rc;
    if ((rc = ensure_dev(h, h->d_cand, (size_t)nq_pad * cap * 8)) != MLV_OK) return rc;
    float* qhi = (float*)h->d_gq.p;
    float* qlo = qhi + (size_t)nq_pad * ld;
    float* qn = qlo + (size_t)nq_pad * ld;
    float* thr = qn + nq_pad;
    uint32_t* cnt = (uint32_t*)(thr + nq_pad);
    uint32_t* flags = cnt + nq_pad;
    split_queries_kernel<<<(nq_pad + 7) / 8, 256, 0, st>>>((const float*)lane_for(h, st)->d_q.p, qhi, qlo, qn, thr, cnt, flags, nq, nq_pad, ld);
    CK(h, cudaGetLastError());
    CUtensorMap mx, mqh, mql;
    if ((rc = make_tile_map(h, &mx, h->d_rows, h->rows, GEMM_BM)) != MLV_OK) return rc;
    if ((rc = make_tile_map(h, &mqh, qhi, nq_pad, GEMM_BN)) != MLV_OK) return rc;
    if ((rc = make_tile_map(h, &mql, qlo, nq_pad, GEMM_BN)) != MLV_OK) return rc;
    GemmParams gp{};
    gp.n_rows = (uint32_t)h->rows;
    gp.nq = nq;
    gp.n_qtiles = nq_pad / GEMM_BN;
    gp.n_kchunks = (ld + GEMM_BK - 1) / GEMM_BK;
    gp.row_norms = h->metric == MLV_L2 ? (const float*)h->d_norms.p : nullptr;
    gp.q_norms = qn;
    gp.thr = thr;
    gp.live = h->n_deleted ? h->d_live : nullptr;
    gp.cand = (uint64_t*)h->d_cand.p;
    gp.cand_cnt = cnt;
    gp.cap = cap;
    gp.row_tile0 = 0;
    gp.row_tile1 = (uint32_t)((h->rows + GEMM_BM - 1) / GEMM_BM);
    const int grid = (int)std::min<uint64_t>((uint64_t)gp.row_tile1 * gp.n_qtiles, (uint64_t)h->sm_count);
    CK(h, h->metric == MLV_L2 ? launch_gemm_t<METRIC_L2>(mx, mqh, mql, gp, grid, st, (int)GEMM_BN)
                              : launch_gemm_t<METRIC_IP>(mx, mqh, mql, gp, grid, st, (int)GEMM_BN));
    h->launches += 3;
    std::vector<uint64_t> keys((size_t)nq * cap);
    std::vector<uint32_t> counts(nq);
    CK(h, cudaMemcpyAsync(keys.data(), h->d_cand.p, keys.size() * 8, cudaMemcpyDeviceToHost, st));
    CK(h, cudaMemcpyAsync(counts.data(), cnt, (size_t)nq * 4, cudaMemcpyDeviceToHost, st));
    CK(h, cudaStreamSynchronize(st));
    for (size_t i = 0; i < (size_t)nq * h->rows; i++) out_approx[i] = std::numeric_limits<float>::quiet_NaN();
    for (uint32_t q = 0; q < nq; q++)
        for (uint32_t i = 0; i < std::min(counts[q], cap); i++) {
            const uint64_t key = keys[(size_t)q * cap + i];
            if (key_row(key) < h->rows) out_approx[(size_t)q * h->rows + key_row(key)] = key_dist(key);
        }
    return MLV_OK;
}

int mlv_index_debug_timeline(mlv_index_t h, uint64_t* out, uint32_t max_ctas, uint32_t* n_ctas) {
    if (!h || !out || !n_ctas) return MLV_E_INVALID;
    DeviceGuard g(h->device);
    const uint32_t n = std::min<uint32_t>(max_ctas, (uint32_t)h->last_grid);
    *n_ctas = n;
    if (n == 0 || !h->d_timeline.p) return MLV_OK;
    CK(h, cudaDeviceSynchronize());
    CK(h, cudaMemcpy(out, h->d_timeline.p, (size_t)n * 4 * 8, cudaMemcpyDeviceToHost));
    return MLV_OK;
}

int mlv_index_kernel_launches(mlv_index_t h, uint64_t* launches) {
    if (!h || !launches) return MLV_E_INVALID;
    *launches = h->launches;
    return MLV_OK;
}

}  // extern "C"
