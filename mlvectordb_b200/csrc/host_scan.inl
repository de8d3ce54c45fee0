// host_scan.inl -- host side of the scan path: launch shape, per-stream lanes, filter plans, scheduler / fused tail / exchange set-up, search_prepared, range ordering.
// Part of the single translation unit mlv_index.cu (included there, in order).
#pragma once

namespace {

// ---- scan configuration ------------------------------------------------------------------------
// Shape of one scan launch.  Defaults come from B200 sweeps (profiles/r01_sweep_*.jsonl): short rows
// are latency-bound in the consumers, so they get more warps, more rows per warp step and larger
// stages; a gathering producer is bound by its copy issue rate (~80 cycles per row copy per warp),
// so short rows get more producer warps.
// slots of one (warp, query) candidate list: k itself up to 32 (replace-max insertion), else an append
// buffer of the next power of two >= 2k (scan_kernel.cuh, ScanParams::list_cap)
uint32_t scan_list_cap(uint32_t k) {
    if (k <= 32) return k;
    uint32_t p = 64;
    while (p < 2 * k) p <<= 1;
    return p;
}

uint32_t f16_ld(const mlv_index* h);
int ensure_f16_shadow(mlv_index* h, cudaStream_t st, bool* usable);
int ensure_row_norms(mlv_index* h, cudaStream_t st);

// half: the launch reads the fp16 shadow (rows of f16_ld halves; the query stays fp32 in shared memory); 2 = with the
// tensor-core consumers (16 rows per warp step, the query a second time as halves)
int choose_cfg(mlv_index* h, uint32_t nq, uint32_t k, bool range, ScanCfg* c, bool gather = false, int half = 0) {
    const uint32_t ld4 = half ? f16_ld(h) / 8 : h->ld / 4;                    // 16-byte units per row
    const size_t rowbytes = (size_t)ld4 * 16;
    const size_t qbytes = half ? (size_t)f16_ld(h) * 4 + (half == 2 ? rowbytes : 0) : rowbytes;   // one query in shared memory
    const uint32_t lcap = scan_list_cap(k);
    int CW = h->tune_cw > 0 ? std::min(h->tune_cw, half == 2 ? 7 : SCAN_MAX_CW) : (half == 2 ? 4 : (ld4 <= 64 ? 16 : 8));
    if (!range && k > 512) CW = std::min(CW, 4);  // 2048-slot buffers: 4 warps keep the lists at 64 KB
    int R = h->tune_r ? h->tune_r : (ld4 <= 256 ? 4 : (ld4 <= 512 ? 2 : 1));
    if (R != 1 && R != 2 && R != 4) R = 1;
    if (half == 2) R = 16;
    if (half == 1) CW = std::min(CW, SCAN_WIDE_CW);   // launch bounds of scan_kernel_half
    int NQ = 1;
    if (!range) {
        while (NQ < 8 && (uint32_t)NQ < nq) NQ <<= 1;
        while (NQ > 1 && ((size_t)NQ * lcap * CW * 8 > 65536 || R * NQ > 32)) NQ >>= 1;
        if (NQ >= 2) CW = std::min(CW, SCAN_WIDE_CW);   // launch bounds of the multi-query instantiations (scan_max_threads)
    }
    const int max_stages = std::min(std::max(h->tune_max_stages, 2), 16);
    const size_t fixed = (size_t)NQ * qbytes + (range ? 0 : (size_t)CW * NQ * lcap * 8) + (size_t)max_stages * 24 + 256;
    if (fixed + 2 * rowbytes > h->smem_optin)
        return fail(h, MLV_E_UNSUPPORTED, "dimension too large for the shared-memory ring of this build");
    const size_t avail = h->smem_optin - fixed;
    const size_t group_bytes = (size_t)R * CW * rowbytes;
    const size_t target = (size_t)(h->tune_stage_kb > 0 ? h->tune_stage_kb : (ld4 <= 32 ? 64 : 32)) * 1024;
    uint64_t m = std::max<uint64_t>(1, target / group_bytes);
    uint64_t T = (uint64_t)R * CW * m;
    // tensor-core consumers: a warp takes 16 rows at a time and two warps keep up with the copies, so a tile is a
    // multiple of 16 rows (~48 KB), not of 16 * CW
    if (half == 2) T = 16 * std::max<uint64_t>(1, ((size_t)(h->tune_stage_kb > 0 ? h->tune_stage_kb : 48) * 1024) / (16 * rowbytes));
    if (T * rowbytes * 2 > avail) {
        T = (avail / 2 / rowbytes) / R * R;
        if (T == 0) {
            R = 1;
            T = avail / 2 / rowbytes;
        }
        if (T == 0) return fail(h, MLV_E_UNSUPPORTED, "dimension too large for the shared-memory ring of this build");
    }
    {
        // small matrices: at least ~2 tiles per SM, otherwise most SMs idle while a few walk several tiles
        const uint64_t want_tiles = 2ull * (uint64_t)(h->tune_ctas > 0 ? h->tune_ctas : h->sm_count);
        if (h->rows && (h->rows + T - 1) / T < want_tiles) {
            uint64_t t2 = (h->rows + want_tiles - 1) / want_tiles;
            t2 = std::max<uint64_t>((t2 + R - 1) / R * R, (uint64_t)R * 8);  // >= 8 warp steps of R rows per tile
            T = std::min(T, t2);
        }
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
    c->smem = (size_t)S * stage + (size_t)NQ * qbytes + (range ? 0 : (size_t)CW * NQ * lcap * 8) + (size_t)S * 24;
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
    // the opt-in shared-memory limit is per function and device: raise it once per (instantiation, device)
    static size_t raised[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || raised[dev] < c.smem) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c.smem);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) raised[dev] = c.smem;
    }
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

template <int METRIC, int R, bool RANGE>
cudaError_t launch_scan_half_t(const ScanParams& p, const ScanCfg& c, cudaStream_t st) {
    auto kern = scan_kernel_half<METRIC, R, RANGE>;
    static size_t raised[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || raised[dev] < c.smem) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c.smem);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) raised[dev] = c.smem;
    }
    kern<<<c.grid, c.threads, c.smem, st>>>(p);
    return cudaGetLastError();
}
template <int METRIC, bool RANGE>
cudaError_t launch_scan_half_mma_t(const ScanParams& p, const ScanCfg& c, cudaStream_t st) {
    auto kern = scan_kernel_half_mma<METRIC, RANGE>;
    static size_t raised[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || raised[dev] < c.smem) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c.smem);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) raised[dev] = c.smem;
    }
    kern<<<c.grid, c.threads, c.smem, st>>>(p);
    return cudaGetLastError();
}
template <int METRIC, bool RANGE = false>
cudaError_t launch_scan_half_m(const ScanParams& p, const ScanCfg& c, cudaStream_t st) {
    if (c.R == 16) return launch_scan_half_mma_t<METRIC, RANGE>(p, c, st);
    if (c.R == 1) return launch_scan_half_t<METRIC, 1, RANGE>(p, c, st);
    if (c.R == 2) return launch_scan_half_t<METRIC, 2, RANGE>(p, c, st);
    if (c.R == 4) return launch_scan_half_t<METRIC, 4, RANGE>(p, c, st);
    return cudaErrorInvalidValue;
}

template <int METRIC, int R, bool RANGE>
cudaError_t launch_scan_iq_t(const ScanParams& p, const InlineQuery& iq, const ScanCfg& c, cudaStream_t st) {
    auto kern = scan_kernel_iq<METRIC, R, RANGE>;
    static size_t raised[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || raised[dev] < c.smem) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c.smem);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) raised[dev] = c.smem;
    }
    kern<<<c.grid, c.threads, c.smem, st>>>(p, iq);
    return cudaGetLastError();
}
template <int METRIC, bool RANGE>
cudaError_t launch_scan_iq_m(const ScanParams& p, const InlineQuery& iq, const ScanCfg& c, cudaStream_t st) {
    if (c.R == 1) return launch_scan_iq_t<METRIC, 1, RANGE>(p, iq, c, st);
    if (c.R == 2) return launch_scan_iq_t<METRIC, 2, RANGE>(p, iq, c, st);
    if (c.R == 4) return launch_scan_iq_t<METRIC, 4, RANGE>(p, iq, c, st);
    return cudaErrorInvalidValue;
}
// one query, raw values in the launch parameters (c.NQ == 1)
cudaError_t launch_scan_iq(mlv_index* h, const ScanParams& p, const InlineQuery& iq, const ScanCfg& c, bool range, cudaStream_t st) {
    const bool l2 = h->metric == MLV_L2;
    h->launches++;
    if (range) return l2 ? launch_scan_iq_m<METRIC_L2, true>(p, iq, c, st) : launch_scan_iq_m<METRIC_IP, true>(p, iq, c, st);
    return l2 ? launch_scan_iq_m<METRIC_L2, false>(p, iq, c, st) : launch_scan_iq_m<METRIC_IP, false>(p, iq, c, st);
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
    if (lru->used && cudaStreamSynchronize(lru->stream) != cudaSuccess)
        cudaGetLastError();  // the previous owner destroyed its stream: nothing of it can still be running
    lru->used = true;
    lru->stream = st;
    lru->last_use = ++h->lane_clock;
    return lru;
}

// Build the ascending list of rows that are live AND pass `bm` (device bitmap, `words` words; rows
// beyond it do not pass) into list/scratch, on `st`.  scratch[0] (u64) receives the list length.
int build_gather_list(mlv_index* h, DevBuf& list, DevBuf& scratch, const uint32_t* bm, uint64_t words, cudaStream_t st) {
    const uint64_t n = h->rows, n_words = (n + 31) / 32;
    const unsigned blocks = (unsigned)std::max<uint64_t>((n_words + LIST_BLOCK_WORDS - 1) / LIST_BLOCK_WORDS, 1);
    int rc;
    if ((rc = ensure_dev(h, scratch, 8 + (size_t)blocks * 4)) != MLV_OK) return rc;
    if ((rc = ensure_dev(h, list, std::max<uint64_t>(n, 1) * 4)) != MLV_OK) return rc;
    uint64_t* d_total = (uint64_t*)scratch.p;                 // [total u64][block popcounts u32 ...]
    uint32_t* d_block_sums = (uint32_t*)(d_total + 1);
    passing_block_sums_kernel<<<blocks, LIST_BLOCK_WORDS, 0, st>>>(h->d_live, bm, words, n, d_block_sums);
    scatter_passing_rows_kernel<<<blocks, LIST_BLOCK_WORDS, 0, st>>>(h->d_live, bm, words, n, d_block_sums, d_total,
                                                                    (uint32_t*)list.p);
    h->launches += 2;
    CK(h, cudaGetLastError());
    return MLV_OK;
}

// What a search reads its rows through: nothing special, a bitmap checked per row, or a gather list.
struct FilterPlan {
    const uint32_t* bitmap = nullptr;      // stream + mask
    const uint32_t* gather = nullptr;      // row list (live AND passing)
    const uint32_t* n_rows_dev = nullptr;  // its length (device)
    uint64_t n_rows_host = 0;              // ... when the host has read it back (prepared filters), else 0
};

// filter_dev: per-call bitmap (ceil(rows/32) words) or null; a bound prepared filter applies when it is null.
// want_list: the caller needs the row list whatever the density (the tensor-core path compacts the rows with it).
int plan_filter(mlv_index* h, Lane* ln, const uint32_t* filter_dev, cudaStream_t st, FilterPlan* out, bool want_list = false) {
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
        const bool want_stream = !want_list && (h->tune_gather == 0 || (h->tune_gather < 0 && dense));
        if (want_stream && covers) {
            out->bitmap = (const uint32_t*)f->d_bitmap.p;  // stream every row, mask in the epilogue
            return MLV_OK;
        }
        out->gather = (const uint32_t*)f->d_list.p;
        out->n_rows_dev = (const uint32_t*)f->d_scratch.p;  // low word of the u64 total
        if (f->counted) out->n_rows_host = f->passing;
        return MLV_OK;
    }
    if (h->tune_gather == 0 && !want_list) {
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

void fill_sched(mlv_index* h, Lane* ln, ScanParams& p, uint64_t gathered_rows = 0) {
    p.sched = h->tune_dynamic ? (uint32_t*)ln->d_sched.p : nullptr;
    // several tiles per claim only when there are plenty of claims per SM (a gathered scan walks the list's tiles, not
    // the matrix's: 1 % of 10M x 384 is 21 tiles per SM, and claims of four left SMs idle for the last ~3 us of 50)
    uint32_t batch = (uint32_t)std::max(h->tune_tile_batch, 1);
    const uint64_t tiles = (p.gather && gathered_rows) ? (gathered_rows + p.tile_rows - 1) / p.tile_rows : (p.gather ? 0 : p.n_tiles);
    while (batch > 1 && tiles < 16ull * batch * (uint64_t)h->sm_count) batch >>= 1;
    p.tile_batch = batch;
}

// slots of the last CTA's candidate array for a fused final select
uint32_t fused_cap(const ScanCfg& c, uint32_t k) { return pow2_ceil(std::max<uint32_t>((uint32_t)c.grid * k, 1024)); }

// can the last CTA fold the whole grid's lists?  Its scratch (candidate array + final lists) overlays the CTA's dynamic
// shared memory from the start -- ring, queries, warp lists and barriers are all dead by then -- and search_prepared
// pads the allocation when a tiny ring (small dimension x small shard) is shorter than that, so the decision depends
// on k and the grid only: every rank of an exchange search takes the same one.
bool fused_ok(const mlv_index* h, const ScanCfg& c, uint32_t k) {
    if (!h->tune_dynamic || !h->tune_fused) return false;
    return (uint64_t)c.grid * k <= SCAN_FUSED_MAX_KEYS;
}
size_t fused_scratch_bytes(const ScanCfg& c, uint32_t k) { return ((size_t)fused_cap(c, k) + (size_t)c.NQ * k) * 8; }

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
    x.timeout_ns = e->timeout_ns;
}

// What every scan launch of one search shares: the matrix, the ring shape, masks / gather list, scheduler.
ScanParams scan_params(mlv_index* h, const ScanCfg& c, const FilterPlan& fp, Lane* ln, uint32_t k) {
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
    p.list_cap = scan_list_cap(k);
    p.live = h->n_deleted ? h->d_live : nullptr;
    p.filter = fp.bitmap;
    p.gather = fp.gather;
    p.n_rows_dev = fp.n_rows_dev;
    p.evict_first = c.evict_first;
    fill_sched(h, ln, p, fp.n_rows_host);
    return p;
}

// qprep: prepared queries [nq, ld] in device memory; all output pointers in device memory.
// exchange: merge with the other ranks' results over peer memory (caller checked exchange_ok).
// inline_q: ONE raw host query carried in the launch parameters instead of qprep (nq == 1, dim <= SCAN_INLINE_MAX_DIM);
// done_flag / done_value: completion flag (mapped host memory) the fused tail writes after the outputs.  Both need the
// fused final select: *took_fast (nullable) reports whether the launch carried them; when it did not, nothing was
// launched and the caller takes the staged path.
struct FastArgs {
    const float* inline_q = nullptr;
    unsigned int* done_flag = nullptr;
    unsigned int done_value = 0;
    bool* took_fast = nullptr;
    uint4* tagged_out = nullptr;   // single GPU: tagged 16-byte result records instead of arrays + flag (scan_kernel.cuh)
};

// ---- shadow scan (scan_kernel_half) ------------------------------------------------------------------------------
// A single query reads the fp16 shadow of the rows -- half the bytes of the HBM-bound pass -- keeps k' = 32 candidates,
// and the last CTA re-scores them from the fp32 matrix in the scan's own arithmetic and certifies the answer like the
// tensor-core tiers do.  An fp32 launch is queued right behind it and returns at once unless the certificate failed
// (ScanParams::cert / run_if): the fallback is decided on the device, the call stays asynchronous and the results
// are the fp32 scan's bit for bit either way.
constexpr uint32_t HALF_SCAN_KPRIME = 32;
constexpr uint32_t HALF_SCAN_MAX_K = 16;
float gemm_delta_rel_f16_host(bool l2, uint32_t d);

// the part of the decision that every rank of an exchange search shares (it fixes the number of launches per search)
bool half_scan_proto(const mlv_index* h, uint32_t nq, uint32_t k) {
    return h->tune_scan_half != 0 && nq == 1 && k <= HALF_SCAN_MAX_K && h->tune_dynamic && h->tune_fused &&
           (uint64_t)h->sm_count * HALF_SCAN_KPRIME <= SCAN_FUSED_MAX_KEYS;
}
// ... and this rank's own: a matrix large enough for the saved bytes to matter (below ~256 MB the pass is a few tens of
// microseconds and the second launch costs more than the bytes), no per-row gather, not sitting out
// gathered_rows: 0 = not a gathered scan; UINT64_MAX = gathered, list length unknown to the host; else the list length.
// A gathered scan copies row by row: half-size rows halve the bytes per copy, not the copies, and a short list is all
// start-up and tail (10M x 384: 1 % of the rows 53 -> 100 us, 10 % 250 -> 230 us, 50 % 1083 -> 705 us), so it pays from
// 256-byte half rows up and from a list worth 1 GB of fp32 rows.
bool half_scan_shape(const mlv_index* h, uint32_t nq, uint32_t k, uint64_t gathered_rows) {
    if (!half_scan_proto(h, nq, k) || h->ld < 8) return false;
    if (gathered_rows) {
        if (!h->tune_scan_half_gather || f16_ld(h) < 128) return false;
        if (h->tune_scan_half == 1) return true;
        return gathered_rows != UINT64_MAX && gathered_rows * h->ld * 4 >= (1ull << 30);
    }
    return h->tune_scan_half == 1 || (uint64_t)h->rows * h->ld * 4 >= (256ull << 20);
}
bool half_scan_local(const mlv_index* h, uint32_t nq, uint32_t k, uint64_t gathered_rows) {
    return half_scan_shape(h, nq, k, gathered_rows) && h->half_skip == 0;
}
uint64_t gathered_rows_of(const FilterPlan& fp) { return fp.gather ? (fp.n_rows_host ? fp.n_rows_host : UINT64_MAX) : 0; }
// feedback from the kernels' pinned mirror: a shadow that certifies less than half of its queries sits out 64, 128, ...
// searches (each uncertified query costs the fp32 pass on top); an overflowed shadow is rebuilt with a fresh scale
void half_scan_policy(mlv_index* h) {
    if (h->half_skip > 0) {
        h->half_skip--;
        return;
    }
    if (!h->h_half_stats.p) return;
    const volatile uint32_t* m = (const volatile uint32_t*)h->h_half_stats.p;
    const uint32_t q = m[0], u = m[1];
    if (m[2]) {
        ((volatile uint32_t*)h->h_half_stats.p)[2] = 0;
        h->f16_valid = 0;
    }
    if (q - h->half_seen_q >= 16) {
        if ((u - h->half_seen_u) * 2 > q - h->half_seen_q) {
            h->half_backoff = std::min<uint32_t>(std::max<uint32_t>(64, h->half_backoff * 2), 1u << 16);
            h->half_skip = h->half_backoff;
        } else {
            h->half_backoff = 0;
        }
        h->half_seen_q = q;
        h->half_seen_u = u;
    }
}

// Shadow RANGE scan: the decision (this rank's own -- nothing about it shows in an exchange) and the launch shape.
// On success *use = true, *c holds the half configuration and p describes the shadow (radius etc. are the caller's).
int half_range_setup(mlv_index* h, const FilterPlan& fp, Lane* ln, cudaStream_t st, bool* use, ScanCfg* c, ScanParams* p) {
    *use = false;
    if (h->tune_scan_half == 0 || fp.gather || h->ld < 8 || !h->tune_dynamic) return MLV_OK;
    if (h->tune_scan_half != 1 && (uint64_t)h->rows * h->ld * 4 < (256ull << 20)) return MLV_OK;
    // every hit is re-scored by a whole warp (one dependent HBM round trip each): radii that return more than a few per
    // cent of the rows are cheaper on the fp32 pass.  The hint is the last host call's largest list.
    if (h->tune_scan_half != 1 && h->range_hits_hint * 32 > h->rows) return MLV_OK;
    int rc;
    if (h->h_half_stats.p && ((volatile uint32_t*)h->h_half_stats.p)[2]) {   // a kernel met an overflowed shadow: rebuild it
        ((volatile uint32_t*)h->h_half_stats.p)[2] = 0;
        h->f16_valid = 0;
    }
    if (h->metric != MLV_COSINE && (rc = ensure_row_norms(h, st)) != MLV_OK) return rc;
    bool usable = false;
    if ((rc = ensure_f16_shadow(h, st, &usable)) != MLV_OK) return rc;
    if (!usable) return MLV_OK;
    if (!h->h_half_stats.p) {
        if ((rc = ensure_dev(h, h->d_half_stats, 8)) != MLV_OK) return rc;
        CK(h, cudaMemsetAsync(h->d_half_stats.p, 0, 8, st));
        if ((rc = ensure_host(h, h->h_half_stats, 16)) != MLV_OK) return rc;
        memset(h->h_half_stats.p, 0, 16);
    }
    const int half_kind = (f16_ld(h) % 64 == 0 && h->tune_scan_half_mma != 0) ? 2 : 1;
    if ((rc = choose_cfg(h, 1, 1, true, c, false, half_kind)) != MLV_OK) return rc;
    *p = scan_params(h, *c, fp, ln, 1);
    p->rows = reinterpret_cast<const float4*>(h->d_rows16.p);
    p->ld4 = f16_ld(h) / 8;
    p->rows_exact = reinterpret_cast<const float4*>(h->d_rows);
    p->ld4_exact = h->ld / 4;
    p->half_state = (const uint32_t*)h->d_f16st.p;
    p->row_norms = h->metric == MLV_L2 ? (const float*)h->d_norms.p : nullptr;
    p->max_norm2_bits = h->metric == MLV_COSINE ? nullptr : (const uint32_t*)h->d_maxn2.p;
    p->delta_rel = gemm_delta_rel_f16_host(h->metric == MLV_L2, h->ld);
    p->cosine = h->metric == MLV_COSINE;
    p->half_stats_host = (volatile uint32_t*)h->h_half_stats.p;
    *use = true;
    return MLV_OK;
}
cudaError_t launch_scan_half_range(mlv_index* h, const ScanParams& p, const ScanCfg& c, cudaStream_t st) {
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
    const cudaError_t e = h->metric == MLV_L2 ? launch_scan_half_m<METRIC_L2, true>(p, c, st) : launch_scan_half_m<METRIC_IP, true>(p, c, st);
    h->launches++;
    h->half_scan_launches++;
    if (h->timing) {
        cudaEventRecord(e1, st);
        h->pending.emplace_back(e0, e1);
    }
    return e;
}

int search_prepared(mlv_index* h, const float* qprep, uint32_t nq, uint32_t k, const uint32_t* filter_dev, float* out_d,
                    int64_t* out_r, int32_t* out_c, cudaStream_t st, bool exchange = false, const FastArgs* fast = nullptr) {
    Lane* ln = lane_for(h, st);
    FilterPlan fp;
    int rc = plan_filter(h, ln, filter_dev, st, &fp);
    if (rc != MLV_OK) return rc;
    // one policy step per search: the latency path asks first (with `fast`) and may come back staged
    if (half_scan_shape(h, nq, k, gathered_rows_of(fp))) {
        if (fast) {
            half_scan_policy(h);
            h->half_stepped = true;
        } else if (h->half_stepped) {
            h->half_stepped = false;
        } else {
            half_scan_policy(h);
        }
    }
    // two launches per search (first tier + conditional fp32)?  Across ranks the protocol decides, alone this rank does
    const bool pair_proto = half_scan_proto(h, nq, k) && (exchange || half_scan_local(h, nq, k, gathered_rows_of(fp)));
    ScanCfg c;
    if ((rc = choose_cfg(h, nq, k, false, &c, fp.gather != nullptr)) != MLV_OK) return rc;
    if ((rc = ensure_sched(h, ln)) != MLV_OK) return rc;
    // on one GPU the single-CTA sort of more than ~4096 keys costs as much as the select launch it saves;
    // across GPUs it still replaces two all-gathers and a merge launch
    const bool fused = fused_ok(h, c, k) && (exchange || (uint64_t)c.grid * k <= 4096);
    if (exchange && !fused) return fail(h, MLV_E_UNSUPPORTED, "exchange search needs the fused final select");
    if (fused) c.smem = std::max(c.smem, fused_scratch_bytes(c, k));
    if (fast) {
        const bool ok = fused && nq == 1 && c.NQ == 1 && h->dim <= SCAN_INLINE_MAX_DIM && !h->timing && !pair_proto;
        if (fast->took_fast) *fast->took_fast = ok;
        if (!ok) return MLV_OK;   // nothing launched: the caller stages the query and takes the general path
    }
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

    ScanParams p = scan_params(h, c, fp, ln, k);
    p.fused = fused ? 1 : 0;
    p.fused_cap = fused_cap(c, k);
    p.row_base = h->row_base;
    if (exchange) fill_exchange(h, p.xchg);
    if (h->tune_timeline) {
        rc = ensure_dev(h, h->d_timeline, (size_t)c.grid * 16 * 8);
        if (rc == MLV_OK) CK(h, cudaMemsetAsync(h->d_timeline.p, 0, (size_t)c.grid * 16 * 8, st));
        if (rc != MLV_OK) return rc;
        p.timeline = (unsigned long long*)h->d_timeline.p;
        h->last_grid = c.grid;
    }

    if (pair_proto && fused) {
        // ---- first tier + conditional fp32 launch (one query) ----
        if ((rc = ensure_dev(h, ln->d_cert, 4)) != MLV_OK) return rc;
        bool use_half = half_scan_local(h, nq, k, gathered_rows_of(fp));
        ScanCfg ch{};
        if (use_half) {
            if (h->metric != MLV_COSINE && (rc = ensure_row_norms(h, st)) != MLV_OK) return rc;
            if ((rc = ensure_f16_shadow(h, st, &use_half)) != MLV_OK) return rc;
        }
        if (use_half) {
            // rows of whole 128-byte chunks: tensor-core consumers (the FMA consumers are FMA-latency bound at the
            // power-capped clock: 6.5 of the fp32 pass's 7.4 TB/s)
            const int half_kind = (f16_ld(h) % 64 == 0 && h->tune_scan_half_mma != 0) ? 2 : 1;
            if ((rc = choose_cfg(h, 1, HALF_SCAN_KPRIME, false, &ch, fp.gather != nullptr, half_kind)) != MLV_OK) return rc;
            // the tail's scratch (candidates | k' approximate | 32 exact | k final keys) overlays the ring and must leave
            // the query behind it alone
            const size_t scratch = ((size_t)fused_cap(ch, HALF_SCAN_KPRIME) + HALF_SCAN_KPRIME + 32 + k) * 8;
            use_half = scratch <= (size_t)ch.S * ch.stage_f4 * 16 && (uint64_t)ch.grid * HALF_SCAN_KPRIME <= SCAN_FUSED_MAX_KEYS;
        }
        if (use_half) {
            if ((rc = ensure_dev(h, h->d_half_stats, 8)) != MLV_OK) return rc;
            if (!h->h_half_stats.p) {
                CK(h, cudaMemsetAsync(h->d_half_stats.p, 0, 8, st));
                if ((rc = ensure_host(h, h->h_half_stats, 16)) != MLV_OK) return rc;
                memset(h->h_half_stats.p, 0, 16);
            }
            if ((rc = ensure_dev(h, ln->d_keys0, (size_t)ch.grid * HALF_SCAN_KPRIME * 8)) != MLV_OK) return rc;
        }
        cudaEvent_t e0 = nullptr, e1 = nullptr;
        const bool timing = h->timing;
        if (timing) {   // one event pair around both launches: a search, not a launch, is what scan_time_ms averages
            for (cudaEvent_t* ev : {&e0, &e1}) {
                if (!h->event_pool.empty()) {
                    *ev = h->event_pool.back();
                    h->event_pool.pop_back();
                } else {
                    CK(h, cudaEventCreate(ev));
                }
            }
            cudaEventRecord(e0, st);
            h->timing = false;
        }
        p.queries = reinterpret_cast<const float4*>(qprep);
        p.nq_valid = 1;
        p.out_dists = out_d;
        p.out_rows = out_r;
        p.out_counts = out_c;
        cudaError_t le;
        if (use_half) {
            ScanParams ph = scan_params(h, ch, fp, ln, HALF_SCAN_KPRIME);
            ph.rows = reinterpret_cast<const float4*>(h->d_rows16.p);
            ph.ld4 = f16_ld(h) / 8;
            ph.rows_exact = reinterpret_cast<const float4*>(h->d_rows);
            ph.ld4_exact = h->ld / 4;
            ph.k_out = k;
            ph.half_state = (const uint32_t*)h->d_f16st.p;
            ph.row_norms = h->metric == MLV_L2 ? (const float*)h->d_norms.p : nullptr;
            ph.max_norm2_bits = h->metric == MLV_COSINE ? nullptr : (const uint32_t*)h->d_maxn2.p;
            ph.delta_rel = gemm_delta_rel_f16_host(h->metric == MLV_L2, h->ld);
            ph.cosine = h->metric == MLV_COSINE;
            ph.cert = (uint32_t*)ln->d_cert.p;
            ph.half_stats = (uint32_t*)h->d_half_stats.p;
            ph.half_stats_host = (volatile uint32_t*)h->h_half_stats.p;
            ph.fused = 1;
            ph.fused_cap = fused_cap(ch, HALF_SCAN_KPRIME);
            ph.row_base = h->row_base;
            ph.queries = p.queries;
            ph.nq_valid = 1;
            ph.out_keys = (uint64_t*)ln->d_keys0.p;
            ph.out_dists = out_d;
            ph.out_rows = out_r;
            ph.out_counts = out_c;
            ph.timeline = p.timeline;
            if (exchange) {
                fill_exchange(h, ph.xchg);
                ph.xchg.seq = ++h->xseq;
            }
            le = h->metric == MLV_L2 ? launch_scan_half_m<METRIC_L2>(ph, ch, st) : launch_scan_half_m<METRIC_IP>(ph, ch, st);
            h->launches++;
            h->half_scan_launches++;
        } else {
            // this rank has no shadow to offer after all (no room for it, or -- in an exchange search -- a shard too small to
            // bother): its first launch is the fp32 scan, which always certifies
            p.out_keys = (uint64_t*)ln->d_keys0.p;
            p.cert = (uint32_t*)ln->d_cert.p;
            if (exchange) p.xchg.seq = ++h->xseq;
            le = launch_scan(h, p, c, false, st);
            p.cert = nullptr;
        }
        if (le != cudaSuccess) {
            h->timing = timing;
            if (timing) {
                h->event_pool.push_back(e0);
                h->event_pool.push_back(e1);
            }
            return fail_cuda(h, le, "scan launch");
        }
        p.out_keys = (uint64_t*)ln->d_keys0.p;
        p.run_if = (const uint32_t*)ln->d_cert.p;
        if (exchange) p.xchg.seq = ++h->xseq;
        le = launch_scan(h, p, c, false, st);
        h->timing = timing;
        if (timing) {
            cudaEventRecord(e1, st);
            h->pending.emplace_back(e0, e1);
        }
        if (le != cudaSuccess) return fail_cuda(h, le, "scan launch");
        return MLV_OK;
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
            if (fast) {
                InlineQuery iq;
                memcpy(iq.v, fast->inline_q, (size_t)h->dim * 4);
                p.queries = nullptr;
                p.dim = h->dim;
                p.normalize = h->metric == MLV_COSINE;
                p.done_flag = fast->tagged_out && !exchange ? nullptr : fast->done_flag;
                p.done_value = fast->done_value;
                p.tagged_out = exchange ? nullptr : fast->tagged_out;
                p.tag = fast->done_value;
                CK(h, launch_scan_iq(h, p, iq, c, false, st));
                continue;
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

// Bitonic network over P keys in device memory (P a power of two >= SELECT_MAX_P): CTA-local sorts of 8192 keys,
// global compare-exchange steps for the strides that span CTAs.
int sort_keys_device(mlv_index* h, uint64_t* a, uint64_t P, cudaStream_t st) {
    CK(h, cudaFuncSetAttribute(bitonic_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SELECT_MAX_P * 8)));
    const unsigned blocks = (unsigned)(P / SELECT_MAX_P);
    bitonic_block_kernel<<<blocks, SELECT_THREADS, (size_t)SELECT_MAX_P * 8, st>>>(a, 2, SELECT_MAX_P);
    for (uint64_t size = 2ull * SELECT_MAX_P; size <= P; size <<= 1) {
        for (uint64_t stride = size >> 1; stride >= SELECT_MAX_P; stride >>= 1)
            bitonic_global_kernel<<<(unsigned)std::min<uint64_t>((P / 2 + 255) / 256, 4096), 256, 0, st>>>(a, (uint32_t)P, (uint32_t)size,
                                                                                                         (uint32_t)stride);
        bitonic_block_kernel<<<blocks, SELECT_THREADS, (size_t)SELECT_MAX_P * 8, st>>>(a, (uint32_t)size, (uint32_t)size);
    }
    h->launches += 2;
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
    if ((rc = sort_keys_device(h, a, P, st)) != MLV_OK) return rc;
    decode_keys_kernel<<<(unsigned)std::min<uint64_t>((n + 255) / 256, 1024), 256, 0, st>>>(a, n, h->row_base, out_d, out_r);
    h->launches++;
    CK(h, cudaGetLastError());
    return MLV_OK;
}


}  // namespace
