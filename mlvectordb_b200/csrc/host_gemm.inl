// host_gemm.inl -- host side of the tensor-core batch path: tensor maps, row norms, rounds, re-rank, certificate, scan fallback.
// Part of the single translation unit mlv_index.cu (included there, in order).
#pragma once

namespace {

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

// halves per row of the fp16 shadow: rows stay 16-byte aligned (TMA) -> a multiple of 8
uint32_t f16_ld(const mlv_index* h) { return (h->ld + 7u) & ~7u; }
float gemm_delta_rel_f16_host(bool l2, uint32_t d) { return gemm_delta_rel_f16(l2, d); }

// fp32 matrix [n_rows, ld] row-major -> boxes of {GEMM_BK floats, box_rows rows}, 128-byte swizzle,
// out-of-range elements read as zero (ragged last row tile, ld not a multiple of 32)
// half = true: an fp16 matrix [n_rows, ld16] with boxes of {64 halves, box_rows rows} (the same 128-byte rows)
int make_tile_map(mlv_index* h, CUtensorMap* map, const void* base, uint64_t n_rows, uint32_t box_rows, bool half = false) {
    encode_tiled_fn enc = get_encode_tiled();
    if (!enc) return fail(h, MLV_E_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    const uint32_t ld = half ? f16_ld(h) : h->ld;
    const cuuint64_t gdim[2] = {ld, n_rows};
    const cuuint64_t gstride[1] = {(cuuint64_t)ld * (half ? 2 : 4)};
    const cuuint32_t box[2] = {(cuuint32_t)(half ? 2 * GEMM_BK : GEMM_BK), box_rows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, half ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, gdim, gstride, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(h, MLV_E_CUDA, "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
    return MLV_OK;
}

// queries per GEMM tile: the narrowest shape that holds the batch (small batches are HBM-bound,
// padding them to 256 columns would make them tensor-bound), 256 for anything larger
uint32_t gemm_tile_width(const mlv_index* h, uint32_t nq) {
    if (h->tune_gemm_bn == 64 || h->tune_gemm_bn == 128 || h->tune_gemm_bn == 256) return (uint32_t)h->tune_gemm_bn;
    return nq <= 64 ? 64 : (nq <= 128 ? 128 : 256);
}

uint32_t gemm_kprime(uint32_t k) {
    const uint32_t slack = std::max<uint32_t>(16, k / 4);
    return (k + slack + 31) & ~31u;
}

bool gemm_1pass_ok(const mlv_index* h, uint32_t k);

// Smallest batch the tensor-core path takes.  With the one-pass tier a pass over a >= 1 GB matrix costs 1.1x a
// one-query scan (10M x 768: 4.57 ms vs 4.17 ms), less than the scan's 5 .. 8-query pass (5.4 - 5.6 ms): measured
// crossover nq = 5 at 1M and 10M rows.  The 3xTF32 tier alone needs more than one 8-query scan pass to pay off.
uint32_t gemm_min_nq(const mlv_index* h, uint32_t k) {
    if (h->tune_gemm_min_nq > 0) return (uint32_t)h->tune_gemm_min_nq;
    const bool big = (uint64_t)h->rows * h->ld * 4 >= (1ull << 30);
    // with the fp16 shadow a pass over the rows costs HALF a one-query fp32 scan (10M x 768: 2.37 ms for up to 64 queries
    // against 4.1 ms): two queries already pay for it -- and two shadow scans cost 4.3 ms (profiles/r02_batch_surface_768.jsonl)
    if (gemm_1pass_ok(h, k) && big && h->tune_gemm_passes != 1 && h->tune_gemm_passes != 3 && h->d_rows16.p) return 2u;
    return gemm_1pass_ok(h, k) && big ? 5u : 9u;
}

bool gemm_eligible(const mlv_index* h, uint32_t nq, uint32_t k) {
    if (h->tune_gemm == 0) return false;
    if (h->ld < (uint32_t)GEMM_BK) return false;
    if (gemm_kprime(k) * 4 > SELECT_MAX_P) return false;
    if (h->tune_gemm == 1) return true;
    return nq >= gemm_min_nq(h, k) && h->rows >= 16384;
}

// Order `st` behind the last norm / shadow maintenance if that ran on another stream; call before using either.
int maint_wait(mlv_index* h, cudaStream_t st) {
    if (!h->maint_gen) return MLV_OK;
    Lane* ln = lane_for(h, st);
    if (ln->seen_maint == h->maint_gen) return MLV_OK;
    if (st != h->maint_stream) CK(h, cudaStreamWaitEvent(st, h->maint_event, 0));
    ln->seen_maint = h->maint_gen;
    return MLV_OK;
}
// ... and publish maintenance work just queued on `st`
int maint_done(mlv_index* h, cudaStream_t st) {
    if (!h->maint_event) CK(h, cudaEventCreateWithFlags(&h->maint_event, cudaEventDisableTiming));
    CK(h, cudaEventRecord(h->maint_event, st));
    h->maint_gen++;
    h->maint_stream = st;
    lane_for(h, st)->seen_maint = h->maint_gen;
    return MLV_OK;
}

int ensure_row_norms(mlv_index* h, cudaStream_t st) {
    int rc;
    if ((rc = maint_wait(h, st)) != MLV_OK) return rc;
    if (!h->d_maxn2.p) {
        if ((rc = ensure_dev(h, h->d_maxn2, 4)) != MLV_OK) return rc;
        CK(h, cudaMemsetAsync(h->d_maxn2.p, 0, 4, st));
        h->norms_valid = 0;
        h->f16_valid = 0;   // the shadow's scale follows the largest norm
    }
    if (h->d_norms.bytes < h->rows * 4) {
        // growing reallocates: recompute everything (rows rarely grow between large batches)
        if ((rc = ensure_dev(h, h->d_norms, std::max<uint64_t>(h->capacity, h->rows) * 4)) != MLV_OK) return rc;
        h->norms_valid = 0;
        h->f16_valid = 0;
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
        if ((rc = maint_done(h, st)) != MLV_OK) return rc;
    }
    return MLV_OK;
}

template <int METRIC, int BN, int PASSES>
cudaError_t launch_gemm_tt(const CUtensorMap& mx, const CUtensorMap& mqh, const CUtensorMap& mql, const GemmParams& gp, int grid,
                           cudaStream_t st) {
    auto kern = gemm_topk_kernel<METRIC, BN, PASSES>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GemmShape<BN, PASSES>::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    kern<<<grid, GEMM_THREADS, GemmShape<BN, PASSES>::SMEM_BYTES, st>>>(mx, mqh, mql, gp);
    return cudaGetLastError();
}
template <int METRIC, int PASSES>
cudaError_t launch_gemm_tp(const CUtensorMap& mx, const CUtensorMap& mqh, const CUtensorMap& mql, const GemmParams& gp, int grid,
                           cudaStream_t st, int bn) {
    if (bn == 64) return launch_gemm_tt<METRIC, 64, PASSES>(mx, mqh, mql, gp, grid, st);
    if (bn == 128) return launch_gemm_tt<METRIC, 128, PASSES>(mx, mqh, mql, gp, grid, st);
    return launch_gemm_tt<METRIC, 256, PASSES>(mx, mqh, mql, gp, grid, st);
}
template <int METRIC>
cudaError_t launch_gemm_t(const CUtensorMap& mx, const CUtensorMap& mqh, const CUtensorMap& mql, const GemmParams& gp, int grid,
                          cudaStream_t st, int bn, int passes = 3) {
    if (passes == GEMM_TIER_F16) return launch_gemm_tp<METRIC, GEMM_TIER_F16>(mx, mqh, mql, gp, grid, st, bn);
    return passes == 1 ? launch_gemm_tp<METRIC, 1>(mx, mqh, mql, gp, grid, st, bn) : launch_gemm_tp<METRIC, 3>(mx, mqh, mql, gp, grid, st, bn);
}

// the CTA-pair kernel (cta_group::2: the pair shares the query tile inside the tensor cores)
template <int METRIC, int PASSES>
cudaError_t launch_gemm_pair_t(const CUtensorMap& mx, const CUtensorMap& mq, const GemmParams& gp, int grid, cudaStream_t st) {
    auto kern = gemm_topk_pair_kernel<METRIC, PASSES>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMMP_SMEM_BYTES);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(GEMM_THREADS);
    cfg.dynamicSmemBytes = GEMMP_SMEM_BYTES;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, mx, mq, gp);
}
template <int METRIC>
cudaError_t launch_gemm_pair(const CUtensorMap& mx, const CUtensorMap& mq, const GemmParams& gp, int grid, cudaStream_t st, int passes) {
    return passes == GEMM_TIER_F16 ? launch_gemm_pair_t<METRIC, GEMM_TIER_F16>(mx, mq, gp, grid, st) : launch_gemm_pair_t<METRIC, 1>(mx, mq, gp, grid, st);
}

// The fp16 shadow of the rows ([capacity, f16_ld] halves of value * 2^s) behind the HALF tier, built lazily like the row
// norms: rows [0, f16_valid) are current; the scale is frozen when row 0 is converted (st[0..2] = 2^-s, s, overflow flag).
// Returns MLV_OK with *usable = false when there is no room for it (the caller takes the TF32 tier).
int ensure_f16_shadow(mlv_index* h, cudaStream_t st, bool* usable) {
    *usable = false;
    {
        const int rcw = maint_wait(h, st);
        if (rcw != MLV_OK) return rcw;
    }
    const uint32_t ld16 = f16_ld(h);
    if (h->d_rows16.bytes < (size_t)h->rows * ld16 * 2) {
        free_dev(h->d_rows16);
        if (ensure_dev(h, h->d_rows16, (size_t)std::max<uint64_t>(h->capacity, h->rows) * ld16 * 2) != MLV_OK) {
            h->err.clear();
            cudaGetLastError();
            return MLV_OK;
        }
        h->f16_valid = 0;
    }
    int rc;
    if ((rc = ensure_dev(h, h->d_f16st, 16)) != MLV_OK) return rc;
    if (h->f16_valid == 0) {
        f16_freeze_scale_kernel<<<1, 1, 0, st>>>(h->metric == MLV_COSINE ? nullptr : (const uint32_t*)h->d_maxn2.p, (uint32_t*)h->d_f16st.p);
        h->launches++;
    }
    if (h->f16_valid < h->rows) {
        const uint64_t n = h->rows - h->f16_valid;
        const uint64_t total = n * (ld16 / 2);
        convert_rows_f16_kernel<<<(unsigned)std::min<uint64_t>((total + 255) / 256, (uint64_t)h->sm_count * 16), 256, 0, st>>>(
            h->d_rows, h->f16_valid, n, h->ld, ld16, (__half*)h->d_rows16.p, (uint32_t*)h->d_f16st.p);
        h->launches++;
        CK(h, cudaGetLastError());
        h->f16_valid = h->rows;
        if ((rc = maint_done(h, st)) != MLV_OK) return rc;
    }
    *usable = true;
    return MLV_OK;
}

// candidates kept per query by the fast tier: its approximate distances are 16x coarser, so the exact k-th best must
// clear a wider margin below the k'-th approximate distance -- twice k (at least k + 64)
uint32_t gemm_kprime_1pass(uint32_t k) { return (std::max<uint32_t>(2 * k, k + 64) + 31) & ~31u; }

bool gemm_1pass_ok(const mlv_index* h, uint32_t k) {
    if (h->tune_gemm_passes == 3) return false;
    return gemm_kprime_1pass(k) * 4 <= SELECT_MAX_P;
}

// What a tier multiplies the queries with: the index's own matrix (rows masked by the tombstone / filter bitmaps in the
// epilogue), or -- for a selective filter -- a compacted copy of the passing-and-live rows, whose positions `rowmap`
// translates back (the list is ascending, so ordering candidates by position orders them by row).
struct GemmView {
    const float* rows = nullptr;
    uint64_t n_rows = 0;
    const float* norms = nullptr;      // |x|^2 per view row (l2 only)
    const uint32_t* live = nullptr;
    const uint32_t* filter = nullptr;
    const uint32_t* rowmap = nullptr;  // view row -> index row (nullptr = identity)
    const void* rows16 = nullptr;      // fp16 shadow of `rows` (HALF tier), or nullptr
    const float* x_unscale = nullptr;  // its frozen 2^-s (device scalar)
};

// One tier of the batch path: tcgen05 GEMM (PASSES = 1 or 3) selects k' candidates per query in geometrically growing
// rounds, rerank_kernel scores them in the reference's arithmetic and certifies.  hflags[q] != 0 = not certified
// (or the candidate buffer overflowed): the caller re-runs those queries.  Synchronises `st` once (to read the flags).
int search_gemm_tier(mlv_index* h, const float* qprep, uint32_t nq, uint32_t k, int passes, const GemmView& view, float* out_d,
                     int64_t* out_r, int32_t* out_c, cudaStream_t st, std::vector<uint32_t>& hflags, bool predict = false) {
    NvtxRange nvtx_range(passes == GEMM_TIER_F16 ? "gemm tier: fp16 shadow" : (passes == 1 ? "gemm tier: 1xTF32" : "gemm tier: 3xTF32"));
    int rc;
    const uint32_t ld = h->ld;
    const uint32_t GEMM_BN = gemm_tile_width(h, nq);
    const uint32_t nq_pad = (nq + GEMM_BN - 1) / GEMM_BN * GEMM_BN;
    const bool half = passes == GEMM_TIER_F16;
    const uint32_t ld16 = f16_ld(h);
    const uint32_t kprime = (passes == 1 || half) ? gemm_kprime_1pass(k) : gemm_kprime(k);
    const uint32_t cap = std::min<uint32_t>(SELECT_MAX_P, pow2_ceil(8 * kprime));
    const uint32_t P = cap;  // power of two
    const bool l2 = h->metric == MLV_L2;
    // scratch: Qhi | Qlo | qn | thr | cnt | flags
    const size_t qmat = (size_t)nq_pad * ld * 4;
    if ((rc = ensure_dev(h, h->d_gq, 2 * qmat + (size_t)nq_pad * 20)) != MLV_OK) return rc;
    if ((rc = ensure_dev(h, h->d_cand, (size_t)nq_pad * cap * 8)) != MLV_OK) return rc;
    float* qhi = (float*)h->d_gq.p;
    float* qlo = qhi + (size_t)nq_pad * ld;
    float* qn = qlo + (size_t)nq_pad * ld;
    float* thr = qn + nq_pad;
    uint32_t* cnt = (uint32_t*)(thr + nq_pad);
    uint32_t* flags = cnt + nq_pad;
    uint32_t* sorted_n = flags + nq_pad;
    uint64_t* cand = (uint64_t*)h->d_cand.p;
    // HALF tier: the Qhi area holds the fp16 queries, the Qlo area their 2^-sq
    float* q_unscale = qlo;
    {
        const int wpb = 8;
        if (half)
            split_queries_f16_kernel<<<(nq_pad + wpb - 1) / wpb, wpb * 32, 0, st>>>(qprep, (__half*)qhi, q_unscale, qn, thr, cnt, flags,
                                                                                    sorted_n, nq, nq_pad, ld, ld16);
        else
            split_queries_kernel<<<(nq_pad + wpb - 1) / wpb, wpb * 32, 0, st>>>(qprep, qhi, qlo, qn, thr, cnt, flags, sorted_n, nq, nq_pad, ld);
        h->launches++;
        CK(h, cudaGetLastError());
    }
    // one-pass tiers of batches wider than 128 queries: CTA pairs (tcgen05 cta_group::2, double-buffered accumulators);
    // tune_gemm_wide 0 = the single-tile kernel
    const bool pair = GEMM_BN == 256 && passes != 3 && h->tune_gemm_wide != 0;
    CUtensorMap mx, mqh, mql;
    if ((rc = make_tile_map(h, &mx, half ? view.rows16 : (const void*)view.rows, view.n_rows, GEMM_BM, half)) != MLV_OK) return rc;
    if ((rc = make_tile_map(h, &mqh, qhi, nq_pad, pair ? GEMM_BN / 2 : GEMM_BN, half)) != MLV_OK) return rc;
    if ((rc = make_tile_map(h, &mql, half ? qhi : qlo, nq_pad, GEMM_BN, half)) != MLV_OK) return rc;
    CK(h, cudaFuncSetAttribute(refine_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((SELECT_MAX_P + SELECT_MAX_P / 4) * 8)));

    GemmParams gp{};
    gp.n_rows = (uint32_t)view.n_rows;
    gp.nq = nq;
    gp.n_qtiles = nq_pad / GEMM_BN;
    gp.n_kchunks = half ? (ld16 + 2 * GEMM_BK - 1) / (2 * GEMM_BK) : (ld + GEMM_BK - 1) / GEMM_BK;
    gp.x_unscale = view.x_unscale;
    gp.q_unscale = q_unscale;
    gp.debug = h->tune_gemm_debug;
    gp.row_norms = l2 ? view.norms : nullptr;
    gp.q_norms = qn;
    gp.thr = thr;
    gp.live = view.live;
    gp.filter = view.filter;
    gp.cand = cand;
    gp.cand_cnt = cnt;
    gp.cap = cap;

    // Rounds.  Plain rule: the first takes as many rows as a candidate buffer holds (no threshold yet), each later one
    // (cap - k') / (4 k') times the rows seen so far, so a buffer is expected to stay at most a quarter full however the
    // thresholds started.  Predicted thresholds (refine_kernel): the first round is two tiles, a round that has seen S of
    // the N rows thresholds at its j-th best, j = max(32, 4 k' S / N) capped at k' -- about 4 k' rows of the whole
    // matrix are expected below it, and it is the k'-th best again once a quarter of the rows are seen -- and the next
    // round takes min(8, (cap - k') / (4 j)) times the rows seen: a 10M-row pass appends ~2,000 candidates per query
    // in 7 rounds instead of ~5,600 in 9.
    const uint32_t total_tiles = (uint32_t)((view.n_rows + GEMM_BM - 1) / GEMM_BM);
    predict = predict && passes != 3 && kprime > 32;
    auto rank_after = [&](uint32_t seen_tiles) -> uint32_t {
        if (!predict || seen_tiles >= total_tiles) return kprime;
        const double j = 4.0 * kprime * (double)seen_tiles / (double)total_tiles;
        return (uint32_t)std::min<double>(kprime, std::max(32.0, std::ceil(j)));
    };
    uint32_t jrank = kprime;   // rank of the thresholds in force
    uint32_t seen = 0;
    // a query's buffer is typically a quarter full: 256 threads sort it, and many CTAs share an SM
    const int refine_threads = (int)std::min<uint32_t>(256, std::max<uint32_t>(P / 2, 32));
    while (seen < total_tiles) {
        const double growth = std::min(predict ? 8.0 : 1e9, (double)(cap - kprime) / (4.0 * jrank));
        uint32_t take = seen == 0 ? (predict ? 2u : std::max<uint32_t>(1, cap / GEMM_BM)) : std::max<uint32_t>(1, (uint32_t)(seen * growth));
        take = std::min(take, total_tiles - seen);
        jrank = rank_after(seen + take);
        gp.row_tile0 = seen;
        gp.row_tile1 = seen + take;
        uint64_t items = (uint64_t)take * gp.n_qtiles;
        int grid = (int)std::min<uint64_t>(items, (uint64_t)h->sm_count);
        if (pair) {
            items = (uint64_t)((take + 1) / 2) * gp.n_qtiles * 2;
            grid = (int)std::min<uint64_t>(items, (uint64_t)h->sm_count);
        }
        if (pair) grid = std::max(2, grid & ~1);
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
        if (pair) {
            cudaError_t le = l2 ? launch_gemm_pair<METRIC_L2>(mx, mqh, gp, grid, st, passes) : launch_gemm_pair<METRIC_IP>(mx, mqh, gp, grid, st, passes);
            if (le != cudaSuccess) {
                // a device / partition that cannot co-schedule CTA pairs: nothing was launched -- this handle uses the
                // single-tile kernel from now on, and this batch starts over with it (no round has completed yet or the
                // thresholds simply carry over: candidate buffers are additive)
                cudaGetLastError();
                h->tune_gemm_wide = 0;
                return search_gemm_tier(h, qprep, nq, k, passes, view, out_d, out_r, out_c, st, hflags, predict);
            }
        } else
            CK(h, l2 ? launch_gemm_t<METRIC_L2>(mx, mqh, mql, gp, grid, st, (int)GEMM_BN, passes)
                     : launch_gemm_t<METRIC_IP>(mx, mqh, mql, gp, grid, st, (int)GEMM_BN, passes));
        if (h->timing) {
            cudaEventRecord(e1, st);
            h->gemm_pending.emplace_back(e0, e1);
        }
        refine_kernel<<<nq, refine_threads, (size_t)(P + kprime) * 8, st>>>(cand, cnt, thr, flags, sorted_n, cap, P, kprime, jrank,
                                                                            seen + take >= total_tiles ? 1 : 0);
        CK(h, cudaGetLastError());
        h->launches += 2;
        h->gemm_launches++;
        h->gemm_rounds++;
        seen += take;
    }

    RerankParams rp{};
    rp.rows = reinterpret_cast<const float4*>(view.rows);
    rp.rowmap = view.rowmap;
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
    rp.delta_rel = half ? gemm_delta_rel_f16(l2, ld) : (passes == 1 ? gemm_delta_rel_1pass(l2, ld) : GEMM_DELTA_REL);
    if (l2)
        rerank_kernel<METRIC_L2><<<nq, 256, (size_t)rp.P * 8, st>>>(rp);
    else
        rerank_kernel<METRIC_IP><<<nq, 256, (size_t)rp.P * 8, st>>>(rp);
    h->launches++;
    CK(h, cudaGetLastError());

    // certificate check: the one synchronisation of this tier
    hflags.resize(nq);
    uint32_t f16st[3] = {0, 0, 0};
    CK(h, cudaMemcpyAsync(hflags.data(), flags, (size_t)nq * 4, cudaMemcpyDeviceToHost, st));
    if (half) CK(h, cudaMemcpyAsync(f16st, h->d_f16st.p, sizeof(f16st), cudaMemcpyDeviceToHost, st));
    CK(h, cudaStreamSynchronize(st));
    if (half && f16st[2]) {
        // rows far larger than any present when the shadow's scale was frozen overflowed fp16: nothing this tier
        // selected can be trusted; the shadow is rebuilt with a fresh scale by the next batch
        std::fill(hflags.begin(), hflags.end(), 2u);
        h->f16_valid = 0;
        h->f16_overflowed = true;
    }
    return MLV_OK;
}

// Large batches.  Tier 1: one-pass TF32 GEMM (a third of the tensor work) with a wide candidate slack; tier 2: the
// 3xTF32 GEMM for the queries tier 1 could not certify; what neither certifies is re-run by the exact scan.  Every
// tier ends in the same exact re-rank, so results never depend on which tier answered.
int search_gemm(mlv_index* h, const float* qprep, uint32_t nq, uint32_t k, const uint32_t* filter_dev, float* out_d,
                int64_t* out_r, int32_t* out_c, cudaStream_t st) {
    int rc;
    const uint32_t ld = h->ld;
    std::vector<uint32_t> hflags;
    std::vector<uint32_t> failing;
    h->gemm_searches++;
    h->gemm_queries += nq;
    if (h->metric != MLV_COSINE && (rc = ensure_row_norms(h, st)) != MLV_OK) return rc;
    GemmView view;
    view.rows = h->d_rows;
    view.n_rows = h->rows;
    view.norms = (const float*)h->d_norms.p;
    view.live = h->n_deleted ? h->d_live : nullptr;
    view.filter = filter_dev;
    bool short_bitmap = false;
    uint32_t gathered_rows = 0;
    mlv_filter* bf = filter_dev ? nullptr : h->bound_filter;
    if (bf) {
        if (bf->compact_gen != h->compact_gen)
            return fail(h, MLV_E_INVALID, "prepared filter predates a compaction / clear of the index (rows were renumbered); create it again");
        view.filter = (const uint32_t*)bf->d_bitmap.p;
        // rows were appended after the filter was made (they do not pass): the bitmap cannot mask a full pass, but the
        // row list still describes it -- such a batch always multiplies the compacted rows (or takes the gathered scan)
        short_bitmap = bf->bitmap_words < (h->rows + 31) / 32;
    }
    if (view.filter && (h->tune_gather != 0 || short_bitmap)) {
        // Selective filter: multiply only the passing-and-live rows.  Their list is the scan's gather list (prepared
        // filters keep it; a per-call bitmap builds it here); the rows are copied once into a dense matrix, which
        // costs a read + write of s x the matrix against a GEMM over all of it.
        Lane* ln = lane_for(h, st);
        FilterPlan fp;
        const int keep_tune = h->tune_gather;
        if ((rc = plan_filter(h, ln, filter_dev, st, &fp, /*want_list=*/true)) != MLV_OK) return rc;
        uint32_t m = 0;
        CK(h, cudaMemcpyAsync(&m, fp.n_rows_dev, 4, cudaMemcpyDeviceToHost, st));
        CK(h, cudaStreamSynchronize(st));
        if (m == 0) {
            fill_empty_kernel<<<32, 256, 0, st>>>(out_d, out_r, out_c, nq, k);
            h->launches++;
            CK(h, cudaGetLastError());
            return MLV_OK;
        }
        const uint64_t live_rows = h->rows - h->n_deleted;
        // The copy reads and writes the passing rows (2 s B bytes for a matrix of B bytes at selectivity s) and saves
        // (1 - s) of a GEMM over everything, which takes c passes' worth of time: c ~ 1.1 while a pass is HBM-bound
        // (<= 64 queries), ~ nq / 120 for wide batches (measured: 256 queries 2.3, 1024 queries 7.7).  Gather when
        // 2.5 s < (1 - s) c, with a 20 % margin for the allocation: 1024 queries up to ~60 % passing, 256 up to ~37 %,
        // small batches up to ~24 % (10M x 384: 256 queries at 50 % lost 10.6 vs 5.1 ms to the copy, 1024 won 12.3 vs 18.5).
        const double c = std::max(1.1, (double)nq / 120.0);
        const double s_max = 0.8 * c / (2.5 + c);
        bool gather = (double)m <= s_max * (double)live_rows || keep_tune == 1 || short_bitmap;
        if (gather && m < 16384 && keep_tune != 1)   // a few thousand rows: eight queries per gathered scan pass are cheaper than GEMM rounds
            return search_prepared(h, qprep, nq, k, filter_dev, out_d, out_r, out_c, st);
        if (gather && ensure_dev(h, h->d_gx, (size_t)m * ld * 4 + (size_t)m * 4) != MLV_OK) {
            h->err.clear();
            if (short_bitmap) return search_prepared(h, qprep, nq, k, filter_dev, out_d, out_r, out_c, st);   // the scan reads through the list
            gather = false;   // no room for the copy beside the matrix: mask in the epilogue instead
        }
        if (gather) {
            gathered_rows = m;
            float* gx = (float*)h->d_gx.p;
            float* gn = gx + (size_t)m * ld;
            const int wpb = 8;
            gather_rows_kernel<<<(unsigned)((m + wpb - 1) / wpb), wpb * 32, 0, st>>>(
                reinterpret_cast<const float4*>(h->d_rows), fp.gather, m, ld / 4, reinterpret_cast<float4*>(gx),
                h->metric == MLV_L2 ? (const float*)h->d_norms.p : nullptr, gn);
            h->launches++;
            CK(h, cudaGetLastError());
            view.rows = gx;
            view.n_rows = m;
            view.norms = gn;
            view.live = nullptr;
            view.filter = nullptr;
            view.rowmap = fp.gather;
            h->gemm_gathered_searches++;
        }
    }
    // data on which the one-pass tier certifies little (neighbours crowded within its error) would pay for both
    // tiers on every batch: after a batch where it failed for more than half of the queries it sits out 8 batches
    bool fast = gemm_1pass_ok(h, k);
    if (fast && h->tune_gemm_passes == 0 && h->gemm_fast_skip > 0) {
        h->gemm_fast_skip--;
        fast = false;
    }
    // first tier: the fp16 shadow (twice the rate of the TF32 pass and half its error) when there is room for it
    int first_tier = fast ? 1 : 3;
    if (fast && (h->tune_gemm_passes == 0 || h->tune_gemm_passes == GEMM_TIER_F16)) {
        bool usable = false;
        if ((rc = ensure_f16_shadow(h, st, &usable)) != MLV_OK) return rc;
        if (usable && gathered_rows) {   // a filtered batch multiplies the compacted rows: their halves, same frozen scale
            const uint32_t ld16 = f16_ld(h);
            if (ensure_dev(h, h->d_gx16, (size_t)gathered_rows * ld16 * 2) != MLV_OK) {
                usable = false;
                h->err.clear();
                cudaGetLastError();
            } else {
                const uint64_t total = (uint64_t)gathered_rows * (ld16 / 2);
                convert_rows_f16_kernel<<<(unsigned)std::min<uint64_t>((total + 255) / 256, (uint64_t)h->sm_count * 16), 256, 0, st>>>(
                    view.rows, 0, gathered_rows, ld, ld16, (__half*)h->d_gx16.p, (uint32_t*)h->d_f16st.p);
                h->launches++;
                CK(h, cudaGetLastError());
                view.rows16 = h->d_gx16.p;
            }
        } else if (usable) {
            view.rows16 = h->d_rows16.p;
        }
        if (usable) {
            view.x_unscale = (const float*)h->d_f16st.p;
            first_tier = GEMM_TIER_F16;
        }
    }
    // predicted thresholds for the one-pass tier (not while they sit out after failing; a forced tier without a tier
    // behind it keeps the plain rule, or a misprediction would send the query to the scan)
    bool predict = fast && h->tune_gemm_predict != 0 && h->tune_gemm_passes == 0;
    if (predict && h->gemm_predict_skip > 0) {
        h->gemm_predict_skip--;
        predict = false;
    }
    if ((rc = search_gemm_tier(h, qprep, nq, k, first_tier, view, out_d, out_r, out_c, st, hflags, predict)) != MLV_OK) return rc;
    size_t mispredicted = 0;
    for (uint32_t q = 0; q < nq; q++)
        if (hflags[q]) {
            failing.push_back(q);
            mispredicted += (hflags[q] & 4u) ? 1 : 0;
        }
    if (predict) {
        h->gemm_mispredicted_queries += mispredicted;
        if (mispredicted * 8 > nq) {   // the rows' order correlates with the queries: the plain rule for 8, 16, ... batches
            h->gemm_predict_backoff = std::min<uint32_t>(std::max<uint32_t>(8, h->gemm_predict_backoff * 2), 1u << 14);
            h->gemm_predict_skip = h->gemm_predict_backoff;
        } else {
            h->gemm_predict_backoff = 0;
        }
    }
    if (fast) {
        h->gemm_fast_queries += nq - failing.size();
        if (first_tier == GEMM_TIER_F16) h->gemm_half_queries += nq - failing.size();
        if (h->tune_gemm_passes == 0 && (failing.size() - mispredicted) * 2 > nq && !h->f16_overflowed) h->gemm_fast_skip = 8;
        h->f16_overflowed = false;   // an overflowed shadow says nothing about the data's neighbourhoods: no sitting out
        if (!failing.empty() && h->tune_gemm_passes != 1 && h->tune_gemm_passes != GEMM_TIER_F16) {
            // second tier on the compacted failing queries; results scattered back to their slots
            const uint32_t nf = (uint32_t)failing.size();
            const size_t need = (size_t)nf * 4 + (size_t)nf * ld * 4 + (size_t)nf * k * 12 + (size_t)nf * 4 + 64;
            if ((rc = ensure_dev(h, h->d_sub, need)) != MLV_OK) return rc;
            uint32_t* d_idx = (uint32_t*)h->d_sub.p;
            float* sub_q = (float*)(d_idx + ((nf + 3) & ~3u));
            int64_t* sub_r = (int64_t*)(sub_q + (size_t)nf * ld + ((size_t)nf * ld & 1));
            float* sub_d = (float*)(sub_r + (size_t)nf * k);
            int32_t* sub_c = (int32_t*)(sub_d + (size_t)nf * k);
            CK(h, cudaMemcpyAsync(d_idx, failing.data(), (size_t)nf * 4, cudaMemcpyHostToDevice, st));
            gather_queries_kernel<<<(unsigned)std::min<uint64_t>(((uint64_t)nf * ld + 255) / 256, 2048), 256, 0, st>>>(qprep, d_idx, sub_q, nf, ld);
            h->launches++;
            CK(h, cudaGetLastError());
            std::vector<uint32_t> f2;
            if ((rc = search_gemm_tier(h, sub_q, nf, k, 3, view, sub_d, sub_r, sub_c, st, f2)) != MLV_OK) return rc;
            scatter_results_kernel<<<(unsigned)std::min<uint64_t>(((uint64_t)nf * k + 255) / 256, 2048), 256, 0, st>>>(
                d_idx, nf, k, sub_d, sub_r, sub_c, out_d, out_r, out_c);
            h->launches++;
            CK(h, cudaGetLastError());
            std::vector<uint32_t> still;
            for (uint32_t i = 0; i < nf; i++)
                if (f2[i]) still.push_back(failing[i]);
            failing.swap(still);
        }
    }
    for (uint32_t q : failing) {
        h->gemm_fallback_queries++;
        rc = search_prepared(h, qprep + (size_t)q * ld, 1, k, filter_dev, out_d + (size_t)q * k, out_r + (size_t)q * k, out_c + q, st);
        if (rc != MLV_OK) return rc;
    }
    // the dense copy of a filtered batch's rows can be a large fraction of the matrix: small ones are kept for the next
    // batch, large ones go back to the allocator (every tier has synchronised `st` after its last use of the copy)
    if (h->d_gx.bytes > ((size_t)1 << 30)) free_dev(h->d_gx);
    if (h->d_gx16.bytes > ((size_t)1 << 29)) free_dev(h->d_gx16);
    return MLV_OK;
}


}  // namespace
