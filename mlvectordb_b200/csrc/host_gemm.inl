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
    if (bn == 64) return launch_gemm_tt<METRIC, 64>(mx, mqh, mql, gp, grid, st);
    if (bn == 128) return launch_gemm_tt<METRIC, 128>(mx, mqh, mql, gp, grid, st);
    return launch_gemm_tt<METRIC, 256>(mx, mqh, mql, gp, grid, st);
}

// Large batches: tcgen05 GEMM selects k' candidates per query in geometrically growing rounds,
// rerank_kernel scores them in the reference's arithmetic and certifies; uncertified queries are
// re-run by the exact scan.  Synchronises `st` once (to read the per-query flags).
int search_gemm(mlv_index* h, const float* qprep, uint32_t nq, uint32_t k, const uint32_t* filter_dev, float* out_d,
                int64_t* out_r, int32_t* out_c, cudaStream_t st) {
    int rc;
    const uint32_t ld = h->ld;
    const uint32_t GEMM_BN = gemm_tile_width(h, nq);
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
    // a query's buffer is typically a quarter full: 256 threads sort it, and many CTAs share an SM
    const int refine_threads = (int)std::min<uint32_t>(256, std::max<uint32_t>(P / 2, 32));
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
