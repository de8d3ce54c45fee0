// host_columns.inl -- columnar metadata, device-evaluated predicates, snapshot export / import.
// Part of the single translation unit mlv_index.cu (included there, in order).
#pragma once

namespace {

unsigned grid_for(const mlv_index* h, uint64_t items, unsigned per_block) {
    return (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((items + per_block - 1) / per_block, (uint64_t)h->sm_count * 16));
}

// Make column c cover `capacity` rows (new rows read as missing); keeps what was written.
int ensure_column(mlv_index* h, uint32_t c) {
    if (h->col_rows[c] >= h->capacity && h->d_cols[c].p) return MLV_OK;
    const uint64_t cap = std::max<uint64_t>(h->capacity, 32);
    int32_t* fresh = nullptr;
    CK(h, cudaMalloc(&fresh, cap * 4));
    const uint64_t keep = h->d_cols[c].p ? std::min(h->col_rows[c], cap) : 0;
    cudaError_t e = cudaSuccess;
    if (keep) e = cudaMemcpyAsync(fresh, h->d_cols[c].p, keep * 4, cudaMemcpyDeviceToDevice, h->stream);
    if (e == cudaSuccess) {
        fill_i32_kernel<<<grid_for(h, cap - keep, 256), 256, 0, h->stream>>>(fresh + keep, cap - keep, COLUMN_MISSING);
        h->launches++;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    if (e != cudaSuccess) {
        cudaFree(fresh);
        return fail_cuda(h, e, "column allocation");
    }
    free_dev(h->d_cols[c]);
    h->d_cols[c].p = fresh;
    h->d_cols[c].bytes = cap * 4;
    h->col_rows[c] = cap;
    return MLV_OK;
}

void drop_columns(mlv_index* h) {
    for (uint32_t c = 0; c < MLV_MAX_COLUMNS; c++) {
        free_dev(h->d_cols[c]);
        h->col_rows[c] = 0;
    }
}

// Called by mlv_index_compact after live_prefix_kernel (d_live still the old bitmap): every column is
// rewritten in the survivors' order.  Enqueued on h->stream; the old buffers are released by the caller's
// synchronise + swap (fresh[] holds the new ones).
int compact_columns(mlv_index* h, const uint64_t* d_wbase, uint64_t n, int32_t* fresh[MLV_MAX_COLUMNS]) {
    for (uint32_t c = 0; c < MLV_MAX_COLUMNS; c++) fresh[c] = nullptr;
    for (uint32_t c = 0; c < MLV_MAX_COLUMNS; c++) {
        if (!h->d_cols[c].p) continue;
        const uint64_t cap = std::max<uint64_t>(h->capacity, 32);
        cudaError_t e = cudaMalloc(&fresh[c], cap * 4);
        if (e == cudaSuccess) {
            fill_i32_kernel<<<grid_for(h, cap, 256), 256, 0, h->stream>>>(fresh[c], cap, COLUMN_MISSING);
            const uint64_t rows = std::min(n, h->col_rows[c]);
            compact_column_kernel<<<grid_for(h, rows, 256), 256, 0, h->stream>>>((const int32_t*)h->d_cols[c].p, fresh[c],
                                                                                 h->d_live, d_wbase, rows);
            h->launches += 2;
            e = cudaGetLastError();
        }
        if (e != cudaSuccess) {
            for (uint32_t j = 0; j <= c; j++)
                if (fresh[j]) cudaFree(fresh[j]), fresh[j] = nullptr;
            return fail_cuda(h, e, "column compaction");
        }
    }
    return MLV_OK;
}

void adopt_compacted_columns(mlv_index* h, int32_t* fresh[MLV_MAX_COLUMNS]) {
    for (uint32_t c = 0; c < MLV_MAX_COLUMNS; c++) {
        if (!fresh[c]) continue;
        const uint64_t cap = std::max<uint64_t>(h->capacity, 32);
        free_dev(h->d_cols[c]);
        h->d_cols[c].p = fresh[c];
        h->d_cols[c].bytes = cap * 4;
        h->col_rows[c] = cap;
    }
}

int set_column_common(mlv_index* h, uint32_t column, uint64_t first_row, const int32_t* values, uint64_t n, cudaMemcpyKind kind) {
    if (!h || (!values && n)) return MLV_E_INVALID;
    if (column >= MLV_MAX_COLUMNS) return fail(h, MLV_E_UNSUPPORTED, "column index exceeds MLV_MAX_COLUMNS");
    if (first_row + n > h->rows || first_row + n < first_row) return fail(h, MLV_E_INVALID, "column write beyond the stored rows");
    if (n == 0) return MLV_OK;
    DeviceGuard g(h->device);
    int rc = ensure_column(h, column);
    if (rc != MLV_OK) return rc;
    CK(h, cudaMemcpyAsync((int32_t*)h->d_cols[column].p + first_row, values, n * 4, kind, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    return MLV_OK;
}

// a fresh filter starts from the buffers of a destroyed one when there are any
mlv_filter* new_filter(mlv_index* h, uint64_t n_words) {
    mlv_filter* f = new (std::nothrow) mlv_filter();
    if (!f) return nullptr;
    f->owner = h;
    f->compact_gen = h->compact_gen;
    f->bitmap_words = n_words;
    if (!h->filter_pool.empty()) {
        FilterBufs b = h->filter_pool.back();
        h->filter_pool.pop_back();
        f->d_bitmap = b.bitmap;
        f->d_list = b.list;
        f->d_scratch = b.scratch;
    }
    return f;
}
void retire_filter(mlv_index* h, mlv_filter* f) {
    if (h->filter_pool.size() < MLV_FILTER_POOL) {
        h->filter_pool.push_back(FilterBufs{f->d_bitmap, f->d_list, f->d_scratch});
    } else {
        for (DevBuf* b : {&f->d_bitmap, &f->d_list, &f->d_scratch}) free_dev(*b);
    }
    delete f;
}
void drop_filter_pool(mlv_index* h) {
    for (FilterBufs& b : h->filter_pool)
        for (DevBuf* d : {&b.bitmap, &b.list, &b.scratch}) free_dev(*d);
    h->filter_pool.clear();
}

// Shared tail of mlv_filter_create / mlv_filter_create_where: the bitmap is in f->d_bitmap (work enqueued on
// h->stream); build the passing-row list and read the count.
int finish_filter(mlv_index* h, mlv_filter* f, mlv_filter_t* out) {
    int rc = MLV_OK;
    if (h->rows) rc = build_gather_list(h, f->d_list, f->d_scratch, (const uint32_t*)f->d_bitmap.p, f->bitmap_words, h->stream);
    uint64_t total = 0;
    if (rc == MLV_OK) {
        cudaError_t e = cudaSuccess;
        if (h->rows) e = cudaMemcpyAsync(&total, f->d_scratch.p, 8, cudaMemcpyDeviceToHost, h->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
        if (e != cudaSuccess) rc = fail_cuda(h, e, "filter build");
    }
    if (rc != MLV_OK) {
        retire_filter(h, f);
        return rc;
    }
    f->passing = total;
    f->counted = true;
    f->epoch = h->rows ? h->epoch : ~0ull;
    *out = f;
    return MLV_OK;
}

}  // namespace
