// column_kernels.cuh -- columnar metadata beside the row matrix.
//
// The reference keeps metadata as an arbitrary host mapping per vector (implementations/vector.py:15) and
// has no filter code at all (README.md:123,477; request shape examples/api_client.py:65-74: a dict of
// equality constraints).  Here a namespace may carry int32 "code" columns, one value per row, resident in
// HBM; a conjunction of comparisons over them is evaluated by where_kernel into the same filter bitmap a
// caller could have uploaded (mlv_filter_create), so a predicate over 10M rows costs one pass over
// 4 bytes per row and predicate instead of a Python loop over 10M dicts.  HBM-bound integer work.
#pragma once
#include "common.cuh"

namespace mlv {

constexpr int32_t COLUMN_MISSING = INT32_MIN;   // == MLV_COLUMN_MISSING: fails every predicate
constexpr uint32_t WHERE_MAX_PREDS = 8;         // == MLV_MAX_PREDICATES

struct WherePred {
    const int32_t* col;   // device column
    uint64_t col_rows;    // rows the column allocation covers; rows beyond are missing
    int32_t op, a, b;
};
struct WhereArgs {
    WherePred p[WHERE_MAX_PREDS];
    uint32_t n;
};

__device__ __forceinline__ bool where_eval(int32_t op, int32_t v, int32_t a, int32_t b) {
    switch (op) {
        case 0: return v == a;            // MLV_OP_EQ
        case 1: return v != a;            // MLV_OP_NE
        case 2: return v < a;             // MLV_OP_LT
        case 3: return v <= a;            // MLV_OP_LE
        case 4: return v > a;             // MLV_OP_GT
        case 5: return v >= a;            // MLV_OP_GE
        default: return v >= a && v <= b; // MLV_OP_BETWEEN
    }
}

// One warp per group of WORDS_PER_STEP bitmap words (32 rows each): lane = row inside the word, every
// column read is a coalesced 128-byte line, __ballot_sync packs the word.  All loads of a step are
// issued before the first comparison.
template <int WORDS_PER_STEP>
__global__ void __launch_bounds__(256) where_kernel(WhereArgs args, uint64_t n_rows, uint32_t* __restrict__ bitmap) {
    const int lane = threadIdx.x & 31;
    const uint64_t n_words = (n_rows + 31) >> 5;
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t w0 = warp * WORDS_PER_STEP; w0 < n_words; w0 += n_warps * WORDS_PER_STEP) {
        bool pass[WORDS_PER_STEP];
#pragma unroll
        for (int j = 0; j < WORDS_PER_STEP; j++) pass[j] = (w0 + j) * 32 + lane < n_rows;
        for (uint32_t pi = 0; pi < args.n; pi++) {
            const WherePred& p = args.p[pi];
            int32_t v[WORDS_PER_STEP];
#pragma unroll
            for (int j = 0; j < WORDS_PER_STEP; j++) {
                const uint64_t r = (w0 + j) * 32 + lane;
                v[j] = r < p.col_rows ? __ldg(p.col + r) : COLUMN_MISSING;
            }
#pragma unroll
            for (int j = 0; j < WORDS_PER_STEP; j++)
                pass[j] = pass[j] && v[j] != COLUMN_MISSING && where_eval(p.op, v[j], p.a, p.b);
        }
#pragma unroll
        for (int j = 0; j < WORDS_PER_STEP; j++) {
            const uint32_t word = __ballot_sync(0xffffffffu, pass[j]);
            if (lane == 0 && w0 + j < n_words) bitmap[w0 + j] = word;
        }
    }
}

__global__ void fill_i32_kernel(int32_t* p, uint64_t n, int32_t value) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) p[i] = value;
}

// compaction of one column with the row matrix: live rows keep their order (word_base from live_prefix_kernel)
__global__ void compact_column_kernel(const int32_t* __restrict__ src, int32_t* __restrict__ dst, const uint32_t* live,
                                      const uint64_t* word_base, uint64_t n_rows) {
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_rows; r += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t word = live[r >> 5];
        const uint32_t bit = (uint32_t)(r & 31);
        if ((word >> bit) & 1u) dst[word_base[r >> 5] + __popc(word & ((1u << bit) - 1))] = src[r];
    }
}

}  // namespace mlv
