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

// A predicate as the kernel evaluates it: the host turns every operator into "lo <= value <= lo + span",
// optionally negated (MLV_OP_NE), so the inner loop is one subtract and two compares per value instead of a
// seven-way switch (the switch made the kernel instruction-bound: 55 % SM vs 37 % DRAM throughput in ncu).
// lo > INT32_MIN for every satisfiable range, so a missing value never lies inside one; an unsatisfiable
// predicate is the range [COLUMN_MISSING, COLUMN_MISSING], which only the missing value hits and the missing
// test rejects.
struct WherePred {
    const int32_t* col;   // device column
    uint64_t col_rows;    // rows the column allocation covers; rows beyond are missing
    int32_t lo;
    uint32_t span;
    uint32_t negate;
};
struct WhereArgs {
    WherePred p[WHERE_MAX_PREDS];
    uint32_t n;
};

// enum mlv_pred_op -> range form
__host__ inline void where_range(int32_t op, int32_t a, int32_t b, WherePred* out) {
    const int64_t MINV = (int64_t)INT32_MIN + 1, MAXV = INT32_MAX;
    int64_t lo = MINV, hi = MAXV;
    out->negate = 0;
    switch (op) {
        case 0: lo = a; hi = a; break;                       // MLV_OP_EQ
        case 1: lo = a; hi = a; out->negate = 1; break;      // MLV_OP_NE
        case 2: hi = (int64_t)a - 1; break;                  // MLV_OP_LT
        case 3: hi = a; break;                               // MLV_OP_LE
        case 4: lo = (int64_t)a + 1; break;                  // MLV_OP_GT
        case 5: lo = a; break;                               // MLV_OP_GE
        default: lo = a; hi = b; break;                      // MLV_OP_BETWEEN
    }
    if (lo < MINV) lo = out->negate ? lo : MINV;             // NE COLUMN_MISSING: nothing stored equals it
    if (lo > hi || lo < (int64_t)INT32_MIN) {                // unsatisfiable
        out->lo = COLUMN_MISSING;
        out->span = 0;
        return;
    }
    out->lo = (int32_t)lo;
    out->span = (uint32_t)(hi - lo);
}

__device__ __forceinline__ uint32_t where_eval(const WherePred& p, int32_t v) {
    const bool inside = (uint32_t)v - (uint32_t)p.lo <= p.span;
    return (uint32_t)((inside != (bool)p.negate) && v != COLUMN_MISSING);
}

// One warp per step of GROUPS x 128 rows: lane l reads rows 4l .. 4l+3 of a group with ONE 16-byte load per
// predicate (512 contiguous bytes per warp and load; every load of a step is issued before the first comparison),
// turns its four verdicts into a nibble and the 8 lanes that share a bitmap word OR their nibbles together with
// three shuffles.  Column allocations are whole multiples of 32 rows (reserve_rows), so a 4-row load is
// either entirely inside a column or entirely beyond it (= missing).
template <int GROUPS>
__global__ void __launch_bounds__(256) where_kernel(WhereArgs args, uint64_t n_rows, uint32_t* __restrict__ bitmap) {
    const int lane = threadIdx.x & 31;
    const uint64_t n_groups = (n_rows + 127) >> 7;
    const uint64_t n_words = (n_rows + 31) >> 5;
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t g0 = warp * GROUPS; g0 < n_groups; g0 += n_warps * GROUPS) {
        uint32_t pass[GROUPS];
#pragma unroll
        for (int j = 0; j < GROUPS; j++) {
            const uint64_t r = ((g0 + j) << 7) + 4 * lane;
            pass[j] = r + 3 < n_rows ? 0xFu : (r < n_rows ? (1u << (n_rows - r)) - 1u : 0u);
        }
        for (uint32_t pi = 0; pi < args.n; pi++) {
            const WherePred p = args.p[pi];
            int4 v[GROUPS];
#pragma unroll
            for (int j = 0; j < GROUPS; j++) {
                const uint64_t r = ((g0 + j) << 7) + 4 * lane;
                v[j] = r < p.col_rows ? __ldg(reinterpret_cast<const int4*>(p.col + r))
                                      : make_int4(COLUMN_MISSING, COLUMN_MISSING, COLUMN_MISSING, COLUMN_MISSING);
            }
#pragma unroll
            for (int j = 0; j < GROUPS; j++) {
                const int32_t e[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
                uint32_t ok = 0;
#pragma unroll
                for (int c = 0; c < 4; c++) ok |= where_eval(p, e[c]) << c;
                pass[j] &= ok;
            }
        }
#pragma unroll
        for (int j = 0; j < GROUPS; j++) {
            uint32_t word = pass[j] << (4 * (lane & 7));
            word |= __shfl_xor_sync(0xffffffffu, word, 1);
            word |= __shfl_xor_sync(0xffffffffu, word, 2);
            word |= __shfl_xor_sync(0xffffffffu, word, 4);
            const uint64_t w = ((g0 + j) << 2) + (lane >> 3);
            if ((lane & 7) == 0 && w < n_words) bitmap[w] = word;
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
