// select_kernel.cuh -- second stage of the search: fold the per-block candidate lists the scan
// kernel wrote (k keys per (query, block)) into the final ascending top-k, and the same for the
// per-shard lists of a sharded search (SURVEY.md section 8e "final select kernel").
//
// One CTA sorts up to SELECT_MAX_P keys in shared memory with a bitonic network and keeps the
// first k.  When a query has more candidates than fit, the host runs the kernel as a tree:
// every CTA folds `lists_per_block` lists into one until a single list remains.
#pragma once
#include "common.cuh"

namespace mlv {

constexpr uint32_t SELECT_MAX_P = 8192;  // keys sorted per CTA (64 KB of shared memory)
constexpr int SELECT_THREADS = 1024;

__device__ __forceinline__ void bitonic_sort_smem(uint64_t* a, uint32_t P) {
    for (uint32_t size = 2; size <= P; size <<= 1) {
        for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            for (uint32_t t = threadIdx.x; t < (P >> 1); t += blockDim.x) {
                const uint32_t i = 2 * t - (t & (stride - 1));
                const uint32_t j = i + stride;
                const bool up = (i & size) == 0;
                const uint64_t x = a[i], y = a[j];
                if ((x > y) == up) {
                    a[i] = y;
                    a[j] = x;
                }
            }
        }
    }
    __syncthreads();
}

struct SelectParams {
    const uint64_t* in_keys;  // [nq][n_lists][k]
    uint32_t n_lists;
    uint32_t k;
    uint32_t lists_per_block;  // F
    uint32_t n_out_lists;      // ceil(n_lists / F); gridDim.x
    uint32_t P;                // power of two >= F * k
    uint64_t* out_keys;        // [nq][n_out_lists][k] when not final
    int final_pass;            // n_out_lists == 1 and outputs below are written instead
    float* out_dists;          // [nq][k]
    int64_t* out_rows;         // [nq][k]
    int32_t* out_counts;       // [nq]
    uint64_t row_base;
};

// grid = (n_out_lists, nq)
__global__ void __launch_bounds__(SELECT_THREADS, 1) select_kernel(const SelectParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* a = reinterpret_cast<uint64_t*>(smem_raw);
    const uint32_t q = blockIdx.y;
    const uint32_t first = blockIdx.x * p.lists_per_block;
    const uint32_t n_here = min(p.lists_per_block, p.n_lists - first);
    const uint32_t n_keys = n_here * p.k;
    const uint64_t* src = p.in_keys + ((size_t)q * p.n_lists + first) * p.k;
    for (uint32_t i = threadIdx.x; i < p.P; i += blockDim.x) a[i] = i < n_keys ? src[i] : KEY_SENTINEL;
    bitonic_sort_smem(a, p.P);
    if (!p.final_pass) {
        uint64_t* dst = p.out_keys + ((size_t)q * p.n_out_lists + blockIdx.x) * p.k;
        for (uint32_t i = threadIdx.x; i < p.k; i += blockDim.x) dst[i] = a[i];
        return;
    }
    __shared__ int cnt;
    if (threadIdx.x == 0) cnt = 0;
    __syncthreads();
    int local = 0;
    for (uint32_t i = threadIdx.x; i < p.k; i += blockDim.x) {
        const uint64_t key = a[i];
        const bool valid = key != KEY_SENTINEL;
        p.out_dists[(size_t)q * p.k + i] = valid ? key_dist(key) : __int_as_float(0x7f800000);
        p.out_rows[(size_t)q * p.k + i] = valid ? (int64_t)(p.row_base + key_row(key)) : -1;
        local += valid;
    }
    if (local) atomicAdd(&cnt, local);
    __syncthreads();
    if (threadIdx.x == 0) p.out_counts[q] = cnt;
}

// ---- sorting more keys than one CTA holds: bitonic network over global memory ------------------
// a[0..P) (P a power of two > SELECT_MAX_P, padded with KEY_SENTINEL).  Blocks of SELECT_MAX_P keys are
// handled in shared memory (all strides < SELECT_MAX_P of the merge steps size_lo..size_hi), the wider
// strides by one global compare-exchange pass each.  Directions follow the GLOBAL element index.
__global__ void __launch_bounds__(SELECT_THREADS, 1) bitonic_block_kernel(uint64_t* a, uint32_t size_lo, uint32_t size_hi) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* s = reinterpret_cast<uint64_t*>(smem_raw);
    const uint32_t base = blockIdx.x * SELECT_MAX_P;
    for (uint32_t i = threadIdx.x; i < SELECT_MAX_P; i += blockDim.x) s[i] = a[base + i];
    for (uint32_t size = size_lo; size <= size_hi; size <<= 1) {
        const uint32_t first = size >> 1 < SELECT_MAX_P ? size >> 1 : SELECT_MAX_P >> 1;
        for (uint32_t stride = first; stride > 0; stride >>= 1) {
            __syncthreads();
            for (uint32_t t = threadIdx.x; t < (SELECT_MAX_P >> 1); t += blockDim.x) {
                const uint32_t i = 2 * t - (t & (stride - 1));
                const uint32_t j = i + stride;
                const bool up = ((base + i) & size) == 0;
                const uint64_t x = s[i], y = s[j];
                if ((x > y) == up) {
                    s[i] = y;
                    s[j] = x;
                }
            }
        }
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < SELECT_MAX_P; i += blockDim.x) a[base + i] = s[i];
}
__global__ void bitonic_global_kernel(uint64_t* a, uint32_t P, uint32_t size, uint32_t stride) {
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < (P >> 1); t += gridDim.x * blockDim.x) {
        const uint32_t i = 2 * t - (t & (stride - 1));
        const uint32_t j = i + stride;
        const bool up = (i & size) == 0;
        const uint64_t x = a[i], y = a[j];
        if ((x > y) == up) {
            a[i] = y;
            a[j] = x;
        }
    }
}
// range hits -> (distance, row) arrays once the keys are in order
__global__ void decode_keys_kernel(const uint64_t* keys, uint64_t n, uint64_t row_base, float* out_dists, int64_t* out_rows) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        out_dists[i] = key_dist(keys[i]);
        out_rows[i] = (int64_t)(row_base + key_row(keys[i]));
    }
}
// (distance, GLOBAL row) pairs -> keys; padding entries (row < 0) and slots beyond n become sentinels
__global__ void encode_pairs_kernel(const float* dists, const int64_t* rows, uint64_t n, uint64_t P, uint64_t* keys) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < P; i += (uint64_t)gridDim.x * blockDim.x)
        keys[i] = (i < n && rows[i] >= 0) ? make_key(dists[i], (uint32_t)rows[i]) : KEY_SENTINEL;
}
__global__ void fill_sentinel_kernel(uint64_t* a, uint64_t from, uint64_t to) {
    for (uint64_t i = from + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < to; i += (uint64_t)gridDim.x * blockDim.x) a[i] = KEY_SENTINEL;
}

// ---- range search: order one query's hits -------------------------------------------------------
// grid = nq.  keys: [nq][slots] as appended by the scan (unordered).  Writes the first
// min(count, max_hits) hits ascending (distance, row) when they fit one CTA's sort (<= SELECT_MAX_P);
// larger hit lists are decoded unordered here and ordered by sort_big_device (mlv_index.cu).
__global__ void __launch_bounds__(SELECT_THREADS, 1)
range_finish_kernel(const uint64_t* keys, const unsigned long long* counts, uint64_t slots, uint64_t max_hits, uint64_t row_base,
                    float* out_dists, int64_t* out_rows, unsigned long long* out_counts) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* a = reinterpret_cast<uint64_t*>(smem_raw);
    const uint32_t q = blockIdx.x;
    const unsigned long long total = counts[q];
    const uint64_t n = total < max_hits ? total : max_hits;
    const uint64_t* mine = keys + (size_t)q * slots;
    if (threadIdx.x == 0) out_counts[q] = total;
    if (n == 0) return;
    if (n <= SELECT_MAX_P) {
        uint32_t P = 2;
        while (P < n) P <<= 1;
        for (uint32_t i = threadIdx.x; i < P; i += blockDim.x) a[i] = i < n ? mine[i] : KEY_SENTINEL;
        bitonic_sort_smem(a, P);
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
            out_dists[(size_t)q * max_hits + i] = key_dist(a[i]);
            out_rows[(size_t)q * max_hits + i] = (int64_t)(row_base + key_row(a[i]));
        }
    } else {  // decoded in arrival order; the host entry point runs the global bitonic network over them
        for (uint64_t i = threadIdx.x; i < n; i += blockDim.x) {
            out_dists[(size_t)q * max_hits + i] = key_dist(mine[i]);
            out_rows[(size_t)q * max_hits + i] = (int64_t)(row_base + key_row(mine[i]));
        }
    }
}

// ---- merge of (distance, row) pair lists from row shards ------------------------------------
struct MergePairsParams {
    const float* dists;   // [n_lists][nq][k]
    const int64_t* rows;  // [n_lists][nq][k]  (-1 = empty slot)
    uint32_t n_lists, nq, k, P;
    float* out_dists;     // [nq][k]
    int64_t* out_rows;    // [nq][k]
    int32_t* out_counts;  // [nq]
};

// grid = nq.  Ties on distance are broken by list position, which equals ascending global row
// because lists arrive rank-major (ascending row_base) and each list is (distance,row)-ascending.
__global__ void __launch_bounds__(SELECT_THREADS, 1) merge_pairs_kernel(const MergePairsParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* a = reinterpret_cast<uint64_t*>(smem_raw);
    const uint32_t q = blockIdx.x;
    const uint32_t n_keys = p.n_lists * p.k;
    for (uint32_t i = threadIdx.x; i < p.P; i += blockDim.x) {
        uint64_t key = KEY_SENTINEL;
        if (i < n_keys) {
            const uint32_t l = i / p.k, j = i - l * p.k;
            const size_t at = ((size_t)l * p.nq + q) * p.k + j;
            if (p.rows[at] >= 0) key = make_key(p.dists[at], i);
        }
        a[i] = key;
    }
    bitonic_sort_smem(a, p.P);
    __shared__ int cnt;
    if (threadIdx.x == 0) cnt = 0;
    __syncthreads();
    int local = 0;
    for (uint32_t i = threadIdx.x; i < p.k; i += blockDim.x) {
        const uint64_t key = a[i];
        const bool valid = key != KEY_SENTINEL;
        float d = __int_as_float(0x7f800000);
        int64_t r = -1;
        if (valid) {
            const uint32_t pos = key_row(key);
            const uint32_t l = pos / p.k, j = pos - l * p.k;
            const size_t at = ((size_t)l * p.nq + q) * p.k + j;
            d = p.dists[at];
            r = p.rows[at];
        }
        p.out_dists[(size_t)q * p.k + i] = d;
        p.out_rows[(size_t)q * p.k + i] = r;
        local += valid;
    }
    if (local) atomicAdd(&cnt, local);
    __syncthreads();
    if (threadIdx.x == 0) p.out_counts[q] = cnt;
}

}  // namespace mlv
