// maint_kernels.cuh -- everything around the scan that touches the device matrix:
// ingest (pad + normalise), the deterministic synthetic generator, tombstone bitmap updates,
// compaction and query preparation.  None of these is on the per-query critical path except
// prep_queries_kernel (one tiny launch per search call).
#pragma once
#include "common.cuh"

namespace mlv {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

// hnswlib bindings.cpp normalize_vector: inv = 1 / (sqrt(sum v_i^2) + 1e-30); v_i *= inv.
// One warp per row; rows already padded to ld (padding is zero and stays zero).
__global__ void normalize_rows_kernel(float* rows, uint64_t first, uint64_t n, uint32_t ld) {
    const uint64_t w = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (w >= n) return;
    float4* r = reinterpret_cast<float4*>(rows + (first + w) * ld);
    const uint32_t ld4 = ld >> 2;
    float s = 0.f;
    for (uint32_t j = lane; j < ld4; j += 32) {
        const float4 x = r[j];
        s = fmaf(x.x, x.x, s);
        s = fmaf(x.y, x.y, s);
        s = fmaf(x.z, x.z, s);
        s = fmaf(x.w, x.w, s);
    }
    s = warp_sum(s);
    const float inv = 1.0f / (sqrtf(s) + 1e-30f);
    for (uint32_t j = lane; j < ld4; j += 32) {
        float4 x = r[j];
        x.x *= inv;
        x.y *= inv;
        x.z *= inv;
        x.w *= inv;
        r[j] = x;
    }
}

// Copy densely packed [n, dim] rows into the padded matrix rows [first, first+n) x ld.
__global__ void pad_rows_kernel(const float* __restrict__ src, float* __restrict__ dst, uint64_t first, uint64_t n,
                                uint32_t dim, uint32_t ld) {
    const uint64_t total = n * ld;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t r = i / ld;
        const uint32_t c = (uint32_t)(i - r * ld);
        dst[(first + r) * ld + c] = c < dim ? src[r * dim + c] : 0.f;
    }
}

// Queries: [nq, dim] raw -> [nq, ld] zero padded, normalised when cosine.  One warp per query.
__global__ void prep_queries_kernel(const float* __restrict__ src, float* __restrict__ dst, uint32_t nq, uint32_t dim,
                                    uint32_t ld, int normalize) {
    const uint32_t w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (w >= nq) return;
    const float* s = src + (size_t)w * dim;
    float* d = dst + (size_t)w * ld;
    float inv = 1.0f;
    if (normalize) {
        float acc = 0.f;
        for (uint32_t j = lane; j < dim; j += 32) acc = fmaf(s[j], s[j], acc);
        acc = warp_sum(acc);
        inv = 1.0f / (sqrtf(acc) + 1e-30f);
    }
    for (uint32_t j = lane; j < ld; j += 32) d[j] = j < dim ? (normalize ? s[j] * inv : s[j]) : 0.f;
}

// ---- deterministic synthetic rows (bit-identical to oracle/exact_scan.c::orc_fill_synthetic) ----
__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    uint64_t z = x + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__global__ void fill_synthetic_kernel(float* rows, uint64_t first_local, uint64_t n, uint32_t dim, uint32_t ld,
                                      uint64_t key, uint64_t key2, uint64_t first_gen_row, int scaled) {
    const uint64_t total = n * ld;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t r = i / ld;
        const uint32_t c = (uint32_t)(i - r * ld);
        float v = 0.f;
        if (c < dim) {
            const uint64_t grow = first_gen_row + r;
            const uint64_t h = splitmix64(key + grow * dim + c);
            v = (float)(uint32_t)(h >> 40) * 1.1920928955078125e-07f - 1.0f;
            if (scaled) {
                const uint64_t hs = splitmix64(key2 + grow);
                const float s = 0.5f + (float)(uint32_t)(hs >> 40) * 5.9604644775390625e-08f;
                v = v * s;
            }
        }
        rows[(first_local + r) * ld + c] = v;
    }
}

// ---- tombstone bitmap ---------------------------------------------------------------------------
// set bits [first, first+n) (rows just appended are live)
__global__ void set_live_range_kernel(uint32_t* live, uint64_t first, uint64_t n) {
    const uint64_t w0 = first >> 5, w1 = (first + n - 1) >> 5;
    for (uint64_t w = w0 + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; w <= w1; w += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t mask = 0xffffffffu;
        if (w == w0) mask &= 0xffffffffu << (first & 31);
        if (w == w1) mask &= 0xffffffffu >> (31 - ((first + n - 1) & 31));
        atomicOr(live + w, mask);
    }
}
__global__ void mark_deleted_kernel(uint32_t* live, const uint64_t* rows, uint64_t n, uint64_t n_rows,
                                    unsigned long long* changed) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t r = rows[i];
        if (r >= n_rows) continue;
        const uint32_t bit = 1u << (r & 31);
        const uint32_t old = atomicAnd(live + (r >> 5), ~bit);
        if (old & bit) atomicAdd(changed, 1ull);
    }
}

// ---- compaction ---------------------------------------------------------------------------------
// word_base[w] = number of live rows before word w (single CTA, chunked scan with carry).
// `mask` (nullable) is ANDed in: the prefix of rows that are live AND pass a filter bitmap.
__global__ void __launch_bounds__(1024, 1) live_prefix_kernel(const uint32_t* live, uint64_t n_rows, uint64_t* word_base,
                                                              uint64_t* total_live, const uint32_t* mask = nullptr,
                                                              uint64_t mask_words = 0) {
    __shared__ uint32_t warp_sums[32];
    __shared__ uint64_t carry;
    const uint64_t n_words = (n_rows + 31) >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint64_t base = 0; base < n_words; base += blockDim.x) {
        const uint64_t w = base + threadIdx.x;
        uint32_t word = 0;
        if (w < n_words) {
            word = live[w];
            if (mask) word &= w < mask_words ? mask[w] : 0u;
            if (w == n_words - 1 && (n_rows & 31)) word &= (1u << (n_rows & 31)) - 1;
        }
        const uint32_t c = __popc(word);
        uint32_t incl = c;  // inclusive warp scan
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, incl, off);
            if (lane >= off) incl += t;
        }
        if (lane == 31) warp_sums[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            uint32_t v = lane < (int)(blockDim.x >> 5) ? warp_sums[lane] : 0;
            uint32_t inc2 = v;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                uint32_t t = __shfl_up_sync(0xffffffffu, inc2, off);
                if (lane >= off) inc2 += t;
            }
            warp_sums[lane] = inc2 - v;  // exclusive
        }
        __syncthreads();
        const uint64_t excl = carry + warp_sums[wid] + (incl - c);
        if (w < n_words) word_base[w] = excl;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry = excl + c;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total_live = carry;
}
// ---- passing-row list (the gather list of the scan kernel): rows that are live AND pass `mask`, ascending ----
// Two launches over the bitmap words, 1024 words per CTA: block popcounts, then every CTA sums the
// popcounts of the CTAs before it (a few hundred values), scans its own words and scatters the row
// numbers.  ~10 us at 10M rows; the single-CTA prefix above took ~300 us, which a search with a per-call
// filter bitmap paid on every query.
constexpr int LIST_BLOCK_WORDS = 1024;

__device__ __forceinline__ uint32_t passing_word(const uint32_t* live, const uint32_t* mask, uint64_t mask_words, uint64_t n_rows,
                                                 uint64_t n_words, uint64_t w) {
    if (w >= n_words) return 0u;
    uint32_t word = live[w];
    if (mask) word &= w < mask_words ? mask[w] : 0u;
    if (w == n_words - 1 && (n_rows & 31)) word &= (1u << (n_rows & 31)) - 1;
    return word;
}

// sum over the CTA (blockDim.x == LIST_BLOCK_WORDS); every thread gets the total, `excl` its exclusive prefix
__device__ __forceinline__ uint32_t block_scan_1024(uint32_t v, uint32_t* excl, uint32_t* warp_sums) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += t;
    }
    if (lane == 31) warp_sums[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        const uint32_t w = warp_sums[lane];
        uint32_t inc2 = w;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, inc2, off);
            if (lane >= off) inc2 += t;
        }
        warp_sums[lane] = inc2 - w;          // exclusive prefix of the warps
        if (lane == 31) warp_sums[32] = inc2;  // total
    }
    __syncthreads();
    *excl = warp_sums[wid] + incl - v;
    const uint32_t total = warp_sums[32];
    __syncthreads();
    return total;
}

__global__ void __launch_bounds__(LIST_BLOCK_WORDS) passing_block_sums_kernel(const uint32_t* live, const uint32_t* mask,
                                                                             uint64_t mask_words, uint64_t n_rows,
                                                                             uint32_t* block_sums) {
    __shared__ uint32_t warp_sums[33];
    const uint64_t n_words = (n_rows + 31) >> 5;
    const uint64_t w = (uint64_t)blockIdx.x * LIST_BLOCK_WORDS + threadIdx.x;
    uint32_t excl;
    const uint32_t total = block_scan_1024(__popc(passing_word(live, mask, mask_words, n_rows, n_words, w)), &excl, warp_sums);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(LIST_BLOCK_WORDS) scatter_passing_rows_kernel(const uint32_t* live, const uint32_t* mask,
                                                                               uint64_t mask_words, uint64_t n_rows,
                                                                               const uint32_t* block_sums, uint64_t* total_out,
                                                                               uint32_t* list) {
    __shared__ uint32_t warp_sums[33];
    const uint64_t n_words = (n_rows + 31) >> 5;
    // rows passing in the CTAs before this one (u32 is enough: a shard holds < 2^32 rows)
    uint32_t before = 0;
    for (uint32_t b = threadIdx.x; b < blockIdx.x; b += LIST_BLOCK_WORDS) before += block_sums[b];
    uint32_t unused;
    const uint32_t base = block_scan_1024(before, &unused, warp_sums);
    const uint64_t w = (uint64_t)blockIdx.x * LIST_BLOCK_WORDS + threadIdx.x;
    uint32_t word = passing_word(live, mask, mask_words, n_rows, n_words, w);
    uint32_t excl;
    const uint32_t mine = block_scan_1024(__popc(word), &excl, warp_sums);
    uint64_t at = (uint64_t)base + excl;
    while (word) {
        const int b = __ffs(word) - 1;
        word &= word - 1;
        list[at++] = (uint32_t)(w * 32 + b);
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) *total_out = (uint64_t)base + mine;
}

// one warp per old row: live rows move to their new position, map[old] = new or -1
__global__ void compact_rows_kernel(const float4* __restrict__ src, float4* __restrict__ dst, const uint32_t* live,
                                    const uint64_t* word_base, uint64_t n_rows, uint32_t ld4, int64_t* map) {
    const uint64_t r = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= n_rows) return;
    const uint32_t word = live[r >> 5];
    const uint32_t bit = (uint32_t)(r & 31);
    if (!((word >> bit) & 1u)) {
        if (lane == 0 && map) map[r] = -1;
        return;
    }
    const uint64_t nr = word_base[r >> 5] + __popc(word & ((1u << bit) - 1));
    for (uint32_t j = lane; j < ld4; j += 32) dst[nr * ld4 + j] = src[r * ld4 + j];
    if (lane == 0 && map) map[r] = (int64_t)nr;
}

}  // namespace mlv
