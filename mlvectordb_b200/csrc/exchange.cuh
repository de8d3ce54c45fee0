// exchange.cuh -- the one exchange step of a row-sharded search done over
// NVLink peer memory inside the search kernel instead of an NCCL all-gather + merge launch.
//
// Every rank owns one exchange buffer (cudaMalloc, opened by all peers through CUDA IPC).  The
// last CTA of a rank's scan writes its k best keys per query straight into every peer's buffer
// (plain stores to peer-mapped addresses), publishes a sequence number with st.release.sys and
// waits with ld.acquire.sys until every peer's sequence number has arrived in its own buffer;
// then it merges the world * k candidates.  Slots are indexed by seq mod XCHG_SLOTS and at most XCHG_MAX_IN_FLIGHT exchange
// searches are in flight per handle (mlv_index_submit enforces it; the synchronous entry points run one
// at a time); the slot count follows from that (see XCHG_SLOTS).
// The reference has no counterpart (single process; README.md:142-155 sketches sharding only).
#pragma once
#include "common.cuh"

namespace mlv {

constexpr uint32_t XCHG_MAX_WORLD = 16;
constexpr uint32_t XCHG_MAX_NQ = 8;   // queries per scan launch
constexpr uint32_t XCHG_MAX_K = 64;   // the fused exchange handles k <= 64 (55 in practice: 148 SMs * k <= 8192 keys)
constexpr uint32_t XCHG_SLOT_KEYS = XCHG_MAX_WORLD * XCHG_MAX_NQ * XCHG_MAX_K;  // u64 keys per parity slot
// A search may take two launches (first tier + conditional fp32), i.e. two sequence numbers, and up to FOUR exchange
// searches may be in flight per handle: a rank launches search i + 8 (the next user of search i's slots) only after it
// collected search i + 4, i.e. after every peer posted i + 4 -- and a peer that posts i + 4 has at most i + 1 .. i + 3 still
// in flight, so its read of search i's slot is over.  16 slots = 4 x (searches in flight).
constexpr uint32_t XCHG_SLOTS = 16;
constexpr int XCHG_MAX_IN_FLIGHT = 4;
// buffer layout (u64 words): keys[XCHG_SLOTS][XCHG_MAX_WORLD][XCHG_MAX_NQ][XCHG_MAX_K], flags[XCHG_SLOTS][XCHG_MAX_WORLD]
constexpr uint32_t XCHG_FLAGS_OFF = XCHG_SLOTS * XCHG_SLOT_KEYS;
// ... and, for range searches, counts[XCHG_SLOTS][XCHG_MAX_WORLD]: how many hits a rank's list holds (bit 63: too many
// for its share of the slot, the low bits then carry the true total)
constexpr uint32_t XCHG_COUNTS_OFF = XCHG_FLAGS_OFF + XCHG_SLOTS * XCHG_MAX_WORLD;
constexpr uint32_t XCHG_WORDS = XCHG_COUNTS_OFF + XCHG_SLOTS * XCHG_MAX_WORLD;
constexpr unsigned long long RANGE_OVERFLOW = 1ull << 63;   // in a range count: the lists did not fit the exchange slot
constexpr unsigned long long RANGE_TIMEOUT = ~0ull;         // range count when a peer never arrived
// a peer that never arrives: flag an error, do not hang.  Default; MLV_EXCHANGE_TIMEOUT_MS / mlv_exchange_set_timeout_ms
// change it (the wait holds the GPU, so a serving process wants it short; a skewed batch job wants it long).
constexpr unsigned long long XCHG_TIMEOUT_NS = 5000000000ull;

struct ExchangeView {
    uint64_t* bufs[XCHG_MAX_WORLD];      // every rank's buffer as mapped into THIS process (bufs[rank] is local)
    uint64_t row_bases[XCHG_MAX_WORLD];  // global row of each rank's local row 0 (ascending with rank)
    uint32_t world, rank;
    uint64_t seq;                        // 1-based, identical on all ranks for the same scan launch
    int* error;                          // set to 1 when a peer did not answer in time
    unsigned long long timeout_ns;
};

__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// Called by `nthr` threads (tid 0..nthr-1, synchronised with named barrier 1).  `top` holds this
// rank's nq * k keys (local rows, ascending per query, KEY_SENTINEL padded).  Writes the global
// top-k of every query.  *s_valid is a zeroed shared counter.
// with_marks: the ranks also exchange one word each -- `my_mark` -- in the last key slot of their query 0 (k < XCHG_MAX_K
// then); the return value is the OR of all ranks' words, identical on every rank (0 without marks or after a timeout).
__device__ __forceinline__ uint32_t exchange_and_merge(const ExchangeView& x, const uint64_t* top, uint32_t nq, uint32_t k,
                                                       float* out_dists, int64_t* out_rows, int32_t* out_counts, uint32_t tid,
                                                       uint32_t nthr, uint32_t* s_valid, bool with_marks = false,
                                                       uint32_t my_mark = 0) {
    const uint32_t parity = (uint32_t)(x.seq % XCHG_SLOTS);
    const uint32_t per_peer = nq * k;
    for (uint32_t i = tid; i < x.world * per_peer; i += nthr) {
        const uint32_t dst = i / per_peer, r = i - dst * per_peer;
        const uint32_t qi = r / k, j = r - qi * k;
        x.bufs[dst][(size_t)parity * XCHG_SLOT_KEYS + ((size_t)x.rank * XCHG_MAX_NQ + qi) * XCHG_MAX_K + j] = top[r];
    }
    if (with_marks && tid < x.world)
        x.bufs[tid][(size_t)parity * XCHG_SLOT_KEYS + ((size_t)x.rank * XCHG_MAX_NQ) * XCHG_MAX_K + (XCHG_MAX_K - 1)] = my_mark;
    __threadfence_system();
    named_bar_sync(1, nthr);
    if (tid < x.world) {
        uint64_t* flag = x.bufs[tid] + XCHG_FLAGS_OFF + (size_t)parity * XCHG_MAX_WORLD + x.rank;
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(flag), "l"(x.seq) : "memory");
        // ... and wait for rank `tid`'s list in my own buffer
        const uint64_t* mine = x.bufs[x.rank] + XCHG_FLAGS_OFF + (size_t)parity * XCHG_MAX_WORLD + tid;
        const unsigned long long t0 = global_timer_ns();
        for (;;) {
            uint64_t v;
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(mine) : "memory");
            if (v >= x.seq) break;
            if (global_timer_ns() - t0 > x.timeout_ns) {
                *x.error = 1;
                atomicOr(s_valid, 0x80000000u);
                break;
            }
        }
    }
    named_bar_sync(1, nthr);
    if (*s_valid & 0x80000000u) {  // a peer never arrived: report count -1 instead of a partial merge
        for (uint32_t qi = tid; qi < nq; qi += nthr) out_counts[qi] = -1;
        return 0;
    }
    const volatile uint64_t* slot = x.bufs[x.rank] + (size_t)parity * XCHG_SLOT_KEYS;
    uint32_t marks = 0;
    if (with_marks)
        for (uint32_t r = 0; r < x.world; r++) marks |= (uint32_t)slot[((size_t)r * XCHG_MAX_NQ) * XCHG_MAX_K + (XCHG_MAX_K - 1)];
    const uint32_t mm = x.world * k;
    for (uint32_t qi = 0; qi < nq; qi++) {
        for (uint32_t e = tid; e < mm; e += nthr) {
            const uint32_t se = e / k, je = e - se * k;
            const uint64_t key = slot[((size_t)se * XCHG_MAX_NQ + qi) * XCHG_MAX_K + je];
            const uint32_t de = (uint32_t)(key >> 32);
            uint32_t rank = 0;
            for (uint32_t i = 0; i < mm; i++) {
                const uint32_t si = i / k, ji = i - si * k;
                const uint64_t o = slot[((size_t)si * XCHG_MAX_NQ + qi) * XCHG_MAX_K + ji];
                const uint32_t d_o = (uint32_t)(o >> 32);
                // (distance, global row) == (distance, rank, local row); the slot index breaks sentinel ties
                rank += (d_o < de) ||
                        (d_o == de && (si < se || (si == se && ((uint32_t)o < (uint32_t)key || ((uint32_t)o == (uint32_t)key && ji < je)))));
            }
            if (rank < k) {
                const bool valid = key != KEY_SENTINEL;
                out_dists[qi * k + rank] = valid ? key_dist(key) : __int_as_float(0x7f800000);
                out_rows[qi * k + rank] = valid ? (int64_t)(x.row_bases[se] + key_row(key)) : -1;
                if (valid) atomicAdd(s_valid, 1u);
            }
        }
        named_bar_sync(1, nthr);
        if (tid == 0) {
            out_counts[qi] = (int32_t)*s_valid;
            *s_valid = 0;
        }
        named_bar_sync(1, nthr);
    }
    return marks;
}

// Bitonic network over a[0..P) (P a power of two) by the nthr threads of named barrier 1.
__device__ __forceinline__ void cta_bitonic_sort(uint64_t* a, uint32_t P, uint32_t tid, uint32_t nthr) {
    for (uint32_t size = 2; size <= P; size <<= 1) {
        for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
            named_bar_sync(1, nthr);
            for (uint32_t t = tid; t < (P >> 1); t += nthr) {
                const uint32_t i = 2 * t - (t & (stride - 1));
                const uint32_t j = i + stride;
                const bool up = (i & size) == 0;
                const uint64_t va = a[i], vb = a[j];
                if ((va > vb) == up) {
                    a[i] = vb;
                    a[j] = va;
                }
            }
        }
    }
    named_bar_sync(1, nthr);
}

// Range search across row shards: the exchange step of SURVEY.md section 8e for hit LISTS.  Every rank's share of an
// exchange slot holds XCHG_SLOT_KEYS / world keys.  Called by `nthr` threads (named barrier 1) of the last CTA of a
// rank's range scan: `a` is shared scratch of XCHG_SLOT_KEYS keys whose first n_local entries are this rank's hits
// (distance | local row, any order); true_total is how many hits the scan found (> share: overflow).  The rank sorts
// its list, stores it into every peer's buffer together with its count, publishes the sequence number, waits for the
// peers and merges the world's lists ordered by (distance, rank, local row) = (distance, global row).  Writes the
// merged hits and *out_count = their number; RANGE_OVERFLOW | (sum of the true totals) when some rank's list did not
// fit (nothing else is written: the caller takes the all-gather path); RANGE_TIMEOUT when a peer never arrived.
// s_cnt: shared, XCHG_MAX_WORLD + 3 words ([0..world] prefix sums of the counts, then overflow marker, timeout flag).
__device__ __forceinline__ void range_exchange_and_merge(const ExchangeView& x, uint64_t* a, uint32_t n_local,
                                                         unsigned long long true_total, float* out_dists, int64_t* out_rows,
                                                         unsigned long long* out_count, uint32_t tid, uint32_t nthr,
                                                         unsigned long long* s_cnt) {
    const uint32_t share = XCHG_SLOT_KEYS / x.world;
    const uint32_t parity = (uint32_t)(x.seq % XCHG_SLOTS);
    const bool overflow = true_total > share;
    if (!overflow && n_local) {
        uint32_t P = 2;
        while (P < n_local) P <<= 1;
        for (uint32_t i = n_local + tid; i < P; i += nthr) a[i] = KEY_SENTINEL;
        cta_bitonic_sort(a, P, tid, nthr);
        for (uint32_t i = tid; i < x.world * n_local; i += nthr) {
            const uint32_t dst = i / n_local, j = i - dst * n_local;
            x.bufs[dst][(size_t)parity * XCHG_SLOT_KEYS + (size_t)x.rank * share + j] = a[j];
        }
    }
    if (tid < x.world)
        x.bufs[tid][XCHG_COUNTS_OFF + (size_t)parity * XCHG_MAX_WORLD + x.rank] = overflow ? (RANGE_OVERFLOW | true_total) : (unsigned long long)n_local;
    if (tid == 0) s_cnt[XCHG_MAX_WORLD + 2] = 0;
    __threadfence_system();
    named_bar_sync(1, nthr);
    if (tid < x.world) {
        uint64_t* flag = x.bufs[tid] + XCHG_FLAGS_OFF + (size_t)parity * XCHG_MAX_WORLD + x.rank;
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(flag), "l"(x.seq) : "memory");
        const uint64_t* mine = x.bufs[x.rank] + XCHG_FLAGS_OFF + (size_t)parity * XCHG_MAX_WORLD + tid;
        const unsigned long long t0 = global_timer_ns();
        for (;;) {
            uint64_t v;
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(mine) : "memory");
            if (v >= x.seq) break;
            if (global_timer_ns() - t0 > x.timeout_ns) {
                *x.error = 1;
                s_cnt[XCHG_MAX_WORLD + 2] = 1;
                break;
            }
        }
    }
    named_bar_sync(1, nthr);
    if (s_cnt[XCHG_MAX_WORLD + 2]) {
        if (tid == 0) *out_count = RANGE_TIMEOUT;
        return;
    }
    const volatile uint64_t* slot = x.bufs[x.rank] + (size_t)parity * XCHG_SLOT_KEYS;
    const volatile uint64_t* counts = x.bufs[x.rank] + XCHG_COUNTS_OFF + (size_t)parity * XCHG_MAX_WORLD;
    if (tid == 0) {   // prefix sums of the ranks' counts; s_cnt[world] = total, s_cnt[XCHG_MAX_WORLD + 1] = overflow marker
        unsigned long long sum = 0, any = 0, true_sum = 0;
        for (uint32_t r = 0; r < x.world; r++) {
            const unsigned long long c = counts[r];
            s_cnt[r] = sum;
            if (c & RANGE_OVERFLOW) any = 1;
            true_sum += c & ~RANGE_OVERFLOW;
            sum += (c & RANGE_OVERFLOW) ? 0 : c;
        }
        s_cnt[x.world] = sum;
        s_cnt[XCHG_MAX_WORLD + 1] = any ? (RANGE_OVERFLOW | true_sum) : 0;
    }
    named_bar_sync(1, nthr);
    if (s_cnt[XCHG_MAX_WORLD + 1]) {
        if (tid == 0) *out_count = s_cnt[XCHG_MAX_WORLD + 1];
        return;
    }
    const uint32_t total = (uint32_t)s_cnt[x.world];
    // merge keys: (distance, position in the rank-major concatenation); every list is ascending (distance, row), so
    // position order inside a rank is row order
    for (uint32_t r = 0; r < x.world; r++) {
        const uint32_t base = (uint32_t)s_cnt[r], n = (uint32_t)(s_cnt[r + 1] - s_cnt[r]);
        for (uint32_t j = tid; j < n; j += nthr)
            a[base + j] = (slot[(size_t)r * share + j] & 0xFFFFFFFF00000000ull) | (uint64_t)(r * share + j);
    }
    uint32_t P = 2;
    while (P < total) P <<= 1;
    named_bar_sync(1, nthr);
    for (uint32_t i = total + tid; i < P; i += nthr) a[i] = KEY_SENTINEL;
    cta_bitonic_sort(a, P, tid, nthr);
    for (uint32_t i = tid; i < total; i += nthr) {
        const uint64_t key = a[i];
        const uint32_t pos = (uint32_t)key, r = pos / share;
        out_dists[i] = key_dist(key);
        out_rows[i] = (int64_t)(x.row_bases[r] + key_row(slot[pos]));
    }
    if (tid == 0) *out_count = total;
}

// A rank with nothing to scan still takes part in a range exchange (count 0).
__global__ void __launch_bounds__(256, 1) range_exchange_only_kernel(const ExchangeView x, float* out_dists, int64_t* out_rows,
                                                                     unsigned long long* out_count) {
    extern __shared__ __align__(16) unsigned char xchg_smem_raw[];
    __shared__ unsigned long long s_cnt[XCHG_MAX_WORLD + 3];
    range_exchange_and_merge(x, reinterpret_cast<uint64_t*>(xchg_smem_raw), 0, 0, out_dists, out_rows, out_count, threadIdx.x,
                             blockDim.x, s_cnt);
}

// A rank with nothing to scan (empty shard, everything tombstoned) still has to take part.
// cert / run_if: this rank's part of a {first tier, conditional fp32} pair of launches (scan_kernel.cuh, ScanParams::cert):
// the first posts "certified" and records whether any rank was not, the second runs only in that case.
__global__ void __launch_bounds__(256, 1) exchange_only_kernel(const ExchangeView x, uint32_t nq, uint32_t k, float* out_dists,
                                                               int64_t* out_rows, int32_t* out_counts, uint32_t* cert = nullptr,
                                                               const uint32_t* run_if = nullptr) {
    __shared__ uint64_t top[XCHG_MAX_NQ * XCHG_MAX_K];
    __shared__ uint32_t s_valid;
    if (run_if && *reinterpret_cast<const volatile uint32_t*>(run_if) == 0) return;
    for (uint32_t i = threadIdx.x; i < nq * k; i += blockDim.x) top[i] = KEY_SENTINEL;
    if (threadIdx.x == 0) s_valid = 0;
    __syncthreads();
    const uint32_t marks = exchange_and_merge(x, top, nq, k, out_dists, out_rows, out_counts, threadIdx.x, blockDim.x, &s_valid,
                                              cert != nullptr, 0);
    if (cert && threadIdx.x == 0) *cert = marks ? 1u : 0u;
}

}  // namespace mlv
