/*
 * exact_scan.c -- plain-C restatement of hnswlib 0.8.0's distance arithmetic plus a
 * brute-force (exact) top-k / range scan.  TEST ORACLE AND CPU BASELINE ONLY.
 *
 * PARITY UNPINNED: hnswlib (reference pyproject.toml:12, poetry.lock:144-152) is a
 * third-party dependency that is absent from /root/reference and from this image; the
 * reference's tests hold no golden vectors for this path.  The functions below restate the
 * published hnswlib 0.8.0 algorithm that the reference selects at
 * src/mlvectordb/implementations/index.py:36 and calls at index.py:65,111:
 *
 *   orc_normalize   bindings.cpp  normalize_vector():  norm += v[i]*v[i];
 *                                 norm = 1.0f/(sqrtf(norm)+1e-30f); out[i] = v[i]*norm
 *   orc_l2sqr       space_l2.h    L2Sqr():             t = a[i]-b[i]; res += t*t
 *   orc_ip          space_ip.h    InnerProduct():      res += a[i]*b[i]
 *                                 InnerProductDistance() = 1.0f - res
 *   *_simd16        space_l2.h / space_ip.h  *SIMD16Ext (AVX-512 build): 16 independent
 *                   partial sums over i mod 16, summed lane 0..15 in order; for dim % 16 != 0
 *                   the *Residuals variants add the scalar tail afterwards.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline / reference legs may
 * load this library.  The product never links it.
 *
 * Build: see oracle/Makefile (gcc -O3 -fopenmp -ffp-contract=off, no -ffast-math so the
 * summation orders written here are the ones executed).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_L2 0
#define ORC_IP 1
#define ORC_COSINE 2

/* ---------------------------------------------------------------- distances (scalar) */
float orc_l2sqr(const float *a, const float *b, size_t d) {
    float res = 0.0f;
    for (size_t i = 0; i < d; i++) {
        float t = a[i] - b[i];
        res += t * t;
    }
    return res;
}

float orc_ip(const float *a, const float *b, size_t d) {
    float res = 0.0f;
    for (size_t i = 0; i < d; i++) res += a[i] * b[i];
    return res;
}

/* ------------------------------------------------- distances (hnswlib SIMD16 ordering) */
float orc_l2sqr_simd16(const float *a, const float *b, size_t d) {
    float lane[16];
    for (int j = 0; j < 16; j++) lane[j] = 0.0f;
    size_t d16 = d & ~(size_t)15;
    for (size_t i = 0; i < d16; i += 16)
        for (int j = 0; j < 16; j++) {
            float t = a[i + j] - b[i + j];
            lane[j] += t * t;
        }
    float res = 0.0f;
    for (int j = 0; j < 16; j++) res += lane[j];
    float tail = 0.0f;
    for (size_t i = d16; i < d; i++) {
        float t = a[i] - b[i];
        tail += t * t;
    }
    return res + tail;
}

float orc_ip_simd16(const float *a, const float *b, size_t d) {
    float lane[16];
    for (int j = 0; j < 16; j++) lane[j] = 0.0f;
    size_t d16 = d & ~(size_t)15;
    for (size_t i = 0; i < d16; i += 16)
        for (int j = 0; j < 16; j++) lane[j] += a[i + j] * b[i + j];
    float res = 0.0f;
    for (int j = 0; j < 16; j++) res += lane[j];
    float tail = 0.0f;
    for (size_t i = d16; i < d; i++) tail += a[i] * b[i];
    return res + tail;
}

/* ---------------------------------------------------------------------- normalisation */
void orc_normalize(const float *in, float *out, size_t n, size_t d) {
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < (int64_t)n; r++) {
        const float *v = in + (size_t)r * d;
        float *o = out + (size_t)r * d;
        float norm = 0.0f;
        for (size_t i = 0; i < d; i++) norm += v[i] * v[i];
        norm = 1.0f / (sqrtf(norm) + 1e-30f);
        for (size_t i = 0; i < d; i++) o[i] = v[i] * norm;
    }
}

static inline float orc_dist(const float *x, const float *q, size_t d, int space, int simd16) {
    if (space == ORC_L2) return simd16 ? orc_l2sqr_simd16(x, q, d) : orc_l2sqr(x, q, d);
    return 1.0f - (simd16 ? orc_ip_simd16(x, q, d) : orc_ip(x, q, d));
}

/* all distances of one query (rows already normalised for cosine) */
void orc_distances(const float *rows, size_t n, size_t d, const float *q, int space, int simd16,
                   float *out) {
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < (int64_t)n; r++) out[r] = orc_dist(rows + (size_t)r * d, q, d, space, simd16);
}

/* ------------------------------------------------------------------- top-k by (d, label) */
typedef struct {
    float d;
    int64_t l;
} orc_cand;

static inline int cand_less(orc_cand a, orc_cand b) { return a.d < b.d || (a.d == b.d && a.l < b.l); }

/* max-heap on (d, l): heap[0] is the worst kept candidate */
static void heap_sift_down(orc_cand *h, int n, int i) {
    for (;;) {
        int l = 2 * i + 1, r = l + 1, m = i;
        if (l < n && cand_less(h[m], h[l])) m = l;
        if (r < n && cand_less(h[m], h[r])) m = r;
        if (m == i) return;
        orc_cand t = h[i];
        h[i] = h[m];
        h[m] = t;
        i = m;
    }
}

static void heap_push(orc_cand *h, int *n, int k, orc_cand c) {
    if (*n < k) {
        int i = (*n)++;
        h[i] = c;
        while (i > 0) {
            int p = (i - 1) / 2;
            if (!cand_less(h[p], h[i])) break;
            orc_cand t = h[i];
            h[i] = h[p];
            h[p] = t;
            i = p;
        }
    } else if (cand_less(c, h[0])) {
        h[0] = c;
        heap_sift_down(h, k, 0);
    }
}

static int cand_cmp(const void *a, const void *b) {
    orc_cand x = *(const orc_cand *)a, y = *(const orc_cand *)b;
    return cand_less(x, y) ? -1 : (cand_less(y, x) ? 1 : 0);
}

/*
 * Exact kNN of nq queries over rows[n, d].
 *   allow       optional bitmap over rows (bit r of word r/32, LSB first; 1 = candidate), or NULL
 *   first_label label of rows[0]
 *   out_labels  [nq, k] (-1 padded), out_dists [nq, k] (+inf padded), out_counts [nq]
 * Rows and queries must already be normalised for cosine.
 */
void orc_knn(const float *rows, size_t n, size_t d, const float *queries, size_t nq, int k, int space,
             int simd16, const uint32_t *allow, int64_t first_label, int64_t *out_labels, float *out_dists,
             int32_t *out_counts) {
    int nthreads = 1;
#ifdef _OPENMP
    nthreads = omp_get_max_threads();
#endif
    for (size_t qi = 0; qi < nq; qi++) {
        const float *q = queries + qi * d;
        orc_cand *heaps = (orc_cand *)malloc(sizeof(orc_cand) * (size_t)k * nthreads);
        int *counts = (int *)calloc(nthreads, sizeof(int));
#pragma omp parallel
        {
            int t = 0;
#ifdef _OPENMP
            t = omp_get_thread_num();
#endif
            orc_cand *h = heaps + (size_t)t * k;
            int cnt = 0;
#pragma omp for schedule(static)
            for (int64_t r = 0; r < (int64_t)n; r++) {
                if (allow && !((allow[r >> 5] >> (r & 31)) & 1u)) continue;
                orc_cand c;
                c.d = orc_dist(rows + (size_t)r * d, q, d, space, simd16);
                c.l = first_label + r;
                heap_push(h, &cnt, k, c);
            }
            counts[t] = cnt;
        }
        size_t total = 0;
        for (int t = 0; t < nthreads; t++) total += counts[t];
        orc_cand *all = (orc_cand *)malloc(sizeof(orc_cand) * (total ? total : 1));
        size_t p = 0;
        for (int t = 0; t < nthreads; t++) {
            memcpy(all + p, heaps + (size_t)t * k, sizeof(orc_cand) * counts[t]);
            p += counts[t];
        }
        qsort(all, total, sizeof(orc_cand), cand_cmp);
        int m = total < (size_t)k ? (int)total : k;
        for (int j = 0; j < k; j++) {
            out_labels[qi * k + j] = j < m ? all[j].l : -1;
            out_dists[qi * k + j] = j < m ? all[j].d : INFINITY;
        }
        out_counts[qi] = m;
        free(all);
        free(heaps);
        free(counts);
    }
}

/* ------------------------------------------------------- deterministic synthetic rows
 * Restated bit-for-bit by the CUDA generator (mlvectordb_b200/csrc, mlv_index_add_synthetic)
 * and by oracle/synthetic.py.  All float steps are exact or a single IEEE rounding, so CPU
 * and GPU produce identical bits.
 */
static inline uint64_t splitmix64(uint64_t x) {
    uint64_t z = x + 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

static inline float syn_elem(uint64_t key, uint64_t row, uint64_t col, uint64_t d) {
    uint64_t h = splitmix64(key + row * d + col);
    return (float)(uint32_t)(h >> 40) * 1.1920928955078125e-07f - 1.0f; /* m * 2^-23 - 1 in [-1, 1) */
}

static inline float syn_row_scale(uint64_t key2, uint64_t row) {
    uint64_t h = splitmix64(key2 + row);
    return 0.5f + (float)(uint32_t)(h >> 40) * 5.9604644775390625e-08f; /* 0.5 + m * 2^-24 in [0.5, 1.5) */
}

void orc_fill_synthetic(float *out, uint64_t seed, uint64_t first_row, uint64_t n, uint64_t d, int scaled) {
    uint64_t key = splitmix64(seed);
    uint64_t key2 = splitmix64(seed ^ 0xA5A5A5A5A5A5A5A5ULL);
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < (int64_t)n; r++) {
        uint64_t row = first_row + (uint64_t)r;
        float s = scaled ? syn_row_scale(key2, row) : 1.0f;
        float *o = out + (size_t)r * d;
        for (uint64_t c = 0; c < d; c++) {
            float v = syn_elem(key, row, c, d);
            o[c] = scaled ? v * s : v;
        }
    }
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void orc_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ------------------------------------------- streamed oracle over the synthetic generator
 * Exact kNN / range search over generator rows first_row .. first_row+n-1 WITHOUT materialising
 * them: every thread regenerates one row at a time (same syn_elem / syn_row_scale as
 * orc_fill_synthetic), normalises it like orc_normalize when space == cosine, and scores it with
 * orc_dist against all queries.  Bit-identical to orc_fill_synthetic -> orc_normalize -> orc_knn on
 * the same rows (tests/test_oracle.py checks that), but a 10M x 768 pass costs seconds and no
 * memory, so parity at BASELINE.json's full sizes (SURVEY.md 8d "run at full N") is affordable.
 *   allow   optional bitmap over (row - first_row), LSB first; 1 = candidate (tombstones AND filter)
 *   labels  generator row numbers
 * Queries must already be normalised for cosine.
 */
static void syn_row(float *o, uint64_t key, uint64_t key2, uint64_t row, uint64_t d, int scaled, int space) {
    float s = scaled ? syn_row_scale(key2, row) : 1.0f;
    for (uint64_t c = 0; c < d; c++) {
        float v = syn_elem(key, row, c, d);
        o[c] = scaled ? v * s : v;
    }
    if (space == ORC_COSINE) {
        float norm = 0.0f;
        for (uint64_t i = 0; i < d; i++) norm += o[i] * o[i];
        norm = 1.0f / (sqrtf(norm) + 1e-30f);
        for (uint64_t i = 0; i < d; i++) o[i] = o[i] * norm;
    }
}

void orc_knn_synthetic(uint64_t seed, uint64_t first_row, uint64_t n, uint64_t d, int scaled, const float *queries,
                       size_t nq, int k, int space, int simd16, const uint32_t *allow, int64_t *out_labels,
                       float *out_dists, int32_t *out_counts) {
    const uint64_t key = splitmix64(seed), key2 = splitmix64(seed ^ 0xA5A5A5A5A5A5A5A5ULL);
    int nthreads = 1;
#ifdef _OPENMP
    nthreads = omp_get_max_threads();
#endif
    orc_cand *heaps = (orc_cand *)malloc(sizeof(orc_cand) * (size_t)k * nq * nthreads);
    int *counts = (int *)calloc((size_t)nthreads * nq, sizeof(int));
#pragma omp parallel
    {
        int t = 0;
#ifdef _OPENMP
        t = omp_get_thread_num();
#endif
        float *row = (float *)malloc(sizeof(float) * d);
        orc_cand *h = heaps + (size_t)t * nq * k;
        int *cnt = counts + (size_t)t * nq;
#pragma omp for schedule(static)
        for (int64_t r = 0; r < (int64_t)n; r++) {
            if (allow && !((allow[r >> 5] >> (r & 31)) & 1u)) continue;
            syn_row(row, key, key2, first_row + (uint64_t)r, d, scaled, space);
            for (size_t qi = 0; qi < nq; qi++) {
                orc_cand c;
                c.d = orc_dist(row, queries + qi * d, d, space, simd16);
                c.l = (int64_t)(first_row + (uint64_t)r);
                heap_push(h + qi * k, &cnt[qi], k, c);
            }
        }
        free(row);
    }
    orc_cand *all = (orc_cand *)malloc(sizeof(orc_cand) * (size_t)k * nthreads);
    for (size_t qi = 0; qi < nq; qi++) {
        size_t total = 0;
        for (int t = 0; t < nthreads; t++) {
            int c = counts[(size_t)t * nq + qi];
            memcpy(all + total, heaps + ((size_t)t * nq + qi) * k, sizeof(orc_cand) * c);
            total += c;
        }
        qsort(all, total, sizeof(orc_cand), cand_cmp);
        int m = total < (size_t)k ? (int)total : k;
        for (int j = 0; j < k; j++) {
            out_labels[qi * k + j] = j < m ? all[j].l : -1;
            out_dists[qi * k + j] = j < m ? all[j].d : INFINITY;
        }
        out_counts[qi] = m;
    }
    free(all);
    free(heaps);
    free(counts);
}

/* Every candidate row with distance <= radius, ascending (distance, label).  out_counts[q] = total hits;
 * at most max_hits of them (the smallest) are written to out_labels / out_dists [nq, max_hits]. */
void orc_range_synthetic(uint64_t seed, uint64_t first_row, uint64_t n, uint64_t d, int scaled, const float *queries,
                         size_t nq, float radius, int space, int simd16, const uint32_t *allow, uint64_t max_hits,
                         int64_t *out_labels, float *out_dists, uint64_t *out_counts) {
    const uint64_t key = splitmix64(seed), key2 = splitmix64(seed ^ 0xA5A5A5A5A5A5A5A5ULL);
    int nthreads = 1;
#ifdef _OPENMP
    nthreads = omp_get_max_threads();
#endif
    /* per (thread, query) growable hit lists */
    orc_cand **lists = (orc_cand **)calloc((size_t)nthreads * nq, sizeof(orc_cand *));
    size_t *lens = (size_t *)calloc((size_t)nthreads * nq, sizeof(size_t));
    size_t *caps = (size_t *)calloc((size_t)nthreads * nq, sizeof(size_t));
#pragma omp parallel
    {
        int t = 0;
#ifdef _OPENMP
        t = omp_get_thread_num();
#endif
        float *row = (float *)malloc(sizeof(float) * d);
#pragma omp for schedule(static)
        for (int64_t r = 0; r < (int64_t)n; r++) {
            if (allow && !((allow[r >> 5] >> (r & 31)) & 1u)) continue;
            syn_row(row, key, key2, first_row + (uint64_t)r, d, scaled, space);
            for (size_t qi = 0; qi < nq; qi++) {
                float dist = orc_dist(row, queries + qi * d, d, space, simd16);
                if (!(dist <= radius)) continue;
                size_t s = (size_t)t * nq + qi;
                if (lens[s] == caps[s]) {
                    caps[s] = caps[s] ? caps[s] * 2 : 64;
                    lists[s] = (orc_cand *)realloc(lists[s], sizeof(orc_cand) * caps[s]);
                }
                lists[s][lens[s]].d = dist;
                lists[s][lens[s]].l = (int64_t)(first_row + (uint64_t)r);
                lens[s]++;
            }
        }
        free(row);
    }
    for (size_t qi = 0; qi < nq; qi++) {
        size_t total = 0;
        for (int t = 0; t < nthreads; t++) total += lens[(size_t)t * nq + qi];
        orc_cand *all = (orc_cand *)malloc(sizeof(orc_cand) * (total ? total : 1));
        size_t p = 0;
        for (int t = 0; t < nthreads; t++) {
            size_t s = (size_t)t * nq + qi;
            if (lens[s]) memcpy(all + p, lists[s], sizeof(orc_cand) * lens[s]);
            p += lens[s];
            free(lists[s]);
        }
        qsort(all, total, sizeof(orc_cand), cand_cmp);
        out_counts[qi] = total;
        for (uint64_t j = 0; j < max_hits && j < total; j++) {
            out_labels[qi * max_hits + j] = all[j].l;
            out_dists[qi * max_hits + j] = all[j].d;
        }
        free(all);
    }
    free(lists);
    free(lens);
    free(caps);
}

/* distances of generator rows `labels[0..m)` (regenerated one by one) to one query: adjudication of ids
 * that are in a result but not in the oracle's top-k (oracle/exact.py::check_topk_parity all_ref_scores) */
void orc_distances_synthetic(uint64_t seed, const int64_t *labels, size_t m, uint64_t d, int scaled, const float *q,
                             int space, int simd16, float *out) {
    const uint64_t key = splitmix64(seed), key2 = splitmix64(seed ^ 0xA5A5A5A5A5A5A5A5ULL);
    float *row = (float *)malloc(sizeof(float) * d);
    for (size_t i = 0; i < m; i++) {
        syn_row(row, key, key2, (uint64_t)labels[i], d, scaled, space);
        out[i] = orc_dist(row, q, d, space, simd16);
    }
    free(row);
}
