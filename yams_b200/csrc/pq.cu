// pq.cu -- the SimeonPqAdc engine scan: product-quantised asymmetric-distance scan + exact rerank.
//
// Reference (paths under /root/reference):
//   simeonPqSearchUnlocked                src/vector/sqlite_vec_backend.cpp:3868-4056
//   index build (normalise, encode)       src/vector/sqlite_vec_backend.cpp:3540-3690, normalizeEmbeddingInPlace :213-226
//   ProductQuantizer::encode / PQInnerProductQuery / inner_product_from_lut
//                                         third_party/simeon/src/pq.cpp:31-57,218-235,391-414
//   simd::dot (AVX2 tier on x86 builds)   third_party/simeon/src/arch/avx2.cpp:63-84, include/simeon/simd.hpp:187-201
//   computeCosineSimilarity (rerank)      src/vector/vector_database.cpp:1786-1810
//
// What the reference does per query: normalise q; LUT[mi][ki] = q_sub . centroid[mi][ki]; for EVERY indexed row the score is
// the sequential float sum of m LUT entries selected by the row's m code bytes; the best approxK = max(k, k * rerank_factor)
// rows by (score desc, tie-break key asc) are re-scored exactly (double cosine against the stored embedding), filtered by
// the threshold, ordered (similarity desc, chunk_id asc) and cut to k.
//
// Device layout: codes [n_idx][m] bytes (32 B per row at the default m = 32: 320 MB per 10 M rows, 1/48 of the fp16 corpus),
// the LUT of the query in shared memory (m * k floats = 32 KiB), one CTA per 4096-row tile: scores -> shared memory ->
// tile-local top-approxK by a bitonic network -> one short list per query -> the exact selection / rerank kernels of the exact
// scan (knn.cu).  Every float operation is in the reference's order, so the approximate scores, the survivor set and the final
// scores are bit-identical to the CPU engine (tests/test_gpu_pq.py pins them against simeon's own pq.cpp compiled in place).
// The scan is HBM-bound at m bytes per row per query; it is the LATENCY engine -- for batches beyond ~150 queries the exact
// tensor-core scan of the same corpus is faster and exact.
#include <cuda_fp16.h>

#include <algorithm>
#include <new>
#include <vector>

#include "knn.cuh"

namespace yb {

constexpr int PQ_TILE = 4096;        // rows per CTA
constexpr int PQ_THREADS = 256;
constexpr uint32_t PQ_MAX_APPROX = 1024;

// normalizeEmbeddingInPlace (sqlite_vec_backend.cpp:213-226): double sum of squares, float reciprocal square root, float scale
__device__ __forceinline__ bool normalize_factor(double norm_sq, float* inv) {
    if (norm_sq <= 1e-20) return false;
    *inv = __fdiv_rn(1.0f, sqrtf((float)norm_sq));
    return true;
}

__device__ __forceinline__ float load_row_elem(const void* rows, int dtype, uint64_t idx) {
    if (dtype == YAMS_B200_F16) return __half2float(reinterpret_cast<const __half*>(rows)[idx]);
    return reinterpret_cast<const float*>(rows)[idx];
}

// one thread per corpus row: is it indexable (norm_sq > 1e-20)?  keep[row] = 1/0
__global__ void pq_keep_kernel(const void* __restrict__ rows, int dtype, uint32_t d, uint64_t n, uint32_t* __restrict__ keep) {
    uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    double ss = 0.0;
    for (uint32_t c = 0; c < d; ++c) {
        float v = load_row_elem(rows, dtype, r * d + c);
        ss += (double)v * (double)v;
    }
    keep[r] = ss > 1e-20 ? 1u : 0u;
}
__global__ void pq_compact_kernel(const uint32_t* __restrict__ keep, const uint32_t* __restrict__ pos, uint64_t n, uint32_t* __restrict__ idx_rows) {
    uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n && keep[r]) idx_rows[pos[r]] = (uint32_t)r;
}

// ProductQuantizer::encode (pq.cpp:218-235) of one normalised row per CTA: thread ki owns centroid ki of every subspace,
// l2_sq is the reference's sequential float loop (separate multiply and add), argmin with the first index winning ties.
__global__ void __launch_bounds__(256) pq_encode_kernel(const void* __restrict__ rows, int dtype, uint32_t d, const uint32_t* __restrict__ idx_rows,
                                                        uint64_t n_idx, const float* __restrict__ codebooks, uint32_t m, uint32_t k,
                                                        uint8_t* __restrict__ codes) {
    extern __shared__ float xs[];                    // the normalised row, d floats
    __shared__ double s_norm;
    __shared__ float s_best[8];
    __shared__ uint32_t s_arg[8];
    const uint64_t j = blockIdx.x;
    if (j >= n_idx) return;
    const uint64_t row = idx_rows[j];
    const uint32_t dsub = d / m;
    for (uint32_t c = threadIdx.x; c < d; c += blockDim.x) xs[c] = load_row_elem(rows, dtype, row * d + c);
    __syncthreads();
    if (threadIdx.x == 0) {
        double ss = 0.0;
        for (uint32_t c = 0; c < d; ++c) ss += (double)xs[c] * (double)xs[c];
        s_norm = ss;
    }
    __syncthreads();
    float inv = 0.f;
    normalize_factor(s_norm, &inv);
    for (uint32_t c = threadIdx.x; c < d; c += blockDim.x) xs[c] = __fmul_rn(xs[c], inv);
    __syncthreads();
    const uint32_t ki = threadIdx.x;
    for (uint32_t mi = 0; mi < m; ++mi) {
        float dist = INFINITY;
        if (ki < k) {
            const float* cb = codebooks + ((size_t)mi * k + ki) * dsub;
            const float* x = xs + mi * dsub;
            float acc = 0.f;
            for (uint32_t e = 0; e < dsub; ++e) {
                const float df = __fsub_rn(x[e], cb[e]);
                acc = __fadd_rn(acc, __fmul_rn(df, df));
            }
            dist = acc;
        }
        // argmin over ki with `d < best_d` semantics: smallest distance, lowest index on equality; NaN never wins
        float bd = dist == dist ? dist : INFINITY;
        uint32_t bi = ki;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float od = __shfl_xor_sync(0xffffffffu, bd, o);
            const uint32_t oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (od < bd || (od == bd && oi < bi)) { bd = od; bi = oi; }
        }
        if ((threadIdx.x & 31) == 0) { s_best[threadIdx.x >> 5] = bd; s_arg[threadIdx.x >> 5] = bi; }
        __syncthreads();
        if (threadIdx.x == 0) {
            float fb = s_best[0];
            uint32_t fi = s_arg[0];
            for (int w = 1; w < 8; ++w)
                if (s_best[w] < fb || (s_best[w] == fb && s_arg[w] < fi)) { fb = s_best[w]; fi = s_arg[w]; }
            // every distance +inf / NaN: the reference keeps best = 0
            codes[j * m + mi] = (uint8_t)(fb < INFINITY ? fi : 0u);
        }
        __syncthreads();
    }
}

// simd::dot of the AVX2 tier (avx2.cpp:63-84) for n >= 16, the scalar loop of pq.cpp:42-48 below that
__device__ __forceinline__ float pq_dot(const float* __restrict__ a, const float* __restrict__ b, uint32_t n) {
    if (n < 16) {
        float acc = 0.f;
        for (uint32_t i = 0; i < n; ++i) acc = __fadd_rn(acc, __fmul_rn(a[i], b[i]));
        return acc;
    }
    float a0[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, a1[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    uint32_t i = 0;
    for (; i + 16 <= n; i += 16) {
#pragma unroll
        for (int l = 0; l < 8; ++l) {
            a0[l] = __fmaf_rn(a[i + l], b[i + l], a0[l]);
            a1[l] = __fmaf_rn(a[i + 8 + l], b[i + 8 + l], a1[l]);
        }
    }
    float s = __fadd_rn(a0[0], a1[0]);
#pragma unroll
    for (int l = 1; l < 8; ++l) s = __fadd_rn(s, __fadd_rn(a0[l], a1[l]));
    for (; i < n; ++i) s = __fmaf_rn(a[i], b[i], s);
    return s;
}

// one CTA per query: normalised query + LUT[mi][ki] = dot(q_sub, centroid) (PQInnerProductQuery, pq.cpp:391-403);
// valid[q] = 0 when the query cannot be normalised (the reference returns no results)
__global__ void __launch_bounds__(256) pq_lut_kernel(const float* __restrict__ q32, uint32_t d, const float* __restrict__ codebooks, uint32_t m,
                                                     uint32_t k, float* __restrict__ lut, uint32_t* __restrict__ valid) {
    extern __shared__ float qs[];
    __shared__ double s_norm;
    const uint32_t q = blockIdx.x;
    const uint32_t dsub = d / m;
    for (uint32_t c = threadIdx.x; c < d; c += blockDim.x) qs[c] = q32[(size_t)q * d + c];
    __syncthreads();
    if (threadIdx.x == 0) {
        double ss = 0.0;
        for (uint32_t c = 0; c < d; ++c) ss += (double)qs[c] * (double)qs[c];
        s_norm = ss;
    }
    __syncthreads();
    float inv = 0.f;
    const bool ok = normalize_factor(s_norm, &inv);
    if (threadIdx.x == 0) valid[q] = ok ? 1u : 0u;
    for (uint32_t c = threadIdx.x; c < d; c += blockDim.x) qs[c] = __fmul_rn(qs[c], inv);
    __syncthreads();
    for (uint32_t e = threadIdx.x; e < m * k; e += blockDim.x) {
        const uint32_t mi = e / k;
        lut[(size_t)q * m * k + e] = pq_dot(qs + mi * dsub, codebooks + (size_t)e * dsub, dsub);
    }
}

// ADC scan: CTA (tile, q).  score = sequential float sum of LUT[mi][code[mi]] (inner_product_from_lut, pq.cpp:50-57);
// key = (score, tie rank) packed so that a descending sort is (score desc, tie-break key asc); the tile's best `approx`
// keys are written to the query's list.
__global__ void __launch_bounds__(PQ_THREADS) pq_adc_kernel(const uint8_t* __restrict__ codes, uint64_t n_idx, uint32_t m, uint32_t k,
                                                            const float* __restrict__ lut_all, const uint32_t* __restrict__ tie_rank,
                                                            uint32_t approx, Cand* __restrict__ lists, uint32_t list_cap,
                                                            uint32_t* __restrict__ counts, uint32_t tile_stride) {
    extern __shared__ unsigned char sm[];
    float* lut = reinterpret_cast<float*>(sm);                                     // m * k floats
    uint64_t* keys = reinterpret_cast<uint64_t*>(sm + (((size_t)m * k * 4 + 7) & ~(size_t)7));   // PQ_TILE keys
    const uint32_t q = blockIdx.y;
    const uint64_t r0 = (uint64_t)blockIdx.x * tile_stride * PQ_TILE;
    const float* lq = lut_all + (size_t)q * m * k;
    for (uint32_t e = threadIdx.x; e < m * k; e += PQ_THREADS) lut[e] = lq[e];
    __syncthreads();
    const bool vec = (m % 16) == 0;
    for (uint32_t t = threadIdx.x; t < PQ_TILE; t += PQ_THREADS) {
        const uint64_t j = r0 + t;
        uint64_t key = 0;
        if (j < n_idx) {
            const uint8_t* cd = codes + j * m;
            float acc = 0.f;
            if (vec) {
                for (uint32_t mi = 0; mi < m; mi += 16) {
                    const uint4 w = *reinterpret_cast<const uint4*>(cd + mi);
                    const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                    for (int u = 0; u < 4; ++u)
#pragma unroll
                        for (int b = 0; b < 4; ++b) acc = __fadd_rn(acc, lut[(mi + u * 4 + b) * k + ((ws[u] >> (8 * b)) & 0xFFu)]);
                }
            } else {
                for (uint32_t mi = 0; mi < m; ++mi) acc = __fadd_rn(acc, lut[mi * k + cd[mi]]);
            }
            uint32_t b = __float_as_uint(acc);
            b = (acc != acc) ? 0u : ((b & 0x80000000u) ? ~b : (b | 0x80000000u));   // monotone key, NaN last (as knn.cu fkey)
            key = ((uint64_t)b << 32) | (uint64_t)(0xFFFFFFFFu - tie_rank[j]);
            if (key == 0) key = 1;
        }
        keys[t] = key;
    }
    __syncthreads();
    // bitonic sort, descending
    for (uint32_t size = 2; size <= PQ_TILE; size <<= 1) {
        for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
            for (uint32_t t = threadIdx.x; t < PQ_TILE / 2; t += PQ_THREADS) {
                const uint32_t lo = (t / stride) * (stride * 2) + (t % stride), hi = lo + stride;
                const bool desc = ((lo & size) == 0);
                const uint64_t x = keys[lo], y = keys[hi];
                if ((x < y) == desc) { keys[lo] = y; keys[hi] = x; }
            }
            __syncthreads();
        }
    }
    const uint64_t in_tile = n_idx > r0 ? min((uint64_t)PQ_TILE, n_idx - r0) : 0;
    const uint32_t nout = (uint32_t)min((uint64_t)approx, in_tile);
    __shared__ uint32_t s_base;
    if (threadIdx.x == 0) s_base = atomicAdd(&counts[q], nout);
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < nout; i += PQ_THREADS) {
        const uint64_t key = keys[i];
        uint32_t b = (uint32_t)(key >> 32);
        b = (b & 0x80000000u) ? (b & 0x7FFFFFFFu) : ~b;
        Cand c;
        c.score = __uint_as_float(b);
        c.row = 0xFFFFFFFFu - (uint32_t)key;     // tie rank: topk_select orders equal scores by ascending `row`
        if (s_base + i < list_cap) lists[(size_t)q * list_cap + s_base + i] = c;
    }
}

// Threshold of the filtered scan: the approx-th best key of the sampled tiles (any lower bound of the global approx-th best
// key keeps the selection exact); 0 = accept every row (fewer than approx sampled rows).
__global__ void pq_tau_kernel(const Cand* __restrict__ sel, const uint32_t* __restrict__ sel_n, uint32_t approx, uint32_t nq,
                              uint64_t* __restrict__ tau, uint32_t* __restrict__ counts) {
    uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    uint64_t t = 0;
    if (sel_n[q] >= approx) {
        const Cand c = sel[(size_t)q * approx + approx - 1];
        uint32_t b = __float_as_uint(c.score);
        b = (c.score != c.score) ? 0u : ((b & 0x80000000u) ? ~b : (b | 0x80000000u));
        t = ((uint64_t)b << 32) | (uint64_t)(0xFFFFFFFFu - c.row);
    }
    tau[q] = t;
    counts[q] = 0;
}

// Filtered ADC scan: CTA (row range, group of QG queries) with the QG lookup tables in shared memory; every row's code bytes
// are read once per group; rows whose key reaches the query's threshold are appended to its list.  Same float order as
// pq_adc_kernel (sequential sum over the m sub-quantisers).
template <int QG>
__global__ void __launch_bounds__(PQ_THREADS) pq_adc_filter_kernel(const uint8_t* __restrict__ codes, uint64_t n_idx, uint32_t m, uint32_t k,
                                                                   const float* __restrict__ lut_all, const uint32_t* __restrict__ tie_rank,
                                                                   const uint64_t* __restrict__ tau, uint32_t nq, uint64_t rows_per_cta,
                                                                   Cand* __restrict__ lists, uint32_t list_cap, uint32_t* __restrict__ counts) {
    extern __shared__ unsigned char sm[];
    float* lut = reinterpret_cast<float*>(sm);   // QG x m x k
    const uint32_t q0 = blockIdx.y * QG;
    const uint32_t lut_len = m * k;
    for (uint32_t e = threadIdx.x; e < lut_len * QG; e += PQ_THREADS) {
        const uint32_t g = e / lut_len;
        lut[e] = (q0 + g < nq) ? lut_all[(size_t)(q0 + g) * lut_len + (e - g * lut_len)] : 0.f;
    }
    uint64_t tq[QG];
#pragma unroll
    for (int g = 0; g < QG; ++g) tq[g] = (q0 + g < nq) ? tau[q0 + g] : ~0ull;
    __syncthreads();
    const uint64_t r0 = (uint64_t)blockIdx.x * rows_per_cta;
    const uint64_t r1 = min(n_idx, r0 + rows_per_cta);
    const bool vec = (m % 16) == 0;
    const unsigned lane = threadIdx.x & 31u;
    for (uint64_t base = r0; base < r1; base += PQ_THREADS) {     // block-uniform trip count (warp collectives below)
        const uint64_t j = base + threadIdx.x;
        const bool valid = j < r1;
        float acc[QG];
#pragma unroll
        for (int g = 0; g < QG; ++g) acc[g] = 0.f;
        if (valid) {
            const uint8_t* cd = codes + j * m;
            if (vec) {
                for (uint32_t mi = 0; mi < m; mi += 16) {
                    const uint4 w = *reinterpret_cast<const uint4*>(cd + mi);
                    const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                    for (int u = 0; u < 4; ++u)
#pragma unroll
                        for (int b = 0; b < 4; ++b) {
                            const uint32_t e = (mi + u * 4 + b) * k + ((ws[u] >> (8 * b)) & 0xFFu);
#pragma unroll
                            for (int g = 0; g < QG; ++g) acc[g] = __fadd_rn(acc[g], lut[g * lut_len + e]);
                        }
                }
            } else {
                for (uint32_t mi = 0; mi < m; ++mi) {
                    const uint32_t e = mi * k + cd[mi];
#pragma unroll
                    for (int g = 0; g < QG; ++g) acc[g] = __fadd_rn(acc[g], lut[g * lut_len + e]);
                }
            }
        }
#pragma unroll
        for (int g = 0; g < QG; ++g) {
            uint32_t b = __float_as_uint(acc[g]);
            b = (acc[g] != acc[g]) ? 0u : ((b & 0x80000000u) ? ~b : (b | 0x80000000u));
            const uint32_t th = (uint32_t)(tq[g] >> 32);
            bool pass = valid && b >= th;           // tq = ~0 for the padding queries of the last group: never passes
            uint32_t rank = 0;
            if (pass) {
                rank = tie_rank[j];
                uint64_t key = ((uint64_t)b << 32) | (uint64_t)(0xFFFFFFFFu - rank);
                if (key == 0) key = 1;
                pass = key >= tq[g] && tq[g] != ~0ull;
            }
            const unsigned mask = __ballot_sync(0xffffffffu, pass);
            if (mask) {
                const int leader = __ffs(mask) - 1;
                uint32_t pos = 0;
                if ((int)lane == leader) pos = atomicAdd(&counts[q0 + g], (uint32_t)__popc(mask));
                pos = __shfl_sync(0xffffffffu, pos, leader) + __popc(mask & ((1u << lane) - 1u));
                if (pass && pos < list_cap) {
                    Cand c;
                    c.score = acc[g];
                    c.row = rank;
                    lists[(size_t)(q0 + g) * list_cap + pos] = c;
                }
            }
        }
    }
}

// counts[q] > cap: the filtered list dropped rows (a sample that misjudged the data): the caller re-runs the unfiltered scan
__global__ void pq_overflow_kernel(const uint32_t* __restrict__ counts, uint32_t nq, uint32_t cap, uint32_t* __restrict__ flag) {
    uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q < nq && counts[q] > cap) atomicExch(flag, 1u);
}

// survivors carry tie ranks: back to corpus rows (rank -> indexed row -> corpus row)
__global__ void pq_rank_to_row_kernel(Cand* __restrict__ sel, const uint32_t* __restrict__ sel_n, uint32_t K, uint32_t nq,
                                      const uint32_t* __restrict__ rank_to_idx, const uint32_t* __restrict__ idx_rows) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nq * K) return;
    uint32_t q = t / K, j = t % K;
    if (j < sel_n[q]) {
        Cand c = sel[t];
        c.row = idx_rows[rank_to_idx[c.row]];
        sel[t] = c;
    }
}

struct PqExact {
    float sim;
    uint32_t row;
};
// VectorDatabase::computeCosineSimilarity (vector_database.cpp:1786-1810) of the ORIGINAL query against the stored row, float
// cast, `similarity < threshold` dropped (sqlite_vec_backend.cpp:4022-4037)
__global__ void pq_rerank_kernel(const void* __restrict__ rows, int dtype, uint32_t d, const float* __restrict__ q32, const Cand* __restrict__ sel,
                                 const uint32_t* __restrict__ sel_n, const uint32_t* __restrict__ valid, uint32_t K, uint32_t nq,
                                 float threshold, PqExact* __restrict__ out) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nq * K) return;
    uint32_t q = t / K, j = t % K;
    PqExact e;
    e.sim = __int_as_float(0x7FC00000);
    e.row = 0xFFFFFFFFu;
    if (valid[q] && j < sel_n[q]) {
        const uint32_t row = sel[t].row;
        const float* a = q32 + (size_t)q * d;
        double dp = 0.0, na = 0.0, nb = 0.0;
        for (uint32_t c = 0; c < d; ++c) {
            const double x = (double)a[c], y = (double)load_row_elem(rows, dtype, (uint64_t)row * d + c);
            dp += x * y;
            na += x * x;
            nb += y * y;
        }
        na = sqrt(na);
        nb = sqrt(nb);
        const float sim = (float)((na == 0.0 || nb == 0.0) ? 0.0 : dp / (na * nb));
        if (sim == sim && !(sim < threshold)) {
            e.sim = sim;
            e.row = row;
        }
    }
    out[t] = e;
}

// defined in knn.cu
void launch_topk_lists(const Cand* lists, const uint32_t* counts, uint32_t cap, uint32_t nq, uint32_t K, Cand* out_sel, uint32_t* out_n,
                       cudaStream_t st);
void launch_final_plain(const void* exact, uint32_t Kp, uint32_t k, uint32_t nq, const int64_t* rowids, int64_t* out_rowids, float* out_scores,
                        uint32_t* out_counts, uint64_t* out_flags, cudaStream_t st);

}  // namespace yb

using namespace yb;

struct yams_b200_pq {
    yams_b200_corpus* corpus = nullptr;
    uint64_t corpus_generation = 0;
    uint32_t m = 0, k = 0, dim = 0;
    uint64_t n_idx = 0;
    DevBuf codebooks, codes, idx_rows, tie_rank, rank_to_idx;
    // search workspace
    DevBuf q32, lut, valid, lists, counts, sel, exact, dout, tau;
    std::mutex mu;
};

extern "C" {

yams_status_t yams_b200_pq_build(yams_b200_corpus* c, uint32_t m, uint32_t k, const float* codebooks, const uint64_t* tie_break_keys,
                                 yams_b200_pq** out) {
    YB_TRY
    YB_ARG(c && out && codebooks, "null argument");
    *out = nullptr;
    std::lock_guard<std::mutex> corpus_lock(c->mu);
    YB_BIND(c);
    YB_ARG(m >= 1 && k >= 2 && k <= 256, "m must be >= 1 and k in 2..256");            // sqlite_vec_backend.cpp:3552-3553
    YB_ARG(c->dim % m == 0, "m must divide the dimension");                            // :3549-3551
    YB_ARG((size_t)m * k * 4 + PQ_TILE * 8 + 64 <= 200 * 1024, "m * k too large for the shared-memory lookup table");
    const uint64_t n = c->n;
    YB_ARG(n < 0xFFFFFFFFull, "too many rows");
    yams_b200_pq* p = new (std::nothrow) yams_b200_pq();
    if (!p) return YAMS_ERR_INTERNAL;
    struct Guard { yams_b200_pq* p; ~Guard() { if (p) yams_b200_pq_destroy(p); } } guard{p};
    p->corpus = c;
    p->corpus_generation = c->generation;
    p->m = m; p->k = k; p->dim = c->dim;
    cudaStream_t st = c->st;
    yams_status_t rc;
    const uint32_t dsub = c->dim / m;
    const size_t cb_bytes = (size_t)m * k * dsub * 4;
    if ((rc = p->codebooks.reserve(cb_bytes)) != YAMS_OK) return rc;
    YB_CUDA(cudaMemcpyAsync(p->codebooks.p, codebooks, cb_bytes, cudaMemcpyHostToDevice, st));
    if (n) {
        // rows the index holds: those normalizeEmbeddingInPlace accepts (:3569-3577)
        DevBuf keep, pos, scratch, total;
        struct Rel { DevBuf *a, *b, *c2, *d; ~Rel() { a->release(); b->release(); c2->release(); d->release(); } } rel{&keep, &pos, &scratch, &total};
        if ((rc = keep.reserve(n * 4)) != YAMS_OK || (rc = pos.reserve(n * 4)) != YAMS_OK || (rc = total.reserve(8)) != YAMS_OK) return rc;
        pq_keep_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(c->rows.p, c->dtype, c->dim, n, keep.as<uint32_t>());
        if ((rc = exclusive_scan_u32(keep.as<uint32_t>(), pos.as<uint32_t>(), n, total.as<uint64_t>(), scratch, st)) != YAMS_OK) return rc;
        uint64_t n_idx = 0;
        YB_CUDA(cudaMemcpyAsync(&n_idx, total.p, 8, cudaMemcpyDeviceToHost, st));
        YB_CUDA(cudaStreamSynchronize(st));
        p->n_idx = n_idx;
        if ((rc = p->idx_rows.reserve(std::max<uint64_t>(n_idx, 1) * 4)) != YAMS_OK) return rc;
        pq_compact_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(keep.as<uint32_t>(), pos.as<uint32_t>(), n, p->idx_rows.as<uint32_t>());
        if ((rc = p->codes.reserve(std::max<uint64_t>(n_idx, 1) * m + 64)) != YAMS_OK) return rc;
        if (n_idx) {
            YB_CUDA(cudaFuncSetAttribute(pq_encode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(c->dim * 4)));
            pq_encode_kernel<<<(unsigned)n_idx, 256, c->dim * 4, st>>>(c->rows.p, c->dtype, c->dim, p->idx_rows.as<uint32_t>(), n_idx,
                                                                      p->codebooks.as<float>(), m, k, p->codes.as<uint8_t>());
        }
        YB_CUDA(cudaGetLastError());
        // tie-break order (sqlite_vec_backend.cpp:3985-3990): rank of every indexed row by (tie_break_key, position)
        std::vector<uint32_t> h_idx(n_idx), rank(n_idx), order(n_idx);
        if (n_idx) YB_CUDA(cudaMemcpyAsync(h_idx.data(), p->idx_rows.p, n_idx * 4, cudaMemcpyDeviceToHost, st));
        YB_CUDA(cudaStreamSynchronize(st));
        for (uint64_t j = 0; j < n_idx; ++j) order[j] = (uint32_t)j;
        if (tie_break_keys)
            std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return tie_break_keys[h_idx[a]] < tie_break_keys[h_idx[b]]; });
        for (uint64_t r = 0; r < n_idx; ++r) rank[order[r]] = (uint32_t)r;
        if ((rc = p->tie_rank.reserve(std::max<uint64_t>(n_idx, 1) * 4)) != YAMS_OK) return rc;
        if ((rc = p->rank_to_idx.reserve(std::max<uint64_t>(n_idx, 1) * 4)) != YAMS_OK) return rc;
        if (n_idx) {
            YB_CUDA(cudaMemcpyAsync(p->tie_rank.p, rank.data(), n_idx * 4, cudaMemcpyHostToDevice, st));
            YB_CUDA(cudaMemcpyAsync(p->rank_to_idx.p, order.data(), n_idx * 4, cudaMemcpyHostToDevice, st));
        }
        YB_CUDA(cudaStreamSynchronize(st));
    }
    guard.p = nullptr;
    *out = p;
    return YAMS_OK;
    YB_CATCH
}

void yams_b200_pq_destroy(yams_b200_pq* p) {
    if (!p) return;
    for (DevBuf* b : {&p->codebooks, &p->codes, &p->idx_rows, &p->tie_rank, &p->rank_to_idx, &p->q32, &p->lut, &p->valid, &p->lists, &p->counts,
                      &p->sel, &p->exact, &p->dout, &p->tau})
        b->release();
    delete p;
}

yams_status_t yams_b200_pq_codes(yams_b200_pq* p, uint8_t* out_codes, int64_t* out_rowids, uint64_t* out_n) {
    YB_TRY
    YB_ARG(p && out_n, "null argument");
    std::lock_guard<std::mutex> lk(p->mu);
    std::lock_guard<std::mutex> corpus_lock(p->corpus->mu);
    YB_BIND(p->corpus);
    *out_n = p->n_idx;
    cudaStream_t st = p->corpus->st;
    if (p->n_idx && out_codes) YB_CUDA(cudaMemcpyAsync(out_codes, p->codes.p, p->n_idx * p->m, cudaMemcpyDeviceToHost, st));
    if (p->n_idx && out_rowids) {
        std::vector<uint32_t> h_idx(p->n_idx);
        std::vector<int64_t> h_rid(p->corpus->n);
        YB_CUDA(cudaMemcpyAsync(h_idx.data(), p->idx_rows.p, p->n_idx * 4, cudaMemcpyDeviceToHost, st));
        YB_CUDA(cudaMemcpyAsync(h_rid.data(), p->corpus->rowids.p, p->corpus->n * 8, cudaMemcpyDeviceToHost, st));
        YB_CUDA(cudaStreamSynchronize(st));
        for (uint64_t j = 0; j < p->n_idx; ++j) out_rowids[j] = h_rid[h_idx[j]];
    }
    YB_CUDA(cudaStreamSynchronize(st));
    return YAMS_OK;
    YB_CATCH
}

}  // extern "C"

// the per-query lists (32 k entries each on the filtered path) bound the batch one pass can take: larger batches are sliced
static constexpr uint32_t kPqMaxBatch = 512;
static yams_status_t pq_search_batch(yams_b200_pq* p, const float* queries, uint32_t nq, uint32_t k, uint32_t rerank_factor, float threshold,
                                     int64_t* out_rowids, float* out_scores, uint32_t* out_counts, uint64_t* out_flags);

extern "C" {

yams_status_t yams_b200_pq_search(yams_b200_pq* p, const float* queries, uint32_t nq, uint32_t k, uint32_t rerank_factor, float threshold,
                                  int64_t* out_rowids, float* out_scores, uint32_t* out_counts, uint64_t* out_flags) {
    YB_TRY
    YB_ARG(p && out_counts, "null argument");
    std::lock_guard<std::mutex> lk(p->mu);
    yams_b200_corpus* c = p->corpus;
    std::lock_guard<std::mutex> corpus_lock(c->mu);
    YB_BIND(c);
    for (uint32_t q = 0; q < nq; ++q) out_counts[q] = 0;
    if (out_flags) for (uint32_t q = 0; q < nq; ++q) out_flags[q] = 0;
    if (nq == 0 || k == 0 || p->n_idx == 0) return YAMS_OK;                           // :3873-3881
    YB_ARG(queries && out_rowids && out_scores, "null argument");
    YB_ARG(c->generation == p->corpus_generation, "the corpus changed after the PQ index was built: rebuild it (the reference marks the index dirty)");
    for (uint32_t q0 = 0; q0 < nq; q0 += kPqMaxBatch) {
        const uint32_t nb = std::min<uint32_t>(kPqMaxBatch, nq - q0);
        yams_status_t rc = pq_search_batch(p, queries + (size_t)q0 * p->dim, nb, k, rerank_factor, threshold, out_rowids + (size_t)q0 * k,
                                           out_scores + (size_t)q0 * k, out_counts + q0, out_flags ? out_flags + q0 : nullptr);
        if (rc != YAMS_OK) return rc;
    }
    return YAMS_OK;
    YB_CATCH
}

}  // extern "C"

// one batch of <= kPqMaxBatch queries; the caller holds both locks and has bound the device
static yams_status_t pq_search_batch(yams_b200_pq* p, const float* queries, uint32_t nq, uint32_t k, uint32_t rerank_factor, float threshold,
                                     int64_t* out_rowids, float* out_scores, uint32_t* out_counts, uint64_t* out_flags) {
    YB_TRY
    yams_b200_corpus* c = p->corpus;
    if (rerank_factor == 0) rerank_factor = 1;                                        // :3670 max(1, rerank_factor)
    const uint64_t budget = (uint64_t)k * rerank_factor;
    const uint32_t approx = (uint32_t)std::min<uint64_t>(p->n_idx, std::max<uint64_t>(k, budget));   // :3956-3959
    YB_ARG(approx <= PQ_MAX_APPROX, "k * rerank_factor > 1024 is not supported by the PQ engine (use the exact scan)");
    cudaStream_t st = c->st;
    yams_status_t rc;
    const uint32_t m = p->m, kc = p->k, d = p->dim;
    const uint32_t tiles = (uint32_t)((p->n_idx + PQ_TILE - 1) / PQ_TILE);
    // Two ways to the same exact top-approx by (score, tie-break key):
    //   unfiltered: every tile keeps its own top-approx (bitonic network in shared memory) -> lists of tiles * approx entries;
    //   filtered  : a strided sample of tiles goes through the unfiltered path and yields the sample's approx-th best key tau
    //               (<= the global approx-th best key), then ONE pass over all rows appends the rows with key >= tau.
    // The filtered pass does no sorting (the unfiltered kernel spends ~10x its lookup work in the network); it is used when
    // the expected survivors (approx * tiles / sample tiles) fit the list with a 4x margin, and falls back when a list overflows.
    static const int force_unfiltered = [] { const char* e = getenv("YAMS_B200_PQ_UNFILTERED"); return e ? atoi(e) : 0; }();
    const uint32_t kFilterCap = 32768;
    uint32_t sample_tiles = 16;
    while ((uint64_t)approx * 4 * tiles > (uint64_t)kFilterCap * sample_tiles && sample_tiles < tiles) sample_tiles *= 2;
    const bool filtered = !force_unfiltered && tiles >= 4 * sample_tiles;
    const uint32_t cap_unf = tiles * approx;
    const uint32_t cap = filtered ? std::max<uint32_t>(kFilterCap, sample_tiles * approx) : cap_unf;
    if ((rc = p->q32.reserve((size_t)nq * d * 4)) != YAMS_OK) return rc;
    if ((rc = p->lut.reserve((size_t)nq * m * kc * 4)) != YAMS_OK) return rc;
    if ((rc = p->valid.reserve((size_t)nq * 4)) != YAMS_OK) return rc;
    if ((rc = p->lists.reserve((size_t)nq * cap * sizeof(Cand))) != YAMS_OK) return rc;
    if ((rc = p->counts.reserve((size_t)nq * 4 + 16)) != YAMS_OK) return rc;
    if ((rc = p->sel.reserve((size_t)nq * approx * sizeof(Cand) + (size_t)nq * 4)) != YAMS_OK) return rc;
    if ((rc = p->exact.reserve((size_t)nq * approx * sizeof(PqExact))) != YAMS_OK) return rc;
    if ((rc = p->dout.reserve((size_t)nq * k * 12 + (size_t)nq * 12 + 64)) != YAMS_OK) return rc;
    if ((rc = p->tau.reserve((size_t)nq * 8)) != YAMS_OK) return rc;
    Cand* d_sel = p->sel.as<Cand>();
    uint32_t* d_sel_n = reinterpret_cast<uint32_t*>(d_sel + (size_t)nq * approx);
    uint32_t* d_counts = p->counts.as<uint32_t>();
    uint32_t* d_overflow = d_counts + nq;
    YB_CUDA(cudaMemcpyAsync(p->q32.p, queries, (size_t)nq * d * 4, cudaMemcpyHostToDevice, st));
    YB_CUDA(cudaFuncSetAttribute(pq_lut_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(d * 4)));
    pq_lut_kernel<<<nq, 256, d * 4, st>>>(p->q32.as<float>(), d, p->codebooks.as<float>(), m, kc, p->lut.as<float>(), p->valid.as<uint32_t>());
    const size_t lut_bytes = (size_t)m * kc * 4;
    const size_t smem = ((lut_bytes + 7) & ~(size_t)7) + (size_t)PQ_TILE * 8;
    YB_CUDA(cudaFuncSetAttribute(pq_adc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    auto unfiltered = [&](uint32_t ntiles, uint32_t stride, uint32_t list_cap) {
        cudaMemsetAsync(d_counts, 0, (size_t)nq * 4 + 16, st);
        pq_adc_kernel<<<dim3(ntiles, nq), PQ_THREADS, smem, st>>>(p->codes.as<uint8_t>(), p->n_idx, m, kc, p->lut.as<float>(), p->tie_rank.as<uint32_t>(),
                                                                  approx, p->lists.as<Cand>(), list_cap, d_counts, stride);
        launch_topk_lists(p->lists.as<Cand>(), d_counts, list_cap, nq, approx, d_sel, d_sel_n, st);
    };
    bool used_filter = false;
    if (filtered) {
        used_filter = true;
        const uint32_t stride = tiles / sample_tiles;
        unfiltered(sample_tiles, stride, cap);
        pq_tau_kernel<<<(nq + 127) / 128, 128, 0, st>>>(d_sel, d_sel_n, approx, nq, p->tau.as<uint64_t>(), d_counts);
        int qg = nq >= 4 ? 4 : (nq >= 2 ? 2 : 1);
        while (qg > 1 && (size_t)qg * lut_bytes > 200 * 1024) qg >>= 1;
        const uint32_t groups = (nq + qg - 1) / qg;
        uint64_t chunks = std::max<uint64_t>(1, (uint64_t)c->dev->sm_count * 6 / groups);
        chunks = std::min<uint64_t>(chunks, (p->n_idx + 2047) / 2048);
        uint64_t rows_per_cta = ((p->n_idx + chunks - 1) / chunks + PQ_THREADS - 1) / PQ_THREADS * PQ_THREADS;
        chunks = (p->n_idx + rows_per_cta - 1) / rows_per_cta;
#define YB_PQF(G)                                                                                                                              \
        YB_CUDA(cudaFuncSetAttribute(pq_adc_filter_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(G * lut_bytes)));             \
        pq_adc_filter_kernel<G><<<dim3((unsigned)chunks, groups), PQ_THREADS, G * lut_bytes, st>>>(                                            \
            p->codes.as<uint8_t>(), p->n_idx, m, kc, p->lut.as<float>(), p->tie_rank.as<uint32_t>(), p->tau.as<uint64_t>(), nq, rows_per_cta,   \
            p->lists.as<Cand>(), cap, d_counts)
        if (qg == 4) { YB_PQF(4); } else if (qg == 2) { YB_PQF(2); } else { YB_PQF(1); }
#undef YB_PQF
        pq_overflow_kernel<<<(nq + 127) / 128, 128, 0, st>>>(d_counts, nq, cap, d_overflow);
        launch_topk_lists(p->lists.as<Cand>(), d_counts, cap, nq, approx, d_sel, d_sel_n, st);
    } else {
        unfiltered(tiles, 1, cap);
    }
    for (int attempt = 0; attempt < 2; ++attempt) {
    if (attempt == 1) {
        // a filtered list overflowed: redo the selection without the filter
        if ((rc = p->lists.reserve((size_t)nq * cap_unf * sizeof(Cand))) != YAMS_OK) return rc;
        unfiltered(tiles, 1, cap_unf);
    }
    const uint32_t tot = nq * approx;
    pq_rank_to_row_kernel<<<(tot + 255) / 256, 256, 0, st>>>(d_sel, d_sel_n, approx, nq, p->rank_to_idx.as<uint32_t>(), p->idx_rows.as<uint32_t>());
    pq_rerank_kernel<<<(tot + 127) / 128, 128, 0, st>>>(c->rows.p, c->dtype, d, p->q32.as<float>(), d_sel, d_sel_n, p->valid.as<uint32_t>(), approx, nq,
                                                        threshold, p->exact.as<PqExact>());
    int64_t* d_or = p->dout.as<int64_t>();
    uint64_t* d_of = reinterpret_cast<uint64_t*>(d_or + (size_t)nq * k);
    float* d_os = reinterpret_cast<float*>(d_of + nq);
    uint32_t* d_oc = reinterpret_cast<uint32_t*>(d_os + (size_t)nq * k);
    launch_final_plain(p->exact.p, approx, k, nq, c->rowids.as<int64_t>(), d_or, d_os, d_oc, d_of, st);
    YB_CUDA(cudaGetLastError());
    YB_CUDA(cudaMemcpyAsync(out_rowids, d_or, (size_t)nq * k * 8, cudaMemcpyDeviceToHost, st));
    YB_CUDA(cudaMemcpyAsync(out_scores, d_os, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, st));
    YB_CUDA(cudaMemcpyAsync(out_counts, d_oc, (size_t)nq * 4, cudaMemcpyDeviceToHost, st));
    if (out_flags) YB_CUDA(cudaMemcpyAsync(out_flags, d_of, (size_t)nq * 8, cudaMemcpyDeviceToHost, st));
    uint32_t h_overflow = 0;
    if (used_filter && attempt == 0) YB_CUDA(cudaMemcpyAsync(&h_overflow, d_overflow, 4, cudaMemcpyDeviceToHost, st));
    YB_CUDA(cudaStreamSynchronize(st));
    if (!h_overflow) break;
    }
    return YAMS_OK;
    YB_CATCH
}
