// knn.cuh -- shared declarations of the exact vector scan (vector_scan_v1).
#pragma once
#include <mutex>

#include "common.cuh"

namespace yb {

// (score, row index) candidate produced by stage 1
struct __align__(8) Cand {
    float score;
    uint32_t row;
};

struct Corpus {
    DeviceCtx* dev = nullptr;
    cudaStream_t st = nullptr;
    uint32_t dim = 0;
    int dtype = YAMS_B200_F16;
    int metric = YAMS_B200_COSINE;
    uint64_t n = 0;
    uint64_t generation = 0;   // bumped by every append / remove / clear: derived indexes (pq.cu) notice a stale corpus
    int64_t last_rowid = INT64_MIN;
    bool rowids_dense = true;  // rowid[i] == rowid[0] + i
    DevBuf rows;       // n x dim elements
    DevBuf rowids;     // n x int64
    DevBuf inv_norm;   // n x float. cosine: 1/|row| (double math), 0 for rows the reference skips; L2: |row|^2
    DevBuf stats;      // 3 floats: max |row|, max |row - tf32(row)|, max of that residual relative to |row| (row_stats_kernel)
    float r_max = 0.f, dr_abs_max = 0.f, dr_rel_max = 0.f;   // host copies of stats (upper bounds; removals keep them)
    // search workspace
    DevBuf q32, q16, qinv, tau, counts, cands, sample_scores, sel, outbuf, dense, mask, misc, dout;
    DevBuf status;     // ScanStatus + bad[nq] + bad2[nq] of the call in flight
    DevBuf bound, eps, eps2;   // per query: certificate inputs (knn.cu final_kernel)
    DevBuf lvl;        // exhaustive level 1 workspace
    HostBuf h_pin;
    // the call enqueued by search_device and not yet finished (search_device_finish)
    struct Pending {
        bool active = false;
        uint32_t nq = 0, k = 0;
        float threshold = 0.f;
        int64_t* d_out_rowids = nullptr;
        float* d_out_scores = nullptr;
    } pending;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_scan[2] = {nullptr, nullptr};  // around the full-corpus filtered scan launch
    bool scan_timed = false;
    float last_ms[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    std::mutex mu;   // public entry points on one corpus are serialised (the workspace buffers are per corpus)
    size_t elem() const { return dtype == YAMS_B200_F16 ? 2 : 4; }
};

// Stage-1 engine interface: approximate scores of `nrows` corpus rows (row_start + i*row_stride) against nq queries.
//   cosine: score = q.r * inv_norm[row] * qinv[q]          (-inf for rows the reference skips)
//   L2    : score = 2 q.r - |row|^2  (inv_norm[] holds |row|^2; larger = closer; |q|^2 is constant per query)
//   STORE : out_scores[q * ld + i] = score
//   FILTER: rows with score > tau[q] (and mask bit set, if mask) are appended to cands[q][..cap)
struct Stage1Args {
    const void* rows;
    const float* inv_norm;
    uint32_t dim;
    int dtype;
    uint64_t row_start;
    uint64_t row_stride;
    uint64_t nrows;
    const float* q32;       // nq x dim fp32
    const float* qinv;      // nq: 1/|q|
    uint32_t nq;
    // STORE
    float* out_scores;
    uint64_t ld;
    // FILTER
    const float* tau;
    Cand* cands;
    uint32_t cap;
    uint32_t* counts;
    const uint32_t* mask;   // nullable: bit (q * mask_ld*32 + row)
    uint64_t mask_ld;       // words per query
    int metric;             // YAMS_B200_COSINE / YAMS_B200_L2
    // tensor engine: operand-typed queries + the per-query error bound eps[] (|score~ - score| <= eps[q], the input of the
    // exactness certificate) are prepared by the first call of a search and reused by the following ones
    bool skip_qprep;
    float* eps;             // nq floats, written by the engine's query preparation
    float r_max, dr_abs_max, dr_rel_max;   // corpus norm statistics (Corpus)
};

yams_status_t stage1_cuda_core(const Stage1Args& a, bool filter, cudaStream_t st);

}  // namespace yb

// the opaque handle of the C ABI
struct yams_b200_corpus : public yb::Corpus {};
