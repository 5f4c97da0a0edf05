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
    int64_t last_rowid = INT64_MIN;
    bool rowids_dense = true;  // rowid[i] == rowid[0] + i
    DevBuf rows;       // n x dim elements
    DevBuf rowids;     // n x int64
    DevBuf inv_norm;   // n x float: 1/|row| (double math), 0 for rows the reference skips
    // search workspace
    DevBuf q32, q16, qinv, tau, counts, cands, sample_scores, sel, outbuf, dense, mask, misc, dout;
    HostBuf h_pin;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_scan[2] = {nullptr, nullptr};  // around the full-corpus filtered scan launch
    bool scan_timed = false;
    float last_ms[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    std::mutex mu;   // public entry points on one corpus are serialised (the workspace buffers are per corpus)
    size_t elem() const { return dtype == YAMS_B200_F16 ? 2 : 4; }
};

// Stage-1 engine interface: scores of `nrows` corpus rows (row_start + i*row_stride) against nq
// queries; cosine scaling (inv_norm[row] * qinv[q]) applied.
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
};

yams_status_t stage1_cuda_core(const Stage1Args& a, bool filter, cudaStream_t st);

}  // namespace yb
