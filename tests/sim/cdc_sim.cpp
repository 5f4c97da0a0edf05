// cdc_sim.cpp -- TEST-ONLY host emulation of the GPU chunking pipeline.  It drives the very same
// integer logic the kernels use (yams_b200/csrc/cdc_logic.h: is_candidate, next_cut, block_exit_seq,
// block_mark_seq, node_emit_count) in the same order as cdc.cu / ingest.cu, but sequentially on the
// CPU, so the logic can be checked against the oracle on machines without a GPU.  Not part of the
// product; never linked into libyams_b200.so.
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../yams_b200/csrc/cdc_logic.h"

using namespace yb;

namespace {
struct Params {
    CdcParams P;
    uint64_t T[256];
    bool no_candidates;
};

bool resolve(uint64_t window, uint64_t minc, uint64_t maxc, uint64_t poly, uint64_t mask, int variant, Params* out) {
    // mirrors yb::resolve_params in cdc.cu
    memset(&out->P, 0, sizeof(out->P));
    if (window > (uint64_t)kMaxWindow) return false;
    uint64_t p = poly ? poly : kDefaultPoly;
    for (int b = 0; b < 256; ++b) out->T[b] = table_entry(p, (uint32_t)b);
    out->P.mask = mask;
    out->P.window = window ? (uint32_t)window : 1u;
    out->P.steps = mask_steps(mask);
    uint64_t force = minc > maxc ? minc : maxc;
    if (variant == 0) {
        out->P.lo = (minc > 1 ? minc : 1) - 1;
        if (force == 0) force = 1;
    } else {
        out->P.lo = minc;
        if (force == 0) return false;
    }
    out->P.force = force;
    out->no_candidates = out->P.lo >= out->P.force;
    return true;
}
}  // namespace

extern "C" {

// Emulates a streaming session fed in slices of `slice` bytes (0 = everything in one feed).
// Returns the number of chunks (may exceed cap), or (size_t)-1 on invalid config.
size_t sim_chunk(const uint8_t* data, size_t n, uint64_t window, uint64_t minc, uint64_t maxc, uint64_t poly,
                 uint64_t mask, int variant, size_t slice, uint64_t* out_offsets, uint64_t* out_sizes, size_t cap) {
    Params pr;
    if (!resolve(window, minc, maxc, poly, mask, variant, &pr)) return (size_t)-1;
    const CdcParams& P = pr.P;
    size_t nout = 0;
    uint64_t chunk_start = 0, stream_pos = 0, keep_from = 0;
    if (slice == 0) slice = n ? n : 1;
    bool done = false;
    while (!done) {
        size_t len = (size_t)((n - stream_pos) < slice ? (n - stream_pos) : slice);
        bool final = stream_pos + len == n;
        uint64_t scan_lo = stream_pos, scan_hi = stream_pos + len;
        // the device sees bytes from keep_from only
        ByteView view{data + keep_from, keep_from, keep_from};
        std::vector<uint64_t> cand;
        if (!pr.no_candidates)
            for (uint64_t p = scan_lo; p < scan_hi; ++p)
                if (is_candidate(view, pr.T, P, p)) cand.push_back(p);
        uint32_t ncand = (uint32_t)cand.size();
        uint32_t nnodes = ncand + 1, END = ncand + 1;
        std::vector<uint32_t> next(nnodes), forced(nnodes), exit_(nnodes);
        const uint64_t* cp = cand.empty() ? nullptr : cand.data();
        uint64_t dummy = 0;
        if (!cp) cp = &dummy;
        for (uint32_t node = 0; node < nnodes; ++node) {
            uint64_t s = node_start(cp, node, chunk_start);
            NextCut r = next_cut(cp, ncand, node, s, P);
            next[node] = r.j + 1;
            forced[node] = (uint32_t)r.forced;
        }
        uint32_t nblocks = (nnodes + kNodeBlock - 1) / kNodeBlock;
        for (uint32_t b = 0; b < nblocks; ++b) {
            uint32_t bs = b * kNodeBlock, be = bs + kNodeBlock < nnodes ? bs + kNodeBlock : nnodes;
            block_exit_seq(next.data() + bs, bs, be, exit_.data() + bs);
        }
        std::vector<uint32_t> entry(nblocks, kNoEntry);
        for (uint32_t cur = 0; cur < nnodes; cur = exit_[cur]) entry[cur / kNodeBlock] = cur;
        std::vector<uint8_t> onchain(nnodes, 0);
        for (uint32_t b = 0; b < nblocks; ++b) {
            uint32_t bs = b * kNodeBlock, be = bs + kNodeBlock < nnodes ? bs + kNodeBlock : nnodes;
            if (entry[b] != kNoEntry) block_mark_seq(next.data() + bs, bs, be, entry[b], onchain.data() + bs);
        }
        uint64_t pending_start = chunk_start;
        for (uint32_t node = 0; node < nnodes; ++node) {
            if (!onchain[node]) continue;
            uint64_t s = node_start(cp, node, chunk_start);
            uint64_t cnt = node_emit_count(next[node], forced[node], END, s, scan_hi, final, P);
            if (next[node] != END) {
                uint64_t F = forced[node];
                for (uint64_t t = 0; t < F; ++t) {
                    if (nout < cap) { out_offsets[nout] = s + t * P.force; out_sizes[nout] = P.force; }
                    ++nout;
                }
                uint64_t ls = s + F * P.force;
                if (nout < cap) { out_offsets[nout] = ls; out_sizes[nout] = cp[next[node] - 1] + 1 - ls; }
                ++nout;
            } else {
                for (uint64_t t = 0; t < cnt; ++t) {
                    uint64_t cs = s + t * P.force;
                    uint64_t sz = scan_hi - cs < P.force ? scan_hi - cs : P.force;
                    if (nout < cap) { out_offsets[nout] = cs; out_sizes[nout] = sz; }
                    ++nout;
                }
                uint64_t ns = s + cnt * P.force;
                // same expression as cdc_emit_kernel
                pending_start = ns < scan_hi ? ns : (final ? scan_hi : ns);
            }
        }
        chunk_start = pending_start;  // (the kernel writes scalars[0]; the host reads it after the pass)
        stream_pos = scan_hi;
        uint64_t nk = chunk_start < (stream_pos > (uint64_t)kHistory ? stream_pos - kHistory : 0)
                          ? chunk_start
                          : (stream_pos > (uint64_t)kHistory ? stream_pos - kHistory : 0);
        if (nk < keep_from) nk = keep_from;
        keep_from = nk;
        if (final) done = true;
    }
    return nout;
}

// candidate positions through the shared predicate (for direct comparison with the oracle)
size_t sim_candidates(const uint8_t* data, size_t n, uint64_t window, uint64_t poly, uint64_t mask, uint64_t* out,
                      size_t cap) {
    Params pr;
    if (!resolve(window, 1, 2, poly, mask, 0, &pr)) return (size_t)-1;
    ByteView view{data, 0, 0};
    size_t cnt = 0;
    for (uint64_t p = 0; p < n; ++p)
        if (is_candidate(view, pr.T, pr.P, p)) {
            if (cnt < cap) out[cnt] = p;
            ++cnt;
        }
    return cnt;
}
}
