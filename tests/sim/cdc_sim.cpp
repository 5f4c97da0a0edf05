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

// Emulates chunk_and_hash_batch's device pass (ingest.cu run_batch_group): the files are laid out in one zeroed buffer
// (256-byte aligned starts, >= 64 zero bytes in front of each), ONE candidate pass over the whole buffer, the merged
// node table, the shared chain kernels, per-node emission.  out_first[f] = index of file f's first chunk.
size_t sim_chunk_batch(const uint8_t* const* files, const size_t* lens, size_t nfiles, uint64_t window, uint64_t minc,
                       uint64_t maxc, uint64_t poly, uint64_t mask, int variant, uint64_t* out_offsets, uint64_t* out_sizes,
                       size_t cap, uint64_t* out_first) {
    Params pr;
    if (!resolve(window, minc, maxc, poly, mask, variant, &pr)) return (size_t)-1;
    const CdcParams& P = pr.P;
    const uint32_t nf = (uint32_t)nfiles;
    std::vector<uint64_t> starts(nf), ends(nf);
    uint64_t pos = 256;
    for (uint32_t f = 0; f < nf; ++f) {
        starts[f] = pos;
        ends[f] = pos + lens[f];
        pos += (lens[f] + 64 + 255) & ~255ull;
    }
    const uint64_t L = pos;
    std::vector<uint8_t> buf((size_t)L + 256, 0);
    for (uint32_t f = 0; f < nf; ++f)
        if (lens[f]) memcpy(buf.data() + starts[f], files[f], lens[f]);
    ByteView view{buf.data(), 0, 0};
    std::vector<uint64_t> cand;
    if (!pr.no_candidates)
        for (uint64_t p = 0; p < L; ++p)
            if (is_candidate(view, pr.T, P, p)) cand.push_back(p);
    const uint32_t ncand = (uint32_t)cand.size();
    uint64_t dummy = 0;
    BatchLayout B{cand.empty() ? &dummy : cand.data(), ncand, starts.data(), ends.data(), nf};
    const uint32_t nnodes = ncand + nf;
    std::vector<uint64_t> npos(nnodes);
    std::vector<uint32_t> nref(nnodes), root_node(nf), next(nnodes), forced(nnodes), cnt(nnodes), exit_(nnodes);
    for (uint32_t j = 0; j < ncand; ++j) {
        uint32_t node = batch_node_of_cand(B, j);
        npos[node] = cand[j] + 1;
        nref[node] = j;
    }
    for (uint32_t f = 0; f < nf; ++f) {
        uint32_t node = batch_node_of_root(B, f);
        npos[node] = starts[f];
        nref[node] = kBatchRootFlag | f;
        root_node[f] = node;
    }
    for (uint32_t i = 0; i < nnodes; ++i) {
        BatchNext r = batch_next(B, P, i, npos[i], nref[i], root_node.data(), nnodes);
        if (r.next <= i) return (size_t)-2;   // the chain must move forward
        next[i] = r.next;
        forced[i] = r.forced;
        cnt[i] = r.count;
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
    std::vector<uint64_t> offsets(nnodes + 1, 0);
    for (uint32_t i = 0; i < nnodes; ++i) offsets[i + 1] = offsets[i] + (onchain[i] ? cnt[i] : 0);
    const size_t total = (size_t)offsets[nnodes];
    for (uint32_t f = 0; f < nf; ++f) {
        if (!onchain[root_node[f]]) return (size_t)-3;   // every root lies on the chain
        out_first[f] = offsets[root_node[f]];
    }
    out_first[nf] = total;
    for (uint32_t i = 0; i < nnodes; ++i) {
        if (!onchain[i]) continue;
        uint64_t base = offsets[i];
        batch_emit(B, P, npos[i], forced[i], next[i] < nnodes ? npos[next[i]] : 0ull, [&](uint64_t k, uint64_t off, uint64_t size) {
            uint32_t fup = upper_bound_u64(starts.data(), nf, off);
            if (base + k < cap) {
                out_offsets[base + k] = off - starts[fup - 1];   // batch_rebase_kernel
                out_sizes[base + k] = size;
            }
        });
    }
    return total;
}
}
