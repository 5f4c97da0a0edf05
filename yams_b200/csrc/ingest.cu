// ingest.cu -- host orchestration of the content-ingest path: CDC boundaries + per-chunk SHA-256.
// Implements the content_ingest_v1 entry points of include/yams_b200.h.
//
// Reference call sites this replaces (paths under /root/reference):
//   StreamingChunker::chunkData       src/chunking/streaming_chunker.cpp:92-137
//   StreamingChunker::processStream   include/yams/chunking/streaming_chunker.h:78-121
//   RabinChunker::chunkDataImpl       src/chunking/rabin_chunker.cpp:120-152
//   ContentStore::store / storeBytes  src/api/content_store_impl.cpp:216-220, 510-545 (callers)
#include <stdlib.h>

#include <algorithm>
#include <chrono>
#include <mutex>
#include <thread>
#include <new>
#include <vector>

#include "cdc_kernels.cuh"

namespace yb {

constexpr uint32_t kTileBytesHost = kTileBytes;
// Device-resident data is processed in segments (two host round trips each to size the selection buffers).  16 GiB: per 64 GiB the
// candidate scan + selection cost 15.7 ms instead of 17.7 ms with 4 GiB segments (profiles/r2_sha_order.md); the candidate
// workspace grows to ~1 GiB (2 x 16 B per 4 KiB tile x 16-entry slack).
static uint64_t segment_bytes() {
    const char* e = getenv("YAMS_B200_SEGMENT_MIB");
    uint64_t mib = e ? strtoull(e, nullptr, 10) : 16384;
    if (mib < 1) mib = 1;
    if (mib > 16384) mib = 16384;
    return mib << 20;
}
static inline double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
constexpr uint64_t kFeedSlice = 256ull << 20;      // host feeds are staged 256 MiB at a time

// All device-side working state of one chunking stream.
struct CdcStream {
    DeviceCtx* dev = nullptr;
    cudaStream_t st = nullptr;
    CdcParams P{};
    bool no_candidates = false;  // lo >= force: no candidate can ever cut
    bool two_pass_only = false;  // YAMS_B200_TWO_PASS=1: skip the single-pass scan (diagnostics)
    uint32_t two_pass_fallbacks = 0;
    DevBuf table, tile_counts, tile_offsets, cand, cand_tmp, next, forced, exit_, entry, onchain, emit_counts,
        emit_offsets, scan_scratch, descs, scalars, sha_order;
    DevBuf npos, nref, roots, fileinfo, first;   // chunk_and_hash_batch: merged node table, file layout, per-file chunk ranges
    HostBuf h_scalars;           // pinned: [0] ncand/ntotal, [1] new chunk start
    uint64_t ndescs = 0;         // chunks accumulated in `descs`
    uint64_t chunk_start = 0;    // stream position where the open chunk starts
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    float ms_scan = 0, ms_select = 0;
    double host_sync1 = 0, host_sync2 = 0, host_alloc = 0, host_total = 0;  // wall-clock ms (diagnostics)

    bool created = false;
    yams_status_t create() {
        if (created) return YAMS_OK;
        yams_status_t rc = ensure_device(&dev);
        if (rc != YAMS_OK) return rc;
        YB_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        for (auto& e : ev) YB_CUDA(cudaEventCreate(&e));
        if ((rc = table.reserve(256 * 8)) != YAMS_OK) return rc;
        if ((rc = scalars.reserve(64)) != YAMS_OK) return rc;
        if ((rc = h_scalars.reserve(64 + 256 * 8)) != YAMS_OK) return rc;
        created = true;
        return YAMS_OK;
    }
    // (re)configure for a new stream: parameters, table, counters. Buffers are kept.
    yams_status_t init(const yams_cdc_config* cfg) {
        yams_status_t rc = create();
        if (rc != YAMS_OK) return rc;
        uint64_t* tbl = h_scalars.as<uint64_t>() + 8;   // pinned staging for the table
        YB_CUDA(cudaStreamSynchronize(st));
        rc = resolve_params(cfg, &P, tbl);
        if (rc != YAMS_OK) return rc;
        no_candidates = P.lo >= P.force;
        two_pass_only = getenv("YAMS_B200_TWO_PASS") != nullptr;
        ndescs = 0;
        chunk_start = 0;
        ms_scan = ms_select = 0;
        host_sync1 = host_sync2 = host_alloc = host_total = 0;
        two_pass_fallbacks = 0;
        YB_CUDA(cudaMemcpyAsync(table.p, tbl, 256 * 8, cudaMemcpyHostToDevice, st));
        return YAMS_OK;
    }
    void destroy() {
        for (DevBuf* b : {&table, &tile_counts, &tile_offsets, &cand, &cand_tmp, &next, &forced, &exit_, &entry, &onchain,
                          &emit_counts, &emit_offsets, &scan_scratch, &descs, &scalars, &sha_order, &npos, &nref, &roots, &fileinfo, &first})
            b->release();
        h_scalars.release();
        for (auto& e : ev)
            if (e) cudaEventDestroy(e);
        if (st) cudaStreamDestroy(st);
        st = nullptr;
        created = false;
    }

    // Candidate scan of stream positions [scan_lo, scan_hi): the ascending candidate positions land in `cand`,
    // their number in *ncand_out.  Single pass; dense (adversarial) data falls back to the exact two-pass kernels.
    yams_status_t scan(const uint8_t* data, uint64_t base_pos, uint64_t lowest, uint64_t scan_lo, uint64_t scan_hi,
                       uint32_t* ncand_out) {
        yams_status_t rc;
        uint64_t* d_sc = scalars.as<uint64_t>();
        volatile uint64_t* h_sc = h_scalars.as<uint64_t>();
        uint32_t ncand = 0;
        if (!no_candidates && scan_hi > scan_lo) {
            // tiles start at a 16-byte-aligned ADDRESS: normally at or below scan_lo (the positions in
            // [origin, scan_lo) are masked and never dereferenced); when that would fall before stream position
            // 0 (misaligned buffer at the very start of a stream) the tiles start at the next aligned address
            // and the < 16 head positions [scan_lo, origin) are tested separately by the kernels.
            uintptr_t addr_lo = reinterpret_cast<uintptr_t>(data) + (scan_lo - base_pos);
            uint64_t mis = (uint64_t)(addr_lo & 15);
            uint64_t origin = scan_lo >= mis ? scan_lo - mis : scan_lo + (16 - mis);
            uint64_t span = scan_hi > origin ? scan_hi - origin : 1;
            uint64_t ntiles64 = (span + kTileBytesHost - 1) / kTileBytesHost;
            YB_ARG(ntiles64 < (1ull << 31), "segment too large");
            uint32_t ntiles = (uint32_t)ntiles64;
            ScanArgs A{data, base_pos, lowest, origin, scan_lo, scan_hi, table.as<uint64_t>(), P};
            bool need_two_pass = two_pass_only;
            if (!two_pass_only) {
                // single pass: every warp scans a contiguous range (4 KiB tiles) and appends to its own slice
                uint32_t ntiles4 = (uint32_t)((span + kSinglePassTile - 1) / kSinglePassTile);
                uint32_t nslices = std::min<uint32_t>(ntiles4, (uint32_t)dev->sm_count * kSinglePassWarpsPerCta);
                uint32_t tpw = (ntiles4 + nslices - 1) / nslices;
                uint32_t slice_cap = std::max<uint32_t>(64u, tpw * 16u);   // 32x the 1/8192 density of random data
                size_t total_cap = (size_t)nslices * slice_cap;
                if ((rc = cand_tmp.reserve(total_cap * 8)) != YAMS_OK) return rc;
                if ((rc = cand.reserve(total_cap * 8)) != YAMS_OK) return rc;
                if ((rc = tile_counts.reserve((size_t)std::max<uint32_t>(ntiles, nslices) * 4)) != YAMS_OK) return rc;
                if ((rc = launch_scan_single_pass(A, ntiles4, dev->sm_count, cand_tmp.as<uint64_t>(), tile_counts.as<uint32_t>(),
                                                  slice_cap, nslices, cand.as<uint64_t>(), d_sc, st)) != YAMS_OK)
                    return rc;
                YB_CUDA(cudaMemcpyAsync((void*)h_sc, d_sc, 24, cudaMemcpyDeviceToHost, st));
                { const double t0 = now_ms(); YB_CUDA(cudaStreamSynchronize(st)); host_sync1 += now_ms() - t0; }
                if (h_sc[2] != 0) {
                    need_two_pass = true;   // dense candidates: exact two-pass kernels take over
                    ++two_pass_fallbacks;
                } else {
                    YB_ARG(h_sc[0] < 0xFFFFFFF0ull, "too many boundary candidates in one segment");
                    ncand = (uint32_t)h_sc[0];
                }
            }
            if (need_two_pass) {
                if ((rc = tile_counts.reserve((size_t)ntiles * 4)) != YAMS_OK) return rc;
                if ((rc = tile_offsets.reserve((size_t)ntiles * 4)) != YAMS_OK) return rc;
                uint32_t grid = std::min<uint32_t>(ntiles, (uint32_t)dev->sm_count * 8u);
                cdc_count_kernel<<<grid, 256, 0, st>>>(A, ntiles, tile_counts.as<uint32_t>());
                if ((rc = exclusive_scan_u32(tile_counts.as<uint32_t>(), tile_offsets.as<uint32_t>(), ntiles,
                                             d_sc + 0, scan_scratch, st)) != YAMS_OK)
                    return rc;
                YB_CUDA(cudaMemcpyAsync((void*)h_sc, d_sc, 8, cudaMemcpyDeviceToHost, st));
                { const double t0 = now_ms(); YB_CUDA(cudaStreamSynchronize(st)); host_sync1 += now_ms() - t0; }
                uint64_t nc = h_sc[0];
                YB_ARG(nc < 0xFFFFFFF0ull, "too many boundary candidates in one segment");
                ncand = (uint32_t)nc;
                if (ncand) {
                    if ((rc = cand.reserve((size_t)ncand * 8)) != YAMS_OK) return rc;
                    cdc_write_kernel<<<grid, 256, 0, st>>>(A, ntiles, tile_counts.as<uint32_t>(),
                                                           tile_offsets.as<uint32_t>(), cand.as<uint64_t>());
                }
            }
        }
        *ncand_out = ncand;
        return YAMS_OK;
    }

    // Scan stream positions [scan_lo, scan_hi) (bytes readable from `lowest`), select cuts from the
    // open chunk at chunk_start, append the completed chunks to descs. `data[0]` is stream position
    // base_pos.  When final, the trailing partial chunk is emitted too.
    yams_status_t process(const uint8_t* data, uint64_t base_pos, uint64_t lowest, uint64_t scan_lo,
                          uint64_t scan_hi, bool final) {
        yams_status_t rc;
        const double t_begin = now_ms();
        uint64_t* d_sc = scalars.as<uint64_t>();
        volatile uint64_t* h_sc = h_scalars.as<uint64_t>();
        uint32_t ncand = 0;
        YB_CUDA(cudaEventRecord(ev[0], st));
        if ((rc = scan(data, base_pos, lowest, scan_lo, scan_hi, &ncand)) != YAMS_OK) return rc;
        YB_CUDA(cudaEventRecord(ev[1], st));
        // ---- selection ------------------------------------------------------------------------
        uint32_t nnodes = ncand + 1;
        uint32_t nblocks = (nnodes + kNodeBlock - 1) / kNodeBlock;
        if ((rc = next.reserve((size_t)nnodes * 4)) != YAMS_OK) return rc;
        if ((rc = forced.reserve((size_t)nnodes * 4)) != YAMS_OK) return rc;
        if ((rc = exit_.reserve((size_t)nnodes * 4)) != YAMS_OK) return rc;
        if ((rc = entry.reserve((size_t)nblocks * 4)) != YAMS_OK) return rc;
        if ((rc = onchain.reserve((size_t)nnodes)) != YAMS_OK) return rc;
        if ((rc = emit_counts.reserve((size_t)nnodes * 4)) != YAMS_OK) return rc;
        if ((rc = emit_offsets.reserve((size_t)nnodes * 4)) != YAMS_OK) return rc;
        if (!cand.p && (rc = cand.reserve(8)) != YAMS_OK) return rc;
        SelectArgs S{cand.as<uint64_t>(), ncand, chunk_start, scan_hi, final ? 1 : 0, P};
        uint32_t tgrid = (nnodes + 255) / 256;
        cdc_next_kernel<<<tgrid, 256, 0, st>>>(S, next.as<uint32_t>(), forced.as<uint32_t>());
        cdc_exit_kernel<<<nblocks, 32, 0, st>>>(next.as<uint32_t>(), nnodes, exit_.as<uint32_t>());
        YB_CUDA(cudaMemsetAsync(entry.p, 0xFF, (size_t)nblocks * 4, st));
        cdc_walk_kernel<<<1, 32, 0, st>>>(exit_.as<uint32_t>(), nnodes, entry.as<uint32_t>());
        cdc_mark_kernel<<<nblocks, 32, 0, st>>>(next.as<uint32_t>(), nnodes, entry.as<uint32_t>(),
                                                onchain.as<uint8_t>());
        cdc_emit_count_kernel<<<tgrid, 256, 0, st>>>(S, next.as<uint32_t>(), forced.as<uint32_t>(),
                                                     onchain.as<uint8_t>(), emit_counts.as<uint32_t>());
        if ((rc = exclusive_scan_u32(emit_counts.as<uint32_t>(), emit_offsets.as<uint32_t>(), nnodes, d_sc + 0,
                                     scan_scratch, st)) != YAMS_OK)
            return rc;
        YB_CUDA(cudaMemcpyAsync((void*)h_sc, d_sc, 8, cudaMemcpyDeviceToHost, st));
        { const double t0 = now_ms(); YB_CUDA(cudaStreamSynchronize(st)); host_sync2 += now_ms() - t0; }
        uint64_t nnew = h_sc[0];
        { const double t0 = now_ms();
          if ((rc = descs.reserve((size_t)(ndescs + nnew + 1) * sizeof(yams_chunk_desc), true, st)) != YAMS_OK)
            return rc;
          host_alloc += now_ms() - t0; }
        cdc_emit_kernel<<<tgrid, 256, 0, st>>>(S, next.as<uint32_t>(), forced.as<uint32_t>(),
                                               onchain.as<uint8_t>(), emit_offsets.as<uint32_t>(),
                                               descs.as<yams_chunk_desc>(), ndescs, d_sc + 1);
        YB_CUDA(cudaGetLastError());
        YB_CUDA(cudaMemcpyAsync((void*)(h_sc + 1), d_sc + 1, 8, cudaMemcpyDeviceToHost, st));
        YB_CUDA(cudaEventRecord(ev[2], st));
        YB_CUDA(cudaStreamSynchronize(st));
        chunk_start = h_sc[1];
        ndescs += nnew;
        float a = 0, b = 0;
        cudaEventElapsedTime(&a, ev[0], ev[1]);
        cudaEventElapsedTime(&b, ev[1], ev[2]);
        ms_scan += a;
        ms_select += b;
        host_total += now_ms() - t_begin;
        return YAMS_OK;
    }
};

static thread_local float g_last_ms[8] = {0};

// Device-side resources of one ingest stream (kernel workspace + the two staging buffers of the host
// path). They are pooled: `yams add` calls chunk_and_hash once per file, and cudaMalloc / cudaFree of the
// workspace would otherwise dominate small inputs.
struct IngestRes {
    CdcStream cs;
    cudaStream_t copy_st = nullptr;
    DevBuf stage[2];
    cudaEvent_t copied[2] = {nullptr, nullptr};
    cudaEvent_t freed[2] = {nullptr, nullptr};
    cudaEvent_t t0 = nullptr, t1 = nullptr, e0 = nullptr, e1 = nullptr, e2 = nullptr;
    HostBuf pin[2];                                   // batch path: pinned images of two device-buffer slices
    cudaEvent_t pin_done[2] = {nullptr, nullptr};     // H2D copy out of pin[b] finished
    bool created = false;
    yams_status_t create() {
        if (created) return YAMS_OK;
        yams_status_t rc = cs.create();
        if (rc != YAMS_OK) return rc;
        YB_CUDA(cudaStreamCreateWithFlags(&copy_st, cudaStreamNonBlocking));
        for (int b = 0; b < 2; ++b) {
            YB_CUDA(cudaEventCreateWithFlags(&copied[b], cudaEventDisableTiming));
            YB_CUDA(cudaEventCreateWithFlags(&freed[b], cudaEventDisableTiming));
        }
        for (cudaEvent_t* e : {&t0, &t1, &e0, &e1, &e2}) YB_CUDA(cudaEventCreate(e));
        for (int b = 0; b < 2; ++b) YB_CUDA(cudaEventCreateWithFlags(&pin_done[b], cudaEventDisableTiming));
        created = true;
        return YAMS_OK;
    }
    void destroy() {
        if (cs.st) cudaStreamSynchronize(cs.st);
        if (copy_st) cudaStreamSynchronize(copy_st);
        for (int b = 0; b < 2; ++b) {
            stage[b].release();
            if (copied[b]) cudaEventDestroy(copied[b]);
            if (freed[b]) cudaEventDestroy(freed[b]);
        }
        for (cudaEvent_t e : {t0, t1, e0, e1, e2, pin_done[0], pin_done[1]})
            if (e) cudaEventDestroy(e);
        pin[0].release();
        pin[1].release();
        if (copy_st) cudaStreamDestroy(copy_st);
        cs.destroy();
        created = false;
    }
};

static std::mutex g_pool_mu;
static std::vector<IngestRes*> g_pool;
// one parked resource set per concurrently ingesting host thread (`yams add` workers); creating/destroying one costs
// cudaMalloc/cudaFree/cudaMallocHost calls that serialise the whole device
static size_t pool_max() {
    static const size_t v = [] { const char* e = getenv("YAMS_B200_POOL_MAX"); long x = e ? atol(e) : 16; return (size_t)(x < 1 ? 1 : x); }();
    return v;
}

static yams_status_t acquire_res(const yams_cdc_config* cfg, IngestRes** out) {
    IngestRes* r = nullptr;
    {
        // a pooled resource may be picked up by a thread that never touched CUDA: bind the plugin's device first
        DeviceCtx* dev = nullptr;
        yams_status_t drc = ensure_device(&dev);
        if (drc != YAMS_OK) return drc;
    }
    {
        std::lock_guard<std::mutex> lk(g_pool_mu);
        if (!g_pool.empty()) {
            r = g_pool.back();
            g_pool.pop_back();
        }
    }
    if (!r) r = new (std::nothrow) IngestRes();
    if (!r) return YAMS_ERR_INTERNAL;
    yams_status_t rc = r->create();
    if (rc == YAMS_OK) rc = r->cs.init(cfg);
    if (rc != YAMS_OK) {
        r->destroy();
        delete r;
        return rc;
    }
    *out = r;
    return YAMS_OK;
}

static void release_res(IngestRes* r) {
    if (!r) return;
    if (r->cs.st) cudaStreamSynchronize(r->cs.st);
    if (r->copy_st) cudaStreamSynchronize(r->copy_st);
    // drop very large buffers instead of parking them in the pool
    const size_t kKeep = 1ull << 30;
    for (DevBuf* b : {&r->stage[0], &r->stage[1], &r->cs.descs, &r->cs.cand, &r->cs.cand_tmp})
        if (b->cap > kKeep) b->release();
    {
        std::lock_guard<std::mutex> lk(g_pool_mu);
        if (g_pool.size() < pool_max()) {
            g_pool.push_back(r);
            return;
        }
    }
    r->destroy();
    delete r;
}

static yams_status_t copy_out(CdcStream& cs, uint64_t first, uint64_t n, yams_chunk_desc** out, size_t* out_n) {
    *out = nullptr;
    *out_n = 0;
    if (n == 0) return YAMS_OK;
    yams_chunk_desc* h = static_cast<yams_chunk_desc*>(malloc((size_t)n * sizeof(yams_chunk_desc)));
    if (!h) {
        set_last_error("out of host memory for %llu chunk descriptors", (unsigned long long)n);
        return YAMS_ERR_INTERNAL;
    }
    cudaError_t e = cudaMemcpyAsync(h, cs.descs.as<yams_chunk_desc>() + first, (size_t)n * sizeof(yams_chunk_desc),
                                    cudaMemcpyDeviceToHost, cs.st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(cs.st);
    if (e != cudaSuccess) {
        free(h);
        set_last_error("D2H of chunk descriptors failed: %s", cudaGetErrorString(e));
        return YAMS_ERR_INTERNAL;
    }
    *out = h;
    *out_n = (size_t)n;
    return YAMS_OK;
}

// One-shot over device-resident data.
static yams_status_t run_device(const uint8_t* d_data, size_t len, const yams_cdc_config* cfg, bool hash,
                                yams_chunk_desc** out, size_t* out_n) {
    IngestRes* r = nullptr;
    yams_status_t rc = acquire_res(cfg, &r);
    if (rc != YAMS_OK) return rc;
    CdcStream& cs = r->cs;
    // size the descriptor table once: every chunk but the last is at least max(min_chunk,1) bytes long
    {
        uint64_t per = std::max<uint64_t>(cs.P.lo + 1, 4096);
        rc = cs.descs.reserve((size_t)(len / per + 16) * sizeof(yams_chunk_desc));
    }
    cudaEventRecord(r->e0, cs.st);
    const uint64_t seg = segment_bytes();
    for (uint64_t lo = 0; lo < len && rc == YAMS_OK; lo += seg) {
        uint64_t hi = std::min<uint64_t>(len, lo + seg);
        rc = cs.process(d_data, 0, 0, lo, hi, hi == len);
    }
    cudaEventRecord(r->e1, cs.st);
    if (rc == YAMS_OK && hash && cs.ndescs) {
        if (cs.ndescs >= 0xFFFFFFFFull) {
            set_last_error("too many chunks");
            rc = YAMS_ERR_INVALID_ARG;
        } else {
            rc = cs.sha_order.reserve(sha256_order_ws_bytes(cs.ndescs));
            if (rc == YAMS_OK)
                rc = launch_sha256_chunks(d_data, 0, cs.descs.as<yams_chunk_desc>(), 0, (uint32_t)cs.ndescs,
                                          reinterpret_cast<unsigned int*>(cs.scalars.as<uint64_t>() + 4), cs.dev->sm_count, cs.st,
                                          cs.sha_order.as<uint32_t>(), len);
        }
    }
    cudaEventRecord(r->e2, cs.st);
    if (rc == YAMS_OK) rc = copy_out(cs, 0, cs.ndescs, out, out_n);
    if (rc == YAMS_OK) {
        float t_sha = 0, t_all = 0;
        cudaEventElapsedTime(&t_sha, r->e1, r->e2);
        cudaEventElapsedTime(&t_all, r->e0, r->e2);
        g_last_ms[0] = cs.ms_scan; g_last_ms[1] = cs.ms_select; g_last_ms[2] = t_sha; g_last_ms[3] = t_all;
        g_last_ms[4] = (float)cs.host_sync1; g_last_ms[5] = (float)cs.host_sync2; g_last_ms[6] = (float)cs.host_alloc;
        g_last_ms[7] = (float)cs.host_total;
    }
    release_res(r);
    return rc;
}

// Upload from ordinary (pageable) host memory.  A plain cudaMemcpy moves pageable memory at ~10 GB/s through the
// driver's single staging thread; here host threads assemble each slice of the destination in one of two pinned buffers
// (fill(dst, lo, hi) writes the image of destination bytes [lo, hi)) while the previous slice is on the wire.
static uint64_t stage_slice_bytes() {
    static const uint64_t v = [] { const char* e = getenv("YAMS_B200_STAGE_MIB"); long x = e ? atol(e) : 64; return (uint64_t)(x < 1 ? 1 : x) << 20; }();
    return v;
}
static unsigned stage_threads() {
    static const unsigned v = [] {
        const char* e = getenv("YAMS_B200_STAGE_THREADS");
        long x = e ? atol(e) : (long)std::min(8u, std::max(1u, std::thread::hardware_concurrency()));
        return (unsigned)(x < 1 ? 1 : (x > 64 ? 64 : x));
    }();
    return v;
}
template <typename Fill>
static yams_status_t staged_upload(IngestRes* r, uint8_t* d_dst, uint64_t total, const Fill& fill, cudaStream_t st) {
    if (total == 0) return YAMS_OK;
    yams_status_t rc;
    const uint64_t S = std::min<uint64_t>(stage_slice_bytes(), (total + 255) & ~255ull);
    for (int b = 0; b < 2; ++b)
        if ((rc = r->pin[b].reserve((size_t)S)) != YAMS_OK) return rc;
    const unsigned T = total < (4u << 20) ? 1u : stage_threads();
    int b = 0;
    for (uint64_t lo = 0; lo < total; lo += S, b ^= 1) {
        const uint64_t hi = std::min<uint64_t>(total, lo + S);
        YB_CUDA(cudaEventSynchronize(r->pin_done[b]));   // the copy that last used this pinned buffer has finished
        uint8_t* img = r->pin[b].as<uint8_t>();
        if (T == 1) {
            fill(img, lo, hi);
        } else {
            std::vector<std::thread> th;
            const uint64_t part = (((hi - lo) + T - 1) / T + 4095) & ~4095ull;
            for (unsigned t = 0; t < T; ++t) {
                uint64_t a = lo + (uint64_t)t * part, e = std::min<uint64_t>(hi, a + part);
                if (a < e) th.emplace_back([&fill, img, lo, a, e] { fill(img + (a - lo), a, e); });
            }
            for (auto& x : th) x.join();
        }
        YB_CUDA(cudaMemcpyAsync(d_dst + lo, img, (size_t)(hi - lo), cudaMemcpyHostToDevice, st));
        YB_CUDA(cudaEventRecord(r->pin_done[b], st));
    }
    return YAMS_OK;
}
static bool is_pageable(const void* p) {
    cudaPointerAttributes a{};
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return true;
    }
    return a.type == cudaMemoryTypeUnregistered;
}

// ---- many files per call ----------------------------------------------------------------------------------------
// One group = consecutive files whose padded total fits the staging budget.  Layout: every file starts at a
// 256-byte-aligned buffer position preceded by >= 64 zero bytes (the whole buffer is zeroed first), so ONE candidate
// scan, ONE selection pass and ONE SHA-256 launch serve every file of the group.
constexpr uint64_t kBatchGap = 64;
static uint64_t batch_group_bytes() {
    static const uint64_t v = [] { const char* e = getenv("YAMS_B200_BATCH_MIB"); long x = e ? atol(e) : 1024; return (uint64_t)(x < 1 ? 1 : x) << 20; }();
    return v;
}
static inline uint64_t batch_padded(uint64_t len) { return (len + kBatchGap + 255) & ~255ull; }

static yams_status_t run_batch_group(IngestRes* r, const uint8_t* const* files, const size_t* lens, size_t f0, size_t f1, bool hash,
                                     std::vector<yams_chunk_desc>& out, uint64_t* out_first, float* ms_acc) {
    CdcStream& cs = r->cs;
    cudaStream_t st = cs.st;
    yams_status_t rc;
    const uint32_t nf = (uint32_t)(f1 - f0);
    std::vector<uint64_t> lay(2 * (size_t)nf);   // starts | ends
    uint64_t pos = 256;
    for (uint32_t f = 0; f < nf; ++f) {
        lay[f] = pos;
        lay[nf + f] = pos + lens[f0 + f];
        pos += batch_padded(lens[f0 + f]);
    }
    const uint64_t L = pos;
    if ((rc = r->stage[0].reserve((size_t)L + 256)) != YAMS_OK) return rc;
    uint8_t* d_buf = r->stage[0].as<uint8_t>();
    if ((rc = cs.fileinfo.reserve((size_t)nf * 16)) != YAMS_OK) return rc;
    if ((rc = cs.roots.reserve((size_t)nf * 4)) != YAMS_OK) return rc;
    if ((rc = cs.first.reserve(((size_t)nf + 1) * 8)) != YAMS_OK) return rc;
    YB_CUDA(cudaEventRecord(r->e0, st));
    // Upload: the files sit in ordinary (pageable) host memory, which a plain cudaMemcpy moves at ~10 GB/s through
    // the driver's single staging thread.  Instead, host threads assemble an exact image of each 32 MiB slice of the
    // device buffer (file bytes + zero gaps) in one of two pinned buffers while the previous slice is on the wire.
    {
        auto fill = [&](uint8_t* dst, uint64_t lo, uint64_t hi) {   // image of buffer positions [lo, hi)
            // first file whose end lies beyond lo
            uint32_t f = (uint32_t)(std::upper_bound(lay.begin() + nf, lay.begin() + 2 * (size_t)nf, lo) - (lay.begin() + nf));
            uint64_t p = lo;
            while (p < hi) {
                if (f < nf && lay[f] <= p) {   // inside file f
                    uint64_t e = std::min<uint64_t>(hi, lay[nf + f]);
                    if (e > p) memcpy(dst + (p - lo), files[f0 + f] + (p - lay[f]), (size_t)(e - p));
                    p = std::max(p, e);
                    if (p >= lay[nf + f]) ++f;
                } else {                       // gap up to the next file start (or the end of the buffer)
                    uint64_t e = std::min<uint64_t>(hi, f < nf ? lay[f] : hi);
                    memset(dst + (p - lo), 0, (size_t)(e - p));
                    p = e;
                }
            }
        };
        if ((rc = staged_upload(r, d_buf, L + 256, fill, st)) != YAMS_OK) return rc;
    }
    YB_CUDA(cudaMemcpyAsync(cs.fileinfo.p, lay.data(), (size_t)nf * 16, cudaMemcpyHostToDevice, st));
    YB_CUDA(cudaEventRecord(cs.ev[0], st));
    uint32_t ncand = 0;
    if ((rc = cs.scan(d_buf, 0, 0, 0, L, &ncand)) != YAMS_OK) return rc;
    YB_CUDA(cudaEventRecord(cs.ev[1], st));
    // ---- selection over the merged node table ----
    YB_ARG((uint64_t)ncand + nf < 0xFFFFFFF0ull, "too many nodes in one batch group");
    const uint32_t nnodes = ncand + nf;
    const uint32_t nblocks = (nnodes + kNodeBlock - 1) / kNodeBlock;
    if ((rc = cs.npos.reserve((size_t)nnodes * 8)) != YAMS_OK) return rc;
    if ((rc = cs.nref.reserve((size_t)nnodes * 4)) != YAMS_OK) return rc;
    if ((rc = cs.next.reserve((size_t)nnodes * 4)) != YAMS_OK) return rc;
    if ((rc = cs.forced.reserve((size_t)nnodes * 4)) != YAMS_OK) return rc;
    if ((rc = cs.exit_.reserve((size_t)nnodes * 4)) != YAMS_OK) return rc;
    if ((rc = cs.entry.reserve((size_t)nblocks * 4)) != YAMS_OK) return rc;
    if ((rc = cs.onchain.reserve((size_t)nnodes)) != YAMS_OK) return rc;
    if ((rc = cs.emit_counts.reserve((size_t)nnodes * 4)) != YAMS_OK) return rc;
    if ((rc = cs.emit_offsets.reserve((size_t)nnodes * 4)) != YAMS_OK) return rc;
    if (!cs.cand.p && (rc = cs.cand.reserve(8)) != YAMS_OK) return rc;
    uint64_t* d_sc = cs.scalars.as<uint64_t>();
    volatile uint64_t* h_sc = cs.h_scalars.as<uint64_t>();
    BatchArgs B{BatchLayout{cs.cand.as<uint64_t>(), ncand, cs.fileinfo.as<uint64_t>(), cs.fileinfo.as<uint64_t>() + nf, nf}, cs.P};
    const uint32_t tgrid = (nnodes + 255) / 256;
    batch_nodes_kernel<<<tgrid, 256, 0, st>>>(B, cs.npos.as<uint64_t>(), cs.nref.as<uint32_t>(), cs.roots.as<uint32_t>());
    batch_next_kernel<<<tgrid, 256, 0, st>>>(B, cs.npos.as<uint64_t>(), cs.nref.as<uint32_t>(), cs.roots.as<uint32_t>(), nnodes,
                                             cs.next.as<uint32_t>(), cs.forced.as<uint32_t>(), cs.emit_counts.as<uint32_t>());
    cdc_exit_kernel<<<nblocks, 32, 0, st>>>(cs.next.as<uint32_t>(), nnodes, cs.exit_.as<uint32_t>());
    YB_CUDA(cudaMemsetAsync(cs.entry.p, 0xFF, (size_t)nblocks * 4, st));
    YB_CUDA(cudaMemsetAsync(cs.onchain.p, 0, (size_t)nnodes, st));
    cdc_walk_kernel<<<1, 32, 0, st>>>(cs.exit_.as<uint32_t>(), nnodes, cs.entry.as<uint32_t>());
    cdc_mark_kernel<<<nblocks, 32, 0, st>>>(cs.next.as<uint32_t>(), nnodes, cs.entry.as<uint32_t>(), cs.onchain.as<uint8_t>());
    batch_mask_counts_kernel<<<tgrid, 256, 0, st>>>(cs.onchain.as<uint8_t>(), cs.emit_counts.as<uint32_t>(), nnodes);
    if ((rc = exclusive_scan_u32(cs.emit_counts.as<uint32_t>(), cs.emit_offsets.as<uint32_t>(), nnodes, d_sc + 0, cs.scan_scratch,
                                 st)) != YAMS_OK)
        return rc;
    YB_CUDA(cudaMemcpyAsync((void*)h_sc, d_sc, 8, cudaMemcpyDeviceToHost, st));
    YB_CUDA(cudaStreamSynchronize(st));
    const uint64_t nnew = h_sc[0];
    YB_ARG(nnew < 0xFFFFFFFFull, "too many chunks in one batch group");
    if ((rc = cs.descs.reserve((size_t)(nnew + 1) * sizeof(yams_chunk_desc))) != YAMS_OK) return rc;
    batch_emit_kernel<<<tgrid, 256, 0, st>>>(B, cs.npos.as<uint64_t>(), cs.next.as<uint32_t>(), cs.forced.as<uint32_t>(),
                                             cs.onchain.as<uint8_t>(), cs.emit_offsets.as<uint32_t>(), nnodes,
                                             cs.descs.as<yams_chunk_desc>());
    batch_first_kernel<<<(nf + 1 + 255) / 256, 256, 0, st>>>(cs.roots.as<uint32_t>(), cs.emit_offsets.as<uint32_t>(), nf, d_sc + 0,
                                                             cs.first.as<uint64_t>());
    YB_CUDA(cudaEventRecord(cs.ev[2], st));
    if (hash && nnew) {
        if ((rc = cs.sha_order.reserve(sha256_order_ws_bytes(nnew))) != YAMS_OK) return rc;
        if ((rc = launch_sha256_chunks(d_buf, 0, cs.descs.as<yams_chunk_desc>(), 0, (uint32_t)nnew,
                                       reinterpret_cast<unsigned int*>(d_sc + 4), cs.dev->sm_count, st, cs.sha_order.as<uint32_t>(),
                                       L)) != YAMS_OK)
            return rc;
    }
    YB_CUDA(cudaEventRecord(cs.ev[3], st));
    if (nnew) batch_rebase_kernel<<<(unsigned)((nnew + 255) / 256), 256, 0, st>>>(cs.descs.as<yams_chunk_desc>(), nnew,
                                                                               cs.fileinfo.as<uint64_t>(), nf);
    YB_CUDA(cudaGetLastError());
    const size_t base = out.size();
    out.resize(base + (size_t)nnew);
    std::vector<uint64_t> first((size_t)nf + 1);
    if (nnew) YB_CUDA(cudaMemcpyAsync(out.data() + base, cs.descs.p, (size_t)nnew * sizeof(yams_chunk_desc), cudaMemcpyDeviceToHost, st));
    YB_CUDA(cudaMemcpyAsync(first.data(), cs.first.p, ((size_t)nf + 1) * 8, cudaMemcpyDeviceToHost, st));
    YB_CUDA(cudaEventRecord(r->e1, st));
    YB_CUDA(cudaStreamSynchronize(st));
    if (out_first)
        for (uint32_t f = 0; f <= nf; ++f) out_first[f0 + f] = base + first[f];
    float a = 0, b = 0, c = 0, d = 0;
    cudaEventElapsedTime(&a, cs.ev[0], cs.ev[1]);
    cudaEventElapsedTime(&b, cs.ev[1], cs.ev[2]);
    cudaEventElapsedTime(&c, cs.ev[2], cs.ev[3]);
    cudaEventElapsedTime(&d, r->e0, r->e1);
    ms_acc[0] += a; ms_acc[1] += b; ms_acc[2] += c; ms_acc[3] += d;
    return YAMS_OK;
}

}  // namespace yb

using namespace yb;

// Streaming session: host bytes are staged [carry | slice] into one of two device buffers; the H2D copy
// of slice i+1 overlaps the kernels of slice i (when the host memory is pinned).
struct yams_b200_ingest {
    IngestRes* r = nullptr;
    uint64_t head = 0;            // bytes reserved in front of a slice for the carry
    uint64_t stream_pos = 0;      // bytes fed so far
    uint64_t keep_from = 0;       // lowest stream position still resident on the device
    const uint8_t* res_ptr = nullptr;  // device address of stream position keep_from
    int cur = -1;                 // stage buffer holding the resident bytes (-1: none yet)
    bool hash = true;
    bool finished = false;
    float ms_sha = 0;
    uint64_t sha_pos = 0;         // stream position up to which chunks have been hashed (mean chunk size of a slice)
};

static yams_status_t session_open(const yams_cdc_config* cfg, bool hash, yams_b200_ingest** out) {
    yams_b200_ingest* s = new (std::nothrow) yams_b200_ingest();
    if (!s) return YAMS_ERR_INTERNAL;
    s->hash = hash;
    yams_status_t rc = acquire_res(cfg, &s->r);
    if (rc != YAMS_OK) {
        delete s;
        return rc;
    }
    // the open chunk is always shorter than `force`, and a position needs <= kHistory bytes behind it
    uint64_t need = std::max<uint64_t>(s->r->cs.P.force, (uint64_t)kHistory) + 16;
    s->head = (need + 255) & ~255ull;
    *out = s;
    return YAMS_OK;
}

static void session_close(yams_b200_ingest* s) {
    if (!s) return;
    release_res(s->r);
    delete s;
}

static yams_status_t session_sha(yams_b200_ingest* s, const uint8_t* data, uint64_t base_pos, uint64_t first) {
    CdcStream& cs = s->r->cs;
    if (!s->hash || cs.ndescs <= first) return YAMS_OK;
    YB_ARG(cs.ndescs - first < 0xFFFFFFFFull, "too many chunks in one slice");
    YB_CUDA(cudaEventRecord(s->r->t0, cs.st));
    yams_status_t rc = cs.sha_order.reserve(sha256_order_ws_bytes(cs.ndescs - first));
    if (rc != YAMS_OK) return rc;
    // mean chunk size of the slice: the new chunks end at the open chunk's start
    const uint64_t slice_bytes = cs.chunk_start > s->sha_pos ? cs.chunk_start - s->sha_pos : 0;
    s->sha_pos = cs.chunk_start;
    rc = launch_sha256_chunks(data, base_pos, cs.descs.as<yams_chunk_desc>(), (uint32_t)first, (uint32_t)(cs.ndescs - first),
                              reinterpret_cast<unsigned int*>(cs.scalars.as<uint64_t>() + 4), cs.dev->sm_count, cs.st,
                              cs.sha_order.as<uint32_t>(), slice_bytes);
    if (rc != YAMS_OK) return rc;
    YB_CUDA(cudaEventRecord(s->r->t1, cs.st));
    YB_CUDA(cudaEventSynchronize(s->r->t1));
    float ms = 0;
    cudaEventElapsedTime(&ms, s->r->t0, s->r->t1);
    s->ms_sha += ms;
    return YAMS_OK;
}

// Feed `len` host bytes; when `final`, also close the stream.  Newly completed chunks are appended to
// cs.descs (digests included when hashing is on).
static yams_status_t session_feed(yams_b200_ingest* s, const uint8_t* data, size_t len, bool final) {
    IngestRes* r = s->r;
    CdcStream& cs = r->cs;
    yams_status_t rc;
    YB_ARG(!s->finished, "ingest session already finished");
    const size_t nslices = (size_t)((len + kFeedSlice - 1) / kFeedSlice);
    if (nslices == 0) {
        if (final) {
            uint64_t first = cs.ndescs;
            rc = cs.process(s->res_ptr, s->keep_from, s->keep_from, s->stream_pos, s->stream_pos, true);
            if (rc != YAMS_OK) return rc;
            if ((rc = session_sha(s, s->res_ptr, s->keep_from, first)) != YAMS_OK) return rc;
            s->finished = true;
        }
        return YAMS_OK;
    }
    auto slice_len = [&](size_t i) { return (size_t)std::min<uint64_t>(kFeedSlice, (uint64_t)len - i * kFeedSlice); };
    const bool pageable = len >= (4u << 20) && is_pageable(data);
    auto issue_copy = [&](size_t i, int b) -> yams_status_t {
        size_t sl = slice_len(i);
        yams_status_t rr = r->stage[b].reserve((size_t)s->head + sl + 64);
        if (rr != YAMS_OK) return rr;
        YB_CUDA(cudaStreamWaitEvent(r->copy_st, r->freed[b], 0));
        const uint8_t* src = data + i * kFeedSlice;
        if (pageable && sl >= (4u << 20)) {
            // ordinary host memory: pinned double-buffered upload filled by host threads (see staged_upload)
            auto fill = [src](uint8_t* dst, uint64_t lo, uint64_t hi) { memcpy(dst, src + lo, (size_t)(hi - lo)); };
            yams_status_t ur = staged_upload(r, r->stage[b].as<uint8_t>() + s->head, sl, fill, r->copy_st);
            if (ur != YAMS_OK) return ur;
        } else {
            YB_CUDA(cudaMemcpyAsync(r->stage[b].as<uint8_t>() + s->head, src, sl, cudaMemcpyHostToDevice, r->copy_st));
        }
        YB_CUDA(cudaEventRecord(r->copied[b], r->copy_st));
        return YAMS_OK;
    };
    int b = s->cur < 0 ? 0 : (s->cur ^ 1);
    if ((rc = issue_copy(0, b)) != YAMS_OK) return rc;
    for (size_t i = 0; i < nslices; ++i) {
        const size_t sl = slice_len(i);
        const uint64_t carry = s->stream_pos - s->keep_from;
        uint8_t* dst0 = r->stage[b].as<uint8_t>() + s->head;  // device address of stream position stream_pos
        if (carry) {
            YB_CUDA(cudaMemcpyAsync(dst0 - carry, s->res_ptr, (size_t)carry, cudaMemcpyDeviceToDevice, cs.st));
        }
        if (s->cur >= 0) YB_CUDA(cudaEventRecord(r->freed[s->cur], cs.st));
        if (i + 1 < nslices) {
            if ((rc = issue_copy(i + 1, b ^ 1)) != YAMS_OK) return rc;
        }
        YB_CUDA(cudaStreamWaitEvent(cs.st, r->copied[b], 0));
        const uint64_t new_pos = s->stream_pos + sl;
        const uint64_t first = cs.ndescs;
        const uint8_t* view = dst0 - carry;  // stream position keep_from
        const bool last = final && (i + 1 == nslices);
        rc = cs.process(view, s->keep_from, s->keep_from, s->stream_pos, new_pos, last);
        if (rc != YAMS_OK) return rc;
        if ((rc = session_sha(s, view, s->keep_from, first)) != YAMS_OK) return rc;
        s->stream_pos = new_pos;
        uint64_t nk = std::min<uint64_t>(cs.chunk_start, new_pos > (uint64_t)kHistory ? new_pos - kHistory : 0);
        if (nk < s->keep_from) nk = s->keep_from;
        s->res_ptr = view + (nk - s->keep_from);
        s->keep_from = nk;
        s->cur = b;
        b ^= 1;
    }
    if (final) s->finished = true;
    return YAMS_OK;
}

static yams_status_t session_take(yams_b200_ingest* s, yams_chunk_desc** out, size_t* out_n) {
    yams_status_t rc = copy_out(s->r->cs, 0, s->r->cs.ndescs, out, out_n);
    if (rc == YAMS_OK) s->r->cs.ndescs = 0;  // descriptors handed over; reuse the device table
    return rc;
}

static yams_status_t run_host(const uint8_t* data, size_t len, const yams_cdc_config* cfg, bool hash,
                              yams_chunk_desc** out, size_t* out_n) {
    yams_b200_ingest* s = nullptr;
    yams_status_t rc = session_open(cfg, hash, &s);
    if (rc != YAMS_OK) return rc;
    CdcStream& cs = s->r->cs;
    {
        uint64_t per = std::max<uint64_t>(cs.P.lo + 1, 4096);
        rc = cs.descs.reserve((size_t)(len / per + 16) * sizeof(yams_chunk_desc));
    }
    cudaEventRecord(s->r->e0, cs.st);
    if (rc == YAMS_OK) rc = session_feed(s, data, len, true);
    if (rc == YAMS_OK) rc = session_take(s, out, out_n);
    cudaEventRecord(s->r->e1, cs.st);
    cudaEventSynchronize(s->r->e1);
    float t = 0;
    cudaEventElapsedTime(&t, s->r->e0, s->r->e1);
    g_last_ms[0] = cs.ms_scan; g_last_ms[1] = cs.ms_select; g_last_ms[2] = s->ms_sha; g_last_ms[3] = t;
    g_last_ms[4] = (float)cs.host_sync1; g_last_ms[5] = (float)cs.host_sync2; g_last_ms[6] = (float)cs.host_alloc;
    g_last_ms[7] = (float)cs.host_total;
    session_close(s);
    return rc;
}

extern "C" {

void yams_b200_cdc_default_config(yams_cdc_config* cfg) {
    // /root/reference/include/yams/chunking/chunker.h:44-51, include/yams/core/types.h:280-285;
    // `yams add` uses StreamingChunker: src/api/content_store_builder.cpp:152-165
    if (!cfg) return;
    cfg->window_size = 48;
    cfg->min_chunk = 16 * 1024;
    cfg->max_chunk = 1024 * 1024;
    cfg->polynomial = kDefaultPoly;
    cfg->mask = 0x1FFF;
    cfg->variant = YAMS_CDC_STREAMING;
    cfg->reserved = 0;
}

yams_status_t yams_b200_chunk_and_hash(void* self, const uint8_t* data, size_t len, const yams_cdc_config* cfg,
                                       yams_chunk_desc** out, size_t* out_n) {
    YB_TRY
    (void)self;
    YB_ARG(out && out_n, "out / out_n is null");
    *out = nullptr;
    *out_n = 0;
    YB_ARG(data || len == 0, "data is null");
    YB_ARG(cfg, "cfg is null");
    return run_host(data, len, cfg, true, out, out_n);
    YB_CATCH
}

yams_status_t yams_b200_chunk_boundaries(void* self, const uint8_t* data, size_t len, const yams_cdc_config* cfg,
                                         yams_chunk_desc** out, size_t* out_n) {
    YB_TRY
    (void)self;
    YB_ARG(out && out_n, "out / out_n is null");
    *out = nullptr;
    *out_n = 0;
    YB_ARG(data || len == 0, "data is null");
    YB_ARG(cfg, "cfg is null");
    yams_status_t rc = run_host(data, len, cfg, false, out, out_n);
    if (rc == YAMS_OK)
        for (size_t i = 0; i < *out_n; ++i) memset((*out)[i].digest, 0, 32);
    return rc;
    YB_CATCH
}

yams_status_t yams_b200_chunk_and_hash_device(void* self, const uint8_t* d_data, size_t len,
                                              const yams_cdc_config* cfg, yams_chunk_desc** out, size_t* out_n) {
    YB_TRY
    (void)self;
    YB_ARG(out && out_n, "out / out_n is null");
    *out = nullptr;
    *out_n = 0;
    YB_ARG(d_data || len == 0, "d_data is null");
    YB_ARG(cfg, "cfg is null");
    return run_device(d_data, len, cfg, true, out, out_n);
    YB_CATCH
}

yams_status_t yams_b200_chunk_and_hash_batch(void* self, const uint8_t* const* files, const size_t* lens, size_t n_files,
                                             const yams_cdc_config* cfg, yams_chunk_desc** out, size_t* out_n, uint64_t* out_first) {
    YB_TRY
    (void)self;
    YB_ARG(out && out_n, "out / out_n is null");
    *out = nullptr;
    *out_n = 0;
    YB_ARG(cfg, "cfg is null");
    if (out_first) out_first[0] = 0;
    if (n_files == 0) return YAMS_OK;
    YB_ARG(files && lens && out_first, "null argument");
    for (size_t f = 0; f < n_files; ++f) YB_ARG(files[f] || lens[f] == 0, "a file pointer is null");
    IngestRes* r = nullptr;
    yams_status_t rc = acquire_res(cfg, &r);
    if (rc != YAMS_OK) return rc;
    std::vector<yams_chunk_desc> all;
    float ms[4] = {0, 0, 0, 0};
    const uint64_t budget = batch_group_bytes();
    size_t f = 0;
    while (f < n_files && rc == YAMS_OK) {
        if (batch_padded(lens[f]) + 256 > budget) {
            // a file larger than a whole group goes through the streaming path on its own
            yams_chunk_desc* d = nullptr;
            size_t n = 0;
            rc = run_host(files[f], lens[f], cfg, true, &d, &n);
            if (rc == YAMS_OK) {
                out_first[f] = all.size();
                all.insert(all.end(), d, d + n);
                out_first[f + 1] = all.size();
                free(d);
                for (int i = 0; i < 4; ++i) ms[i] += g_last_ms[i];
            }
            ++f;
            continue;
        }
        size_t g = f;
        uint64_t tot = 256;
        while (g < n_files && tot + batch_padded(lens[g]) <= budget) tot += batch_padded(lens[g++]);
        rc = run_batch_group(r, files, lens, f, g, true, all, out_first, ms);
        f = g;
    }
    release_res(r);
    if (rc != YAMS_OK) return rc;
    for (int i = 0; i < 4; ++i) g_last_ms[i] = ms[i];
    g_last_ms[4] = g_last_ms[5] = g_last_ms[6] = g_last_ms[7] = 0;
    if (!all.empty()) {
        yams_chunk_desc* h = static_cast<yams_chunk_desc*>(malloc(all.size() * sizeof(yams_chunk_desc)));
        if (!h) {
            set_last_error("out of host memory for %llu chunk descriptors", (unsigned long long)all.size());
            return YAMS_ERR_INTERNAL;
        }
        memcpy(h, all.data(), all.size() * sizeof(yams_chunk_desc));
        *out = h;
        *out_n = all.size();
    }
    return YAMS_OK;
    YB_CATCH
}

void yams_b200_free_chunks(void* self, yams_chunk_desc* chunks, size_t n) {
    (void)self;
    (void)n;
    free(chunks);
}

yams_status_t yams_b200_ingest_open(void* self, const yams_cdc_config* cfg, yams_b200_ingest** out) {
    YB_TRY
    (void)self;
    YB_ARG(out, "out is null");
    *out = nullptr;
    YB_ARG(cfg, "cfg is null");
    return session_open(cfg, true, out);
    YB_CATCH
}

yams_status_t yams_b200_ingest_feed(yams_b200_ingest* s, const uint8_t* data, size_t len, yams_chunk_desc** out,
                                    size_t* out_n) {
    YB_TRY
    YB_ARG(s && out && out_n, "null argument");
    if (s->r && s->r->cs.dev) cudaSetDevice(s->r->cs.dev->device);   // a session may be fed from another thread than the one that opened it
    *out = nullptr;
    *out_n = 0;
    YB_ARG(data || len == 0, "data is null");
    yams_status_t rc = session_feed(s, data, len, false);
    if (rc != YAMS_OK) return rc;
    return session_take(s, out, out_n);
    YB_CATCH
}

yams_status_t yams_b200_ingest_finish(yams_b200_ingest* s, yams_chunk_desc** out, size_t* out_n) {
    YB_TRY
    YB_ARG(s && out && out_n, "null argument");
    if (s->r && s->r->cs.dev) cudaSetDevice(s->r->cs.dev->device);   // a session may be fed from another thread than the one that opened it
    *out = nullptr;
    *out_n = 0;
    yams_status_t rc = session_feed(s, nullptr, 0, true);
    if (rc != YAMS_OK) return rc;
    return session_take(s, out, out_n);
    YB_CATCH
}

void yams_b200_ingest_close(yams_b200_ingest* s) { session_close(s); }

static yams_status_t sha_batch_impl(const uint8_t* d_base, size_t base_len, const uint64_t* offsets,
                                    const uint64_t* sizes, size_t n, uint8_t* digests, DeviceCtx* dev,
                                    cudaStream_t st) {
    YB_ARG(n < 0xFFFFFFFFull, "too many spans");
    std::vector<yams_chunk_desc> h(n);
    for (size_t i = 0; i < n; ++i) {
        YB_ARG(offsets[i] <= base_len && sizes[i] <= base_len - offsets[i], "span outside the base buffer");
        h[i].offset = offsets[i];
        h[i].size = sizes[i];
    }
    DevBuf d_descs, d_cnt;
    yams_status_t rc = d_descs.reserve(n * sizeof(yams_chunk_desc));
    if (rc == YAMS_OK) rc = d_cnt.reserve(16);
    if (rc == YAMS_OK) {
        cudaError_t e = cudaMemcpyAsync(d_descs.p, h.data(), n * sizeof(yams_chunk_desc), cudaMemcpyHostToDevice, st);
        if (e != cudaSuccess) { set_last_error("H2D failed: %s", cudaGetErrorString(e)); rc = YAMS_ERR_INTERNAL; }
    }
    if (rc == YAMS_OK)
        rc = launch_sha256_chunks(d_base, 0, d_descs.as<yams_chunk_desc>(), 0, (uint32_t)n,
                                  d_cnt.as<unsigned int>(), dev->sm_count, st);
    if (rc == YAMS_OK) {
        cudaError_t e = cudaMemcpyAsync(h.data(), d_descs.p, n * sizeof(yams_chunk_desc), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) { set_last_error("sha256 batch failed: %s", cudaGetErrorString(e)); rc = YAMS_ERR_INTERNAL; }
    }
    if (rc == YAMS_OK)
        for (size_t i = 0; i < n; ++i) memcpy(digests + 32 * i, h[i].digest, 32);
    d_descs.release();
    d_cnt.release();
    return rc;
}

yams_status_t yams_b200_sha256_batch_device(void* self, const uint8_t* d_base, size_t base_len,
                                            const uint64_t* offsets, const uint64_t* sizes, size_t n,
                                            uint8_t* digests) {
    YB_TRY
    (void)self;
    if (n == 0) return YAMS_OK;
    YB_ARG(offsets && sizes && digests, "null argument");
    YB_ARG(d_base || base_len == 0, "d_base is null");
    DeviceCtx* dev = nullptr;
    yams_status_t rc = ensure_device(&dev);
    if (rc != YAMS_OK) return rc;
    cudaStream_t st;
    YB_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    rc = sha_batch_impl(d_base, base_len, offsets, sizes, n, digests, dev, st);
    cudaStreamDestroy(st);
    return rc;
    YB_CATCH
}

yams_status_t yams_b200_sha256_batch(void* self, const uint8_t* base, size_t base_len, const uint64_t* offsets,
                                     const uint64_t* sizes, size_t n, uint8_t* digests) {
    YB_TRY
    (void)self;
    if (n == 0) return YAMS_OK;
    YB_ARG(offsets && sizes && digests, "null argument");
    YB_ARG(base || base_len == 0, "base is null");
    DeviceCtx* dev = nullptr;
    yams_status_t rc = ensure_device(&dev);
    if (rc != YAMS_OK) return rc;
    cudaStream_t st;
    YB_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    DevBuf d_base;
    rc = d_base.reserve(base_len + 64);
    if (rc == YAMS_OK && base_len) {
        cudaError_t e = cudaMemcpyAsync(d_base.p, base, base_len, cudaMemcpyHostToDevice, st);
        if (e != cudaSuccess) { set_last_error("H2D failed: %s", cudaGetErrorString(e)); rc = YAMS_ERR_INTERNAL; }
    }
    if (rc == YAMS_OK) rc = sha_batch_impl(d_base.as<uint8_t>(), base_len, offsets, sizes, n, digests, dev, st);
    d_base.release();
    cudaStreamDestroy(st);
    return rc;
    YB_CATCH
}

// Many separate messages (IContentHasher::hash per span, ChunkValidator::validateChunks): packed 16-byte aligned into
// the pooled staging buffer by the pinned upload path, hashed by one launch per <= 1 GiB group.
yams_status_t yams_b200_sha256_many(void* self, const uint8_t* const* msgs, const size_t* lens, size_t n, uint8_t* digests) {
    YB_TRY
    (void)self;
    if (n == 0) return YAMS_OK;
    YB_ARG(msgs && lens && digests, "null argument");
    for (size_t i = 0; i < n; ++i) YB_ARG(msgs[i] || lens[i] == 0, "a message pointer is null");
    yams_cdc_config cfg;
    yams_b200_cdc_default_config(&cfg);
    IngestRes* r = nullptr;
    yams_status_t rc = acquire_res(&cfg, &r);
    if (rc != YAMS_OK) return rc;
    CdcStream& cs = r->cs;
    cudaStream_t st = cs.st;
    const uint64_t budget = batch_group_bytes();
    size_t i0 = 0;
    while (i0 < n && rc == YAMS_OK) {
        // group [i0, i1): packed size within the budget (a single larger message forms its own group)
        size_t i1 = i0;
        uint64_t tot = 0;
        std::vector<yams_chunk_desc> h;
        while (i1 < n && (i1 == i0 || tot + lens[i1] + 16 <= budget) && h.size() < 0xFFFFFFF0ull) {
            yams_chunk_desc d{};
            d.offset = tot;
            d.size = lens[i1];
            h.push_back(d);
            tot = (tot + lens[i1] + 15) & ~15ull;
            ++i1;
        }
        const size_t m = i1 - i0;
        rc = r->stage[0].reserve((size_t)tot + 256);
        if (rc == YAMS_OK) rc = cs.descs.reserve(m * sizeof(yams_chunk_desc));
        if (rc != YAMS_OK) break;
        uint8_t* d_buf = r->stage[0].as<uint8_t>();
        auto fill = [&](uint8_t* dst, uint64_t lo, uint64_t hi) {
            // first message whose packed range ends beyond lo
            size_t a = 0, b = m;
            while (a < b) {
                size_t mid = a + ((b - a) >> 1);
                if (h[mid].offset + h[mid].size <= lo) a = mid + 1; else b = mid;
            }
            uint64_t p = lo;
            for (size_t j = a; j < m && p < hi; ++j) {
                if (h[j].offset > p) {   // alignment padding
                    uint64_t e = std::min<uint64_t>(hi, h[j].offset);
                    memset(dst + (p - lo), 0, (size_t)(e - p));
                    p = e;
                }
                uint64_t e = std::min<uint64_t>(hi, h[j].offset + h[j].size);
                if (e > p) {
                    memcpy(dst + (p - lo), msgs[i0 + j] + (p - h[j].offset), (size_t)(e - p));
                    p = e;
                }
            }
            if (p < hi) memset(dst + (p - lo), 0, (size_t)(hi - p));
        };
        rc = staged_upload(r, d_buf, tot, fill, st);
        if (rc != YAMS_OK) break;
        cudaError_t e = cudaMemcpyAsync(cs.descs.p, h.data(), m * sizeof(yams_chunk_desc), cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) rc = cs.sha_order.reserve(sha256_order_ws_bytes(m));
        if (e == cudaSuccess && rc == YAMS_OK)
            rc = launch_sha256_chunks(d_buf, 0, cs.descs.as<yams_chunk_desc>(), 0, (uint32_t)m,
                                      reinterpret_cast<unsigned int*>(cs.scalars.as<uint64_t>() + 4), cs.dev->sm_count, st,
                                      cs.sha_order.as<uint32_t>(), tot);
        if (rc == YAMS_OK && e == cudaSuccess) e = cudaMemcpyAsync(h.data(), cs.descs.p, m * sizeof(yams_chunk_desc), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) {
            set_last_error("sha256_many failed: %s", cudaGetErrorString(e));
            rc = YAMS_ERR_INTERNAL;
        }
        if (rc == YAMS_OK)
            for (size_t j = 0; j < m; ++j) memcpy(digests + 32 * (i0 + j), h[j].digest, 32);
        i0 = i1;
    }
    release_res(r);
    return rc;
    YB_CATCH
}

yams_status_t yams_b200_dedup_stats(void* self, const yams_chunk_desc* chunks, size_t n, yams_dedup_stats* out) {
    YB_TRY
    (void)self;
    YB_ARG(out, "out is null");
    memset(out, 0, sizeof(*out));
    if (n == 0) return YAMS_OK;
    YB_ARG(chunks, "chunks is null");
    YB_ARG(n < (1ull << 31), "too many chunks");
    DeviceCtx* dev = nullptr;
    yams_status_t rc = ensure_device(&dev);
    if (rc != YAMS_OK) return rc;
    cudaStream_t st;
    YB_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    uint64_t slots = 1;
    while (slots < 2 * (uint64_t)n) slots <<= 1;   // load factor <= 0.5
    DevBuf d_descs, d_table, d_out;
    rc = d_descs.reserve(n * sizeof(yams_chunk_desc));
    if (rc == YAMS_OK) rc = d_table.reserve(slots * 4);
    if (rc == YAMS_OK) rc = d_out.reserve(sizeof(yams_dedup_stats));
    if (rc == YAMS_OK) {
        cudaMemcpyAsync(d_descs.p, chunks, n * sizeof(yams_chunk_desc), cudaMemcpyHostToDevice, st);
        cudaMemsetAsync(d_table.p, 0xFF, slots * 4, st);
        cudaMemsetAsync(d_out.p, 0, sizeof(yams_dedup_stats), st);
        rc = launch_dedup_stats(d_descs.as<yams_chunk_desc>(), (uint32_t)n, d_table.as<uint32_t>(), slots,
                                reinterpret_cast<unsigned long long*>(d_out.p), st);
        if (rc == YAMS_OK) {
            cudaError_t e = cudaMemcpyAsync(out, d_out.p, sizeof(yams_dedup_stats), cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess) e = cudaStreamSynchronize(st);
            if (e != cudaSuccess) {
                set_last_error("dedup_stats failed: %s", cudaGetErrorString(e));
                rc = YAMS_ERR_INTERNAL;
            }
        }
    }
    for (DevBuf* b : {&d_descs, &d_table, &d_out}) b->release();
    cudaStreamDestroy(st);
    return rc;
    YB_CATCH
}

yams_status_t yams_b200_ingest_last_timings(void* self, float out_ms[8]) {
    (void)self;
    YB_ARG(out_ms, "out_ms is null");
    for (int i = 0; i < 8; ++i) out_ms[i] = g_last_ms[i];
    return YAMS_OK;
}

// bench / test utility: SURVEY.md §8d byte stream generated straight into HBM
yams_status_t yams_b200_synth_bytes_device(uint64_t seed, uint64_t start, uint64_t n, uint8_t* d_out) {
    DeviceCtx* dev = nullptr;
    yams_status_t rc = ensure_device(&dev);
    if (rc != YAMS_OK) return rc;
    YB_ARG(d_out || n == 0, "d_out is null");
    rc = launch_synth_bytes(seed, start, n, d_out, dev->sm_count, 0);
    if (rc != YAMS_OK) return rc;
    YB_CUDA(cudaStreamSynchronize(0));
    return YAMS_OK;
}

}  // extern "C"
