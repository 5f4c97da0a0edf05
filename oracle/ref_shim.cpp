// ref_shim.cpp -- extern "C" door onto the UNMODIFIED reference implementation (test infrastructure).
//
// Compiled by oracle/Makefile together with the reference's own sources *where they lie* under
// /root/reference (src/chunking/{rabin,streaming}_chunker.cpp, src/crypto/sha256_hasher.cpp and the
// header-only third_party/sqlite-vec-cpp distances, src/manifest/manifest_manager.cpp, third_party/simeon/src/{pq,arch/avx2,
// arch/scalar}.cpp) into oracle/_ref/libyams_ref.so.  No reference
// source is copied into this repository; this file only *calls* the reference's public C++ API.
// Used to (1) pin oracle/yams_oracle.c, (2) generate tests/golden/*, (3) serve as bench.py's
// "reference" CPU arm when present.
#include <yams/chunking/chunker.h>
#include <yams/chunking/streaming_chunker.h>
#include <yams/crypto/hasher.h>

#include <sqlite-vec-cpp/distances/batch.hpp>
#include <sqlite-vec-cpp/distances/cosine.hpp>
#include <sqlite-vec-cpp/distances/l2.hpp>
#include <yams/manifest/manifest_manager.h>
#include <simeon/pq.hpp>
#include <simeon/simeon.hpp>
#include <sqlite-vec-cpp/utils/float16.hpp>

#include <algorithm>
#include <cstring>
#include <span>
#include <vector>

namespace {
yams::chunking::ChunkingConfig make_cfg(uint64_t window, uint64_t minc, uint64_t maxc,
                                        uint64_t poly, uint64_t mask) {
    yams::chunking::ChunkingConfig c;
    c.windowSize = window;
    c.minChunkSize = minc;
    c.maxChunkSize = maxc;
    c.targetChunkSize = std::max<uint64_t>(1, std::min<uint64_t>(maxc, 256 * 1024));
    c.polynomial = poly;
    c.chunkMask = mask;
    return c;
}

int hexval(char c) { return c <= '9' ? c - '0' : c - 'a' + 10; }
} // namespace

extern "C" {

// variant 0 = StreamingChunker::chunkData, 1 = RabinChunker::chunkDataLazy, 2 = RabinChunker::chunkData
size_t ref_chunk(const uint8_t* data, size_t n, uint64_t window, uint64_t minc, uint64_t maxc,
                 uint64_t poly, uint64_t mask, int variant, uint64_t* out_offsets,
                 uint64_t* out_sizes, uint8_t* out_digests /* cap*32, nullable */, size_t cap) {
    auto cfg = make_cfg(window, minc, maxc, poly, mask);
    std::span<const std::byte> span(reinterpret_cast<const std::byte*>(data), n);
    std::vector<yams::chunking::Chunk> chunks;
    if (variant == 0) {
        yams::chunking::StreamingChunker ch(cfg);
        chunks = ch.chunkData(span);
    } else {
        yams::chunking::RabinChunker ch(cfg);
        chunks = variant == 1 ? ch.chunkDataLazy(span) : ch.chunkData(span);
    }
    for (size_t i = 0; i < chunks.size() && i < cap; ++i) {
        out_offsets[i] = chunks[i].offset;
        out_sizes[i] = chunks[i].size;
        if (out_digests) {
            const std::string& h = chunks[i].hash;
            for (size_t b = 0; b < 32 && 2 * b + 1 < h.size(); ++b) {
                out_digests[32 * i + b] =
                    static_cast<uint8_t>((hexval(h[2 * b]) << 4) | hexval(h[2 * b + 1]));
            }
        }
    }
    return chunks.size();
}

// calculateDeduplication (rabin_chunker.cpp:224-239) over the chunks of the reference chunker itself.
// out[4] = totalSize, uniqueSize, chunkCount, uniqueChunks
void ref_dedup_stats(const uint8_t* data, size_t n, uint64_t window, uint64_t minc, uint64_t maxc, uint64_t poly,
                     uint64_t mask, int variant, uint64_t* out) {
    auto cfg = make_cfg(window, minc, maxc, poly, mask);
    std::span<const std::byte> span(reinterpret_cast<const std::byte*>(data), n);
    std::vector<yams::chunking::Chunk> chunks;
    if (variant == 0) {
        yams::chunking::StreamingChunker ch(cfg);
        chunks = ch.chunkData(span);
    } else {
        yams::chunking::RabinChunker ch(cfg);
        chunks = ch.chunkDataLazy(span);
    }
    yams::chunking::DeduplicationStats st = yams::chunking::calculateDeduplication(chunks);
    out[0] = st.totalSize;
    out[1] = st.uniqueSize;
    out[2] = st.chunkCount;
    out[3] = st.uniqueChunks;
}

// ManifestManager::createManifest (src/manifest/manifest_manager.cpp:411-436) over a chunk table given as raw digests:
// returns Manifest::checksum (calculateChecksum :705-730); *out_valid = validateManifest (:438-486) of that manifest.
uint32_t ref_manifest_checksum(const uint8_t* file_digest32, uint64_t file_size, const uint8_t* digests /* n x 32 */,
                               const uint64_t* offsets, const uint64_t* sizes, size_t n, int* out_valid) {
    auto hex = [](const uint8_t* d) {
        static const char* k = "0123456789abcdef";
        std::string s(64, '0');
        for (int i = 0; i < 32; ++i) { s[2 * i] = k[d[i] >> 4]; s[2 * i + 1] = k[d[i] & 15]; }
        return s;
    };
    yams::FileInfo info;
    info.hash = hex(file_digest32);
    info.size = file_size;
    info.originalName = "x";
    std::vector<yams::manifest::ChunkRef> refs(n);
    for (size_t i = 0; i < n; ++i) {
        refs[i].hash = hex(digests + 32 * i);
        refs[i].offset = offsets[i];
        refs[i].size = static_cast<uint32_t>(sizes[i]);
    }
    yams::manifest::ManifestManager mm{yams::manifest::ManifestManager::Config{}};
    auto res = mm.createManifest(info, refs);
    if (!res) { if (out_valid) *out_valid = -1; return 0; }
    if (out_valid) {
        auto v = mm.validateManifest(res.value());
        *out_valid = v ? (v.value() ? 1 : 0) : -1;
    }
    return res.value().checksum;
}

// ---- SimeonPqAdc pieces: simeon's own ProductQuantizer / PQInnerProductQuery (third_party/simeon/src/pq.cpp) ----
// encode_batch of ALREADY NORMALISED rows with imported codebooks [m][k][dim/m]
void ref_pq_encode(const float* codebooks, uint32_t dim, uint32_t m, uint32_t k, const float* vecs, uint32_t n, uint8_t* codes) {
    simeon::ProductQuantizer pq(simeon::PQConfig{.dim = dim, .m = m, .k = k});
    pq.import_codebooks(std::span<const float>(codebooks, (size_t)m * k * (dim / m)));
    pq.encode_batch(vecs, n, codes);
}
// lookup table + ADC inner products of one (already normalised) query over n codes
void ref_pq_scores(const float* codebooks, uint32_t dim, uint32_t m, uint32_t k, const float* query, const uint8_t* codes, size_t n,
                   float* out_scores, float* out_lut) {
    simeon::ProductQuantizer pq(simeon::PQConfig{.dim = dim, .m = m, .k = k});
    pq.import_codebooks(std::span<const float>(codebooks, (size_t)m * k * (dim / m)));
    simeon::PQInnerProductQuery qy(pq, query);
    for (size_t i = 0; i < n; ++i) out_scores[i] = qy.inner_product(codes + i * m);
    if (out_lut) std::memcpy(out_lut, qy.lut_ip().data(), qy.lut_ip().size() * sizeof(float));
}
// simeon::Encoder of the profile YAMS runs (src/simeon.cpp:75-93 simeon_v1_384_config with the given overrides): n texts -> n x output_dim
void ref_simeon_encode(uint32_t ngram_min, uint32_t ngram_max, uint32_t sketch_dim, uint32_t output_dim, uint64_t hash_seed,
                       uint64_t projection_seed, int l2_normalize, const char* const* texts, const size_t* lens, size_t n, float* out) {
    simeon::EncoderConfig cfg = simeon::simeon_v1_384_config();
    cfg.ngram_min = ngram_min;
    cfg.ngram_max = ngram_max;
    cfg.sketch_dim = sketch_dim;
    cfg.output_dim = output_dim;
    cfg.hash_seed = hash_seed;
    cfg.projection_seed = projection_seed;
    cfg.l2_normalize = l2_normalize != 0;
    simeon::Encoder enc(cfg);
    for (size_t i = 0; i < n; ++i) enc.encode(std::string_view(texts[i], lens[i]), out + i * enc.output_dim());
}
// The encoder YAMS builds when [embeddings.simeon] is left unconfigured (the "configurable" profile,
// /root/reference/src/embedding_simeon/simeon_embedding_backend.cpp:118-135 with the parse_* defaults at :18-47): library defaults
// plus the given ngram mode (0 CharOnly, 2 CharAndWord) and projection (1 AchlioptasSparse, 5 Fwht -- simeon.hpp enum values are
// passed through as integers).
void ref_simeon_encode_modes(int ngram_mode, int projection, uint32_t ngram_min, uint32_t ngram_max, uint32_t sketch_dim, uint32_t output_dim,
                             uint64_t hash_seed, uint64_t projection_seed, int l2_normalize, const char* const* texts, const size_t* lens,
                             size_t n, float* out) {
    simeon::EncoderConfig cfg;
    cfg.ngram_mode = static_cast<simeon::NGramMode>(ngram_mode);
    cfg.projection = static_cast<simeon::ProjectionMode>(projection);
    cfg.ngram_min = ngram_min;
    cfg.ngram_max = ngram_max;
    cfg.sketch_dim = sketch_dim;
    cfg.output_dim = output_dim;
    cfg.hash_seed = hash_seed;
    cfg.projection_seed = projection_seed;
    cfg.l2_normalize = l2_normalize != 0;
    simeon::Encoder enc(cfg);
    for (size_t i = 0; i < n; ++i) enc.encode(std::string_view(texts[i], lens[i]), out + i * enc.output_dim());
}
int ref_simeon_enum(const char* name) {
    const std::string s(name);
    if (s == "CharOnly") return (int)simeon::NGramMode::CharOnly;
    if (s == "CharAndWord") return (int)simeon::NGramMode::CharAndWord;
    if (s == "AchlioptasSparse") return (int)simeon::ProjectionMode::AchlioptasSparse;
    if (s == "Fwht") return (int)simeon::ProjectionMode::Fwht;
    return -1;
}

// Lloyd training of the reference (ProductQuantizer::train) -> codebooks, for realistic test indexes
void ref_pq_train(uint32_t dim, uint32_t m, uint32_t k, const float* training, uint32_t n_train, float* out_codebooks) {
    simeon::ProductQuantizer pq(simeon::PQConfig{.dim = dim, .m = m, .k = k});
    pq.train(training, n_train);
    auto cb = pq.codebooks();
    std::memcpy(out_codebooks, cb.data(), cb.size() * sizeof(float));
}

// SHA256Hasher::hash (static one-shot) -> 64-char lowercase hex + NUL
void ref_sha256_hex(const uint8_t* data, size_t n, char* out_hex65) {
    auto h = yams::crypto::SHA256Hasher::hash(
        std::span<const std::byte>(reinterpret_cast<const std::byte*>(data), n));
    std::memcpy(out_hex65, h.c_str(), h.size() + 1);
}

// init/update.../finalize with the given split points (streaming == one-shot contract)
void ref_sha256_stream_hex(const uint8_t* data, size_t n, size_t piece, char* out_hex65) {
    yams::crypto::SHA256Hasher hs;
    hs.init();
    if (piece == 0) piece = n ? n : 1;
    for (size_t off = 0; off < n; off += piece) {
        size_t len = std::min(piece, n - off);
        hs.update(std::span<const std::byte>(reinterpret_cast<const std::byte*>(data + off), len));
    }
    auto h = hs.finalize();
    std::memcpy(out_hex65, h.c_str(), h.size() + 1);
}

void ref_sha256_batch(const uint8_t* base, const uint64_t* offsets, const uint64_t* sizes, size_t n,
                      uint8_t* out_digests) {
#pragma omp parallel for schedule(dynamic, 16)
    for (long i = 0; i < static_cast<long>(n); ++i) {
        auto h = yams::crypto::SHA256Hasher::hash(std::span<const std::byte>(
            reinterpret_cast<const std::byte*>(base + offsets[i]), sizes[i]));
        for (size_t b = 0; b < 32; ++b) {
            out_digests[32 * i + b] =
                static_cast<uint8_t>((hexval(h[2 * b]) << 4) | hexval(h[2 * b + 1]));
        }
    }
}

float ref_l2_distance(const float* a, const float* b, size_t d) {
    return sqlite_vec_cpp::distances::l2_distance(std::span<const float>(a, d),
                                                  std::span<const float>(b, d));
}

float ref_cosine_distance(const float* a, const float* b, size_t d) {
    return sqlite_vec_cpp::distances::cosine_distance(std::span<const float>(a, d),
                                                      std::span<const float>(b, d));
}

// batch_distance_contiguous + partial_sort top-k (distances/batch.hpp:47-92), metric 0 cosine 1 L2
size_t ref_batch_top_k(const float* query, const float* rows, size_t n, size_t d, int metric,
                       size_t k, uint64_t* out_idx, float* out_dist) {
    using namespace sqlite_vec_cpp::distances;
    std::vector<float> dist;
    if (metric == 0) {
        dist = batch::batch_distance_contiguous(std::span<const float>(query, d),
                                                std::span<const float>(rows, n * d), n, d,
                                                CosineMetric<float>{});
    } else {
        dist = batch::batch_distance_contiguous(std::span<const float>(query, d),
                                                std::span<const float>(rows, n * d), n, d,
                                                L2Metric<float>{});
    }
    std::vector<size_t> idx(n);
    for (size_t i = 0; i < n; ++i) idx[i] = i;
    size_t m = std::min(k, n);
    std::partial_sort(idx.begin(), idx.begin() + m, idx.end(),
                      [&dist](size_t a, size_t b) { return dist[a] < dist[b]; });
    for (size_t i = 0; i < m; ++i) {
        out_idx[i] = idx[i];
        if (out_dist) out_dist[i] = dist[idx[i]];
    }
    return m;
}

// Q queries, OpenMP over queries: the sqlite-vec-cpp SIMD arm of BASELINE.md §3 ("knn-ref-simd").
// rows fp32. Writes Q x k indices/distances.
void ref_batch_top_k_queries(const float* queries, size_t nq, const float* rows, size_t n, size_t d,
                             int metric, size_t k, uint64_t* out_idx, float* out_dist) {
#pragma omp parallel for schedule(dynamic, 1)
    for (long q = 0; q < static_cast<long>(nq); ++q) {
        ref_batch_top_k(queries + q * d, rows, n, d, metric, k, out_idx + q * k,
                        out_dist ? out_dist + q * k : nullptr);
    }
}

uint16_t ref_f16_from_float(float f) { return sqlite_vec_cpp::utils::float16_t::from_float(f).bits; }
float ref_f16_to_float(uint16_t h) { return sqlite_vec_cpp::utils::float16_t(h).to_float(); }

} // extern "C"
