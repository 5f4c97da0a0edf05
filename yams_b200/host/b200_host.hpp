// b200_host.hpp -- C++ host-side mirror of the reference's operator interfaces for the hot path, written
// ONLY against the C ABI (include/yams_b200.h).  A YAMS maintainer would derive these from the real
// IChunker / IContentHasher / IVectorStore; here they are self-contained so they build without
// the YAMS tree, keeping the reference's member names, argument meaning and error behaviour:
//   IChunker         /root/reference/include/yams/chunking/chunker.h:65-92   (Chunk: :18-31, ChunkingConfig: :44-51)
//   IContentHasher   /root/reference/include/yams/crypto/hasher.h:14-46      (hex: src/crypto/sha256_hasher.cpp:19-30)
//   IVectorStore     /root/reference/include/yams/vector/vector_store.h:44-53, :121-138
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <filesystem>
#include <span>
#include <stdexcept>
#include <algorithm>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/yams_b200.h"

namespace yams_b200::host {

using Hash = std::string;  // core/types.h:17

struct Chunk {  // chunker.h:18-31
    std::vector<std::byte> data;
    Hash hash;
    size_t offset = 0;
    size_t size = 0;
};

struct ChunkingConfig {  // chunker.h:44-51 (defaults: core/types.h:280-285)
    size_t windowSize = 48;
    size_t minChunkSize = 16 * 1024;
    size_t targetChunkSize = 256 * 1024;
    size_t maxChunkSize = 1024 * 1024;
    uint64_t polynomial = 0x3DA3358B4DC173ULL;
    uint64_t chunkMask = 0x1FFF;
};

inline std::string bytesToHex(const uint8_t* d, size_t n) {  // sha256_hasher.cpp:19-30
    static constexpr char kHex[] = "0123456789abcdef";
    std::string out(n * 2, '\0');
    for (size_t i = 0; i < n; ++i) {
        out[2 * i] = kHex[(d[i] >> 4) & 0xF];
        out[2 * i + 1] = kHex[d[i] & 0xF];
    }
    return out;
}

inline void throw_status(const char* what, yams_status_t st) {
    throw std::runtime_error(std::string(what) + " failed (status " + std::to_string(st) + "): " + yams_b200_last_error());
}

// IChunker over the GPU.  variant: YAMS_CDC_STREAMING mirrors StreamingChunker (what ContentStore uses,
// src/api/content_store_builder.cpp:152-165), YAMS_CDC_RABIN mirrors RabinChunker.
class B200Chunker {
public:
    explicit B200Chunker(ChunkingConfig config = {}, int variant = YAMS_CDC_STREAMING) : config_(config), variant_(variant) {}
    const ChunkingConfig& getConfig() const { return config_; }

    std::vector<Chunk> chunkData(std::span<const std::byte> data) { return run(data, false); }
    std::vector<Chunk> chunkDataLazy(std::span<const std::byte> data) { return run(data, true); }

    // streaming_chunker.cpp:71-90 / streaming_chunker.h:78-121: 64 KiB reads fed to a session
    std::vector<Chunk> chunkFile(const std::filesystem::path& path) {
        std::ifstream file(path, std::ios::binary);
        if (!file) throw std::runtime_error("Failed to open file: " + path.string());  // rabin_chunker.cpp:156-158
        yams_cdc_config c = cfg();
        yams_b200_ingest* s = nullptr;
        yams_status_t st = yams_b200_ingest_open(nullptr, &c, &s);
        if (st != YAMS_OK) throw_status("ingest_open", st);
        std::vector<Chunk> chunks;
        std::vector<std::byte> whole;  // Chunk::data is filled from the bytes read (non-lazy contract of chunkFile)
        std::vector<char> buf(1 << 22);
        try {
            for (;;) {
                file.clear();
                file.read(buf.data(), (std::streamsize)buf.size());
                std::streamsize got = file.gcount();
                if (got <= 0) break;
                whole.insert(whole.end(), reinterpret_cast<std::byte*>(buf.data()), reinterpret_cast<std::byte*>(buf.data()) + got);
                yams_chunk_desc* d = nullptr;
                size_t n = 0;
                st = yams_b200_ingest_feed(s, reinterpret_cast<const uint8_t*>(buf.data()), (size_t)got, &d, &n);
                if (st != YAMS_OK) throw_status("ingest_feed", st);
                append(chunks, d, n, whole, false);
                yams_b200_free_chunks(nullptr, d, n);
            }
            yams_chunk_desc* d = nullptr;
            size_t n = 0;
            st = yams_b200_ingest_finish(s, &d, &n);
            if (st != YAMS_OK) throw_status("ingest_finish", st);
            append(chunks, d, n, whole, false);
            yams_b200_free_chunks(nullptr, d, n);
        } catch (...) {
            yams_b200_ingest_close(s);
            throw;
        }
        yams_b200_ingest_close(s);
        return chunks;
    }

private:
    yams_cdc_config cfg() const {
        yams_cdc_config c{};
        c.window_size = config_.windowSize;
        c.min_chunk = config_.minChunkSize;
        c.max_chunk = config_.maxChunkSize;
        c.polynomial = config_.polynomial;
        c.mask = config_.chunkMask;
        c.variant = variant_;
        return c;
    }
    static void append(std::vector<Chunk>& out, const yams_chunk_desc* d, size_t n, const std::vector<std::byte>& bytes, bool lazy) {
        for (size_t i = 0; i < n; ++i) {
            Chunk ch;
            ch.offset = (size_t)d[i].offset;
            ch.size = (size_t)d[i].size;
            ch.hash = bytesToHex(d[i].digest, 32);
            if (!lazy) ch.data.assign(bytes.begin() + (ptrdiff_t)ch.offset, bytes.begin() + (ptrdiff_t)(ch.offset + ch.size));
            out.push_back(std::move(ch));
        }
    }
    std::vector<Chunk> run(std::span<const std::byte> data, bool lazy) {
        yams_cdc_config c = cfg();
        yams_chunk_desc* d = nullptr;
        size_t n = 0;
        yams_status_t st = yams_b200_chunk_and_hash(nullptr, reinterpret_cast<const uint8_t*>(data.data()), data.size(), &c, &d, &n);
        if (st != YAMS_OK) throw_status("chunk_and_hash", st);
        std::vector<Chunk> out;
        out.reserve(n);
        for (size_t i = 0; i < n; ++i) {
            Chunk ch;
            ch.offset = (size_t)d[i].offset;
            ch.size = (size_t)d[i].size;
            ch.hash = bytesToHex(d[i].digest, 32);
            if (!lazy) ch.data.assign(data.begin() + (ptrdiff_t)ch.offset, data.begin() + (ptrdiff_t)(ch.offset + ch.size));
            out.push_back(std::move(ch));
        }
        yams_b200_free_chunks(nullptr, d, n);
        return out;
    }
    ChunkingConfig config_;
    int variant_;
};

// IContentHasher: init / update / finalize accumulate on the host side (the digest of ONE message is a serial
// chain; the GPU earns its keep when many messages are hashed at once -> hashMany).
class B200ContentHasher {
public:
    void init() { buf_.clear(); }
    void update(std::span<const std::byte> data) { buf_.insert(buf_.end(), data.begin(), data.end()); }
    std::string finalize() {
        std::string h = hash(std::span<const std::byte>(buf_.data(), buf_.size()));
        buf_.clear();  // sha256_hasher.cpp:104: finalize re-initialises
        return h;
    }
    static std::string hash(std::span<const std::byte> data) {  // sha256_hasher.cpp:167-195
        uint64_t off = 0, sz = data.size();
        uint8_t dg[32];
        yams_status_t st = yams_b200_sha256_batch(nullptr, reinterpret_cast<const uint8_t*>(data.data()), data.size(), &off, &sz, 1, dg);
        if (st != YAMS_OK) throw_status("sha256_batch", st);
        return bytesToHex(dg, 32);
    }
    static std::vector<std::string> hashMany(std::span<const std::byte> base, const std::vector<uint64_t>& offsets,
                                             const std::vector<uint64_t>& sizes) {
        std::vector<uint8_t> dg(offsets.size() * 32);
        yams_status_t st = yams_b200_sha256_batch(nullptr, reinterpret_cast<const uint8_t*>(base.data()), base.size(), offsets.data(),
                                                  sizes.data(), offsets.size(), dg.data());
        if (st != YAMS_OK) throw_status("sha256_batch", st);
        std::vector<std::string> out(offsets.size());
        for (size_t i = 0; i < offsets.size(); ++i) out[i] = bytesToHex(dg.data() + 32 * i, 32);
        return out;
    }

private:
    std::vector<std::byte> buf_;
};

// The part of IVectorStore the hot path covers.  Records live in SQLite in the real backend; here a record is
// (rowid, chunk_id) and `relevance_score` -- enough to express the tie-break contract of
// sqlite_vec_backend.cpp:4218-4223 (similarity desc, then chunk_id asc).
struct VectorHit {
    int64_t rowid = -1;
    std::string chunk_id;
    float relevance_score = 0.f;
};

class B200VectorStore {
public:
    B200VectorStore(uint32_t dim, int dtype = YAMS_B200_F32, int metric = YAMS_B200_COSINE) : dim_(dim) {
        yams_status_t st = yams_b200_corpus_create(nullptr, dim, dtype, metric, 0, &c_);
        if (st != YAMS_OK) throw_status("corpus_create", st);
    }
    ~B200VectorStore() { yams_b200_corpus_destroy(c_); }
    B200VectorStore(const B200VectorStore&) = delete;
    B200VectorStore& operator=(const B200VectorStore&) = delete;

    // insertVectorsBatch: rows are fp32 (the reference BLOB layout); rowids ascending
    void insertVectorsBatch(const std::vector<float>& rows, const std::vector<int64_t>& rowids, const std::vector<std::string>& chunk_ids) {
        yams_status_t st = yams_b200_corpus_append(c_, rows.data(), rowids.size(), rowids.data());
        if (st != YAMS_OK) throw_status("corpus_append", st);
        for (size_t i = 0; i < rowids.size(); ++i) chunk_of_[rowids[i]] = chunk_ids[i];
    }

    // searchSimilar (vector_store.h:44-49): InvalidArgument (std::invalid_argument here) for a non-finite or
    // zero-norm query; k == 0 -> empty.
    std::vector<VectorHit> searchSimilar(const std::vector<float>& query, size_t k, float similarity_threshold = 0.0f) {
        auto all = searchSimilarBatch({query}, k, similarity_threshold);
        return all.empty() ? std::vector<VectorHit>{} : all[0];
    }

    std::vector<std::vector<VectorHit>> searchSimilarBatch(const std::vector<std::vector<float>>& queries, size_t k,
                                                           float similarity_threshold = 0.0f) {
        std::vector<std::vector<VectorHit>> out(queries.size());
        if (queries.empty() || k == 0) return out;
        std::vector<float> flat;
        for (const auto& q : queries) {
            if (q.size() != dim_) throw std::invalid_argument("All query embeddings must have the same dimension");  // :1620-1627
            flat.insert(flat.end(), q.begin(), q.end());
        }
        // ask for slack so that an equal-score run straddling k can be re-ordered by chunk_id on the host
        size_t kk = k + 8;
        for (int attempt = 0; attempt < 4; ++attempt) {
            std::vector<int64_t> rid(queries.size() * kk);
            std::vector<float> sc(queries.size() * kk);
            std::vector<uint32_t> cnt(queries.size());
            std::vector<uint64_t> flg(queries.size());
            yams_status_t st = yams_b200_search(c_, flat.data(), (uint32_t)queries.size(), (uint32_t)kk, similarity_threshold, nullptr,
                                                nullptr, rid.data(), sc.data(), cnt.data(), flg.data());
            if (st == YAMS_ERR_INVALID_ARG) throw std::invalid_argument("Exact vector search requires a finite, non-zero query embedding");
            if (st != YAMS_OK) throw_status("search", st);
            bool need_more = false;
            for (size_t q = 0; q < queries.size(); ++q)
                if ((flg[q] & YAMS_B200_FLAG_TIE_AT_K) && cnt[q] == kk && kk < 768) need_more = true;
            if (need_more) { kk = std::min<size_t>(768, kk * 2); continue; }
            for (size_t q = 0; q < queries.size(); ++q) {
                std::vector<VectorHit> hits(cnt[q]);
                for (uint32_t i = 0; i < cnt[q]; ++i) {
                    hits[i].rowid = rid[q * kk + i];
                    hits[i].relevance_score = sc[q * kk + i];
                    auto it = chunk_of_.find(hits[i].rowid);
                    hits[i].chunk_id = it == chunk_of_.end() ? std::string() : it->second;
                }
                // sqlite_vec_backend.cpp:4218-4223: similarity desc, then chunk_id asc
                std::stable_sort(hits.begin(), hits.end(), [](const VectorHit& a, const VectorHit& b) {
                    if (a.relevance_score != b.relevance_score) return a.relevance_score > b.relevance_score;
                    return a.chunk_id < b.chunk_id;
                });
                if (hits.size() > k) hits.resize(k);
                out[q] = std::move(hits);
            }
            return out;
        }
        throw std::runtime_error("too many equal-score rows at the k boundary");
    }

    size_t size() const {
        uint64_t n = 0;
        yams_b200_corpus_size(c_, &n);
        return (size_t)n;
    }

private:
    yams_b200_corpus* c_ = nullptr;
    uint32_t dim_;
    std::unordered_map<int64_t, std::string> chunk_of_;
};

}  // namespace yams_b200::host
