// GPU test of the C++ host adapters (yams_b200/host/b200_host.hpp) through the C ABI.  The expectations mirror the
// reference's own unit tests: chunk invariants + per-chunk hash of the pattern stream
// (/root/reference/tests/unit/chunking/chunking_test.cpp:55-62,146-216), lazy == full (:588-609), SHA-256 KATs
// (/root/reference/tests/unit/crypto/crypto_test.cpp:92-99), exact-scan contract incl. chunk_id tie-break
// (/root/reference/tests/unit/vector/vector_smoke_catch2_test.cpp:188-340).  Prints "ALL OK" on success.
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "../../yams_b200/host/b200_host.hpp"

using namespace yams_b200::host;

#define CHECK(cond)                                                                  \
    do {                                                                             \
        if (!(cond)) {                                                               \
            std::fprintf(stderr, "CHECK failed %s:%d: %s\n", __FILE__, __LINE__, #cond); \
            std::exit(1);                                                            \
        }                                                                            \
    } while (0)

int main() {
    if (yams_plugin_init("{}", nullptr) != YAMS_PLUGIN_OK) {
        std::fprintf(stderr, "plugin init failed: %s\n", yams_b200_last_error());
        return 2;
    }
    // ---- hasher KATs ----
    auto bytes = [](const char* s) { return std::span<const std::byte>(reinterpret_cast<const std::byte*>(s), std::strlen(s)); };
    CHECK(B200ContentHasher::hash(bytes("")) == "e3b0c44298fc1c149afbf4c8996fb92427ae41e4649b934ca495991b7852b855");
    CHECK(B200ContentHasher::hash(bytes("abc")) == "ba7816bf8f01cfea414140de5dae2223b00361a396177a9cb410ff61f20015ad");
    B200ContentHasher h;
    h.init();
    h.update(bytes("Hello "));
    h.update(bytes("World"));
    CHECK(h.finalize() == "a591a6d40bf420404a011733cfb7b190d62c65bf0bcda32b57b277d9ad9f146e");
    {   // hashFile / hashFiles (sha256_hasher.cpp:111-150): file contents, empty file, missing file throws
        namespace fs = std::filesystem;
        const fs::path dir = fs::temp_directory_path() / "yams_b200_hashfile_test";
        fs::create_directories(dir);
        { std::ofstream(dir / "a.txt", std::ios::binary) << "Hello World"; }
        { std::ofstream(dir / "empty.bin", std::ios::binary); }
        CHECK(B200ContentHasher::hashFile(dir / "a.txt") == "a591a6d40bf420404a011733cfb7b190d62c65bf0bcda32b57b277d9ad9f146e");
        auto many = B200ContentHasher::hashFiles({dir / "empty.bin", dir / "a.txt"});
        CHECK(many.size() == 2 && many[0] == "e3b0c44298fc1c149afbf4c8996fb92427ae41e4649b934ca495991b7852b855" && many[1] == B200ContentHasher::hashFile(dir / "a.txt"));
        bool threw = false;
        try { B200ContentHasher::hashFile(dir / "does_not_exist"); } catch (const std::runtime_error&) { threw = true; }
        CHECK(threw);
        fs::remove_all(dir);
    }
    // ---- chunker: pattern data, chunk invariants, per-chunk hash, lazy == full ----
    std::vector<std::byte> data(3 * 1024 * 1024 + 123);
    for (size_t i = 0; i < data.size(); ++i) data[i] = (std::byte)((i * 1315423911u + 0x9E3779B9u) & 0xFF);
    ChunkingConfig cfg;
    cfg.minChunkSize = 4096;
    cfg.maxChunkSize = 65536;
    B200Chunker chunker(cfg);
    auto full = chunker.chunkData(data);
    auto lazy = chunker.chunkDataLazy(data);
    CHECK(!full.empty() && full.size() == lazy.size());
    size_t pos = 0;
    for (size_t i = 0; i < full.size(); ++i) {
        CHECK(full[i].offset == pos && lazy[i].offset == pos && full[i].size == lazy[i].size && full[i].hash == lazy[i].hash);
        CHECK(lazy[i].data.empty() && full[i].data.size() == full[i].size);
        if (i + 1 < full.size()) CHECK(full[i].size >= cfg.minChunkSize && full[i].size <= cfg.maxChunkSize);
        CHECK(full[i].hash == B200ContentHasher::hash(std::span<const std::byte>(data.data() + pos, full[i].size)));
        CHECK(std::memcmp(full[i].data.data(), data.data() + pos, full[i].size) == 0);
        pos += full[i].size;
    }
    CHECK(pos == data.size());
    // chunkFile == chunkData (streamed in 4 MiB reads)
    auto path = std::filesystem::temp_directory_path() / "yams_b200_host_test.bin";
    {
        std::ofstream f(path, std::ios::binary);
        f.write(reinterpret_cast<const char*>(data.data()), (std::streamsize)data.size());
    }
    auto filed = chunker.chunkFile(path);
    CHECK(filed.size() == full.size());
    for (size_t i = 0; i < full.size(); ++i) CHECK(filed[i].offset == full[i].offset && filed[i].hash == full[i].hash && filed[i].data == full[i].data);
    std::filesystem::remove(path);
    bool threw = false;
    try { chunker.chunkFile("/nonexistent/yams_b200"); } catch (const std::runtime_error&) { threw = true; }
    CHECK(threw);
    // ---- vector store: top-1, threshold, invalid query, deterministic chunk_id tie-break ----
    B200VectorStore vs(4);
    std::vector<float> rows = {1, 0, 0, 0, /**/ 0.9f, 0.1f, 0, 0, /**/ 2, 0, 0, 0, /**/ -1, 0, 0, 0};
    vs.insertVectorsBatch(rows, {1, 2, 3, 4}, {"zz", "mm", "aa", "bb"});   // rows 1 and 3 tie at similarity 1.0
    auto hits = vs.searchSimilar({1, 0, 0, 0}, 1, -1.0f);
    CHECK(hits.size() == 1 && hits[0].chunk_id == "aa" && hits[0].relevance_score == 1.0f);   // chunk_id order, not rowid
    hits = vs.searchSimilar({1, 0, 0, 0}, 10, 0.5f);
    CHECK(hits.size() == 3 && hits[0].chunk_id == "aa" && hits[1].chunk_id == "zz" && hits[2].chunk_id == "mm");
    CHECK(vs.searchSimilar({1, 0, 0, 0}, 0).empty());
    threw = false;
    try { vs.searchSimilar({0, 0, 0, 0}, 3); } catch (const std::invalid_argument&) { threw = true; }
    CHECK(threw);
    threw = false;
    try { vs.searchSimilar({NAN, 0, 0, 0}, 3); } catch (const std::invalid_argument&) { threw = true; }
    CHECK(threw);
    CHECK(vs.size() == 4);
    // ---- candidate documents, all-matching, deletes (vector_store.h:39-40,121-138) ----
    B200VectorStore cs(4);
    std::vector<float> crow = {1, 0, 0, 0, /**/ 0.9f, 0.1f, 0, 0, /**/ 0.5f, 0.5f, 0, 0, /**/ 0, 1, 0, 0, /**/ 0.8f, 0.2f, 0, 0};
    cs.insertVectorsBatch(crow, {10, 11, 12, 13, 14}, {"c0", "c1", "c2", "c3", "c4"}, {"docA", "docB", "docB", "docC", "docA"});
    auto ch = cs.searchExactCandidates({1, 0, 0, 0}, 2, -1.0f, {"docB", "docC"});
    CHECK(ch.size() == 2 && ch[0].chunk_id == "c1" && ch[1].chunk_id == "c2");
    CHECK(cs.searchExactCandidates({1, 0, 0, 0}, 3, -1.0f, {"nope"}).empty());
    auto am = cs.searchAllExactCandidateRows({1, 0, 0, 0}, 0.5f, {"docA", "docB", "docC"});
    CHECK(am.size() == 4 && am[0].chunk_id == "c0" && am[1].chunk_id == "c1" && am[2].chunk_id == "c4" && am[3].chunk_id == "c2");
    cs.deleteVectorsByDocument("docB");
    CHECK(cs.size() == 3);
    hits = cs.searchSimilar({1, 0, 0, 0}, 10, -1.0f);
    CHECK(hits.size() == 3 && hits[0].chunk_id == "c0" && hits[1].chunk_id == "c4" && hits[2].chunk_id == "c3");
    cs.deleteVector("c0");
    hits = cs.searchSimilar({1, 0, 0, 0}, 1, -1.0f);
    CHECK(cs.size() == 2 && hits.size() == 1 && hits[0].chunk_id == "c4");
    // ---- every argument of IVectorStore::searchSimilar (vector_store.h:44-49) + VectorSearchDiagnostics ----
    {
        B200VectorStore ms(4);
        std::vector<float> mrow = {1, 0, 0, 0, /**/ 0.9f, 0.1f, 0, 0, /**/ 0.5f, 0.5f, 0, 0, /**/ 0, 0, 0, 0, /**/ 0.8f, 0.2f, 0, 0, /**/ 0.7f, 0.3f, 0, 0};
        ms.insertVectorsBatch(mrow, {1, 2, 3, 4, 5, 6}, {"m0", "m1", "m2", "m3", "m4", "m5"}, {"docA", "docA", "docB", "docB", "docB", "docC"},
                              {{{"lang", "en"}}, {{"lang", "de"}}, {{"lang", "en"}}, {{"lang", "en"}}, {{"lang", "en"}, {"kind", "code"}}, {}});
        VectorSearchDiagnostics dg;
        auto h = ms.searchSimilar({1, 0, 0, 0}, 10, -1.0f, std::string("docB"), {}, {}, &dg);   // WHERE document_hash = 'docB'
        CHECK(h.size() == 2 && h[0].chunk_id == "m4" && h[1].chunk_id == "m2");                 // m3 is a zero row: skipped
        CHECK(dg.usedExactScan && dg.rowsVisitedObserved && dg.rowsVisited == 3 && dg.exactDistanceEvaluations == 2 && dg.returnedRows == 2);
        dg = {};
        h = ms.searchSimilar({1, 0, 0, 0}, 10, -1.0f, std::nullopt, {}, {{"lang", "en"}}, &dg);     // metadata filter over the whole table
        CHECK(h.size() == 3 && h[0].chunk_id == "m0" && h[1].chunk_id == "m4" && h[2].chunk_id == "m2");
        CHECK(dg.rowsVisited == 6 && dg.exactDistanceEvaluations == 3 && dg.returnedRows == 3);
        h = ms.searchSimilar({1, 0, 0, 0}, 10, -1.0f, std::nullopt, {"docA", "docB"}, {{"lang", "en"}, {"kind", "code"}});
        CHECK(h.size() == 1 && h[0].chunk_id == "m4");
        h = ms.searchSimilar({1, 0, 0, 0}, 10, -1.0f, std::string("docC"), {"docA"}, {});          // document_hash outside the candidate set
        CHECK(h.empty());
        dg = {};
        h = ms.searchSimilar({1, 0, 0, 0}, 2, -1.0f, std::nullopt, {}, {}, &dg);
        CHECK(h.size() == 2 && h[0].chunk_id == "m0" && h[1].chunk_id == "m1" && dg.rowsVisited == 6 && dg.exactDistanceEvaluations == 5);
    }
    // ---- fp16 store (BASELINE config C2's layout) fed with the fp32 rows the reference holds: same neighbours as the fp32 store ----
    {
        const uint32_t D = 64, N = 3000;
        std::vector<float> big(N * D);
        uint64_t sd = 99;
        for (auto& v : big) { sd = sd * 6364136223846793005ULL + 1442695040888963407ULL; v = (float)((sd >> 40) & 0xFFFF) / 65536.0f - 0.5f; }
        std::vector<int64_t> ids(N);
        std::vector<std::string> cids(N);
        for (uint32_t i = 0; i < N; ++i) { ids[i] = 100 + i; cids[i] = "k" + std::to_string(i); }
        B200VectorStore s32(D, YAMS_B200_F32), s16(D, YAMS_B200_F16);
        s32.insertVectorsBatch(big, ids, cids);
        s16.insertVectorsBatch(big, ids, cids);
        std::vector<float> q(big.begin() + 17 * D, big.begin() + 18 * D);
        auto a32 = s32.searchSimilar(q, 5, -1.0f), a16 = s16.searchSimilar(q, 5, -1.0f);
        CHECK(a32.size() == 5 && a16.size() == 5 && a32[0].chunk_id == "k17" && a16[0].chunk_id == "k17");
        for (int i = 0; i < 5; ++i) CHECK(std::fabs(a32[i].relevance_score - a16[i].relevance_score) < 2e-3f);   // truncated halves: ranks 2.. may swap, scores cannot move
    }
    // ---- an equal-score run longer than the slack: widened until the chunk_id order is right (:4218-4223) ----
    {
        B200VectorStore ts(4);
        const int T = 40;
        std::vector<float> trow;
        std::vector<int64_t> tid;
        std::vector<std::string> tcid;
        for (int i = 0; i < T; ++i) {
            trow.insert(trow.end(), {2.0f, 0, 0, 0});
            tid.push_back(i + 1);
            char buf[16];
            snprintf(buf, sizeof buf, "t%03d", T - i);      // chunk_id order is the reverse of the rowid order
            tcid.push_back(buf);
        }
        ts.insertVectorsBatch(trow, tid, tcid);
        auto th = ts.searchSimilar({1, 0, 0, 0}, 3, -1.0f);
        CHECK(th.size() == 3 && th[0].chunk_id == "t001" && th[1].chunk_id == "t002" && th[2].chunk_id == "t003");
        bool k_threw = false;
        try { ts.searchSimilar({1, 0, 0, 0}, 4000, -1.0f); } catch (const std::invalid_argument& e) { k_threw = std::string(e.what()).find("3072") != std::string::npos; }
        CHECK(k_threw);
    }
    CHECK(computeCosineSimilarity({1, 2, 3}, {1, 2, 3, 4}) == 0.0);
    CHECK(std::fabs(computeCosineSimilarity({1, 0}, {1, 1}) - 0.70710678118654757) < 1e-15);
    // ---- many buffers in one device pass == one chunkDataLazy per buffer ----
    {
        std::vector<std::span<const std::byte>> many;
        const size_t cuts[] = {0, 1, 100000, 100000, 1500000, data.size()};
        for (size_t i = 0; i + 1 < sizeof(cuts) / sizeof(cuts[0]); ++i) many.emplace_back(data.data() + cuts[i], cuts[i + 1] - cuts[i]);
        auto res = chunker.chunkManyLazy(many);
        CHECK(res.size() == many.size());
        for (size_t i = 0; i < many.size(); ++i) {
            auto want = chunker.chunkDataLazy(many[i]);
            CHECK(res[i].size() == want.size());
            for (size_t j = 0; j < want.size(); ++j) CHECK(res[i][j].offset == want[j].offset && res[i][j].size == want[j].size && res[i][j].hash == want[j].hash);
        }
    }
    // ---- validateChunks: re-hash stored chunks, flag the corrupted one ----
    {
        std::vector<std::pair<std::span<const std::byte>, std::string>> stored;
        for (size_t i = 0; i < std::min<size_t>(full.size(), 40); ++i) stored.emplace_back(std::span<const std::byte>(full[i].data), full[i].hash);
        std::vector<std::byte> bad(full[3].data);
        bad[bad.size() / 2] ^= std::byte{0x01};
        stored[3].first = std::span<const std::byte>(bad);
        auto rep = validateChunks(stored);
        CHECK(rep.size() == stored.size());
        for (size_t i = 0; i < rep.size(); ++i) CHECK(rep[i].isValid == (i != 3));
        CHECK(rep[3].errorMessage.rfind("Hash mismatch", 0) == 0);
    }
    // ---- dedup accounting + the exists/store loop over repeated content ----
    {
        std::vector<std::byte> twice(data.begin(), data.end());
        twice.insert(twice.end(), data.begin(), data.end());
        auto cc = chunker.chunkDataLazy(twice);
        auto ds = calculateDeduplication(cc);
        CHECK(ds.chunkCount == cc.size() && ds.totalSize == twice.size() && ds.uniqueChunks < ds.chunkCount && ds.getRatio() > 0.3);
        B200ChunkIndex index;
        auto acc = index.addChunks(cc);
        CHECK(acc.bytesStored == ds.uniqueSize && acc.bytesStored + acc.bytesDeduped == twice.size() && index.size() == ds.uniqueChunks);
        auto again = index.addChunks(cc);
        CHECK(again.bytesStored == 0 && again.bytesDeduped == twice.size());
    }
    // ---- manifest of the chunked buffer: refs mirror the table, the checksum is CRC-32 of the reference's text (zlib polynomial) ----
    {
        auto hasher_hex = B200ContentHasher::hash(std::span<const std::byte>(data));
        auto m = createManifest(hasher_hex, data.size(), full);
        CHECK(m.valid && m.chunks.size() == full.size() && m.chunks[3].hash == full[3].hash && m.chunks[3].offset == full[3].offset);
        std::string text = hasher_hex + std::to_string(data.size());
        for (const auto& c : full) text += c.hash + std::to_string(c.offset) + std::to_string((uint32_t)c.size);
        uint32_t crc = 0xFFFFFFFFu;   // src/manifest/manifest_manager.cpp:705-730, the loop itself
        for (char ch : text) { crc ^= (uint32_t)ch; for (int b = 0; b < 8; ++b) crc = (crc >> 1) ^ (0xEDB88320u * (crc & 1u)); }
        CHECK(m.checksum == ~crc);
        auto broken = full;
        broken[2].offset += 1;
        CHECK(!createManifest(hasher_hex, data.size(), broken).valid);
    }
    // ---- Simeon backend adapter: unit-norm embeddings, batch == single, the text itself is its own nearest neighbour ----
    {
        B200SimeonBackend sb;
        std::vector<std::string> texts = {"content addressed storage", "vector similarity search", "", "content addressed storage!"};
        auto embs = sb.generateEmbeddings(texts);
        CHECK(sb.getEmbeddingDimension() == 384 && embs.size() == 4 && embs[0].size() == 384);
        CHECK(sb.generateEmbedding(texts[1]) == embs[1]);
        double n0 = 0, dot03 = 0, dot01 = 0, n2 = 0;
        for (int i = 0; i < 384; ++i) { n0 += embs[0][i] * embs[0][i]; dot03 += embs[0][i] * embs[3][i]; dot01 += embs[0][i] * embs[1][i]; n2 += embs[2][i] * embs[2][i]; }
        CHECK(std::fabs(n0 - 1.0) < 1e-5 && n2 == 0.0 && dot03 > 0.9 && dot01 < dot03);
        CHECK(sb.getEmbeddingSpaceIdentity() == "simeon-v1-384");
        // the profile of an unconfigured YAMS (Configurable): CharAndWord + Fwht, embedding_dim coordinates, its identity string
        B200SimeonBackend yd(B200SimeonBackend::Profile::Configurable, 1024);
        CHECK(yd.getEmbeddingDimension() == 1024);
        CHECK(yd.getEmbeddingSpaceIdentity() == "simeon-config-v1:char_and_word:3-5:sketch=4096:output=1024:projection=fwht:l2=1");
        auto e2 = yd.generateEmbeddings(texts);
        double m0 = 0, d03 = 0, d01 = 0;
        for (int i = 0; i < 1024; ++i) { m0 += e2[0][i] * e2[0][i]; d03 += e2[0][i] * e2[3][i]; d01 += e2[0][i] * e2[1][i]; }
        CHECK(std::fabs(m0 - 1.0) < 1e-5 && d03 > 0.9 && d01 < d03);
        B200SimeonBackend fixed(B200SimeonBackend::Profile::FixedHash384, 1024);
        CHECK(fixed.getEmbeddingDimension() == 384 && fixed.generateEmbedding(texts[0]) == embs[0]);
    }
    std::puts("ALL OK");
    return 0;
}
