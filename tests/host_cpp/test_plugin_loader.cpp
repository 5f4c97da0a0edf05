// Loader-shaped test: what YAMS's AbiPluginLoader does with a plugin (/root/reference/src/daemon/resource/abi_plugin_loader.cpp:
// dlopen(RTLD_LAZY | RTLD_LOCAL) :303-305, dlsym of the envelope symbols :306-341, ABI version check, init once with a JSON config,
// manifest parse :365-392, get_interface by id + version :657-680), followed by a chunk + search THROUGH THE VTABLES -- no symbol of
// the library is linked, everything is reached the way the daemon would reach it.
//   usage: test_plugin_loader /path/to/libyams_b200.so
#include <dlfcn.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/yams_b200.h"

#define CHECK(x)                                                            \
    do {                                                                    \
        if (!(x)) {                                                         \
            std::fprintf(stderr, "CHECK failed %s:%d: %s\n", __FILE__, __LINE__, #x); \
            return 1;                                                       \
        }                                                                   \
    } while (0)

int main(int argc, char** argv) {
    CHECK(argc == 2);
    void* h = dlopen(argv[1], RTLD_LAZY | RTLD_LOCAL);
    if (!h) { std::fprintf(stderr, "dlopen: %s\n", dlerror()); return 1; }
    auto get_abi = reinterpret_cast<int (*)()>(dlsym(h, "yams_plugin_get_abi_version"));
    auto get_name = reinterpret_cast<const char* (*)()>(dlsym(h, "yams_plugin_get_name"));
    auto get_version = reinterpret_cast<const char* (*)()>(dlsym(h, "yams_plugin_get_version"));
    auto get_manifest = reinterpret_cast<const char* (*)()>(dlsym(h, "yams_plugin_get_manifest_json"));
    auto init = reinterpret_cast<int (*)(const char*, const void*)>(dlsym(h, "yams_plugin_init"));
    auto shutdown = reinterpret_cast<void (*)()>(dlsym(h, "yams_plugin_shutdown"));
    auto get_iface = reinterpret_cast<int (*)(const char*, uint32_t, void**)>(dlsym(h, "yams_plugin_get_interface"));
    auto get_health = reinterpret_cast<int (*)(char**)>(dlsym(h, "yams_plugin_get_health_json"));
    CHECK(get_abi && get_name && get_version && get_manifest && init && shutdown && get_iface && get_health);
    CHECK(get_abi() == YAMS_PLUGIN_ABI_VERSION);
    CHECK(std::string(get_name()) == "yams_b200" && std::strlen(get_version()) > 0);
    const std::string manifest = get_manifest();
    CHECK(manifest.find("\"interfaces\"") != std::string::npos && manifest.find("vector_scan_v1") != std::string::npos &&
          manifest.find("content_ingest_v1") != std::string::npos);
    // before init: interfaces are not handed out
    void* none = nullptr;
    CHECK(get_iface(YAMS_IFACE_VECTOR_SCAN_V1, 1, &none) == YAMS_PLUGIN_ERR_INIT_FAILED && none == nullptr);
    CHECK(init("{\"device\": 0}", nullptr) == YAMS_PLUGIN_OK);
    char* health = nullptr;
    CHECK(get_health(&health) == YAMS_PLUGIN_OK && health && std::strstr(health, "\"status\": \"ok\""));
    std::free(health);
    CHECK(get_iface("no_such_iface", 1, &none) == YAMS_PLUGIN_ERR_NOT_FOUND);
    CHECK(get_iface(YAMS_IFACE_VECTOR_SCAN_V1, 99, &none) == YAMS_PLUGIN_ERR_NOT_FOUND);
    void* p = nullptr;
    CHECK(get_iface(YAMS_IFACE_CONTENT_INGEST_V1, YAMS_IFACE_CONTENT_INGEST_V1_VERSION, &p) == YAMS_PLUGIN_OK && p);
    auto* ing = static_cast<yams_content_ingest_v1*>(p);
    CHECK(get_iface(YAMS_IFACE_VECTOR_SCAN_V1, YAMS_IFACE_VECTOR_SCAN_V1_VERSION, &p) == YAMS_PLUGIN_OK && p);
    auto* scan = static_cast<yams_vector_scan_v1*>(p);
    CHECK(ing->abi_version == YAMS_IFACE_CONTENT_INGEST_V1_VERSION && scan->abi_version == YAMS_IFACE_VECTOR_SCAN_V1_VERSION);

    // ---- ingest through the vtable: chunk_and_hash, the SHA-256 KAT of "abc" (tests/unit/crypto/crypto_test.cpp:92-99) ----
    std::vector<uint8_t> data(3u << 20);
    uint64_t sd = 12345;
    for (auto& b : data) { sd = sd * 6364136223846793005ULL + 1442695040888963407ULL; b = (uint8_t)(sd >> 56); }
    yams_cdc_config cfg{};
    cfg.window_size = 48; cfg.min_chunk = 16384; cfg.max_chunk = 1 << 20; cfg.polynomial = 0; cfg.mask = 0x1FFF; cfg.variant = YAMS_CDC_STREAMING;
    yams_chunk_desc* chunks = nullptr;
    size_t n_chunks = 0;
    CHECK(ing->chunk_and_hash(ing->self, data.data(), data.size(), &cfg, &chunks, &n_chunks) == YAMS_OK && n_chunks > 0);
    uint64_t covered = 0;
    for (size_t i = 0; i < n_chunks; ++i) {
        CHECK(chunks[i].offset == covered && chunks[i].size > 0 && chunks[i].size <= cfg.max_chunk);
        covered += chunks[i].size;
    }
    CHECK(covered == data.size());
    // every digest of the table equals sha256_batch over the same spans (second vtable entry, same data)
    std::vector<uint64_t> offs(n_chunks), szs(n_chunks);
    for (size_t i = 0; i < n_chunks; ++i) { offs[i] = chunks[i].offset; szs[i] = chunks[i].size; }
    std::vector<uint8_t> dg(n_chunks * 32);
    CHECK(ing->sha256_batch(ing->self, data.data(), data.size(), offs.data(), szs.data(), n_chunks, dg.data()) == YAMS_OK);
    for (size_t i = 0; i < n_chunks; ++i) CHECK(std::memcmp(dg.data() + 32 * i, chunks[i].digest, 32) == 0);
    ing->free_chunks(ing->self, chunks, n_chunks);
    const uint8_t abc[3] = {'a', 'b', 'c'};
    const uint8_t abc_want[32] = {0xba, 0x78, 0x16, 0xbf, 0x8f, 0x01, 0xcf, 0xea, 0x41, 0x41, 0x40, 0xde, 0x5d, 0xae, 0x22, 0x23,
                                  0xb0, 0x03, 0x61, 0xa3, 0x96, 0x17, 0x7a, 0x9c, 0xb4, 0x10, 0xff, 0x61, 0xf2, 0x00, 0x15, 0xad};
    uint64_t o0 = 0, s3 = 3;
    uint8_t d1[32];
    CHECK(ing->sha256_batch(ing->self, abc, 3, &o0, &s3, 1, d1) == YAMS_OK && std::memcmp(d1, abc_want, 32) == 0);

    // ---- scan through the vtable: the exact-scan contract of tests/unit/vector/vector_smoke_catch2_test.cpp:188-305 ----
    yams_b200_corpus* c = nullptr;
    CHECK(scan->corpus_create(scan->self, 4, YAMS_B200_F32, YAMS_B200_COSINE, 0, &c) == YAMS_OK && c);
    const float rows[6 * 4] = {1, 0, 0, 0, 0.9f, 0.1f, 0, 0, 0, 0, 0, 0, NAN, 1, 0, 0, 1e19f, 0, 0, 0, -1, 0, 0, 0};
    const int64_t rowids[6] = {10, 11, 12, 13, 14, 15};
    CHECK(scan->corpus_append(c, rows, 6, rowids) == YAMS_OK);
    uint64_t sz = 0;
    CHECK(scan->corpus_size(c, &sz) == YAMS_OK && sz == 6);
    const float q[4] = {1, 0, 0, 0};
    int64_t out_r[10];
    float out_s[10];
    uint32_t cnt = 0;
    uint64_t flags = 0;
    CHECK(scan->search(c, q, 1, 10, -1.0f, nullptr, nullptr, out_r, out_s, &cnt, &flags) == YAMS_OK);
    CHECK(cnt == 4 && out_r[0] == 10 && out_r[1] == 14 && out_r[2] == 11 && out_r[3] == 15 && out_s[0] == 1.0f && out_s[3] == -1.0f);
    const float zero[4] = {0, 0, 0, 0};
    CHECK(scan->search(c, zero, 1, 3, -1.0f, nullptr, nullptr, out_r, out_s, &cnt, &flags) == YAMS_ERR_INVALID_ARG);
    uint64_t removed = 0;
    const int64_t gone[1] = {14};
    CHECK(scan->corpus_remove(c, gone, 1, &removed) == YAMS_OK && removed == 1);
    CHECK(scan->search(c, q, 1, 2, -1.0f, nullptr, nullptr, out_r, out_s, &cnt, &flags) == YAMS_OK && cnt == 2 && out_r[0] == 10 && out_r[1] == 11);
    double cs = 0;
    const float a2[2] = {1, 0}, b2[2] = {1, 1};
    CHECK(scan->compute_cosine_similarity(scan->self, a2, 2, b2, 2, &cs) == YAMS_OK && std::fabs(cs - 0.70710678118654757) < 1e-15);
    scan->corpus_destroy(c);
    shutdown();
    CHECK(get_iface(YAMS_IFACE_VECTOR_SCAN_V1, 1, &none) == YAMS_PLUGIN_ERR_INIT_FAILED);
    dlclose(h);
    std::puts("LOADER OK");
    return 0;
}
