/*
 * yams_b200.h -- C ABI of libyams_b200.so: the B200-native drop-in for YAMS's data-parallel hot
 * path (brute-force vector scan behind src/vector + the sqlite-vec-cpp operator surface, and the
 * CDC + SHA-256 ingest path behind src/chunking + src/crypto).
 *
 * Conventions follow the reference's plugin ABI (all path:line under /root/reference):
 *   - envelope: include/yams/plugins/abi.h:17-33 (8 symbols, YAMS_PLUGIN_* return codes);
 *   - vtables:  include/yams/plugins/model_provider_v1.h:44-130 -- struct starts with
 *     {uint32_t abi_version; void* self;}, every fn takes self first and returns yams_status_t,
 *     plugin-allocated buffers are released by a paired free_* fn (host never calls free()),
 *     contiguous row-major [batch, dim] float buffers;
 *   - status codes: model_provider_v1.h:17-25.
 * No exceptions cross this boundary; CUDA failures map to YAMS_ERR_INTERNAL and the text is
 * available through yams_plugin_get_health_json().
 *
 * There is NO CPU fallback: every compute entry point returns YAMS_ERR_INTERNAL (and says why in the
 * health JSON) when no sm_100 device is usable.
 *
 * Threading (model_provider_v1.h:45 "all functions must be thread-safe unless documented"): every entry point may
 * be called from any host thread.  Calls on one corpus or one digest set are serialised by a per-handle mutex; the
 * handle-less ingest entries (chunk_and_hash*, sha256_*) draw a private device workspace from a pool, so concurrent
 * callers run on separate CUDA streams; an ingest SESSION (ingest_open..ingest_close) belongs to one thread at a time.
 * Input pointers are borrowed for the duration of the call only.
 */
#ifndef YAMS_B200_H
#define YAMS_B200_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__) || defined(__clang__)
#define YAMS_B200_API __attribute__((visibility("default")))
#else
#define YAMS_B200_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes (model_provider_v1.h:17-25) --------------------------------------------- */
#ifndef YAMS_PLUGINS_MODEL_PROVIDER_V1_H
enum yams_status_e {
    YAMS_OK = 0,
    YAMS_ERR_INVALID_ARG = 1,
    YAMS_ERR_NOT_FOUND = 2,
    YAMS_ERR_IO = 3,
    YAMS_ERR_INTERNAL = 4,
    YAMS_ERR_UNSUPPORTED = 5
};
typedef int yams_status_t;
#endif

/* ---- plugin envelope (abi.h:17-33) ----------------------------------------------------------
 * Loader: src/daemon/resource/abi_plugin_loader.cpp:303-341 (dlopen RTLD_LAZY|RTLD_LOCAL, dlsym of
 * each symbol), manifest regex-parsed :54-79, interface fetched by id+version :657-680. */
#define YAMS_PLUGIN_ABI_VERSION 1
#define YAMS_PLUGIN_OK 0
#define YAMS_PLUGIN_ERR_INCOMPATIBLE -1
#define YAMS_PLUGIN_ERR_NOT_FOUND -2
#define YAMS_PLUGIN_ERR_INIT_FAILED -3
#define YAMS_PLUGIN_ERR_INVALID -4

YAMS_B200_API int yams_plugin_get_abi_version(void);
YAMS_B200_API const char* yams_plugin_get_name(void);
YAMS_B200_API const char* yams_plugin_get_version(void);
YAMS_B200_API const char* yams_plugin_get_manifest_json(void);
/* config_json keys (all optional): "device": int (default: LOCAL_RANK env or 0) */
YAMS_B200_API int yams_plugin_init(const char* config_json, const void* host_context);
YAMS_B200_API void yams_plugin_shutdown(void);
YAMS_B200_API int yams_plugin_get_interface(const char* iface_id, uint32_t version, void** out_iface);
/* *out_json is malloc'd (strdup), as s3_plugin.cpp does; release with free() */
YAMS_B200_API int yams_plugin_get_health_json(char** out_json);

#define YAMS_IFACE_VECTOR_SCAN_V1 "vector_scan_v1"
#define YAMS_IFACE_VECTOR_SCAN_V1_VERSION 1u
#define YAMS_IFACE_CONTENT_INGEST_V1 "content_ingest_v1"
#define YAMS_IFACE_CONTENT_INGEST_V1_VERSION 1u

/* =============================================================================================
 * content_ingest_v1 -- replaces IChunker (include/yams/chunking/chunker.h:65-92) +
 * IContentHasher (include/yams/crypto/hasher.h:14-46) as injected at
 * src/api/content_store_builder.cpp:152-170.
 * ============================================================================================= */

enum yams_cdc_variant_e {
    YAMS_CDC_STREAMING = 0, /* StreamingChunker (streaming_chunker.h:146-181) -- what `yams add` uses */
    YAMS_CDC_RABIN = 1      /* RabinChunker (rabin_chunker.cpp:63-152) -- bench/test class          */
};

/* ChunkingConfig (chunker.h:44-51) + which reference class to mirror */
typedef struct yams_cdc_config {
    uint64_t window_size; /* 1..48 (the reference ring is a fixed 48-byte array, chunker.h:151) */
    uint64_t min_chunk;
    uint64_t max_chunk;
    uint64_t polynomial;  /* 0 -> default 0x3DA3358B4DC173 (rabin_chunker.cpp:30-36)             */
    uint64_t mask;
    int32_t variant;      /* yams_cdc_variant_e */
    int32_t reserved;
} yams_cdc_config;

/* One chunk: Chunk{hash,offset,size} (chunker.h:18-23) with the raw 32-byte digest; the host
 * adapter hex-encodes (sha256_hasher.cpp:19-30) */
typedef struct yams_chunk_desc {
    uint64_t offset;
    uint64_t size;
    uint8_t digest[32];
} yams_chunk_desc;

typedef struct yams_b200_ingest yams_b200_ingest; /* opaque streaming session */

YAMS_B200_API void yams_b200_cdc_default_config(yams_cdc_config* cfg);

/* IChunker::chunkData / chunkDataLazy (rabin_chunker.cpp:112-152, streaming_chunker.cpp:92-137):
 * boundaries + per-chunk SHA-256 of a HOST buffer. *out is plugin-allocated; free with
 * yams_b200_free_chunks. len == 0 -> *out_n = 0. */
YAMS_B200_API yams_status_t yams_b200_chunk_and_hash(void* self, const uint8_t* data, size_t len,
                                                     const yams_cdc_config* cfg,
                                                     yams_chunk_desc** out, size_t* out_n);
/* Many files per call -- the shape of `yams add -r` (one StreamingChunker run per file; the reference spreads files
 * over ingest workers, src/app/services/indexing_service.cpp, src/api/content_store_impl.cpp:199-220).  Every file is
 * an independent stream (rolling hash and cut positions restart at its first byte), but the whole batch shares ONE
 * candidate scan, ONE cut-selection pass and ONE SHA-256 launch, so small files reach the throughput of one large
 * stream.  files[i] / lens[i]: HOST buffers.  out: plugin-allocated table (free_chunks) with offsets relative to the
 * owning file; chunks of file i are out[out_first[i] .. out_first[i+1]); out_first holds n_files + 1 entries. */
YAMS_B200_API yams_status_t yams_b200_chunk_and_hash_batch(void* self, const uint8_t* const* files,
                                                           const size_t* lens, size_t n_files,
                                                           const yams_cdc_config* cfg, yams_chunk_desc** out,
                                                           size_t* out_n, uint64_t* out_first);

YAMS_B200_API void yams_b200_free_chunks(void* self, yams_chunk_desc* chunks, size_t n);

/* Same with the input already resident in HBM (device pointer). Results come back on the host. */
YAMS_B200_API yams_status_t yams_b200_chunk_and_hash_device(void* self, const uint8_t* d_data,
                                                            size_t len, const yams_cdc_config* cfg,
                                                            yams_chunk_desc** out, size_t* out_n);

/* Streaming form: StreamingChunker::processStream/processBuffer (streaming_chunker.h:78-181).
 * Fragmentation is invisible (tests/unit/chunking/chunking_test.cpp:481-523): feeding any split of
 * a stream yields the same chunks as one call.  Each feed returns the chunks completed so far;
 * finish returns the trailing partial chunk (streaming_chunker.h:115-118). Offsets are stream
 * offsets. */
YAMS_B200_API yams_status_t yams_b200_ingest_open(void* self, const yams_cdc_config* cfg,
                                                  yams_b200_ingest** out);
YAMS_B200_API yams_status_t yams_b200_ingest_feed(yams_b200_ingest* s, const uint8_t* data,
                                                  size_t len, yams_chunk_desc** out, size_t* out_n);
YAMS_B200_API yams_status_t yams_b200_ingest_finish(yams_b200_ingest* s, yams_chunk_desc** out,
                                                    size_t* out_n);
YAMS_B200_API void yams_b200_ingest_close(yams_b200_ingest* s);

/* SHA256Hasher::hash over many spans of one HOST buffer (sha256_hasher.cpp:167-195);
 * digests: n x 32 bytes, caller-owned. */
YAMS_B200_API yams_status_t yams_b200_sha256_batch(void* self, const uint8_t* base, size_t base_len,
                                                   const uint64_t* offsets, const uint64_t* sizes,
                                                   size_t n, uint8_t* digests);
/* base is a device pointer; offsets/sizes/digests are host arrays */
YAMS_B200_API yams_status_t yams_b200_sha256_batch_device(void* self, const uint8_t* d_base,
                                                          size_t base_len, const uint64_t* offsets,
                                                          const uint64_t* sizes, size_t n,
                                                          uint8_t* digests);

/* Only the CDC boundaries (no digests): (offset,size) pairs, digest field zero. */
YAMS_B200_API yams_status_t yams_b200_chunk_boundaries(void* self, const uint8_t* data, size_t len,
                                                       const yams_cdc_config* cfg,
                                                       yams_chunk_desc** out, size_t* out_n);

/* SHA-256 of n SEPARATE host messages in one device pass (IContentHasher::hash per span, hasher.h:14-46;
 * ChunkValidator::validateChunks hashes chunks fetched one by one, src/integrity/chunk_validator.cpp:173-213).
 * digests: n x 32 bytes. */
YAMS_B200_API yams_status_t yams_b200_sha256_many(void* self, const uint8_t* const* msgs, const size_t* lens,
                                                  size_t n, uint8_t* digests);

/* calculateDeduplication (src/chunking/rabin_chunker.cpp:224-239): unique-hash accounting over a chunk table.
 * A chunk is "unique" the first time its digest appears.  chunks: HOST array (as returned by chunk_and_hash). */
typedef struct yams_dedup_stats {
    uint64_t total_size;     /* DeduplicationStats::totalSize   (chunker.h:204-209) */
    uint64_t unique_size;    /* uniqueSize   */
    uint64_t chunk_count;    /* chunkCount   */
    uint64_t unique_chunks;  /* uniqueChunks */
} yams_dedup_stats;
YAMS_B200_API yams_status_t yams_b200_dedup_stats(void* self, const yams_chunk_desc* chunks, size_t n,
                                                  yams_dedup_stats* out);

/* ---- device-resident digest set --------------------------------------------------------
 * Batched form of the per-chunk `storage_->exists(hash)` / `storage_->store(hash, data)` loop that follows
 * chunking (src/api/content_store_impl.cpp:245-288, src/storage/storage_engine.cpp:281-305): one call answers
 * the existence question for every chunk of a file. digests are HOST pointers, `stride` bytes apart (32 for a
 * packed array; sizeof(yams_chunk_desc) with digests = chunks[0].digest for a chunk table). */
typedef struct yams_b200_digest_set yams_b200_digest_set;
YAMS_B200_API yams_status_t yams_b200_digest_set_create(void* self, uint64_t capacity_hint,
                                                        yams_b200_digest_set** out);
/* exists-then-store: out_existed[i] = 1 iff digest i was in the set before the call or equals an EARLIER
 * digest of this batch (exactly what the sequential loop observes); afterwards every digest is in the set.
 * out_existed and out_new (number of digests that were new) are nullable. */
YAMS_B200_API yams_status_t yams_b200_digest_set_insert(yams_b200_digest_set* s, const uint8_t* digests,
                                                        size_t stride, size_t n, uint8_t* out_existed,
                                                        uint64_t* out_new);
/* read-only membership (storage_->exists) */
YAMS_B200_API yams_status_t yams_b200_digest_set_contains(yams_b200_digest_set* s, const uint8_t* digests,
                                                          size_t stride, size_t n, uint8_t* out_exists);
YAMS_B200_API yams_status_t yams_b200_digest_set_size(yams_b200_digest_set* s, uint64_t* out);
/* device time (ms) of the kernels of the last insert/contains call */
YAMS_B200_API yams_status_t yams_b200_digest_set_last_ms(yams_b200_digest_set* s, float* out_ms);
YAMS_B200_API void yams_b200_digest_set_destroy(yams_b200_digest_set* s);

/* ---- manifest of one file ---------------------------------------------------------------
 * The consumer of the chunk table: ManifestManager::createManifest (src/manifest/manifest_manager.cpp:411-436) turns it
 * into ChunkRef{hash (64 lowercase hex chars), offset, u32 size, flags} (include/yams/manifest/manifest_manager.h:48-61)
 * and seals it with calculateChecksum (:705-730): a bit-serial CRC-32 over fileHash || to_string(fileSize) || for each
 * chunk hash || to_string(offset) || to_string(size) -- ~80 text bytes per chunk on one dependent chain.  Here the hex
 * rendering, the per-record CRCs and their GF(2) combination run on the device; the checksum is bit-identical. */
typedef struct yams_chunk_ref {
    char hash[64];      /* lowercase hex, NOT NUL-terminated (HASH_STRING_SIZE, core/types.h:279) */
    uint64_t offset;
    uint32_t size;      /* static_cast<uint32_t>(chunk.size), manifest_manager.h:209 */
    uint32_t flags;
} yams_chunk_ref;
typedef struct yams_manifest_summary {
    uint32_t checksum;             /* Manifest::checksum */
    uint32_t valid;                /* Manifest::isValid (manifest_manager.h:98-104) && the offset / size rules of validateManifest (:452-468) */
    uint32_t offsets_sequential;   /* every chunk.offset == running sum of the sizes */
    uint32_t sizes_valid;          /* every size > 0 and representable in 32 bits */
    uint64_t chunk_count;
    uint64_t total_size;           /* Manifest::calculateTotalSize */
    uint64_t checksum_text_bytes;  /* length of the text the CRC ran over */
} yams_manifest_summary;
/* chunks: HOST table as returned by chunk_and_hash (n may be 0); file_digest: SHA-256 of the whole file (FileInfo::hash,
 * raw 32 bytes); out_refs: nullable HOST array of n entries. */
YAMS_B200_API yams_status_t yams_b200_manifest_build(void* self, const yams_chunk_desc* chunks, size_t n,
                                                     const uint8_t file_digest[32], uint64_t file_size,
                                                     yams_chunk_ref* out_refs, yams_manifest_summary* out);

/* Per-stage device timings (ms) of the last chunk_and_hash* call on this thread's context:
 * [0] candidate scan, [1] cut selection, [2] sha256, [3] total device, [4] h2d (0 for _device) */
YAMS_B200_API yams_status_t yams_b200_ingest_last_timings(void* self, float out_ms[8]);

typedef struct yams_content_ingest_v1 {
    uint32_t abi_version; /* YAMS_IFACE_CONTENT_INGEST_V1_VERSION */
    void* self;
    yams_status_t (*chunk_and_hash)(void* self, const uint8_t* data, size_t len,
                                    const yams_cdc_config* cfg, yams_chunk_desc** out,
                                    size_t* out_n);
    void (*free_chunks)(void* self, yams_chunk_desc* chunks, size_t n);
    yams_status_t (*ingest_open)(void* self, const yams_cdc_config* cfg, yams_b200_ingest** out);
    yams_status_t (*ingest_feed)(yams_b200_ingest* s, const uint8_t* data, size_t len,
                                 yams_chunk_desc** out, size_t* out_n);
    yams_status_t (*ingest_finish)(yams_b200_ingest* s, yams_chunk_desc** out, size_t* out_n);
    void (*ingest_close)(yams_b200_ingest* s);
    yams_status_t (*sha256_batch)(void* self, const uint8_t* base, size_t base_len,
                                  const uint64_t* offsets, const uint64_t* sizes, size_t n,
                                  uint8_t* digests);
    yams_status_t (*dedup_stats)(void* self, const yams_chunk_desc* chunks, size_t n,
                                 yams_dedup_stats* out);
    yams_status_t (*chunk_and_hash_batch)(void* self, const uint8_t* const* files, const size_t* lens,
                                          size_t n_files, const yams_cdc_config* cfg, yams_chunk_desc** out,
                                          size_t* out_n, uint64_t* out_first);
    yams_status_t (*sha256_many)(void* self, const uint8_t* const* msgs, const size_t* lens, size_t n,
                                 uint8_t* digests);
    yams_status_t (*digest_set_create)(void* self, uint64_t capacity_hint, yams_b200_digest_set** out);
    yams_status_t (*digest_set_insert)(yams_b200_digest_set* s, const uint8_t* digests, size_t stride,
                                       size_t n, uint8_t* out_existed, uint64_t* out_new);
    yams_status_t (*digest_set_contains)(yams_b200_digest_set* s, const uint8_t* digests, size_t stride,
                                         size_t n, uint8_t* out_exists);
    yams_status_t (*digest_set_size)(yams_b200_digest_set* s, uint64_t* out);
    void (*digest_set_destroy)(yams_b200_digest_set* s);
    yams_status_t (*manifest_build)(void* self, const yams_chunk_desc* chunks, size_t n, const uint8_t file_digest[32],
                                    uint64_t file_size, yams_chunk_ref* out_refs, yams_manifest_summary* out);
} yams_content_ingest_v1;

/* =============================================================================================
 * vector_scan_v1 -- replaces the exact scan behind IVectorStore::searchSimilar /
 * searchSimilarBatch (include/yams/vector/vector_store.h:44-53) and the exact-candidate seams
 * (:121-138), i.e. SqliteVecBackend::Impl::bruteForceSearchUnlocked
 * (src/vector/sqlite_vec_backend.cpp:4115-4410), and vec0_run_exact_query
 * (third_party/sqlite-vec-cpp/include/sqlite-vec-cpp/sqlite/vec0_module.hpp:376-430).
 * ============================================================================================= */

enum yams_b200_dtype_e { YAMS_B200_F32 = 0, YAMS_B200_F16 = 1 };
enum yams_b200_metric_e { YAMS_B200_COSINE = 0, YAMS_B200_L2 = 1 };

typedef struct yams_b200_corpus yams_b200_corpus; /* opaque device-resident mirror of `vectors` */

/* rows are stored row-major [n, dim] in HBM in `dtype` (fp32 = the reference's BLOB layout,
 * sqlite_vec_backend.cpp:343-363; fp16 = IEEE half bit patterns, utils/float16.hpp). */
YAMS_B200_API yams_status_t yams_b200_corpus_create(void* self, uint32_t dim, int dtype, int metric,
                                                    uint64_t capacity_hint, yams_b200_corpus** out);
/* rows: HOST pointer, n x dim elements of the corpus dtype; rowids nullable (then consecutive,
 * continuing from the current size). Rowids must be non-negative (-1 marks an unused result slot) and appended in
 * strictly ascending order (the reference scans ORDER BY rowid, sqlite_vec_backend.cpp:4175). A rejected batch leaves
 * the corpus unchanged. */
YAMS_B200_API yams_status_t yams_b200_corpus_append(yams_b200_corpus* c, const void* rows, uint64_t n,
                                                    const int64_t* rowids);
/* fp32 HOST rows converted with the reference's TRUNCATING float16_t::from_float
 * (utils/float16.hpp:20-40) into an fp16 corpus */
YAMS_B200_API yams_status_t yams_b200_corpus_append_f32_as_f16(yams_b200_corpus* c, const float* rows,
                                                               uint64_t n, const int64_t* rowids);
/* Synthetic rows generated on the device (SURVEY.md §8d generator; bit-identical to
 * oracle yo_gen_rows_f32 [+ truncating fp16]); rowid = first_row + i */
YAMS_B200_API yams_status_t yams_b200_corpus_append_synthetic(yams_b200_corpus* c, uint64_t seed,
                                                              uint64_t first_row, uint64_t n);
/* Remove rows by rowid (IVectorStore::deleteVector / deleteVectorsByDocument after the SQLite delete,
 * vector_store.h; the scan then no longer sees them, exactly as `SELECT ... FROM vectors` would not).
 * Unknown rowids are ignored; the remaining rows keep their order (ORDER BY rowid). Stable on-device compaction.
 * out_removed (nullable) receives the number of rows actually removed. */
YAMS_B200_API yams_status_t yams_b200_corpus_remove(yams_b200_corpus* c, const int64_t* rowids, uint64_t n,
                                                    uint64_t* out_removed);
YAMS_B200_API yams_status_t yams_b200_corpus_clear(yams_b200_corpus* c);
YAMS_B200_API yams_status_t yams_b200_corpus_size(const yams_b200_corpus* c, uint64_t* out_n);
YAMS_B200_API void yams_b200_corpus_destroy(yams_b200_corpus* c);

#define YAMS_B200_FLAG_TIE_AT_K 1ull       /* equal scores straddle the k boundary: the host   */
                                           /* adapter must re-break by chunk_id (:4218-4223)   */
#define YAMS_B200_FLAG_FALLBACK_PATH 2ull  /* the fast path could not PROVE its answer exact;   */
                                           /* the query was answered by an exhaustive level     */

/* Exact top-k of every query against the corpus (bruteForceSearchUnlocked fast path semantics,
 * sqlite_vec_backend.cpp:4203-4331):
 *   cosine: sim = float(dot / (|row| * |q|)) accumulated in double; rows with a non-finite element
 *   or |row|^2 <= 1e-12 are skipped; sim < threshold dropped; order (sim desc, rowid asc).
 *   l2 (vec0 surface, vec0_module.hpp:376-430): dist = sqrtf(sum (q-r)^2) in float, evaluated as the reference build does
 *   (AVX lane order when dim % 16 == 0, simd/avx.hpp:20-66); order (dist asc, rowid asc); threshold ignored.
 * How "exact" is guaranteed: a tensor-core pass ranks all rows by an approximate score with a known error bound eps, the
 * best K' = k + max(16, k/4) rows are re-scored exactly, and a device-side certificate checks that no other row can
 * reach the k-th exact score (bound + eps < k-th score).  A query that fails the check (near-duplicate rows, an
 * overflowing or short candidate list) is re-run exhaustively inside the same call -- CUDA-core scores of all rows +
 * exact re-scoring of the best 4096, then, if even that cannot be certified, the exact score of every row -- and carries
 * YAMS_B200_FLAG_FALLBACK_PATH.  The returned ids/scores are the reference's in every case.
 * queries: HOST, Q x dim fp32 row-major.  A query that is non-finite or has |q|^2 < 1e-10 makes
 * the whole call return YAMS_ERR_INVALID_ARG (:4127-4130).  k == 0 -> all counts 0 (:4123-4126).
 * allowed_rowids: nullable; when given, query i may only match rowids in
 * allowed_rowids[allowed_offsets[i] .. allowed_offsets[i+1]) (ascending) -- CandidateFilterMode::
 * Exact (src/vector/vector_database.cpp:570-597). allowed_offsets has Q+1 entries.
 * Outputs (caller-owned HOST): out_rowids/out_scores Q x k (unused slots: rowid -1, score 0),
 * out_counts Q, out_flags Q (nullable). * k <= 3072 (cosine) / 4096 (L2). */
YAMS_B200_API yams_status_t yams_b200_search(yams_b200_corpus* c, const float* queries, uint32_t nq,
                                             uint32_t k, float threshold,
                                             const int64_t* allowed_rowids,
                                             const uint64_t* allowed_offsets, int64_t* out_rowids,
                                             float* out_scores, uint32_t* out_counts,
                                             uint64_t* out_flags);

/* ExactRowSelection::AllMatching (sqlite_vec_backend.cpp:4283-4288,4315-4316; IAllExactCandidateVectorStore,
 * vector_store.h:131-138): EVERY row of the candidate set that passes the skip rules and the threshold, ordered
 * (sim desc, rowid asc), scored exactly (double accumulation). One query. out arrays hold n_allowed entries.
 * allowed_rowids == NULL means the whole corpus (out arrays then hold corpus_size entries). */
YAMS_B200_API yams_status_t yams_b200_search_all_matching(yams_b200_corpus* c, const float* query, float threshold,
                                                          const int64_t* allowed_rowids, uint64_t n_allowed,
                                                          int64_t* out_rowids, float* out_scores,
                                                          uint64_t* out_count);

/* Device-pointer form for multi-GPU sharding: queries and outputs are DEVICE pointers, the whole pipeline is enqueued on
 * the corpus stream and the call returns WITHOUT synchronising (query validity, list overflow and the exactness
 * certificate are evaluated on the device and parked in a status block).  Outputs are the rank-local partial top-k in
 * the padded Q x k layout (unused: score -inf / +inf(l2), rowid -1).
 * yams_b200_search_device_finish waits for the stream, returns YAMS_ERR_INVALID_ARG if a query was non-finite / zero,
 * and re-runs the queries the certificate rejected through the exhaustive levels, patching the device outputs in place
 * (*out_resolved = how many; 0 for ordinary data).  A caller that pipelines batches calls it before it consumes the
 * outputs of that batch; only the most recent search_device of a corpus can be finished. */
YAMS_B200_API yams_status_t yams_b200_search_device(yams_b200_corpus* c, const float* d_queries,
                                                    uint32_t nq, uint32_t k, float threshold,
                                                    int64_t* d_out_rowids, float* d_out_scores);
YAMS_B200_API yams_status_t yams_b200_search_device_finish(yams_b200_corpus* c, uint32_t* out_resolved);
/* The exhaustive reference inside the library: exact score of EVERY row (the reference's loop, row-parallel) + a global
 * sort, no tensor-core stage, no thresholds.  Same outputs as yams_b200_search.  This is what the fast path is checked
 * against at corpus sizes a CPU oracle cannot reach (bench.py parity_full, tests); ~3 ms per query per 10 M rows. */
YAMS_B200_API yams_status_t yams_b200_search_exhaustive(yams_b200_corpus* c, const float* queries, uint32_t nq,
                                                        uint32_t k, float threshold, int64_t* out_rowids,
                                                        float* out_scores, uint32_t* out_counts,
                                                        uint64_t* out_flags);
/* Merge R partial results laid out [R][Q][k] (as produced by an all-gather of search_device
 * outputs) into the global top-k, same order. All DEVICE pointers. */
YAMS_B200_API yams_status_t yams_b200_merge_partials_device(yams_b200_corpus* c,
                                                            const int64_t* d_rowids,
                                                            const float* d_scores, uint32_t nranks,
                                                            uint32_t nq, uint32_t k,
                                                            int64_t* d_out_rowids,
                                                            float* d_out_scores,
                                                            uint32_t* d_out_counts);
/* Same merge over the PACKED record an all-gather of one buffer produces: rank r's record is
 * [nq*k int64 rowids][nq*k float scores] (12*nq*k bytes), records back to back.  search_device writes
 * such a record when d_out_scores = (float*)(d_out_rowids + nq*k).  stream: cudaStream_t to launch on (NULL = the
 * corpus stream) so that gather + merge of batch i can overlap the scan of batch i+1. */
YAMS_B200_API yams_status_t yams_b200_merge_packed_device(yams_b200_corpus* c, const void* d_packed,
                                                          uint32_t nranks, uint32_t nq, uint32_t k,
                                                          int64_t* d_out_rowids, float* d_out_scores,
                                                          uint32_t* d_out_counts, void* stream);
YAMS_B200_API yams_status_t yams_b200_corpus_sync(yams_b200_corpus* c);
/* raw cudaStream_t of the corpus (for event timing by the bench) */
YAMS_B200_API void* yams_b200_corpus_stream(yams_b200_corpus* c);

/* vec0_run_exact_query (vec0_module.hpp:376-430): L2 + sqrt in float over fp32 HOST rows, ascending
 * by distance (ties rowid asc), truncated to k when k > 0 (k == 0: all rows); optional inclusive
 * rowid range. out arrays hold min(n, k?k:n) entries. */
YAMS_B200_API yams_status_t yams_b200_vec0_exact(void* self, const float* query, uint32_t dim,
                                                 const float* rows, const int64_t* rowids,
                                                 uint64_t n, uint64_t k, int use_range,
                                                 int64_t rowid_lo, int64_t rowid_hi,
                                                 int64_t* out_rowids, float* out_dist,
                                                 uint64_t* out_count);

/* sqlite-vec-cpp batch surface (distances/batch.hpp:24-146): ONE query against n contiguous fp32 rows, float accumulation
 * in the operation ORDER of the reference build (AVX lane order, simd/avx.hpp; see csrc/ref_order.cuh), so distances are
 * bit-identical to distances::{cosine,l2}_distance<float> (cosine DISTANCE = 1 - cos, 1.0 when the norm product < 1e-8).
 *   mode ALL      batch_distance / batch_distance_contiguous / batch_distance_parallel: out_dist[n] in row order
 *   mode TOP_K    batch_top_k: indices of the k smallest distances, ascending (ties by index -- a legal refinement
 *                 of the reference's unstable partial_sort); out_idx[min(k,n)], out_dist optional
 *   mode FILTERED batch_distance_filtered: every (index, dist) with dist < threshold, ascending; out arrays hold n
 * database is a HOST pointer (n x dim, row-major). */
#define YAMS_B200_BATCH_ALL 0
#define YAMS_B200_BATCH_TOP_K 1
#define YAMS_B200_BATCH_FILTERED 2
YAMS_B200_API yams_status_t yams_b200_batch_distance(void* self, int metric, const float* query, uint32_t dim,
                                                     const float* database, uint64_t n, int mode, uint64_t k,
                                                     float threshold, uint64_t* out_idx, float* out_dist,
                                                     uint64_t* out_count);

/* VectorDatabase::computeCosineSimilarity (vector_database.cpp:1786-1810): double dot / (sqrt(na) * sqrt(nb)),
 * 0 on a size mismatch, an empty vector or a zero norm. Sequential double accumulation -> bit-identical. */
YAMS_B200_API yams_status_t yams_b200_compute_cosine_similarity(void* self, const float* a, size_t na,
                                                                const float* b, size_t nb, double* out);
/* The rerank loops that call computeCosineSimilarity once per candidate (sqlite_vec_backend.cpp:4025,4374,4507) as ONE
 * device pass: n pairs (a_i, b_i), row-major a[n][dim] and b[n][dim] HOST arrays -> out[n]. Same arithmetic per pair. */
YAMS_B200_API yams_status_t yams_b200_compute_cosine_similarity_many(void* self, const float* a, const float* b,
                                                                     size_t n, size_t dim, double* out);

/* ---- SimeonPqAdc engine -----------------------------------------------------------------
 * The reference's default engine (src/vector/sqlite_vec_backend.cpp:3868-4056): a product-quantised index over the
 * L2-normalised embeddings (third_party/simeon/include/simeon/pq.hpp:20-136: m sub-quantisers x k <= 256 centroids,
 * one byte per subspace), scanned with an asymmetric-distance lookup table per query, the best
 * max(k, k * rerank_factor) rows re-scored exactly.  Here the index lives beside a corpus: codes [rows][m] bytes in HBM
 * (32 B per row at the defaults, 1/48 of the fp16 corpus), the exact rerank reads the corpus rows.
 *   pq_build   rebuildSimeonPqIndex (:3540-3690) minus the training: rows with |row|^2 <= 1e-20 are left out, every other
 *              row is normalised like normalizeEmbeddingInPlace (:213-226) and encoded like ProductQuantizer::encode
 *              (simeon/src/pq.cpp:218-235) with the given codebooks [m][k][dim/m] (ProductQuantizer::import_codebooks; the
 *              k-means training of <= 4096 samples stays where it is).  tie_break_keys: nullable, one per corpus row (the
 *              FNV-1a of chunk_id the reference orders equal approximate scores by); NULL = row order.
 *   pq_search  simeonPqSearchUnlocked: normalised query, lookup table (simd::dot order of the x86 build), ADC scores as
 *              sequential float sums, best approxK by (score desc, tie key asc), exact double cosine of the ORIGINAL query
 *              against the stored row (computeCosineSimilarity), threshold, order (similarity desc, rowid asc; TIE_AT_K
 *              flag for the host's chunk_id re-break), k results.  Approximate scores, survivors and final scores are
 *              bit-identical to the reference engine.  A query that cannot be normalised returns nothing.
 * The index is invalidated by any corpus mutation (the reference marks it dirty): pq_search then returns INVALID_ARG. */
typedef struct yams_b200_pq yams_b200_pq;
YAMS_B200_API yams_status_t yams_b200_pq_build(yams_b200_corpus* c, uint32_t m, uint32_t k, const float* codebooks,
                                               const uint64_t* tie_break_keys, yams_b200_pq** out);
YAMS_B200_API yams_status_t yams_b200_pq_search(yams_b200_pq* pq, const float* queries, uint32_t nq, uint32_t k,
                                                uint32_t rerank_factor, float threshold, int64_t* out_rowids,
                                                float* out_scores, uint32_t* out_counts, uint64_t* out_flags);
/* the index content, for parity checks: out_codes [n_indexed][m], out_rowids [n_indexed] (both nullable), *out_n */
YAMS_B200_API yams_status_t yams_b200_pq_codes(yams_b200_pq* pq, uint8_t* out_codes, int64_t* out_rowids,
                                               uint64_t* out_n);
YAMS_B200_API void yams_b200_pq_destroy(yams_b200_pq* pq);

/* ---- Simeon text encoder, default profile ---------------------------------------------------
 * third_party/simeon Encoder::encode (src/simeon.cpp:190-262) for the profile YAMS runs (`simeon-v1-384`, src/simeon.cpp:75-93;
 * src/embedding_simeon/simeon_embedding_backend.cpp:118-135): byte n-grams of length ngram_min..ngram_max over the raw text,
 * SplitMix64 hashing, integer count sketch (+-2 per gram), Achlioptas sparse projection (int64 sums, one float scale),
 * L2 normalisation in the AVX2 tier's lane order.  Bit-identical to the reference encoder.  flags select the other recipe YAMS
 * runs: word tokens (CharAndWord) and the Fwht projection.  Other encoder modes (sub-word tokens, ASCII lowering, word-bounded
 * n-grams, IDF / PMI / sqrt-TF weighting, Gaussian / sparse-JL projections, matryoshka) are not served here. */
typedef struct yams_simeon_config {
    uint32_t ngram_min, ngram_max;     /* 3, 5 */
    uint32_t sketch_dim, output_dim;   /* 4096, 384 */
    uint64_t hash_seed;                /* 0xA5A5A5A5A5A5A5A5 */
    uint64_t projection_seed;          /* 0xDEADBEEFCAFEBABE */
    int32_t l2_normalize;              /* 1 */
    int32_t flags;                     /* YAMS_SIMEON_* below; 0 = the simeon-v1-384 recipe */
} yams_simeon_config;
/* NGramMode::CharAndWord (third_party/simeon/src/tokenizer.cpp:40-52): besides the byte n-grams, every maximal [A-Za-z0-9_]+
 * run is one feature of weight 0.5 (+-1 in the integer sketch) */
#define YAMS_SIMEON_WORD_TOKENS 1
/* ProjectionMode::Fwht (third_party/simeon/src/projection.cpp:149-186,359-374): sign diagonal, Walsh-Hadamard transform of the
 * zero-padded sketch, output_dim sampled coordinates scaled by 1/sqrt(output_dim); requires output_dim <= next_pow2(sketch_dim).
 * Without this flag the projection is AchlioptasSparse. */
#define YAMS_SIMEON_PROJECTION_FWHT 2
typedef struct yams_b200_encoder yams_b200_encoder;
/* simeon_v1_384_config (third_party/simeon/src/simeon.cpp:75-93): EmbeddingConfig::SimeonEncoderProfile::FixedHash384 */
YAMS_B200_API void yams_b200_simeon_default_config(yams_simeon_config* cfg);
/* the encoder YAMS builds when [embeddings.simeon] is left unconfigured (SimeonEncoderProfile::Configurable, the default;
 * /root/reference/src/embedding_simeon/simeon_embedding_backend.cpp:18-47,118-135): CharAndWord, n-grams 3..5, sketch 4096,
 * Fwht, output_dim = embedding_dim (0 -> 1024, EmbeddingConfig's default), L2-normalised */
YAMS_B200_API void yams_b200_simeon_yams_config(yams_simeon_config* cfg, uint32_t embedding_dim);
YAMS_B200_API yams_status_t yams_b200_simeon_create(void* self, const yams_simeon_config* cfg /* NULL = default */,
                                                    yams_b200_encoder** out);
/* IEmbeddingBackend::generateEmbeddings (simeon_embedding_backend.cpp:194-209): n HOST texts (bytes, not NUL-terminated)
 * -> out[n][output_dim] HOST floats, one device pass */
YAMS_B200_API yams_status_t yams_b200_simeon_encode(yams_b200_encoder* e, const char* const* texts, const size_t* lens,
                                                    size_t n, float* out);
YAMS_B200_API void yams_b200_simeon_destroy(yams_b200_encoder* e);

/* [0] stage-1 scan ms, [1] rescoring+select ms, [2] total device ms, [3] h2d+d2h ms of the last
 * yams_b200_search on this corpus; [4] = which stage-1 kernel ran (0 cuda-core, 1 tcgen05);
 * [5] = duration of the full-corpus filtered scan launch alone (the dominant kernel);
 * [6] = queries of that call the certificate sent to the exhaustive levels */
YAMS_B200_API yams_status_t yams_b200_search_last_timings(yams_b200_corpus* c, float out_ms[8]);

typedef struct yams_vector_scan_v1 {
    uint32_t abi_version; /* YAMS_IFACE_VECTOR_SCAN_V1_VERSION */
    void* self;
    yams_status_t (*corpus_create)(void* self, uint32_t dim, int dtype, int metric,
                                   uint64_t capacity_hint, yams_b200_corpus** out);
    yams_status_t (*corpus_append)(yams_b200_corpus* c, const void* rows, uint64_t n,
                                   const int64_t* rowids);
    yams_status_t (*corpus_clear)(yams_b200_corpus* c);
    yams_status_t (*corpus_size)(const yams_b200_corpus* c, uint64_t* out_n);
    void (*corpus_destroy)(yams_b200_corpus* c);
    yams_status_t (*search)(yams_b200_corpus* c, const float* queries, uint32_t nq, uint32_t k,
                            float threshold, const int64_t* allowed_rowids,
                            const uint64_t* allowed_offsets, int64_t* out_rowids, float* out_scores,
                            uint32_t* out_counts, uint64_t* out_flags);
    yams_status_t (*vec0_exact)(void* self, const float* query, uint32_t dim, const float* rows,
                                const int64_t* rowids, uint64_t n, uint64_t k, int use_range,
                                int64_t rowid_lo, int64_t rowid_hi, int64_t* out_rowids,
                                float* out_dist, uint64_t* out_count);
    yams_status_t (*corpus_remove)(yams_b200_corpus* c, const int64_t* rowids, uint64_t n,
                                   uint64_t* out_removed);
    yams_status_t (*search_all_matching)(yams_b200_corpus* c, const float* query, float threshold,
                                         const int64_t* allowed_rowids, uint64_t n_allowed,
                                         int64_t* out_rowids, float* out_scores, uint64_t* out_count);
    yams_status_t (*batch_distance)(void* self, int metric, const float* query, uint32_t dim,
                                    const float* database, uint64_t n, int mode, uint64_t k,
                                    float threshold, uint64_t* out_idx, float* out_dist,
                                    uint64_t* out_count);
    yams_status_t (*compute_cosine_similarity)(void* self, const float* a, size_t na, const float* b,
                                               size_t nb, double* out);
    yams_status_t (*compute_cosine_similarity_many)(void* self, const float* a, const float* b, size_t n,
                                                    size_t dim, double* out);
    yams_status_t (*search_exhaustive)(yams_b200_corpus* c, const float* queries, uint32_t nq, uint32_t k,
                                       float threshold, int64_t* out_rowids, float* out_scores,
                                       uint32_t* out_counts, uint64_t* out_flags);
    yams_status_t (*pq_build)(yams_b200_corpus* c, uint32_t m, uint32_t k, const float* codebooks,
                              const uint64_t* tie_break_keys, yams_b200_pq** out);
    yams_status_t (*pq_search)(yams_b200_pq* pq, const float* queries, uint32_t nq, uint32_t k,
                               uint32_t rerank_factor, float threshold, int64_t* out_rowids, float* out_scores,
                               uint32_t* out_counts, uint64_t* out_flags);
    void (*pq_destroy)(yams_b200_pq* pq);
} yams_vector_scan_v1;

/* ---- sqlite-vec-cpp C API kept bit-for-bit in signature and error behaviour ------------------
 * third_party/sqlite-vec-cpp/src/sqlite_vec_c_api.cpp:57-105 (declared sqlite_vec.hpp:34-50):
 * sizes in BYTES, returns SQLITE_OK(0) / SQLITE_ERROR(1) on null pointer or dimension mismatch.
 * The library has no CPU arithmetic at all: a pair is evaluated by one device thread in the operation order of the
 * reference build (csrc/ref_order.cuh), through a pooled workspace (pinned staging + stream; no allocation per call).
 * One pair costs a launch round trip (~20 us) -- callers with many pairs should use yams_b200_batch_distance /
 * yams_b200_compute_cosine_similarity_many.  The SQL registration (sqlite3_vec_init, vec_distance_l2/cosine/l1 scalar
 * functions) is csrc/sqlite_glue.cpp, compiled only where sqlite3.h exists. */
YAMS_B200_API int sqlite3_vec_distance_l2(const void* vec1, size_t size1, const void* vec2,
                                          size_t size2, float* result);
YAMS_B200_API int sqlite3_vec_distance_cosine(const void* vec1, size_t size1, const void* vec2,
                                              size_t size2, float* result);
/* same contract for the L1 metric of the SQL scalar vec_distance_l1 (sqlite/functions.hpp:144-208, distances/l1.hpp:62-92);
 * the reference C API has no such symbol -- the SQL glue (csrc/sqlite_glue.cpp) needs it */
YAMS_B200_API int yams_b200_vec_distance_l1(const void* vec1, size_t size1, const void* vec2, size_t size2,
                                            float* result);

/* ---- bench / test utilities (SURVEY.md §8d synthetic inputs, generated straight into HBM) ------ */
/* byte[i] = (splitmix64(seed ^ (i>>3)) >> (8*(i&7))) & 0xFF for stream positions [start, start+n) */
YAMS_B200_API yams_status_t yams_b200_synth_bytes_device(uint64_t seed, uint64_t start, uint64_t n,
                                                         uint8_t* d_out);

/* rows [first_row, first_row + n) of the synthetic vector generator (x = (splitmix64(seed ^ (row*dim + c)) >> 40) * 2^-23 - 1,
 * L2-normalised in fp32; bit-identical to oracle yo_gen_rows_f32) as fp32 into a DEVICE buffer of n x dim floats */
YAMS_B200_API yams_status_t yams_b200_synth_rows_device(uint64_t seed, uint64_t first_row, uint64_t n, uint32_t dim,
                                                        float* d_out);

/* diagnostics: dense stage-1 (approximate) scores of rows row_start + i*row_stride, i < nrows, with the
 * chosen engine (0 = CUDA-core, 1 = tcgen05); out[q * nrows + i] on the HOST. YAMS_ERR_UNSUPPORTED when the
 * engine cannot take the shape. */
YAMS_B200_API yams_status_t yams_b200_debug_stage1_scores(yams_b200_corpus* c, const float* queries, uint32_t nq,
                                                          int engine, uint64_t row_start, uint64_t row_stride,
                                                          uint64_t nrows, float* out);

/* diagnostics: the per-query stage-1 error bound eps[q] of the last search / debug_stage1_scores call (HOST out) */
YAMS_B200_API yams_status_t yams_b200_debug_last_eps(yams_b200_corpus* c, uint32_t nq, float* out);

/* ---- misc ------------------------------------------------------------------------------------ */
YAMS_B200_API int yams_b200_device_count(void);
/* last error text of the calling thread (static storage) */
YAMS_B200_API const char* yams_b200_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* YAMS_B200_H */
