// manifest.cu -- the step that consumes the chunk table `yams add` just produced: the manifest of one file.
//
// Reference (paths under /root/reference):
//   ChunkRef{hash (64 hex chars), offset, u32 size, flags}     include/yams/manifest/manifest_manager.h:48-61
//   ManifestManager::createManifest                             src/manifest/manifest_manager.cpp:411-436
//   ManifestManager::calculateChecksum                          src/manifest/manifest_manager.cpp:705-730
//   ManifestManager::validateManifest (offset / size rules)     src/manifest/manifest_manager.cpp:438-470
//
// calculateChecksum is a bit-serial reflected CRC-32 (polynomial 0xEDB88320, register 0xFFFFFFFF, final complement) over
//     fileHash || to_string(fileSize) || for each chunk: hash || to_string(offset) || to_string(size)
// i.e. ~80 bytes of text per chunk -- 216 MB for the 2.7 M chunks of a 64 GiB stream, one dependent chain on the CPU.
// A CRC is linear over GF(2): for a register s and a message M, reg(s, M) = s * x^(8|M|) + reg(0, M) (mod P), so
// (|M|, reg(0, M)) pairs combine associatively: (la, ra) o (lb, rb) = (la + lb, ra * x^(8 lb) + rb).  The device therefore
//   1. renders every record's text and takes its CRC from a zero register (one thread per chunk; byte table in shared memory),
//      writing the ChunkRef table (hex digests) on the way,
//   2. folds the pairs in order: 64 records per thread, then a shared-memory tree, then one short chain,
//   3. applies the initial register and the final complement.
// The result is bit-identical to the reference's loop (tests/test_gpu_manifest.py pins it against the reference's own
// ManifestManager compiled in place).
#include <algorithm>
#include <mutex>

#include "common.cuh"

namespace yb {

constexpr uint32_t kCrcPoly = 0xEDB88320u;

// polynomial arithmetic in the reflected representation (bit 31 = x^0), as in zlib's crc32_combine
__host__ __device__ __forceinline__ uint32_t gf_mulmod(uint32_t a, uint32_t b) {
    uint32_t m = 1u << 31, p = 0;
    for (;;) {
        if (a & m) {
            p ^= b;
            if ((a & (m - 1)) == 0) break;
        }
        m >>= 1;
        b = (b & 1u) ? (b >> 1) ^ kCrcPoly : b >> 1;
    }
    return p;
}

struct CrcTables {
    uint32_t byte_table[256];   // register after one byte from a zero register
    uint32_t x2n[64];           // x^(2^k) mod P
    uint32_t xpow8[128];        // x^(8 n) mod P, n < 128: the shift over one short record
};

static void build_tables(CrcTables& t) {
    for (uint32_t b = 0; b < 256; ++b) {
        uint32_t c = b;
        for (int i = 0; i < 8; ++i) c = (c >> 1) ^ (kCrcPoly * (c & 1u));
        t.byte_table[b] = c;
    }
    uint32_t p = 1u << 30;   // x^1
    t.x2n[0] = p;
    for (int k = 1; k < 64; ++k) t.x2n[k] = p = gf_mulmod(p, p);
    t.xpow8[0] = 1u << 31;   // x^0
    for (int n = 1; n < 128; ++n) t.xpow8[n] = gf_mulmod(t.xpow8[n - 1], t.x2n[3]);   // * x^8
}

// x^(8 * nbytes) mod P
__device__ __forceinline__ uint32_t x_pow_bytes(const CrcTables* __restrict__ t, uint64_t nbytes) {
    if (nbytes < 128) return t->xpow8[nbytes];
    uint32_t p = 1u << 31;
    uint64_t n = nbytes;
    for (int k = 3; n; n >>= 1, ++k)
        if (n & 1) p = gf_mulmod(t->x2n[k & 63], p);
    return p;
}

struct CrcPart {
    uint64_t len;
    uint32_t reg;
};
__device__ __forceinline__ CrcPart crc_join(const CrcTables* __restrict__ t, CrcPart a, CrcPart b) {
    CrcPart r;
    r.len = a.len + b.len;
    r.reg = (a.reg ? gf_mulmod(x_pow_bytes(t, b.len), a.reg) : 0u) ^ b.reg;
    return r;
}

__device__ __forceinline__ uint32_t crc_byte(const uint32_t* __restrict__ tab, uint32_t reg, uint32_t byte) {
    return (reg >> 8) ^ tab[(reg ^ byte) & 0xFFu];
}
// decimal text of v, most significant digit first, without a digit buffer
__device__ __forceinline__ uint32_t crc_decimal(const uint32_t* __restrict__ tab, uint32_t reg, uint64_t v, uint32_t* len) {
    uint64_t p10 = 1;
    uint32_t digits = 1;
    while (digits < 20 && v / p10 >= 10) {
        p10 *= 10;
        ++digits;
    }
    *len += digits;
    for (uint32_t i = 0; i < digits; ++i) {
        const uint32_t dgt = (uint32_t)(v / p10);
        reg = crc_byte(tab, reg, (uint32_t)'0' + dgt);
        v -= (uint64_t)dgt * p10;
        p10 /= 10;
    }
    return reg;
}
__device__ __forceinline__ uint32_t hex_char(uint32_t nib) { return nib < 10 ? '0' + nib : 'a' + (nib - 10); }

// record 0 = file header (fileHash || to_string(fileSize)); record i + 1 = chunk i
__global__ void manifest_records_kernel(const yams_chunk_desc* __restrict__ chunks, uint32_t n, const uint8_t* __restrict__ file_digest,
                                        uint64_t file_size, const CrcTables* __restrict__ tabs, yams_chunk_ref* __restrict__ refs,
                                        CrcPart* __restrict__ parts, unsigned long long* __restrict__ checks /* [0] bad offsets, [1] bad sizes */) {
    __shared__ uint32_t tab[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) tab[i] = tabs->byte_table[i];
    __syncthreads();
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r > n) return;
    const uint8_t* dg = r == 0 ? file_digest : chunks[r - 1].digest;
    uint32_t reg = 0, len = 64;
    uint32_t* ref_words = (refs && r > 0) ? reinterpret_cast<uint32_t*>(refs[r - 1].hash) : nullptr;   // 80-byte records: 4-byte aligned
    for (int i = 0; i < 32; i += 2) {   // two digest bytes -> four hex characters = one 32-bit word of ChunkRef::hash
        const uint32_t b0 = dg[i], b1 = dg[i + 1];
        const uint32_t c0 = hex_char(b0 >> 4), c1 = hex_char(b0 & 15), c2 = hex_char(b1 >> 4), c3 = hex_char(b1 & 15);
        reg = crc_byte(tab, reg, c0);
        reg = crc_byte(tab, reg, c1);
        reg = crc_byte(tab, reg, c2);
        reg = crc_byte(tab, reg, c3);
        if (ref_words) ref_words[i >> 1] = c0 | (c1 << 8) | (c2 << 16) | (c3 << 24);
    }
    if (r == 0) {
        reg = crc_decimal(tab, reg, file_size, &len);
    } else {
        const uint64_t c_offset = chunks[r - 1].offset, c_size = chunks[r - 1].size;
        // ChunkRef::size is 32 bits (manifest_manager.h:51): createManifest's static_cast<uint32_t>(chunk.size)
        const uint32_t sz32 = (uint32_t)c_size;
        reg = crc_decimal(tab, reg, c_offset, &len);
        reg = crc_decimal(tab, reg, (uint64_t)sz32, &len);
        if (refs) {
            refs[r - 1].offset = c_offset;
            refs[r - 1].size = sz32;
            refs[r - 1].flags = 0;
        }
        // validateManifest (:452-461): offsets are the running sum of the sizes; ChunkRef::isValid: size > 0
        const uint64_t expect = r == 1 ? 0 : chunks[r - 2].offset + (uint64_t)(uint32_t)chunks[r - 2].size;
        if (c_offset != expect) atomicAdd(&checks[0], 1ull);
        if (sz32 == 0 || c_size != (uint64_t)sz32) atomicAdd(&checks[1], 1ull);
    }
    CrcPart p;
    p.len = len;
    p.reg = reg;
    parts[r] = p;
}

// in-order fold of `group` consecutive parts per thread
__global__ void manifest_fold_kernel(const CrcPart* __restrict__ in, uint64_t n, uint32_t group, const CrcTables* __restrict__ tabs,
                                     CrcPart* __restrict__ out) {
    const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t lo = g * group;
    if (lo >= n) return;
    const uint64_t hi = min(n, lo + group);
    CrcPart acc = in[lo];
    for (uint64_t i = lo + 1; i < hi; ++i) acc = crc_join(tabs, acc, in[i]);
    out[g] = acc;
}

// final: one CTA folds what is left (<= 1024 * 64 parts), applies the 0xFFFFFFFF register and the complement
__global__ void manifest_final_kernel(const CrcPart* __restrict__ in, uint64_t n, const CrcTables* __restrict__ tabs,
                                      uint32_t* __restrict__ out_crc, uint64_t* __restrict__ out_len) {
    __shared__ CrcPart sh[1024];
    const uint32_t t = threadIdx.x;
    const uint64_t per = (n + 1023) / 1024;
    CrcPart acc;
    acc.len = 0;
    acc.reg = 0;
    const uint64_t lo = (uint64_t)t * per, hi = min(n, lo + per);
    for (uint64_t i = lo; i < hi; ++i) acc = crc_join(tabs, acc, in[i]);
    sh[t] = acc;
    __syncthreads();
    for (uint32_t stride = 1; stride < 1024; stride <<= 1) {
        CrcPart v = sh[t];
        if ((t % (2 * stride)) == 0 && t + stride < 1024) v = crc_join(tabs, sh[t], sh[t + stride]);
        __syncthreads();
        sh[t] = v;
        __syncthreads();
    }
    if (t == 0) {
        const CrcPart total = sh[0];
        const uint32_t reg = gf_mulmod(x_pow_bytes(tabs, total.len), 0xFFFFFFFFu) ^ total.reg;
        *out_crc = ~reg;
        *out_len = total.len;
    }
}

static CrcTables* g_tables_dev = nullptr;
static std::mutex g_tables_mu;

static yams_status_t tables_on_device(const CrcTables** out) {
    std::lock_guard<std::mutex> lk(g_tables_mu);
    if (!g_tables_dev) {
        CrcTables h;
        build_tables(h);
        CrcTables* d = nullptr;
        YB_CUDA(cudaMalloc(&d, sizeof(CrcTables)));
        cudaError_t e = cudaMemcpy(d, &h, sizeof(CrcTables), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) {
            cudaFree(d);
            set_last_error("crc table upload failed: %s", cudaGetErrorString(e));
            return YAMS_ERR_INTERNAL;
        }
        g_tables_dev = d;
    }
    *out = g_tables_dev;
    return YAMS_OK;
}

}  // namespace yb

using namespace yb;

extern "C" yams_status_t yams_b200_manifest_build(void* self, const yams_chunk_desc* chunks, size_t n, const uint8_t file_digest[32],
                                                  uint64_t file_size, yams_chunk_ref* out_refs, yams_manifest_summary* out) {
    YB_TRY
    (void)self;
    YB_ARG(out && file_digest, "null argument");
    memset(out, 0, sizeof(*out));
    YB_ARG(chunks || n == 0, "chunks is null");
    YB_ARG(n < (1ull << 31), "too many chunks");
    OpWsLease lease;
    OpWs* w = lease.w;
    if (!w) return YAMS_ERR_INTERNAL;
    const CrcTables* tabs = nullptr;
    yams_status_t rc = tables_on_device(&tabs);
    if (rc != YAMS_OK) return rc;
    cudaStream_t st = w->st;
    const uint64_t nrec = (uint64_t)n + 1;
    const uint32_t group = 64;
    const uint64_t n1 = (nrec + group - 1) / group;
    if ((rc = w->d[0].reserve(std::max<size_t>(n, 1) * sizeof(yams_chunk_desc))) != YAMS_OK) return rc;
    if ((rc = w->d[1].reserve(std::max<size_t>(n, 1) * sizeof(yams_chunk_ref))) != YAMS_OK) return rc;
    if ((rc = w->d[2].reserve(nrec * sizeof(CrcPart))) != YAMS_OK) return rc;
    if ((rc = w->d[3].reserve(n1 * sizeof(CrcPart))) != YAMS_OK) return rc;
    if ((rc = w->d[4].reserve(128)) != YAMS_OK) return rc;
    if ((rc = w->h.reserve(128)) != YAMS_OK) return rc;
    uint8_t* d_small = w->d[4].as<uint8_t>();   // [0,32) file digest, [32,48) checks, [48,52) crc, [56,64) len
    YB_CUDA(cudaMemsetAsync(d_small, 0, 128, st));
    memcpy(w->h.p, file_digest, 32);
    YB_CUDA(cudaMemcpyAsync(d_small, w->h.p, 32, cudaMemcpyHostToDevice, st));
    if (n) YB_CUDA(cudaMemcpyAsync(w->d[0].p, chunks, n * sizeof(yams_chunk_desc), cudaMemcpyHostToDevice, st));
    manifest_records_kernel<<<(unsigned)((nrec + 127) / 128), 128, 0, st>>>(w->d[0].as<yams_chunk_desc>(), (uint32_t)n, d_small, file_size, tabs,
                                                                          out_refs ? w->d[1].as<yams_chunk_ref>() : nullptr,
                                                                          w->d[2].as<CrcPart>(), reinterpret_cast<unsigned long long*>(d_small + 32));
    manifest_fold_kernel<<<(unsigned)((n1 + 127) / 128), 128, 0, st>>>(w->d[2].as<CrcPart>(), nrec, group, tabs, w->d[3].as<CrcPart>());
    manifest_final_kernel<<<1, 1024, 0, st>>>(w->d[3].as<CrcPart>(), n1, tabs, reinterpret_cast<uint32_t*>(d_small + 48),
                                              reinterpret_cast<uint64_t*>(d_small + 56));
    YB_CUDA(cudaGetLastError());
    if (out_refs && n) YB_CUDA(cudaMemcpyAsync(out_refs, w->d[1].p, n * sizeof(yams_chunk_ref), cudaMemcpyDeviceToHost, st));
    YB_CUDA(cudaMemcpyAsync(w->h.p, d_small, 128, cudaMemcpyDeviceToHost, st));
    YB_CUDA(cudaStreamSynchronize(st));
    const uint8_t* hs = w->h.as<uint8_t>();
    unsigned long long bad_off = 0, bad_sz = 0;
    memcpy(&bad_off, hs + 32, 8);
    memcpy(&bad_sz, hs + 40, 8);
    memcpy(&out->checksum, hs + 48, 4);
    memcpy(&out->checksum_text_bytes, hs + 56, 8);
    out->chunk_count = n;
    uint64_t total = 0;
    if (n) total = chunks[n - 1].offset + (uint64_t)(uint32_t)chunks[n - 1].size;
    out->total_size = total;
    out->offsets_sequential = bad_off == 0;
    out->sizes_valid = bad_sz == 0;
    // Manifest::isValid (manifest_manager.h:98-104) + validateManifest (:452-468)
    out->valid = n > 0 && file_size > 0 && bad_off == 0 && bad_sz == 0 && total == file_size;
    return YAMS_OK;
    YB_CATCH
}
