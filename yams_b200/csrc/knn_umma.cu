// knn_umma.cu -- stage-1 scan engine on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), sm_100a.
//
// Same contract as the CUDA-core engine in knn.cu (Stage1Args): approximate cosine scores of corpus rows
// against a query batch with a fused epilogue (store, or per-query threshold filter -> candidate lists).
// The scan of `rows x dim` fp16 against `nq x dim` queries IS a dense GEMM (tensor-bound above
// ~Q=281, HBM-bound below), so it runs as:
//
//   warp 0      TMA producer : cp.async.bulk.tensor 2D loads of a 128-row corpus tile (A, K-major, 128B swizzle)
//                               and an N-query tile (B) per 64-element K block into a multi-stage smem ring
//   warp 1      MMA issuer   : one elected thread issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N<=256,
//                               K=16 per instruction) accumulating fp32 in TMEM; tcgen05.commit releases smem
//                               stages and publishes finished accumulators
//   warps 2..5  epilogue     : tcgen05.ld the accumulator (lane = corpus row, column = query), scale by
//                               1/|row|, compare with the per-query threshold, append survivors
//   two TMEM accumulator buffers (2 x 256 columns) let the epilogue of tile i overlap the MMAs of tile i+1.
//
// Queries are pre-scaled by 1/|q| and rounded to fp16; the resulting ~1e-5 score error only affects which
// rows reach stage 2, where the survivors are re-scored exactly in fp64 (knn.cu).
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "knn.cuh"

namespace yb {

constexpr int UM_BLOCK_M = 128;
constexpr int UM_BLOCK_K_BYTES = 128;    // one swizzle row of K per stage: 64 fp16 or 32 fp32 (tf32) elements
constexpr int UM_MMAS_PER_KBLOCK = 4;    // 4 x 32 bytes of K: 4 x (K=16 fp16) or 4 x (K=8 tf32)
constexpr int UM_MAX_N = 256;
constexpr int UM_THREADS = 192;          // 6 warps
constexpr uint32_t UM_A_STAGE = UM_BLOCK_M * UM_BLOCK_K_BYTES;   // 16 KiB
constexpr uint32_t UM_TMEM_COLS = 512;
constexpr long long UM_WAIT_LIMIT_CYCLES = 4000000000ll;
constexpr uint32_t UM_EPI_CAP = 128;      // survivors staged per epilogue warp before a batched flush

struct EpiEntry {
    uint32_t q;
    uint32_t row;
    float score;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, long long* waited = nullptr) {
    uint32_t addr = smem_u32(bar);
    uint32_t done = 0;
    // try_wait may itself suspend the thread until the phase completes, so time from before the first attempt
    const long long t_begin = clock64();
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        // never hang the GPU: a wait longer than ~2 s of SM clocks aborts the kernel
        if (!done && clock64() - t_begin > UM_WAIT_LIMIT_CYCLES) __trap();
    }
    if (waited) *waited += clock64() - t_begin;
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
// pull a tile into L2 ahead of time (no shared-memory destination, no barrier)
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* map, int c0, int c1) {
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];"
                 ::"l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// fp32 operands read as tf32 (the tensor core uses sign, 8 exponent and the top 10 mantissa bits), K = 8 per instruction
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// 32 lanes x 32 columns of fp32 accumulator -> 32 registers per thread
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_ld_32x32_nowait(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major operand tile, 128-byte swizzle: rows of 128 B, 8-row groups 1024 B apart (SBO); version 1 (sm_100)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFFu);   // start address, 16-byte units
    d |= (uint64_t)1 << 16;                        // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;              // stride byte offset between 8-row core groups
    d |= (uint64_t)1 << 46;                        // descriptor version
    d |= (uint64_t)2 << 61;                        // SWIZZLE_128B
    return d;
}

struct UmmaArgs {
    Stage1Args a;
    uint32_t n_tile;      // queries per tile (multiple of 16, <= 256)
    uint32_t nqt;         // query tiles
    uint32_t nrt;         // corpus row tiles (128 rows x CTAS)
    uint32_t kblocks;     // ceil(dim / block_k)
    uint32_t block_k;     // K elements per stage: 64 (fp16) or 32 (fp32 rows read as tf32)
    uint32_t tf32;        // 1: kind::tf32 (fp32 corpus), 0: kind::f16
    uint32_t stages;
    uint32_t b_stage;     // bytes of one B stage held by ONE CTA = (n_tile / CTAS) * 128
    uint32_t idesc;
    unsigned long long* prof;   // nullable diagnostics: per CTA {producer wait, mma wait full, mma wait tempty, epi wait, epi work, total}
    // experiment knobs (tools/exp_umma.py; all 0 in production)
    uint32_t mode;        // 0 normal; 1 loads only (no MMAs); 2 MMAs only (ring filled once, then no loads)
    uint32_t nopf;        // 1: no L2 prefetch of the next row tile
    uint32_t peer_arrive; // CTAS == 2: the peer CTA arrives on the leader's full barrier (count 2) instead of count 1
    uint32_t epi_relaxed; // CTAS == 2: accumulator hand-back without a cluster-scope release
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t rank) {
    uint32_t raddr;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(smem_u32(bar)), "r"(rank));
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
}
// same without the cluster-scope release (a cluster-scope release drains the SM's outstanding memory traffic and costs
// thousands of cycles next to in-flight TMA loads): enough for hand-backs whose data travelled through tcgen05/TMEM and
// were ordered by tcgen05.fence::before_thread_sync
__device__ __forceinline__ void mbar_arrive_remote_relaxed(uint64_t* bar, uint32_t rank) {
    uint32_t raddr;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(smem_u32(bar)), "r"(rank));
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
}
// 2-SM TMA load: bytes are credited to the barrier of the pair's leader CTA (peer bit cleared)
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {   // arrive on this barrier in BOTH CTAs of the pair
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void umma_f16_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_tf32_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}


// ---- warp-uniform role loops -------------------------------------------------------------------------------------------
// The producer and the MMA issuer run with the WHOLE warp converged (every lane polls the barriers) and only the
// tcgen05 / TMA instructions sit under elect.sync.  A role written as `if (lane == 0) { loop }` compiles into a divergent
// region: every UTCHMMA / UTMALDG gets its own elect-and-retry loop and the barrier waits carry their clock reads, ~165
// issued instructions per 64-deep K block -- the single issuing thread then needs ~900 cycles per block for 512 cycles
// of tensor work (tensor pipe 57 % active in profiles/r2_full_umma.md, and the same time per block at N = 128 as at
// N = 256).  Converged, the loop state lives in uniform registers and a K block costs a fraction of that.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ bool mbar_try(uint32_t addr, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    return done != 0;
}
// fast path: one try_wait, no clock reads; the slow path keeps the "never hang the GPU" limit (checked every 1024 spins)
template <bool PROF>
__device__ __forceinline__ void mbar_wait_u(uint32_t addr, uint32_t parity, long long* waited) {
    if (PROF) {
        const long long t_begin = clock64();
        while (!mbar_try(addr, parity))
            if (clock64() - t_begin > UM_WAIT_LIMIT_CYCLES) __trap();
        *waited += clock64() - t_begin;
    } else {
        if (mbar_try(addr, parity)) return;
        long long t_begin = 0;
        uint32_t spins = 0;
        while (!mbar_try(addr, parity)) {
            if ((++spins & 1023u) == 0) {
                const long long now = clock64();
                if (t_begin == 0) t_begin = now;
                else if (now - t_begin > UM_WAIT_LIMIT_CYCLES) __trap();
            }
        }
    }
}
__device__ __forceinline__ void mbar_expect_tx_u(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_u(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote_u(uint32_t bar, uint32_t rank, bool relaxed) {
    uint32_t raddr;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(bar), "r"(rank));
    if (relaxed) asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
    else asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
}
template <int CTAS>
__device__ __forceinline__ void tma_load_u(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    if (CTAS == 2)
        asm volatile(
            "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
            ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar & 0xFEFFFFFFu), "r"(c0), "r"(c1)
            : "memory");
    else
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
            ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
            : "memory");
}
template <int CTAS>
__device__ __forceinline__ void umma_commit_u(uint32_t bar) {
    if (CTAS == 2)
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                     ::"r"(bar), "h"((uint16_t)3) : "memory");
    else
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
template <int CTAS, bool TF32>
__device__ __forceinline__ void umma_u(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    if (TF32) {
        if (CTAS == 2) umma_tf32_2sm(tmem_d, adesc, bdesc, idesc, accumulate); else umma_tf32(tmem_d, adesc, bdesc, idesc, accumulate);
    } else {
        if (CTAS == 2) umma_f16_2sm(tmem_d, adesc, bdesc, idesc, accumulate); else umma_f16(tmem_d, adesc, bdesc, idesc, accumulate);
    }
}

__device__ __noinline__ void append_candidate(uint32_t* counts, Cand* cands, uint32_t cap, uint32_t q, uint32_t row, float sc) {
    uint32_t pos = atomicAdd(&counts[q], 1u);
    if (pos < cap) {
        Cand cd;
        cd.score = sc;
        cd.row = row;
        cands[(uint64_t)q * cap + pos] = cd;
    }
}

// CTAS == 1: one CTA per 128-row tile (cta_group::1).  CTAS == 2: a CTA pair (cluster of 2) shares a 256-row x
// n_tile unit: each CTA stages its own 128 corpus rows and HALF of the query tile, the leader issues
// tcgen05.mma.cta_group::2 (M = 256), each CTA's TMEM receives the accumulator rows of its own corpus rows.
// Halving the query bytes each SM pulls through L2 is what lifts the L2-bound 1-CTA version.
template <bool FILTER, int CTAS, bool TF32, bool PROF>
__global__ void __launch_bounds__(UM_THREADS, 1)
stage1_umma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, UmmaArgs u) {
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment for the 128B-swizzled tiles (identical offsets in both CTAs of a pair)
    // (pointer arithmetic, not integer casts, so the compiler keeps these in the shared address space: LDS/STS/ATOMS)
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const uint32_t stage_bytes = UM_A_STAGE + u.b_stage;
    uint8_t* tiles = smem;
    uint64_t* bars = reinterpret_cast<uint64_t*>(tiles + (size_t)u.stages * stage_bytes);
    uint64_t* full = bars;                      // [stages]
    uint64_t* empty = bars + u.stages;          // [stages]
    uint64_t* tfull = bars + 2 * u.stages;      // [2]
    uint64_t* tempty = tfull + 2;               // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
    float* tmin_s = reinterpret_cast<float*>(tmem_slot + 4);             // [32] smallest threshold of each query tile
    uint32_t* wcnt_s = reinterpret_cast<uint32_t*>(tmin_s + 32);           // [4] staged survivors per epilogue warp
    float* tau_all = reinterpret_cast<float*>(wcnt_s + 4);                 // [nqt * n_tile] thresholds (+inf past nq)
    // (padded to a whole 32-column TMEM chunk: columns past the tile hold stale accumulators and must never pass)
    EpiEntry* stage_s = reinterpret_cast<EpiEntry*>(tau_all + (((size_t)u.nqt * u.n_tile + 31) & ~(size_t)31));   // [4][UM_EPI_CAP]

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = CTAS == 2 ? cluster_ctarank() : 0u;
    const bool leader = rank == 0;
    const uint32_t group = blockIdx.x / CTAS, ngroups = gridDim.x / CTAS;
    // row tiles group, group+ngroups, ... belong to this CTA (pair); each is run against all nqt query tiles
    const uint64_t my_units = group < u.nrt ? (uint64_t)((u.nrt - group + ngroups - 1) / ngroups) * u.nqt : 0;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
        for (uint32_t s = 0; s < u.stages; ++s) {
            mbar_init(&full[s], (CTAS == 2 && u.peer_arrive) ? 2 : 1);   // leader's expect_tx arrive (+ the peer's remote arrive)
            mbar_init(&empty[s], 1);
        }
        mbar_init(&tfull[0], 1);
        mbar_init(&tfull[1], 1);
        mbar_init(&tempty[0], 4 * CTAS);        // 4 epilogue warps per CTA
        mbar_init(&tempty[1], 4 * CTAS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        if (CTAS == 2) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(UM_TMEM_COLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(UM_TMEM_COLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    if (FILTER) {
        const uint32_t nqpad = (u.nqt * u.n_tile + 31u) & ~31u;
        for (uint32_t i = threadIdx.x; i < nqpad; i += UM_THREADS) tau_all[i] = i < u.a.nq ? __ldg(&u.a.tau[i]) : INFINITY;
        if (threadIdx.x < 4) wcnt_s[threadIdx.x] = 0;
        __syncthreads();
        for (uint32_t t = threadIdx.x; t < u.nqt; t += UM_THREADS) {
            float m = INFINITY;
            for (uint32_t c = 0; c < u.n_tile; ++c) m = fminf(m, tau_all[t * u.n_tile + c]);
            tmin_s[t] = m;
        }
    }
    tcgen05_fence_before();
    if (CTAS == 2) cluster_sync_all(); else __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // experiment modes exist in the diagnostic build only (tools/exp_umma.py sets YAMS_B200_UMMA_PROF=1 with them)
    const uint32_t mode = PROF ? u.mode : 0u;
    const uint32_t tiles_u = smem_u32(tiles), full_u = smem_u32(full), empty_u = smem_u32(empty);
    const uint32_t tfull_u = smem_u32(tfull), tempty_u = smem_u32(tempty);
    if (warp == 0) {
        // ===================== TMA producer (every CTA loads its own rows and its share of the queries) =====
        // A CTA (pair) owns whole row tiles and runs every query tile against each: the corpus tile is fetched
        // from HBM once and re-read from L2 by the same SM; the next row tile is prefetched into L2 meanwhile.
        uint32_t s = 0, ph = 0, iter = 0, sa = tiles_u;
        long long w_prod = 0;
        for (uint32_t rt = group; rt < u.nrt; rt += ngroups) {
            const int a_row = (int)((rt * CTAS + rank) * UM_BLOCK_M);
            const int a_next = (int)(((rt + ngroups) * CTAS + rank) * UM_BLOCK_M);
            const bool have_next = rt + ngroups < u.nrt && !u.nopf;
            for (uint32_t qt = 0; qt < u.nqt; ++qt) {
                const int b_row = (int)(qt * u.n_tile + rank * (u.n_tile / CTAS));
                const bool pf = qt == 0 && have_next;
                int kc = 0;
                for (uint32_t kb = 0; kb < u.kblocks; ++kb, kc += (int)u.block_k) {
                    mbar_wait_u<PROF>(empty_u + 8u * s, ph ^ 1u, &w_prod);
                    if (elect_one()) {
                        const uint32_t fb = full_u + 8u * s;
                        if (mode == 2 && iter >= u.stages) {   // experiment: operands stay in place, barriers only
                            if (leader) mbar_arrive_u(fb);
                            else if (u.peer_arrive) mbar_arrive_remote_u(fb, 0, false);
                        } else {
                            if (CTAS == 2) {
                                if (leader) mbar_expect_tx_u(fb, stage_bytes * 2);
                                else if (u.peer_arrive) mbar_arrive_remote_u(fb, 0, false);
                            } else {
                                mbar_expect_tx_u(fb, stage_bytes);
                            }
                            tma_load_u<CTAS>(sa, &tmA, fb, kc, a_row);
                            tma_load_u<CTAS>(sa + UM_A_STAGE, &tmB, fb, kc, b_row);
                        }
                        if (pf) tma_prefetch_2d(&tmA, kc, a_next);
                    }
                    ++iter;
                    sa += stage_bytes;
                    if (++s == u.stages) { s = 0; ph ^= 1u; sa = tiles_u; }
                }
            }
        }
        if (PROF && lane == 0 && u.prof) u.prof[blockIdx.x * 8 + 0] = (unsigned long long)w_prod;
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA only) =====================
        if (leader) {
            uint32_t s = 0, ph = 0, it = 0, sa = tiles_u;
            long long w_full = 0, w_tempty = 0;
            const long long t_start = PROF ? clock64() : 0;
            // K-major operand tile, 128-byte swizzle (make_smem_desc): everything but the start address is constant
            const uint64_t desc_hi = make_smem_desc(0);
            for (uint64_t unit = 0; unit < my_units; ++unit, ++it) {
                const uint32_t as = it & 1u, aph = (it >> 1) & 1u;
                mbar_wait_u<PROF>(tempty_u + 8u * as, aph ^ 1u, &w_tempty);
                tcgen05_fence_after();
                const uint32_t tmem_d = tmem_base + as * UM_MAX_N;
                for (uint32_t kb = 0; kb < u.kblocks; ++kb) {
                    mbar_wait_u<PROF>(full_u + 8u * s, ph, &w_full);
                    tcgen05_fence_after();
                    if (elect_one()) {
                        if (mode == 1) {   // experiment: loads only
                            mbar_arrive_u(empty_u + 8u * s);
                            if (CTAS == 2) mbar_arrive_remote_u(empty_u + 8u * s, 1, false);
                        } else {
                            const uint64_t adesc = desc_hi | (uint64_t)((sa >> 4) & 0x3FFFu);
                            const uint64_t bdesc = desc_hi | (uint64_t)(((sa + UM_A_STAGE) >> 4) & 0x3FFFu);
                            // advance 32 bytes (16 fp16 / 8 tf32) along K inside the swizzle row: +2 in 16-byte units
                            umma_u<CTAS, TF32>(tmem_d, adesc, bdesc, u.idesc, kb != 0 ? 1u : 0u);
#pragma unroll
                            for (uint32_t k = 1; k < UM_MMAS_PER_KBLOCK; ++k) umma_u<CTAS, TF32>(tmem_d, adesc + 2 * k, bdesc + 2 * k, u.idesc, 1u);
                            // smem stage reusable (in both CTAs) once these MMAs have read it
                            umma_commit_u<CTAS>(empty_u + 8u * s);
                        }
                    }
                    sa += stage_bytes;
                    if (++s == u.stages) { s = 0; ph ^= 1u; sa = tiles_u; }
                }
                if (elect_one()) {
                    if (mode == 1) {
                        mbar_arrive_u(tfull_u + 8u * as);
                        if (CTAS == 2) mbar_arrive_remote_u(tfull_u + 8u * as, 1, false);
                    } else {
                        umma_commit_u<CTAS>(tfull_u + 8u * as);   // accumulator complete
                    }
                }
            }
            if (PROF && lane == 0 && u.prof) {
                u.prof[blockIdx.x * 8 + 1] = (unsigned long long)w_full;
                u.prof[blockIdx.x * 8 + 2] = (unsigned long long)w_tempty;
                u.prof[blockIdx.x * 8 + 5] = (unsigned long long)(clock64() - t_start);
            }
        }
    } else {
        // ===================== epilogue (warps 2..5) =====================
        const uint32_t quad = warp & 3;              // TMEM lane quadrant this warp may access
        EpiEntry* my_stage = stage_s + quad * UM_EPI_CAP;
        uint32_t* my_cnt = wcnt_s + quad;
        const bool direct = u.a.mask != nullptr;     // candidate-set mode: most allowed rows survive, no staging
        uint32_t it = 0;
        long long w_epi = 0, w_ld = 0, t_ldissue = 0, t_proc = 0, t_tail = 0, t_head = 0;
        unsigned long long n_slow = 0;
        const long long e_start = PROF ? clock64() : 0;
        constexpr bool profiling = PROF;

        auto append_global = [&](uint32_t q, uint32_t row, float sc) { append_candidate(u.a.counts, u.a.cands, u.a.cap, q, row, sc); };
        auto flush = [&]() {   // warp-uniform: drain the staged survivors with the atomics in flight together
            __syncwarp();
            uint32_t n = min(*my_cnt, UM_EPI_CAP);
            for (uint32_t i = lane; i < n; i += 32) {
                EpiEntry e = my_stage[i];
                append_global(e.q, e.row, e.score);
            }
            __syncwarp();
            if (lane == 0) *my_cnt = 0;
            __syncwarp();
        };
        // row-level data of the first unit
        auto row_of = [&](uint64_t unit, uint64_t& li_out) {
            const uint32_t rt_ = group + (uint32_t)(unit / u.nqt) * ngroups;
            li_out = ((uint64_t)rt_ * CTAS + rank) * UM_BLOCK_M + quad * 32 + lane;
        };
        uint64_t li = 0;
        float inr = 0.f;
        if (my_units) {
            row_of(0, li);
            inr = li < u.a.nrows ? __ldg(&u.a.inv_norm[u.a.row_start + li * u.a.row_stride]) : 0.f;
        }
        for (uint64_t unit = 0; unit < my_units; ++unit, ++it) {
            const long long t_h0 = profiling ? clock64() : 0;
            const uint32_t qt = (uint32_t)(unit % u.nqt);
            const uint32_t as = it & 1, aph = (it >> 1) & 1;
            const bool rvalid = li < u.a.nrows;
            const uint64_t grow = u.a.row_start + li * u.a.row_stride;
            // prefetch the next unit's row scale while this unit is processed
            uint64_t li_next = li;
            float inr_next = inr;
            if (unit + 1 < my_units && qt + 1 == u.nqt) {
                row_of(unit + 1, li_next);
                inr_next = li_next < u.a.nrows ? __ldg(&u.a.inv_norm[u.a.row_start + li_next * u.a.row_stride]) : 0.f;
            }
            const uint32_t q0 = qt * u.n_tile;
            const uint32_t taddr = tmem_base + ((quad * 32u) << 16) + as * UM_MAX_N;
            if (profiling) t_head += clock64() - t_h0;
            mbar_wait_u<PROF>(tfull_u + 8u * as, aph, &w_epi);
            tcgen05_fence_after();
            // score = acc * alpha + beta.  cosine: alpha = 1/|row| (0 marks a row the reference skips), beta = 0;
            // L2: alpha = 2, beta = -|row|^2 (the row slot holds |row|^2; +inf marks a non-finite row)
            const bool l2 = u.a.metric == YAMS_B200_L2;
            const float alpha = l2 ? 2.f : inr;
            const float beta = l2 ? -inr : 0.f;
            const bool rowok = rvalid && (l2 ? inr < INFINITY : inr > 0.f);
            if (mode == 1) {
                // experiment: no accumulators to read
            } else if (FILTER) {
                const float* tau_t = tau_all + q0;
                // conservative row-level bound: acc * alpha + beta > tau_q  =>  acc > (tmin - beta) / alpha (slightly relaxed)
                float bound = INFINITY;
                if (rowok) {
                    float b0 = (tmin_s[qt] - beta) * (1.0f / alpha);
                    bound = b0 - (fabsf(tmin_s[qt]) + fabsf(beta)) * (2e-6f / alpha) - 1e-30f;
                }
                auto process = [&](const uint32_t (&v)[32], uint32_t c0) {
                    float m = __uint_as_float(v[0]);
#pragma unroll
                    for (int c = 1; c + 1 < 32; c += 2)
                        asm("max.f32 %0, %0, %1, %2;" : "+f"(m) : "f"(__uint_as_float(v[c])), "f"(__uint_as_float(v[c + 1])));
                    m = fmaxf(m, __uint_as_float(v[31]));
                    if (m > bound) {   // rare (~1% of rows per chunk): some column of this row may survive
                        ++n_slow;
                        // exact per-column test, branch-free: thresholds come in as 8 independent LDS.128
                        uint32_t pm = 0;
#pragma unroll
                        for (int c = 0; c < 32; c += 4) {
                            const float4 t4 = *reinterpret_cast<const float4*>(tau_t + c0 + c);
                            pm |= (fmaf(__uint_as_float(v[c + 0]), alpha, beta) > t4.x ? 1u : 0u) << (c + 0);
                            pm |= (fmaf(__uint_as_float(v[c + 1]), alpha, beta) > t4.y ? 1u : 0u) << (c + 1);
                            pm |= (fmaf(__uint_as_float(v[c + 2]), alpha, beta) > t4.z ? 1u : 0u) << (c + 2);
                            pm |= (fmaf(__uint_as_float(v[c + 3]), alpha, beta) > t4.w ? 1u : 0u) << (c + 3);
                        }
                        // survivors (about one per entry): compact loop, the column is extracted with a select tree
                        // so the code stays small (an unrolled per-column body thrashed the instruction cache)
                        while (pm) {
                            const int c = __ffs(pm) - 1;
                            pm &= pm - 1;
                            uint32_t x16[16], x8[8], x4[4], x2[2];
#pragma unroll
                            for (int i = 0; i < 16; ++i) x16[i] = (c & 16) ? v[i + 16] : v[i];
#pragma unroll
                            for (int i = 0; i < 8; ++i) x8[i] = (c & 8) ? x16[i + 8] : x16[i];
#pragma unroll
                            for (int i = 0; i < 4; ++i) x4[i] = (c & 4) ? x8[i + 4] : x8[i];
#pragma unroll
                            for (int i = 0; i < 2; ++i) x2[i] = (c & 2) ? x4[i + 2] : x4[i];
                            const float sc = fmaf(__uint_as_float((c & 1) ? x2[1] : x2[0]), alpha, beta);
                            const uint32_t q = q0 + c0 + (uint32_t)c;
                            if (direct) {
                                uint32_t w = __ldg(&u.a.mask[(uint64_t)q * u.a.mask_ld + (grow >> 5)]);
                                if ((w >> (grow & 31)) & 1u) append_global(q, (uint32_t)grow, sc);
                            } else {
                                uint32_t idx = atomicAdd(my_cnt, 1u);
                                if (idx < UM_EPI_CAP) {
                                    EpiEntry e;
                                    e.q = q; e.row = (uint32_t)grow; e.score = sc;
                                    my_stage[idx] = e;
                                } else {
                                    append_global(q, (uint32_t)grow, sc);
                                }
                            }
                        }
                    }
                };
                // software-pipelined TMEM reads: the load of chunk c+1 is in flight while chunk c is tested
                uint32_t va[32], vb[32];
                auto timed_wait = [&]() {
                    if (profiling) {
                        long long t0 = clock64();
                        tmem_ld_wait();
                        w_ld += clock64() - t0;
                    } else {
                        tmem_ld_wait();
                    }
                };
                long long tq = profiling ? clock64() : 0;
                auto lap = [&](long long& acc) {
                    if (profiling) {
                        long long now = clock64();
                        acc += now - tq;
                        tq = now;
                    }
                };
                tmem_ld_32x32_nowait(taddr, va);
                lap(t_ldissue);
                timed_wait();
                lap(w_ld);
                for (uint32_t c0 = 0; c0 < u.n_tile; c0 += 64) {
                    const bool has_b = c0 + 32 < u.n_tile;
                    if (has_b) tmem_ld_32x32_nowait(taddr + c0 + 32, vb);
                    lap(t_ldissue);
                    process(va, c0);
                    lap(t_proc);
                    tmem_ld_wait();
                    lap(w_ld);
                    if (c0 + 64 < u.n_tile) tmem_ld_32x32_nowait(taddr + c0 + 64, va);
                    lap(t_ldissue);
                    if (has_b) process(vb, c0 + 32);
                    lap(t_proc);
                    tmem_ld_wait();
                    lap(w_ld);
                }
            } else {
                for (uint32_t c0 = 0; c0 < u.n_tile; c0 += 32) {
                    uint32_t v[32];
                    tmem_ld_32x32(taddr + c0, v);
                    if (q0 + c0 >= u.a.nq || !rvalid) continue;
#pragma unroll
                    for (int c = 0; c < 32; ++c) {
                        const uint32_t q = q0 + c0 + c;
                        if (q < u.a.nq)
                            u.a.out_scores[(uint64_t)q * u.a.ld + li] = rowok ? fmaf(__uint_as_float(v[c]), alpha, beta) : -INFINITY;
                    }
                }
            }
            const long long t_t0 = profiling ? clock64() : 0;
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (CTAS == 2) {
                    if (u.epi_relaxed) mbar_arrive_remote_relaxed(&tempty[as], 0); else mbar_arrive_remote(&tempty[as], 0);
                } else mbar_arrive(&tempty[as]);
            }
            if (FILTER && !direct && *my_cnt >= (UM_EPI_CAP * 3) / 4) flush();   // warp-uniform (smem value)
            li = li_next;
            inr = inr_next;
            if (profiling) t_tail += clock64() - t_t0;
        }
        if (FILTER && !direct) flush();
        if (PROF && u.prof && quad == 0 && lane == 0) {
            u.prof[blockIdx.x * 8 + 3] = (unsigned long long)w_epi;
            u.prof[blockIdx.x * 8 + 4] = (unsigned long long)(clock64() - e_start);
            u.prof[blockIdx.x * 8 + 6] = (unsigned long long)w_ld;
            u.prof[blockIdx.x * 8 + 7] = n_slow;
            // detail slots live after the 8 per-CTA slots of all CTAs
            unsigned long long* d = u.prof + (size_t)gridDim.x * 8 + (size_t)blockIdx.x * 4;
            d[0] = (unsigned long long)t_ldissue; d[1] = (unsigned long long)t_proc; d[2] = (unsigned long long)t_tail; d[3] = (unsigned long long)t_head;
        }
    }
    tcgen05_fence_before();
    if (CTAS == 2) cluster_sync_all(); else __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        if (CTAS == 2)
            asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(UM_TMEM_COLS) : "memory");
        else
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(UM_TMEM_COLS) : "memory");
    }
}

// Query preparation, one CTA per query: the operand-typed copy the tensor cores read (fp16, or fp32 read as tf32), pre-scaled
// by 1/|q| for the cosine metric, and the per-query stage-1 error bound eps[q] that the exactness certificate uses.
// With x = the scaled query, y = what the tensor core multiplies (fp16-rounded / tf32-truncated x), dq = |x - y|, and the
// rows r read as r' (exact for fp16 rows; tf32-truncated for fp32 rows, residual norms tracked per corpus):
//   |x.r - y.r'| <= dq |r'| + |x| |r - r'|            (Cauchy-Schwarz with the ACTUAL rounding residuals)
//   cosine (|x| = 1, score = x.r / |r|):  eps = 1.25 (dq + dr_rel_max) + 2^-23 (d + 4)
//   L2 (score = 2 x.r - |r|^2)         :  eps = 2.5 (dq Rmax + |x| dr_abs_max) + 2^-22 (d + 4) |x| Rmax + 2^-22 Rmax^2
// The 1.25 covers the second-order term and the fp32 accumulation inside the tensor core is the 2^-23 (d + 4) term.
template <bool TF32>
__global__ void __launch_bounds__(128) queries_prep_kernel(const float* __restrict__ q32, const float* __restrict__ qinv, uint32_t nq,
                                                           uint32_t d, void* __restrict__ out, int metric, float r_max, float dr_abs_max,
                                                           float dr_rel_max, float* __restrict__ eps) {
    const uint32_t q = blockIdx.x;
    const float scale = qinv[q];   // 1 for L2
    float ss = 0.f, xx = 0.f;
    for (uint32_t c = threadIdx.x; c < d; c += 128) {
        const float x = q32[(uint64_t)q * d + c] * scale;
        float y;
        if (TF32) {
            reinterpret_cast<float*>(out)[(uint64_t)q * d + c] = x;
            y = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
        } else {
            const __half h = __float2half_rn(x);
            reinterpret_cast<__half*>(out)[(uint64_t)q * d + c] = h;
            y = __half2float(h);
        }
        const float r = x - y;
        ss = fmaf(r, r, ss);
        xx = fmaf(x, x, xx);
    }
    __shared__ float red[2][4];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        ss += __shfl_xor_sync(0xffffffffu, ss, o);
        xx += __shfl_xor_sync(0xffffffffu, xx, o);
    }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = ss; red[1][threadIdx.x >> 5] = xx; }
    __syncthreads();
    if (threadIdx.x == 0) {
        const float dq = sqrtf(red[0][0] + red[0][1] + red[0][2] + red[0][3]) * 1.000001f;
        const float xn = sqrtf(red[1][0] + red[1][1] + red[1][2] + red[1][3]) * 1.000001f;
        const float u = 1.1920929e-07f;   // 2^-23
        float e;
        if (metric == YAMS_B200_L2) e = 2.5f * (dq * r_max + xn * dr_abs_max) + 2.f * u * (float)(d + 4) * xn * r_max + 2.f * u * r_max * r_max + 1e-30f;
        else e = 1.25f * (dq + xn * dr_rel_max) + u * (float)(d + 4) + 1e-7f;
        eps[q] = e;
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

static bool make_map_2d(CUtensorMap* m, const void* base, uint64_t inner, uint64_t outer, uint64_t outer_stride_bytes,
                        uint32_t box_inner, uint32_t box_outer, bool f32 = false) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return false;
    cuuint64_t dims[2] = {inner, outer};
    cuuint64_t strides[1] = {outer_stride_bytes};
    cuuint32_t box[2] = {box_inner, box_outer};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(m, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

bool tcgen05_supported(const Corpus* c, uint32_t nq) {
    (void)nq;
    if (getenv("YAMS_B200_DISABLE_TCGEN05")) return false;
    if (get_encode_fn() == nullptr) return false;
    // fp16 rows: kind::f16; fp32 rows (the reference's BLOB layout): kind::tf32.  TMA needs 16-byte row pitches.
    return c->dtype == YAMS_B200_F16 ? (c->dim % 8 == 0) : (c->dim % 4 == 0 && !getenv("YAMS_B200_DISABLE_TF32"));
}

yams_status_t stage1_tcgen05(Corpus* c, const Stage1Args& a, bool filter, cudaStream_t st) {
    const bool tf32 = a.dtype == YAMS_B200_F32;
    const uint32_t esz = tf32 ? 4 : 2;
    if ((a.dim * esz) % 16 != 0 || a.nrows == 0 || a.nq == 0) return YAMS_ERR_UNSUPPORTED;
    if (tf32 && getenv("YAMS_B200_DISABLE_TF32")) return YAMS_ERR_UNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(a.rows) & 15) != 0) return YAMS_ERR_UNSUPPORTED;
    if (a.nrows >= (1ull << 32)) return YAMS_ERR_UNSUPPORTED;
    yams_status_t rc;
    // queries pre-scaled by 1/|q| (so the epilogue only multiplies by 1/|row|), in the operand type of the engine
    size_t qb = (size_t)a.nq * a.dim * esz;
    if ((rc = c->q16.reserve(qb + 256)) != YAMS_OK) return rc;
    void* d_q16 = c->q16.p;
    if (!a.skip_qprep) {
        if (!a.eps) return YAMS_ERR_UNSUPPORTED;
        if (tf32) queries_prep_kernel<true><<<a.nq, 128, 0, st>>>(a.q32, a.qinv, a.nq, a.dim, d_q16, a.metric, a.r_max, a.dr_abs_max,
                                                                   a.dr_rel_max, a.eps);
        else queries_prep_kernel<false><<<a.nq, 128, 0, st>>>(a.q32, a.qinv, a.nq, a.dim, d_q16, a.metric, a.r_max, 0.f, 0.f, a.eps);
    }

    auto env_int = [](const char* name, int dflt) { const char* e = getenv(name); return e ? atoi(e) : dflt; };
    // 2-CTA pairs (cta_group::2, M = 256) for tensor-bound batches: ~3 % faster under the power cap (half the query-operand shared-memory
    // reads).  HBM-bound batches (below ~2 query tiles) stream slightly faster with independent CTAs (measured 2.84 vs 2.89 ms per 10 M rows).
    const int ctas_env = env_int("YAMS_B200_UMMA_CTAS", a.nq >= 512 ? 2 : 1);
    const int ctas = (ctas_env == 1 || c->dev->sm_count < 2) ? 1 : 2;
    UmmaArgs u{};
    u.a = a;
    u.mode = (uint32_t)env_int("YAMS_B200_UMMA_MODE", 0);
    u.nopf = (uint32_t)env_int("YAMS_B200_UMMA_NOPF", 0);
    u.peer_arrive = (uint32_t)env_int("YAMS_B200_UMMA_PEER_ARRIVE", 0);
    u.epi_relaxed = (uint32_t)env_int("YAMS_B200_UMMA_EPI_RELAXED", 1);
    const uint32_t nmult = 16 * ctas;   // each CTA's share of the query tile must be a multiple of 8 rows (16 for M=128)
    const uint32_t max_n = (uint32_t)std::min(UM_MAX_N, std::max(32, env_int("YAMS_B200_UMMA_NTILE", UM_MAX_N)));
    u.n_tile = a.nq >= max_n ? max_n : ((a.nq + nmult - 1) / nmult) * nmult;
    u.nqt = (a.nq + u.n_tile - 1) / u.n_tile;
    const uint32_t rows_per_unit = UM_BLOCK_M * ctas;
    u.nrt = (uint32_t)((a.nrows + rows_per_unit - 1) / rows_per_unit);
    u.tf32 = tf32 ? 1u : 0u;
    u.block_k = UM_BLOCK_K_BYTES / esz;
    u.kblocks = (a.dim + u.block_k - 1) / u.block_k;
    u.b_stage = (u.n_tile / ctas) * 128;
    uint32_t stage_bytes = UM_A_STAGE + u.b_stage;
    if (u.nqt > 32 || (size_t)u.nqt * u.n_tile > 4096) return YAMS_ERR_UNSUPPORTED;   // thresholds live in shared memory
    const uint32_t budget = 224 * 1024 - (uint32_t)(32 * 4 + 16 + (((size_t)u.nqt * u.n_tile + 31) & ~(size_t)31) * 4 + 4 * UM_EPI_CAP * sizeof(EpiEntry)) - 2048;
    u.stages = std::min<uint32_t>(8, budget / stage_bytes);
    if (int st_env = env_int("YAMS_B200_UMMA_STAGES", 0)) u.stages = std::min<uint32_t>(u.stages, (uint32_t)st_env);
    if (u.stages < 2) return YAMS_ERR_UNSUPPORTED;
    // instruction descriptor: D=f32 (bits 4-5), A/B format (bits 7-9 / 10-12: 0 = f16, 2 = tf32), both K-major,
    // N = n_tile (bits 17-22, /8), M = 128 per CTA (bits 24-28, /16)
    u.idesc = (1u << 4) | (tf32 ? ((2u << 7) | (2u << 10)) : 0u) | ((u.n_tile >> 3) << 17) | (((UM_BLOCK_M * ctas) >> 4) << 24);

    CUtensorMap tmA, tmB;
    const uint8_t* a_base = reinterpret_cast<const uint8_t*>(a.rows) + (size_t)a.row_start * a.dim * esz;
    if (!make_map_2d(&tmA, a_base, a.dim, a.nrows, (uint64_t)a.row_stride * a.dim * esz, u.block_k, UM_BLOCK_M, tf32))
        return YAMS_ERR_UNSUPPORTED;
    if (!make_map_2d(&tmB, d_q16, a.dim, a.nq, (uint64_t)a.dim * esz, u.block_k, u.n_tile / ctas, tf32)) return YAMS_ERR_UNSUPPORTED;

    const size_t epi_bytes = 32 * 4 + 16 + (((size_t)u.nqt * u.n_tile + 31) & ~(size_t)31) * 4 + 4 * UM_EPI_CAP * sizeof(EpiEntry);
    size_t smem = (size_t)u.stages * stage_bytes + 1024 /*align slack*/ + (2 * u.stages + 4) * 8 + 16 + epi_bytes;
    unsigned groups = (unsigned)std::min<uint64_t>(u.nrt, (uint64_t)(c->dev->sm_count / ctas));
    if (getenv("YAMS_B200_UMMA_PROF")) {
        YB_CUDA(cudaMalloc(&u.prof, (size_t)groups * ctas * 12 * 8));
        YB_CUDA(cudaMemsetAsync(u.prof, 0, (size_t)groups * ctas * 12 * 8, st));
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(groups * ctas);
    cfg.blockDim = dim3(UM_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = ctas;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
#define YB_LAUNCH_UMMA(F, C, T, P)                                                                                          \
    do {                                                                                                                    \
        YB_CUDA(cudaFuncSetAttribute(stage1_umma_kernel<F, C, T, P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        YB_CUDA(cudaLaunchKernelEx(&cfg, stage1_umma_kernel<F, C, T, P>, tmA, tmB, u));                                     \
    } while (0)
#define YB_LAUNCH_UMMA_T(F, C)                                                                                              \
    do {                                                                                                                    \
        if (u.prof) { if (tf32) YB_LAUNCH_UMMA(F, C, true, true); else YB_LAUNCH_UMMA(F, C, false, true); }                 \
        else { if (tf32) YB_LAUNCH_UMMA(F, C, true, false); else YB_LAUNCH_UMMA(F, C, false, false); }                      \
    } while (0)
    if (ctas == 2) {
        if (filter) YB_LAUNCH_UMMA_T(true, 2); else YB_LAUNCH_UMMA_T(false, 2);
    } else {
        if (filter) YB_LAUNCH_UMMA_T(true, 1); else YB_LAUNCH_UMMA_T(false, 1);
    }
#undef YB_LAUNCH_UMMA_T
#undef YB_LAUNCH_UMMA
    YB_CUDA(cudaGetLastError());
    if (u.prof) {
        // diagnostics (YAMS_B200_UMMA_PROF=1): aggregate per-role wait cycles
        YB_CUDA(cudaStreamSynchronize(st));
        std::vector<unsigned long long> h((size_t)groups * ctas * 12);
        YB_CUDA(cudaMemcpy(h.data(), u.prof, h.size() * 8, cudaMemcpyDeviceToHost));
        double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        int nlead = 0;
        for (unsigned b = 0; b < groups * ctas; ++b) {
            acc[0] += (double)h[b * 8 + 0];
            acc[3] += (double)h[b * 8 + 3];
            acc[4] += (double)h[b * 8 + 4];
            acc[6] += (double)h[b * 8 + 6];
            acc[7] += (double)h[b * 8 + 7];
            if (h[b * 8 + 5]) { acc[1] += (double)h[b * 8 + 1]; acc[2] += (double)h[b * 8 + 2]; acc[5] += (double)h[b * 8 + 5]; ++nlead; }
        }
        unsigned nb = groups * ctas;
        double det[4] = {0, 0, 0, 0};
        for (unsigned b = 0; b < nb; ++b)
            for (int j = 0; j < 4; ++j) det[j] += (double)h[(size_t)nb * 8 + (size_t)b * 4 + j];
        fprintf(stderr, "[umma prof] epilogue detail per-CTA Mcycles: ld-issue %.2f process %.2f tail(arrive/flush) %.2f head %.2f\n", det[0] / nb / 1e6,
                det[1] / nb / 1e6, det[2] / nb / 1e6, det[3] / nb / 1e6);
        fprintf(stderr, "[umma prof] ctas=%d filter=%d units/group=%llu | per-CTA Mcycles: producer-wait-empty %.2f | mma total %.2f wait-full %.2f wait-tempty %.2f | epilogue total %.2f wait-tfull %.2f tmem-ld-wait %.2f slow-path entries/lane %.0f\n",
                ctas, (int)filter, (unsigned long long)(((uint64_t)u.nrt + groups - 1) / groups * u.nqt), acc[0] / nb / 1e6, acc[5] / nlead / 1e6,
                acc[1] / nlead / 1e6, acc[2] / nlead / 1e6, acc[4] / nb / 1e6, acc[3] / nb / 1e6, acc[6] / nb / 1e6, acc[7] / nb);
        cudaFree(u.prof);
    }
    return YAMS_OK;
}

}  // namespace yb
