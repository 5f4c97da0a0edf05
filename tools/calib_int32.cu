// calib_int32.cu -- calibrates the INT32 issue rates the SHA-256 roofline rests on ("calibrate lanes/clk and clock
// with a microbenchmark on the box").  Measures, per SM and per clock, how many thread-instructions of each class retire when
// every SM runs 1024 threads of independent dependency chains:
//     ALU pipe : LOP3 (xor/and mix), SHF (funnel shift = rotate)           -> the SHA-256 round function
//     FMA pipe : IMAD (x * 1 + y, the form sha256.cu issues its additions in)
//     both     : interleaved LOP3/SHF + IMAD in SHA-256's own ratio (1110 ALU : 693 FMA per block)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/calib_int32 tools/calib_int32.cu ; run on the GPU box.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <vector>

constexpr int kIters = 4096, kChains = 8;

template <int MODE>
__global__ void __launch_bounds__(1024, 1) chains(uint32_t* out, uint32_t seed, unsigned long long* cycles) {
    uint32_t v[kChains], w[kChains];
#pragma unroll
    for (int i = 0; i < kChains; ++i) { v[i] = seed + threadIdx.x * 31 + i; w[i] = seed ^ (threadIdx.x + i * 977); }
    const uint32_t one = seed | 1u;   // opaque 1 (the kernel is launched with seed = 0): keeps IMAD an IMAD
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < kIters; ++it) {
#pragma unroll
        for (int i = 0; i < kChains; ++i) {
            // inline PTX, volatile: ptxas can neither merge two LOP3 of one chain into one nor drop anything
            if (MODE == 0) {          // LOP3 only: 4 per chain step
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(v[i]) : "r"(w[i]), "r"(one));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0xCA;" : "+r"(w[i]) : "r"(v[i]), "r"(one));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0xE8;" : "+r"(v[i]) : "r"(w[i]), "r"(one));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(w[i]) : "r"(v[i]), "r"(one));
            } else if (MODE == 1) {   // SHF only: 4 funnel shifts (rotates) per chain step
                asm volatile("shf.r.wrap.b32 %0, %0, %0, 7;" : "+r"(v[i]));
                asm volatile("shf.r.wrap.b32 %0, %0, %0, 13;" : "+r"(w[i]));
                asm volatile("shf.r.wrap.b32 %0, %0, %1, 3;" : "+r"(v[i]) : "r"(w[i]));
                asm volatile("shf.r.wrap.b32 %0, %0, %1, 19;" : "+r"(w[i]) : "r"(v[i]));
            } else if (MODE == 2) {   // IMAD only: 4 per chain step (x * 1 + y with an opaque 1, as sha256.cu issues its additions)
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(v[i]) : "r"(one), "r"(w[i]));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(w[i]) : "r"(one), "r"(v[i]));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(v[i]) : "r"(one), "r"(w[i]));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(w[i]) : "r"(one), "r"(v[i]));
            } else {                  // SHA-256's mix per chain step: 3 SHF + 2 LOP3 (ALU pipe) and 3 IMAD (FMA pipe)  ~ 1110 : 693
                uint32_t r0, r1, r2;
                asm volatile("shf.r.wrap.b32 %0, %1, %1, 6;" : "=r"(r0) : "r"(v[i]));
                asm volatile("shf.r.wrap.b32 %0, %1, %1, 11;" : "=r"(r1) : "r"(v[i]));
                asm volatile("shf.r.wrap.b32 %0, %1, %1, 25;" : "=r"(r2) : "r"(v[i]));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r0) : "r"(r1), "r"(r2));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0xCA;" : "+r"(w[i]) : "r"(v[i]), "r"(r0));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(v[i]) : "r"(one), "r"(w[i]));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(w[i]) : "r"(one), "r"(r0));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(v[i]) : "r"(one), "r"(w[i]));
            }
        }
    }
    const long long t1 = clock64();
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < kChains; ++i) acc ^= v[i] ^ w[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
}

int main() {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    uint32_t* out;
    unsigned long long* cyc;
    cudaMalloc(&out, (size_t)sms * 1024 * 4);
    cudaMalloc(&cyc, (size_t)sms * 8);
    const char* names[4] = {"LOP3 (ALU pipe)", "SHF rotate (ALU pipe)", "IMAD x*1+y (FMA pipe)", "SHA-256 mix 5 ALU : 3 FMA"};
    const double per_step[4] = {4, 4, 4, 8};   // thread-instructions per chain step (mode 3: 5 on the ALU pipe + 3 on the FMA pipe)
    std::printf("# %d SMs, 1024 threads/SM, %d independent chains/thread, %d iterations\n", sms, kChains, kIters);
    std::printf("| instruction mix | thread-instr / clk / SM | ms | SM clock MHz (cycles / time) |\n|---|---:|---:|---:|\n");
    for (int mode = 0; mode < 4; ++mode) {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) chains<0><<<sms, 1024>>>(out, 0, cyc);
            if (mode == 1) chains<1><<<sms, 1024>>>(out, 0, cyc);
            if (mode == 2) chains<2><<<sms, 1024>>>(out, 0, cyc);
            if (mode == 3) chains<3><<<sms, 1024>>>(out, 0, cyc);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
        }
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        std::vector<unsigned long long> h(sms);
        cudaMemcpy(h.data(), cyc, (size_t)sms * 8, cudaMemcpyDeviceToHost);
        double mean = 0;
        for (auto c : h) mean += (double)c;
        mean /= sms;
        const double instr = per_step[mode] * kChains * (double)kIters * 1024.0;
        std::printf("| %s | %.1f | %.3f | %.0f |\n", names[mode], instr / mean, ms, mean / (ms * 1e3));
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { std::printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
