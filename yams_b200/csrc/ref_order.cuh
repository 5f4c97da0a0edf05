// ref_order.cuh -- the float distance arithmetic of sqlite-vec-cpp in the exact operation ORDER of the reference build, so
// that the pairwise / batch / vec0 operators return the reference's bits, not merely its value within a tolerance.
//
// YAMS compiles sqlite-vec-cpp with -mavx -mfma -mavx2 -DSQLITE_VEC_ENABLE_AVX (src/vector/meson.build:79-87; the FMA
// macro is not set).  What that build evaluates (pinned against the reference compiled in place, tests/test_oracle_pin.py):
//   l2_distance<float>      d >= 16 && d % 16 == 0: simd/avx.hpp:20-66 -- eight lane-strided partial sums, diff*diff and the
//                           add are SEPARATE roundings, lanes summed left to right; else the scalar loop l2.hpp:108-118
//                           (also separate roundings);
//   cosine_distance<float>  d >= 8: simd/avx.hpp:111-171 -- eight lane-strided partial sums of a.b, a.a, b.b with the multiply
//                           and add FUSED (the compiler contracts avx_fmadd_ps' mul+add under -mfma), horizontal reduction
//                           (l4+l0 + l5+l1) + (l6+l2 + l7+l3), the d % 8 tail fused as well; d < 8: the scalar loop
//                           cosine.hpp:48-69, fused; denom < 1e-8 -> 1.0;
//   l1_distance<float>      d >= 8: simd/avx.hpp:66-104 -- eight lane-strided float sums of |a-b|, the same horizontal
//                           reduction, float tail; d < 8: the double-accumulating loop l1.hpp:33-42, cast to float.
// One thread evaluates one pair: the partial sums live in registers.
#pragma once
#include <stdint.h>

namespace yb {

template <typename LA, typename LB>
__device__ __forceinline__ float ref_l2_distance(LA a, LB b, uint32_t d) {
    float sum;
    if (d >= 16 && d % 16 == 0) {
        float p[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (uint32_t c = 0; c < d; c += 8) {
#pragma unroll
            for (int l = 0; l < 8; ++l) {
                const float df = __fsub_rn(a(c + l), b(c + l));
                p[l] = __fadd_rn(p[l], __fmul_rn(df, df));
            }
        }
        sum = p[0];
#pragma unroll
        for (int l = 1; l < 8; ++l) sum = __fadd_rn(sum, p[l]);
    } else {
        sum = 0.f;
        for (uint32_t c = 0; c < d; ++c) {
            const float df = __fsub_rn(a(c), b(c));
            sum = __fadd_rn(sum, __fmul_rn(df, df));
        }
    }
    return sqrtf(sum);
}

template <typename LA, typename LB>
__device__ __forceinline__ float ref_cosine_distance(LA a, LB b, uint32_t d) {
    float dot = 0.f, am = 0.f, bm = 0.f;
    uint32_t e = 0;
    if (d >= 8) {
        float pd[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, pa[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f},
              pb[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        e = d & ~7u;
        for (uint32_t c = 0; c < e; c += 8) {
#pragma unroll
            for (int l = 0; l < 8; ++l) {
                const float x = a(c + l), y = b(c + l);
                pd[l] = __fmaf_rn(x, y, pd[l]);
                pa[l] = __fmaf_rn(x, x, pa[l]);
                pb[l] = __fmaf_rn(y, y, pb[l]);
            }
        }
        auto hsum = [](const float (&p)[8]) {
            const float c0 = __fadd_rn(p[4], p[0]), c1 = __fadd_rn(p[5], p[1]), c2 = __fadd_rn(p[6], p[2]), c3 = __fadd_rn(p[7], p[3]);
            return __fadd_rn(__fadd_rn(c0, c1), __fadd_rn(c2, c3));
        };
        dot = hsum(pd);
        am = hsum(pa);
        bm = hsum(pb);
    }
    for (uint32_t c = e; c < d; ++c) {
        const float x = a(c), y = b(c);
        dot = __fmaf_rn(x, y, dot);
        am = __fmaf_rn(x, x, am);
        bm = __fmaf_rn(y, y, bm);
    }
    const float denom = __fmul_rn(sqrtf(am), sqrtf(bm));
    if (denom < 1e-8f) return 1.0f;
    return __fsub_rn(1.0f, __fdiv_rn(dot, denom));
}

template <typename LA, typename LB>
__device__ __forceinline__ float ref_l1_distance(LA a, LB b, uint32_t d) {
    if (d < 8) {
        double sum = 0.0;
        for (uint32_t c = 0; c < d; ++c) sum += fabs((double)a(c) - (double)b(c));
        return (float)sum;
    }
    float p[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const uint32_t e = d & ~7u;
    for (uint32_t c = 0; c < e; c += 8) {
#pragma unroll
        for (int l = 0; l < 8; ++l) p[l] = __fadd_rn(p[l], fabsf(__fsub_rn(a(c + l), b(c + l))));
    }
    const float c0 = __fadd_rn(p[4], p[0]), c1 = __fadd_rn(p[5], p[1]), c2 = __fadd_rn(p[6], p[2]), c3 = __fadd_rn(p[7], p[3]);
    float r = __fadd_rn(__fadd_rn(c0, c1), __fadd_rn(c2, c3));
    for (uint32_t c = e; c < d; ++c) r = __fadd_rn(r, fabsf(__fsub_rn(a(c), b(c))));
    return r;
}

struct F32At {
    const float* p;
    __device__ __forceinline__ float operator()(uint32_t i) const { return p[i]; }
};

}  // namespace yb
