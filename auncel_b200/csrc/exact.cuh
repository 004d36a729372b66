// The reference's distance arithmetic, reproduced bit for bit on the device.
//
// fvec_L2sqr / fvec_inner_product in the reference's default (-msse4, no FMA) build
// (/root/reference/Auncel/utils_simd.cpp:391-443) keep four lane accumulators; lane l sums
// the elements i == l (mod 4) in index order with a separately rounded multiply and add,
// and the result is (s0+s1)+(s2+s3) (two _mm_hadd_ps).  Rows are zero-padded to a multiple
// of 4 floats, which is what masked_read does for the tail (utils_simd.cpp:118-136); adding
// the resulting +0 terms is exact.  One float4 step of one (query, vector) pair:
#pragma once
#include "errmodel.h"

namespace auncel {

template <int METRIC>
__device__ __forceinline__ void exact_step(float (&s)[4], const float4& x, const float4& y) {
    if (METRIC == METRIC_L2) {
        float t0 = __fsub_rn(x.x, y.x), t1 = __fsub_rn(x.y, y.y);
        float t2 = __fsub_rn(x.z, y.z), t3 = __fsub_rn(x.w, y.w);
        s[0] = __fadd_rn(s[0], __fmul_rn(t0, t0));
        s[1] = __fadd_rn(s[1], __fmul_rn(t1, t1));
        s[2] = __fadd_rn(s[2], __fmul_rn(t2, t2));
        s[3] = __fadd_rn(s[3], __fmul_rn(t3, t3));
    } else {
        s[0] = __fadd_rn(s[0], __fmul_rn(x.x, y.x));
        s[1] = __fadd_rn(s[1], __fmul_rn(x.y, y.y));
        s[2] = __fadd_rn(s[2], __fmul_rn(x.z, y.z));
        s[3] = __fadd_rn(s[3], __fmul_rn(x.w, y.w));
    }
}

__device__ __forceinline__ float exact_finish(const float (&s)[4]) {
    return __fadd_rn(__fadd_rn(s[0], s[1]), __fadd_rn(s[2], s[3]));
}

}  // namespace auncel
