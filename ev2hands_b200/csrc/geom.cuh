// The reference's expanded squared distance (src/Ev2Hands/model/pointnet2_utils.py:19-40), rounding for rounding:
// ball-query membership and the 3-NN order of feature propagation both hinge on these exact bits.
#pragma once
#include <cuda_runtime.h>

namespace ev2h {

__device__ __forceinline__ float sq_norm3(float x, float y, float z) {
    // torch.sum(v ** 2, -1): (x*x + y*y) + z*z, nothing fused
    return __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
}

__device__ __forceinline__ float sqdist_expanded(float qx, float qy, float qz, float qn, const float4 p) {
    float dot = __fmul_rn(qx, p.x);          // sgemm with K = 3: x*x', then two FMAs
    dot = __fmaf_rn(qy, p.y, dot);
    dot = __fmaf_rn(qz, p.z, dot);
    float t = __fmul_rn(-2.0f, dot);         // dist = -2 * matmul           (:37)
    t = __fadd_rn(t, qn);                    // dist += sum(src**2)          (:38)
    t = __fadd_rn(t, p.w);                   // dist += sum(dst**2)          (:39)
    return t;
}

}  // namespace ev2h
