// Host restatement of seganygaussians_b200/csrc/mma.cuh for the CPU execution shim.  TEST INFRASTRUCTURE ONLY.
// mma.sync.m16n8k8 (tf32 x tf32 -> f32) is warp-collective: every lane posts its A / B fragments, then computes its own four
// outputs from the assembled 16x8 and 8x8 matrices.  Operands are read as the tensor core reads them (top 19 bits), products
// and the 8-term sum are formed in double and rounded to fp32 once -- at least as accurate as the hardware's accumulation.
#pragma once
#include <cstring>

namespace sagars {

inline void split_tf32(float x, uint32_t& hi, uint32_t& lo)
{
    hi = __float_as_uint(x) & 0xFFFFE000u;
    lo = __float_as_uint(x - __uint_as_float(hi));
}

inline void mma_16n8k8(float* d, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1)
{
    static uint32_t frag[64][32][6];     // per warp of the block (blocks run one at a time), per lane: a0..a3, b0, b1
    const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t* mine = frag[warp][lane];
    mine[0] = a0; mine[1] = a1; mine[2] = a2; mine[3] = a3; mine[4] = b0; mine[5] = b1;
    __syncwarp();
    auto tf = [](uint32_t v) { return (double)__uint_as_float(v & 0xFFFFE000u); };
    // A[row][k]: a0 = (g, t), a1 = (g + 8, t), a2 = (g, t + 4), a3 = (g + 8, t + 4) of lane 4 g + t;  B[k][n]: b0 = (t, g), b1 = (t + 4, g)
    auto A = [&](int row, int k) { const uint32_t* f = frag[warp][4 * (row & 7) + (k & 3)]; return tf(f[(row >> 3) + 2 * (k >> 2)]); };
    auto B = [&](int k, int n) { const uint32_t* f = frag[warp][4 * n + (k & 3)]; return tf(f[4 + (k >> 2)]); };
    const int g = lane >> 2, t = lane & 3;
    const int rows[4] = {g, g, g + 8, g + 8}, cols[4] = {2 * t, 2 * t + 1, 2 * t, 2 * t + 1};
    float out[4];
    for (int i = 0; i < 4; i++) {
        double acc = (double)d[i];
        for (int k = 0; k < 8; k++) acc += A(rows[i], k) * B(k, cols[i]);
        out[i] = (float)acc;
    }
    __syncwarp();                        // every lane has read the fragments before the next mma overwrites them
    for (int i = 0; i < 4; i++) d[i] = out[i];
}

inline float rcp_approx(float x) { return 1.0f / x; }   // rcp.approx.ftz: <= 1 ulp on the device

}  // namespace sagars
