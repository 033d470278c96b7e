// multimem_allreduce.cu -- the library's own all-reduce of the per-Gaussian feature gradient over the NVSwitch multicast
// mapping (opt-in; seganygaussians_b200/data_parallel.py `MulticastAllReduce`, bench.py --allreduce multimem).
//
// Every rank holds the gradient in a buffer that is part of ONE symmetric allocation mapped twice: at its own address and at a
// multicast address that names the same offset on all N GPUs.  Rank r owns the r-th slice of the tensor:
//     multimem.ld_reduce.add  [mc + i]   -- the switch reads element i from all N GPUs and returns their sum (in-switch reduction:
//                                           the bytes of the slice cross this GPU's link ONCE, not N - 1 times)
//     multimem.st             [mc + i]   -- the switch writes the sum to all N GPUs
// so one pass over 1/N of the tensor per GPU is the whole all-reduce: 2 x 16 B of link traffic per 16 B owned.  The caller
// separates it from producers / consumers of the buffer with the symmetric allocation's device barriers (before: every rank's
// gradient is complete; after: every slice has been written everywhere).  Accumulation order inside the switch is fixed per
// address, so all ranks receive bit-identical sums (NCCL's ring gives that too; a pull-based P2P reduction would as well).
//
// Why this and not a reduction fused into the backward kernel's red.global traffic: DESIGN.md section 6 (that would push the
// 1.1 GB of partial sums through NVLink instead of the 128 MB of finished ones).
#include "common.cuh"

namespace sagars {

__global__ void __launch_bounds__(512)
multimem_allreduce_f32_kernel(float* __restrict__ mc, size_t quad_begin, size_t quad_end)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t q = quad_begin + (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < quad_end; q += stride) {
        float* p = mc + 4 * q;
        float x, y, z, w;
        asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(x), "=f"(y), "=f"(z), "=f"(w) : "l"(p) : "memory");
        asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};"
                     :: "l"(p), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
    }
}

}  // namespace sagars

using namespace sagars;

extern "C" int sagars_multimem_allreduce_f32(int32_t device, void* multicast_ptr, int64_t numel, int32_t rank, int32_t world,
                                             void* stream)
{
    if (!multicast_ptr || numel < 0 || (numel & 3) || world < 1 || rank < 0 || rank >= world || ((uintptr_t)multicast_ptr & 15)) {
        set_error("sagars_multimem_allreduce_f32: bad argument (numel %% 4 == 0, 16-byte aligned multicast pointer, 0 <= rank < world)");
        return SAGARS_EINVAL;
    }
    if (numel == 0) return SAGARS_OK;
    SAGARS_CUDA(cudaSetDevice(device));
    const size_t quads = (size_t)numel / 4;
    const size_t per = (quads + (size_t)world - 1) / (size_t)world;
    const size_t begin = per * (size_t)rank < quads ? per * (size_t)rank : quads;
    const size_t end = begin + per < quads ? begin + per : quads;
    if (end > begin) {
        // a modest grid: the pass is bound by the link, not by issue -- and the SMs stay free for the kernels it overlaps with
        const size_t want = (end - begin + 511) / 512;
        const int grid = (int)(want < 148 ? want : 148);
        multimem_allreduce_f32_kernel<<<grid, 512, 0, (cudaStream_t)stream>>>((float*)multicast_ptr, begin, end);
        SAGARS_LAUNCH_CHECK((cudaStream_t)stream, false);
    }
    return SAGARS_OK;
}
