// mma_probe.cu -- (1) checks the m16n8k8 TF32 mma.sync fragment layout against a scalar reference,
// (2) checks the 3xTF32 split accuracy, (3) measures mma.sync throughput on this GPU.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cmath>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t f2tf32(float x) { uint32_t r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x)); return r; }
__device__ __forceinline__ void mma_tf32(float* d, const uint32_t* a, const uint32_t* b)
{
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// A[16][8] row-major, B[8][8] (k x n), D[16][8]
__global__ void layout_kernel(const float* A, const float* B, float* D, float* D3)
{
    int lane = threadIdx.x, g = lane >> 2, t = lane & 3;
    float a[4] = {A[g * 8 + t], A[(g + 8) * 8 + t], A[g * 8 + t + 4], A[(g + 8) * 8 + t + 4]};
    float b[2] = {B[t * 8 + g], B[(t + 4) * 8 + g]};
    uint32_t ah[4], al[4], bh[2], bl[2];
    for (int i = 0; i < 4; i++) { ah[i] = f2tf32(a[i]); al[i] = f2tf32(a[i] - __uint_as_float(ah[i])); }
    for (int i = 0; i < 2; i++) { bh[i] = f2tf32(b[i]); bl[i] = f2tf32(b[i] - __uint_as_float(bh[i])); }
    float d[4] = {0, 0, 0, 0}, d3[4] = {0, 0, 0, 0};
    mma_tf32(d, ah, bh);
    mma_tf32(d3, al, bh); mma_tf32(d3, ah, bl); mma_tf32(d3, ah, bh);
    D[g * 8 + 2 * t] = d[0]; D[g * 8 + 2 * t + 1] = d[1]; D[(g + 8) * 8 + 2 * t] = d[2]; D[(g + 8) * 8 + 2 * t + 1] = d[3];
    D3[g * 8 + 2 * t] = d3[0]; D3[g * 8 + 2 * t + 1] = d3[1]; D3[(g + 8) * 8 + 2 * t] = d3[2]; D3[(g + 8) * 8 + 2 * t + 1] = d3[3];
}

template <int ILP>
__global__ void tput_kernel(float* out, int iters)
{
    float d[ILP][4];
    uint32_t a[4] = {0x3f800000u, 0x3f800000u, 0x3f800000u, 0x3f800000u}, b[2] = {0x3f800000u, 0x3f000000u};
    for (int i = 0; i < ILP; i++) for (int j = 0; j < 4; j++) d[i][j] = threadIdx.x * 1e-9f;
    for (int it = 0; it < iters; it++)
#pragma unroll
        for (int i = 0; i < ILP; i++) mma_tf32(d[i], a, b);
    float s = 0;
    for (int i = 0; i < ILP; i++) for (int j = 0; j < 4; j++) s += d[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main()
{
    float hA[128], hB[64], hD[128], hD3[128], ref[128];
    srand(1);
    for (int i = 0; i < 128; i++) hA[i] = (rand() / (float)RAND_MAX - 0.5f) * 3.f;
    for (int i = 0; i < 64; i++) hB[i] = (rand() / (float)RAND_MAX - 0.5f) * 3.f;
    for (int m = 0; m < 16; m++) for (int n = 0; n < 8; n++) { double s = 0; for (int k = 0; k < 8; k++) s += (double)hA[m * 8 + k] * hB[k * 8 + n]; ref[m * 8 + n] = (float)s; }
    float *dA, *dB, *dD, *dD3;
    cudaMalloc(&dA, 512); cudaMalloc(&dB, 256); cudaMalloc(&dD, 512); cudaMalloc(&dD3, 512);
    cudaMemcpy(dA, hA, 512, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, 256, cudaMemcpyHostToDevice);
    layout_kernel<<<1, 32>>>(dA, dB, dD, dD3);
    cudaMemcpy(hD, dD, 512, cudaMemcpyDeviceToHost); cudaMemcpy(hD3, dD3, 512, cudaMemcpyDeviceToHost);
    double e1 = 0, e3 = 0, mx = 0;
    for (int i = 0; i < 128; i++) { e1 = fmax(e1, fabs(hD[i] - ref[i])); e3 = fmax(e3, fabs(hD3[i] - ref[i])); mx = fmax(mx, fabs(ref[i])); }
    printf("layout: max|ref|=%.3f  1xTF32 max err=%.3e  3xTF32 max err=%.3e  (%s)\n", mx, e1, e3, (e1 < 2e-2 && e3 < 2e-5) ? "LAYOUT OK" : "LAYOUT WRONG");
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) printf("cuda error %s\n", cudaGetErrorString(err));

    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float* out; cudaMalloc(&out, sms * 8 * 1024 * 4);
    cudaEvent_t e0, e1e; cudaEventCreate(&e0); cudaEventCreate(&e1e);
    for (int warps = 4; warps <= 32; warps *= 2) {
        int iters = 20000;
        tput_kernel<4><<<sms, warps * 32>>>(out, 100);
        cudaDeviceSynchronize();
        cudaEventRecord(e0);
        tput_kernel<4><<<sms, warps * 32>>>(out, iters);
        cudaEventRecord(e1e); cudaEventSynchronize(e1e);
        float ms; cudaEventElapsedTime(&ms, e0, e1e);
        double mmas = (double)sms * warps * iters * 4;
        printf("tput: %2d warps/SM  %.3f ms  %.1f TFLOP/s tf32 (m16n8k8)  = %.2f mma/clk/SM @1.965GHz\n", warps, ms, mmas * 2 * 16 * 8 * 8 / ms / 1e9,
               mmas / sms / (ms * 1e-3 * 1.965e9));
    }
    return 0;
}
