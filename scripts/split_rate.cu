// split_rate.cu -- dev micro-benchmark: instruction throughput of the bf16 hi / lo operand split (cvt.rn.bf16x2.f32 based, as
// in the kernels) against integer-arithmetic variants, per SM.  nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t pack_rn(float a, float b) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
}
// the kernels' split: two F2FP, two expands, two subtractions
__device__ __forceinline__ void split_cvt(float a, float b, uint32_t& hi, uint32_t& lo) {
    hi = pack_rn(a, b);
    const float ha = __uint_as_float(hi << 16), hb = __uint_as_float(hi & 0xFFFF0000u);
    lo = pack_rn(a - ha, b - hb);
}
// truncating hi (PRMT), round-to-nearest lo (one F2FP)
__device__ __forceinline__ void split_trunc_hi(float a, float b, uint32_t& hi, uint32_t& lo) {
    const uint32_t ua = __float_as_uint(a), ub = __float_as_uint(b);
    hi = __byte_perm(ua, ub, 0x7632);
    lo = pack_rn(a - __uint_as_float(ua & 0xFFFF0000u), b - __uint_as_float(ub & 0xFFFF0000u));
}
// no F2FP at all: truncating hi and lo
__device__ __forceinline__ void split_trunc_both(float a, float b, uint32_t& hi, uint32_t& lo) {
    const uint32_t ua = __float_as_uint(a), ub = __float_as_uint(b);
    hi = __byte_perm(ua, ub, 0x7632);
    const float la = a - __uint_as_float(ua & 0xFFFF0000u), lb = b - __uint_as_float(ub & 0xFFFF0000u);
    lo = __byte_perm(__float_as_uint(la), __float_as_uint(lb), 0x7632);
}

template <int MODE>
__global__ void __launch_bounds__(256) rate_kernel(int iters, float seed, uint32_t* out) {
    float x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = seed * (threadIdx.x + 1) + i;
    uint32_t acc = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
            uint32_t h, l;
            if (MODE == 0) split_cvt(x[i], x[i + 1], h, l);
            else if (MODE == 1) split_trunc_hi(x[i], x[i + 1], h, l);
            else if (MODE == 2) split_trunc_both(x[i], x[i + 1], h, l);
            else { h = __float_as_uint(x[i] * 1.0001f + x[i + 1]); l = __float_as_uint(x[i + 1] * 0.9999f - x[i]); }   // 2 FFMA reference
            acc ^= h + l;
            x[i] = __uint_as_float((h & 0x007FFFFFu) | 0x3F800000u);          // keep the chain data dependent, values sane
            x[i + 1] = __uint_as_float((l & 0x007FFFFFu) | 0x3F800000u);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    uint32_t* out;
    cudaMalloc(&out, (size_t)sms * 8 * 256 * 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int iters = 20000;
    const char* names[4] = {"cvt.rn x2 (kernels)", "prmt hi + cvt.rn lo", "prmt hi + prmt lo", "2 FFMA (reference)"};
    for (int mode = 0; mode < 4; ++mode) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) rate_kernel<0><<<sms * 8, 256>>>(iters, 1.37f, out);
            if (mode == 1) rate_kernel<1><<<sms * 8, 256>>>(iters, 1.37f, out);
            if (mode == 2) rate_kernel<2><<<sms * 8, 256>>>(iters, 1.37f, out);
            if (mode == 3) rate_kernel<3><<<sms * 8, 256>>>(iters, 1.37f, out);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
        }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        // per SM: 8 CTAs x 8 warps, iters x 4 pair-splits per warp
        const double splits_per_sm = 8.0 * 8.0 * iters * 4.0;
        const double clk = ms * 1e-3 * 1.9e9;
        printf("%-22s %8.3f ms   %.2f cycles per warp-level pair split per SM (all 64 warps)\n", names[mode], ms, clk / splits_per_sm);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
