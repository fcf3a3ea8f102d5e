// gemm_fp32.cu -- plain fp32 FFMA GEMMs (mode 0 of vgtkb_gemm_nt / vgtkb_gemm_tn).
//
// This is the CUDA-core formulation of BasicSO3Conv.forward (vgtk/vgtk/so3conv/modules.py:48-55:
// W[Co, Ci*K] @ x) and of its weight gradient.  It is the bring-up / verification path and the
// fallback for shapes the tcgen05 kernels (gemm_tc.cu) do not take (tiny K, unaligned rows);
// the tensor-core path is the one the backbone uses on shapes that matter.
#include "common.cuh"

namespace vgtkb {

constexpr int GB = 128;   // tile edge (M and N)
constexpr int GK = 8;     // k-slab
constexpr int GT = 256;   // threads, 16x16, 8x8 outputs each
constexpr int GP = 4;     // smem row padding (floats)

__device__ __forceinline__ void micro_kernel(const float (*As)[GB + GP], const float (*Bs)[GB + GP], int ty, int tx,
                                             float (&acc)[8][8]) {
#pragma unroll
    for (int k = 0; k < GK; ++k) {
        float a[8], b[8];
        const float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
        const float4 a1 = *reinterpret_cast<const float4*>(&As[k][64 + ty * 4]);
        const float4 b0 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
        const float4 b1 = *reinterpret_cast<const float4*>(&Bs[k][64 + tx * 4]);
        a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
        b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
}

// thread (ty,tx) owns rows {ty*4..+3, 64+ty*4..+3} and cols {tx*4..+3, 64+tx*4..+3}
__device__ __forceinline__ int own_row(int ty, int i) { return (i < 4 ? 0 : 64) + ty * 4 + (i & 3); }

// C[M,N] = A[M,K] * B[N,K]^T + bias[N]
__global__ void __launch_bounds__(GT)
sgemm_nt_kernel(int64_t M, int N, int K, const float* __restrict__ A, const float* __restrict__ B,
                const float* __restrict__ bias, float* __restrict__ C) {
    __shared__ __align__(16) float As[GK][GB + GP];
    __shared__ __align__(16) float Bs[GK][GB + GP];
    const int64_t m0 = (int64_t)blockIdx.x * GB;
    const int n0 = blockIdx.y * GB;
    const int t = threadIdx.x, ty = t / 16, tx = t % 16;
    const int lr = t / 2, lk = (t % 2) * 4;  // loader: row lr, k offset lk..lk+3
    const bool vec = (K % 4 == 0) && ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B)) & 15) == 0;
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    for (int k0 = 0; k0 < K; k0 += GK) {
        float av[4] = {0.f, 0.f, 0.f, 0.f}, bv[4] = {0.f, 0.f, 0.f, 0.f};
        const int64_t ar = m0 + lr;
        const int br = n0 + lr;
        if (vec) {
            if (ar < M && k0 + lk < K) *reinterpret_cast<float4*>(av) = *reinterpret_cast<const float4*>(A + ar * K + k0 + lk);
            if (br < N && k0 + lk < K) *reinterpret_cast<float4*>(bv) = *reinterpret_cast<const float4*>(B + (int64_t)br * K + k0 + lk);
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (ar < M && k0 + lk + i < K) av[i] = A[ar * K + k0 + lk + i];
                if (br < N && k0 + lk + i < K) bv[i] = B[(int64_t)br * K + k0 + lk + i];
            }
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            As[lk + i][lr] = av[i];
            Bs[lk + i][lr] = bv[i];
        }
        __syncthreads();
        micro_kernel(As, Bs, ty, tx, acc);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int64_t r = m0 + own_row(ty, i);
        if (r >= M) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int cidx = n0 + own_row(tx, j);
            if (cidx < N) C[r * N + cidx] = acc[i][j] + (bias ? bias[cidx] : 0.f);
        }
    }
}

// C[M,N] += sum_{r in slice} A[r,M]^T B[r,N]   (split over grid.z; atomics into C)
__global__ void __launch_bounds__(GT)
sgemm_tn_kernel(int M, int N, int64_t R, int64_t rows_per_split, const float* __restrict__ A,
                const float* __restrict__ B, float* __restrict__ C) {
    __shared__ __align__(16) float As[GK][GB + GP];
    __shared__ __align__(16) float Bs[GK][GB + GP];
    const int m0 = blockIdx.x * GB, n0 = blockIdx.y * GB;
    const int64_t r_begin = (int64_t)blockIdx.z * rows_per_split;
    const int64_t r_end = min(R, r_begin + rows_per_split);
    const int t = threadIdx.x, ty = t / 16, tx = t % 16;
    const int lr = t / 32, lc = (t % 32) * 4;  // loader: slab row lr (0..7), columns lc..lc+3
    const bool vecA = (M % 4 == 0) && (reinterpret_cast<uintptr_t>(A) & 15) == 0;
    const bool vecB = (N % 4 == 0) && (reinterpret_cast<uintptr_t>(B) & 15) == 0;
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    for (int64_t r0 = r_begin; r0 < r_end; r0 += GK) {
        float av[4] = {0.f, 0.f, 0.f, 0.f}, bv[4] = {0.f, 0.f, 0.f, 0.f};
        const int64_t r = r0 + lr;
        if (r < r_end) {
            if (vecA && m0 + lc + 3 < M) *reinterpret_cast<float4*>(av) = *reinterpret_cast<const float4*>(A + r * M + m0 + lc);
            else
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (m0 + lc + i < M) av[i] = A[r * M + m0 + lc + i];
            if (vecB && n0 + lc + 3 < N) *reinterpret_cast<float4*>(bv) = *reinterpret_cast<const float4*>(B + r * N + n0 + lc);
            else
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (n0 + lc + i < N) bv[i] = B[r * N + n0 + lc + i];
        }
        __syncthreads();
        *reinterpret_cast<float4*>(&As[lr][lc]) = *reinterpret_cast<float4*>(av);
        *reinterpret_cast<float4*>(&Bs[lr][lc]) = *reinterpret_cast<float4*>(bv);
        __syncthreads();
        micro_kernel(As, Bs, ty, tx, acc);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int r = m0 + own_row(ty, i);
        if (r >= M) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int cidx = n0 + own_row(tx, j);
            if (cidx < N) atomicAdd(C + (int64_t)r * N + cidx, acc[i][j]);
        }
    }
}

// K == 1 (the first layer's Ci = 1 skip conv): C[m, n] = fma(A[m], B[n], 0) + bias[n], a write stream of M*N floats.  The tiled
// kernel spends its time on empty k-slabs there (47 us for a 63 MB output).  Same arithmetic as the tiled kernel: one fmaf onto a
// zero accumulator, then the bias.
__global__ void __launch_bounds__(256)
sgemm_nt_rank1_kernel(int64_t M, int N4, const float* __restrict__ A, const float4* __restrict__ B, const float4* __restrict__ bias,
                      float4* __restrict__ C) {
    const int64_t total = M * N4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t m = i / N4;
        const int n4 = (int)(i - m * N4);
        const float a = __ldg(A + m);
        const float4 b = __ldg(B + n4);
        float4 o = make_float4(fmaf(a, b.x, 0.f), fmaf(a, b.y, 0.f), fmaf(a, b.z, 0.f), fmaf(a, b.w, 0.f));
        if (bias != nullptr) {
            const float4 bb = __ldg(bias + n4);
            o.x += bb.x; o.y += bb.y; o.z += bb.z; o.w += bb.w;
        }
        C[i] = o;
    }
}

int sgemm_nt(int64_t M, int N, int K, const float* A, const float* B, const float* bias, float* C, cudaStream_t st) {
    if (K == 1 && N % 4 == 0 &&
        ((reinterpret_cast<uintptr_t>(B) | reinterpret_cast<uintptr_t>(C) | reinterpret_cast<uintptr_t>(bias)) & 15) == 0) {
        const int64_t total = M * (N / 4);
        const unsigned grid = (unsigned)(ceil_div64(total, 256) < (int64_t)kNumSMs * 16 ? ceil_div64(total, 256) : kNumSMs * 16);
        sgemm_nt_rank1_kernel<<<grid, 256, 0, st>>>(M, N / 4, A, reinterpret_cast<const float4*>(B),
                                                   reinterpret_cast<const float4*>(bias), reinterpret_cast<float4*>(C));
        return check_launch("gemm_nt(fp32, K = 1)");
    }
    dim3 grid((unsigned)ceil_div64(M, GB), ceil_div(N, GB));
    VGTKB_REQUIRE(grid.y <= 65535, "gemm_nt: N too large");
    sgemm_nt_kernel<<<grid, GT, 0, st>>>(M, N, K, A, B, bias, C);
    return check_launch("gemm_nt(fp32)");
}

// C[m, n] += sum_r A[r, m] B[r, n] for a very narrow B (N <= 4: the first layer's Ci = 1 skip conv):
// thread <-> column m of A, a CTA owns a slice of rows, one atomic per (CTA, m, n)
constexpr int NARROW_T = 256;
__global__ void __launch_bounds__(NARROW_T)
sgemm_tn_narrow_kernel(int M, int N, int64_t R, int64_t rows_per_cta, const float* __restrict__ A,
                       const float* __restrict__ B, float* __restrict__ C) {
    const int tx = min(M, NARROW_T), ty = NARROW_T / tx;
    const int lx = threadIdx.x % tx, ly = threadIdx.x / tx;
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_cta;
    const int64_t r1 = min(R, r0 + rows_per_cta);
    __shared__ float red[NARROW_T][4];
    for (int m = lx; m < M; m += tx) {
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        if (ly < ty)
            for (int64_t r = r0 + ly; r < r1; r += ty) {
                const float av = A[r * M + m];
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (j < N) acc[j] = fmaf(av, B[r * N + j], acc[j]);
            }
#pragma unroll
        for (int j = 0; j < 4; ++j) red[threadIdx.x][j] = acc[j];
        __syncthreads();
        if (ly == 0) {
            for (int q = 1; q < ty; ++q)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[j] += red[q * tx + lx][j];
            for (int j = 0; j < N; ++j) atomicAdd(C + (size_t)m * N + j, acc[j]);
        }
        __syncthreads();
    }
}

int sgemm_tn(int M, int N, int64_t R, const float* A, const float* B, float* C, int accumulate, cudaStream_t st) {
    if (!accumulate) VGTKB_CUDA(cudaMemsetAsync(C, 0, sizeof(float) * (size_t)M * N, st));
    if (N <= 4 && M <= 4096) {
        int64_t rpc = ceil_div64(R, (int64_t)kNumSMs * 4);
        if (rpc < 64) rpc = 64;
        const unsigned grid = (unsigned)ceil_div64(R, rpc);
        sgemm_tn_narrow_kernel<<<grid, NARROW_T, 0, st>>>(M, N, R, rpc, A, B, C);
        return check_launch("gemm_tn(fp32, narrow)");
    }
    const int tiles = ceil_div(M, GB) * ceil_div(N, GB);
    int64_t splits = ceil_div64((int64_t)kNumSMs * 4, tiles);
    if (splits < 1) splits = 1;
    int64_t rps = ceil_div64(R, splits);
    rps = ceil_div64(rps, GK) * GK;
    if (rps < 256) rps = 256;
    splits = ceil_div64(R, rps);
    VGTKB_REQUIRE(splits <= 65535, "gemm_tn: too many splits");
    dim3 grid(ceil_div(M, GB), ceil_div(N, GB), (unsigned)splits);
    VGTKB_REQUIRE(grid.y <= 65535, "gemm_tn: N too large");
    sgemm_tn_kernel<<<grid, GT, 0, st>>>(M, N, R, rps, A, B, C);
    return check_launch("gemm_tn(fp32)");
}

}  // namespace vgtkb
