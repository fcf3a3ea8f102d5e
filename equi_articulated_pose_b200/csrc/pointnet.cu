// pointnet.cu -- PointnetSO3Conv pooling head for sm_100a (vgtk/vgtk/so3conv/modules.py:376-413).
//
// The reference concatenates the anchor-rotated, centred coordinates to the features ([B, C+3, N, A], a copy of the
// whole feature tensor), runs a 1x1 Conv2d and takes the max over the points.  Here the feature part of the 1x1 conv
// is the tcgen05 contraction on the channels-last rows (vgtkb_gemm_nt, bias folded in), and the coordinate part is
// three FMAs per output folded into the pooling pass:
//     out[b, o, a] = max_n ( e[b, n, a, o] + v[a, o, :] . xc[b, :, n] ),   v[a, o, j] = sum_i W_x[o, i] anchors[a, j, i]
// e is read exactly once (HBM-bound); the arg-max is kept for the backward scatter.
#include "common.cuh"

namespace vgtkb {

constexpr int PN_THREADS = 128;

// grid (ceil(co/128), a, b): thread <-> output channel, loop over the points
template <bool POOL>
__global__ void __launch_bounds__(PN_THREADS)
pointnet_kernel(int n, int na, int co, float* __restrict__ e, const float* __restrict__ v, const float* __restrict__ xc,
                float* __restrict__ out, int32_t* __restrict__ arg) {
    const int o = blockIdx.x * PN_THREADS + threadIdx.x, a = blockIdx.y, b = blockIdx.z;
    if (o >= co) return;
    const float v0 = v[((size_t)a * co + o) * 3 + 0], v1 = v[((size_t)a * co + o) * 3 + 1], v2 = v[((size_t)a * co + o) * 3 + 2];
    const float* x = xc + (size_t)b * 3 * n;
    float* row = e + (((size_t)b * n) * na + a) * co + o;
    const size_t stride = (size_t)na * co;
    float best = 0.f;
    int besti = 0;
    constexpr int U = 8;
    for (int i0 = 0; i0 < n; i0 += U) {
        float val[U];
#pragma unroll
        for (int u = 0; u < U; ++u) val[u] = i0 + u < n ? row[(size_t)(i0 + u) * stride] : 0.f;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int i = i0 + u;
            if (i < n) {
                const float t = val[u] + __fmaf_rn(v2, x[2 * n + i], __fmaf_rn(v1, x[n + i], __fmul_rn(v0, x[i])));
                if (POOL) {
                    const bool take = i == 0 || t > best;          // first maximum, like torch.max
                    best = take ? t : best;
                    besti = take ? i : besti;
                } else {
                    row[(size_t)i * stride] = t;
                }
            }
        }
    }
    if (POOL) {
        out[((size_t)b * co + o) * na + a] = best;
        arg[((size_t)b * na + a) * co + o] = besti;
    }
}

// grad_e[b, arg[b,a,o], a, o] = grad_out[b, o, a]   (grad_e zero-filled by the entry point)
__global__ void pointnet_pool_bwd_kernel(int64_t total, int n, int na, int co, const float* __restrict__ gout,
                                         const int32_t* __restrict__ arg, float* __restrict__ ge) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;      // (b, a, o)
    if (t >= total) return;
    const int o = (int)(t % co);
    const int64_t ba = t / co;
    const int a = (int)(ba % na);
    const int64_t b = ba / na;
    ge[((b * n + arg[t]) * na + a) * co + o] = gout[(b * co + o) * na + a];
}

}  // namespace vgtkb

using namespace vgtkb;

extern "C" int vgtkb_pointnet_pool_forward(int b, int n, int a, int co, const float* e, const float* v, const float* xc,
                                           float* out, int32_t* arg, void* stream) {
    VGTKB_REQUIRE(b >= 0 && n >= 1 && a >= 1 && co >= 1, "pointnet_pool: bad size");
    VGTKB_REQUIRE(b <= 65535 && a <= 65535, "pointnet_pool: batch/anchors > 65535");
    if (b == 0) return VGTKB_OK;
    pointnet_kernel<true><<<dim3(ceil_div(co, PN_THREADS), a, b), PN_THREADS, 0, (cudaStream_t)stream>>>(
        n, a, co, const_cast<float*>(e), v, xc, out, arg);
    return check_launch("pointnet_pool_forward");
}

extern "C" int vgtkb_pointnet_embed_xyz(int b, int n, int a, int co, float* e, const float* v, const float* xc, void* stream) {
    VGTKB_REQUIRE(b >= 0 && n >= 1 && a >= 1 && co >= 1, "pointnet_embed_xyz: bad size");
    VGTKB_REQUIRE(b <= 65535 && a <= 65535, "pointnet_embed_xyz: batch/anchors > 65535");
    if (b == 0) return VGTKB_OK;
    pointnet_kernel<false><<<dim3(ceil_div(co, PN_THREADS), a, b), PN_THREADS, 0, (cudaStream_t)stream>>>(n, a, co, e, v, xc,
                                                                                                        nullptr, nullptr);
    return check_launch("pointnet_embed_xyz");
}

extern "C" int vgtkb_pointnet_pool_backward(int b, int n, int a, int co, const float* grad_out, const int32_t* arg,
                                            float* grad_e, void* stream) {
    VGTKB_REQUIRE(b >= 0 && n >= 1 && a >= 1 && co >= 1, "pointnet_pool_backward: bad size");
    if (b == 0) return VGTKB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    VGTKB_CUDA(cudaMemsetAsync(grad_e, 0, sizeof(float) * (size_t)b * n * a * co, st));
    const int64_t total = (int64_t)b * a * co;
    pointnet_pool_bwd_kernel<<<(unsigned)ceil_div64(total, 256), 256, 0, st>>>(total, n, a, co, grad_out, arg, grad_e);
    return check_launch("pointnet_pool_backward");
}
