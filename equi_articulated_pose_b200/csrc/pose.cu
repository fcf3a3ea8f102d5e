// pose.cu -- pose-aware inter grouping with arbitrary per-point rotations (InterSO3PoseConv: both branches of
// inter_so3poseconv_grouping_strided, vgtk/vgtk/so3conv/functional.py:896-1060 strided, :1061-1261 no stride; a strided
// layer has p = ceil(n / stride) centres taken from the cloud by sample_idx, whose poses are pose[sample_idx]).
//
// The reference rotates every neighbour offset by R_rel = R_p R_j^T and picks, per (point, neighbour, anchor), the
// anchor pi(a) = argmax_a' tr((R_rel^T R_a) R_a'^T) through a [B,N,nn,A,A,3,3] temporary (68 GB at config-2 size).
// Here a first kernel produces the rotated offsets [B,N,nn,3] and the permutation table [B,N,nn,A] (one byte per
// entry; one warp per (point, neighbour), the 60x60 trace table lives in registers), and the grouping kernels take
// both as inputs.  All shipped configurations keep the identity pose and never reach this file (the module then
// runs the plain, tuned kernels of grouping.cu); these kernels favour simplicity.
#include "common.cuh"

namespace vgtkb {

constexpr int PG_WARPS = 8;
constexpr int PG_MAXNN = 128;
constexpr int PG_KP = 24;
constexpr int PG_MAXA = 64;

// one warp per (b, p, n)
__global__ void __launch_bounds__(256)
pose_neighbourhood_kernel(int64_t total, int n, int p, int nn, int a, const float* __restrict__ xyz, const float* __restrict__ pose,
                          const float* __restrict__ sxyz, const int32_t* __restrict__ sidx,
                          const int32_t* __restrict__ idx, const float* __restrict__ anchors, float* __restrict__ rel_xyz,
                          uint8_t* __restrict__ perm) {
    extern __shared__ float s_anch[];   // [a][9]
    for (int i = threadIdx.x; i < a * 9; i += blockDim.x) s_anch[i] = anchors[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t t = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (t >= total) return;
    const int ni = (int)(t % nn);
    const int pi = (int)((t / nn) % p);
    const int64_t b = t / ((int64_t)nn * p);
    const int j = idx[t];
    const int ci = sidx != nullptr ? sidx[b * p + pi] : pi;            // the centre as a point of the cloud
    const float* Rp = pose + (b * n + ci) * 16;
    const float* Rj = pose + (b * n + j) * 16;
    float rel[9];   // R_p R_j^T
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c)
            rel[r * 3 + c] = Rp[r * 4 + 0] * Rj[c * 4 + 0] + Rp[r * 4 + 1] * Rj[c * 4 + 1] + Rp[r * 4 + 2] * Rj[c * 4 + 2];
    if (lane == 0) {
        const float* X = xyz + b * 3 * n;
        const float* S = sxyz + b * 3 * p;
        const float gx = X[j] - S[pi], gy = X[n + j] - S[p + pi], gz = X[2 * n + j] - S[2 * p + pi];
        rel_xyz[t * 3 + 0] = rel[0] * gx + rel[1] * gy + rel[2] * gz;
        rel_xyz[t * 3 + 1] = rel[3] * gx + rel[4] * gy + rel[5] * gz;
        rel_xyz[t * 3 + 2] = rel[6] * gx + rel[7] * gy + rel[8] * gz;
    }
    if (perm == nullptr) return;
    for (int ai = lane; ai < a; ai += 32) {
        const float* Ra = s_anch + ai * 9;
        float m[9];   // R_rel^T R_a
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) m[r * 3 + c] = rel[0 * 3 + r] * Ra[0 * 3 + c] + rel[1 * 3 + r] * Ra[1 * 3 + c] + rel[2 * 3 + r] * Ra[2 * 3 + c];
        float best = -1e30f;
        int besti = 0;
        for (int a2 = 0; a2 < a; ++a2) {
            const float* R2 = s_anch + a2 * 9;
            float tr = 0.f;
#pragma unroll
            for (int e = 0; e < 9; ++e) tr = fmaf(m[e], R2[e], tr);
            if (tr > best) {   // first maximum, like torch.argmax
                best = tr;
                besti = a2;
            }
        }
        perm[t * a + ai] = (uint8_t)besti;
    }
}

// CTA per point, warp per anchor, lane per channel (stride 32); weights in shared memory per warp
template <bool FWD>
__global__ void __launch_bounds__(PG_WARPS * 32)
pose_group_kernel(int n, int p, int nn, int a, int k, int ci, const int32_t* __restrict__ idx, const float* __restrict__ rel_xyz,
                  const uint8_t* __restrict__ perm, const float* __restrict__ rk, float inv_sigma,
                  const float* __restrict__ in, float* __restrict__ out) {
    extern __shared__ __align__(16) float smem[];
    float* s_w = smem;                                          // [PG_WARPS][nn][PG_KP]
    float* s_g = s_w + PG_WARPS * nn * PG_KP;                   // [nn][3]
    int* s_j = reinterpret_cast<int*>(s_g + nn * 3);            // [nn]
    uint8_t* s_p = reinterpret_cast<uint8_t*>(s_j + nn);        // [nn][a]
    const int pi = blockIdx.x, b = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t row = (int64_t)b * p + pi;
    for (int i = threadIdx.x; i < nn; i += blockDim.x) {
        s_j[i] = idx[row * nn + i];
        s_g[i * 3 + 0] = rel_xyz[(row * nn + i) * 3 + 0];
        s_g[i * 3 + 1] = rel_xyz[(row * nn + i) * 3 + 1];
        s_g[i * 3 + 2] = rel_xyz[(row * nn + i) * 3 + 2];
    }
    for (int i = threadIdx.x; i < nn * a; i += blockDim.x) s_p[i] = perm != nullptr ? perm[row * nn * a + i] : (uint8_t)(i % a);
    __syncthreads();
    float* w_a = s_w + warp * nn * PG_KP;
    for (int ai = warp; ai < a; ai += PG_WARPS) {
        __syncwarp();
        for (int e = lane; e < nn * PG_KP; e += 32) {
            const int ni = e / PG_KP, ki = e % PG_KP;
            float w = 0.f;
            if (ki < k) {
                const float* kp = rk + (ai * k + ki) * 3;
                const float dx = s_g[ni * 3] - kp[0], dy = s_g[ni * 3 + 1] - kp[1], dz = s_g[ni * 3 + 2] - kp[2];
                w = fmaxf(0.f, 1.f - (dx * dx + dy * dy + dz * dz) * inv_sigma);
            }
            w_a[e] = w;
        }
        __syncwarp();
        const int64_t grow = (row * a + ai) * (int64_t)k * ci;      // G / dG row of this (point, anchor)
        for (int c = lane; c < ci; c += 32) {
            if (FWD) {
                float acc[PG_KP];
#pragma unroll
                for (int ki = 0; ki < PG_KP; ++ki) acc[ki] = 0.f;
                for (int ni = 0; ni < nn; ++ni) {
                    const float x = in[(((int64_t)b * n + s_j[ni]) * a + s_p[ni * a + ai]) * ci + c];
#pragma unroll
                    for (int ki = 0; ki < PG_KP; ++ki) acc[ki] = fmaf(w_a[ni * PG_KP + ki], x, acc[ki]);
                }
#pragma unroll
                for (int ki = 0; ki < PG_KP; ++ki)
                    if (ki < k) out[grow + (int64_t)ki * ci + c] = acc[ki];
            } else {
                float dg[PG_KP];
#pragma unroll
                for (int ki = 0; ki < PG_KP; ++ki) dg[ki] = ki < k ? in[grow + (int64_t)ki * ci + c] : 0.f;
                for (int ni = 0; ni < nn; ++ni) {
                    float v = 0.f;
#pragma unroll
                    for (int ki = 0; ki < PG_KP; ++ki) v = fmaf(w_a[ni * PG_KP + ki], dg[ki], v);
                    atomicAdd(out + (((int64_t)b * n + s_j[ni]) * a + s_p[ni * a + ai]) * ci + c, v);
                }
            }
        }
    }
}

}  // namespace vgtkb

using namespace vgtkb;

extern "C" int vgtkb_pose_neighbourhood_strided(int b, int n, int p, int nn, int a, const float* xyz, const float* pose,
                                                const float* sample_xyz, const int32_t* sample_idx, const int32_t* idx,
                                                const float* anchors, float* rel_xyz, uint8_t* perm, void* stream) {
    VGTKB_REQUIRE(b >= 0 && n > 0 && p > 0 && nn > 0 && a > 0 && a <= 255, "pose_neighbourhood: bad size");
    VGTKB_REQUIRE(sample_idx != nullptr || p == n, "pose_neighbourhood: p != n needs sample_idx");
    const int64_t total = (int64_t)b * p * nn;
    if (total == 0) return VGTKB_OK;
    const unsigned grid = (unsigned)ceil_div64(total, 8);
    pose_neighbourhood_kernel<<<grid, 256, sizeof(float) * a * 9, (cudaStream_t)stream>>>(total, n, p, nn, a, xyz, pose, sample_xyz,
                                                                                         sample_idx, idx, anchors, rel_xyz, perm);
    return check_launch("pose_neighbourhood");
}

extern "C" int vgtkb_pose_neighbourhood(int b, int n, int nn, int a, const float* xyz, const float* pose, const int32_t* idx,
                                        const float* anchors, float* rel_xyz, uint8_t* perm, void* stream) {
    return vgtkb_pose_neighbourhood_strided(b, n, n, nn, a, xyz, pose, xyz, nullptr, idx, anchors, rel_xyz, perm, stream);
}

static int launch_pose_group(bool fwd, int b, int n, int p, int nn, int a, int k, int ci, const int32_t* idx, const float* rel_xyz,
                             const uint8_t* perm, const float* rk, float sigma, const float* in, float* out, cudaStream_t st) {
    VGTKB_REQUIRE(b >= 0 && n > 0 && p > 0 && nn > 0 && nn <= PG_MAXNN && a > 0 && a <= 255 && k > 0 && k <= PG_KP && ci > 0 &&
                      b <= 65535,
                  "inter_pose_group: bad size (nn <= 128, k <= 24, a <= 255)");
    if (b == 0) return VGTKB_OK;
    const size_t smem = ((size_t)PG_WARPS * nn * PG_KP + nn * 3 + nn) * 4 + (((size_t)nn * a + 15) & ~(size_t)15);
    auto kf = pose_group_kernel<true>;
    auto kb = pose_group_kernel<false>;
    VGTKB_CUDA(cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    VGTKB_CUDA(cudaFuncSetAttribute(kb, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    if (fwd) kf<<<dim3(p, b), PG_WARPS * 32, smem, st>>>(n, p, nn, a, k, ci, idx, rel_xyz, perm, rk, 1.0f / sigma, in, out);
    else kb<<<dim3(p, b), PG_WARPS * 32, smem, st>>>(n, p, nn, a, k, ci, idx, rel_xyz, perm, rk, 1.0f / sigma, in, out);
    return check_launch(fwd ? "inter_pose_group_forward" : "inter_pose_group_backward");
}

extern "C" int vgtkb_inter_pose_group_forward_strided(int b, int n, int p, int nn, int a, int k, int ci, const int32_t* idx,
                                                      const float* rel_xyz, const uint8_t* perm, const float* rot_kernels,
                                                      float sigma, const float* feats, float* grouped, void* stream) {
    return launch_pose_group(true, b, n, p, nn, a, k, ci, idx, rel_xyz, perm, rot_kernels, sigma, feats, grouped, (cudaStream_t)stream);
}

extern "C" int vgtkb_inter_pose_group_backward_strided(int b, int n, int p, int nn, int a, int k, int ci, const int32_t* idx,
                                                       const float* rel_xyz, const uint8_t* perm, const float* rot_kernels,
                                                       float sigma, const float* grad_grouped, float* grad_feats, void* stream) {
    return launch_pose_group(false, b, n, p, nn, a, k, ci, idx, rel_xyz, perm, rot_kernels, sigma, grad_grouped, grad_feats,
                             (cudaStream_t)stream);
}

extern "C" int vgtkb_inter_pose_group_forward(int b, int n, int nn, int a, int k, int ci, const int32_t* idx, const float* rel_xyz,
                                              const uint8_t* perm, const float* rot_kernels, float sigma, const float* feats,
                                              float* grouped, void* stream) {
    return launch_pose_group(true, b, n, n, nn, a, k, ci, idx, rel_xyz, perm, rot_kernels, sigma, feats, grouped, (cudaStream_t)stream);
}

extern "C" int vgtkb_inter_pose_group_backward(int b, int n, int nn, int a, int k, int ci, const int32_t* idx, const float* rel_xyz,
                                               const uint8_t* perm, const float* rot_kernels, float sigma, const float* grad_grouped,
                                               float* grad_feats, void* stream) {
    return launch_pose_group(false, b, n, n, nn, a, k, ci, idx, rel_xyz, perm, rot_kernels, sigma, grad_grouped, grad_feats,
                             (cudaStream_t)stream);
}
