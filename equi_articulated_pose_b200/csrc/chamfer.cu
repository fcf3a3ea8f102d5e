// chamfer.cu -- bidirectional nearest-neighbour squared distance + gradient for sm_100a.
//
// Replaces extensions/chamfer_dist/chamfer.cu of the reference (forward :15-145, backward
// :173-229).  Same values (the squared distance is evaluated with the reference's contraction
// FMUL,FFMA,FFMA and strict '<' keeps the lowest index on ties) with a grid sized from the
// problem (the reference launches a fixed 32x16 grid), cloud-2 tiles staged by the TMA
// bulk-copy engine and two queries per thread sharing every shared-memory read.
#include "common.cuh"

namespace vgtkb {

constexpr int CH_THREADS = 256;
constexpr int CH_QPT = 2;      // queries per thread
constexpr int CH_TILE = 2048;  // cloud-2 points per shared-memory tile (24 KB)

__global__ void __launch_bounds__(CH_THREADS)
chamfer_nn_kernel(int n, const float* __restrict__ xyz1, int m, const float* __restrict__ xyz2,
                  float* __restrict__ dist, int32_t* __restrict__ index) {
    __shared__ __align__(16) float buf[CH_TILE * 3];
    __shared__ __align__(8) uint64_t bar;
    const int b = blockIdx.y;
    const float* p1 = xyz1 + (size_t)b * n * 3;
    const float* p2 = xyz2 + (size_t)b * m * 3;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_barrier_init();
    }
    float x1[CH_QPT], y1[CH_QPT], z1[CH_QPT], best[CH_QPT];
    int besti[CH_QPT], q[CH_QPT];
#pragma unroll
    for (int i = 0; i < CH_QPT; ++i) {
        q[i] = (blockIdx.x * CH_QPT + i) * CH_THREADS + threadIdx.x;
        const int qq = min(q[i], n - 1);
        x1[i] = p1[qq * 3 + 0];
        y1[i] = p1[qq * 3 + 1];
        z1[i] = p1[qq * 3 + 2];
        best[i] = 0.f;
        besti[i] = 0;
    }
    const bool aligned = (reinterpret_cast<uintptr_t>(p2) & 15) == 0;
    uint32_t phase = 0;
    __syncthreads();
    for (int k2 = 0; k2 < m; k2 += CH_TILE) {
        const int len = min(CH_TILE, m - k2);
        const uint32_t bytes = (uint32_t)len * 12u;
        // tile start k2*12 bytes is a multiple of 16 because CH_TILE*12 is
        if (aligned && (bytes & 15u) == 0) {
            if (threadIdx.x == 0) {
                fence_proxy_async();
                mbar_arrive_expect_tx(&bar, bytes);
                bulk_g2s(buf, p2 + (size_t)k2 * 3, bytes, &bar);
            }
            mbar_wait(&bar, phase);
            phase ^= 1;
        } else {
            for (int j = threadIdx.x; j < len * 3; j += CH_THREADS) buf[j] = p2[(size_t)k2 * 3 + j];
            __syncthreads();
        }
#pragma unroll 4
        for (int k = 0; k < len; ++k) {
            const float bx = buf[k * 3 + 0], by = buf[k * 3 + 1], bz = buf[k * 3 + 2];
#pragma unroll
            for (int i = 0; i < CH_QPT; ++i) {
                // reference SASS (all 43 sites): FMUL on y, FFMA x, FFMA z -> fma(z,z,fma(x,x,y*y))
                const float d = sq3(by - y1[i], bx - x1[i], bz - z1[i]);
                const bool take = (k2 + k == 0) || d < best[i];
                best[i] = take ? d : best[i];
                besti[i] = take ? k2 + k : besti[i];
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < CH_QPT; ++i)
        if (q[i] < n) {
            dist[(size_t)b * n + q[i]] = best[i];
            index[(size_t)b * n + q[i]] = besti[i];
        }
}

// own part: g1[b,j,:] = 2 g (p1_j - p2_idx1[j])           (plain stores, initialises the buffer)
__global__ void chamfer_grad_own_kernel(int64_t total, int n, int m, const float* __restrict__ xyz1,
                                        const float* __restrict__ xyz2, const float* __restrict__ gd,
                                        const int32_t* __restrict__ idx, float* __restrict__ g1) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int64_t b = t / n;
    const int j2 = idx[t];
    const float* a = xyz1 + t * 3;
    const float* c = xyz2 + (b * m + j2) * 3;
    const float g = gd[t] * 2;
#pragma unroll
    for (int d = 0; d < 3; ++d) g1[t * 3 + d] = g * (a[d] - c[d]);
}
// cross part: g2[b,idx1[j],:] -= 2 g (p1_j - p2_idx1[j])  (atomics; after BOTH own parts)
__global__ void chamfer_grad_cross_kernel(int64_t total, int n, int m, const float* __restrict__ xyz1,
                                          const float* __restrict__ xyz2, const float* __restrict__ gd,
                                          const int32_t* __restrict__ idx, float* __restrict__ g2) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int64_t b = t / n;
    const int j2 = idx[t];
    const float* a = xyz1 + t * 3;
    const float* c = xyz2 + (b * m + j2) * 3;
    const float g = gd[t] * 2;
#pragma unroll
    for (int d = 0; d < 3; ++d) atomicAdd(g2 + (b * m + j2) * 3 + d, -(g * (a[d] - c[d])));
}

// ---------------------------------------------------------------------------------- anchor chamfer
// Model 38's reconstruction loss (SPConvNets/models/unsup_seg_so3_pose_conv_pn_38_multi_stage.py:429-436): the
// reconstruction is rigidly transformed by every one of the A anchor poses, Y[b,a,i] = R[b,a] p[b,i] + T[b,a], and each
// copy is chamfer-matched against the SAME input cloud, which the reference replicates A times ([B*A, N, 3]) next to
// the materialised [B, A, M, 3] transformed tensor.  Here the transform is applied on the fly -- to the queries
// (direction recon -> ori) or to the shared-memory tile right after it lands (direction ori -> recon) -- so neither
// tensor exists.  Distances as in chamfer_nn_kernel (same contraction, lowest index on ties).
__device__ __forceinline__ void rigid(const float* __restrict__ R, const float* __restrict__ T, float px, float py, float pz,
                                      float& x, float& y, float& z) {
    x = __fmaf_rn(R[2], pz, __fmaf_rn(R[1], py, __fmul_rn(R[0], px))) + T[0];
    y = __fmaf_rn(R[5], pz, __fmaf_rn(R[4], py, __fmul_rn(R[3], px))) + T[1];
    z = __fmaf_rn(R[8], pz, __fmaf_rn(R[7], py, __fmul_rn(R[6], px))) + T[2];
}

// XQ = true : queries = transformed reconstruction (nq = M), database = input cloud (nd = N)
// XQ = false: queries = input cloud (nq = N), database = transformed reconstruction (nd = M)
template <bool XQ>
__global__ void __launch_bounds__(CH_THREADS)
anchor_chamfer_nn_kernel(int na, int nq, const float* __restrict__ qpts, int nd, const float* __restrict__ dpts,
                         const float* __restrict__ rot, const float* __restrict__ trans, float* __restrict__ dist,
                         int32_t* __restrict__ index) {
    __shared__ __align__(16) float buf[CH_TILE * 3];
    __shared__ float sR[9], sT[3];
    const int a = blockIdx.y, b = blockIdx.z;
    const float* p1 = qpts + (size_t)b * nq * 3;
    const float* p2 = dpts + (size_t)b * nd * 3;
    if (threadIdx.x < 9) sR[threadIdx.x] = rot[((size_t)b * na + a) * 9 + threadIdx.x];
    if (threadIdx.x < 3) sT[threadIdx.x] = trans[((size_t)b * na + a) * 3 + threadIdx.x];
    __syncthreads();
    float x1[CH_QPT], y1[CH_QPT], z1[CH_QPT], best[CH_QPT];
    int besti[CH_QPT], q[CH_QPT];
#pragma unroll
    for (int i = 0; i < CH_QPT; ++i) {
        q[i] = (blockIdx.x * CH_QPT + i) * CH_THREADS + threadIdx.x;
        const int qq = min(q[i], nq - 1);
        const float px = p1[qq * 3 + 0], py = p1[qq * 3 + 1], pz = p1[qq * 3 + 2];
        if (XQ) rigid(sR, sT, px, py, pz, x1[i], y1[i], z1[i]);
        else { x1[i] = px; y1[i] = py; z1[i] = pz; }
        best[i] = 0.f;
        besti[i] = 0;
    }
    for (int k2 = 0; k2 < nd; k2 += CH_TILE) {
        const int len = min(CH_TILE, nd - k2);
        for (int k = threadIdx.x; k < len; k += CH_THREADS) {
            const float px = p2[(size_t)(k2 + k) * 3 + 0], py = p2[(size_t)(k2 + k) * 3 + 1], pz = p2[(size_t)(k2 + k) * 3 + 2];
            if (XQ) { buf[k * 3 + 0] = px; buf[k * 3 + 1] = py; buf[k * 3 + 2] = pz; }
            else rigid(sR, sT, px, py, pz, buf[k * 3 + 0], buf[k * 3 + 1], buf[k * 3 + 2]);
        }
        __syncthreads();
#pragma unroll 4
        for (int k = 0; k < len; ++k) {
            const float bx = buf[k * 3 + 0], by = buf[k * 3 + 1], bz = buf[k * 3 + 2];
#pragma unroll
            for (int i = 0; i < CH_QPT; ++i) {
                const float d = sq3(by - y1[i], bx - x1[i], bz - z1[i]);
                const bool take = (k2 + k == 0) || d < best[i];
                best[i] = take ? d : best[i];
                besti[i] = take ? k2 + k : besti[i];
            }
        }
        __syncthreads();
    }
    const size_t o = ((size_t)b * na + a) * nq;
#pragma unroll
    for (int i = 0; i < CH_QPT; ++i)
        if (q[i] < nq) {
            dist[o + q[i]] = best[i];
            index[o + q[i]] = besti[i];
        }
}

// gradients w.r.t. the transformed reconstruction Y [B,A,M,3] and the input cloud [B,N,3] (both zero-initialised by
// the caller): thread t = (b, a, i) for direction 1 (recon -> ori), (b, a, j) for direction 2.
template <bool DIR1>
__global__ void anchor_chamfer_grad_kernel(int64_t total, int na, int m, int n, const float* __restrict__ canon,
                                           const float* __restrict__ rot, const float* __restrict__ trans,
                                           const float* __restrict__ ori, const int32_t* __restrict__ idx,
                                           const float* __restrict__ gd, float* __restrict__ gy, float* __restrict__ gori) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int nq = DIR1 ? m : n;
    const int64_t ba = t / nq;
    const int own = (int)(t - ba * nq), other = idx[t];
    const int64_t b = ba / na;
    const int i = DIR1 ? own : other, j = DIR1 ? other : own;     // i: reconstruction point, j: input point
    const float* p = canon + (b * m + i) * 3;
    float y[3];
    rigid(rot + ba * 9, trans + ba * 3, p[0], p[1], p[2], y[0], y[1], y[2]);
    const float* o = ori + (b * n + j) * 3;
    const float g = gd[t] * 2;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const float v = g * (y[d] - o[d]);
        if (DIR1) gy[(ba * m + i) * 3 + d] += v;                  // own slot: exclusive in this launch
        else atomicAdd(gy + (ba * m + i) * 3 + d, v);
        if (gori != nullptr) atomicAdd(gori + (b * n + j) * 3 + d, -v);
    }
}

}  // namespace vgtkb

using namespace vgtkb;

extern "C" int vgtkb_chamfer_forward(int b, int n, const float* xyz1, int m, const float* xyz2, float* dist1,
                                     float* dist2, int32_t* idx1, int32_t* idx2, void* stream) {
    VGTKB_REQUIRE(b >= 0 && n >= 0 && m >= 0, "chamfer: negative size");
    VGTKB_REQUIRE(b <= 65535, "chamfer: batch %d > 65535", b);
    cudaStream_t st = (cudaStream_t)stream;
    if (b == 0) return VGTKB_OK;
    if (n == 0 || m == 0) {  // reference returns the zero-initialised outputs
        if (n) {
            VGTKB_CUDA(cudaMemsetAsync(dist1, 0, sizeof(float) * (size_t)b * n, st));
            VGTKB_CUDA(cudaMemsetAsync(idx1, 0, sizeof(int32_t) * (size_t)b * n, st));
        }
        if (m) {
            VGTKB_CUDA(cudaMemsetAsync(dist2, 0, sizeof(float) * (size_t)b * m, st));
            VGTKB_CUDA(cudaMemsetAsync(idx2, 0, sizeof(int32_t) * (size_t)b * m, st));
        }
        return VGTKB_OK;
    }
    const int per = CH_THREADS * CH_QPT;
    chamfer_nn_kernel<<<dim3(ceil_div(n, per), b), CH_THREADS, 0, st>>>(n, xyz1, m, xyz2, dist1, idx1);
    chamfer_nn_kernel<<<dim3(ceil_div(m, per), b), CH_THREADS, 0, st>>>(m, xyz2, n, xyz1, dist2, idx2);
    return check_launch("chamfer_forward");
}

extern "C" int vgtkb_chamfer_backward(int b, int n, const float* xyz1, int m, const float* xyz2,
                                      const int32_t* idx1, const int32_t* idx2, const float* grad_dist1,
                                      const float* grad_dist2, float* grad_xyz1, float* grad_xyz2, void* stream) {
    VGTKB_REQUIRE(b >= 0 && n >= 0 && m >= 0, "chamfer: negative size");
    cudaStream_t st = (cudaStream_t)stream;
    if (b == 0) return VGTKB_OK;
    if (n == 0 || m == 0) {
        if (n) VGTKB_CUDA(cudaMemsetAsync(grad_xyz1, 0, sizeof(float) * (size_t)b * n * 3, st));
        if (m) VGTKB_CUDA(cudaMemsetAsync(grad_xyz2, 0, sizeof(float) * (size_t)b * m * 3, st));
        return VGTKB_OK;
    }
    const int64_t t1 = (int64_t)b * n, t2 = (int64_t)b * m;
    chamfer_grad_own_kernel<<<(unsigned)ceil_div64(t1, 256), 256, 0, st>>>(t1, n, m, xyz1, xyz2, grad_dist1, idx1, grad_xyz1);
    chamfer_grad_own_kernel<<<(unsigned)ceil_div64(t2, 256), 256, 0, st>>>(t2, m, n, xyz2, xyz1, grad_dist2, idx2, grad_xyz2);
    chamfer_grad_cross_kernel<<<(unsigned)ceil_div64(t1, 256), 256, 0, st>>>(t1, n, m, xyz1, xyz2, grad_dist1, idx1, grad_xyz2);
    chamfer_grad_cross_kernel<<<(unsigned)ceil_div64(t2, 256), 256, 0, st>>>(t2, m, n, xyz2, xyz1, grad_dist2, idx2, grad_xyz1);
    return check_launch("chamfer_backward");
}

extern "C" int vgtkb_anchor_chamfer_forward(int b, int a, int m, const float* canon, const float* rot, const float* trans,
                                            int n, const float* ori, float* dist1, float* dist2, int32_t* idx1,
                                            int32_t* idx2, void* stream) {
    VGTKB_REQUIRE(b >= 0 && a >= 1 && m >= 1 && n >= 1, "anchor_chamfer: bad size");
    VGTKB_REQUIRE(b <= 65535 && a <= 65535, "anchor_chamfer: batch/anchors > 65535");
    cudaStream_t st = (cudaStream_t)stream;
    if (b == 0) return VGTKB_OK;
    const int per = CH_THREADS * CH_QPT;
    anchor_chamfer_nn_kernel<true><<<dim3(ceil_div(m, per), a, b), CH_THREADS, 0, st>>>(a, m, canon, n, ori, rot, trans, dist1, idx1);
    anchor_chamfer_nn_kernel<false><<<dim3(ceil_div(n, per), a, b), CH_THREADS, 0, st>>>(a, n, ori, m, canon, rot, trans, dist2, idx2);
    return check_launch("anchor_chamfer_forward");
}

extern "C" int vgtkb_anchor_chamfer_backward(int b, int a, int m, const float* canon, const float* rot, const float* trans,
                                             int n, const float* ori, const int32_t* idx1, const int32_t* idx2,
                                             const float* grad_dist1, const float* grad_dist2, float* grad_y,
                                             float* grad_ori, void* stream) {
    VGTKB_REQUIRE(b >= 0 && a >= 1 && m >= 1 && n >= 1, "anchor_chamfer: bad size");
    cudaStream_t st = (cudaStream_t)stream;
    if (b == 0) return VGTKB_OK;
    VGTKB_CUDA(cudaMemsetAsync(grad_y, 0, sizeof(float) * (size_t)b * a * m * 3, st));
    if (grad_ori != nullptr) VGTKB_CUDA(cudaMemsetAsync(grad_ori, 0, sizeof(float) * (size_t)b * n * 3, st));
    const int64_t t1 = (int64_t)b * a * m, t2 = (int64_t)b * a * n;
    anchor_chamfer_grad_kernel<true><<<(unsigned)ceil_div64(t1, 256), 256, 0, st>>>(t1, a, m, n, canon, rot, trans, ori, idx1,
                                                                                  grad_dist1, grad_y, grad_ori);
    anchor_chamfer_grad_kernel<false><<<(unsigned)ceil_div64(t2, 256), 256, 0, st>>>(t2, a, m, n, canon, rot, trans, ori, idx2,
                                                                                   grad_dist2, grad_y, grad_ori);
    return check_launch("anchor_chamfer_backward");
}
