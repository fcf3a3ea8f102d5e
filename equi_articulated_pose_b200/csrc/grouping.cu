// grouping.cu -- kernel-point correlation + inter-/intra-anchor grouping for sm_100a.
//
// The reference evaluates these in eager PyTorch (vgtk/vgtk/so3conv/functional.py:2508-2549,
// 2553-2567 and vgtk/vgtk/spconv/functional.py:375-406) and materialises, per layer, the
// weights w[B,P,A,K,nn] (755 MB at config 2), three permuted copies of the gathered neighbour
// features and the 12x index_select copy.  Here the weights live only in shared memory:
// one CTA owns an output point, a warp owns an anchor, computes its [nn x K] weight tile and
// streams the neighbour rows (channels-last, one coalesced 256/512-byte row segment per load)
// through K*CPL register accumulators.
#include "common.cuh"

namespace vgtkb {

constexpr int IG_WARPS = 8;
constexpr int IG_MAXNN = 128;
constexpr int IG_KP = 24;  // padded kernel-point count handled by the fast path (K <= 24)

// materialised weights (API parity / tests only)
__global__ void inter_weights_kernel(int n, int p, int nn, int a, int k, const float* __restrict__ xyz,
                                     const float* __restrict__ sxyz, const int32_t* __restrict__ idx,
                                     const float* __restrict__ rk, float inv_sigma, float* __restrict__ w) {
    const int b = blockIdx.y;
    const int64_t per = (int64_t)a * k * nn;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)p * per) return;
    const int pi = (int)(t / per);
    const int r = (int)(t % per);
    const int ai = r / (k * nn), ki = (r / nn) % k, ni = r % nn;
    const int j = idx[((size_t)b * p + pi) * nn + ni];
    const float* X = xyz + (size_t)b * 3 * n;
    const float* S = sxyz + (size_t)b * 3 * p;
    const float gx = X[j] - S[pi], gy = X[n + j] - S[p + pi], gz = X[2 * n + j] - S[2 * p + pi];
    const float* kp = rk + ((size_t)ai * k + ki) * 3;
    const float dx = gx - kp[0], dy = gy - kp[1], dz = gz - kp[2];
    const float d2 = dx * dx + dy * dy + dz * dz;
    w[(size_t)b * p * per + t] = fmaxf(0.f, 1.f - d2 * inv_sigma);
}

// shared prologue: neighbour indices and centre-relative offsets of point (b,pi)
__device__ __forceinline__ void load_neighbourhood(int b, int pi, int n, int p, int nn, const float* xyz,
                                                   const float* sxyz, const int32_t* idx, int* s_j, float* s_g) {
    const float* X = xyz + (size_t)b * 3 * n;
    const float* S = sxyz + (size_t)b * 3 * p;
    for (int i = threadIdx.x; i < nn; i += blockDim.x) {
        const int j = idx[((size_t)b * p + pi) * nn + i];
        s_j[i] = j;
        s_g[i * 3 + 0] = X[j] - S[pi];
        s_g[i * 3 + 1] = X[n + j] - S[p + pi];
        s_g[i * 3 + 2] = X[2 * n + j] - S[2 * p + pi];
    }
}

// w_a[nn][IG_KP] for anchor ai, computed by one warp: lane <-> kernel point (its rotated position stays in
// three registers), loop over the neighbours; lanes >= k write the zero padding
__device__ __forceinline__ void warp_weights(int lane, int ai, int nn, int k, const float* rk, float inv_sigma,
                                             const float* s_g, float* w_a) {
    if (lane < IG_KP) {
        const bool live = lane < k;
        float kx = 0.f, ky = 0.f, kz = 0.f;
        if (live) {
            const float* kp = rk + (ai * k + lane) * 3;
            kx = __ldg(kp);
            ky = __ldg(kp + 1);
            kz = __ldg(kp + 2);
        }
#pragma unroll 4
        for (int ni = 0; ni < nn; ++ni) {
            const float dx = s_g[ni * 3] - kx, dy = s_g[ni * 3 + 1] - ky, dz = s_g[ni * 3 + 2] - kz;
            const float w = fmaxf(0.f, 1.f - (dx * dx + dy * dy + dz * dz) * inv_sigma);
            w_a[ni * IG_KP + lane] = live ? w : 0.f;
        }
    }
}

// element-parallel variant (lane <-> (neighbour, kernel point) pairs): faster inside the register-capped
// forward kernel, slower in the backward one (measured both ways)
__device__ __forceinline__ void warp_weights_flat(int lane, int ai, int nn, int k, const float* rk, float inv_sigma,
                                                  const float* s_g, float* w_a) {
    for (int e = lane; e < nn * IG_KP; e += 32) {
        const int ni = e / IG_KP, ki = e % IG_KP;
        float w = 0.f;
        if (ki < k) {
            const float* kp = rk + (ai * k + ki) * 3;
            const float dx = s_g[ni * 3] - __ldg(kp), dy = s_g[ni * 3 + 1] - __ldg(kp + 1),
                        dz = s_g[ni * 3 + 2] - __ldg(kp + 2);
            w = fmaxf(0.f, 1.f - (dx * dx + dy * dy + dz * dz) * inv_sigma);
        }
        w_a[e] = w;
    }
}

template <int CPL>
struct VecT;
template <>
struct VecT<2> { using type = float2; };
template <>
struct VecT<4> { using type = float4; };

// fast path: ci % (32*CPL) == 0, k <= 24.  grid (p, b), 8 warps, warp <-> anchor.
// The 24 kernel points are processed in two groups of 12 so that the accumulators (12 x CPL registers)
// leave room for 3 CTAs per SM (the first version kept 24 x CPL: 153/205 registers, 1 CTA per SM, 12 %
// occupancy, latency bound -- profiles/r1_ncu_inter_group.txt); the second pass re-reads the neighbour
// rows from L1/L2.
constexpr int IG_KG = 12;
template <int CPL>
__global__ void __launch_bounds__(IG_WARPS * 32, 3)
inter_group_fwd_kernel(int n, int p, int nn, int a, int k, int ci, const float* __restrict__ xyz,
                       const float* __restrict__ sxyz, const int32_t* __restrict__ idx,
                       const float* __restrict__ rk, float inv_sigma, const float* __restrict__ feats,
                       float* __restrict__ grouped) {
    using V = typename VecT<CPL>::type;
    extern __shared__ __align__(16) float smem[];
    float* s_w = smem;                                   // [IG_WARPS][nn][IG_KP]
    float* s_g = s_w + IG_WARPS * nn * IG_KP;            // [nn][3]
    int* s_j = reinterpret_cast<int*>(s_g + nn * 3);     // [nn]
    const int pi = blockIdx.x, b = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    load_neighbourhood(b, pi, n, p, nn, xyz, sxyz, idx, s_j, s_g);
    __syncthreads();
    // neighbour rows as 32-bit element offsets j * A * Ci (the 64-bit index arithmetic per neighbour was a
    // quarter of the issued instructions); s_j is reused in place, every thread only rewrites its own entries
    uint32_t* s_off = reinterpret_cast<uint32_t*>(s_j);
    for (int i = threadIdx.x; i < nn; i += blockDim.x) s_off[i] = (uint32_t)s_j[i] * (uint32_t)(a * ci);
    __syncthreads();
    float* w_a = s_w + warp * nn * IG_KP;
    const int chunks = ci / (32 * CPL);
    const float* fb = feats + (size_t)b * n * a * ci;
    for (int ai = warp; ai < a; ai += IG_WARPS) {
        __syncwarp();
        warp_weights_flat(lane, ai, nn, k, rk, inv_sigma, s_g, w_a);
        __syncwarp();
        float* out = grouped + (((size_t)b * p + pi) * a + ai) * (size_t)k * ci;
        for (int ch = 0; ch < chunks; ++ch) {
            const int c0 = ch * 32 * CPL + lane * CPL;
            const float* fcol = fb + (size_t)ai * ci + c0;
#pragma unroll 1
            for (int kg = 0; kg < IG_KP; kg += IG_KG) {
                float acc[IG_KG][CPL];
#pragma unroll
                for (int i = 0; i < IG_KG; ++i)
#pragma unroll
                    for (int u = 0; u < CPL; ++u) acc[i][u] = 0.f;
#pragma unroll 4
                for (int ni = 0; ni < nn; ++ni) {
                    const V xv = __ldg(reinterpret_cast<const V*>(fcol + s_off[ni]));
                    const float* xs = reinterpret_cast<const float*>(&xv);
                    const float4* wr = reinterpret_cast<const float4*>(w_a + ni * IG_KP + kg);
#pragma unroll
                    for (int k4 = 0; k4 < IG_KG / 4; ++k4) {
                        const float4 w4 = wr[k4];
                        const float ws[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                        for (int i = 0; i < 4; ++i)
#pragma unroll
                            for (int u = 0; u < CPL; ++u) acc[k4 * 4 + i][u] = fmaf(ws[i], xs[u], acc[k4 * 4 + i][u]);
                    }
                }
#pragma unroll
                for (int ki = 0; ki < IG_KG; ++ki)
                    if (kg + ki < k) {
                        V o;
                        float* os = reinterpret_cast<float*>(&o);
#pragma unroll
                        for (int u = 0; u < CPL; ++u) os[u] = acc[ki][u];
                        __stcs(reinterpret_cast<V*>(out + (size_t)(kg + ki) * ci + c0), o);   // streaming: G is read once, later
                    }
            }
        }
    }
}

// backward of the fast path: dX[b, j_n, a, c] += sum_k w[n][k] dG[b,p,a,k,c]
// (bound by the red.global.add traffic: the 16-byte vector atomics of CPL = 4 beat higher occupancy with
// CPL = 2 -- measured 2.9 ms vs 3.4 ms per step)
template <int CPL>
__global__ void __launch_bounds__(IG_WARPS * 32)
inter_group_bwd_kernel(int n, int p, int nn, int a, int k, int ci, const float* __restrict__ xyz,
                       const float* __restrict__ sxyz, const int32_t* __restrict__ idx,
                       const float* __restrict__ rk, float inv_sigma, const float* __restrict__ ggrouped,
                       float* __restrict__ gfeats) {
    using V = typename VecT<CPL>::type;
    extern __shared__ __align__(16) float smem[];
    float* s_w = smem;
    float* s_g = s_w + IG_WARPS * nn * IG_KP;
    int* s_j = reinterpret_cast<int*>(s_g + nn * 3);
    const int pi = blockIdx.x, b = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    load_neighbourhood(b, pi, n, p, nn, xyz, sxyz, idx, s_j, s_g);
    __syncthreads();
    uint32_t* s_off = reinterpret_cast<uint32_t*>(s_j);   // 32-bit element offsets j * A * Ci (see the forward kernel)
    for (int i = threadIdx.x; i < nn; i += blockDim.x) s_off[i] = (uint32_t)s_j[i] * (uint32_t)(a * ci);
    __syncthreads();
    float* gb = gfeats + (size_t)b * n * a * ci;
    float* w_a = s_w + warp * nn * IG_KP;
    const int chunks = ci / (32 * CPL);
    for (int ai = warp; ai < a; ai += IG_WARPS) {
        __syncwarp();
        warp_weights(lane, ai, nn, k, rk, inv_sigma, s_g, w_a);
        __syncwarp();
        const float* gin = ggrouped + (((size_t)b * p + pi) * a + ai) * (size_t)k * ci;
        for (int ch = 0; ch < chunks; ++ch) {
            const int c0 = ch * 32 * CPL + lane * CPL;
            float dg[IG_KP][CPL];
#pragma unroll
            for (int ki = 0; ki < IG_KP; ++ki) {
                V v;
                float* vs = reinterpret_cast<float*>(&v);
                if (ki < k) v = __ldcs(reinterpret_cast<const V*>(gin + (size_t)ki * ci + c0));   // dG is read once
                else
#pragma unroll
                    for (int u = 0; u < CPL; ++u) vs[u] = 0.f;
#pragma unroll
                for (int u = 0; u < CPL; ++u) dg[ki][u] = vs[u];
            }
#pragma unroll 2
            for (int ni = 0; ni < nn; ++ni) {
                const float4* wr = reinterpret_cast<const float4*>(w_a + ni * IG_KP);
                float val[CPL];
#pragma unroll
                for (int u = 0; u < CPL; ++u) val[u] = 0.f;
#pragma unroll
                for (int k4 = 0; k4 < IG_KP / 4; ++k4) {
                    const float4 w4 = wr[k4];
                    const float ws[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int u = 0; u < CPL; ++u) val[u] = fmaf(ws[i], dg[k4 * 4 + i][u], val[u]);
                }
                float* dst = gb + (size_t)ai * ci + c0 + s_off[ni];
                V o;
                float* os = reinterpret_cast<float*>(&o);
#pragma unroll
                for (int u = 0; u < CPL; ++u) os[u] = val[u];
                atomicAdd(reinterpret_cast<V*>(dst), o);  // vector red.global.add (sm_90+)
            }
        }
    }
}

// generic path (any ci, any k): one thread per (a,k,c) of a point; weights recomputed per item.
__global__ void inter_group_fwd_generic_kernel(int n, int p, int nn, int a, int k, int ci,
                                               const float* __restrict__ xyz, const float* __restrict__ sxyz,
                                               const int32_t* __restrict__ idx, const float* __restrict__ rk,
                                               float inv_sigma, const float* __restrict__ feats,
                                               float* __restrict__ grouped) {
    __shared__ float s_g[IG_MAXNN * 3];
    __shared__ int s_j[IG_MAXNN];
    const int pi = blockIdx.x, b = blockIdx.y;
    load_neighbourhood(b, pi, n, p, nn, xyz, sxyz, idx, s_j, s_g);
    __syncthreads();
    const int items = a * k * ci;
    for (int t = threadIdx.x; t < items; t += blockDim.x) {
        const int c = t % ci, ki = (t / ci) % k, ai = t / (ci * k);
        const float* kp = rk + ((size_t)ai * k + ki) * 3;
        const float kx = kp[0], ky = kp[1], kz = kp[2];
        float acc = 0.f;
        for (int ni = 0; ni < nn; ++ni) {
            const float dx = s_g[ni * 3] - kx, dy = s_g[ni * 3 + 1] - ky, dz = s_g[ni * 3 + 2] - kz;
            const float w = fmaxf(0.f, 1.f - (dx * dx + dy * dy + dz * dz) * inv_sigma);
            acc = fmaf(w, feats[(((size_t)b * n + s_j[ni]) * a + ai) * ci + c], acc);
        }
        grouped[(((size_t)b * p + pi) * a + ai) * (size_t)k * ci + (size_t)ki * ci + c] = acc;
    }
}

// ci == 1 (first layer: occupancy features): warp <-> anchor, lane <-> kernel point, the rotated kernel point
// stays in registers and the neighbours are a loop -- no per-item weight recomputation
__global__ void __launch_bounds__(IG_WARPS * 32)
inter_group_fwd_c1_kernel(int n, int p, int nn, int a, int k, const float* __restrict__ xyz,
                          const float* __restrict__ sxyz, const int32_t* __restrict__ idx, const float* __restrict__ rk,
                          float inv_sigma, const float* __restrict__ feats, float* __restrict__ grouped) {
    __shared__ float s_g[IG_MAXNN * 3];
    __shared__ int s_j[IG_MAXNN];
    const int pi = blockIdx.x, b = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    load_neighbourhood(b, pi, n, p, nn, xyz, sxyz, idx, s_j, s_g);
    __syncthreads();
    const float* fb = feats + (size_t)b * n * a;
    for (int ai = warp; ai < a; ai += IG_WARPS) {
        if (lane < k) {
            const float* kp = rk + (ai * k + lane) * 3;
            const float kx = __ldg(kp), ky = __ldg(kp + 1), kz = __ldg(kp + 2);
            float acc = 0.f;
            for (int ni = 0; ni < nn; ++ni) {
                const float dx = s_g[ni * 3] - kx, dy = s_g[ni * 3 + 1] - ky, dz = s_g[ni * 3 + 2] - kz;
                const float w = fmaxf(0.f, 1.f - (dx * dx + dy * dy + dz * dz) * inv_sigma);
                acc = fmaf(w, __ldg(fb + (size_t)s_j[ni] * a + ai), acc);
            }
            grouped[(((size_t)b * p + pi) * a + ai) * (size_t)k + lane] = acc;
        }
    }
}

__global__ void inter_group_bwd_generic_kernel(int n, int p, int nn, int a, int k, int ci,
                                               const float* __restrict__ xyz, const float* __restrict__ sxyz,
                                               const int32_t* __restrict__ idx, const float* __restrict__ rk,
                                               float inv_sigma, const float* __restrict__ ggrouped,
                                               float* __restrict__ gfeats) {
    __shared__ float s_g[IG_MAXNN * 3];
    __shared__ int s_j[IG_MAXNN];
    const int pi = blockIdx.x, b = blockIdx.y;
    load_neighbourhood(b, pi, n, p, nn, xyz, sxyz, idx, s_j, s_g);
    __syncthreads();
    const int items = a * nn * ci;
    for (int t = threadIdx.x; t < items; t += blockDim.x) {
        const int c = t % ci, ni = (t / ci) % nn, ai = t / (ci * nn);
        const float* gin = ggrouped + (((size_t)b * p + pi) * a + ai) * (size_t)k * ci;
        float acc = 0.f;
        for (int ki = 0; ki < k; ++ki) {
            const float* kp = rk + ((size_t)ai * k + ki) * 3;
            const float dx = s_g[ni * 3] - kp[0], dy = s_g[ni * 3 + 1] - kp[1], dz = s_g[ni * 3 + 2] - kp[2];
            const float w = fmaxf(0.f, 1.f - (dx * dx + dy * dy + dz * dz) * inv_sigma);
            acc = fmaf(w, gin[(size_t)ki * ci + c], acc);
        }
        atomicAdd(gfeats + (((size_t)b * n + s_j[ni]) * a + ai) * ci + c, acc);
    }
}

// ------------------------------------------------------------------------------ intra grouping
// G[r,a,k,:] = Y[r, intra[a,k], :]   (V = float4 when c % 4 == 0)
template <typename V>
__global__ void intra_group_fwd_kernel(int64_t rows, int a, int kk, int cv, const int32_t* __restrict__ intra,
                                       const V* __restrict__ y, V* __restrict__ g) {
    extern __shared__ int s_idx[];  // [a*kk]
    for (int i = threadIdx.x; i < a * kk; i += blockDim.x) s_idx[i] = intra[i];
    __syncthreads();
    const int64_t per_row = (int64_t)a * kk * cv;
    const int64_t total = rows * per_row;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = t / per_row;
        const int e = (int)(t % per_row);
        const int c = e % cv, ak = e / cv;
        g[t] = y[(r * a + s_idx[ak]) * cv + c];
    }
}

// dY[r,a',:] = sum_k dG[r, inv_k(a'), k, :] when every column of intra is a permutation
// (true for the icosahedral table); otherwise a scan over all (a,k) pairs.
template <typename V>
__device__ __forceinline__ V vadd(V x, V y);
template <>
__device__ __forceinline__ float vadd<float>(float x, float y) { return x + y; }
template <>
__device__ __forceinline__ float4 vadd<float4>(float4 x, float4 y) {
    return make_float4(x.x + y.x, x.y + y.y, x.z + y.z, x.w + y.w);
}
template <typename V>
__device__ __forceinline__ V vzero();
template <>
__device__ __forceinline__ float vzero<float>() { return 0.f; }
template <>
__device__ __forceinline__ float4 vzero<float4>() { return make_float4(0.f, 0.f, 0.f, 0.f); }

template <typename V>
__global__ void intra_group_bwd_kernel(int64_t rows, int a, int kk, int cv, const int32_t* __restrict__ intra,
                                       const V* __restrict__ gg, V* __restrict__ gy) {
    extern __shared__ int s_tab[];  // intra [a*kk] | inv [kk*a]
    int* s_idx = s_tab;
    int* s_inv = s_tab + a * kk;
    __shared__ int s_bad;
    if (threadIdx.x == 0) s_bad = 0;
    for (int i = threadIdx.x; i < a * kk; i += blockDim.x) {
        s_idx[i] = intra[i];
        s_inv[i] = -1;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < a * kk; i += blockDim.x) {
        const int ai = i / kk, ki = i % kk;
        if (atomicExch(&s_inv[ki * a + s_idx[i]], ai) != -1) s_bad = 1;
    }
    __syncthreads();
    const bool perm = s_bad == 0;
    const int64_t per_row = (int64_t)a * cv;
    const int64_t total = rows * per_row;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = t / per_row;
        const int e = (int)(t % per_row);
        const int c = e % cv, at = e / cv;
        V acc = vzero<V>();
        if (perm) {
            for (int ki = 0; ki < kk; ++ki) {
                const int ai = s_inv[ki * a + at];
                acc = vadd<V>(acc, gg[((r * a + ai) * kk + ki) * cv + c]);
            }
        } else {
            for (int i = 0; i < a * kk; ++i)
                if (s_idx[i] == at) acc = vadd<V>(acc, gg[(r * a * kk + i) * cv + c]);
        }
        gy[t] = acc;
    }
}

}  // namespace vgtkb

using namespace vgtkb;

static int check_inter_args(int b, int n, int p, int nn, int a, int k, int ci) {
    VGTKB_REQUIRE(b >= 0 && n > 0 && p >= 0 && nn > 0 && a > 0 && k > 0 && ci > 0, "inter_group: bad size");
    VGTKB_REQUIRE(nn <= IG_MAXNN, "inter_group: nn=%d > %d", nn, IG_MAXNN);
    VGTKB_REQUIRE(b <= 65535, "inter_group: batch %d > 65535", b);
    return VGTKB_OK;
}

extern "C" int vgtkb_inter_weights(int b, int n, int p, int nn, int a, int k, const float* xyz,
                                   const float* sample_xyz, const int32_t* idx, const float* rot_kernels,
                                   float sigma, float* w, void* stream) {
    int rc = check_inter_args(b, n, p, nn, a, k, 1);
    if (rc) return rc;
    if (b == 0 || p == 0) return VGTKB_OK;
    const int64_t per_b = (int64_t)p * a * k * nn;
    dim3 grid((unsigned)ceil_div64(per_b, 256), b);
    inter_weights_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(n, p, nn, a, k, xyz, sample_xyz, idx, rot_kernels,
                                                                 1.0f / sigma, w);
    return check_launch("inter_weights");
}

template <int CPL>
static int launch_inter(bool fwd, int b, int n, int p, int nn, int a, int k, int ci, const float* xyz,
                        const float* sxyz, const int32_t* idx, const float* rk, float sigma, const float* in,
                        float* out, cudaStream_t st) {
    const size_t smem = ((size_t)IG_WARPS * nn * IG_KP + nn * 3 + nn) * 4;
    auto kf = inter_group_fwd_kernel<CPL>;
    auto kb = inter_group_bwd_kernel<CPL>;
    VGTKB_CUDA(cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
    VGTKB_CUDA(cudaFuncSetAttribute(kb, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
    dim3 grid(p, b);
    if (fwd) kf<<<grid, IG_WARPS * 32, smem, st>>>(n, p, nn, a, k, ci, xyz, sxyz, idx, rk, 1.0f / sigma, in, out);
    else kb<<<grid, IG_WARPS * 32, smem, st>>>(n, p, nn, a, k, ci, xyz, sxyz, idx, rk, 1.0f / sigma, in, out);
    return check_launch(fwd ? "inter_group_forward" : "inter_group_backward");
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

extern "C" int vgtkb_inter_group_forward(int b, int n, int p, int nn, int a, int k, int ci, const float* xyz,
                                         const float* sample_xyz, const int32_t* idx, const float* rot_kernels,
                                         float sigma, const float* feats, float* grouped, void* stream) {
    int rc = check_inter_args(b, n, p, nn, a, k, ci);
    if (rc) return rc;
    if (b == 0 || p == 0) return VGTKB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const bool al = aligned16(feats) && aligned16(grouped);
    if (k <= IG_KP && ci % 128 == 0 && al)
        return launch_inter<4>(true, b, n, p, nn, a, k, ci, xyz, sample_xyz, idx, rot_kernels, sigma, feats, grouped, st);
    if (k <= IG_KP && ci % 64 == 0 && al)
        return launch_inter<2>(true, b, n, p, nn, a, k, ci, xyz, sample_xyz, idx, rot_kernels, sigma, feats, grouped, st);
    if (ci == 1 && k <= 32) {
        inter_group_fwd_c1_kernel<<<dim3(p, b), IG_WARPS * 32, 0, st>>>(n, p, nn, a, k, xyz, sample_xyz, idx, rot_kernels,
                                                                         1.0f / sigma, feats, grouped);
        return check_launch("inter_group_forward(ci=1)");
    }
    inter_group_fwd_generic_kernel<<<dim3(p, b), 256, 0, st>>>(n, p, nn, a, k, ci, xyz, sample_xyz, idx, rot_kernels,
                                                              1.0f / sigma, feats, grouped);
    return check_launch("inter_group_forward(generic)");
}

extern "C" int vgtkb_inter_group_backward(int b, int n, int p, int nn, int a, int k, int ci, const float* xyz,
                                          const float* sample_xyz, const int32_t* idx, const float* rot_kernels,
                                          float sigma, const float* grad_grouped, float* grad_feats, void* stream) {
    int rc = check_inter_args(b, n, p, nn, a, k, ci);
    if (rc) return rc;
    if (b == 0 || p == 0) return VGTKB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const bool al = aligned16(grad_feats) && aligned16(grad_grouped);
    if (k <= IG_KP && ci % 128 == 0 && al)
        return launch_inter<4>(false, b, n, p, nn, a, k, ci, xyz, sample_xyz, idx, rot_kernels, sigma, grad_grouped, grad_feats, st);
    if (k <= IG_KP && ci % 64 == 0 && al)
        return launch_inter<2>(false, b, n, p, nn, a, k, ci, xyz, sample_xyz, idx, rot_kernels, sigma, grad_grouped, grad_feats, st);
    inter_group_bwd_generic_kernel<<<dim3(p, b), 256, 0, st>>>(n, p, nn, a, k, ci, xyz, sample_xyz, idx, rot_kernels,
                                                              1.0f / sigma, grad_grouped, grad_feats);
    return check_launch("inter_group_backward(generic)");
}

extern "C" int vgtkb_intra_group_forward(int64_t rows, int a, int kk, int c, const int32_t* intra_idx, const float* y,
                                         float* grouped, void* stream) {
    VGTKB_REQUIRE(rows >= 0 && a > 0 && kk > 0 && c > 0, "intra_group: bad size");
    if (rows == 0) return VGTKB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t smem = (size_t)a * kk * 4;
    VGTKB_REQUIRE(smem <= 40 * 1024, "intra_group: index table too large");
    const bool v4 = c % 4 == 0 && aligned16(y) && aligned16(grouped);
    const int cv = v4 ? c / 4 : c;
    const int64_t total = rows * a * kk * cv;
    const unsigned grid = (unsigned)(ceil_div64(total, 256) < (int64_t)kNumSMs * 16 ? ceil_div64(total, 256) : kNumSMs * 16);
    if (v4) intra_group_fwd_kernel<float4><<<grid, 256, smem, st>>>(rows, a, kk, cv, intra_idx, (const float4*)y, (float4*)grouped);
    else intra_group_fwd_kernel<float><<<grid, 256, smem, st>>>(rows, a, kk, cv, intra_idx, y, grouped);
    return check_launch("intra_group_forward");
}

extern "C" int vgtkb_intra_group_backward(int64_t rows, int a, int kk, int c, const int32_t* intra_idx,
                                          const float* grad_grouped, float* grad_y, void* stream) {
    VGTKB_REQUIRE(rows >= 0 && a > 0 && kk > 0 && c > 0, "intra_group: bad size");
    if (rows == 0) return VGTKB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t smem = (size_t)a * kk * 8;
    VGTKB_REQUIRE(smem <= 40 * 1024, "intra_group: index table too large");
    const bool v4 = c % 4 == 0 && aligned16(grad_y) && aligned16(grad_grouped);
    const int cv = v4 ? c / 4 : c;
    const int64_t total = rows * a * cv;
    const unsigned grid = (unsigned)(ceil_div64(total, 256) < (int64_t)kNumSMs * 16 ? ceil_div64(total, 256) : kNumSMs * 16);
    if (v4) intra_group_bwd_kernel<float4><<<grid, 256, smem, st>>>(rows, a, kk, cv, intra_idx, (const float4*)grad_grouped, (float4*)grad_y);
    else intra_group_bwd_kernel<float><<<grid, 256, smem, st>>>(rows, a, kk, cv, intra_idx, grad_grouped, grad_y);
    return check_launch("intra_group_backward");
}
