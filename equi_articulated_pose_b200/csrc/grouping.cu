// grouping.cu -- kernel-point correlation + inter-/intra-anchor grouping for sm_100a.
//
// The reference evaluates these in eager PyTorch (vgtk/vgtk/so3conv/functional.py:2508-2549,
// 2553-2567 and vgtk/vgtk/spconv/functional.py:375-406) and materialises, per layer, the
// weights w[B,P,A,K,nn] (755 MB at config 2), three permuted copies of the gathered neighbour
// features and the 12x index_select copy.  Here the weights live only in shared memory:
// one CTA owns an output point, a warp owns an anchor, computes its [nn x K] weight tile and
// streams the neighbour rows (channels-last, one coalesced 256/512-byte row segment per load)
// through K*CPL register accumulators.
#include <stdlib.h>

#include "tc_common.cuh"

namespace vgtkb {
using namespace tc;

// timing experiments (VGTKB_DBG_GROUP; results are garbage): fwd 1 = no tensor-map stores, 2 = no gathers, 4 = no MMAs;
// bwd 1 = no tensor-map loads, 2 = no atomics, 4 = no MMAs
__constant__ int c_dbg_group;

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

// ---------------------------------------------------------------------------------- forward on the tensor cores
// bf16x3 variant of the forward grouping (the default contraction arithmetic, mode 3): per (point, anchor) the small
// product  G[c, kp] = sum_n F[c, n] W[n, kp]  (F = gathered neighbour rows, W = kernel-point correlation weights) runs
// as warp-level mma.sync.m16n8k16 bf16 MMAs with fp32 accumulation, both operands split hi/lo exactly like the
// tcgen05 contraction that consumes G (lo*hi + hi*lo + hi*hi: ~1e-5 of max|G|).  Everything stays in registers:
//   * B fragments (weights) are COMPUTED straight into their fragment slots -- lane (gid, tig) owns kernel points
//     8t + gid and neighbours 2tig, 2tig+1, 2tig+8, 2tig+9 of every 16-neighbour step -- no shared-memory weight table;
//   * A fragments come from one 16-byte load per neighbour row and 32-channel chunk: lane gid reads channels
//     4gid..4gid+3 (8 lanes = one 128-byte line) and the MMA rows are PERMUTED so that those four channels are rows
//     (gid, gid+8) of two M tiles; the same permutation turns the accumulator fragments into 16-byte stores.
// The FFMA kernel above issues 12*CPL FFMA per neighbour and lane (45 % of all issued instructions, 70 % issue
// utilisation: instruction-bound); this one needs about a third of the instructions and leaves the HBM write of G
// as the bound.  K <= 24 kernel points (3 N tiles), nn <= 32 (KS = 1 or 2 k-steps), Ci % 32 == 0.
__device__ __forceinline__ uint32_t ig_pack_bf16x2(float lo_elem, float hi_elem) {   // lo_elem -> low half
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi_elem), "f"(lo_elem));
    return r;
}
__device__ __forceinline__ void ig_split2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
    hi = ig_pack_bf16x2(x0, x1);
    lo = ig_pack_bf16x2(x0 - __uint_as_float(hi << 16), x1 - __uint_as_float(hi & 0xFFFF0000u));
}
__device__ __forceinline__ void ig_mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int KS>
__global__ void __launch_bounds__(IG_WARPS * 32, 2)
inter_group_fwd_mma_kernel(int n, int p, int nn, int a, int k, int ci, const float* __restrict__ xyz,
                           const float* __restrict__ sxyz, const int32_t* __restrict__ idx,
                           const float* __restrict__ rk, float inv_sigma, const float* __restrict__ feats,
                           float* __restrict__ grouped) {
    __shared__ float s_g[16 * KS * 3];
    __shared__ uint32_t s_off[16 * KS];
    const int pi = blockIdx.x, b = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gid = lane >> 2, tig = lane & 3;
    {   // neighbour offsets (centre-relative) and row offsets; slots >= nn: weight 0 (offset 1e18 -> relu 0), row 0
        const float* X = xyz + (size_t)b * 3 * n;
        const float* S = sxyz + (size_t)b * 3 * p;
        for (int i = threadIdx.x; i < 16 * KS; i += blockDim.x) {
            if (i < nn) {
                const int j = idx[((size_t)b * p + pi) * nn + i];
                s_off[i] = (uint32_t)j * (uint32_t)(a * ci);
                s_g[i * 3 + 0] = X[j] - S[pi];
                s_g[i * 3 + 1] = X[n + j] - S[p + pi];
                s_g[i * 3 + 2] = X[2 * n + j] - S[2 * p + pi];
            } else {
                s_off[i] = 0;
                s_g[i * 3 + 0] = s_g[i * 3 + 1] = s_g[i * 3 + 2] = 1e12f;
            }
        }
    }
    __syncthreads();
    // this lane's four neighbours per k-step: n0 = 16s + 2tig, n0+1, n0+8, n0+9
    float gx[KS][4], gy[KS][4], gz[KS][4];
    uint32_t off[KS][4];
#pragma unroll
    for (int s = 0; s < KS; ++s)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int ni = 16 * s + 2 * tig + (j & 1) + 8 * (j >> 1);
            gx[s][j] = s_g[ni * 3];
            gy[s][j] = s_g[ni * 3 + 1];
            gz[s][j] = s_g[ni * 3 + 2];
            off[s][j] = s_off[ni];
        }
    const float* fb = feats + (size_t)b * n * a * ci + 4 * gid;
    for (int ai = warp; ai < a; ai += IG_WARPS) {
        // ---- B fragments: weights of kernel points 8t + gid against the lane's neighbours, hi / lo
        uint32_t bh[KS][3][2], bl[KS][3][2];
#pragma unroll
        for (int t = 0; t < 3; ++t) {
            const int kp = 8 * t + gid;
            const bool live = kp < k;
            float kx = 0.f, ky = 0.f, kz = 0.f;
            if (live) {
                const float* q = rk + (ai * k + kp) * 3;
                kx = __ldg(q);
                ky = __ldg(q + 1);
                kz = __ldg(q + 2);
            }
#pragma unroll
            for (int s = 0; s < KS; ++s) {
                float w[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float dx = gx[s][j] - kx, dy = gy[s][j] - ky, dz = gz[s][j] - kz;
                    const float v = fmaxf(0.f, 1.f - (dx * dx + dy * dy + dz * dz) * inv_sigma);
                    w[j] = live ? v : 0.f;
                }
                ig_split2(w[0], w[1], bh[s][t][0], bl[s][t][0]);
                ig_split2(w[2], w[3], bh[s][t][1], bl[s][t][1]);
            }
        }
        float* out = grouped + (((size_t)b * p + pi) * a + ai) * (size_t)k * ci + 4 * gid;
        const float* fa = fb + (size_t)ai * ci;
        for (int c0 = 0; c0 < ci; c0 += 32) {
            float4 v[KS][4];
#pragma unroll
            for (int s = 0; s < KS; ++s)
#pragma unroll
                for (int j = 0; j < 4; ++j) v[s][j] = __ldg(reinterpret_cast<const float4*>(fa + off[s][j] + c0));
            float d[2][3][4];
#pragma unroll
            for (int m = 0; m < 2; ++m)
#pragma unroll
                for (int t = 0; t < 3; ++t)
#pragma unroll
                    for (int e = 0; e < 4; ++e) d[m][t][e] = 0.f;
#pragma unroll
            for (int s = 0; s < KS; ++s) {
                // M tile 0: rows (gid, gid+8) = channels (.x, .y); M tile 1: (.z, .w)
                uint32_t ah[2][4], al[2][4];
                ig_split2(v[s][0].x, v[s][1].x, ah[0][0], al[0][0]);
                ig_split2(v[s][0].y, v[s][1].y, ah[0][1], al[0][1]);
                ig_split2(v[s][2].x, v[s][3].x, ah[0][2], al[0][2]);
                ig_split2(v[s][2].y, v[s][3].y, ah[0][3], al[0][3]);
                ig_split2(v[s][0].z, v[s][1].z, ah[1][0], al[1][0]);
                ig_split2(v[s][0].w, v[s][1].w, ah[1][1], al[1][1]);
                ig_split2(v[s][2].z, v[s][3].z, ah[1][2], al[1][2]);
                ig_split2(v[s][2].w, v[s][3].w, ah[1][3], al[1][3]);
#pragma unroll
                for (int m = 0; m < 2; ++m)
#pragma unroll
                    for (int t = 0; t < 3; ++t) {
                        ig_mma(d[m][t], al[m], bh[s][t][0], bh[s][t][1]);
                        ig_mma(d[m][t], ah[m], bl[s][t][0], bl[s][t][1]);
                        ig_mma(d[m][t], ah[m], bh[s][t][0], bh[s][t][1]);
                    }
            }
            // accumulator (row gid | gid+8, col 2tig | 2tig+1) -> channels 4gid.. of kernel points 8t + 2tig (+1)
#pragma unroll
            for (int t = 0; t < 3; ++t) {
                const int kp = 8 * t + 2 * tig;
                if (kp < k)
                    __stcs(reinterpret_cast<float4*>(out + (size_t)kp * ci + c0),
                           make_float4(d[0][t][0], d[0][t][2], d[1][t][0], d[1][t][2]));
                if (kp + 1 < k)
                    __stcs(reinterpret_cast<float4*>(out + (size_t)(kp + 1) * ci + c0),
                           make_float4(d[0][t][1], d[0][t][3], d[1][t][1], d[1][t][3]));
            }
        }
    }
}

// backward on the tensor cores (mode 3): dF[c, n] = sum_kp dG[kp, c] W[n, kp] per (point, anchor), same scheme as the
// forward kernel with the roles turned: M = channels (permuted rows, 16-byte loads of dG rows), N = neighbours (tiles
// of 8), K = kernel points = one k16 step + one k8 step (24 exactly).  The accumulator fragments are four adjacent
// channels of one neighbour: one 16-byte red.global.add per fragment pair.  (The FFMA kernel below is instruction
// bound -- 70 % issue utilisation, L2 at 30 % -- not atomics bound.)
__device__ __forceinline__ void ig_mma_k8(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a0), "r"(a1), "r"(b0));
}

template <int KS>
__global__ void __launch_bounds__(IG_WARPS * 32, 2)
inter_group_bwd_mma_kernel(int n, int p, int nn, int a, int k, int ci, const float* __restrict__ xyz,
                           const float* __restrict__ sxyz, const int32_t* __restrict__ idx,
                           const float* __restrict__ rk, float inv_sigma, const float* __restrict__ ggrouped,
                           float* __restrict__ gfeats) {
    constexpr int NT = 2 * KS;                       // neighbour tiles of 8
    __shared__ float s_g[16 * KS * 3];
    __shared__ uint32_t s_off[16 * KS];
    const int pi = blockIdx.x, b = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gid = lane >> 2, tig = lane & 3;
    {
        const float* X = xyz + (size_t)b * 3 * n;
        const float* S = sxyz + (size_t)b * 3 * p;
        for (int i = threadIdx.x; i < 16 * KS; i += blockDim.x) {
            if (i < nn) {
                const int j = idx[((size_t)b * p + pi) * nn + i];
                s_off[i] = (uint32_t)j * (uint32_t)(a * ci);
                s_g[i * 3 + 0] = X[j] - S[pi];
                s_g[i * 3 + 1] = X[n + j] - S[p + pi];
                s_g[i * 3 + 2] = X[2 * n + j] - S[2 * p + pi];
            } else {
                s_off[i] = 0;
                s_g[i * 3 + 0] = s_g[i * 3 + 1] = s_g[i * 3 + 2] = 1e12f;
            }
        }
    }
    __syncthreads();
    // B side: this lane's neighbour of tile t is 8t + gid; D side: neighbours 8t + 2tig, 8t + 2tig + 1
    float gx[NT], gy[NT], gz[NT];
    uint32_t off[NT][2];
#pragma unroll
    for (int t = 0; t < NT; ++t) {
        gx[t] = s_g[(8 * t + gid) * 3];
        gy[t] = s_g[(8 * t + gid) * 3 + 1];
        gz[t] = s_g[(8 * t + gid) * 3 + 2];
        off[t][0] = s_off[8 * t + 2 * tig];
        off[t][1] = s_off[8 * t + 2 * tig + 1];
    }
    // this lane's kernel points: k16 step 2tig, 2tig+1, 2tig+8, 2tig+9; k8 step 16+2tig, 17+2tig
    int kps[6];
#pragma unroll
    for (int j = 0; j < 4; ++j) kps[j] = 2 * tig + (j & 1) + 8 * (j >> 1);
    kps[4] = 16 + 2 * tig;
    kps[5] = 17 + 2 * tig;
    float* gb = gfeats + (size_t)b * n * a * ci + 4 * gid;
    for (int ai = warp; ai < a; ai += IG_WARPS) {
        uint32_t bh[NT][3], bl[NT][3];
        {
            float kx[6], ky[6], kz[6];
#pragma unroll
            for (int j = 0; j < 6; ++j) {
                kx[j] = ky[j] = kz[j] = 1e12f;                      // kp >= k: weight 0
                if (kps[j] < k) {
                    const float* q = rk + (ai * k + kps[j]) * 3;
                    kx[j] = __ldg(q);
                    ky[j] = __ldg(q + 1);
                    kz[j] = __ldg(q + 2);
                }
            }
#pragma unroll
            for (int t = 0; t < NT; ++t) {
                float w[6];
#pragma unroll
                for (int j = 0; j < 6; ++j) {
                    const float dx = gx[t] - kx[j], dy = gy[t] - ky[j], dz = gz[t] - kz[j];
                    const float v = fmaxf(0.f, 1.f - (dx * dx + dy * dy + dz * dz) * inv_sigma);
                    w[j] = (kps[j] < k && 8 * t + gid < nn) ? v : 0.f;
                }
                ig_split2(w[0], w[1], bh[t][0], bl[t][0]);
                ig_split2(w[2], w[3], bh[t][1], bl[t][1]);
                ig_split2(w[4], w[5], bh[t][2], bl[t][2]);
            }
        }
        const float* gin = ggrouped + (((size_t)b * p + pi) * a + ai) * (size_t)k * ci + 4 * gid;
        float* ga = gb + (size_t)ai * ci;
        for (int c0 = 0; c0 < ci; c0 += 32) {
            float4 v[6];
#pragma unroll
            for (int j = 0; j < 6; ++j)
                v[j] = kps[j] < k ? __ldcs(reinterpret_cast<const float4*>(gin + (size_t)kps[j] * ci + c0))
                                  : make_float4(0.f, 0.f, 0.f, 0.f);
            uint32_t ah[2][6], al[2][6];            // per M tile: k16 a0..a3, k8 a0, a1
            ig_split2(v[0].x, v[1].x, ah[0][0], al[0][0]);
            ig_split2(v[0].y, v[1].y, ah[0][1], al[0][1]);
            ig_split2(v[2].x, v[3].x, ah[0][2], al[0][2]);
            ig_split2(v[2].y, v[3].y, ah[0][3], al[0][3]);
            ig_split2(v[4].x, v[5].x, ah[0][4], al[0][4]);
            ig_split2(v[4].y, v[5].y, ah[0][5], al[0][5]);
            ig_split2(v[0].z, v[1].z, ah[1][0], al[1][0]);
            ig_split2(v[0].w, v[1].w, ah[1][1], al[1][1]);
            ig_split2(v[2].z, v[3].z, ah[1][2], al[1][2]);
            ig_split2(v[2].w, v[3].w, ah[1][3], al[1][3]);
            ig_split2(v[4].z, v[5].z, ah[1][4], al[1][4]);
            ig_split2(v[4].w, v[5].w, ah[1][5], al[1][5]);
#pragma unroll
            for (int t = 0; t < NT; ++t) {
                float d[2][4];
#pragma unroll
                for (int m = 0; m < 2; ++m) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) d[m][e] = 0.f;
                    const uint32_t (&h)[6] = ah[m];
                    const uint32_t (&l)[6] = al[m];
                    const uint32_t h16[4] = {h[0], h[1], h[2], h[3]}, l16[4] = {l[0], l[1], l[2], l[3]};
                    ig_mma(d[m], l16, bh[t][0], bh[t][1]);
                    ig_mma(d[m], h16, bl[t][0], bl[t][1]);
                    ig_mma(d[m], h16, bh[t][0], bh[t][1]);
                    ig_mma_k8(d[m], l[4], l[5], bh[t][2]);
                    ig_mma_k8(d[m], h[4], h[5], bl[t][2]);
                    ig_mma_k8(d[m], h[4], h[5], bh[t][2]);
                }
                const int n0 = 8 * t + 2 * tig;
                if (n0 < nn)
                    atomicAdd(reinterpret_cast<float4*>(ga + off[t][0] + c0), make_float4(d[0][0], d[0][2], d[1][0], d[1][2]));
                if (n0 + 1 < nn)
                    atomicAdd(reinterpret_cast<float4*>(ga + off[t][1] + c0), make_float4(d[0][1], d[0][3], d[1][1], d[1][3]));
            }
        }
    }
}

// ---------------------------------------------------------------------------------- bulk-copy (TMA) variants
// ncu on the two kernels above: l1tex__data_pipe_lsu_wavefronts at 80-90 % with one 32-byte sector per wavefront --
// a quad's four lanes own four different neighbours / kernel points, so every 16-byte gather or scatter instruction
// is split into 16 wavefronts -- while issue slots, the tensor pipe, L2 and HBM idle.  These variants take the global
// traffic off the LSU: a warp's neighbour rows (256 B = 64 channels each) arrive in shared memory through
// cp.async.bulk (completion on a per-warp mbarrier, double buffered over the warp's (anchor, channel-block) items),
// the fragments are read from / written to padded shared rows with conflict-free 16-byte accesses (4 wavefronts per
// instruction), and the results leave through cp.async.bulk stores (forward: 24 rows of G) or cp.reduce.async.bulk
// .add.f32 (backward: one row per neighbour -- the scatter-add runs in the L2, no per-lane atomics).
constexpr int IGT_WARPS = 4;           // 60 anchors = 15 per warp
constexpr int IGT_CB = 64;             // channels per item
constexpr int IGT_LD = IGT_CB + 4;     // padded row (floats): fragment accesses of a quarter warp hit 32 distinct banks

__device__ __forceinline__ void bulk_s2g(void* gdst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_s2g_add_f32(void* gdst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(gdst),
                 "r"(smem_u32(smem_src)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ float4 lds128(const float* p) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(smem_u32(p)));
    return v;
}
__device__ __forceinline__ void sts128(float* p, float a, float b, float c, float d) {
    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(smem_u32(p)), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// shared prologue of the two kernels: neighbour offsets / row offsets of point (b, pi) for the whole CTA
template <int KS>
__device__ __forceinline__ void igt_neighbourhood(int b, int pi, int n, int p, int nn, int a, int ci, const float* xyz,
                                                  const float* sxyz, const int32_t* idx, float* s_g, uint32_t* s_off) {
    const float* X = xyz + (size_t)b * 3 * n;
    const float* S = sxyz + (size_t)b * 3 * p;
    for (int i = threadIdx.x; i < 16 * KS; i += blockDim.x) {
        if (i < nn) {
            const int j = idx[((size_t)b * p + pi) * nn + i];
            s_off[i] = (uint32_t)j * (uint32_t)(a * ci);
            s_g[i * 3 + 0] = X[j] - S[pi];
            s_g[i * 3 + 1] = X[n + j] - S[p + pi];
            s_g[i * 3 + 2] = X[2 * n + j] - S[2 * p + pi];
        } else {
            s_off[i] = 0;
            s_g[i * 3 + 0] = s_g[i * 3 + 1] = s_g[i * 3 + 2] = 1e12f;
        }
    }
}

template <int KS>
__global__ void __launch_bounds__(IGT_WARPS * 32)
inter_group_fwd_tma_kernel(int n, int p, int nn, int a, int k, int ci, const float* __restrict__ xyz,
                           const float* __restrict__ sxyz, const int32_t* __restrict__ idx,
                           const float* __restrict__ rk, float inv_sigma, const float* __restrict__ feats,
                           float* __restrict__ grouped) {
    constexpr int NNP = 16 * KS;
    constexpr int IN_FLOATS = NNP * IGT_LD, OUT_FLOATS = 24 * IGT_LD, WARP_FLOATS = 2 * IN_FLOATS + OUT_FLOATS;
    extern __shared__ __align__(16) float smem[];
    __shared__ float s_g[NNP * 3];
    __shared__ uint32_t s_off[NNP];
    __shared__ __align__(8) uint64_t s_bar[IGT_WARPS][2];
    const int pi = blockIdx.x, b = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gid = lane >> 2, tig = lane & 3;
    float* inb = smem + warp * WARP_FLOATS;             // [2][NNP][IGT_LD]
    float* outb = inb + 2 * IN_FLOATS;                  // [24][IGT_LD]
    igt_neighbourhood<KS>(b, pi, n, p, nn, a, ci, xyz, sxyz, idx, s_g, s_off);
    for (int i = lane; i < 2 * IN_FLOATS; i += 32) inb[i] = 0.f;      // rows >= nn are never loaded: 0 * weight 0
    if (lane == 0) {
        mbar_init(&s_bar[warp][0], 1);
        mbar_init(&s_bar[warp][1], 1);
        fence_barrier_init();
    }
    fence_proxy_async();
    __syncthreads();
    float gx[KS][4], gy[KS][4], gz[KS][4];
#pragma unroll
    for (int s = 0; s < KS; ++s)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int ni = 16 * s + 2 * tig + (j & 1) + 8 * (j >> 1);
            gx[s][j] = s_g[ni * 3];
            gy[s][j] = s_g[ni * 3 + 1];
            gz[s][j] = s_g[ni * 3 + 2];
        }
    const uint32_t my_off = lane < nn ? s_off[lane] : 0u;              // lane <-> neighbour row it fetches
    const float* fb = feats + (size_t)b * n * a * ci;
    const int cbs = ci / IGT_CB;
    const int my_anchors = (a - warp + IGT_WARPS - 1) / IGT_WARPS;      // anchors warp, warp + 4, ...
    const int items = my_anchors * cbs;                                // item -> (anchor slot, channel block)
    auto issue = [&](int item) {
        const int ai = warp + (item / cbs) * IGT_WARPS, cb = item % cbs, buf = item & 1;
        if (lane == 0) mbar_arrive_expect_tx(&s_bar[warp][buf], (uint32_t)nn * IGT_CB * 4u);
        __syncwarp();
        if (lane < nn)
            bulk_g2s(inb + buf * IN_FLOATS + lane * IGT_LD, fb + my_off + (size_t)ai * ci + cb * IGT_CB, IGT_CB * 4, &s_bar[warp][buf]);
    };
    if (items > 0) issue(0);
    uint32_t bh[KS][3][2], bl[KS][3][2];
    for (int item = 0; item < items; ++item) {
        const int ai = warp + (item / cbs) * IGT_WARPS, cb = item % cbs, buf = item & 1;
        if (item + 1 < items) issue(item + 1);          // its buffer was read two items ago (program order + __syncwarp)
        if (cb == 0) {
            // ---- B fragments of this anchor (weights), computed in registers, hi / lo
#pragma unroll
            for (int t = 0; t < 3; ++t) {
                const int kp = 8 * t + gid;
                const bool live = kp < k;
                float kx = 0.f, ky = 0.f, kz = 0.f;
                if (live) {
                    const float* q = rk + (ai * k + kp) * 3;
                    kx = __ldg(q);
                    ky = __ldg(q + 1);
                    kz = __ldg(q + 2);
                }
#pragma unroll
                for (int s = 0; s < KS; ++s) {
                    float w[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float dx = gx[s][j] - kx, dy = gy[s][j] - ky, dz = gz[s][j] - kz;
                        const float v = fmaxf(0.f, 1.f - (dx * dx + dy * dy + dz * dz) * inv_sigma);
                        w[j] = live ? v : 0.f;
                    }
                    ig_split2(w[0], w[1], bh[s][t][0], bl[s][t][0]);
                    ig_split2(w[2], w[3], bh[s][t][1], bl[s][t][1]);
                }
            }
        }
        mbar_wait(&s_bar[warp][buf], (uint32_t)(item >> 1) & 1u);
        const float* in = inb + buf * IN_FLOATS;
        // the previous item's bulk stores must have finished READING the staging rows before they are rewritten
        if (lane < 24) bulk_wait_read_all();
        __syncwarp();
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int c0 = half * 32 + 4 * gid;
            float d[2][3][4];
#pragma unroll
            for (int m = 0; m < 2; ++m)
#pragma unroll
                for (int t = 0; t < 3; ++t)
#pragma unroll
                    for (int e = 0; e < 4; ++e) d[m][t][e] = 0.f;
#pragma unroll
            for (int s = 0; s < KS; ++s) {
                float4 v[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) v[j] = lds128(in + (16 * s + 2 * tig + (j & 1) + 8 * (j >> 1)) * IGT_LD + c0);
                uint32_t ah[2][4], al[2][4];
                ig_split2(v[0].x, v[1].x, ah[0][0], al[0][0]);
                ig_split2(v[0].y, v[1].y, ah[0][1], al[0][1]);
                ig_split2(v[2].x, v[3].x, ah[0][2], al[0][2]);
                ig_split2(v[2].y, v[3].y, ah[0][3], al[0][3]);
                ig_split2(v[0].z, v[1].z, ah[1][0], al[1][0]);
                ig_split2(v[0].w, v[1].w, ah[1][1], al[1][1]);
                ig_split2(v[2].z, v[3].z, ah[1][2], al[1][2]);
                ig_split2(v[2].w, v[3].w, ah[1][3], al[1][3]);
#pragma unroll
                for (int m = 0; m < 2; ++m)
#pragma unroll
                    for (int t = 0; t < 3; ++t) {
                        ig_mma(d[m][t], al[m], bh[s][t][0], bh[s][t][1]);
                        ig_mma(d[m][t], ah[m], bl[s][t][0], bl[s][t][1]);
                        ig_mma(d[m][t], ah[m], bh[s][t][0], bh[s][t][1]);
                    }
            }
#pragma unroll
            for (int t = 0; t < 3; ++t) {
                const int kp = 8 * t + 2 * tig;
                sts128(outb + kp * IGT_LD + c0, d[0][t][0], d[0][t][2], d[1][t][0], d[1][t][2]);
                sts128(outb + (kp + 1) * IGT_LD + c0, d[0][t][1], d[0][t][3], d[1][t][1], d[1][t][3]);
            }
        }
        fence_proxy_async();                             // staging rows -> visible to the bulk-copy engine
        __syncwarp();
        if (lane < k) {                                  // one 256-byte row of G per kernel point
            float* out = grouped + ((((size_t)b * p + pi) * a + ai) * (size_t)k + lane) * ci + cb * IGT_CB;
            bulk_s2g(out, outb + lane * IGT_LD, IGT_CB * 4);
        }
        bulk_commit();
    }
    bulk_wait_all();                                     // smem must outlive the last stores
}

template <int KS>
__global__ void __launch_bounds__(IGT_WARPS * 32)
inter_group_bwd_tma_kernel(int n, int p, int nn, int a, int k, int ci, const float* __restrict__ xyz,
                           const float* __restrict__ sxyz, const int32_t* __restrict__ idx,
                           const float* __restrict__ rk, float inv_sigma, const float* __restrict__ ggrouped,
                           float* __restrict__ gfeats) {
    constexpr int NNP = 16 * KS, NT = 2 * KS;
    constexpr int IN_FLOATS = 24 * IGT_LD, OUT_FLOATS = NNP * IGT_LD, WARP_FLOATS = 2 * IN_FLOATS + OUT_FLOATS;
    extern __shared__ __align__(16) float smem[];
    __shared__ float s_g[NNP * 3];
    __shared__ uint32_t s_off[NNP];
    __shared__ __align__(8) uint64_t s_bar[IGT_WARPS][2];
    const int pi = blockIdx.x, b = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gid = lane >> 2, tig = lane & 3;
    float* inb = smem + warp * WARP_FLOATS;             // [2][24][IGT_LD]  dG rows
    float* outb = inb + 2 * IN_FLOATS;                  // [NNP][IGT_LD]    dF rows, one per neighbour
    igt_neighbourhood<KS>(b, pi, n, p, nn, a, ci, xyz, sxyz, idx, s_g, s_off);
    for (int i = lane; i < 2 * IN_FLOATS; i += 32) inb[i] = 0.f;      // kernel-point rows >= k stay zero
    if (lane == 0) {
        mbar_init(&s_bar[warp][0], 1);
        mbar_init(&s_bar[warp][1], 1);
        fence_barrier_init();
    }
    fence_proxy_async();
    __syncthreads();
    float gx[NT], gy[NT], gz[NT];
#pragma unroll
    for (int t = 0; t < NT; ++t) {
        gx[t] = s_g[(8 * t + gid) * 3];
        gy[t] = s_g[(8 * t + gid) * 3 + 1];
        gz[t] = s_g[(8 * t + gid) * 3 + 2];
    }
    const uint32_t my_off = lane < nn ? s_off[lane] : 0u;              // lane <-> neighbour row it scatters
    int kps[6];
#pragma unroll
    for (int j = 0; j < 4; ++j) kps[j] = 2 * tig + (j & 1) + 8 * (j >> 1);
    kps[4] = 16 + 2 * tig;
    kps[5] = 17 + 2 * tig;
    float* gb = gfeats + (size_t)b * n * a * ci;
    const int cbs = ci / IGT_CB;
    const int my_anchors = (a - warp + IGT_WARPS - 1) / IGT_WARPS;
    const int items = my_anchors * cbs;
    auto issue = [&](int item) {
        const int ai = warp + (item / cbs) * IGT_WARPS, cb = item % cbs, buf = item & 1;
        if (lane == 0) mbar_arrive_expect_tx(&s_bar[warp][buf], (uint32_t)k * IGT_CB * 4u);
        __syncwarp();
        if (lane < k)
            bulk_g2s(inb + buf * IN_FLOATS + lane * IGT_LD,
                     ggrouped + ((((size_t)b * p + pi) * a + ai) * (size_t)k + lane) * ci + cb * IGT_CB, IGT_CB * 4,
                     &s_bar[warp][buf]);
    };
    if (items > 0) issue(0);
    uint32_t bh[NT][3], bl[NT][3];
    for (int item = 0; item < items; ++item) {
        const int ai = warp + (item / cbs) * IGT_WARPS, cb = item % cbs, buf = item & 1;
        if (item + 1 < items) issue(item + 1);
        if (cb == 0) {
            float kx[6], ky[6], kz[6];
#pragma unroll
            for (int j = 0; j < 6; ++j) {
                kx[j] = ky[j] = kz[j] = 1e12f;
                if (kps[j] < k) {
                    const float* q = rk + (ai * k + kps[j]) * 3;
                    kx[j] = __ldg(q);
                    ky[j] = __ldg(q + 1);
                    kz[j] = __ldg(q + 2);
                }
            }
#pragma unroll
            for (int t = 0; t < NT; ++t) {
                float w[6];
#pragma unroll
                for (int j = 0; j < 6; ++j) {
                    const float dx = gx[t] - kx[j], dy = gy[t] - ky[j], dz = gz[t] - kz[j];
                    const float v = fmaxf(0.f, 1.f - (dx * dx + dy * dy + dz * dz) * inv_sigma);
                    w[j] = (kps[j] < k && 8 * t + gid < nn) ? v : 0.f;
                }
                ig_split2(w[0], w[1], bh[t][0], bl[t][0]);
                ig_split2(w[2], w[3], bh[t][1], bl[t][1]);
                ig_split2(w[4], w[5], bh[t][2], bl[t][2]);
            }
        }
        mbar_wait(&s_bar[warp][buf], (uint32_t)(item >> 1) & 1u);
        const float* in = inb + buf * IN_FLOATS;
        if (lane < nn) bulk_wait_read_all();            // previous item's reduce-adds have read the staging rows
        __syncwarp();
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int c0 = half * 32 + 4 * gid;
            float4 v[6];
#pragma unroll
            for (int j = 0; j < 6; ++j) v[j] = lds128(in + kps[j] * IGT_LD + c0);
            uint32_t ah[2][6], al[2][6];
            ig_split2(v[0].x, v[1].x, ah[0][0], al[0][0]);
            ig_split2(v[0].y, v[1].y, ah[0][1], al[0][1]);
            ig_split2(v[2].x, v[3].x, ah[0][2], al[0][2]);
            ig_split2(v[2].y, v[3].y, ah[0][3], al[0][3]);
            ig_split2(v[4].x, v[5].x, ah[0][4], al[0][4]);
            ig_split2(v[4].y, v[5].y, ah[0][5], al[0][5]);
            ig_split2(v[0].z, v[1].z, ah[1][0], al[1][0]);
            ig_split2(v[0].w, v[1].w, ah[1][1], al[1][1]);
            ig_split2(v[2].z, v[3].z, ah[1][2], al[1][2]);
            ig_split2(v[2].w, v[3].w, ah[1][3], al[1][3]);
            ig_split2(v[4].z, v[5].z, ah[1][4], al[1][4]);
            ig_split2(v[4].w, v[5].w, ah[1][5], al[1][5]);
#pragma unroll
            for (int t = 0; t < NT; ++t) {
                float d[2][4];
#pragma unroll
                for (int m = 0; m < 2; ++m) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) d[m][e] = 0.f;
                    const uint32_t h16[4] = {ah[m][0], ah[m][1], ah[m][2], ah[m][3]};
                    const uint32_t l16[4] = {al[m][0], al[m][1], al[m][2], al[m][3]};
                    ig_mma(d[m], l16, bh[t][0], bh[t][1]);
                    ig_mma(d[m], h16, bl[t][0], bl[t][1]);
                    ig_mma(d[m], h16, bh[t][0], bh[t][1]);
                    ig_mma_k8(d[m], al[m][4], al[m][5], bh[t][2]);
                    ig_mma_k8(d[m], ah[m][4], ah[m][5], bl[t][2]);
                    ig_mma_k8(d[m], ah[m][4], ah[m][5], bh[t][2]);
                }
                const int n0 = 8 * t + 2 * tig;
                sts128(outb + n0 * IGT_LD + c0, d[0][0], d[0][2], d[1][0], d[1][2]);
                sts128(outb + (n0 + 1) * IGT_LD + c0, d[0][1], d[0][3], d[1][1], d[1][3]);
            }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane < nn)                                   // dF row of neighbour `lane`: reduce-add in the L2
            bulk_s2g_add_f32(gb + my_off + (size_t)ai * ci + cb * IGT_CB, outb + lane * IGT_LD, IGT_CB * 4);
        bulk_commit();
    }
    bulk_wait_all();
}

// ---------------------------------------------------------------------------------- tensor-map variants (default)
// Measured: 256-byte cp.async.bulk copies cost ~15 cycles each in the copy engine, so the all-bulk kernels above are
// slower than the register gathers.  These keep the register path for the scattered side (neighbour rows forward,
// red.global.add backward) and move only the CONTIGUOUS side -- the G / dG block of one (point, anchor, 32 channels):
// k rows of 128 bytes -- through ONE 2-D tensor-map copy (3 KB) into / out of a 128B-swizzled staging tile, which the
// fragment accesses hit conflict free (chunk = gid ^ (row & 7)).  LSU wavefronts per step drop by ~45 %.
// PL = true: G leaves as two bf16 planes (hi = bf16_rn(g), lo = bf16_rn(g - hi): the operand split of the bf16x3
// contraction, done HERE once instead of by the converter warps of every kernel that consumes G); map_g = hi plane,
// map_lo = lo plane, both [rows*k, ci] bf16 with boxes of k rows x 32 channels (64 bytes, no swizzle).  A staging buffer
// holds the hi tile (24 x 64 B) followed by the lo tile; the 8-byte fragment stores alternate between the even and the
// odd kernel-point row across the quad pairs so that one store instruction covers all 32 banks (2 wavefronts).
template <int KS, bool PL = false, int NW = IG_WARPS, int MINB = 2>
__global__ void __launch_bounds__(NW * 32, MINB)
inter_group_fwd_mma_ts_kernel(const __grid_constant__ TmaMap map_g, int n, int p, int nn, int a, int k, int ci,
                              const float* __restrict__ xyz, const float* __restrict__ sxyz, const int32_t* __restrict__ idx,
                              const float* __restrict__ rk, float inv_sigma, const float* __restrict__ feats,
                              const __grid_constant__ TmaMap map_lo) {
    extern __shared__ unsigned char smem_raw[];
    __shared__ float s_g[16 * KS * 3];
    __shared__ uint32_t s_off[16 * KS];
    const int pi = blockIdx.x, b = blockIdx.y;
    // warp index through a shuffle: the compiler then treats it (and the staging address, the row coordinate ...) as
    // warp-uniform, so the tensor-map stores below take their operands from uniform registers
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const int gid = lane >> 2, tig = lane & 3;
    // per warp: two staging tiles [24 rows][128 B], 1024-byte aligned (swizzle atom = 8 rows x 128 B)
    float* stage = reinterpret_cast<float*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u)) + warp * (2 * 24 * 32);
    igt_neighbourhood<KS>(b, pi, n, p, nn, a, ci, xyz, sxyz, idx, s_g, s_off);
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&map_g);
        if (PL) tma_prefetch_desc(&map_lo);
    }
    __syncthreads();
    float gx[KS][4], gy[KS][4], gz[KS][4];
    uint32_t off[KS][4];
#pragma unroll
    for (int s = 0; s < KS; ++s)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int ni = 16 * s + 2 * tig + (j & 1) + 8 * (j >> 1);
            gx[s][j] = s_g[ni * 3];
            gy[s][j] = s_g[ni * 3 + 1];
            gz[s][j] = s_g[ni * 3 + 2];
            off[s][j] = s_off[ni];
        }
    const float* fb = feats + (size_t)b * n * a * ci + 4 * gid;
    int sb = 0;                                                    // staging tile of the next chunk
    float4 v[KS][4];                                               // gathered neighbour rows of the chunk in flight
    const int dbg = c_dbg_group;
    if (dbg & 2) {
#pragma unroll
        for (int s = 0; s < KS; ++s)
#pragma unroll
            for (int j = 0; j < 4; ++j) v[s][j] = make_float4(1.f, 2.f, 3.f, 4.f);
    }
    for (int ai = warp; ai < a; ai += NW) {
        uint32_t bh[KS][3][2], bl[KS][3][2];
#pragma unroll
        for (int t = 0; t < 3; ++t) {
            const int kp = 8 * t + gid;
            const bool live = kp < k;
            float kx = 0.f, ky = 0.f, kz = 0.f;
            if (live) {
                const float* q = rk + (ai * k + kp) * 3;
                kx = __ldg(q);
                ky = __ldg(q + 1);
                kz = __ldg(q + 2);
            }
#pragma unroll
            for (int s = 0; s < KS; ++s) {
                float w[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float dx = gx[s][j] - kx, dy = gy[s][j] - ky, dz = gz[s][j] - kz;
                    const float v = fmaxf(0.f, 1.f - (dx * dx + dy * dy + dz * dz) * inv_sigma);
                    w[j] = live ? v : 0.f;
                }
                ig_split2(w[0], w[1], bh[s][t][0], bl[s][t][0]);
                ig_split2(w[2], w[3], bh[s][t][1], bl[s][t][1]);
            }
        }
        const int row0 = (((b * p + pi) * a) + ai) * k;            // first row of this (point, anchor) in G [rows, ci]
        const float* fa = fb + (size_t)ai * ci;
        if (ai == warp && !(dbg & 2)) {                            // first chunk of the warp; later ones are prefetched
#pragma unroll
            for (int s = 0; s < KS; ++s)
#pragma unroll
                for (int j = 0; j < 4; ++j) v[s][j] = __ldg(reinterpret_cast<const float4*>(fa + off[s][j]));
        }
        for (int c0 = 0; c0 < ci; c0 += 32, sb ^= 1) {
            uint32_t ah[KS][2][4], al[KS][2][4];
#pragma unroll
            for (int s = 0; s < KS; ++s) {
                ig_split2(v[s][0].x, v[s][1].x, ah[s][0][0], al[s][0][0]);
                ig_split2(v[s][0].y, v[s][1].y, ah[s][0][1], al[s][0][1]);
                ig_split2(v[s][2].x, v[s][3].x, ah[s][0][2], al[s][0][2]);
                ig_split2(v[s][2].y, v[s][3].y, ah[s][0][3], al[s][0][3]);
                ig_split2(v[s][0].z, v[s][1].z, ah[s][1][0], al[s][1][0]);
                ig_split2(v[s][0].w, v[s][1].w, ah[s][1][1], al[s][1][1]);
                ig_split2(v[s][2].z, v[s][3].z, ah[s][1][2], al[s][1][2]);
                ig_split2(v[s][2].w, v[s][3].w, ah[s][1][3], al[s][1][3]);
            }
            {   // prefetch the next chunk (same anchor, or the first chunk of the warp's next anchor): the barriers of the
                // store sequence below would otherwise keep these loads from overlapping it
                const bool more_c = c0 + 32 < ci;
                const float* nf = more_c ? fa + c0 + 32 : fa + (size_t)NW * ci;
                if ((more_c || ai + NW < a) && !(dbg & 2)) {
#pragma unroll
                    for (int s = 0; s < KS; ++s)
#pragma unroll
                        for (int j = 0; j < 4; ++j) v[s][j] = __ldg(reinterpret_cast<const float4*>(nf + off[s][j]));
                }
            }
            float d[2][3][4];
#pragma unroll
            for (int m = 0; m < 2; ++m)
#pragma unroll
                for (int t = 0; t < 3; ++t)
#pragma unroll
                    for (int e = 0; e < 4; ++e) d[m][t][e] = 0.f;
            if (!(dbg & 4))
#pragma unroll
            for (int s = 0; s < KS; ++s)
#pragma unroll
                for (int m = 0; m < 2; ++m)
#pragma unroll
                    for (int t = 0; t < 3; ++t) {
                        ig_mma(d[m][t], al[s][m], bh[s][t][0], bh[s][t][1]);
                        ig_mma(d[m][t], ah[s][m], bl[s][t][0], bl[s][t][1]);
                        ig_mma(d[m][t], ah[s][m], bh[s][t][0], bh[s][t][1]);
                    }
            // the store that used this staging tile two chunks ago must have read it
            // (elect.sync names the same lane every time: bulk async-groups belong to the issuing thread)
            if (!(dbg & 1) && tc::elect_one()) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            __syncwarp();
            float* st = stage + sb * (24 * 32);
            if (PL) {
                const uint32_t sh = smem_u32(st), sl = sh + 24 * 64;
                const int flip = tig >> 1;                         // quads 2, 3 store the odd row first
#pragma unroll
                for (int t = 0; t < 3; ++t) {
                    const int kp = 8 * t + 2 * tig;
                    uint32_t h[2][2], l[2][2];                     // [row parity][channel pair]
                    ig_split2(d[0][t][0], d[0][t][2], h[0][0], l[0][0]);
                    ig_split2(d[1][t][0], d[1][t][2], h[0][1], l[0][1]);
                    ig_split2(d[0][t][1], d[0][t][3], h[1][0], l[1][0]);
                    ig_split2(d[1][t][1], d[1][t][3], h[1][1], l[1][1]);
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        const int par = q ^ flip;
                        const uint32_t o = (uint32_t)((kp + par) * 64 + gid * 8);
                        st_shared_u2(sh + o, par ? h[1][0] : h[0][0], par ? h[1][1] : h[0][1]);
                        st_shared_u2(sl + o, par ? l[1][0] : l[0][0], par ? l[1][1] : l[0][1]);
                    }
                }
            } else
#pragma unroll
            for (int t = 0; t < 3; ++t) {
                const int kp = 8 * t + 2 * tig;
                sts128(st + kp * 32 + ((gid ^ (kp & 7)) << 2), d[0][t][0], d[0][t][2], d[1][t][0], d[1][t][2]);
                sts128(st + (kp + 1) * 32 + ((gid ^ ((kp + 1) & 7)) << 2), d[0][t][1], d[0][t][3], d[1][t][1], d[1][t][3]);
            }
            fence_proxy_async();
            __syncwarp();
            if (!(dbg & 1) && tc::elect_one()) {
                tma_store_2d(&map_g, st, c0, row0);                // k rows x 128 B (box rows = k); PL: k rows x 64 B, hi
                if (PL) tma_store_2d(&map_lo, reinterpret_cast<unsigned char*>(st) + 24 * 64, c0, row0);
                bulk_commit();
            }
        }
    }
    if (tc::elect_one()) bulk_wait_all();
}

template <int KS, int NW = IG_WARPS, int MINB = 2>
__global__ void __launch_bounds__(NW * 32, MINB)
inter_group_bwd_mma_tl_kernel(const __grid_constant__ TmaMap map_dg, int n, int p, int nn, int a, int k, int ci,
                              const float* __restrict__ xyz, const float* __restrict__ sxyz, const int32_t* __restrict__ idx,
                              const float* __restrict__ rk, float inv_sigma, float* __restrict__ gfeats) {
    constexpr int NT = 2 * KS;
    extern __shared__ unsigned char smem_raw[];
    __shared__ float s_g[16 * KS * 3];
    __shared__ uint32_t s_off[16 * KS];
    __shared__ __align__(8) uint64_t s_bar[NW][2];
    const int pi = blockIdx.x, b = blockIdx.y;
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // warp-uniform
    const int gid = lane >> 2, tig = lane & 3;
    float* stage = reinterpret_cast<float*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u)) + warp * (2 * 24 * 32);
    igt_neighbourhood<KS>(b, pi, n, p, nn, a, ci, xyz, sxyz, idx, s_g, s_off);
    for (int i = lane; i < 2 * 24 * 32; i += 32) stage[i] = 0.f;    // rows >= k are never loaded
    if (lane == 0) {
        mbar_init(&s_bar[warp][0], 1);
        mbar_init(&s_bar[warp][1], 1);
        fence_barrier_init();
    }
    if (threadIdx.x == 0) tma_prefetch_desc(&map_dg);
    fence_proxy_async();
    __syncthreads();
    float gx[NT], gy[NT], gz[NT];
    uint32_t off[NT][2];
#pragma unroll
    for (int t = 0; t < NT; ++t) {
        gx[t] = s_g[(8 * t + gid) * 3];
        gy[t] = s_g[(8 * t + gid) * 3 + 1];
        gz[t] = s_g[(8 * t + gid) * 3 + 2];
        off[t][0] = s_off[8 * t + 2 * tig];
        off[t][1] = s_off[8 * t + 2 * tig + 1];
    }
    int kps[6];
#pragma unroll
    for (int j = 0; j < 4; ++j) kps[j] = 2 * tig + (j & 1) + 8 * (j >> 1);
    kps[4] = 16 + 2 * tig;
    kps[5] = 17 + 2 * tig;
    float* gb = gfeats + (size_t)b * n * a * ci + 4 * gid;
    const int chunks = ci / 32;
    const int my_anchors = (a - warp + NW - 1) / NW;
    const int items = my_anchors * chunks;                          // item -> (anchor slot, 32-channel chunk)
    const int dbg = c_dbg_group;
    auto issue = [&](int item) {
        if (!(dbg & 1) && tc::elect_one()) {
            const int ai = warp + (item / chunks) * NW, c0 = (item % chunks) * 32, buf = item & 1;
            mbar_arrive_expect_tx(&s_bar[warp][buf], (uint32_t)k * 128u);
            tma_load_2d(stage + buf * (24 * 32), &map_dg, c0, (((b * p + pi) * a) + ai) * k, &s_bar[warp][buf]);
        }
    };
    if (items > 0) issue(0);
    uint32_t bh[NT][3], bl[NT][3];
    for (int item = 0; item < items; ++item) {
        const int ai = warp + (item / chunks) * NW, c0 = (item % chunks) * 32, buf = item & 1;
        __syncwarp();                                   // every lane is done reading the tile the next load overwrites
        if (item + 1 < items) issue(item + 1);
        if (c0 == 0) {
            float kx[6], ky[6], kz[6];
#pragma unroll
            for (int j = 0; j < 6; ++j) {
                kx[j] = ky[j] = kz[j] = 1e12f;
                if (kps[j] < k) {
                    const float* q = rk + (ai * k + kps[j]) * 3;
                    kx[j] = __ldg(q);
                    ky[j] = __ldg(q + 1);
                    kz[j] = __ldg(q + 2);
                }
            }
#pragma unroll
            for (int t = 0; t < NT; ++t) {
                float w[6];
#pragma unroll
                for (int j = 0; j < 6; ++j) {
                    const float dx = gx[t] - kx[j], dy = gy[t] - ky[j], dz = gz[t] - kz[j];
                    const float v = fmaxf(0.f, 1.f - (dx * dx + dy * dy + dz * dz) * inv_sigma);
                    w[j] = (kps[j] < k && 8 * t + gid < nn) ? v : 0.f;
                }
                ig_split2(w[0], w[1], bh[t][0], bl[t][0]);
                ig_split2(w[2], w[3], bh[t][1], bl[t][1]);
                ig_split2(w[4], w[5], bh[t][2], bl[t][2]);
            }
        }
        if (!(dbg & 1)) mbar_wait(&s_bar[warp][buf], (uint32_t)(item >> 1) & 1u);
        const float* in = stage + buf * (24 * 32);
        float4 v[6];
#pragma unroll
        for (int j = 0; j < 6; ++j) v[j] = lds128(in + kps[j] * 32 + ((gid ^ (kps[j] & 7)) << 2));
        uint32_t ah[2][6], al[2][6];
        ig_split2(v[0].x, v[1].x, ah[0][0], al[0][0]);
        ig_split2(v[0].y, v[1].y, ah[0][1], al[0][1]);
        ig_split2(v[2].x, v[3].x, ah[0][2], al[0][2]);
        ig_split2(v[2].y, v[3].y, ah[0][3], al[0][3]);
        ig_split2(v[4].x, v[5].x, ah[0][4], al[0][4]);
        ig_split2(v[4].y, v[5].y, ah[0][5], al[0][5]);
        ig_split2(v[0].z, v[1].z, ah[1][0], al[1][0]);
        ig_split2(v[0].w, v[1].w, ah[1][1], al[1][1]);
        ig_split2(v[2].z, v[3].z, ah[1][2], al[1][2]);
        ig_split2(v[2].w, v[3].w, ah[1][3], al[1][3]);
        ig_split2(v[4].z, v[5].z, ah[1][4], al[1][4]);
        ig_split2(v[4].w, v[5].w, ah[1][5], al[1][5]);
        float* ga = gb + (size_t)ai * ci + c0;
#pragma unroll
        for (int t = 0; t < NT; ++t) {
            float d[2][4];
#pragma unroll
            for (int m = 0; m < 2; ++m) {
#pragma unroll
                for (int e = 0; e < 4; ++e) d[m][e] = 0.f;
                const uint32_t h16[4] = {ah[m][0], ah[m][1], ah[m][2], ah[m][3]};
                const uint32_t l16[4] = {al[m][0], al[m][1], al[m][2], al[m][3]};
                if (dbg & 4) continue;
                ig_mma(d[m], l16, bh[t][0], bh[t][1]);
                ig_mma(d[m], h16, bl[t][0], bl[t][1]);
                ig_mma(d[m], h16, bh[t][0], bh[t][1]);
                ig_mma_k8(d[m], al[m][4], al[m][5], bh[t][2]);
                ig_mma_k8(d[m], ah[m][4], ah[m][5], bl[t][2]);
                ig_mma_k8(d[m], ah[m][4], ah[m][5], bh[t][2]);
            }
            const int n0 = 8 * t + 2 * tig;
            if (dbg & 2) {
                if (d[0][0] + d[1][1] == 12345.678f) ga[0] = d[0][2];      // keep the MMAs alive
                continue;
            }
            if (n0 < nn) atomicAdd(reinterpret_cast<float4*>(ga + off[t][0]), make_float4(d[0][0], d[0][2], d[1][0], d[1][2]));
            if (n0 + 1 < nn) atomicAdd(reinterpret_cast<float4*>(ga + off[t][1]), make_float4(d[0][1], d[0][3], d[1][1], d[1][3]));
        }
    }
}

// ---------------------------------------------------------------------------------- general warp-MMA variants
// KS = 1..4 k-steps of 16 neighbours (nn <= 64: model 38 builds every layer with 64 neighbours,
// SPConvNets/models/unsup_seg_so3_pose_conv_pn_38_multi_stage.py:2174-2191) and FAST = single-pass bf16 (contraction mode
// 4, BASELINE config 3: operands rounded to bf16 once, one MMA per product instead of three, G leaves as ONE bf16 plane).
// Same fragment scheme as inter_group_fwd_mma_ts_kernel / inter_group_bwd_mma_tl_kernel, written for register economy
// instead of prefetch depth: the weight fragments of the anchor stay in registers (KS x 6 words per plane), the neighbour
// offsets are re-read from shared memory while the weights are computed, and the feature fragments of one k-step live
// only until its MMAs are issued.  The tuned kernels above keep the KS <= 2, fp32-parity shapes of the classic backbone.
template <int KS, bool FAST>
__global__ void __launch_bounds__(IG_WARPS * 32, KS <= 2 ? 2 : 1)
inter_group_fwd_mma_gen_kernel(const __grid_constant__ TmaMap map_hi, const __grid_constant__ TmaMap map_lo, int n, int p, int nn,
                               int a, int k, int ci, const float* __restrict__ xyz, const float* __restrict__ sxyz,
                               const int32_t* __restrict__ idx, const float* __restrict__ rk, float inv_sigma,
                               const float* __restrict__ feats) {
    extern __shared__ unsigned char smem_raw[];
    __shared__ float s_g[16 * KS * 3];
    __shared__ uint32_t s_off[16 * KS];
    const int pi = blockIdx.x, b = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gid = lane >> 2, tig = lane & 3;
    // per warp: two staging buffers, each hi tile [24][64 B] + lo tile [24][64 B]
    unsigned char* stage = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u) + warp * (2 * 24 * 128);
    igt_neighbourhood<KS>(b, pi, n, p, nn, a, ci, xyz, sxyz, idx, s_g, s_off);
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&map_hi);
        if (!FAST) tma_prefetch_desc(&map_lo);
    }
    __syncthreads();
    uint32_t off[KS][4];
#pragma unroll
    for (int s = 0; s < KS; ++s)
#pragma unroll
        for (int j = 0; j < 4; ++j) off[s][j] = s_off[16 * s + 2 * tig + (j & 1) + 8 * (j >> 1)];
    const float* fb = feats + (size_t)b * n * a * ci + 4 * gid;
    int sb = 0;
    for (int ai = warp; ai < a; ai += IG_WARPS) {
        uint32_t bh[KS][3][2], bl[FAST ? 1 : KS][3][2];
#pragma unroll
        for (int t = 0; t < 3; ++t) {
            const int kp = 8 * t + gid;
            const bool live = kp < k;
            float kx = 0.f, ky = 0.f, kz = 0.f;
            if (live) {
                const float* q = rk + (ai * k + kp) * 3;
                kx = __ldg(q);
                ky = __ldg(q + 1);
                kz = __ldg(q + 2);
            }
#pragma unroll
            for (int s = 0; s < KS; ++s) {
                float w[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int ni = 16 * s + 2 * tig + (j & 1) + 8 * (j >> 1);
                    const float dx = s_g[ni * 3] - kx, dy = s_g[ni * 3 + 1] - ky, dz = s_g[ni * 3 + 2] - kz;
                    const float v = fmaxf(0.f, 1.f - (dx * dx + dy * dy + dz * dz) * inv_sigma);
                    w[j] = live ? v : 0.f;
                }
                if (FAST) {
                    bh[s][t][0] = ig_pack_bf16x2(w[0], w[1]);
                    bh[s][t][1] = ig_pack_bf16x2(w[2], w[3]);
                } else {
                    ig_split2(w[0], w[1], bh[s][t][0], bl[s][t][0]);
                    ig_split2(w[2], w[3], bh[s][t][1], bl[s][t][1]);
                }
            }
        }
        const int row0 = (((b * p + pi) * a) + ai) * k;
        const float* fa = fb + (size_t)ai * ci;
        for (int c0 = 0; c0 < ci; c0 += 32, sb ^= 1) {
            float d[2][3][4];
#pragma unroll
            for (int m = 0; m < 2; ++m)
#pragma unroll
                for (int t = 0; t < 3; ++t)
#pragma unroll
                    for (int e = 0; e < 4; ++e) d[m][t][e] = 0.f;
#pragma unroll
            for (int s = 0; s < KS; ++s) {
                float4 v[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) v[j] = __ldg(reinterpret_cast<const float4*>(fa + off[s][j] + c0));
                uint32_t ah[2][4], al[2][4];
                if (FAST) {
                    ah[0][0] = ig_pack_bf16x2(v[0].x, v[1].x);
                    ah[0][1] = ig_pack_bf16x2(v[0].y, v[1].y);
                    ah[0][2] = ig_pack_bf16x2(v[2].x, v[3].x);
                    ah[0][3] = ig_pack_bf16x2(v[2].y, v[3].y);
                    ah[1][0] = ig_pack_bf16x2(v[0].z, v[1].z);
                    ah[1][1] = ig_pack_bf16x2(v[0].w, v[1].w);
                    ah[1][2] = ig_pack_bf16x2(v[2].z, v[3].z);
                    ah[1][3] = ig_pack_bf16x2(v[2].w, v[3].w);
                } else {
                    ig_split2(v[0].x, v[1].x, ah[0][0], al[0][0]);
                    ig_split2(v[0].y, v[1].y, ah[0][1], al[0][1]);
                    ig_split2(v[2].x, v[3].x, ah[0][2], al[0][2]);
                    ig_split2(v[2].y, v[3].y, ah[0][3], al[0][3]);
                    ig_split2(v[0].z, v[1].z, ah[1][0], al[1][0]);
                    ig_split2(v[0].w, v[1].w, ah[1][1], al[1][1]);
                    ig_split2(v[2].z, v[3].z, ah[1][2], al[1][2]);
                    ig_split2(v[2].w, v[3].w, ah[1][3], al[1][3]);
                }
#pragma unroll
                for (int m = 0; m < 2; ++m)
#pragma unroll
                    for (int t = 0; t < 3; ++t) {
                        if (!FAST) {
                            ig_mma(d[m][t], al[m], bh[s][t][0], bh[s][t][1]);
                            ig_mma(d[m][t], ah[m], bl[s][t][0], bl[s][t][1]);
                        }
                        ig_mma(d[m][t], ah[m], bh[s][t][0], bh[s][t][1]);
                    }
            }
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            __syncwarp();
            const uint32_t sh = smem_u32(stage + sb * (24 * 128)), sl = sh + 24 * 64;
            const int flip = tig >> 1;
#pragma unroll
            for (int t = 0; t < 3; ++t) {
                const int kp = 8 * t + 2 * tig;
                uint32_t h[2][2], l[2][2];
                if (FAST) {
                    h[0][0] = ig_pack_bf16x2(d[0][t][0], d[0][t][2]);
                    h[0][1] = ig_pack_bf16x2(d[1][t][0], d[1][t][2]);
                    h[1][0] = ig_pack_bf16x2(d[0][t][1], d[0][t][3]);
                    h[1][1] = ig_pack_bf16x2(d[1][t][1], d[1][t][3]);
                } else {
                    ig_split2(d[0][t][0], d[0][t][2], h[0][0], l[0][0]);
                    ig_split2(d[1][t][0], d[1][t][2], h[0][1], l[0][1]);
                    ig_split2(d[0][t][1], d[0][t][3], h[1][0], l[1][0]);
                    ig_split2(d[1][t][1], d[1][t][3], h[1][1], l[1][1]);
                }
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int par = q ^ flip;
                    const uint32_t o = (uint32_t)((kp + par) * 64 + gid * 8);
                    st_shared_u2(sh + o, par ? h[1][0] : h[0][0], par ? h[1][1] : h[0][1]);
                    if (!FAST) st_shared_u2(sl + o, par ? l[1][0] : l[0][0], par ? l[1][1] : l[0][1]);
                }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
                tma_store_2d(&map_hi, stage + sb * (24 * 128), c0, row0);
                if (!FAST) tma_store_2d(&map_lo, stage + sb * (24 * 128) + 24 * 64, c0, row0);
                bulk_commit();
            }
        }
    }
    if (lane == 0) bulk_wait_all();
}

template <int KS, bool FAST>
__global__ void __launch_bounds__(IG_WARPS * 32, (KS <= 2 || FAST) ? 2 : 1)
inter_group_bwd_mma_gen_kernel(const __grid_constant__ TmaMap map_dg, int n, int p, int nn, int a, int k, int ci,
                               const float* __restrict__ xyz, const float* __restrict__ sxyz, const int32_t* __restrict__ idx,
                               const float* __restrict__ rk, float inv_sigma, float* __restrict__ gfeats) {
    constexpr int NT = 2 * KS;
    extern __shared__ unsigned char smem_raw[];
    __shared__ float s_g[16 * KS * 3];
    __shared__ uint32_t s_off[16 * KS];
    __shared__ __align__(8) uint64_t s_bar[IG_WARPS][2];
    const int pi = blockIdx.x, b = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gid = lane >> 2, tig = lane & 3;
    float* stage = reinterpret_cast<float*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u)) + warp * (2 * 24 * 32);
    igt_neighbourhood<KS>(b, pi, n, p, nn, a, ci, xyz, sxyz, idx, s_g, s_off);
    for (int i = lane; i < 2 * 24 * 32; i += 32) stage[i] = 0.f;    // rows >= k are never loaded
    if (lane == 0) {
        mbar_init(&s_bar[warp][0], 1);
        mbar_init(&s_bar[warp][1], 1);
        fence_barrier_init();
    }
    if (threadIdx.x == 0) tma_prefetch_desc(&map_dg);
    fence_proxy_async();
    __syncthreads();
    uint32_t off[NT][2];
#pragma unroll
    for (int t = 0; t < NT; ++t) {
        off[t][0] = s_off[8 * t + 2 * tig];
        off[t][1] = s_off[8 * t + 2 * tig + 1];
    }
    int kps[6];
#pragma unroll
    for (int j = 0; j < 4; ++j) kps[j] = 2 * tig + (j & 1) + 8 * (j >> 1);
    kps[4] = 16 + 2 * tig;
    kps[5] = 17 + 2 * tig;
    float* gb = gfeats + (size_t)b * n * a * ci + 4 * gid;
    const int chunks = ci / 32;
    const int my_anchors = (a - warp + IG_WARPS - 1) / IG_WARPS;
    const int items = my_anchors * chunks;
    auto issue = [&](int item) {
        if (lane == 0) {
            const int ai = warp + (item / chunks) * IG_WARPS, c0 = (item % chunks) * 32, buf = item & 1;
            mbar_arrive_expect_tx(&s_bar[warp][buf], (uint32_t)k * 128u);
            tma_load_2d(stage + buf * (24 * 32), &map_dg, c0, (((b * p + pi) * a) + ai) * k, &s_bar[warp][buf]);
        }
    };
    if (items > 0) issue(0);
    uint32_t bh[NT][3], bl[FAST ? 1 : NT][3];
    for (int item = 0; item < items; ++item) {
        const int ai = warp + (item / chunks) * IG_WARPS, c0 = (item % chunks) * 32, buf = item & 1;
        __syncwarp();
        if (item + 1 < items) issue(item + 1);
        if (c0 == 0) {
            float kx[6], ky[6], kz[6];
#pragma unroll
            for (int j = 0; j < 6; ++j) {
                kx[j] = ky[j] = kz[j] = 1e12f;
                if (kps[j] < k) {
                    const float* q = rk + (ai * k + kps[j]) * 3;
                    kx[j] = __ldg(q);
                    ky[j] = __ldg(q + 1);
                    kz[j] = __ldg(q + 2);
                }
            }
#pragma unroll
            for (int t = 0; t < NT; ++t) {
                const float gx = s_g[(8 * t + gid) * 3], gy = s_g[(8 * t + gid) * 3 + 1], gz = s_g[(8 * t + gid) * 3 + 2];
                float w[6];
#pragma unroll
                for (int j = 0; j < 6; ++j) {
                    const float dx = gx - kx[j], dy = gy - ky[j], dz = gz - kz[j];
                    const float v = fmaxf(0.f, 1.f - (dx * dx + dy * dy + dz * dz) * inv_sigma);
                    w[j] = (kps[j] < k && 8 * t + gid < nn) ? v : 0.f;
                }
                if (FAST) {
                    bh[t][0] = ig_pack_bf16x2(w[0], w[1]);
                    bh[t][1] = ig_pack_bf16x2(w[2], w[3]);
                    bh[t][2] = ig_pack_bf16x2(w[4], w[5]);
                } else {
                    ig_split2(w[0], w[1], bh[t][0], bl[t][0]);
                    ig_split2(w[2], w[3], bh[t][1], bl[t][1]);
                    ig_split2(w[4], w[5], bh[t][2], bl[t][2]);
                }
            }
        }
        mbar_wait(&s_bar[warp][buf], (uint32_t)(item >> 1) & 1u);
        const float* in = stage + buf * (24 * 32);
        float4 v[6];
#pragma unroll
        for (int j = 0; j < 6; ++j) v[j] = lds128(in + kps[j] * 32 + ((gid ^ (kps[j] & 7)) << 2));
        uint32_t ah[2][6], al[2][6];
        if (FAST) {
            ah[0][0] = ig_pack_bf16x2(v[0].x, v[1].x);
            ah[0][1] = ig_pack_bf16x2(v[0].y, v[1].y);
            ah[0][2] = ig_pack_bf16x2(v[2].x, v[3].x);
            ah[0][3] = ig_pack_bf16x2(v[2].y, v[3].y);
            ah[0][4] = ig_pack_bf16x2(v[4].x, v[5].x);
            ah[0][5] = ig_pack_bf16x2(v[4].y, v[5].y);
            ah[1][0] = ig_pack_bf16x2(v[0].z, v[1].z);
            ah[1][1] = ig_pack_bf16x2(v[0].w, v[1].w);
            ah[1][2] = ig_pack_bf16x2(v[2].z, v[3].z);
            ah[1][3] = ig_pack_bf16x2(v[2].w, v[3].w);
            ah[1][4] = ig_pack_bf16x2(v[4].z, v[5].z);
            ah[1][5] = ig_pack_bf16x2(v[4].w, v[5].w);
        } else {
            ig_split2(v[0].x, v[1].x, ah[0][0], al[0][0]);
            ig_split2(v[0].y, v[1].y, ah[0][1], al[0][1]);
            ig_split2(v[2].x, v[3].x, ah[0][2], al[0][2]);
            ig_split2(v[2].y, v[3].y, ah[0][3], al[0][3]);
            ig_split2(v[4].x, v[5].x, ah[0][4], al[0][4]);
            ig_split2(v[4].y, v[5].y, ah[0][5], al[0][5]);
            ig_split2(v[0].z, v[1].z, ah[1][0], al[1][0]);
            ig_split2(v[0].w, v[1].w, ah[1][1], al[1][1]);
            ig_split2(v[2].z, v[3].z, ah[1][2], al[1][2]);
            ig_split2(v[2].w, v[3].w, ah[1][3], al[1][3]);
            ig_split2(v[4].z, v[5].z, ah[1][4], al[1][4]);
            ig_split2(v[4].w, v[5].w, ah[1][5], al[1][5]);
        }
        float* ga = gb + (size_t)ai * ci + c0;
#pragma unroll
        for (int t = 0; t < NT; ++t) {
            float d[2][4];
#pragma unroll
            for (int m = 0; m < 2; ++m) {
#pragma unroll
                for (int e = 0; e < 4; ++e) d[m][e] = 0.f;
                const uint32_t h16[4] = {ah[m][0], ah[m][1], ah[m][2], ah[m][3]};
                if (!FAST) {
                    const uint32_t l16[4] = {al[m][0], al[m][1], al[m][2], al[m][3]};
                    ig_mma(d[m], l16, bh[t][0], bh[t][1]);
                    ig_mma(d[m], h16, bl[t][0], bl[t][1]);
                    ig_mma_k8(d[m], al[m][4], al[m][5], bh[t][2]);
                    ig_mma_k8(d[m], ah[m][4], ah[m][5], bl[t][2]);
                }
                ig_mma(d[m], h16, bh[t][0], bh[t][1]);
                ig_mma_k8(d[m], ah[m][4], ah[m][5], bh[t][2]);
            }
            const int n0 = 8 * t + 2 * tig;
            if (n0 < nn) atomicAdd(reinterpret_cast<float4*>(ga + off[t][0]), make_float4(d[0][0], d[0][2], d[1][0], d[1][2]));
            if (n0 + 1 < nn) atomicAdd(reinterpret_cast<float4*>(ga + off[t][1]), make_float4(d[0][1], d[0][3], d[1][1], d[1][3]));
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
        float kx = 0.f, ky = 0.f, kz = 0.f;
        if (lane < k) {
            const float* kp = rk + (ai * k + lane) * 3;
            kx = __ldg(kp), ky = __ldg(kp + 1), kz = __ldg(kp + 2);
        }
        float acc = 0.f;
        for (int base = 0; base < nn; base += 32) {
            // the 32 neighbour features of this anchor in ONE load instruction (lane <-> neighbour), handed to the kernel-
            // point lanes by shuffles: the first version issued a dependent global load per (kernel point, neighbour)
            const float f = base + lane < nn ? __ldg(fb + (size_t)s_j[base + lane] * a + ai) : 0.f;
            const int cnt = nn - base < 32 ? nn - base : 32;
#pragma unroll 4
            for (int i = 0; i < cnt; ++i) {
                const int ni = base + i;
                const float dx = s_g[ni * 3] - kx, dy = s_g[ni * 3 + 1] - ky, dz = s_g[ni * 3 + 2] - kz;
                const float w = fmaxf(0.f, 1.f - (dx * dx + dy * dy + dz * dz) * inv_sigma);
                acc = fmaf(w, __shfl_sync(0xffffffffu, f, i), acc);
            }
        }
        if (lane < k) grouped[(((size_t)b * p + pi) * a + ai) * (size_t)k + lane] = acc;
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

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

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

namespace vgtkb {
template <int KS, bool FAST>
static int launch_group_fwd_gen(const TmaMap& mh, const TmaMap& ml, int b, int n, int p, int nn, int a, int k, int ci, const float* xyz,
                                const float* sxyz, const int32_t* idx, const float* rk, float sigma, const float* feats,
                                cudaStream_t st) {
    const size_t smem = (size_t)IG_WARPS * 2 * 24 * 128 + 1024;
    auto kern = inter_group_fwd_mma_gen_kernel<KS, FAST>;
    VGTKB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<dim3(p, b), IG_WARPS * 32, smem, st>>>(mh, ml, n, p, nn, a, k, ci, xyz, sxyz, idx, rk, 1.0f / sigma, feats);
    return check_launch("inter_group_forward(mma, general)");
}

// G as bf16 planes [b*p*a, k*ci]: hi and lo (fast = 0, the bf16x3 operand format) or hi only (fast = 1, single-pass bf16;
// g_lo unused) -- the forward grouping of vgtkb_inter_conv_forward.  VGTKB_EUNSUP for shapes the warp-MMA kernels do not
// take (k > 24, nn > 64, ci % 32 != 0, misaligned / too large tensors).
static void set_dbg_group() {
    static bool done = false;
    if (done) return;
    done = true;
    const int v = getenv("VGTKB_DBG_GROUP") ? atoi(getenv("VGTKB_DBG_GROUP")) : 0;
    cudaMemcpyToSymbol(c_dbg_group, &v, sizeof(int));
}

int inter_group_forward_planes(int b, int n, int p, int nn, int a, int k, int ci, const float* xyz, const float* sample_xyz,
                               const int32_t* idx, const float* rot_kernels, float sigma, const float* feats, void* g_hi,
                               void* g_lo, int fast, cudaStream_t st) {
    set_dbg_group();
    const int64_t g_rows = (int64_t)b * p * a * k;
    if (k > 24 || nn > 64 || ci % 32 != 0 || !aligned16(feats) || !aligned16(g_hi) || (!fast && !aligned16(g_lo)) ||
        (int64_t)n * a * ci >= ((int64_t)1 << 32) || g_rows >= ((int64_t)1 << 31) || b > 65535)
        return VGTKB_EUNSUP;
    TmaMap mh, ml;
    int rc = make_plane_map(&mh, g_hi, g_rows, ci, k);
    if (rc) return rc;
    ml = mh;
    if (!fast) {
        rc = make_plane_map(&ml, g_lo, g_rows, ci, k);
        if (rc) return rc;
    }
    const int ks = (nn + 15) / 16;
    if (!fast && ks <= 2) {       // the tuned kernels (prefetch, two CTAs per SM)
        // warps per CTA / CTAs per SM of the tuned kernels: VGTKB_IG_FWD1 / _FWD2 pick another compiled geometry (A/B runs)
#define VGTKB_FWD_PL(KS_, NW_, MB_)                                                                                              \
    do {                                                                                                                         \
        const size_t smem = (size_t)(NW_) * 2 * 24 * 128 + 1024;                                                                 \
        auto kern = inter_group_fwd_mma_ts_kernel<KS_, true, NW_, MB_>;                                                           \
        VGTKB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                          \
        kern<<<dim3(p, b), (NW_) * 32, smem, st>>>(mh, n, p, nn, a, k, ci, xyz, sample_xyz, idx, rot_kernels, 1.0f / sigma, feats, ml); \
    } while (0)
        static const int g1 = getenv("VGTKB_IG_FWD1") ? atoi(getenv("VGTKB_IG_FWD1")) : 0;
        static const int g2 = getenv("VGTKB_IG_FWD2") ? atoi(getenv("VGTKB_IG_FWD2")) : 0;
        // measured on bench.py (same box, per-entry-point CUDA events; profiles/r2_grouping_geometry_ab.txt): 10 warps x 2 CTAs
        // per SM (20 warps, 6 anchors per warp: no ragged last round as with 8 warps on 60 anchors) beats 8 x 2, 6 x 3, 5 x 3
        // and 4 x 4 for the 16-neighbour kernels (forward 2.12 -> 2.03 ms per step, backward 2.92 -> 2.87) and for the
        // 32-neighbour backward (1.90 -> 1.87); the 32-neighbour forward needs 128 registers and stays at 8 x 2
        if (ks == 1) {
            if (g1 == 63) VGTKB_FWD_PL(1, 6, 3);
            else if (g1 == 44) VGTKB_FWD_PL(1, 4, 4);
            else if (g1 == 82) VGTKB_FWD_PL(1, IG_WARPS, 2);
            else if (g1 == 53) VGTKB_FWD_PL(1, 5, 3);
            else VGTKB_FWD_PL(1, 10, 2);
        } else {
            if (g2 == 62) VGTKB_FWD_PL(2, 6, 2);
            else if (g2 == 53) VGTKB_FWD_PL(2, 5, 3);
            else if (g2 == 43) VGTKB_FWD_PL(2, 4, 3);
            else if (g2 == 44) VGTKB_FWD_PL(2, 4, 4);
            else VGTKB_FWD_PL(2, IG_WARPS, 2);
        }
#undef VGTKB_FWD_PL
        return check_launch("inter_group_forward(mma, bf16 planes)");
    }
#define VGTKB_GEN_FWD(KS_, F_) \
    return launch_group_fwd_gen<KS_, F_>(mh, ml, b, n, p, nn, a, k, ci, xyz, sample_xyz, idx, rot_kernels, sigma, feats, st)
    if (fast) {
        if (ks == 1) VGTKB_GEN_FWD(1, true);
        if (ks == 2) VGTKB_GEN_FWD(2, true);
        if (ks == 3) VGTKB_GEN_FWD(3, true);
        VGTKB_GEN_FWD(4, true);
    }
    if (ks == 3) VGTKB_GEN_FWD(3, false);
    VGTKB_GEN_FWD(4, false);
#undef VGTKB_GEN_FWD
}

template <int KS, bool FAST>
static int launch_group_bwd_gen(const TmaMap& map, int b, int n, int p, int nn, int a, int k, int ci, const float* xyz,
                                const float* sxyz, const int32_t* idx, const float* rk, float sigma, float* gfeats, cudaStream_t st) {
    const size_t smem = (size_t)IG_WARPS * 2 * 24 * 128 + 1024;
    auto kern = inter_group_bwd_mma_gen_kernel<KS, FAST>;
    VGTKB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<dim3(p, b), IG_WARPS * 32, smem, st>>>(map, n, p, nn, a, k, ci, xyz, sxyz, idx, rk, 1.0f / sigma, gfeats);
    return check_launch("inter_group_backward(mma, general)");
}

// scatter of the fp32 dG through the neighbourhoods for the shapes the tuned kernel does not take (33..64 neighbours) and for
// the single-pass bf16 mode; grad_feats is accumulated into.  VGTKB_EUNSUP: caller falls back to vgtkb_inter_group_backward.
int inter_group_backward_gen(int b, int n, int p, int nn, int a, int k, int ci, const float* xyz, const float* sample_xyz,
                             const int32_t* idx, const float* rot_kernels, float sigma, const float* grad_grouped,
                             float* grad_feats, int fast, cudaStream_t st) {
    set_dbg_group();
    const int64_t g_rows = (int64_t)b * p * a * k;
    if (k > 24 || nn > 64 || ci % 32 != 0 || !aligned16(grad_feats) || !aligned16(grad_grouped) ||
        (int64_t)n * a * ci >= ((int64_t)1 << 32) || g_rows >= ((int64_t)1 << 31) || b > 65535)
        return VGTKB_EUNSUP;
    TmaMap map;
    const int rc = make_rows_map(&map, grad_grouped, g_rows, ci, k);
    if (rc) return rc;
    const int ks = (nn + 15) / 16;
#define VGTKB_GEN_BWD(KS_, F_) \
    return launch_group_bwd_gen<KS_, F_>(map, b, n, p, nn, a, k, ci, xyz, sample_xyz, idx, rot_kernels, sigma, grad_feats, st)
    if (fast) {
        if (ks == 1) VGTKB_GEN_BWD(1, true);
        if (ks == 2) VGTKB_GEN_BWD(2, true);
        if (ks == 3) VGTKB_GEN_BWD(3, true);
        VGTKB_GEN_BWD(4, true);
    }
    if (ks <= 2) return VGTKB_EUNSUP;      // the tuned kernels behind vgtkb_inter_group_backward take these
    if (ks == 3) VGTKB_GEN_BWD(3, false);
    VGTKB_GEN_BWD(4, false);
#undef VGTKB_GEN_BWD
}
}  // namespace vgtkb

extern "C" int vgtkb_inter_group_forward(int b, int n, int p, int nn, int a, int k, int ci, const float* xyz,
                                         const float* sample_xyz, const int32_t* idx, const float* rot_kernels,
                                         float sigma, const float* feats, float* grouped, int mode, void* stream) {
    int rc = check_inter_args(b, n, p, nn, a, k, ci);
    if (rc) return rc;
    if (b == 0 || p == 0) return VGTKB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const bool al = aligned16(feats) && aligned16(grouped);
    // VGTKB_GROUP_TMA: 2 (default) = register gathers + tensor-map store of G, 1 = all-bulk variant, 0 = registers only
    static const int use_tma = getenv("VGTKB_GROUP_TMA") ? atoi(getenv("VGTKB_GROUP_TMA")) : 2;
    const int64_t g_rows = (int64_t)b * p * a * k;
    if (use_tma == 2 && mode == 3 && k <= 24 && nn <= 32 && ci % 32 == 0 && al && (int64_t)n * a * ci < ((int64_t)1 << 32) &&
        g_rows < ((int64_t)1 << 31)) {
        TmaMap map;
        if (make_rows_map(&map, grouped, g_rows, ci, k) == VGTKB_OK) {
            const size_t smem = (size_t)IG_WARPS * 2 * 24 * 128 + 1024;
            if (nn <= 16) {
                VGTKB_CUDA(cudaFuncSetAttribute(inter_group_fwd_mma_ts_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                inter_group_fwd_mma_ts_kernel<1><<<dim3(p, b), IG_WARPS * 32, smem, st>>>(map, n, p, nn, a, k, ci, xyz, sample_xyz, idx,
                                                                                          rot_kernels, 1.0f / sigma, feats, map);
            } else {
                VGTKB_CUDA(cudaFuncSetAttribute(inter_group_fwd_mma_ts_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                inter_group_fwd_mma_ts_kernel<2><<<dim3(p, b), IG_WARPS * 32, smem, st>>>(map, n, p, nn, a, k, ci, xyz, sample_xyz, idx,
                                                                                          rot_kernels, 1.0f / sigma, feats, map);
            }
            return check_launch("inter_group_forward(mma, tensor-map store)");
        }
    }
    if (use_tma == 1 && mode == 3 && k <= 24 && nn <= 32 && ci % IGT_CB == 0 && al && (int64_t)n * a * ci < ((int64_t)1 << 32)) {
        const int ks = nn <= 16 ? 1 : 2;
        const size_t smem = (size_t)IGT_WARPS * (2 * 16 * ks + 24) * IGT_LD * sizeof(float);
        if (ks == 1) {
            VGTKB_CUDA(cudaFuncSetAttribute(inter_group_fwd_tma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            inter_group_fwd_tma_kernel<1><<<dim3(p, b), IGT_WARPS * 32, smem, st>>>(n, p, nn, a, k, ci, xyz, sample_xyz, idx, rot_kernels,
                                                                                    1.0f / sigma, feats, grouped);
        } else {
            VGTKB_CUDA(cudaFuncSetAttribute(inter_group_fwd_tma_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            inter_group_fwd_tma_kernel<2><<<dim3(p, b), IGT_WARPS * 32, smem, st>>>(n, p, nn, a, k, ci, xyz, sample_xyz, idx, rot_kernels,
                                                                                    1.0f / sigma, feats, grouped);
        }
        return check_launch("inter_group_forward(mma+tma)");
    }
    if (mode == 3 && k <= 24 && nn <= 32 && ci % 32 == 0 && al && (int64_t)n * a * ci < ((int64_t)1 << 32)) {
        if (nn <= 16)
            inter_group_fwd_mma_kernel<1><<<dim3(p, b), IG_WARPS * 32, 0, st>>>(n, p, nn, a, k, ci, xyz, sample_xyz, idx, rot_kernels,
                                                                                1.0f / sigma, feats, grouped);
        else
            inter_group_fwd_mma_kernel<2><<<dim3(p, b), IG_WARPS * 32, 0, st>>>(n, p, nn, a, k, ci, xyz, sample_xyz, idx, rot_kernels,
                                                                                1.0f / sigma, feats, grouped);
        return check_launch("inter_group_forward(mma)");
    }
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
                                          float sigma, const float* grad_grouped, float* grad_feats, int mode, void* stream) {
    int rc = check_inter_args(b, n, p, nn, a, k, ci);
    if (rc) return rc;
    if (b == 0 || p == 0) return VGTKB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const bool al = aligned16(grad_feats) && aligned16(grad_grouped);
    static const int use_tma = getenv("VGTKB_GROUP_TMA") ? atoi(getenv("VGTKB_GROUP_TMA")) : 2;
    const int64_t g_rows = (int64_t)b * p * a * k;
    if (use_tma == 2 && mode == 3 && k <= 24 && nn <= 32 && ci % 32 == 0 && al && (int64_t)n * a * ci < ((int64_t)1 << 32) &&
        g_rows < ((int64_t)1 << 31)) {
        TmaMap map;
        if (make_rows_map(&map, grad_grouped, g_rows, ci, k) == VGTKB_OK) {
#define VGTKB_BWD_TL(KS_, NW_, MB_)                                                                                              \
    do {                                                                                                                         \
        const size_t smem = (size_t)(NW_) * 2 * 24 * 128 + 1024;                                                                 \
        auto kern = inter_group_bwd_mma_tl_kernel<KS_, NW_, MB_>;                                                                 \
        VGTKB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                          \
        kern<<<dim3(p, b), (NW_) * 32, smem, st>>>(map, n, p, nn, a, k, ci, xyz, sample_xyz, idx, rot_kernels, 1.0f / sigma, grad_feats); \
    } while (0)
            static const int g1 = getenv("VGTKB_IG_BWD1") ? atoi(getenv("VGTKB_IG_BWD1")) : 0;
            static const int g2 = getenv("VGTKB_IG_BWD2") ? atoi(getenv("VGTKB_IG_BWD2")) : 0;
            if (nn <= 16) {
                if (g1 == 63) VGTKB_BWD_TL(1, 6, 3);
                else if (g1 == 44) VGTKB_BWD_TL(1, 4, 4);
                else if (g1 == 82) VGTKB_BWD_TL(1, IG_WARPS, 2);
                else if (g1 == 53) VGTKB_BWD_TL(1, 5, 3);
                else VGTKB_BWD_TL(1, 10, 2);
            } else {
                if (g2 == 63) VGTKB_BWD_TL(2, 6, 3);
                else if (g2 == 53) VGTKB_BWD_TL(2, 5, 3);
                else if (g2 == 44) VGTKB_BWD_TL(2, 4, 4);
                else if (g2 == 82) VGTKB_BWD_TL(2, IG_WARPS, 2);
                else VGTKB_BWD_TL(2, 10, 2);
            }
#undef VGTKB_BWD_TL
            return check_launch("inter_group_backward(mma, tensor-map load)");
        }
    }
    if (use_tma == 1 && mode == 3 && k <= 24 && nn <= 32 && ci % IGT_CB == 0 && al && (int64_t)n * a * ci < ((int64_t)1 << 32)) {
        const int ks = nn <= 16 ? 1 : 2;
        const size_t smem = (size_t)IGT_WARPS * (2 * 24 + 16 * ks) * IGT_LD * sizeof(float);
        if (ks == 1) {
            VGTKB_CUDA(cudaFuncSetAttribute(inter_group_bwd_tma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            inter_group_bwd_tma_kernel<1><<<dim3(p, b), IGT_WARPS * 32, smem, st>>>(n, p, nn, a, k, ci, xyz, sample_xyz, idx, rot_kernels,
                                                                                    1.0f / sigma, grad_grouped, grad_feats);
        } else {
            VGTKB_CUDA(cudaFuncSetAttribute(inter_group_bwd_tma_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            inter_group_bwd_tma_kernel<2><<<dim3(p, b), IGT_WARPS * 32, smem, st>>>(n, p, nn, a, k, ci, xyz, sample_xyz, idx, rot_kernels,
                                                                                    1.0f / sigma, grad_grouped, grad_feats);
        }
        return check_launch("inter_group_backward(mma+tma)");
    }
    if (mode == 3 && k <= 24 && nn <= 32 && ci % 32 == 0 && al && (int64_t)n * a * ci < ((int64_t)1 << 32)) {
        if (nn <= 16)
            inter_group_bwd_mma_kernel<1><<<dim3(p, b), IG_WARPS * 32, 0, st>>>(n, p, nn, a, k, ci, xyz, sample_xyz, idx, rot_kernels,
                                                                                1.0f / sigma, grad_grouped, grad_feats);
        else
            inter_group_bwd_mma_kernel<2><<<dim3(p, b), IG_WARPS * 32, 0, st>>>(n, p, nn, a, k, ci, xyz, sample_xyz, idx, rot_kernels,
                                                                                1.0f / sigma, grad_grouped, grad_feats);
        return check_launch("inter_group_backward(mma)");
    }
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
