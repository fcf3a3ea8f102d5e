// pointnet2.cu -- PointNet++ set abstraction / feature propagation for sm_100a
// (SPConvNets/models/PointNet2.py:78-196, the PointnetPP encoder-decoder of the reference).
//
// The reference materialises the [B,S,N] distance matrix, runs torch.topk(k=64), three batched_index_select
// copies and a torch.cat per level, and masks the MLP output in place before the max.  Here:
//   knn_kernel             one CTA per centre: squared distances of the whole cloud go straight into a shared-memory
//                          array of 64-bit (distance bits, index) keys that is bitonic-sorted in place; only the k
//                          winners ever reach HBM (sample_and_group :85-87)
//   sa_group_kernel        gather + centre subtraction + concat written ONCE as the K-major operand rows of the MLP's
//                          first contraction (rows padded to the tensor-core granule) (:92-100)
//   sa_maxpool_*           the radius mask folded into the max over the k neighbours, arg-max kept for backward
//                          (max_pooling_with_r :102-112)
//   three_nn / three_interpolate_*   feature propagation (interpolate_features :114-129)
// The MLPs themselves run on the tcgen05 contraction (vgtkb_gemm_nt/tn) and the fused norm kernels.
// Arithmetic that decides indices follows torch's CPU evaluation order (checked against fixtures of the reference's own
// module): torch.sum((a-b)**2,-1) = (dx*dx + dy*dy) + dz*dz without contraction; torch.norm = fma(dz,dz,fma(dy,dy,dx*dx)).
#include "common.cuh"

namespace vgtkb {

constexpr int KNN_THREADS = 256;

// ---- k nearest neighbours ---------------------------------------------------------------------------------------
// grid (S, B).  keys[np2]: (bits(d2) << 32) | index; d2 >= 0 so the unsigned order of the bits is the order of the
// values, and ties resolve to the smaller index.
__global__ void __launch_bounds__(KNN_THREADS)
knn_kernel(int n, int s, int k, int np2, const float* __restrict__ pos, const float* __restrict__ centers,
           int32_t* __restrict__ idx, float* __restrict__ dist) {
    extern __shared__ __align__(16) unsigned long long knn_keys[];
    const int sc = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
    const float* c = centers + ((size_t)b * s + sc) * 3;
    const float cx = c[0], cy = c[1], cz = c[2];
    const float* p = pos + (size_t)b * n * 3;
    for (int i = tid; i < np2; i += KNN_THREADS) {
        unsigned long long key = ~0ull;
        if (i < n) {
            const float dx = cx - p[3 * i], dy = cy - p[3 * i + 1], dz = cz - p[3 * i + 2];
            const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
            key = ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned)i;
        }
        knn_keys[i] = key;
    }
    for (int size = 2; size <= np2; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            for (int t = tid; t < (np2 >> 1); t += KNN_THREADS) {
                const int lo = 2 * t - (t & (stride - 1));
                const int hi = lo + stride;
                const unsigned long long a = knn_keys[lo], bkey = knn_keys[hi];
                const bool up = (lo & size) == 0;
                if ((a > bkey) == up) {
                    knn_keys[lo] = bkey;
                    knn_keys[hi] = a;
                }
            }
        }
    }
    __syncthreads();
    for (int j = tid; j < k; j += KNN_THREADS) {
        const unsigned long long key = knn_keys[j];
        const size_t o = ((size_t)b * s + sc) * k + j;
        idx[o] = (int32_t)(key & 0xffffffffu);
        dist[o] = __fsqrt_rn(__uint_as_float((unsigned)(key >> 32)));
    }
}

// ---- grouping: out[b,s,j,:] = [pos[b,idx]-center (3) | feat[b,idx,:] (c) | 0 (pad)] -------------------------------
// one warp per row, grid-stride; idx == nullptr: identity neighbourhood (j = point), centers == nullptr: origin
__global__ void __launch_bounds__(256)
sa_group_kernel(int64_t rows, int n, int s, int k, int c, int cpad, const float* __restrict__ pos,
                const float* __restrict__ feat, const float* __restrict__ centers, const int32_t* __restrict__ idx,
                float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp0; r < rows; r += nwarps) {
        const int64_t bs = r / k;
        const int j = (int)(r - bs * k);
        const int64_t b = bs / s;
        const int pt = idx ? idx[r] : j;
        const float* pp = pos + (b * n + pt) * 3;
        const float* ff = feat ? feat + (b * n + pt) * (int64_t)c : nullptr;
        float* o = out + r * cpad;
        for (int ch = lane; ch < cpad; ch += 32) {
            float v = 0.f;
            if (ch < 3) v = centers ? __fsub_rn(pp[ch], centers[bs * 3 + ch]) : pp[ch];
            else if (ch < 3 + c) v = ff[ch - 3];
            o[ch] = v;
        }
    }
}

// grad_feat[b, idx, :] += grad_out[row, 3:3+c]   (grad_feat zero-filled by the entry point)
__global__ void __launch_bounds__(256)
sa_group_bwd_kernel(int64_t rows, int n, int s, int k, int c, int cpad, const float* __restrict__ gout,
                    const int32_t* __restrict__ idx, float* __restrict__ gfeat) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp0; r < rows; r += nwarps) {
        const int64_t bs = r / k;
        const int j = (int)(r - bs * k);
        const int64_t b = bs / s;
        const int pt = idx ? idx[r] : j;
        float* g = gfeat + (b * n + pt) * (int64_t)c;
        const float* go = gout + r * cpad + 3;
        for (int ch = lane; ch < c; ch += 32) atomicAdd(g + ch, go[ch]);
    }
}

// ---- masked max over the k neighbours -----------------------------------------------------------------------------
// grid (groups, ceil(c/128)): thread <-> channel
__global__ void __launch_bounds__(128)
sa_maxpool_kernel(int k, int c, const float* __restrict__ y, const float* __restrict__ dist, float r,
                  float* __restrict__ out, int32_t* __restrict__ arg) {
    const int ch = blockIdx.y * 128 + threadIdx.x;
    const int64_t g = blockIdx.x;
    if (ch >= c) return;
    const float* row = y + g * (int64_t)k * c + ch;
    const float* d = dist ? dist + g * (int64_t)k : nullptr;
    float best = 0.f;
    int besti = 0;
    constexpr int U = 8;
    for (int j0 = 0; j0 < k; j0 += U) {
        float v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = j0 + u < k ? row[(int64_t)(j0 + u) * c] : 0.f;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int j = j0 + u;
            if (j < k) {
                const float t = (d && !(d[j] <= r)) ? -1e8f : v[u];
                const bool take = j == 0 || t > best;        // first maximum, like torch.max on the host
                best = take ? t : best;
                besti = take ? j : besti;
            }
        }
    }
    out[g * c + ch] = best;
    arg[g * c + ch] = besti;
}

// grad_y[g, j, ch] = (arg[g,ch] == j) ? grad_out[g,ch] : 0     -- one write pass, no memset + scatter
__global__ void __launch_bounds__(256)
sa_maxpool_bwd_kernel(int64_t total, int k, int c, const float* __restrict__ gout, const int32_t* __restrict__ arg,
                      float* __restrict__ gy) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int ch = (int)(t % c);
        const int64_t gj = t / c;
        const int j = (int)(gj % k);
        const int64_t g = gj / k;
        gy[t] = arg[g * c + ch] == j ? gout[g * c + ch] : 0.f;
    }
}

// ---- three nearest neighbours + inverse-distance weights ----------------------------------------------------------
// grid (ceil(n2/128), B); the source cloud p1 streams through shared memory in tiles
constexpr int NN3_TILE = 1024;
__global__ void __launch_bounds__(128)
three_nn_kernel(int n1, int n2, const float* __restrict__ p1, const float* __restrict__ p2, int32_t* __restrict__ idx,
                float* __restrict__ w) {
    __shared__ float tile[NN3_TILE * 3];
    const int b = blockIdx.y, q = blockIdx.x * 128 + threadIdx.x;
    const bool live = q < n2;
    float qx = 0.f, qy = 0.f, qz = 0.f;
    if (live) {
        const float* qq = p2 + ((size_t)b * n2 + q) * 3;
        qx = qq[0], qy = qq[1], qz = qq[2];
    }
    float d0 = INFINITY, d1 = INFINITY, d2 = INFINITY;
    int i0 = 0, i1 = 0, i2 = 0;
    for (int base = 0; base < n1; base += NN3_TILE) {
        const int cnt = min(NN3_TILE, n1 - base);
        __syncthreads();
        for (int t = threadIdx.x; t < cnt * 3; t += 128) tile[t] = p1[((size_t)b * n1 + base) * 3 + t];
        __syncthreads();
        if (live) {
            for (int i = 0; i < cnt; ++i) {
                const float d = __fsqrt_rn(sq3(qx - tile[3 * i], qy - tile[3 * i + 1], qz - tile[3 * i + 2]));
                if (d < d2) {
                    if (d < d1) {
                        d2 = d1, i2 = i1;
                        if (d < d0) d1 = d0, i1 = i0, d0 = d, i0 = base + i;
                        else d1 = d, i1 = base + i;
                    } else d2 = d, i2 = base + i;
                }
            }
        }
    }
    if (!live) return;
    const int kk = n1 < 3 ? n1 : 3;
    const float r0 = __fdiv_rn(1.f, __fadd_rn(d0, 1e-8f));
    const float r1 = kk > 1 ? __fdiv_rn(1.f, __fadd_rn(d1, 1e-8f)) : 0.f;
    const float r2 = kk > 2 ? __fdiv_rn(1.f, __fadd_rn(d2, 1e-8f)) : 0.f;
    const float norm = __fadd_rn(__fadd_rn(r0, r1), r2);
    const size_t o = ((size_t)b * n2 + q) * 3;
    idx[o] = i0, idx[o + 1] = kk > 1 ? i1 : 0, idx[o + 2] = kk > 2 ? i2 : 0;
    w[o] = __fdiv_rn(r0, norm), w[o + 1] = __fdiv_rn(r1, norm), w[o + 2] = __fdiv_rn(r2, norm);
}

// out[b,q,:] = (f[i0]*w0 + f[i1]*w1) + f[i2]*w2 (products rounded, like the reference's mul + sum); warp per query
__global__ void __launch_bounds__(256)
three_interpolate_kernel(int64_t rows, int n1, int n2, int c, const float* __restrict__ feat, const int32_t* __restrict__ idx,
                         const float* __restrict__ w, float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp0; r < rows; r += nwarps) {
        const int64_t b = r / n2;
        const float* f = feat + b * n1 * (int64_t)c;
        const float* f0 = f + idx[r * 3] * (int64_t)c, *f1 = f + idx[r * 3 + 1] * (int64_t)c, *f2 = f + idx[r * 3 + 2] * (int64_t)c;
        const float w0 = w[r * 3], w1 = w[r * 3 + 1], w2 = w[r * 3 + 2];
        for (int ch = lane; ch < c; ch += 32)
            out[r * c + ch] = __fadd_rn(__fadd_rn(__fmul_rn(f0[ch], w0), __fmul_rn(f1[ch], w1)), __fmul_rn(f2[ch], w2));
    }
}

__global__ void __launch_bounds__(256)
three_interpolate_bwd_kernel(int64_t rows, int n1, int n2, int c, const float* __restrict__ gout, const int32_t* __restrict__ idx,
                             const float* __restrict__ w, float* __restrict__ gfeat) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp0; r < rows; r += nwarps) {
        const int64_t b = r / n2;
        float* g = gfeat + b * n1 * (int64_t)c;
        for (int j = 0; j < 3; ++j) {
            const float wj = w[r * 3 + j];
            if (wj == 0.f) continue;
            float* gj = g + idx[r * 3 + j] * (int64_t)c;
            for (int ch = lane; ch < c; ch += 32) atomicAdd(gj + ch, gout[r * c + ch] * wj);
        }
    }
}

static inline int rows_grid(int64_t rows) {   // warp per row, 8 warps per CTA, a few CTAs per SM
    const int64_t want = ceil_div64(rows, 8);
    const int64_t cap = (int64_t)kNumSMs * 16;
    return (int)(want < cap ? (want < 1 ? 1 : want) : cap);
}

}  // namespace vgtkb

using namespace vgtkb;

extern "C" int vgtkb_knn_query(int b, int n, int s, int k, const float* pos, const float* centers, int32_t* idx, float* dist,
                               void* stream) {
    VGTKB_REQUIRE(b >= 0 && n >= 1 && s >= 0 && k >= 1, "knn_query: bad size");
    VGTKB_REQUIRE(k <= n, "knn_query: k = %d neighbours of a cloud of %d points", k, n);
    VGTKB_REQUIRE(n <= 16384, "knn_query: clouds above 16384 points are not supported (n = %d)", n);
    VGTKB_REQUIRE(b <= 65535, "knn_query: batch > 65535");
    if (b == 0 || s == 0) return VGTKB_OK;
    int np2 = 2;
    while (np2 < n) np2 <<= 1;
    const size_t smem = (size_t)np2 * sizeof(unsigned long long);
    if (smem > 48 * 1024) VGTKB_CUDA(cudaFuncSetAttribute(knn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    knn_kernel<<<dim3(s, b), KNN_THREADS, smem, (cudaStream_t)stream>>>(n, s, k, np2, pos, centers, idx, dist);
    return check_launch("knn_query");
}

extern "C" int vgtkb_sa_group_forward(int b, int n, int s, int k, int c, int cpad, const float* pos, const float* feat,
                                      const float* centers, const int32_t* idx, float* out, void* stream) {
    VGTKB_REQUIRE(b >= 0 && n >= 1 && s >= 0 && k >= 1 && c >= 0 && cpad >= 3 + c, "sa_group: bad size");
    VGTKB_REQUIRE(c == 0 || feat != nullptr, "sa_group: c > 0 without features");
    VGTKB_REQUIRE(idx != nullptr || k == n, "sa_group: identity neighbourhood needs k == n");
    const int64_t rows = (int64_t)b * s * k;
    if (rows == 0) return VGTKB_OK;
    sa_group_kernel<<<rows_grid(rows), 256, 0, (cudaStream_t)stream>>>(rows, n, s, k, c, cpad, pos, feat, centers, idx, out);
    return check_launch("sa_group_forward");
}

extern "C" int vgtkb_sa_group_backward(int b, int n, int s, int k, int c, int cpad, const float* grad_out, const int32_t* idx,
                                       float* grad_feat, void* stream) {
    VGTKB_REQUIRE(b >= 0 && n >= 1 && s >= 0 && k >= 1 && c >= 1 && cpad >= 3 + c, "sa_group_backward: bad size");
    VGTKB_REQUIRE(idx != nullptr || k == n, "sa_group_backward: identity neighbourhood needs k == n");
    cudaStream_t st = (cudaStream_t)stream;
    if ((size_t)b * n * c > 0) VGTKB_CUDA(cudaMemsetAsync(grad_feat, 0, sizeof(float) * (size_t)b * n * c, st));
    const int64_t rows = (int64_t)b * s * k;
    if (rows == 0) return VGTKB_OK;
    sa_group_bwd_kernel<<<rows_grid(rows), 256, 0, st>>>(rows, n, s, k, c, cpad, grad_out, idx, grad_feat);
    return check_launch("sa_group_backward");
}

extern "C" int vgtkb_sa_maxpool_forward(int64_t groups, int k, int c, const float* y, const float* dist, float radius,
                                        float* out, int32_t* arg, void* stream) {
    VGTKB_REQUIRE(groups >= 0 && k >= 1 && c >= 1, "sa_maxpool: bad size");
    VGTKB_REQUIRE(groups <= 0x7fffffff && c <= 128 * 65535, "sa_maxpool: grid too large");
    if (groups == 0) return VGTKB_OK;
    sa_maxpool_kernel<<<dim3((unsigned)groups, ceil_div(c, 128)), 128, 0, (cudaStream_t)stream>>>(k, c, y, dist, radius, out, arg);
    return check_launch("sa_maxpool_forward");
}

extern "C" int vgtkb_sa_maxpool_backward(int64_t groups, int k, int c, const float* grad_out, const int32_t* arg, float* grad_y,
                                         void* stream) {
    VGTKB_REQUIRE(groups >= 0 && k >= 1 && c >= 1, "sa_maxpool_backward: bad size");
    const int64_t total = groups * k * c;
    if (total == 0) return VGTKB_OK;
    const int64_t want = ceil_div64(total, 256 * 4), cap = (int64_t)kNumSMs * 16;
    sa_maxpool_bwd_kernel<<<(unsigned)(want < cap ? (want < 1 ? 1 : want) : cap), 256, 0, (cudaStream_t)stream>>>(total, k, c, grad_out,
                                                                                                              arg, grad_y);
    return check_launch("sa_maxpool_backward");
}

extern "C" int vgtkb_three_nn(int b, int n1, int n2, const float* p1, const float* p2, int32_t* idx, float* w, void* stream) {
    VGTKB_REQUIRE(b >= 0 && n1 >= 1 && n2 >= 0, "three_nn: bad size");
    VGTKB_REQUIRE(b <= 65535, "three_nn: batch > 65535");
    if (b == 0 || n2 == 0) return VGTKB_OK;
    three_nn_kernel<<<dim3(ceil_div(n2, 128), b), 128, 0, (cudaStream_t)stream>>>(n1, n2, p1, p2, idx, w);
    return check_launch("three_nn");
}

extern "C" int vgtkb_three_interpolate_forward(int b, int n1, int n2, int c, const float* feat, const int32_t* idx, const float* w,
                                               float* out, void* stream) {
    VGTKB_REQUIRE(b >= 0 && n1 >= 1 && n2 >= 0 && c >= 1, "three_interpolate: bad size");
    const int64_t rows = (int64_t)b * n2;
    if (rows == 0) return VGTKB_OK;
    three_interpolate_kernel<<<rows_grid(rows), 256, 0, (cudaStream_t)stream>>>(rows, n1, n2, c, feat, idx, w, out);
    return check_launch("three_interpolate_forward");
}

extern "C" int vgtkb_three_interpolate_backward(int b, int n1, int n2, int c, const float* grad_out, const int32_t* idx,
                                                const float* w, float* grad_feat, void* stream) {
    VGTKB_REQUIRE(b >= 0 && n1 >= 1 && n2 >= 0 && c >= 1, "three_interpolate_backward: bad size");
    cudaStream_t st = (cudaStream_t)stream;
    if ((size_t)b * n1 * c > 0) VGTKB_CUDA(cudaMemsetAsync(grad_feat, 0, sizeof(float) * (size_t)b * n1 * c, st));
    const int64_t rows = (int64_t)b * n2;
    if (rows == 0) return VGTKB_OK;
    three_interpolate_bwd_kernel<<<rows_grid(rows), 256, 0, st>>>(rows, n1, n2, c, grad_out, idx, w, grad_feat);
    return check_launch("three_interpolate_backward");
}
