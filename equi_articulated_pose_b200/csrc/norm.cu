// norm.cu -- BatchNorm2d / InstanceNorm2d statistics, normalise + leaky_relu (+ residual) and
// their backward, on channels-last rows X[groups][rows][c] for sm_100a.
//
// Stands for torch's BatchNorm2d(train) / InstanceNorm2d(affine=False) + F.leaky_relu around the
// reference blocks (SPConvNets/utils/base_so3conv.py:47-64,113-131,198-217).  HBM-bound
// streaming kernels: float4 accesses, fp64 in-thread accumulation of the per-channel sums (the
// skip branch of the first layer normalises a constant tensor, so cancellation matters), one
// fp64 atomic per (CTA, channel).
#include "common.cuh"

namespace vgtkb {

constexpr int NT = 256;  // threads per CTA

struct OpStats {  // sum x, sum x^2
    __device__ __forceinline__ void operator()(float x, int, int, double& s1, double& s2) const {
        const double d = (double)x;
        s1 += d;
        s2 = fma(d, d, s2);
    }
};

struct OpBwd {  // sum dyp, sum dyp*xhat with dyp = grad_y * lrelu'(xhat*gamma+beta)
    const float* gy;
    const float* stats;
    const float* gamma;
    const float* beta;
    float slope;
    int c;
    __device__ __forceinline__ float dyp(float x, float g, int grp, int ch, float& xh) const {
        const float mean = stats[(size_t)grp * 2 * c + ch], inv = stats[(size_t)grp * 2 * c + c + ch];
        xh = (x - mean) * inv;
        const float pre = xh * (gamma ? gamma[ch] : 1.f) + (beta ? beta[ch] : 0.f);
        return pre > 0.f ? g : g * slope;
    }
};

// column reduction: for every (group, channel) accumulate two fp64 sums over rows.
// MODE 0: stats (x, x^2)   MODE 1: backward sums (needs gy)   MODE 2: plain column sum
template <int MODE>
__global__ void __launch_bounds__(NT)
col_reduce_kernel(int64_t rows, int c, int rows_per_cta, const float* __restrict__ x, OpBwd bw,
                  double* __restrict__ scratch) {
    const int g = blockIdx.y;
    const int tx = min(c, NT), ty = NT / tx;
    const int lx = threadIdx.x % tx, ly = threadIdx.x / tx;
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_cta;
    const int64_t r1 = min(rows, r0 + rows_per_cta);
    const float* xg = x + (size_t)g * rows * c;
    const float* gyg = MODE == 1 ? bw.gy + (size_t)g * rows * c : nullptr;
    __shared__ double sh1[NT], sh2[NT];
    for (int ch = lx; ch < c; ch += tx) {
        double s1 = 0.0, s2 = 0.0;
        if (ly < ty) {
            for (int64_t r = r0 + ly; r < r1; r += ty) {
                const float v = xg[r * c + ch];
                if (MODE == 0) {
                    const double d = (double)v;
                    s1 += d;
                    s2 = fma(d, d, s2);
                } else if (MODE == 1) {
                    float xh;
                    const float d = bw.dyp(v, gyg[r * c + ch], g, ch, xh);
                    s1 += (double)d;
                    s2 += (double)d * (double)xh;
                } else {
                    s1 += (double)v;
                }
            }
        }
        sh1[threadIdx.x] = s1;
        sh2[threadIdx.x] = s2;
        __syncthreads();
        if (ly == 0) {
            for (int j = 1; j < ty; ++j) {
                s1 += sh1[j * tx + lx];
                s2 += sh2[j * tx + lx];
            }
            atomicAdd(scratch + ((size_t)g * 2 + 0) * c + ch, s1);
            if (MODE != 2) atomicAdd(scratch + ((size_t)g * 2 + 1) * c + ch, s2);
        }
        __syncthreads();
    }
}

// float4 variant (c % 4 == 0, 16-byte aligned): thread <-> 4 adjacent channels, rows strided over ty
template <int MODE>
__global__ void __launch_bounds__(NT)
col_reduce4_kernel(int64_t rows, int c, int rows_per_cta, const float* __restrict__ x, OpBwd bw, double* __restrict__ scratch) {
    const int g = blockIdx.y;
    const int c4 = c >> 2;
    const int tx = min(c4, NT), ty = NT / tx;
    const int lx = threadIdx.x % tx, ly = threadIdx.x / tx;
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_cta;
    const int64_t r1 = min(rows, r0 + rows_per_cta);
    const float4* xg = reinterpret_cast<const float4*>(x + (size_t)g * rows * c);
    const float4* gyg = MODE == 1 ? reinterpret_cast<const float4*>(bw.gy + (size_t)g * rows * c) : nullptr;
    __shared__ double sh[NT][8];
    for (int q = lx; q < c4; q += tx) {
        const int ch = q * 4;
        double s1[4] = {0.0, 0.0, 0.0, 0.0}, s2[4] = {0.0, 0.0, 0.0, 0.0};
        float mean[4] = {0.f, 0.f, 0.f, 0.f}, inv[4] = {1.f, 1.f, 1.f, 1.f}, ga[4] = {1.f, 1.f, 1.f, 1.f}, be[4] = {0.f, 0.f, 0.f, 0.f};
        if (MODE == 1) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                mean[i] = bw.stats[(size_t)g * 2 * c + ch + i];
                inv[i] = bw.stats[(size_t)g * 2 * c + c + ch + i];
                if (bw.gamma) ga[i] = bw.gamma[ch + i];
                if (bw.beta) be[i] = bw.beta[ch + i];
            }
        }
        if (ly < ty) {
#pragma unroll 4
            for (int64_t r = r0 + ly; r < r1; r += ty) {
                const float4 v4 = xg[r * c4 + q];
                const float v[4] = {v4.x, v4.y, v4.z, v4.w};
                if (MODE == 0) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const double d = (double)v[i];
                        s1[i] += d;
                        s2[i] = fma(d, d, s2[i]);
                    }
                } else if (MODE == 1) {
                    const float4 g4 = gyg[r * c4 + q];
                    const float gv[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float xh = (v[i] - mean[i]) * inv[i];
                        const float pre = xh * ga[i] + be[i];
                        const float d = pre > 0.f ? gv[i] : gv[i] * bw.slope;
                        s1[i] += (double)d;
                        s2[i] += (double)d * (double)xh;
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 4; ++i) s1[i] += (double)v[i];
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            sh[threadIdx.x][i] = s1[i];
            sh[threadIdx.x][4 + i] = s2[i];
        }
        __syncthreads();
        if (ly == 0) {
            for (int j = 1; j < ty; ++j)
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    s1[i] += sh[j * tx + lx][i];
                    s2[i] += sh[j * tx + lx][4 + i];
                }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                atomicAdd(scratch + ((size_t)g * 2 + 0) * c + ch + i, s1[i]);
                if (MODE != 2) atomicAdd(scratch + ((size_t)g * 2 + 1) * c + ch + i, s2[i]);
            }
        }
        __syncthreads();
    }
}

// batched variant for c/4 dividing NT (every power-of-two channel count up to 1024): a thread keeps ONE channel quad,
// issues U independent 16-byte loads per array before touching any of them (the kernel is latency-bound otherwise:
// 30% of the HBM rate with one load in flight per thread), and folds them into its fp64 sums.
template <int MODE, int U>
__global__ void __launch_bounds__(NT, 2)
col_reduce4b_kernel(int64_t rows, int c, int rows_per_cta, const float* __restrict__ x, OpBwd bw, double* __restrict__ scratch) {
    const int g = blockIdx.y;
    const int c4 = c >> 2;
    const int ty = NT / c4;
    const int q = threadIdx.x % c4, ly = threadIdx.x / c4;
    const int ch = q * 4;
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_cta;
    const int64_t r1 = min(rows, r0 + rows_per_cta);
    const float4* xg = reinterpret_cast<const float4*>(x + (size_t)g * rows * c) + q;
    const float4* gyg = MODE == 1 ? reinterpret_cast<const float4*>(bw.gy + (size_t)g * rows * c) + q : nullptr;
    __shared__ double sh[NT][8];
    double s1[4] = {0.0, 0.0, 0.0, 0.0}, s2[4] = {0.0, 0.0, 0.0, 0.0};
    float mean[4] = {0.f, 0.f, 0.f, 0.f}, inv[4] = {1.f, 1.f, 1.f, 1.f}, ga[4] = {1.f, 1.f, 1.f, 1.f}, be[4] = {0.f, 0.f, 0.f, 0.f};
    if (MODE == 1) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            mean[i] = bw.stats[(size_t)g * 2 * c + ch + i];
            inv[i] = bw.stats[(size_t)g * 2 * c + c + ch + i];
            if (bw.gamma) ga[i] = bw.gamma[ch + i];
            if (bw.beta) be[i] = bw.beta[ch + i];
        }
    }
    // two register buffers: the loads of batch i+1 are in flight while batch i is folded into the fp64 sums (ncu on the
    // one-buffer version: long-scoreboard stalls 9-13 per issue, 44 % of the DRAM rate)
    constexpr int UG = MODE == 1 ? U : 1;
    auto load = [&](float4 (&v4)[U], float4 (&g4)[UG], int64_t r) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t rr = r + (int64_t)u * ty;
            const bool ok = rr < r1;
            v4[u] = ok ? __ldg(xg + rr * c4) : make_float4(0.f, 0.f, 0.f, 0.f);
            if (MODE == 1) g4[u] = ok ? __ldg(gyg + rr * c4) : make_float4(0.f, 0.f, 0.f, 0.f);   // gy = 0: no contribution
        }
    };
    auto fold = [&](const float4 (&v4)[U], const float4 (&g4)[UG]) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const float v[4] = {v4[u].x, v4[u].y, v4[u].z, v4[u].w};
            if (MODE == 0) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const double d = (double)v[i];
                    s1[i] += d;
                    s2[i] = fma(d, d, s2[i]);
                }
            } else if (MODE == 1) {
                const float gv[4] = {g4[u].x, g4[u].y, g4[u].z, g4[u].w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float xh = (v[i] - mean[i]) * inv[i];
                    const float pre = xh * ga[i] + be[i];
                    const float d = pre > 0.f ? gv[i] : gv[i] * bw.slope;
                    s1[i] += (double)d;                              // padded rows: gy = 0, so d = 0
                    s2[i] += (double)d * (double)xh;
                }
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i) s1[i] += (double)v[i];
            }
        }
    };
    {
        const int64_t step = (int64_t)ty * U;
        int64_t r = r0 + ly;
        float4 va[U], ga4[UG], vb[U], gb4[UG];
        if (r < r1) load(va, ga4, r);
        while (r < r1) {
            const int64_t rb = r + step;
            if (rb < r1) load(vb, gb4, rb);
            fold(va, ga4);
            if (rb >= r1) break;
            r = rb + step;
            if (r < r1) load(va, ga4, r);
            fold(vb, gb4);
        }
    }
    if (ty > 1) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            sh[threadIdx.x][i] = s1[i];
            sh[threadIdx.x][4 + i] = s2[i];
        }
        __syncthreads();
        // the NT/c4 partial sums of a channel quad are folded by 8 threads (one per sum) instead of one
        for (int t = threadIdx.x; t < c4 * 8; t += NT) {
            const int qq = t >> 3, i = t & 7;
            if (MODE == 2 && i >= 4) continue;
            double s = 0.0;
            for (int j = 0; j < ty; ++j) s += sh[j * c4 + qq][i];
            atomicAdd(scratch + ((size_t)g * 2 + (i >> 2)) * c + qq * 4 + (i & 3), s);
        }
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            atomicAdd(scratch + ((size_t)g * 2 + 0) * c + ch + i, s1[i]);
            if (MODE != 2) atomicAdd(scratch + ((size_t)g * 2 + 1) * c + ch + i, s2[i]);
        }
    }
}

template <int MODE>
static void launch_col_reduce(int groups, int64_t rows, int c, int rpc, unsigned gx, const float* x, const OpBwd& bw,
                              double* scratch, cudaStream_t st) {
    const bool v4 = c % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
                    (MODE != 1 || (reinterpret_cast<uintptr_t>(bw.gy) & 15) == 0);
    if (v4 && NT % (c / 4) == 0) col_reduce4b_kernel<MODE, (MODE == 1 ? 4 : 8)><<<dim3(gx, groups), NT, 0, st>>>(rows, c, rpc, x, bw, scratch);
    else if (v4) col_reduce4_kernel<MODE><<<dim3(gx, groups), NT, 0, st>>>(rows, c, rpc, x, bw, scratch);
    else col_reduce_kernel<MODE><<<dim3(gx, groups), NT, 0, st>>>(rows, c, rpc, x, bw, scratch);
}

__global__ void stats_finalize_kernel(int groups, int64_t rows, int c, float eps, const double* __restrict__ scratch,
                                      float* __restrict__ stats, float* running_mean, float* running_var,
                                      float momentum) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= groups * c) return;
    const int g = t / c, ch = t % c;
    const double n = rows > 0 ? (double)rows : scratch[(size_t)groups * 2 * c];   // rows <= 0: count stored after the sums
    const double mean = scratch[((size_t)g * 2 + 0) * c + ch] / n;
    double var = scratch[((size_t)g * 2 + 1) * c + ch] / n - mean * mean;
    if (var < 0.0) var = 0.0;
    stats[(size_t)g * 2 * c + ch] = (float)mean;
    stats[(size_t)g * 2 * c + c + ch] = (float)(1.0 / sqrt(var + (double)eps));
    if (running_mean != nullptr && g == 0) {  // BatchNorm (groups == 1): torch uses the unbiased variance here
        const double unbiased = n > 1.0 ? var * n / (n - 1.0) : var;
        running_mean[ch] = (float)((1.0 - momentum) * running_mean[ch] + momentum * mean);
        running_var[ch] = (float)((1.0 - momentum) * running_var[ch] + momentum * unbiased);
    }
}

__global__ void norm_act_fwd_kernel(int64_t total4, int64_t rows, int c4, const float4* __restrict__ x,
                                    const float* __restrict__ stats, const float* __restrict__ gamma,
                                    const float* __restrict__ beta, float slope, const float4* __restrict__ res,
                                    float4* __restrict__ y) {
    const int c = c4 * 4;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total4; t += (int64_t)gridDim.x * blockDim.x) {
        const int ch = (int)(t % c4) * 4;
        const int g = (int)(t / (rows * c4));
        const float4 v = x[t];
        const float4 mean = *reinterpret_cast<const float4*>(stats + (size_t)g * 2 * c + ch);
        const float4 inv = *reinterpret_cast<const float4*>(stats + (size_t)g * 2 * c + c + ch);
        float4 ga = make_float4(1.f, 1.f, 1.f, 1.f), be = make_float4(0.f, 0.f, 0.f, 0.f);
        if (gamma) ga = *reinterpret_cast<const float4*>(gamma + ch);
        if (beta) be = *reinterpret_cast<const float4*>(beta + ch);
        float o[4] = {(v.x - mean.x) * inv.x * ga.x + be.x, (v.y - mean.y) * inv.y * ga.y + be.y,
                      (v.z - mean.z) * inv.z * ga.z + be.z, (v.w - mean.w) * inv.w * ga.w + be.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) o[i] = o[i] > 0.f ? o[i] : o[i] * slope;
        if (res) {
            const float4 r = res[t];
            o[0] += r.x; o[1] += r.y; o[2] += r.z; o[3] += r.w;
        }
        y[t] = make_float4(o[0], o[1], o[2], o[3]);
    }
}

__global__ void norm_act_fwd_scalar_kernel(int64_t total, int64_t rows, int c, const float* __restrict__ x,
                                           const float* __restrict__ stats, const float* __restrict__ gamma,
                                           const float* __restrict__ beta, float slope, const float* __restrict__ res,
                                           float* __restrict__ y) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int ch = (int)(t % c);
        const int g = (int)(t / (rows * c));
        float o = (x[t] - stats[(size_t)g * 2 * c + ch]) * stats[(size_t)g * 2 * c + c + ch];
        o = o * (gamma ? gamma[ch] : 1.f) + (beta ? beta[ch] : 0.f);
        o = o > 0.f ? o : o * slope;
        if (res) o += res[t];
        y[t] = o;
    }
}

// dx = gamma*invstd*(dyp - S1/n - xhat*S2/n)
__global__ void norm_act_bwd_apply_kernel(int64_t total, int64_t rows, int64_t nrows, int c, const float* __restrict__ x,
                                          OpBwd bw, const double* __restrict__ scratch, float* __restrict__ gx) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int ch = (int)(t % c);
        const int g = (int)(t / (rows * c));
        float xh;
        const float d = bw.dyp(x[t], bw.gy[t], g, ch, xh);
        const double nn = nrows > 0 ? (double)nrows : scratch[(size_t)(total / (rows * c)) * 2 * c];   // count behind the sums
        const float m1 = (float)(scratch[((size_t)g * 2 + 0) * c + ch] / nn);
        const float m2 = (float)(scratch[((size_t)g * 2 + 1) * c + ch] / nn);
        const float inv = bw.stats[(size_t)g * 2 * c + c + ch];
        const float ga = bw.gamma ? bw.gamma[ch] : 1.f;
        gx[t] = ga * inv * (d - m1 - xh * m2);
    }
}

__global__ void norm_act_bwd_apply4_kernel(int64_t total4, int64_t rows, int64_t nrows, int c4, const float4* __restrict__ x, OpBwd bw,
                                           const double* __restrict__ scratch, float4* __restrict__ gx) {
    const int c = c4 * 4;
    const float4* gy = reinterpret_cast<const float4*>(bw.gy);
    const double inv_rows = 1.0 / (nrows > 0 ? (double)nrows : scratch[(size_t)(total4 / (rows * c4)) * 2 * c]);
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total4; t += (int64_t)gridDim.x * blockDim.x) {
        const int ch = (int)(t % c4) * 4;
        const int g = (int)(t / (rows * c4));
        const float4 xv = x[t], gv = gy[t];
        const float xs[4] = {xv.x, xv.y, xv.z, xv.w}, gs[4] = {gv.x, gv.y, gv.z, gv.w};
        float o[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float mean = bw.stats[(size_t)g * 2 * c + ch + i], inv = bw.stats[(size_t)g * 2 * c + c + ch + i];
            const float ga = bw.gamma ? bw.gamma[ch + i] : 1.f, be = bw.beta ? bw.beta[ch + i] : 0.f;
            const float xh = (xs[i] - mean) * inv;
            const float d = (xh * ga + be) > 0.f ? gs[i] : gs[i] * bw.slope;
            const float m1 = (float)(scratch[((size_t)g * 2 + 0) * c + ch + i] * inv_rows);
            const float m2 = (float)(scratch[((size_t)g * 2 + 1) * c + ch + i] * inv_rows);
            o[i] = ga * inv * (d - m1 - xh * m2);
        }
        gx[t] = make_float4(o[0], o[1], o[2], o[3]);
    }
}

// batched forward / backward apply for c/4 dividing NT: grid (x, groups); the channel quad of a thread never changes
// (the grid stride is a multiple of c/4), so the per-channel coefficients live in registers and a thread keeps U
// independent 16-byte loads per array in flight.
// bf16 hi / lo planes of four adjacent values (the operand split of the bf16x3 contractions, tc_common.cuh): written
// next to the fp32 result when the consumer of this tensor is a tensor-core contraction, which then loads its operand
// by TMA in the MMA's layout instead of converting it in the kernel (once here instead of once per k-block there).
__device__ __forceinline__ uint32_t pack_bf16x2_rn(float a, float b) {   // a -> low half
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
}
__device__ __forceinline__ void store_planes4(uint2* hi, uint2* lo, size_t i, const float (&o)[4]) {
    const uint32_t h0 = pack_bf16x2_rn(o[0], o[1]), h1 = pack_bf16x2_rn(o[2], o[3]);
    const uint32_t l0 = pack_bf16x2_rn(o[0] - __uint_as_float(h0 << 16), o[1] - __uint_as_float(h0 & 0xFFFF0000u));
    const uint32_t l1 = pack_bf16x2_rn(o[2] - __uint_as_float(h1 << 16), o[3] - __uint_as_float(h1 & 0xFFFF0000u));
    hi[i] = make_uint2(h0, h1);
    lo[i] = make_uint2(l0, l1);
}

template <int U>
__global__ void __launch_bounds__(NT, 4)
norm_act_fwd4b_kernel(int64_t rows4, int c4, const float4* __restrict__ x, const float* __restrict__ stats,
                      const float* __restrict__ gamma, const float* __restrict__ beta, float slope,
                      const float4* __restrict__ res, float4* __restrict__ y, uint2* __restrict__ y_hi,
                      uint2* __restrict__ y_lo) {
    const int c = c4 * 4, g = blockIdx.y;
    const int64_t t0 = (int64_t)blockIdx.x * NT + threadIdx.x, stride = (int64_t)gridDim.x * NT;
    const int ch = (int)(t0 % c4) * 4;
    const size_t base = (size_t)g * rows4;
    const float4 mean = *reinterpret_cast<const float4*>(stats + (size_t)g * 2 * c + ch);
    const float4 inv = *reinterpret_cast<const float4*>(stats + (size_t)g * 2 * c + c + ch);
    float4 ga = make_float4(1.f, 1.f, 1.f, 1.f), be = make_float4(0.f, 0.f, 0.f, 0.f);
    if (gamma) ga = *reinterpret_cast<const float4*>(gamma + ch);
    if (beta) be = *reinterpret_cast<const float4*>(beta + ch);
    for (int64_t t = t0; t < rows4; t += stride * U) {
        float4 v[U], r[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t tt = t + u * stride;
            if (tt < rows4) {
                v[u] = __ldg(x + base + tt);
                if (res) r[u] = __ldg(res + base + tt);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t tt = t + u * stride;
            if (tt < rows4) {
                float o[4] = {(v[u].x - mean.x) * inv.x * ga.x + be.x, (v[u].y - mean.y) * inv.y * ga.y + be.y,
                              (v[u].z - mean.z) * inv.z * ga.z + be.z, (v[u].w - mean.w) * inv.w * ga.w + be.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) o[i] = o[i] > 0.f ? o[i] : o[i] * slope;
                if (res) { o[0] += r[u].x; o[1] += r[u].y; o[2] += r[u].z; o[3] += r[u].w; }
                y[base + tt] = make_float4(o[0], o[1], o[2], o[3]);
                if (y_hi != nullptr) store_planes4(y_hi, y_lo, base + tt, o);
            }
        }
    }
}

template <int U>
__global__ void __launch_bounds__(NT, 3)
norm_act_bwd_apply4b_kernel(int64_t rows4, int64_t rows, int64_t nrows, int c4, const float4* __restrict__ x, OpBwd bw,
                            const double* __restrict__ scratch, float4* __restrict__ gx, uint2* __restrict__ gx_hi,
                            uint2* __restrict__ gx_lo, float* __restrict__ ggamma = nullptr, float* __restrict__ gbeta = nullptr) {
    const int c = c4 * 4, g = blockIdx.y;
    // affine gradients (sum dyp * xhat, sum dyp: the two sums themselves, over all groups) written by one CTA of this launch
    // instead of a launch of their own (affine_grad_kernel) -- single-process statistics only, where scratch holds local sums
    if ((ggamma != nullptr || gbeta != nullptr) && blockIdx.x == 0 && blockIdx.y == 0) {
        for (int cc = threadIdx.x; cc < c; cc += NT) {
            double s1 = 0.0, s2 = 0.0;
            for (int gg = 0; gg < (int)gridDim.y; ++gg) {
                s1 += scratch[((size_t)gg * 2 + 0) * c + cc];
                s2 += scratch[((size_t)gg * 2 + 1) * c + cc];
            }
            if (gbeta) gbeta[cc] = (float)s1;
            if (ggamma) ggamma[cc] = (float)s2;
        }
    }
    const int64_t t0 = (int64_t)blockIdx.x * NT + threadIdx.x, stride = (int64_t)gridDim.x * NT;
    const int ch = (int)(t0 % c4) * 4;
    const size_t base = (size_t)g * rows4;
    const float4* gy = reinterpret_cast<const float4*>(bw.gy);
    const double inv_rows = 1.0 / (nrows > 0 ? (double)nrows : scratch[(size_t)gridDim.y * 2 * c]);
    float mean[4], inv[4], ga[4], be[4], m1[4], m2[4], a[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        mean[i] = bw.stats[(size_t)g * 2 * c + ch + i];
        inv[i] = bw.stats[(size_t)g * 2 * c + c + ch + i];
        ga[i] = bw.gamma ? bw.gamma[ch + i] : 1.f;
        be[i] = bw.beta ? bw.beta[ch + i] : 0.f;
        m1[i] = (float)(scratch[((size_t)g * 2 + 0) * c + ch + i] * inv_rows);
        m2[i] = (float)(scratch[((size_t)g * 2 + 1) * c + ch + i] * inv_rows);
        a[i] = ga[i] * inv[i];
    }
    for (int64_t t = t0; t < rows4; t += stride * U) {
        float4 xv[U], gv[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t tt = t + u * stride;
            if (tt < rows4) {
                xv[u] = __ldcs(x + base + tt);      // last use of both tensors
                gv[u] = __ldcs(gy + base + tt);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t tt = t + u * stride;
            if (tt < rows4) {
                const float xs[4] = {xv[u].x, xv[u].y, xv[u].z, xv[u].w}, gs[4] = {gv[u].x, gv[u].y, gv[u].z, gv[u].w};
                float o[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float xh = (xs[i] - mean[i]) * inv[i];
                    const float d = (xh * ga[i] + be[i]) > 0.f ? gs[i] : gs[i] * bw.slope;
                    o[i] = a[i] * (d - m1[i] - xh * m2[i]);
                }
                gx[base + tt] = make_float4(o[0], o[1], o[2], o[3]);
                if (gx_hi != nullptr) store_planes4(gx_hi, gx_lo, base + tt, o);
            }
        }
    }
}

__global__ void affine_grad_kernel(int groups, int c, const double* __restrict__ scratch, float* ggamma, float* gbeta) {
    const int ch = blockIdx.x * blockDim.x + threadIdx.x;
    if (ch >= c) return;
    double s1 = 0.0, s2 = 0.0;
    for (int g = 0; g < groups; ++g) {
        s1 += scratch[((size_t)g * 2 + 0) * c + ch];
        s2 += scratch[((size_t)g * 2 + 1) * c + ch];
    }
    if (gbeta) gbeta[ch] = (float)s1;
    if (ggamma) ggamma[ch] = (float)s2;
}

__global__ void col_sum_finalize_kernel(int c, const double* __restrict__ scratch, float* out) {
    const int ch = blockIdx.x * blockDim.x + threadIdx.x;
    if (ch < c) out[ch] = (float)scratch[ch];
}

static int reduce_geometry(int64_t rows, int c, int groups, int& rows_per_cta, unsigned& gx, int ctas_per_sm = 4) {
    // exactly one wave of resident CTAs overall (ncu: 591 CTAs on 444 slots = 1.33 waves left the second wave a third full),
    // at least 64 rows per CTA
    int64_t want = (int64_t)kNumSMs * ctas_per_sm / (groups > 0 ? groups : 1);
    if (want < 1) want = 1;
    int64_t rpc = ceil_div64(rows, want);
    if (rpc < 64) rpc = 64;
    rows_per_cta = (int)rpc;
    gx = (unsigned)ceil_div64(rows, rpc);
    return 0;
}

}  // namespace vgtkb

using namespace vgtkb;

// phase 1 of the statistics: local fp64 sums (sum x, sum x^2) per (group, channel) into scratch [groups][2][c]
extern "C" int vgtkb_norm_sums(int groups, int64_t rows, int c, const float* x, double* scratch, void* stream) {
    VGTKB_REQUIRE(groups > 0 && rows > 0 && c > 0, "norm_sums: bad size");
    VGTKB_REQUIRE(groups <= 65535, "norm_sums: too many groups");
    cudaStream_t st = (cudaStream_t)stream;
    VGTKB_CUDA(cudaMemsetAsync(scratch, 0, sizeof(double) * (size_t)groups * 2 * c, st));
    int rpc;
    unsigned gx;
    reduce_geometry(rows, c, groups, rpc, gx, 2);
    OpBwd dummy{};
    launch_col_reduce<0>(groups, rows, c, rpc, gx, x, dummy, scratch, st);
    return check_launch("norm_sums");
}

// phase 2: (mean, invstd) and the running statistics from sums over `total_rows` rows (all ranks' rows under SyncBN)
extern "C" int vgtkb_norm_finalize(int groups, int64_t total_rows, int c, float eps, const double* scratch, float* stats,
                                   float* running_mean, float* running_var, float momentum, void* stream) {
    VGTKB_REQUIRE(groups > 0 && c > 0, "norm_finalize: bad size");
    cudaStream_t st = (cudaStream_t)stream;
    stats_finalize_kernel<<<ceil_div(groups * c, 128), 128, 0, st>>>(groups, total_rows, c, eps, scratch, stats, running_mean,
                                                                    running_var, momentum);
    return check_launch("norm_finalize");
}

extern "C" int vgtkb_norm_stats(int groups, int64_t rows, int c, const float* x, float eps, double* scratch,
                                float* stats, float* running_mean, float* running_var, float momentum, void* stream) {
    const int rc = vgtkb_norm_sums(groups, rows, c, x, scratch, stream);
    if (rc != 0) return rc;
    return vgtkb_norm_finalize(groups, rows, c, eps, scratch, stats, running_mean, running_var, momentum, stream);
}

extern "C" int vgtkb_norm_act_forward(int groups, int64_t rows, int c, const float* x, const float* stats,
                                      const float* gamma, const float* beta, float slope, const float* residual,
                                      float* y, void* stream) {
    return vgtkb_norm_act_forward_planes(groups, rows, c, x, stats, gamma, beta, slope, residual, y, nullptr, nullptr, stream);
}

extern "C" int vgtkb_norm_act_forward_planes(int groups, int64_t rows, int c, const float* x, const float* stats,
                                             const float* gamma, const float* beta, float slope, const float* residual,
                                             float* y, void* y_hi, void* y_lo, void* stream) {
    VGTKB_REQUIRE(groups > 0 && rows > 0 && c > 0, "norm_act: bad size");
    VGTKB_REQUIRE((y_hi == nullptr) == (y_lo == nullptr), "norm_act: both planes or none");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t total = (int64_t)groups * rows * c;
    auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    const bool v4 = c % 4 == 0 && al(x) && al(y) && al(stats) && (!gamma || al(gamma)) && (!beta || al(beta)) &&
                    (!residual || al(residual));
    const int64_t work = v4 ? total / 4 : total;
    const unsigned grid = (unsigned)(ceil_div64(work, NT) < (int64_t)kNumSMs * 16 ? ceil_div64(work, NT) : kNumSMs * 16);
    const bool planes_ok = v4 && NT % (c / 4) == 0 && groups <= 65535 &&
                           ((reinterpret_cast<uintptr_t>(y_hi) | reinterpret_cast<uintptr_t>(y_lo)) & 7) == 0;
    VGTKB_REQUIRE(y_hi == nullptr || planes_ok, "norm_act: planes need c %% 4 == 0, 256 %% (c/4) == 0 and aligned tensors");
    if (v4 && NT % (c / 4) == 0 && groups <= 65535) {
        const int64_t rows4 = rows * (c / 4);
        int64_t gx = ceil_div64(rows4, (int64_t)NT * 4);
        const int64_t cap = ceil_div64((int64_t)kNumSMs * 8, groups);
        if (gx > cap) gx = cap;
        norm_act_fwd4b_kernel<4><<<dim3((unsigned)gx, groups), NT, 0, st>>>(rows4, c / 4, (const float4*)x, stats, gamma, beta, slope,
                                                                           (const float4*)residual, (float4*)y, (uint2*)y_hi,
                                                                           (uint2*)y_lo);
    } else if (v4)
        norm_act_fwd_kernel<<<grid, NT, 0, st>>>(work, rows, c / 4, (const float4*)x, stats, gamma, beta, slope,
                                                 (const float4*)residual, (float4*)y);
    else
        norm_act_fwd_scalar_kernel<<<grid, NT, 0, st>>>(total, rows, c, x, stats, gamma, beta, slope, residual, y);
    return check_launch("norm_act_forward");
}

// backward phase 1: local fp64 sums (sum dyp, sum dyp*xhat) into scratch; grad_gamma / grad_beta are LOCAL sums
// (under data parallelism they are reduced with the other parameter gradients, as torch's SyncBatchNorm does)
extern "C" int vgtkb_norm_bwd_sums(int groups, int64_t rows, int c, const float* x, const float* stats, const float* gamma,
                                   const float* beta, float slope, const float* grad_y, double* scratch, float* grad_gamma,
                                   float* grad_beta, void* stream) {
    VGTKB_REQUIRE(groups > 0 && rows > 0 && c > 0, "norm_bwd_sums: bad size");
    VGTKB_REQUIRE(groups <= 65535, "norm_bwd_sums: too many groups");
    cudaStream_t st = (cudaStream_t)stream;
    VGTKB_CUDA(cudaMemsetAsync(scratch, 0, sizeof(double) * (size_t)groups * 2 * c, st));
    OpBwd bw{grad_y, stats, gamma, beta, slope, c};
    int rpc;
    unsigned gx;
    reduce_geometry(rows, c, groups, rpc, gx, 2);
    launch_col_reduce<1>(groups, rows, c, rpc, gx, x, bw, scratch, st);
    if (grad_gamma || grad_beta)
        affine_grad_kernel<<<ceil_div(c, 128), 128, 0, st>>>(groups, c, scratch, grad_gamma, grad_beta);
    return check_launch("norm_bwd_sums");
}

// backward phase 2: dx = gamma*invstd*(dyp - S1/n - xhat*S2/n) with n = total_rows (all ranks' rows under SyncBN)
extern "C" int vgtkb_norm_bwd_apply(int groups, int64_t rows, int64_t total_rows, int c, const float* x, const float* stats,
                                    const float* gamma, const float* beta, float slope, const float* grad_y,
                                    const double* scratch, float* grad_x, void* stream) {
    return vgtkb_norm_bwd_apply_planes(groups, rows, total_rows, c, x, stats, gamma, beta, slope, grad_y, scratch, grad_x, nullptr,
                                       nullptr, stream);
}

// fused: 1 = the 4b kernel ran and also wrote grad_gamma / grad_beta (when given); 0 = another kernel ran, the caller launches
// affine_grad_kernel itself
static int norm_bwd_apply_impl(int groups, int64_t rows, int64_t total_rows, int c, const float* x, const float* stats,
                               const float* gamma, const float* beta, float slope, const float* grad_y, const double* scratch,
                               float* grad_x, void* gx_hi, void* gx_lo, float* grad_gamma, float* grad_beta, int* fused,
                               void* stream);

extern "C" int vgtkb_norm_bwd_apply_planes(int groups, int64_t rows, int64_t total_rows, int c, const float* x, const float* stats,
                                           const float* gamma, const float* beta, float slope, const float* grad_y,
                                           const double* scratch, float* grad_x, void* gx_hi, void* gx_lo, void* stream) {
    int fused = 0;
    return norm_bwd_apply_impl(groups, rows, total_rows, c, x, stats, gamma, beta, slope, grad_y, scratch, grad_x, gx_hi, gx_lo,
                               nullptr, nullptr, &fused, stream);
}

static int norm_bwd_apply_impl(int groups, int64_t rows, int64_t total_rows, int c, const float* x, const float* stats,
                               const float* gamma, const float* beta, float slope, const float* grad_y, const double* scratch,
                               float* grad_x, void* gx_hi, void* gx_lo, float* grad_gamma, float* grad_beta, int* fused,
                               void* stream) {
    *fused = 0;
    VGTKB_REQUIRE(groups > 0 && rows > 0 && (total_rows <= 0 || total_rows >= rows) && c > 0, "norm_bwd_apply: bad size");
    VGTKB_REQUIRE((gx_hi == nullptr) == (gx_lo == nullptr), "norm_bwd_apply: both planes or none");
    VGTKB_REQUIRE(groups <= 65535, "norm_bwd_apply: too many groups");
    cudaStream_t st = (cudaStream_t)stream;
    OpBwd bw{grad_y, stats, gamma, beta, slope, c};
    const int64_t total = (int64_t)groups * rows * c;
    auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    if (c % 4 == 0 && NT % (c / 4) == 0 && al16(x) && al16(grad_y) && al16(grad_x)) {
        const int64_t rows4 = rows * (c / 4);
        int64_t gx4 = ceil_div64(rows4, (int64_t)NT * 4);
        const int64_t cap = ceil_div64((int64_t)kNumSMs * 6, groups);
        if (gx4 > cap) gx4 = cap;
        norm_act_bwd_apply4b_kernel<4><<<dim3((unsigned)gx4, groups), NT, 0, st>>>(rows4, rows, total_rows, c / 4, (const float4*)x, bw,
                                                                                  scratch, (float4*)grad_x, (uint2*)gx_hi, (uint2*)gx_lo,
                                                                                  grad_gamma, grad_beta);
        *fused = 1;
    } else if (gx_hi != nullptr) {
        set_error("norm_bwd_apply: planes need c %% 4 == 0, 256 %% (c/4) == 0 and 16-byte aligned tensors");
        return VGTKB_EINVAL;
    } else if (c % 4 == 0 && al16(x) && al16(grad_y) && al16(grad_x)) {
        const int64_t total4 = total / 4;
        const unsigned grid4 = (unsigned)(ceil_div64(total4, NT) < (int64_t)kNumSMs * 16 ? ceil_div64(total4, NT) : kNumSMs * 16);
        norm_act_bwd_apply4_kernel<<<grid4, NT, 0, st>>>(total4, rows, total_rows, c / 4, (const float4*)x, bw, scratch, (float4*)grad_x);
    } else {
        const unsigned grid = (unsigned)(ceil_div64(total, NT) < (int64_t)kNumSMs * 16 ? ceil_div64(total, NT) : kNumSMs * 16);
        norm_act_bwd_apply_kernel<<<grid, NT, 0, st>>>(total, rows, total_rows, c, x, bw, scratch, grad_x);
    }
    return check_launch("norm_bwd_apply");
}

extern "C" int vgtkb_norm_act_backward(int groups, int64_t rows, int c, const float* x, const float* stats,
                                       const float* gamma, const float* beta, float slope, const float* grad_y,
                                       double* scratch, float* grad_x, float* grad_gamma, float* grad_beta,
                                       void* stream) {
    return vgtkb_norm_act_backward_planes(groups, rows, c, x, stats, gamma, beta, slope, grad_y, scratch, grad_x, grad_gamma,
                                          grad_beta, nullptr, nullptr, stream);
}

extern "C" int vgtkb_norm_act_backward_planes(int groups, int64_t rows, int c, const float* x, const float* stats,
                                              const float* gamma, const float* beta, float slope, const float* grad_y,
                                              double* scratch, float* grad_x, float* grad_gamma, float* grad_beta,
                                              void* gx_hi, void* gx_lo, void* stream) {
    // the affine gradients are the two local sums: written by the apply kernel (one launch less per BatchNorm backward)
    int rc = vgtkb_norm_bwd_sums(groups, rows, c, x, stats, gamma, beta, slope, grad_y, scratch, nullptr, nullptr, stream);
    if (rc != 0) return rc;
    int fused = 0;
    rc = norm_bwd_apply_impl(groups, rows, rows, c, x, stats, gamma, beta, slope, grad_y, scratch, grad_x, gx_hi, gx_lo, grad_gamma,
                             grad_beta, &fused, stream);
    if (rc != 0) return rc;
    if (!fused && (grad_gamma || grad_beta)) {
        affine_grad_kernel<<<ceil_div(c, 128), 128, 0, (cudaStream_t)stream>>>(groups, c, scratch, grad_gamma, grad_beta);
        return check_launch("norm_act_backward");
    }
    return VGTKB_OK;
}

extern "C" int vgtkb_col_sum(int64_t rows, int c, const float* x, double* scratch, float* out, void* stream) {
    VGTKB_REQUIRE(rows > 0 && c > 0, "col_sum: bad size");
    cudaStream_t st = (cudaStream_t)stream;
    VGTKB_CUDA(cudaMemsetAsync(scratch, 0, sizeof(double) * (size_t)2 * c, st));
    int rpc;
    unsigned gx;
    reduce_geometry(rows, c, 1, rpc, gx, 2);
    OpBwd dummy{};
    launch_col_reduce<2>(1, rows, c, rpc, gx, x, dummy, scratch, st);
    col_sum_finalize_kernel<<<ceil_div(c, 128), 128, 0, st>>>(c, scratch, out);
    return check_launch("col_sum");
}
