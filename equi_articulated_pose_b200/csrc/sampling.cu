// sampling.cu -- ball query, furthest point sampling, point gather for sm_100a.
//
// Replaces vgtk/vgtk/cuda/grouping_cuda_kernel.cu (ball_query :67-113, FPS :339-466) and
// gathering_cuda_kernel.cu (:43-98) of the reference.  Same results bit for bit, different
// machine mapping: the reference runs ONE CTA per sample and scans global memory; here the
// support cloud is staged once per CTA into shared memory with the TMA bulk-copy engine, a warp
// owns a query (ballot + popcount appends hits in index order), and FPS keeps its points and
// running distances in registers with a single barrier per round.
#include "common.cuh"

namespace vgtkb {

// ------------------------------------------------------------------------------ ball query
constexpr int BQ_WARPS = 16;
constexpr int BQ_CHUNK = 8192;  // support points staged per pass (96 KB)

__global__ void __launch_bounds__(BQ_WARPS * 32)
ball_query_kernel(int n, int m, float radius2, int nsample, const float* __restrict__ new_xyz,
                  const float* __restrict__ xyz, int32_t* __restrict__ idx) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int chunk = n < BQ_CHUNK ? ((n + 3) & ~3) : BQ_CHUNK;
    float* sx = reinterpret_cast<float*>(smem_raw);
    float* sy = sx + chunk;
    float* sz = sy + chunk;
    int32_t* hits_all = reinterpret_cast<int32_t*>(sz + chunk);
    __shared__ __align__(8) uint64_t bar;

    const int b = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = blockIdx.x * BQ_WARPS + warp;
    const float* sxyz = xyz + (size_t)b * 3 * n;
    int32_t* hits = hits_all + warp * nsample;

    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_barrier_init();
    }
    float qx = 0.f, qy = 0.f, qz = 0.f;
    const bool have_q = q < m;
    if (have_q) {
        const float* nq = new_xyz + (size_t)b * 3 * m;
        qx = nq[q];
        qy = nq[m + q];
        qz = nq[2 * m + q];
    }
    int cnt = 0;
    uint32_t phase = 0;
    // bulk copies need 16-byte aligned rows: base pointer, n and chunk offsets multiples of 4 floats
    const bool aligned = ((reinterpret_cast<uintptr_t>(sxyz) & 15) == 0) && (n % 4 == 0);
    __syncthreads();

    for (int c0 = 0; c0 < n; c0 += chunk) {
        const int len = min(chunk, n - c0);
        if (aligned) {
            if (threadIdx.x == 0) {
                const uint32_t bytes = (uint32_t)len * 4u;
                fence_proxy_async();  // generic-proxy reads of the previous chunk precede the async writes
                mbar_arrive_expect_tx(&bar, 3 * bytes);
                bulk_g2s(sx, sxyz + c0, bytes, &bar);
                bulk_g2s(sy, sxyz + n + c0, bytes, &bar);
                bulk_g2s(sz, sxyz + 2 * (size_t)n + c0, bytes, &bar);
            }
            mbar_wait(&bar, phase);
            phase ^= 1;
        } else {
            for (int i = threadIdx.x; i < len; i += blockDim.x) {
                sx[i] = sxyz[c0 + i];
                sy[i] = sxyz[n + c0 + i];
                sz[i] = sxyz[2 * (size_t)n + c0 + i];
            }
            __syncthreads();
        }
        if (have_q && cnt < nsample) {
            for (int k0 = 0; k0 < len; k0 += 32) {
                const int k = k0 + lane;
                bool hit = false;
                if (k < len) {
                    const float d2 = sq3(qx - sx[k], qy - sy[k], qz - sz[k]);
                    hit = d2 < radius2;
                }
                const unsigned mask = __ballot_sync(0xffffffffu, hit);
                if (hit) {
                    const int pos = cnt + __popc(mask & ((1u << lane) - 1u));
                    if (pos < nsample) hits[pos] = c0 + k;
                }
                cnt += __popc(mask);
                if (cnt >= nsample) break;
            }
        }
        // everyone must be done with this chunk before it is overwritten; stop early when no
        // warp needs more support points
        const int more = __syncthreads_or(have_q && cnt < nsample);
        if (!more) break;
    }
    if (!have_q) return;
    __syncwarp();
    cnt = min(cnt, nsample);
    int32_t* out = idx + ((size_t)b * m + q) * nsample;
    const bool cyclic = cnt > 0 && cnt < nsample - 1;
    for (int i = lane; i < nsample; i += 32) {
        int v = 0;
        if (i < cnt) v = hits[i];
        else if (cyclic) v = hits[i % cnt];
        out[i] = v;
    }
}

// ------------------------------------------------------------------------------ FPS
__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long t = __shfl_xor_sync(0xffffffffu, v, o);
        v = t > v ? t : v;
    }
    return v;
}

// Tie-break of the reference (grouping_cuda_kernel.cu:339-346,382-459): inside a reference
// thread the first strict maximum over k = tid, tid+bs, ... wins; the shared-memory tree keeps
// the LEFT operand on ties, which orders threads by the bit-reversal of tid.  rank(k) =
// bitrev(k mod bs) * iters + (k div bs); the smallest rank among equal distances wins.
template <int PPT, int THREADS>
__global__ void __launch_bounds__(THREADS)
fps_kernel(int n, int m, int log2bs, int iters, int plain, const float* __restrict__ xyz, int32_t* __restrict__ idxs) {
    extern __shared__ __align__(16) float sp[];  // x[n] y[n] z[n]
    __shared__ unsigned long long red[2][THREADS / 32];
    const int b = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* src = xyz + (size_t)b * 3 * n;
    int32_t* out = idxs + (size_t)b * m;
    const int bs = 1 << log2bs;

    for (int i = tid; i < 3 * n; i += THREADS) sp[i] = src[i];
    __syncthreads();
    const float* sx = sp;
    const float* sy = sp + n;
    const float* sz = sp + 2 * n;

    float px[PPT], py[PPT], pz[PPT], tmp[PPT];
    uint32_t inv[PPT];
    bool valid[PPT];
#pragma unroll
    for (int i = 0; i < PPT; ++i) {
        const int k = tid + i * THREADS;
        valid[i] = false;
        px[i] = py[i] = pz[i] = 0.f;
        tmp[i] = 1e10f;
        inv[i] = 0;
        if (k < n) {
            px[i] = sx[k];
            py[i] = sy[k];
            pz[i] = sz[k];
            const float mag = sq3(px[i], py[i], pz[i]);
            valid[i] = plain || !((double)mag <= 1e-3);  // double literal in the reference (:386)
            const uint32_t tr = (uint32_t)k & (uint32_t)(bs - 1);
            const uint32_t br = log2bs ? (__brev(tr) >> (32 - log2bs)) : 0u;
            // plain (torch_cluster-style) sampling: no skipped points, the lowest index wins ties
            inv[i] = 0xffffffffu - (plain ? (uint32_t)k : br * (uint32_t)iters + ((uint32_t)k >> log2bs));
        }
    }
    int old = 0;
    if (tid == 0) out[0] = 0;
    // rank / iters, rank % iters: iters is 1 or 2 up to 2047 points (a shift and a mask); general division otherwise
    const bool ip2 = (iters & (iters - 1)) == 0;
    const uint32_t ilog = 31u - (uint32_t)__clz(iters), imask = (uint32_t)iters - 1u;
    for (int j = 1; j < m; ++j) {
        const float x1 = sx[old], y1 = sy[old], z1 = sz[old];
        uint32_t bhi = 0u, blo = 0u;                     // best (distance bits | sign flag, inverted rank) of this thread
#pragma unroll
        for (int i = 0; i < PPT; ++i) {
            if (valid[i]) {
                const float d = sq3(px[i] - x1, py[i] - y1, pz[i] - z1);
                const float d2 = fminf(d, tmp[i]);
                tmp[i] = d2;
                const uint32_t hi = __float_as_uint(d2) | 0x80000000u;
                const bool better = hi > bhi || (hi == bhi && inv[i] > blo);
                bhi = better ? hi : bhi;
                blo = better ? inv[i] : blo;
            }
        }
        // lexicographic maximum over the warp in two hardware reductions (redux.sync) instead of five 64-bit shuffle steps
        uint32_t whi = __reduce_max_sync(0xffffffffu, bhi);
        uint32_t wlo = __reduce_max_sync(0xffffffffu, bhi == whi ? blo : 0u);
        if (lane == 0) red[j & 1][warp] = ((unsigned long long)whi << 32) | wlo;
        __syncthreads();
        const unsigned long long rv = red[j & 1][lane < THREADS / 32 ? lane : 0];
        const uint32_t rhi = (uint32_t)(rv >> 32), rlo = (uint32_t)rv;
        whi = __reduce_max_sync(0xffffffffu, rhi);
        wlo = __reduce_max_sync(0xffffffffu, rhi == whi ? rlo : 0u);
        if (whi >> 31) {
            const uint32_t rank = 0xffffffffu - wlo;
            const uint32_t br = ip2 ? rank >> ilog : rank / (uint32_t)iters, it = ip2 ? rank & imask : rank % (uint32_t)iters;
            const uint32_t tr = log2bs ? (__brev(br) >> (32 - log2bs)) : 0u;
            old = plain ? (int)rank : (int)(it * (uint32_t)bs + tr);
        } else {
            old = 0;  // every point skipped: the reference's tree returns besti = 0
        }
        if (tid == 0) out[j] = old;
    }
}

// ------------------------------------------------------------------------------ gather
__global__ void gather_fwd_kernel(int c, int n, int m, const float* __restrict__ pts,
                                  const int32_t* __restrict__ idx, float* __restrict__ out) {
    const int b = blockIdx.z, ci = blockIdx.y;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    const int a = idx[(size_t)b * m + j];
    out[((size_t)b * c + ci) * m + j] = pts[((size_t)b * c + ci) * n + a];
}

__global__ void gather_bwd_kernel(int c, int n, int m, const float* __restrict__ gout,
                                  const int32_t* __restrict__ idx, float* __restrict__ gpts) {
    const int b = blockIdx.z, ci = blockIdx.y;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    const int a = idx[(size_t)b * m + j];
    atomicAdd(gpts + ((size_t)b * c + ci) * n + a, gout[((size_t)b * c + ci) * m + j]);
}

// channels-last row gather: out[b,j,:] = x[b,idx[b,j],:], width % 4 == 0 uses float4
template <typename V>
__global__ void row_gather_fwd_kernel(int n, int m, int wv, const V* __restrict__ x,
                                      const int32_t* __restrict__ idx, V* __restrict__ out) {
    const int b = blockIdx.y;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)m * wv) return;
    const int j = (int)(t / wv), e = (int)(t % wv);
    const int src = idx[(size_t)b * m + j];
    out[((size_t)b * m + j) * wv + e] = x[((size_t)b * n + src) * wv + e];
}

__global__ void row_gather_bwd_kernel(int n, int m, int width, const float* __restrict__ gout,
                                      const int32_t* __restrict__ idx, float* __restrict__ gx) {
    const int b = blockIdx.y;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)m * width) return;
    const int j = (int)(t / width), e = (int)(t % width);
    const int src = idx[(size_t)b * m + j];
    atomicAdd(gx + ((size_t)b * n + src) * width + e, gout[((size_t)b * m + j) * width + e]);
}

}  // namespace vgtkb

using namespace vgtkb;

extern "C" int vgtkb_ball_query(int b, int n, int m, float radius, int nsample, const float* new_xyz,
                                const float* xyz, int32_t* idx, void* stream) {
    VGTKB_REQUIRE(b >= 0 && n >= 0 && m >= 0 && nsample >= 0, "ball_query: negative size");
    if (b == 0 || m == 0 || nsample == 0) return VGTKB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0) {  // no support points: reference leaves the zero-initialised output
        VGTKB_CUDA(cudaMemsetAsync(idx, 0, sizeof(int32_t) * (size_t)b * m * nsample, st));
        return VGTKB_OK;
    }
    VGTKB_REQUIRE(b <= 65535, "ball_query: batch %d > 65535", b);
    const int chunk = n < BQ_CHUNK ? ((n + 3) & ~3) : BQ_CHUNK;
    const size_t smem = (size_t)chunk * 12 + (size_t)BQ_WARPS * nsample * 4;
    VGTKB_REQUIRE(smem <= 220 * 1024, "ball_query: nsample %d too large", nsample);
    static bool attr_set = false;
    if (!attr_set) {
        VGTKB_CUDA(cudaFuncSetAttribute(ball_query_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        attr_set = true;
    }
    const float radius2 = radius * radius;  // float product, as in the reference kernel (:82)
    dim3 grid(ceil_div(m, BQ_WARPS), b);
    ball_query_kernel<<<grid, BQ_WARPS * 32, smem, st>>>(n, m, radius2, nsample, new_xyz, xyz, idx);
    return check_launch("ball_query");
}

template <int PPT, int THREADS>
static int launch_fps(int b, int n, int m, int log2bs, int iters, int plain, const float* xyz, int32_t* idx, cudaStream_t st) {
    const size_t smem = (size_t)n * 12;
    auto kern = fps_kernel<PPT, THREADS>;
    VGTKB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    kern<<<b, THREADS, smem, st>>>(n, m, log2bs, iters, plain, xyz, idx);
    return check_launch("furthest_point_sampling");
}

static int fps_dispatch(int b, int n, int m, int plain, const float* xyz, int32_t* idx, void* stream) {
    VGTKB_REQUIRE(b >= 0 && n >= 0 && m >= 0, "fps: negative size");
    if (b == 0 || m <= 0) return VGTKB_OK;
    VGTKB_REQUIRE(n >= 1, "fps: empty cloud");
    VGTKB_REQUIRE(n <= 16384, "fps: n=%d > 16384 points per cloud not supported", n);
    cudaStream_t st = (cudaStream_t)stream;
    // reference block size: largest power of two <= n, at most 1024 (grouping_cuda_kernel.cu:29-33)
    int log2bs = 0;
    while ((2 << log2bs) <= n && log2bs < 10) ++log2bs;
    const int bs = 1 << log2bs;
    const int iters = (n + bs - 1) / bs;
    // few warps per cloud: a round is a latency chain (distance update, two warp reductions, one barrier), not throughput
    if (n <= 512) return launch_fps<4, 128>(b, n, m, log2bs, iters, plain, xyz, idx, st);
    if (n <= 1024) return launch_fps<4, 256>(b, n, m, log2bs, iters, plain, xyz, idx, st);
    if (n <= 2048) return launch_fps<8, 256>(b, n, m, log2bs, iters, plain, xyz, idx, st);
    if (n <= 4096) return launch_fps<8, 512>(b, n, m, log2bs, iters, plain, xyz, idx, st);
    if (n <= 8192) return launch_fps<8, 1024>(b, n, m, log2bs, iters, plain, xyz, idx, st);
    return launch_fps<16, 1024>(b, n, m, log2bs, iters, plain, xyz, idx, st);
}

extern "C" int vgtkb_furthest_point_sampling(int b, int n, int m, const float* xyz, int32_t* idx, void* stream) {
    return fps_dispatch(b, n, m, 0, xyz, idx, stream);
}

// plain farthest point sampling (start at point 0, every point eligible, lowest index on ties): what
// torch_cluster.fps(random_start=False) computes for the reference's wrappers (SPConvNets/models/model_util.py:183-200)
extern "C" int vgtkb_fps_plain(int b, int n, int m, const float* xyz, int32_t* idx, void* stream) {
    return fps_dispatch(b, n, m, 1, xyz, idx, stream);
}

extern "C" int vgtkb_gather_points_forward(int b, int c, int n, int m, const float* points, const int32_t* idx,
                                           float* out, void* stream) {
    VGTKB_REQUIRE(b >= 0 && c >= 0 && n >= 0 && m >= 0, "gather: negative size");
    if (b == 0 || c == 0 || m == 0) return VGTKB_OK;
    VGTKB_REQUIRE(c <= 65535 && b <= 65535, "gather: b or c > 65535");
    dim3 grid(ceil_div(m, 256), c, b);
    gather_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(c, n, m, points, idx, out);
    return check_launch("gather_points_forward");
}

extern "C" int vgtkb_gather_points_backward(int b, int c, int n, int m, const float* grad_out, const int32_t* idx,
                                            float* grad_points, void* stream) {
    VGTKB_REQUIRE(b >= 0 && c >= 0 && n >= 0 && m >= 0, "gather: negative size");
    cudaStream_t st = (cudaStream_t)stream;
    if ((size_t)b * c * n) VGTKB_CUDA(cudaMemsetAsync(grad_points, 0, sizeof(float) * (size_t)b * c * n, st));
    if (b == 0 || c == 0 || m == 0) return VGTKB_OK;
    VGTKB_REQUIRE(c <= 65535 && b <= 65535, "gather: b or c > 65535");
    dim3 grid(ceil_div(m, 256), c, b);
    gather_bwd_kernel<<<grid, 256, 0, st>>>(c, n, m, grad_out, idx, grad_points);
    return check_launch("gather_points_backward");
}

extern "C" int vgtkb_row_gather_forward(int b, int n, int m, int width, const float* x, const int32_t* idx,
                                        float* out, void* stream) {
    VGTKB_REQUIRE(b >= 0 && n >= 0 && m >= 0 && width >= 0, "row_gather: negative size");
    if (b == 0 || m == 0 || width == 0) return VGTKB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const bool v4 = width % 4 == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
    if (v4) {
        const int wv = width / 4;
        dim3 grid((unsigned)ceil_div64((int64_t)m * wv, 256), b);
        row_gather_fwd_kernel<float4><<<grid, 256, 0, st>>>(n, m, wv, (const float4*)x, idx, (float4*)out);
    } else {
        dim3 grid((unsigned)ceil_div64((int64_t)m * width, 256), b);
        row_gather_fwd_kernel<float><<<grid, 256, 0, st>>>(n, m, width, x, idx, out);
    }
    return check_launch("row_gather_forward");
}

extern "C" int vgtkb_row_gather_backward(int b, int n, int m, int width, const float* grad_out, const int32_t* idx,
                                         float* grad_x, void* stream) {
    VGTKB_REQUIRE(b >= 0 && n >= 0 && m >= 0 && width >= 0, "row_gather: negative size");
    if (b == 0 || m == 0 || width == 0) return VGTKB_OK;
    dim3 grid((unsigned)ceil_div64((int64_t)m * width, 256), b);
    row_gather_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(n, m, width, grad_out, idx, grad_x);
    return check_launch("row_gather_backward");
}
