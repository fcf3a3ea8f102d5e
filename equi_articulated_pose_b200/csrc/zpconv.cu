// zpconv.cu -- the four entry points of the reference's `vgtk.cuda.zpconv` extension
// (vgtk/vgtk/cuda/zpconv_cuda.cpp:113-118, kernels zpconv_cuda_kernel.cu:33-195) with the reference tensor layouts:
//   inter_zpconv_forward (idx[B,P,A,K,ann] int32, w[B,P,A,K,ann], feats[B,C,Nq,A])  -> out[B,C,K,P,A]
//   intra_zpconv_forward (idx[Aout,ann] int32,   w[Aout,K,ann],   feats[B,C,P,Ain]) -> out[B,C,K,P,Aout]
// and their transposed scatters.  In the reference these are dead code on the SO(3) path (the Python calls the
// `*_naive` torch versions); they are the live grouping of the legacy S^2 ZPConv modules (config 1b:
// vgtk/vgtk/spconv/modules.py:61-98).  The reference forward kernels scatter with atomicAdd over (C, ann); here the
// forward is a gather (one thread per output element, deterministic), only the backward uses atomics.
#include "common.cuh"

namespace vgtkb {

__global__ void inter_zp_fwd_kernel(int64_t total, int c, int nq, int p, int a, int k, int ann,
                                    const int32_t* __restrict__ idx, const float* __restrict__ w,
                                    const float* __restrict__ feats, float* __restrict__ out) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int ai = (int)(t % a);
        const int pi = (int)((t / a) % p);
        const int ki = (int)((t / ((int64_t)a * p)) % k);
        const int ci = (int)((t / ((int64_t)a * p * k)) % c);
        const int64_t b = t / ((int64_t)a * p * k * c);
        const int64_t base = ((((b * p + pi) * a + ai) * k) + ki) * ann;
        float acc = 0.f;
        for (int n = 0; n < ann; ++n)
            acc = fmaf(w[base + n], feats[((b * c + ci) * nq + idx[base + n]) * a + ai], acc);
        out[t] = acc;
    }
}

__global__ void inter_zp_bwd_kernel(int64_t total, int c, int nq, int p, int a, int k, int ann,
                                    const int32_t* __restrict__ idx, const float* __restrict__ w,
                                    const float* __restrict__ gout, float* __restrict__ gfeats) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int ai = (int)(t % a);
        const int pi = (int)((t / a) % p);
        const int ki = (int)((t / ((int64_t)a * p)) % k);
        const int ci = (int)((t / ((int64_t)a * p * k)) % c);
        const int64_t b = t / ((int64_t)a * p * k * c);
        const int64_t base = ((((b * p + pi) * a + ai) * k) + ki) * ann;
        const float g = gout[t];
        for (int n = 0; n < ann; ++n)
            atomicAdd(gfeats + ((b * c + ci) * nq + idx[base + n]) * a + ai, w[base + n] * g);
    }
}

__global__ void intra_zp_fwd_kernel(int64_t total, int c, int p, int ain, int aout, int k, int ann,
                                    const int32_t* __restrict__ idx, const float* __restrict__ w,
                                    const float* __restrict__ feats, float* __restrict__ out) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int ao = (int)(t % aout);
        const int pi = (int)((t / aout) % p);
        const int ki = (int)((t / ((int64_t)aout * p)) % k);
        const int64_t bc = t / ((int64_t)aout * p * k);
        const float* f = feats + (bc * p + pi) * ain;
        float acc = 0.f;
        for (int n = 0; n < ann; ++n) acc = fmaf(w[((int64_t)ao * k + ki) * ann + n], f[idx[ao * ann + n]], acc);
        out[t] = acc;
    }
}

__global__ void intra_zp_bwd_kernel(int64_t total, int c, int p, int ain, int aout, int k, int ann,
                                    const int32_t* __restrict__ idx, const float* __restrict__ w,
                                    const float* __restrict__ gout, float* __restrict__ gfeats) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int ao = (int)(t % aout);
        const int pi = (int)((t / aout) % p);
        const int ki = (int)((t / ((int64_t)aout * p)) % k);
        const int64_t bc = t / ((int64_t)aout * p * k);
        float* gf = gfeats + (bc * p + pi) * ain;
        const float g = gout[t];
        for (int n = 0; n < ann; ++n) atomicAdd(gf + idx[ao * ann + n], w[((int64_t)ao * k + ki) * ann + n] * g);
    }
}

static unsigned grid_for(int64_t total) {
    const int64_t blocks = ceil_div64(total, 256);
    return (unsigned)(blocks < (int64_t)kNumSMs * 32 ? blocks : kNumSMs * 32);
}

}  // namespace vgtkb

using namespace vgtkb;

extern "C" int vgtkb_inter_zpconv_forward(int b, int c, int nq, int p, int a, int k, int ann, const int32_t* idx,
                                          const float* w, const float* feats, float* out, void* stream) {
    VGTKB_REQUIRE(b >= 0 && c >= 0 && nq > 0 && p >= 0 && a > 0 && k > 0 && ann > 0, "inter_zpconv: bad size");
    const int64_t total = (int64_t)b * c * k * p * a;
    if (total == 0) return VGTKB_OK;
    inter_zp_fwd_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(total, c, nq, p, a, k, ann, idx, w, feats, out);
    return check_launch("inter_zpconv_forward");
}

extern "C" int vgtkb_inter_zpconv_backward(int b, int c, int nq, int p, int a, int k, int ann, const int32_t* idx,
                                           const float* w, const float* grad_out, float* grad_feats, void* stream) {
    VGTKB_REQUIRE(b >= 0 && c >= 0 && nq > 0 && p >= 0 && a > 0 && k > 0 && ann > 0, "inter_zpconv: bad size");
    cudaStream_t st = (cudaStream_t)stream;
    if ((int64_t)b * c * nq * a > 0) VGTKB_CUDA(cudaMemsetAsync(grad_feats, 0, sizeof(float) * (size_t)b * c * nq * a, st));
    const int64_t total = (int64_t)b * c * k * p * a;
    if (total == 0) return VGTKB_OK;
    inter_zp_bwd_kernel<<<grid_for(total), 256, 0, st>>>(total, c, nq, p, a, k, ann, idx, w, grad_out, grad_feats);
    return check_launch("inter_zpconv_backward");
}

extern "C" int vgtkb_intra_zpconv_forward(int b, int c, int p, int ain, int aout, int k, int ann, const int32_t* idx,
                                          const float* w, const float* feats, float* out, void* stream) {
    VGTKB_REQUIRE(b >= 0 && c >= 0 && p >= 0 && ain > 0 && aout > 0 && k > 0 && ann > 0, "intra_zpconv: bad size");
    const int64_t total = (int64_t)b * c * k * p * aout;
    if (total == 0) return VGTKB_OK;
    intra_zp_fwd_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(total, c, p, ain, aout, k, ann, idx, w, feats, out);
    return check_launch("intra_zpconv_forward");
}

extern "C" int vgtkb_intra_zpconv_backward(int b, int c, int p, int ain, int aout, int k, int ann, const int32_t* idx,
                                           const float* w, const float* grad_out, float* grad_feats, void* stream) {
    VGTKB_REQUIRE(b >= 0 && c >= 0 && p >= 0 && ain > 0 && aout > 0 && k > 0 && ann > 0, "intra_zpconv: bad size");
    cudaStream_t st = (cudaStream_t)stream;
    if ((int64_t)b * c * p * ain > 0) VGTKB_CUDA(cudaMemsetAsync(grad_feats, 0, sizeof(float) * (size_t)b * c * p * ain, st));
    const int64_t total = (int64_t)b * c * k * p * aout;
    if (total == 0) return VGTKB_OK;
    intra_zp_bwd_kernel<<<grid_for(total), 256, 0, st>>>(total, c, p, ain, aout, k, ann, idx, w, grad_out, grad_feats);
    return check_launch("intra_zpconv_backward");
}
