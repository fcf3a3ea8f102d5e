// inter_conv.cu -- InterSO3Conv as ONE call per direction: kernel-point correlation + neighbour grouping + contraction
// with W (vgtk/vgtk/so3conv/functional.py:144-203 -> vgtk/vgtk/spconv/functional.py:375-406 ->
// vgtk/vgtk/so3conv/modules.py:48-55) and its backward.
//
// The grouped tensor G [rows, k*ci] is produced ONCE, already in the operand format of the bf16x3 tensor-core
// contraction (two bf16 planes hi / lo, written by the grouping kernel through tensor-map stores), and is consumed as
// pure TMA -> tcgen05 streams by the forward contraction and by the weight gradient: no fp32 copy of G exists and no
// kernel converts it again.  (Why G is not kept on chip: DESIGN.md section 6.1 -- per 128-row MMA tile the producer
// state that must stay resident, the per-row correlation weights, plus the G slab exceed shared memory + TMEM.)
#include "tc_common.cuh"

namespace vgtkb {

int inter_group_forward_planes(int b, int n, int p, int nn, int a, int k, int ci, const float* xyz, const float* sample_xyz,
                               const int32_t* idx, const float* rot_kernels, float sigma, const float* feats, void* g_hi,
                               void* g_lo, int fast, cudaStream_t st);
int inter_group_backward_gen(int b, int n, int p, int nn, int a, int k, int ci, const float* xyz, const float* sample_xyz,
                             const int32_t* idx, const float* rot_kernels, float sigma, const float* grad_grouped,
                             float* grad_feats, int fast, cudaStream_t st);
int tc_gemm_nt_planes(int64_t M, int N, int K, const void* a_hi, const void* a_lo, const float* B, const float* bias, float* C,
                      float* workspace, cudaStream_t st, int fast);
int tc_gemm_tn_planes(int M, int N, int64_t R, const float* A, const void* a_hi, const void* a_lo, const float* B,
                      const void* b_hi, const void* b_lo, float* C, int accumulate, float* workspace, cudaStream_t st, int fast);
int tc_gemm_nt(int64_t M, int N, int K, const float* A, const float* B, const float* bias, float* C, int passes,
               float* workspace, cudaStream_t st);

// out[c][r] = in[r][c]   (W [co, k*ci] -> W^T [k*ci, co]: the B operand of dG = gy W; a few MB at most)
__global__ void transpose_kernel(int rows, int cols, const float* __restrict__ in, float* __restrict__ out) {
    __shared__ float tile[32][33];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int r = r0 + i, c = c0 + threadIdx.x;
        tile[i][threadIdx.x] = (r < rows && c < cols) ? in[(size_t)r * cols + c] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int c = c0 + i, r = r0 + threadIdx.x;
        if (r < rows && c < cols) out[(size_t)c * rows + r] = tile[threadIdx.x][i];
    }
}

}  // namespace vgtkb

using namespace vgtkb;

static bool conv_shape_ok(int b, int n, int p, int nn, int a, int k, int ci, int co) {
    return b > 0 && n > 0 && p > 0 && nn > 0 && a > 0 && k > 0 && ci > 0 && co > 0 && k <= 24 && nn <= 64 && ci % 32 == 0 &&
           co % 8 == 0 && co <= 1024 && (int64_t)b * p * a >= 64;
}

extern "C" int vgtkb_inter_conv_supported(int b, int n, int p, int nn, int a, int k, int ci, int co) {
    return conv_shape_ok(b, n, p, nn, a, k, ci, co) && (int64_t)n * a * ci < ((int64_t)1 << 32) &&
                   (int64_t)b * p * a * k < ((int64_t)1 << 31) && b <= 65535
               ? 1
               : 0;
}

extern "C" int vgtkb_inter_conv_forward(int b, int n, int p, int nn, int a, int k, int ci, int co, const float* xyz,
                                        const float* sample_xyz, const int32_t* idx, const float* rot_kernels, float sigma,
                                        const float* feats, const float* w_kc, void* g_hi, void* g_lo, float* workspace,
                                        float* out, int mode, void* stream) {
    VGTKB_REQUIRE(conv_shape_ok(b, n, p, nn, a, k, ci, co),
                  "inter_conv_forward: unsupported shape (needs k <= 24, nn <= 64, ci %% 32 == 0, co %% 8 == 0, co <= 1024)");
    VGTKB_REQUIRE(mode == 3 || mode == 4, "inter_conv_forward: mode %d (3 = bf16x3, 4 = single-pass bf16)", mode);
    const int fast = mode == 4;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = inter_group_forward_planes(b, n, p, nn, a, k, ci, xyz, sample_xyz, idx, rot_kernels, sigma, feats, g_hi, g_lo, fast, st);
    if (rc == VGTKB_EUNSUP) set_error("inter_conv_forward: tensors misaligned or too large for the plane grouping kernel");
    if (rc) return rc;
    rc = tc_gemm_nt_planes((int64_t)b * p * a, co, k * ci, g_hi, g_lo, w_kc, nullptr, out, workspace, st, fast);
    if (rc == VGTKB_EUNSUP) set_error("inter_conv_forward: contraction operands must be 16-byte aligned (workspace: co*k*ci floats)");
    return rc;
}

extern "C" int vgtkb_inter_conv_backward(int b, int n, int p, int nn, int a, int k, int ci, int co, const float* xyz,
                                         const float* sample_xyz, const int32_t* idx, const float* rot_kernels, float sigma,
                                         const float* w_kc, const void* g_hi, const void* g_lo, const float* grad_out,
                                         const void* grad_out_hi, const void* grad_out_lo, float* grad_grouped,
                                         float* grad_feats, float* grad_w, float* workspace, int mode, void* stream) {
    VGTKB_REQUIRE(conv_shape_ok(b, n, p, nn, a, k, ci, co), "inter_conv_backward: unsupported shape");
    // mode | 256: grad_feats already holds a gradient (the skip branch's) and the scatter ADDS to it -- no memset here, no
    // separate add afterwards
    const bool accumulate = (mode & 256) != 0;
    mode &= 255;
    VGTKB_REQUIRE(mode == 3 || mode == 4, "inter_conv_backward: mode %d (3 = bf16x3, 4 = single-pass bf16)", mode);
    const int fast = mode == 4;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t rows = (int64_t)b * p * a;
    const int kc = k * ci;
    int rc;
    // w_kc == NULL (prepared W^T planes in `workspace`): the workspace must not be used as split scratch
    VGTKB_REQUIRE(w_kc != nullptr || grad_w == nullptr || (grad_out_hi != nullptr && (grad_out_lo != nullptr || fast)),
                  "inter_conv_backward: prepared weight planes (w_kc == NULL) need the planes of grad_out for the weight gradient");
    if (grad_w != nullptr) {
        // dW [co, k*ci] = grad_out^T G: wide operand G straight from its planes; workspace = split of grad_out
        rc = tc_gemm_tn_planes(co, kc, rows, grad_out, grad_out_hi, grad_out_lo, nullptr, g_hi, g_lo, grad_w, 0, workspace, st, fast);
        if (rc == VGTKB_EUNSUP) set_error("inter_conv_backward: weight-gradient operands must be 16-byte aligned, rows >= 64");
        if (rc) return rc;
    }
    if (grad_feats != nullptr) {
        VGTKB_REQUIRE(grad_grouped != nullptr, "inter_conv_backward: grad_grouped scratch [rows, k*ci] needed for grad_feats");
        // dG = grad_out W  (B operand = W^T [k*ci, co]), then the scatter through the neighbourhoods
        const bool gy_planes = grad_out_hi != nullptr && (grad_out_lo != nullptr || fast) && co >= 64;
        if (w_kc == nullptr) {
            // prepared weights: `workspace` holds the bf16 planes of W^T [kc, co] (vgtkb_weight_planes) -- no transpose, no split
            rc = gy_planes ? tc_gemm_nt_planes(rows, kc, co, grad_out_hi, grad_out_lo, nullptr, nullptr, grad_grouped, workspace, st, fast)
                           : tc_gemm_nt(rows, kc, co, grad_out, nullptr, nullptr, grad_grouped, fast ? 7 : 6, workspace, st);
        } else {
        float* wt = workspace;                       // [kc, co]
        float* wsplit = workspace + (size_t)kc * co; // hi/lo split of W^T (kc*co floats)
        transpose_kernel<<<dim3(ceil_div(kc, 32), ceil_div(co, 32)), dim3(32, 8), 0, st>>>(co, kc, w_kc, wt);
        // (the planes of grad_out, when the producer wrote them, feed the contraction without conversion)
        rc = gy_planes ? tc_gemm_nt_planes(rows, kc, co, grad_out_hi, grad_out_lo, wt, nullptr, grad_grouped, wsplit, st, fast)
                       : tc_gemm_nt(rows, kc, co, grad_out, wt, nullptr, grad_grouped, fast ? 7 : 6, wsplit, st);
        }
        if (rc == VGTKB_EUNSUP) set_error("inter_conv_backward: dG contraction shape not covered (co %% 8 == 0, aligned operands)");
        if (rc) return rc;
        if (!accumulate) VGTKB_CUDA(cudaMemsetAsync(grad_feats, 0, sizeof(float) * (size_t)b * n * a * ci, st));
        rc = inter_group_backward_gen(b, n, p, nn, a, k, ci, xyz, sample_xyz, idx, rot_kernels, sigma, grad_grouped, grad_feats, fast, st);
        if (rc == VGTKB_EUNSUP)
            rc = vgtkb_inter_group_backward(b, n, p, nn, a, k, ci, xyz, sample_xyz, idx, rot_kernels, sigma, grad_grouped, grad_feats,
                                            3, stream);
        if (rc) return rc;
    }
    return check_launch("inter_conv_backward");
}
