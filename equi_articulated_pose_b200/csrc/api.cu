// api.cu -- error reporting, device check and the GEMM mode dispatch of libvgtkb200.
#include "common.cuh"

#include <string.h>

namespace vgtkb {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int sgemm_nt(int64_t M, int N, int K, const float* A, const float* B, const float* bias, float* C, cudaStream_t st);
int sgemm_tn(int M, int N, int64_t R, const float* A, const float* B, float* C, int accumulate, cudaStream_t st);
// gemm_tc.cu: return VGTKB_EUNSUP (without setting an error) when the shape is not covered
int tc_gemm_nt(int64_t M, int N, int K, const float* A, const float* B, const float* bias, float* C, int passes,
               float* workspace, cudaStream_t st);
int tc_gemm_tn(int M, int N, int64_t R, const float* A, const float* B, float* C, int accumulate, int passes,
               float* workspace, cudaStream_t st);

int tc_gemm_nt_gather(int64_t points, int anchors, int kk_n, int c_n, int N, const int32_t* table, const float* X,
                      const float* B, const float* bias, float* C, int passes, float* workspace, cudaStream_t st);
int tc_gemm_tn_gather(int64_t points, int anchors, int kk_n, int c_n, int M, const int32_t* table, const float* X,
                      const float* Y, float* C, int accumulate, int passes, float* workspace, cudaStream_t st);

}  // namespace vgtkb

using namespace vgtkb;

// contraction mode -> kernel arithmetic code: 1 = 3xTF32 (3), 2 = 1xTF32 (1), 3 = bf16x3 (6), 4 = bf16 single pass (7)
static int passes_of(int mode) { return mode == 1 ? 3 : (mode == 3 ? 6 : (mode == 4 ? 7 : 1)); }

extern "C" const char* vgtkb_last_error(void) { return g_err; }
extern "C" int vgtkb_version(void) { return VGTKB_ABI_VERSION; }

extern "C" int vgtkb_device_check(void) {
    int dev = 0;
    VGTKB_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    VGTKB_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (prop.major != 10) {
        set_error("libvgtkb200 is built for sm_100a only; device %d is sm_%d%d", dev, prop.major, prop.minor);
        return VGTKB_EUNSUP;
    }
    return VGTKB_OK;
}

extern "C" int vgtkb_gemm_nt(int64_t M, int N, int K, const float* A, const float* B, const float* bias, float* C,
                             int mode, float* workspace, void* stream) {
    VGTKB_REQUIRE(M >= 0 && N > 0 && K > 0, "gemm_nt: bad size");
    VGTKB_REQUIRE(mode >= 0 && mode <= 4, "gemm_nt: bad mode %d", mode);
    if (M == 0) return VGTKB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (mode != 0) {
        const int rc = tc_gemm_nt(M, N, K, A, B, bias, C, passes_of(mode), workspace, st);
        if (rc != VGTKB_EUNSUP) return rc;
    }
    // B == NULL: `workspace` holds the prepared bf16 planes of B (vgtkb_weight_planes) -- tensor-core path only
    VGTKB_REQUIRE(B != nullptr, "gemm_nt: B is NULL (prepared planes in `workspace`) but the shape / mode is not taken by the "
                                "bf16 tensor-core path (modes 3 / 4, K %% 8 == 0, 16-byte aligned operands)");
    return sgemm_nt(M, N, K, A, B, bias, C, st);
}

extern "C" int vgtkb_gemm_tn(int M, int N, int64_t R, const float* A, const float* B, float* C, int accumulate,
                             int mode, float* workspace, void* stream) {
    VGTKB_REQUIRE(M > 0 && N > 0 && R >= 0, "gemm_tn: bad size");
    VGTKB_REQUIRE(mode >= 0 && mode <= 4, "gemm_tn: bad mode %d", mode);
    cudaStream_t st = (cudaStream_t)stream;
    if (R == 0) {
        if (!accumulate) VGTKB_CUDA(cudaMemsetAsync(C, 0, sizeof(float) * (size_t)M * N, st));
        return VGTKB_OK;
    }
    if (mode != 0) {
        const int rc = tc_gemm_tn(M, N, R, A, B, C, accumulate, passes_of(mode), workspace, st);
        if (rc != VGTKB_EUNSUP) return rc;
    }
    return sgemm_tn(M, N, R, A, B, C, accumulate, st);
}

extern "C" int vgtkb_gather_gemm_nt(int64_t points, int anchors, int kk, int c, int n, const int32_t* table, const float* x,
                                    const float* w, const float* bias, float* out, int mode, float* workspace, void* stream) {
    VGTKB_REQUIRE(points >= 0 && anchors > 0 && kk > 0 && c > 0 && n > 0, "gather_gemm_nt: bad size");
    VGTKB_REQUIRE(mode >= 1 && mode <= 4, "gather_gemm_nt: mode %d (the gather GEMM only exists on the tensor-core path)", mode);
    if (points == 0) return VGTKB_OK;
    const int rc = tc_gemm_nt_gather(points, anchors, kk, c, n, table, x, w, bias, out, passes_of(mode), workspace,
                                     (cudaStream_t)stream);
    if (rc == VGTKB_EUNSUP) set_error("gather_gemm_nt: unsupported shape (needs c %% 64 == 0, 16-byte aligned operands)");
    return rc;
}

extern "C" int vgtkb_gather_gemm_tn(int64_t points, int anchors, int kk, int c, int m, const int32_t* table, const float* x,
                                    const float* y, float* out, int accumulate, int mode, float* workspace, void* stream) {
    VGTKB_REQUIRE(points >= 0 && anchors > 0 && kk > 0 && c > 0 && m > 0, "gather_gemm_tn: bad size");
    VGTKB_REQUIRE(mode >= 1 && mode <= 4, "gather_gemm_tn: mode %d (the gather GEMM only exists on the tensor-core path)", mode);
    const int rc = tc_gemm_tn_gather(points, anchors, kk, c, m, table, x, y, out, accumulate, passes_of(mode), workspace,
                                     (cudaStream_t)stream);
    if (rc == VGTKB_EUNSUP) set_error("gather_gemm_tn: unsupported shape (needs c %% 32 == 0, m %% 4 == 0, points >= 64)");
    return rc;
}
