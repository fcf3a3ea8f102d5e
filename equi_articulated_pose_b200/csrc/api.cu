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

}  // namespace vgtkb

using namespace vgtkb;

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
    VGTKB_REQUIRE(mode >= 0 && mode <= 3, "gemm_nt: bad mode %d", mode);
    if (M == 0) return VGTKB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (mode != 0) {
        const int rc = tc_gemm_nt(M, N, K, A, B, bias, C, mode == 1 ? 3 : (mode == 3 ? 6 : 1), workspace, st);
        if (rc != VGTKB_EUNSUP) return rc;
    }
    return sgemm_nt(M, N, K, A, B, bias, C, st);
}

extern "C" int vgtkb_gemm_tn(int M, int N, int64_t R, const float* A, const float* B, float* C, int accumulate,
                             int mode, float* workspace, void* stream) {
    VGTKB_REQUIRE(M > 0 && N > 0 && R >= 0, "gemm_tn: bad size");
    VGTKB_REQUIRE(mode >= 0 && mode <= 3, "gemm_tn: bad mode %d", mode);
    cudaStream_t st = (cudaStream_t)stream;
    if (R == 0) {
        if (!accumulate) VGTKB_CUDA(cudaMemsetAsync(C, 0, sizeof(float) * (size_t)M * N, st));
        return VGTKB_OK;
    }
    if (mode != 0) {
        const int rc = tc_gemm_tn(M, N, R, A, B, C, accumulate, mode == 1 ? 3 : (mode == 3 ? 6 : 1), workspace, st);
        if (rc != VGTKB_EUNSUP) return rc;
    }
    return sgemm_tn(M, N, R, A, B, C, accumulate, st);
}
