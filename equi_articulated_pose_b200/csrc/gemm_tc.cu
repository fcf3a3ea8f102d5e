// gemm_tc.cu -- tcgen05 (5th-gen tensor core) GEMMs for sm_100a.  [bring-up stub: see below]
#include "common.cuh"

namespace vgtkb {

int tc_gemm_nt(int64_t, int, int, const float*, const float*, const float*, float*, int, cudaStream_t) {
    return VGTKB_EUNSUP;
}
int tc_gemm_tn(int, int, int64_t, const float*, const float*, float*, int, int, cudaStream_t) { return VGTKB_EUNSUP; }

}  // namespace vgtkb
