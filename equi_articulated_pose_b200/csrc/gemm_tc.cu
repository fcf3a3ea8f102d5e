// gemm_tc.cu -- tcgen05 (5th-gen tensor core) GEMMs for sm_100a: the anchor/kernel contraction
// BasicSO3Conv.forward (vgtk/vgtk/so3conv/modules.py:48-55, W[Co, Ci*K] @ x) and its gradients.
//
//   tc_gemm_nt : C[M,N] = A[M,K] * B[N,K]^T (+ bias)          (forward, and dX with B = W^T)
//   tc_gemm_tn : C[Mo,No] (+)= A[R,Mo]^T * B[R,No]             (weight gradient, split over R)
//
// fp32 in, fp32 out.  `passes` = 3 is the fp32-parity mode (3xTF32: operands split into a 19-bit
// head and an exact fp32 remainder, three kind::tf32 MMAs per k-step accumulate hi*hi + lo*hi +
// hi*lo in fp32 TMEM); `passes` = 1 is single-pass TF32.
//
// Kernel shape (both): persistent CTAs (one per SM), warp-specialised:
//   warps 0-3   epilogue: tcgen05.ld accumulator rows -> registers -> global
//   warp  4     one elected thread issues tcgen05.mma; owns the TMEM allocation
//   warps 5-12  operand producers: coalesced float4 loads of the activation operand, hi/lo split in
//               registers, st.shared into the 128B-swizzled K-major (nt) / MN-major (tn) canonical
//               layout; the weight operand of the nt kernel arrives pre-split through TMA
// mbarrier rings: full/empty per smem stage, full/empty per TMEM accumulator (double buffered so the
// epilogue of tile i overlaps the MMAs of tile i+1).
#include "tc_common.cuh"

#include <cuda.h>  // CUtensorMap types only; the encoder is fetched through cudaGetDriverEntryPoint

namespace vgtkb {

using namespace tc;

constexpr int TC_BM = 128;         // rows of A per tile = UMMA M = TMEM lanes
constexpr int TC_BK = 32;          // fp32 per k-block = one 128-byte swizzle row
constexpr int TC_EPI_WARPS = 4;
constexpr int TC_PROD_WARPS = 8;
constexpr int TC_THREADS = (TC_EPI_WARPS + 1 + TC_PROD_WARPS) * 32;  // 416
constexpr int TC_SMEM_BUDGET = 200 * 1024;

template <int BN>
struct NtCfg {
    static constexpr int A_BYTES = TC_BM * TC_BK * 4;  // 16 KB (one of hi / lo)
    static constexpr int B_BYTES = BN * TC_BK * 4;
    static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
    static constexpr int STAGES = (TC_SMEM_BUDGET / STAGE_BYTES) > 6 ? 6 : (TC_SMEM_BUDGET / STAGE_BYTES);
    static constexpr int TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;  // two accumulators
    static_assert(STAGES >= 2, "need at least two smem stages");
};

// ---------------------------------------------------------------------------------- NT kernel
template <int BN>
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_gemm_nt_kernel(const __grid_constant__ CUtensorMap map_bhi, const __grid_constant__ CUtensorMap map_blo,
                  const float* __restrict__ A, const float* __restrict__ bias, float* __restrict__ C, int64_t M, int N,
                  int K, int passes) {
    using Cfg = NtCfg<BN>;
    constexpr int STAGES = Cfg::STAGES;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    // the dynamic window is only guaranteed 16-byte aligned: round up to the 1024 B the swizzle needs
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));

    __shared__ __align__(8) uint64_t full_bar[STAGES];
    __shared__ __align__(8) uint64_t empty_bar[STAGES];
    __shared__ __align__(8) uint64_t tfull_bar[2];
    __shared__ __align__(8) uint64_t tempty_bar[2];
    __shared__ uint32_t tmem_base_slot;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m_tiles = (int)((M + TC_BM - 1) / TC_BM);
    const int n_tiles = (N + BN - 1) / BN;
    const int total_tiles = m_tiles * n_tiles;
    const int nkb = (K + TC_BK - 1) / TC_BK;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], TC_PROD_WARPS + 1);  // 8 producer warps + the TMA issuer's expect_tx arrive
            mbar_init(&empty_bar[s], 1);                 // tcgen05.commit
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tfull_bar[a], 1);                     // tcgen05.commit
            mbar_init(&tempty_bar[a], TC_EPI_WARPS * 32);    // every epilogue thread
        }
        fence_barrier_init();
    }
    if (warp == TC_EPI_WARPS) {
        tmem_alloc(&tmem_base_slot, Cfg::TMEM_COLS);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_slot;

    if (warp < TC_EPI_WARPS) {
        // ============================ epilogue ============================
        int t = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++t) {
            const int mt = tile / n_tiles, nt = tile % n_tiles;
            const int acc = t & 1;
            mbar_wait_guard(&tfull_bar[acc], (t >> 1) & 1);
            tc_fence_after();
            const int64_t row = (int64_t)mt * TC_BM + warp * 32 + lane;
            const int n0 = nt * BN;
            const bool vec_ok = (N % 4 == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0);
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 32) {
                if (n0 + c0 >= N) break;  // warp-uniform
                float v[32];
                tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(acc * BN + c0), v);
                if (row < M) {
                    float* dst = C + row * N + n0 + c0;
                    if (vec_ok && n0 + c0 + 32 <= N) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            float4 o = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                            if (bias != nullptr) {
                                const float4 bb = *reinterpret_cast<const float4*>(bias + n0 + c0 + j);
                                o.x += bb.x; o.y += bb.y; o.z += bb.z; o.w += bb.w;
                            }
                            *reinterpret_cast<float4*>(dst + j) = o;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (n0 + c0 + j < N) dst[j] = v[j] + (bias != nullptr ? bias[n0 + c0 + j] : 0.f);
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(&tempty_bar[acc]);
        }
    } else if (warp == TC_EPI_WARPS) {
        // ============================ MMA issuer ============================
        if (lane == 0) {
            const uint32_t idesc = make_idesc_tf32(TC_BM, BN, 0, 0);
            int it = 0, t = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++t) {
                const int acc = t & 1;
                mbar_wait_guard(&tempty_bar[acc], ((t >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int s = it % STAGES;
                    mbar_wait_guard(&full_bar[s], (it / STAGES) & 1);
                    tc_fence_after();
                    const uint32_t a_hi = smem_base + s * Cfg::STAGE_BYTES;
                    const uint32_t a_lo = a_hi + Cfg::A_BYTES;
                    const uint32_t b_hi = a_lo + Cfg::A_BYTES;
                    const uint32_t b_lo = b_hi + Cfg::B_BYTES;
                    const int krem = K - kb * TC_BK;
                    const int ksteps = krem >= TC_BK ? TC_BK / 8 : (krem + 7) / 8;
                    for (int ks = 0; ks < ksteps; ++ks) {
                        const uint32_t koff = ks * 32;  // 8 tf32 = 32 bytes inside the 128 B swizzle row
                        const uint64_t da_hi = make_smem_desc(a_hi + koff, 16, 1024);
                        const uint64_t db_hi = make_smem_desc(b_hi + koff, 16, 1024);
                        const uint32_t first = (kb | ks) != 0;
                        if (passes == 3) {
                            const uint64_t da_lo = make_smem_desc(a_lo + koff, 16, 1024);
                            const uint64_t db_lo = make_smem_desc(b_lo + koff, 16, 1024);
                            umma_tf32(d_tmem, da_lo, db_hi, idesc, first);   // small terms first
                            umma_tf32(d_tmem, da_hi, db_lo, idesc, 1);
                            umma_tf32(d_tmem, da_hi, db_hi, idesc, 1);
                        } else {
                            umma_tf32(d_tmem, da_hi, db_hi, idesc, first);
                        }
                    }
                    umma_commit(&empty_bar[s]);   // smem stage reusable once these MMAs have read it
                }
                umma_commit(&tfull_bar[acc]);     // accumulator complete
            }
        }
        __syncwarp();
    } else {
        // ============================ producers ============================
        const int pw = warp - TC_EPI_WARPS - 1;             // 0..7
        const int pt = pw * 32 + lane;                      // 0..255
        const int prow = pt >> 3;                           // 0..31 (+32*i)
        const int chunk = pt & 7;                           // 16-byte chunk inside the 128 B row
        const bool tma_thread = (pw == 0 && lane == 0);
        if (tma_thread) {
            tma_prefetch_desc(&map_bhi);
            tma_prefetch_desc(&map_blo);
        }
        const uint32_t b_tx_bytes = (uint32_t)Cfg::B_BYTES * (passes == 3 ? 2u : 1u);
        // swizzled offsets of this thread's four 16-byte slots inside an A tile
        uint32_t a_off[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = prow + 32 * i;
            a_off[i] = (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((chunk ^ (r & 7)) << 4));
        }
        const int kcol = chunk * 4;
        int it = 0;
        float4 cur[4];
        auto load_a = [&](int tile, int kb, float4 (&dst)[4]) {
            const int mt = tile / n_tiles;
            const int k = kb * TC_BK + kcol;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int64_t row = (int64_t)mt * TC_BM + prow + 32 * i;
                dst[i] = (row < M && k < K) ? __ldg(reinterpret_cast<const float4*>(A + row * K + k))
                                            : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };
        int tile = blockIdx.x;
        if (tile < total_tiles) load_a(tile, 0, cur);
        while (tile < total_tiles) {
            const int nt = tile % n_tiles;
            for (int kb = 0; kb < nkb; ++kb, ++it) {
                // prefetch the next k-block (or the first of the next tile) into registers
                float4 nxt[4];
                int ntile = tile, nkb_i = kb + 1;
                if (nkb_i == nkb) { ntile = tile + gridDim.x; nkb_i = 0; }
                const bool have_next = ntile < total_tiles;
                if (have_next) load_a(ntile, nkb_i, nxt);

                const int s = it % STAGES;
                mbar_wait_guard(&empty_bar[s], ((it / STAGES) & 1) ^ 1);
                const uint32_t a_hi = smem_base + s * Cfg::STAGE_BYTES;
                const uint32_t a_lo = a_hi + Cfg::A_BYTES;
                if (tma_thread) {
                    unsigned char* st = smem_al + (size_t)s * Cfg::STAGE_BYTES;
                    mbar_arrive_expect_tx(&full_bar[s], b_tx_bytes);
                    tma_load_2d(st + 2 * Cfg::A_BYTES, &map_bhi, kb * TC_BK, nt * BN, &full_bar[s]);
                    if (passes == 3) tma_load_2d(st + 2 * Cfg::A_BYTES + Cfg::B_BYTES, &map_blo, kb * TC_BK, nt * BN, &full_bar[s]);
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float4 hi, lo;
                    split_tf32(cur[i], hi, lo);
                    st_shared_v4(a_hi + a_off[i], hi);
                    if (passes == 3) st_shared_v4(a_lo + a_off[i], lo);
                }
                fence_proxy_async();   // generic-proxy stores -> visible to the tensor-core (async) proxy
                __syncwarp();
                if (lane == 0) mbar_arrive(&full_bar[s]);
                if (have_next) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) cur[i] = nxt[i];
                }
            }
            tile += gridDim.x;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == TC_EPI_WARPS) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

// hi/lo split of the (small) weight operand into a workspace [2][n]
__global__ void split_tf32_kernel(int64_t n, const float* __restrict__ x, float* __restrict__ hi, float* __restrict__ lo) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float v = x[i];
        const float h = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
        hi[i] = h;
        lo[i] = v - h;
    }
}

// ---------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encoder() {
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// row-major fp32 matrix [rows, cols]; box = [box_rows, 32 floats], 128-byte swizzle, zero fill out of bounds
static int make_map_2d(CUtensorMap* map, const float* base, int64_t rows, int64_t cols, int box_rows) {
    EncodeTiledFn enc = get_encoder();
    if (enc == nullptr) {
        set_error("cuTensorMapEncodeTiled not available from the driver");
        return VGTKB_EUNSUP;
    }
    const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t gstr[1] = {(cuuint64_t)cols * 4};
    const cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstr, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d)", (int)r);
        return VGTKB_ECUDA;
    }
    return VGTKB_OK;
}

static int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = kNumSMs;
    }
    return n;
}

template <int BN>
static int launch_nt(int64_t M, int N, int K, const float* A, const float* Bhi, const float* Blo, const float* bias,
                     float* C, int passes, cudaStream_t st) {
    using Cfg = NtCfg<BN>;
    CUtensorMap mhi, mlo;
    int rc = make_map_2d(&mhi, Bhi, N, K, BN);
    if (rc) return rc;
    rc = make_map_2d(&mlo, Blo, N, K, BN);
    if (rc) return rc;
    const size_t smem = (size_t)Cfg::STAGES * Cfg::STAGE_BYTES + 1024;
    auto kern = tc_gemm_nt_kernel<BN>;
    VGTKB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t tiles = ceil_div64(M, TC_BM) * ceil_div(N, BN);
    const int grid = (int)(tiles < num_sms() ? tiles : num_sms());
    kern<<<grid, TC_THREADS, smem, st>>>(mhi, mlo, A, bias, C, M, N, K, passes);
    return check_launch("gemm_nt(tcgen05)");
}

int tc_gemm_nt(int64_t M, int N, int K, const float* A, const float* B, const float* bias, float* C, int passes,
               float* workspace, cudaStream_t st) {
    // shapes the tensor-core path takes; everything else falls back to the FFMA kernel
    if (K % 4 != 0 || K < 8 || M < 1 || ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B)) & 15) != 0)
        return VGTKB_EUNSUP;
    if (bias != nullptr && (reinterpret_cast<uintptr_t>(bias) & 15) != 0) return VGTKB_EUNSUP;
    if (M >= (int64_t)1 << 31) return VGTKB_EUNSUP;
    const int64_t nb = (int64_t)N * K;
    float* ws = nullptr;
    float* owned = nullptr;
    const float* Bhi = B;
    const float* Blo = B;
    if (passes == 3) {
        ws = workspace;
        if (ws == nullptr || (reinterpret_cast<uintptr_t>(ws) & 15) != 0 || nb % 4 != 0) {
            VGTKB_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&owned), sizeof(float) * 2 * (size_t)nb, st));
            ws = owned;
        }
        const int blocks = (int)(ceil_div64(nb, 256) < 1184 ? ceil_div64(nb, 256) : 1184);
        split_tf32_kernel<<<blocks, 256, 0, st>>>(nb, B, ws, ws + nb);
        Bhi = ws;
        Blo = ws + nb;
    }
    int rc;
    if (N <= 64) rc = launch_nt<64>(M, N, K, A, Bhi, Blo, bias, C, passes, st);
    else if (N <= 128) rc = launch_nt<128>(M, N, K, A, Bhi, Blo, bias, C, passes, st);
    else rc = launch_nt<256>(M, N, K, A, Bhi, Blo, bias, C, passes, st);
    if (owned != nullptr) cudaFreeAsync(owned, st);
    return rc;
}

// ---------------------------------------------------------------------------------- TN kernel
// T[i, j] = sum_r P[r, p0+i] * Q[r, q0+j]   (i < 128, j < BN), both operands MN-major in shared
// memory: canonical SWIZZLE_128B_BASE32B atoms (the layout MN-major tf32 operands require) of
// 32 (MN) x 4 (K) fp32 = 512 B, 32-byte chunks XOR-permuted by the k-row; stage layout
// [k-atom (8)][mn-atom][512 B], so LBO (next MN atom) = 512 B and SBO (next k-atom) = MN/32 * 512 B.
// Work item = (R slice, P tile, Q tile); partial tiles are added into C with red.global.add.
//   C[(q0+j) * ldc + (p0+i)] += T[i, j]      (lanes run along i: coalesced)
template <int BN>
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_gemm_tn_kernel(const float* __restrict__ P, int Pw, const float* __restrict__ Q, int Qw, float* __restrict__ C, int ldc,
                  int64_t R, int64_t rows_per_split, int splits, int passes) {
    using Cfg = NtCfg<BN>;
    constexpr int STAGES = Cfg::STAGES;
    constexpr uint32_t P_SBO = (TC_BM / 32) * 512;   // bytes between 4-row k-atoms of the P tile
    constexpr uint32_t Q_SBO = (BN / 32) * 512;
    constexpr uint32_t MN_LBO = 512;
    constexpr uint32_t L32 = 1;                       // SWIZZLE_128B_BASE32B
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;

    __shared__ __align__(8) uint64_t full_bar[STAGES];
    __shared__ __align__(8) uint64_t empty_bar[STAGES];
    __shared__ __align__(8) uint64_t tfull_bar[2];
    __shared__ __align__(8) uint64_t tempty_bar[2];
    __shared__ uint32_t tmem_base_slot;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int p_tiles = (Pw + TC_BM - 1) / TC_BM;
    const int q_tiles = (Qw + BN - 1) / BN;
    const int tiles = p_tiles * q_tiles;
    const int items = tiles * splits;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], TC_PROD_WARPS);
            mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tfull_bar[a], 1);
            mbar_init(&tempty_bar[a], TC_EPI_WARPS * 32);
        }
        fence_barrier_init();
    }
    if (warp == TC_EPI_WARPS) tmem_alloc(&tmem_base_slot, Cfg::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_slot;

    auto item_rows = [&](int item, int64_t& r0, int64_t& r1) {
        const int sp = item / tiles;
        r0 = (int64_t)sp * rows_per_split;
        r1 = r0 + rows_per_split < R ? r0 + rows_per_split : R;
    };

    if (warp < TC_EPI_WARPS) {
        // ============================ epilogue ============================
        int t = 0;
        for (int item = blockIdx.x; item < items; item += gridDim.x, ++t) {
            const int tile = item % tiles;
            const int p0 = (tile / q_tiles) * TC_BM, q0 = (tile % q_tiles) * BN;
            const int acc = t & 1;
            mbar_wait_guard(&tfull_bar[acc], (t >> 1) & 1);
            tc_fence_after();
            const int i = p0 + warp * 32 + lane;
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 32) {
                if (q0 + c0 >= Qw) break;  // warp-uniform
                float v[32];
                tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(acc * BN + c0), v);
                if (i < Pw) {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (q0 + c0 + j < Qw) atomicAdd(C + (size_t)(q0 + c0 + j) * ldc + i, v[j]);
                }
            }
            tc_fence_before();
            mbar_arrive(&tempty_bar[acc]);
        }
    } else if (warp == TC_EPI_WARPS) {
        // ============================ MMA issuer ============================
        if (lane == 0) {
            const uint32_t idesc = make_idesc_tf32(TC_BM, BN, 1, 1);   // both operands MN-major
            int it = 0, t = 0;
            for (int item = blockIdx.x; item < items; item += gridDim.x, ++t) {
                int64_t r0, r1;
                item_rows(item, r0, r1);
                const int nkb = (int)((r1 - r0 + TC_BK - 1) / TC_BK);
                const int acc = t & 1;
                mbar_wait_guard(&tempty_bar[acc], ((t >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int s = it % STAGES;
                    mbar_wait_guard(&full_bar[s], (it / STAGES) & 1);
                    tc_fence_after();
                    const uint32_t p_hi = smem_base + s * Cfg::STAGE_BYTES;
                    const uint32_t p_lo = p_hi + Cfg::A_BYTES;
                    const uint32_t q_hi = p_lo + Cfg::A_BYTES;
                    const uint32_t q_lo = q_hi + Cfg::B_BYTES;
                    const int64_t rrem = r1 - r0 - (int64_t)kb * TC_BK;
                    const int ksteps = rrem >= TC_BK ? TC_BK / 8 : (int)((rrem + 7) / 8);
                    for (int ks = 0; ks < ksteps; ++ks) {
                        // one MMA (K = 8) spans two k-atoms
                        const uint64_t dp_hi = make_smem_desc(p_hi + ks * 2 * P_SBO, MN_LBO, P_SBO, L32);
                        const uint64_t dq_hi = make_smem_desc(q_hi + ks * 2 * Q_SBO, MN_LBO, Q_SBO, L32);
                        const uint32_t first = (kb | ks) != 0;
                        if (passes == 3) {
                            const uint64_t dp_lo = make_smem_desc(p_lo + ks * 2 * P_SBO, MN_LBO, P_SBO, L32);
                            const uint64_t dq_lo = make_smem_desc(q_lo + ks * 2 * Q_SBO, MN_LBO, Q_SBO, L32);
                            umma_tf32(d_tmem, dp_lo, dq_hi, idesc, first);
                            umma_tf32(d_tmem, dp_hi, dq_lo, idesc, 1);
                            umma_tf32(d_tmem, dp_hi, dq_hi, idesc, 1);
                        } else {
                            umma_tf32(d_tmem, dp_hi, dq_hi, idesc, first);
                        }
                    }
                    umma_commit(&empty_bar[s]);
                }
                umma_commit(&tfull_bar[acc]);
            }
        }
        __syncwarp();
    } else {
        // ============================ producers ============================
        const int pw = warp - TC_EPI_WARPS - 1;
        const int pt = pw * 32 + lane;         // 0..255
        const int krow = pt >> 3;              // 0..31: row of R inside the k-block
        const int sub = pt & 7;                // 16-byte chunk inside a 128 B atom row
        // inside an atom: row (krow & 3) of 128 B; the 32-byte chunk index is XORed with the row
        const uint32_t row_off = (uint32_t)((krow & 3) * 128 + ((((sub >> 1) ^ (krow & 3)) << 5) | ((sub & 1) << 4)));
        const uint32_t p_koff = (uint32_t)(krow >> 2) * P_SBO + row_off;
        const uint32_t q_koff = (uint32_t)(krow >> 2) * Q_SBO + row_off;
        int it = 0;
        for (int item = blockIdx.x; item < items; item += gridDim.x) {
            const int tile = item % tiles;
            const int p0 = (tile / q_tiles) * TC_BM, q0 = (tile % q_tiles) * BN;
            int64_t r0, r1;
            item_rows(item, r0, r1);
            const int nkb = (int)((r1 - r0 + TC_BK - 1) / TC_BK);
            for (int kb = 0; kb < nkb; ++kb, ++it) {
                const int64_t r = r0 + (int64_t)kb * TC_BK + krow;
                const bool rok = r < r1;
                float4 pv[TC_BM / 32], qv[BN / 32];
#pragma unroll
                for (int i = 0; i < TC_BM / 32; ++i) {
                    const int col = p0 + i * 32 + sub * 4;
                    pv[i] = (rok && col < Pw) ? __ldg(reinterpret_cast<const float4*>(P + r * Pw + col))
                                              : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int i = 0; i < BN / 32; ++i) {
                    const int col = q0 + i * 32 + sub * 4;
                    qv[i] = (rok && col < Qw) ? __ldg(reinterpret_cast<const float4*>(Q + r * Qw + col))
                                              : make_float4(0.f, 0.f, 0.f, 0.f);
                }
                const int s = it % STAGES;
                mbar_wait_guard(&empty_bar[s], ((it / STAGES) & 1) ^ 1);
                const uint32_t p_hi = smem_base + s * Cfg::STAGE_BYTES;
                const uint32_t p_lo = p_hi + Cfg::A_BYTES;
                const uint32_t q_hi = p_lo + Cfg::A_BYTES;
                const uint32_t q_lo = q_hi + Cfg::B_BYTES;
#pragma unroll
                for (int i = 0; i < TC_BM / 32; ++i) {
                    float4 hi, lo;
                    split_tf32(pv[i], hi, lo);
                    st_shared_v4(p_hi + p_koff + i * MN_LBO, hi);
                    if (passes == 3) st_shared_v4(p_lo + p_koff + i * MN_LBO, lo);
                }
#pragma unroll
                for (int i = 0; i < BN / 32; ++i) {
                    float4 hi, lo;
                    split_tf32(qv[i], hi, lo);
                    st_shared_v4(q_hi + q_koff + i * MN_LBO, hi);
                    if (passes == 3) st_shared_v4(q_lo + q_koff + i * MN_LBO, lo);
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(&full_bar[s]);
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == TC_EPI_WARPS) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

template <int BN>
static int launch_tn(const float* P, int Pw, const float* Q, int Qw, float* C, int ldc, int64_t R, int passes,
                     cudaStream_t st) {
    using Cfg = NtCfg<BN>;
    const int tiles = ceil_div(Pw, TC_BM) * ceil_div(Qw, BN);
    int64_t splits = ceil_div64((int64_t)2 * num_sms(), tiles);
    const int64_t max_splits = ceil_div64(R, 512);
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    int64_t rps = ceil_div64(ceil_div64(R, splits), TC_BK) * TC_BK;
    splits = ceil_div64(R, rps);
    const int64_t items = splits * tiles;
    const size_t smem = (size_t)Cfg::STAGES * Cfg::STAGE_BYTES + 1024;
    auto kern = tc_gemm_tn_kernel<BN>;
    VGTKB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = (int)(items < num_sms() ? items : num_sms());
    kern<<<grid, TC_THREADS, smem, st>>>(P, Pw, Q, Qw, C, ldc, R, rps, (int)splits, passes);
    return check_launch("gemm_tn(tcgen05)");
}

// C[M,N] (+)= A[R,M]^T B[R,N]:  P = B (tiles of 128 over N), Q = A (tiles of <= 256 over M)
int tc_gemm_tn(int M, int N, int64_t R, const float* A, const float* B, float* C, int accumulate, int passes,
               cudaStream_t st) {
    if (M % 4 != 0 || N % 4 != 0 || R < 64 || ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B)) & 15) != 0)
        return VGTKB_EUNSUP;
    if ((int64_t)M * N < 64 * 64 / 4) return VGTKB_EUNSUP;   // tiny outputs: the FFMA split-R kernel is fine
    if (!accumulate) VGTKB_CUDA(cudaMemsetAsync(C, 0, sizeof(float) * (size_t)M * N, st));
    if (M <= 64) return launch_tn<64>(B, N, A, M, C, N, R, passes, st);
    if (M <= 128) return launch_tn<128>(B, N, A, M, C, N, R, passes, st);
    return launch_tn<256>(B, N, A, M, C, N, R, passes, st);
}

}  // namespace vgtkb
