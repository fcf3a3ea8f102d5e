// gemm_tc.cu -- tcgen05 (5th-gen tensor core) GEMMs for sm_100a: the anchor/kernel contraction
// BasicSO3Conv.forward (vgtk/vgtk/so3conv/modules.py:48-55, W[Co, Ci*K] @ x) and its gradients.
//
//   tc_gemm_nt : C[M,N] = A[M,K] * B[N,K]^T (+ bias)          (forward, and dX with B = W^T)
//   tc_gemm_tn : C[Mo,No] (+)= A[R,Mo]^T * B[R,No]             (weight gradient, split over R)
//
// fp32 in, fp32 out.  `passes` = 3 is the fp32-parity mode (3xTF32: operands split into a 19-bit
// head and an exact fp32 remainder, three kind::tf32 MMAs per k-step give hi*hi + lo*hi + hi*lo);
// `passes` = 1 is single-pass TF32.
//
// The TMEM accumulator rounds toward zero on every accumulate (measured: a shrink proportional to
// the number of MMAs).  So the tensor core only ever accumulates a short CHUNK of the reduction
// (chunk_kb k-blocks of 32) in TMEM; the epilogue warps drain every chunk with tcgen05.ld and keep
// the running sum in fp32 registers with round-to-nearest adds.  Two TMEM buffers let chunk c+1
// accumulate while chunk c is drained.
//
// Kernel shape (both kernels): persistent CTAs, one per SM, 14 warps, warp-specialised, register
// file re-partitioned with setmaxnreg:
//   warps 0-7   epilogue (lane quadrant = warp % 4, column half = warp / 4): tcgen05.ld + FADD into
//               BN/2 register accumulators per thread; stores / red.global.add at the end of a tile
//   warps 8-11  converters: hi/lo split of the raw fp32 tile the TMA delivered, in place in shared memory
//   warp  12    one elected thread issues tcgen05.mma; owns the TMEM allocation
//   warp  13    one elected thread issues the TMA loads (cp.async.bulk.tensor.2d)
// mbarrier rings: raw_full (TMA landed) / full (converted) / empty per smem stage, full / empty per
// TMEM buffer.
#include "tc_common.cuh"

#include <stdlib.h>

#include <cuda.h>  // CUtensorMap types only; the encoder is fetched through cudaGetDriverEntryPoint

namespace vgtkb {

using namespace tc;

constexpr int TC_BM = 128;  // rows of A per tile = UMMA M = TMEM lanes
constexpr int TC_BK = 32;   // fp32 per k-block = one 128-byte swizzle row
// 16 warps = 4 complete warpgroups (setmaxnreg is a warpgroup-wide .sync.aligned instruction: a partial
// warpgroup never completes it).  warps 0-7 epilogue, 8-11 and 14-15 converters, 12 MMA, 13 TMA.
constexpr int TC_EPI_WARPS = 8;
constexpr int TC_CONV_WARPS = 6;
constexpr int TC_MMA_WARP = 12;
constexpr int TC_TMA_WARP = 13;
constexpr int TC_THREADS = 16 * 32;                        // 512
constexpr int TC_CONV_THREADS = TC_CONV_WARPS * 32;
constexpr int TC_SMEM_BUDGET = 200 * 1024;

// register budget after setmaxnreg: 256 x 184 + 256 x 64 = 63488 <= 512 x 128 (the pool a CTA can re-partition
// is what it was launched with: threads x the ptxas register count, 128 here)
// (the pool a CTA can re-partition is what it was launched with: threads x ptxas register count)
#define VGTKB_REG_EPI 184
#define VGTKB_REG_OTHER 64
#define VGTKB_STR2(x) #x
#define VGTKB_STR(x) VGTKB_STR2(x)

__device__ __forceinline__ void reg_inc_epi() {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 " VGTKB_STR(VGTKB_REG_EPI) ";" ::: "memory");
}
__device__ __forceinline__ void reg_dec_other() {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 " VGTKB_STR(VGTKB_REG_OTHER) ";" ::: "memory");
}
__device__ __forceinline__ float4 ld_shared_v4(uint32_t saddr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr) : "memory");
    return v;
}

template <int BN>
struct TcCfg {
    static constexpr int A_BYTES = TC_BM * TC_BK * 4;  // 16 KB (one of hi / lo)
    static constexpr int B_BYTES = BN * TC_BK * 4;
    static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
    static constexpr int STAGES = (TC_SMEM_BUDGET / STAGE_BYTES) > 6 ? 6 : (TC_SMEM_BUDGET / STAGE_BYTES);
    static constexpr int TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;  // two chunk buffers
    static constexpr int HALF = BN / 2;                          // accumulator columns per epilogue thread
    static_assert(STAGES >= 2, "need at least two smem stages");
    static_assert(BN % 64 == 0, "BN must be a multiple of 64");
};

// Gather-GEMM addressing (intra-anchor convolution, vgtk/vgtk/so3conv/functional.py:2553-2567 fused into the
// contraction): the activation operand is never materialised.  X is [points, anchors, c]; a tile owns ONE
// output anchor `an` and 128 (64) consecutive points, so the k-block (kk, c0..) of its operand is the regular
// box X[points, table[an*kk_n + kk], c0..] -- one 3-D TMA load with the middle coordinate looked up.
struct TcGather {
    int anchors;            // 0: plain GEMM
    int kk_n;               // neighbours per anchor (12)
    int c;                  // channels of X
    const int32_t* table;   // [anchors, kk_n]
};

struct TcBarriers {
    uint64_t raw_full[6], full[6], empty[6], tfull[2], tempty[2];
    uint32_t tmem_base;
};

template <int STAGES>
__device__ __forceinline__ void tc_init_barriers(TcBarriers& b, int raw_count = 1) {
    for (int s = 0; s < STAGES; ++s) {
        mbar_init(&b.raw_full[s], raw_count);    // the TMA issuers' expect_tx arrives (one per issuing warp)
        mbar_init(&b.full[s], TC_CONV_WARPS);    // one elected arrive per converter warp
        mbar_init(&b.empty[s], 1);               // tcgen05.commit
    }
    for (int a = 0; a < 2; ++a) {
        mbar_init(&b.tfull[a], 1);                     // tcgen05.commit
        mbar_init(&b.tempty[a], TC_EPI_WARPS * 32);    // every epilogue thread
    }
    fence_barrier_init();
}

// converters: split `n16` 16-byte slots starting at `raw` in place (hi) and into raw + lo_off (lo)
__device__ __forceinline__ void convert_region(uint32_t raw, uint32_t lo_off, int n16, int ct) {
#pragma unroll 4
    for (int i = ct; i < n16; i += TC_CONV_THREADS) {
        const uint32_t a = raw + (uint32_t)i * 16u;
        const float4 v = ld_shared_v4(a);
        float4 hi, lo;
        split_tf32(v, hi, lo);
        st_shared_v4(a, hi);
        st_shared_v4(a + lo_off, lo);
    }
}

// bf16x3 converter for a K-major tile: the raw fp32 k-block (128 rows x 64 columns, two TMA boxes of
// 16 KB) becomes, IN PLACE, the bf16 hi tile (first 16 KB) and lo tile (second 16 KB), both 128 rows x
// 128 B in the 128B-swizzled K-major layout.  Row r of both output tiles only depends on row r of the
// two boxes, and the eight lanes that own a row sit in one warp: load, convert, __syncwarp, store.
__device__ __forceinline__ void convert_rows_bf16(uint32_t stage, int cw, int lane) {
    const int j = lane & 7, sub = lane >> 3;
    for (int g = cw; g < TC_BM / 4; g += TC_CONV_WARPS) {
        const int r = g * 4 + sub;
        const uint32_t rowbase = stage + (uint32_t)((r >> 3) * 1024 + (r & 7) * 128);
        const uint32_t src = rowbase + (uint32_t)(j >> 2) * 16384u;
        const int c0 = 2 * (j & 3);
        const float4 v0 = ld_shared_v4(src + (uint32_t)(((c0) ^ (r & 7)) << 4));
        const float4 v1 = ld_shared_v4(src + (uint32_t)(((c0 + 1) ^ (r & 7)) << 4));
        uint4 hi, lo;
        split_bf16x8(v0, v1, hi, lo);
        __syncwarp();
        const uint32_t dst = rowbase + (uint32_t)((j ^ (r & 7)) << 4);
        st_shared_u4(dst, hi);
        st_shared_u4(dst + 16384u, lo);
    }
}

// ---------------------------------------------------------------------------------- NT kernel
// BF = false: kind::tf32 (passes 1 or 3), k-block = 32 columns.  BF = true: bf16x3 (kind::f16), k-block = 64.
template <int BN, bool BF>
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_gemm_nt_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_bhi,
                  const __grid_constant__ CUtensorMap map_blo, const float* __restrict__ bias, float* __restrict__ C,
                  int64_t M, int N, int K, int passes, int chunk_kb, TcGather ga) {
    using Cfg = TcCfg<BN>;
    constexpr int STAGES = Cfg::STAGES;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
    __shared__ __align__(8) TcBarriers bars;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m_tiles = (int)((M + TC_BM - 1) / TC_BM);
    const int n_tiles = (N + BN - 1) / BN;
    const int a_cnt = ga.anchors > 0 ? ga.anchors : 1;      // gather: M counts points, a tile owns one anchor
    const int total_tiles = m_tiles * a_cnt * n_tiles;      // tile -> (mt, an, nt), nt fastest
    constexpr int KBE = BF ? 64 : TC_BK;     // K elements per k-block (one 128-byte operand row)
    constexpr int UK = BF ? 16 : 8;          // K elements per MMA
    const int nkb = (K + KBE - 1) / KBE;
    // every CTA walks the reduction dimension from a different starting k-block: all SMs stream the SAME weight
    // operand, and in lockstep they would hammer the same few L2 slices (the sum does not care about the order)
    const int krot = (int)(((int64_t)blockIdx.x * nkb) / gridDim.x);

    if (threadIdx.x == 0) tc_init_barriers<STAGES>(bars);
    if (warp == TC_MMA_WARP) tmem_alloc(&bars.tmem_base, Cfg::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars.tmem_base;

    if (warp < TC_EPI_WARPS) {
        // ============================ epilogue: drain chunks, RN accumulate in registers ============
        reg_inc_epi();
        const int q = warp & 3, h = warp >> 2;
        float acc[Cfg::HALF];
#pragma unroll
        for (int i = 0; i < Cfg::HALF; ++i) acc[i] = 0.f;
        int ci = 0;
        const bool vec_ok = (N % 4 == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0);
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const int nt = tile % n_tiles, an = (tile / n_tiles) % a_cnt, mt = tile / (n_tiles * a_cnt);
            for (int kb0 = 0; kb0 < nkb; kb0 += chunk_kb, ++ci) {
                const int buf = ci & 1;
                mbar_wait_guard(&bars.tfull[buf], (ci >> 1) & 1);
                tc_fence_after();
#pragma unroll
                for (int j = 0; j < Cfg::HALF / 32; ++j) {
                    float v[32];
                    tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN + h * Cfg::HALF + j * 32), v);
#pragma unroll
                    for (int i = 0; i < 32; ++i) acc[j * 32 + i] += v[i];
                }
                tc_fence_before();
                mbar_arrive(&bars.tempty[buf]);
            }
            // transpose each 32x32 block through a private 4 KB staging tile so that a store instruction
            // covers 4 rows x 128 contiguous bytes (the accumulator layout is one row per lane)
            const uint32_t stg = smem_base + (uint32_t)(STAGES * Cfg::STAGE_BYTES) + (uint32_t)warp * 4096u;
            const int64_t row0 = (int64_t)mt * TC_BM + q * 32;
            const int c_base = nt * BN + h * Cfg::HALF;
#pragma unroll
            for (int j = 0; j < Cfg::HALF / 32; ++j) {
#pragma unroll
                for (int c4 = 0; c4 < 8; ++c4)
                    st_shared_v4(stg + (uint32_t)(lane * 128 + ((c4 ^ (lane & 7)) << 4)),
                                 make_float4(acc[j * 32 + c4 * 4], acc[j * 32 + c4 * 4 + 1], acc[j * 32 + c4 * 4 + 2],
                                             acc[j * 32 + c4 * 4 + 3]));
                __syncwarp();
                const int c4 = lane & 7;
                const int col = c_base + j * 32 + c4 * 4;
                float bb[4] = {0.f, 0.f, 0.f, 0.f};
                if (bias != nullptr) {
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        if (col + e < N) bb[e] = bias[col + e];
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int rr = i * 4 + (lane >> 3);
                    float4 o = ld_shared_v4(stg + (uint32_t)(rr * 128 + ((c4 ^ (rr & 7)) << 4)));
                    o.x += bb[0]; o.y += bb[1]; o.z += bb[2]; o.w += bb[3];
                    const int64_t grow = row0 + rr;
                    if (grow < M) {
                        float* dst = C + (grow * a_cnt + an) * N + col;
                        if (vec_ok && col + 4 <= N) {
                            *reinterpret_cast<float4*>(dst) = o;
                        } else {
                            if (col < N) dst[0] = o.x;
                            if (col + 1 < N) dst[1] = o.y;
                            if (col + 2 < N) dst[2] = o.z;
                            if (col + 3 < N) dst[3] = o.w;
                        }
                    }
                }
                __syncwarp();
            }
#pragma unroll
            for (int i = 0; i < Cfg::HALF; ++i) acc[i] = 0.f;
        }
    } else if (warp != TC_MMA_WARP && warp != TC_TMA_WARP) {
        // ============================ converters ============================
        reg_dec_other();
        const int ct = warp < TC_MMA_WARP ? threadIdx.x - TC_EPI_WARPS * 32 : threadIdx.x - (TC_TMA_WARP + 1) * 32 + 128;
        int it = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            for (int kb = 0; kb < nkb; ++kb, ++it) {
                const int s = it % STAGES;
                mbar_wait_guard(&bars.raw_full[s], (it / STAGES) & 1);
                if (BF) {
                    convert_rows_bf16(smem_base + s * Cfg::STAGE_BYTES, ct >> 5, lane);
                    fence_proxy_async();
                } else if (passes == 3) {
                    convert_region(smem_base + s * Cfg::STAGE_BYTES, Cfg::A_BYTES, Cfg::A_BYTES / 16, ct);
                    fence_proxy_async();   // generic-proxy stores -> visible to the tensor-core (async) proxy
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars.full[s]);
            }
        }
    } else if (warp == TC_MMA_WARP) {
        // ============================ MMA issuer ============================
        // whole warp in the loop, one elected lane issues (uniform-register operands: see tc_gemm_nt_pair_kernel)
        reg_dec_other();
        {
            const uint32_t idesc = BF ? make_idesc_bf16(TC_BM, BN, 0, 0) : make_idesc_tf32(TC_BM, BN, 0, 0);
            int it = 0, ci = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                for (int kb0 = 0; kb0 < nkb; kb0 += chunk_kb, ++ci) {
                    const int buf = ci & 1;
                    mbar_wait_guard(&bars.tempty[buf], ((ci >> 1) & 1) ^ 1);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + (uint32_t)(buf * BN);
                    const int kb1 = kb0 + chunk_kb < nkb ? kb0 + chunk_kb : nkb;
                    for (int kb = kb0; kb < kb1; ++kb, ++it) {
                        const int s = it % STAGES;
                        mbar_wait_guard(&bars.raw_full[s], (it / STAGES) & 1);
                        mbar_wait_guard(&bars.full[s], (it / STAGES) & 1);
                        tc_fence_after();
                        const uint32_t a_hi = smem_base + s * Cfg::STAGE_BYTES;
                        const uint32_t a_lo = a_hi + Cfg::A_BYTES;
                        const uint32_t b_hi = a_lo + Cfg::A_BYTES;
                        const uint32_t b_lo = b_hi + Cfg::B_BYTES;
                        const int kba = (kb + krot) % nkb;
                        const int krem = K - kba * KBE;
                        const int ksteps = krem >= KBE ? KBE / UK : (krem + UK - 1) / UK;
                        const uint64_t da_hi0 = make_smem_desc(a_hi, 16, 1024), db_hi0 = make_smem_desc(b_hi, 16, 1024);
                        const uint64_t da_lo0 = make_smem_desc(a_lo, 16, 1024), db_lo0 = make_smem_desc(b_lo, 16, 1024);
                        const bool first_kb = kb == kb0;
                        if (elect_one()) {
#pragma unroll
                            for (int ks = 0; ks < KBE / UK; ++ks) {
                                if (ks < ksteps) {
                                    const uint64_t o = (uint64_t)(ks * 2);  // 8 tf32 / 16 bf16 = 32 bytes inside the 128 B swizzle row
                                    const uint32_t acc = (first_kb && ks == 0) ? 0u : 1u;
                                    if (BF) {
                                        umma_bf16(d_tmem, da_lo0 + o, db_hi0 + o, idesc, acc);
                                        umma_bf16(d_tmem, da_hi0 + o, db_lo0 + o, idesc, 1);
                                        umma_bf16(d_tmem, da_hi0 + o, db_hi0 + o, idesc, 1);
                                    } else if (passes == 3) {
                                        umma_tf32(d_tmem, da_lo0 + o, db_hi0 + o, idesc, acc);   // small terms first
                                        umma_tf32(d_tmem, da_hi0 + o, db_lo0 + o, idesc, 1);
                                        umma_tf32(d_tmem, da_hi0 + o, db_hi0 + o, idesc, 1);
                                    } else {
                                        umma_tf32(d_tmem, da_hi0 + o, db_hi0 + o, idesc, acc);
                                    }
                                }
                            }
                            umma_commit(&bars.empty[s]);   // smem stage reusable once these MMAs have read it
                        }
                        __syncwarp();
                    }
                    if (elect_one()) umma_commit(&bars.tfull[buf]);     // chunk complete
                    __syncwarp();
                }
            }
        }
        __syncwarp();
    } else {
        // ============================ TMA producer ============================
        reg_dec_other();
        if (lane == 0) {
            tma_prefetch_desc(&map_a);
            tma_prefetch_desc(&map_bhi);
            tma_prefetch_desc(&map_blo);
            const uint32_t tx = BF ? 2u * (uint32_t)Cfg::A_BYTES + 2u * (uint32_t)Cfg::B_BYTES
                                   : (uint32_t)Cfg::A_BYTES + (uint32_t)Cfg::B_BYTES * (passes == 3 ? 2u : 1u);
            int it = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int nt = tile % n_tiles, an = (tile / n_tiles) % a_cnt, mt = tile / (n_tiles * a_cnt);
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int s = it % STAGES;
                    mbar_wait_guard(&bars.empty[s], ((it / STAGES) & 1) ^ 1);
                    unsigned char* st = smem_al + (size_t)s * Cfg::STAGE_BYTES;
                    mbar_arrive_expect_tx(&bars.raw_full[s], tx);
                    const int kba = (kb + krot) % nkb;
                    if (ga.anchors > 0) {
                        const int kcol = kba * KBE, kk = kcol / ga.c, c0 = kcol - kk * ga.c;
                        const int mid = __ldg(ga.table + an * ga.kk_n + kk);
                        tma_load_3d(st, &map_a, c0, mid, mt * TC_BM, &bars.raw_full[s]);
                        if (BF) tma_load_3d(st + Cfg::A_BYTES, &map_a, c0 + 32, mid, mt * TC_BM, &bars.raw_full[s]);
                    } else {
                        tma_load_2d(st, &map_a, kba * KBE, mt * TC_BM, &bars.raw_full[s]);
                        if (BF) tma_load_2d(st + Cfg::A_BYTES, &map_a, kba * KBE + 32, mt * TC_BM, &bars.raw_full[s]);
                    }
                    tma_load_2d(st + 2 * Cfg::A_BYTES, &map_bhi, kba * KBE, nt * BN, &bars.raw_full[s]);
                    if (BF || passes == 3)
                        tma_load_2d(st + 2 * Cfg::A_BYTES + Cfg::B_BYTES, &map_blo, kba * KBE, nt * BN, &bars.raw_full[s]);
                }
            }
        }
        __syncwarp();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == TC_MMA_WARP) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

// ---------------------------------------------------------------------------------- NT kernel, CTA pairs
// bf16x3, N tile 256, two CTAs of a cluster share ONE tcgen05.mma.cta_group::2 (M = 256): CTA r owns rows
// [2*pair_tile + r]*128.. and HALF of the weight tile (128 of the 256 weight rows), so the weight operand -- two
// thirds of the L2 -> SM traffic this kernel is bound by -- is fetched once per pair, and a stage shrinks to 64 KB
// (three stages instead of two).  Only the leader (rank 0) issues MMAs; tcgen05.commit multicasts the "stage free"
// and "chunk complete" arrivals to both CTAs; the converters and epilogues of both CTAs report to the leader's
// barriers with remote mbarrier arrives.  Everything else (TMA-fed raw operand split in place, chunked
// round-to-nearest accumulation, smem-transposed stores, gather addressing) is the single-CTA kernel.
template <int BN_>
struct PairCfg {
    static constexpr int BN = BN_;
    static constexpr int A_BYTES = TC_BM * 64 * 2;          // 16 KB: one bf16 tile (hi or lo) = one raw fp32 box
    static constexpr int BH_BYTES = (BN / 2) * 64 * 2;      // this CTA's half of the weight tile, hi or lo
    static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * BH_BYTES;   // 64 KB at BN = 256
    static constexpr int STAGES = (TC_SMEM_BUDGET - 8 * 1024) / STAGE_BYTES > 6 ? 6 : (TC_SMEM_BUDGET - 8 * 1024) / STAGE_BYTES;
    static constexpr int TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;
    static constexpr int HALF = BN / 2;
    static_assert(STAGES >= 3 && BN % 64 == 0, "pair configuration");
};

// PRE = true (the default path of vgtkb_inter_conv_* and of the intra-conv gather-GEMMs): the activation operand arrives ALREADY split
// into bf16 hi / lo planes in global memory (map_a = hi plane, map_a2 = lo plane; same shape as the fp32 operand), so the
// stage is filled by TMA in its final layout and the converter warps only forward the barrier -- the operand-conversion
// work that bounds the narrow layers (profiles/r1_ncu_gemm_tn_stall_hotspots.txt) disappears from the kernel.
template <int BN_, bool PRE = false>
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_gemm_nt_pair_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_bhi,
                       const __grid_constant__ CUtensorMap map_blo, const __grid_constant__ CUtensorMap map_c, int tma_out,
                       const float* __restrict__ bias, float* __restrict__ C, int64_t M, int N, int K, int chunk_kb, TcGather ga,
                       const __grid_constant__ CUtensorMap map_a2, int fast_in) {
    // bits 4.. of fast_in: timing experiments (VGTKB_DBG; results are garbage): 1 = no MMAs are issued, 2 = no TMA loads, 4 = no output stores
    const int fast = fast_in & 15, dbg = fast_in >> 4;
    // fast != 0: single-pass bf16 (contraction mode 4, BASELINE config 3): only the hi planes take part -- one MMA per k-step
    // instead of three, and the lo planes are neither loaded (PRE) nor used
    using Cfg = PairCfg<BN_>;
    constexpr int STAGES = Cfg::STAGES, BN = Cfg::BN, KBE = 64, UK = 16;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
    __shared__ __align__(8) TcBarriers bars;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
    const int m_tiles = (int)((M + TC_BM - 1) / TC_BM);
    const int mp_tiles = (m_tiles + 1) / 2;
    const int n_tiles = (N + BN - 1) / BN;
    const int a_cnt = ga.anchors > 0 ? ga.anchors : 1;
    const int total_tiles = mp_tiles * a_cnt * n_tiles;     // pair tile -> (mtp, an, nt), nt fastest
    const int nkb = (K + KBE - 1) / KBE;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            // PRE: the stage's boxes are issued by up to four different warps (one expect_tx arrive each), see below
            mbar_init(&bars.raw_full[s], PRE ? (fast ? 2 : 4) : 1);
            // converter warps of BOTH CTAs (PRE: one forwarding warp per CTA); used in the leader
            mbar_init(&bars.full[s], PRE ? 2 : 2 * TC_CONV_WARPS);
            mbar_init(&bars.empty[s], 1);                     // multicast tcgen05.commit
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&bars.tfull[a], 1);                     // multicast tcgen05.commit
            mbar_init(&bars.tempty[a], 2 * TC_EPI_WARPS);     // epilogue warps of BOTH CTAs (used in the leader)
        }
        fence_barrier_init();
    }
    if (warp == TC_MMA_WARP) tmem_alloc_pair(&bars.tmem_base, Cfg::TMEM_COLS);
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = bars.tmem_base;

    // TMA issue loop of one thread.  which = -1: all boxes of every stage (raw fp32 operand, converter path);
    // which = 0..3 (PRE): only box `which` of every stage -- 0: A hi, 1: A lo, 2: weights hi, 3: weights lo
    auto issue_loads = [&](int which) {
        if (which < 0 || which == 0) tma_prefetch_desc(&map_a);
        if (PRE && which == 1) tma_prefetch_desc(&map_a2);
        if (which < 0 || which == 2) tma_prefetch_desc(&map_bhi);
        if (which < 0 || which == 3) tma_prefetch_desc(&map_blo);
        if (fast && (which == 1 || which == 3)) return;          // single-pass bf16: the lo planes are not loaded
        const uint32_t tx = which < 0 ? 2u * (uint32_t)Cfg::A_BYTES + (fast ? 1u : 2u) * (uint32_t)Cfg::BH_BYTES
                                      : (which < 2 ? (uint32_t)Cfg::A_BYTES : (uint32_t)Cfg::BH_BYTES);
        int it = 0;
        for (int tile = pair; tile < total_tiles; tile += npairs) {
            const int nt = tile % n_tiles, an = (tile / n_tiles) % a_cnt, mt = 2 * (tile / (n_tiles * a_cnt)) + (int)rank;
            const int brow = nt * BN + (int)rank * (BN / 2);
            for (int kb = 0; kb < nkb; ++kb, ++it) {
                const int s = it % STAGES;
                mbar_wait_guard(&bars.empty[s], ((it / STAGES) & 1) ^ 1);
                if (!elect_one()) continue;      // the warp runs the loop, one lane issues (uniform-register operands)
                unsigned char* st = smem_al + (size_t)s * Cfg::STAGE_BYTES;
                if (dbg & 2) {
                    mbar_arrive(&bars.raw_full[s]);
                    continue;
                }
                mbar_arrive_expect_tx(&bars.raw_full[s], tx);
                if (PRE) {          // bf16 planes: one box of 64 columns per plane, already the MMA's layout
                    if (which < 2) {
                        const CUtensorMap* mp = which == 0 ? &map_a : &map_a2;
                        if (ga.anchors > 0) {
                            const int kcol = kb * KBE, kk = kcol / ga.c, c0 = kcol - kk * ga.c;
                            const int mid = __ldg(ga.table + an * ga.kk_n + kk);
                            tma_load_3d(st + which * Cfg::A_BYTES, mp, c0, mid, mt * TC_BM, &bars.raw_full[s]);
                        } else {
                            tma_load_2d(st + which * Cfg::A_BYTES, mp, kb * KBE, mt * TC_BM, &bars.raw_full[s]);
                        }
                    } else if (which == 2) {
                        tma_load_2d(st + 2 * Cfg::A_BYTES, &map_bhi, kb * KBE, brow, &bars.raw_full[s]);
                    } else {
                        tma_load_2d(st + 2 * Cfg::A_BYTES + Cfg::BH_BYTES, &map_blo, kb * KBE, brow, &bars.raw_full[s]);
                    }
                    continue;
                }
                if (ga.anchors > 0) {
                    const int kcol = kb * KBE, kk = kcol / ga.c, c0 = kcol - kk * ga.c;
                    const int mid = __ldg(ga.table + an * ga.kk_n + kk);
                    tma_load_3d(st, &map_a, c0, mid, mt * TC_BM, &bars.raw_full[s]);
                    tma_load_3d(st + Cfg::A_BYTES, &map_a, c0 + 32, mid, mt * TC_BM, &bars.raw_full[s]);
                } else {
                    tma_load_2d(st, &map_a, kb * KBE, mt * TC_BM, &bars.raw_full[s]);
                    tma_load_2d(st + Cfg::A_BYTES, &map_a, kb * KBE + 32, mt * TC_BM, &bars.raw_full[s]);
                }
                tma_load_2d(st + 2 * Cfg::A_BYTES, &map_bhi, kb * KBE, brow, &bars.raw_full[s]);
                if (!fast) tma_load_2d(st + 2 * Cfg::A_BYTES + Cfg::BH_BYTES, &map_blo, kb * KBE, brow, &bars.raw_full[s]);
            }
        }
    };

    if (warp < TC_EPI_WARPS) {
        // ============================ epilogue ============================
        reg_inc_epi();
        const int q = warp & 3, h = warp >> 2;
        float acc[Cfg::HALF];
#pragma unroll
        for (int i = 0; i < Cfg::HALF; ++i) acc[i] = 0.f;
        int ci = 0;
        const bool vec_ok = (N % 4 == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0);
        for (int tile = pair; tile < total_tiles; tile += npairs) {
            const int nt = tile % n_tiles, an = (tile / n_tiles) % a_cnt, mt = 2 * (tile / (n_tiles * a_cnt)) + (int)rank;
            for (int kb0 = 0; kb0 < nkb; kb0 += chunk_kb, ++ci) {
                const int buf = ci & 1;
                mbar_wait_guard(&bars.tfull[buf], (ci >> 1) & 1);
                tc_fence_after();
#pragma unroll
                for (int j = 0; j < Cfg::HALF / 32; ++j) {
                    float v[32];
                    tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN + h * Cfg::HALF + j * 32), v);
#pragma unroll
                    for (int i = 0; i < 32; ++i) acc[j * 32 + i] += v[i];
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_remote(&bars.tempty[buf], 0);
            }
            const uint32_t stg = smem_base + (uint32_t)(STAGES * Cfg::STAGE_BYTES) + (uint32_t)warp * 4096u;
            const int64_t row0 = (int64_t)mt * TC_BM + q * 32;
            const int c_base = nt * BN + h * Cfg::HALF;
            if (tma_out) {
                // tensor-map store: the 32x32 block goes registers -> 128B-swizzled staging tile -> ONE bulk tensor copy, instead
                // of being read back and stored by every lane (ncu on the wide-output launches dG = gy W: l1tex 90 %, LSU
                // wavefronts 87 %, tensor pipe 62 % -- the epilogue's shared-memory transposes were the bound)
#pragma unroll
                for (int j = 0; j < Cfg::HALF / 32; ++j) {
                    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // staging tile read out
                    __syncwarp();
                    const int colj = c_base + j * 32;
#pragma unroll
                    for (int c4 = 0; c4 < 8; ++c4) {
                        float4 o = make_float4(acc[j * 32 + c4 * 4], acc[j * 32 + c4 * 4 + 1], acc[j * 32 + c4 * 4 + 2],
                                               acc[j * 32 + c4 * 4 + 3]);
                        if (bias != nullptr) {
                            const int col = colj + c4 * 4;
                            if (col < N) o.x += __ldg(bias + col);
                            if (col + 1 < N) o.y += __ldg(bias + col + 1);
                            if (col + 2 < N) o.z += __ldg(bias + col + 2);
                            if (col + 3 < N) o.w += __ldg(bias + col + 3);
                        }
                        st_shared_v4(stg + (uint32_t)(lane * 128 + ((c4 ^ (lane & 7)) << 4)), o);
                    }
                    fence_proxy_async();                      // generic-proxy writes -> visible to the bulk-copy engine
                    __syncwarp();
                    if (lane == 0 && colj < N && row0 < M && !(dbg & 4)) {  // rows / columns beyond M / N are clipped by the tensor map
                        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&map_c),
                                     "r"(stg), "r"(colj), "r"((int)row0)
                                     : "memory");
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                }
            } else
#pragma unroll
            for (int j = 0; j < Cfg::HALF / 32; ++j) {
#pragma unroll
                for (int c4 = 0; c4 < 8; ++c4)
                    st_shared_v4(stg + (uint32_t)(lane * 128 + ((c4 ^ (lane & 7)) << 4)),
                                 make_float4(acc[j * 32 + c4 * 4], acc[j * 32 + c4 * 4 + 1], acc[j * 32 + c4 * 4 + 2],
                                             acc[j * 32 + c4 * 4 + 3]));
                __syncwarp();
                const int c4 = lane & 7;
                const int col = c_base + j * 32 + c4 * 4;
                float bb[4] = {0.f, 0.f, 0.f, 0.f};
                if (bias != nullptr) {
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        if (col + e < N) bb[e] = bias[col + e];
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int rr = i * 4 + (lane >> 3);
                    float4 o = ld_shared_v4(stg + (uint32_t)(rr * 128 + ((c4 ^ (rr & 7)) << 4)));
                    o.x += bb[0]; o.y += bb[1]; o.z += bb[2]; o.w += bb[3];
                    const int64_t grow = row0 + rr;
                    if (grow < M && !(dbg & 4)) {
                        float* dst = C + (grow * a_cnt + an) * N + col;
                        if (vec_ok && col + 4 <= N) {
                            *reinterpret_cast<float4*>(dst) = o;
                        } else {
                            if (col < N) dst[0] = o.x;
                            if (col + 1 < N) dst[1] = o.y;
                            if (col + 2 < N) dst[2] = o.z;
                            if (col + 3 < N) dst[3] = o.w;
                        }
                    }
                }
                __syncwarp();
            }
#pragma unroll
            for (int i = 0; i < Cfg::HALF; ++i) acc[i] = 0.f;
        }
        if (tma_out && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // all tile stores landed
    } else if (warp != TC_MMA_WARP && warp != TC_TMA_WARP) {
        // ============================ converters ============================
        reg_dec_other();
        const int ct = warp < TC_MMA_WARP ? threadIdx.x - TC_EPI_WARPS * 32 : threadIdx.x - (TC_TMA_WARP + 1) * 32 + 128;
        if (PRE) {
            // nothing to convert.  Measured (scripts/tma_rate.cu, profiles/r2_tma_issue_rate.csv): ONE issuing thread gets at
            // most one 128-byte-row box through the TMA unit per ~700 cycles whatever its height (21 B/clk/SM for 128-row
            // boxes: the 3.3-4.5 TB/s every plane-fed contraction was stuck at), and the rate scales with the number of
            // issuing WARPS.  So the four boxes of a stage are issued by four warps: the TMA warp (A hi) and converter
            // warps 0..2 (A lo, B hi, B lo); converter warp 3 forwards "stage landed" to the leader CTA.
            const int cw = ct >> 5;
            if (cw < 3) {
                issue_loads(cw + 1);
            } else if (cw == 3) {
                int it = 0;
                for (int tile = pair; tile < total_tiles; tile += npairs)
                    for (int kb = 0; kb < nkb; ++kb, ++it) {
                        const int s = it % STAGES;
                        mbar_wait_guard(&bars.raw_full[s], (it / STAGES) & 1);
                        if (lane == 0) mbar_arrive_remote(&bars.full[s], 0);
                    }
            }
            __syncwarp();
        } else {
            int it = 0;
            for (int tile = pair; tile < total_tiles; tile += npairs) {
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int s = it % STAGES;
                    mbar_wait_guard(&bars.raw_full[s], (it / STAGES) & 1);
                    convert_rows_bf16(smem_base + s * Cfg::STAGE_BYTES, ct >> 5, lane);
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_remote(&bars.full[s], 0);
                }
            }
        }
    } else if (warp == TC_MMA_WARP) {
        // ============================ MMA issuer (leader CTA only) ============================
        // The WHOLE warp runs the loop (waits, descriptor arithmetic: warp-uniform, kept in uniform registers) and one
        // elected lane issues the tcgen05 instructions.  The first version ran everything under `lane == 0`: every MMA
        // then needed an ELECT + 5 x R2UR + branch sequence to move its operands into uniform registers, ~100 cycles per
        // MMA -- with the loads switched off (VGTKB_DBG=2) the N = 64 / 128 launches took as long as with them.
        reg_dec_other();
        if (rank == 0) {
            const uint32_t idesc = make_idesc_bf16(2 * TC_BM, BN, 0, 0);
            int it = 0, ci = 0;
            for (int tile = pair; tile < total_tiles; tile += npairs) {
                for (int kb0 = 0; kb0 < nkb; kb0 += chunk_kb, ++ci) {
                    const int buf = ci & 1;
                    mbar_wait_guard_cluster(&bars.tempty[buf], ((ci >> 1) & 1) ^ 1);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + (uint32_t)(buf * BN);
                    const int kb1 = kb0 + chunk_kb < nkb ? kb0 + chunk_kb : nkb;
                    for (int kb = kb0; kb < kb1; ++kb, ++it) {
                        const int s = it % STAGES;
                        mbar_wait_guard_cluster(&bars.full[s], (it / STAGES) & 1);
                        tc_fence_after();
                        const uint32_t a_hi = smem_base + s * Cfg::STAGE_BYTES;
                        const uint32_t a_lo = a_hi + Cfg::A_BYTES;
                        const uint32_t b_hi = a_lo + Cfg::A_BYTES;
                        const uint32_t b_lo = b_hi + Cfg::BH_BYTES;
                        const int krem = K - kb * KBE;
                        const int ksteps = (dbg & 1) ? 0 : (krem >= KBE ? KBE / UK : (krem + UK - 1) / UK);
                        // descriptors of k-step 0; a k-step advances the start-address field (bits 0-13, address >> 4) by 2
                        const uint64_t da_hi0 = make_smem_desc(a_hi, 16, 1024), db_hi0 = make_smem_desc(b_hi, 16, 1024);
                        const uint64_t da_lo0 = make_smem_desc(a_lo, 16, 1024), db_lo0 = make_smem_desc(b_lo, 16, 1024);
                        const bool first_kb = kb == kb0;
                        if (elect_one()) {
#pragma unroll
                            for (int ks = 0; ks < KBE / UK; ++ks) {
                                if (ks < ksteps) {
                                    const uint64_t o = (uint64_t)(ks * 2);
                                    const uint32_t acc = (first_kb && ks == 0) ? 0u : 1u;
                                    if (fast) {
                                        umma_bf16_pair(d_tmem, da_hi0 + o, db_hi0 + o, idesc, acc);
                                    } else {
                                        umma_bf16_pair(d_tmem, da_lo0 + o, db_hi0 + o, idesc, acc);
                                        umma_bf16_pair(d_tmem, da_hi0 + o, db_lo0 + o, idesc, 1);
                                        umma_bf16_pair(d_tmem, da_hi0 + o, db_hi0 + o, idesc, 1);
                                    }
                                }
                            }
                            umma_commit_pair(&bars.empty[s]);
                        }
                        __syncwarp();
                    }
                    if (elect_one()) umma_commit_pair(&bars.tfull[buf]);
                    __syncwarp();
                }
            }
        }
        __syncwarp();
    } else {
        // ============================ TMA producer (both CTAs: own rows, own half of the weights) =========
        reg_dec_other();
        issue_loads(PRE ? 0 : -1);
        __syncwarp();
    }

    tc_fence_before();
    cluster_sync_all();
    if (warp == TC_MMA_WARP) {
        tc_fence_after();
        tmem_dealloc_pair(tmem_base, Cfg::TMEM_COLS);
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

// bf16 hi/lo split of the weight operand into a workspace [2][n] of bf16
__global__ void split_bf16_kernel(int64_t n, const float* __restrict__ x, uint16_t* __restrict__ hi, uint16_t* __restrict__ lo) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float v = x[i];
        const uint32_t h = pack_bf16x2(v, 0.f) & 0xFFFFu;
        const float hf = __uint_as_float(h << 16);
        hi[i] = (uint16_t)h;
        lo[i] = (uint16_t)(pack_bf16x2(v - hf, 0.f) & 0xFFFFu);
    }
}

// bf16x3 converter for an MN-major operand (weight-gradient kernel).  The raw fp32 k-block (64 rows of R
// x W columns) arrives as W/32 TMA boxes of 64 rows x 128 B (8 KB each, 128B swizzle).  It becomes, IN
// PLACE, W/64 bf16 hi atoms-columns (64 k-rows x 128 B = 64 columns each, 8 KB) followed by W/64 lo ones, in
// the canonical 128B-swizzled MN-major layout.  Row k of every output tile only depends on row k of the raw
// boxes, and one warp owns a row: load everything, convert, __syncwarp, store.
template <int W>
__device__ __forceinline__ void convert_tn_row_bf16(uint32_t region, int k, int lane) {
    constexpr int NP = (W + 127) / 128;
    constexpr uint32_t LO_OFF = (uint32_t)(W / 64) * 8192u;
    float4 v[NP];
    const uint32_t rowoff = (uint32_t)(k * 128);
#pragma unroll
    for (int t = 0; t < NP; ++t) {
        const int mn = t * 128 + 4 * lane;
        if (mn < W) v[t] = ld_shared_v4(region + (uint32_t)(t * 4 + (lane >> 3)) * 8192u + rowoff + (uint32_t)(((lane & 7) ^ (k & 7)) << 4));
        else v[t] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    uint32_t h0[NP], h1[NP], l0[NP], l1[NP];
#pragma unroll
    for (int t = 0; t < NP; ++t) {
        h0[t] = pack_bf16x2(v[t].x, v[t].y);
        h1[t] = pack_bf16x2(v[t].z, v[t].w);
        l0[t] = pack_bf16x2(v[t].x - __uint_as_float(h0[t] << 16), v[t].y - __uint_as_float(h0[t] & 0xFFFF0000u));
        l1[t] = pack_bf16x2(v[t].z - __uint_as_float(h1[t] << 16), v[t].w - __uint_as_float(h1[t] & 0xFFFF0000u));
    }
    __syncwarp();
#pragma unroll
    for (int t = 0; t < NP; ++t) {
        const int mn = t * 128 + 4 * lane;
        if (mn < W) {
            const int e = mn & 63;
            const uint32_t dst = region + (uint32_t)(mn >> 6) * 8192u + rowoff + (uint32_t)((((e >> 3) ^ (k & 7)) << 4) | (((e >> 2) & 1) << 3));
            st_shared_u2(dst, h0[t], h1[t]);
            st_shared_u2(dst + LO_OFF, l0[t], l1[t]);
        }
    }
}

// four rows of a 128-column operand per call (more loads in flight per warp)
__device__ __forceinline__ void convert_tn_rows4_bf16(uint32_t region, int k0, int lane) {
    float4 v[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int k = k0 + r;
        v[r] = ld_shared_v4(region + (uint32_t)(lane >> 3) * 8192u + (uint32_t)(k * 128) + (uint32_t)(((lane & 7) ^ (k & 7)) << 4));
    }
    uint32_t h0[4], h1[4], l0[4], l1[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        h0[r] = pack_bf16x2(v[r].x, v[r].y);
        h1[r] = pack_bf16x2(v[r].z, v[r].w);
        l0[r] = pack_bf16x2(v[r].x - __uint_as_float(h0[r] << 16), v[r].y - __uint_as_float(h0[r] & 0xFFFF0000u));
        l1[r] = pack_bf16x2(v[r].z - __uint_as_float(h1[r] << 16), v[r].w - __uint_as_float(h1[r] & 0xFFFF0000u));
    }
    __syncwarp();
    const int mn = 4 * lane, e = mn & 63;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int k = k0 + r;
        const uint32_t dst = region + (uint32_t)(mn >> 6) * 8192u + (uint32_t)(k * 128) + (uint32_t)((((e >> 3) ^ (k & 7)) << 4) | (((e >> 2) & 1) << 3));
        st_shared_u2(dst, h0[r], h1[r]);
        st_shared_u2(dst + 16384u, l0[r], l1[r]);
    }
}

// ---------------------------------------------------------------------------------- TN kernel
// T[i, j] = sum_r P[r, p0+i] * Q[r, q0+j]   (i < 128, j < BN), both operands MN-major in shared
// memory in the SWIZZLE_128B_BASE32B canonical layout (the layout MN-major tf32 operands require):
// atoms of 32 (MN) x 4 (K) fp32 = 512 B, 32-byte chunks XOR-permuted by the k-row.  One TMA box
// (32 rows of R x 32 columns, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B) fills one MN atom column for a
// whole k-block, so a stage is [mn-atom][32 k-rows x 128 B]: LBO (next MN atom) = 4096 B, SBO (next
// k-atom of 4 rows) = 512 B, and one MMA (K = 8) starts 1024 B after the previous one.
// Work item = (R slice, P tile, Q tile); finished tiles are added into C with red.global.add:
//   C[(q0+j) * ldc + (p0+i)] += T[i, j]      (lanes run along i: coalesced)
// BF = false: kind::tf32 (k-block = 32 rows of R).  BF = true: bf16x3, kind::f16, k-block = 64 rows of R,
// operands in the canonical 128B-swizzled MN-major bf16 layout: stage = [mn-atom of 64 columns][64 k-rows x 128 B],
// LBO (next MN atom) = 8192 B, SBO (next 8 k-rows) = 1024 B, one MMA (K = 16) starts 2048 B after the previous.
// PRE = true (with BF; the weight gradients of vgtkb_inter_conv_backward and of the intra conv): the wide operand P arrives as two
// bf16 planes (map_p = hi, map_p2 = lo) and is loaded by TMA in the MMA's layout, like Q; the converters only forward the barrier.
template <int BN, bool BF, bool PRE = false>
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_gemm_tn_kernel(const __grid_constant__ CUtensorMap map_p, const __grid_constant__ CUtensorMap map_q,
                  const __grid_constant__ CUtensorMap map_q2, int Pw, int Qw,
                  float* __restrict__ C, int ldc, int64_t R, int64_t rows_per_split, int splits, int passes, int chunk_kb,
                  TcGather ga, int anchors_per_item, const __grid_constant__ CUtensorMap map_p2, int fast) {
    using Cfg = TcCfg<BN>;
    constexpr int STAGES = Cfg::STAGES;
    constexpr uint32_t MN_LBO = BF ? 8192 : 4096, K_SBO = BF ? 1024 : 512, L32 = BF ? 2 : 1;
    constexpr int KR = BF ? 64 : TC_BK;       // rows of R per k-block
    constexpr int UK = BF ? 16 : 8;           // rows of R per MMA
    constexpr uint32_t KSTEP_BYTES = BF ? 2048 : 1024;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
    __shared__ __align__(8) TcBarriers bars;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int p_tiles = (Pw + TC_BM - 1) / TC_BM;
    const int q_tiles = (Qw + BN - 1) / BN;
    const int tiles = p_tiles * q_tiles;
    // gather mode (intra-conv weight gradient): R counts points, an item also owns a group of output anchors and
    // runs through them one after the other with the SAME register accumulators (one red.global.add pass per item)
    const int n_groups = ga.anchors > 0 ? (ga.anchors + anchors_per_item - 1) / anchors_per_item : 1;
    const int items = tiles * n_groups * splits;        // item -> (sp, grp, tile), tile fastest

    if (threadIdx.x == 0) tc_init_barriers<STAGES>(bars, (BF && PRE) ? (fast ? 2 : 4) : 1);   // PRE: four issuing warps
    if (warp == TC_MMA_WARP) tmem_alloc(&bars.tmem_base, Cfg::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars.tmem_base;

    auto item_nkb = [&](int item, int64_t& r0) {
        const int sp = item / (tiles * n_groups);
        r0 = (int64_t)sp * rows_per_split;
        const int64_t r1 = r0 + rows_per_split < R ? r0 + rows_per_split : R;
        return (int)((r1 - r0 + KR - 1) / KR);
    };
    // anchors [a0, a0 + nan) of an item (plain GEMM: one pseudo anchor)
    auto item_anchors = [&](int item, int& a0) {
        if (ga.anchors <= 0) { a0 = 0; return 1; }
        const int grp = (item / tiles) % n_groups;
        a0 = grp * anchors_per_item;
        return min(anchors_per_item, ga.anchors - a0);
    };

    // TMA issue loop of one thread: which = -1 every box of every stage; BF && PRE: 0 = P hi, 1 = P lo, 2 = Q hi, 3 = Q lo
    auto issue_loads = [&](int which) {
        if (which <= 0) tma_prefetch_desc(&map_p);
        if (PRE && which == 1) tma_prefetch_desc(&map_p2);
        if (which < 0 || which == 2) tma_prefetch_desc(&map_q);
        if (BF && (which < 0 || which == 3)) tma_prefetch_desc(&map_q2);
        if (fast && (which == 1 || which == 3)) return;
        // fast (bf16 single pass): the lo planes of Q, and of P when it arrives as planes, are not loaded
        const uint32_t tx = which >= 0 ? (which < 2 ? (uint32_t)Cfg::A_BYTES : (uint32_t)Cfg::B_BYTES)
                            : BF     ? 2u * (uint32_t)Cfg::A_BYTES + (fast ? 1u : 2u) * (uint32_t)Cfg::B_BYTES
                                     : (uint32_t)Cfg::A_BYTES + (uint32_t)Cfg::B_BYTES;
        constexpr uint32_t BOX_BYTES = BF ? 8192 : 4096;
        int it = 0;
        for (int item = blockIdx.x; item < items; item += gridDim.x) {
            const int tile = item % tiles;
            const int p0 = (tile / q_tiles) * TC_BM, q0 = (tile % q_tiles) * BN;
            int64_t r0;
            int a0;
            const int nkb = item_nkb(item, r0);
            const int nblk = nkb * item_anchors(item, a0);
            for (int kb = 0; kb < nblk; ++kb, ++it) {
                const int s = it % STAGES;
                mbar_wait_guard(&bars.empty[s], ((it / STAGES) & 1) ^ 1);
                if (!elect_one()) continue;      // the warp runs the loop, one lane issues (uniform-register operands)
                unsigned char* st = smem_al + (size_t)s * Cfg::STAGE_BYTES;
                const int row = (int)(r0 + (int64_t)(kb % nkb) * KR);
                mbar_arrive_expect_tx(&bars.raw_full[s], tx);
                if (ga.anchors > 0) {
                    // gathered P: column (kk, c) of anchor `an` lives at X[point, table[an, kk], c]
                    const int an = a0 + kb / nkb;
                    if (PRE) {
                        if (which < 2) {
#pragma unroll
                            for (int a = 0; a < TC_BM / 64; ++a) {       // bf16 planes: atoms of 64 columns, final layout
                                const int cc = p0 + a * 64;
                                const int kk = cc < Pw ? cc / ga.c : 0;
                                const int c0 = cc < Pw ? cc - kk * ga.c : ga.c;                 // out of bounds: zeros
                                const int mid = cc < Pw ? __ldg(ga.table + an * ga.kk_n + kk) : 0;
                                tma_load_3d(st + which * Cfg::A_BYTES + a * 8192, which == 0 ? &map_p : &map_p2, c0, mid, row, &bars.raw_full[s]);
                            }
                        } else {
#pragma unroll
                            for (int a = 0; a < BN / 64; ++a)
                                tma_load_3d(st + 2 * Cfg::A_BYTES + (which - 2) * Cfg::B_BYTES + a * 8192, which == 2 ? &map_q : &map_q2,
                                            q0 + a * 64, an, row, &bars.raw_full[s]);
                        }
                        continue;
                    } else
#pragma unroll
                    for (int a = 0; a < TC_BM / 32; ++a) {
                        const int cc = p0 + a * 32;
                        if (cc < Pw) {
                            const int kk = cc / ga.c;
                            tma_load_3d(st + a * BOX_BYTES, &map_p, cc - kk * ga.c, __ldg(ga.table + an * ga.kk_n + kk), row, &bars.raw_full[s]);
                        } else {
                            tma_load_3d(st + a * BOX_BYTES, &map_p, ga.c, 0, row, &bars.raw_full[s]);   // out of bounds: zeros
                        }
                    }
                    if (BF) {
#pragma unroll
                        for (int a = 0; a < BN / 64; ++a) {
                            tma_load_3d(st + 2 * Cfg::A_BYTES + a * 8192, &map_q, q0 + a * 64, an, row, &bars.raw_full[s]);
                            if (!fast) tma_load_3d(st + 2 * Cfg::A_BYTES + Cfg::B_BYTES + a * 8192, &map_q2, q0 + a * 64, an, row, &bars.raw_full[s]);
                        }
                    } else {
#pragma unroll
                        for (int a = 0; a < BN / 32; ++a)
                            tma_load_3d(st + 2 * Cfg::A_BYTES + a * BOX_BYTES, &map_q, q0 + a * 32, an, row, &bars.raw_full[s]);
                    }
                    continue;
                }
                if (PRE) {
                    if (which < 2) {
#pragma unroll
                        for (int a = 0; a < TC_BM / 64; ++a)
                            tma_load_2d(st + which * Cfg::A_BYTES + a * 8192, which == 0 ? &map_p : &map_p2, p0 + a * 64, row, &bars.raw_full[s]);
                    } else {
#pragma unroll
                        for (int a = 0; a < BN / 64; ++a)
                            tma_load_2d(st + 2 * Cfg::A_BYTES + (which - 2) * Cfg::B_BYTES + a * 8192, which == 2 ? &map_q : &map_q2,
                                        q0 + a * 64, row, &bars.raw_full[s]);
                    }
                    continue;
                } else
#pragma unroll
                for (int a = 0; a < TC_BM / 32; ++a)
                    tma_load_2d(st + a * BOX_BYTES, &map_p, p0 + a * 32, row, &bars.raw_full[s]);
                if (BF) {
#pragma unroll
                    for (int a = 0; a < BN / 64; ++a) {   // bf16 hi / lo atoms of 64 columns, final layout
                        tma_load_2d(st + 2 * Cfg::A_BYTES + a * 8192, &map_q, q0 + a * 64, row, &bars.raw_full[s]);
                        if (!fast) tma_load_2d(st + 2 * Cfg::A_BYTES + Cfg::B_BYTES + a * 8192, &map_q2, q0 + a * 64, row, &bars.raw_full[s]);
                    }
                } else {
#pragma unroll
                    for (int a = 0; a < BN / 32; ++a)
                        tma_load_2d(st + 2 * Cfg::A_BYTES + a * BOX_BYTES, &map_q, q0 + a * 32, row, &bars.raw_full[s]);
                }
            }
        }
    };

    if (warp < TC_EPI_WARPS) {
        // ============================ epilogue ============================
        reg_inc_epi();
        const int q = warp & 3, h = warp >> 2;
        float acc[Cfg::HALF];
#pragma unroll
        for (int i = 0; i < Cfg::HALF; ++i) acc[i] = 0.f;
        int ci = 0;
        for (int item = blockIdx.x; item < items; item += gridDim.x) {
            const int tile = item % tiles;
            const int p0 = (tile / q_tiles) * TC_BM, q0 = (tile % q_tiles) * BN;
            int64_t r0;
            int a0;
            const int nblk = item_nkb(item, r0) * item_anchors(item, a0);
            for (int kb0 = 0; kb0 < nblk; kb0 += chunk_kb, ++ci) {
                const int buf = ci & 1;
                mbar_wait_guard(&bars.tfull[buf], (ci >> 1) & 1);
                tc_fence_after();
#pragma unroll
                for (int j = 0; j < Cfg::HALF / 32; ++j) {
                    float v[32];
                    tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN + h * Cfg::HALF + j * 32), v);
#pragma unroll
                    for (int i = 0; i < 32; ++i) acc[j * 32 + i] += v[i];
                }
                tc_fence_before();
                mbar_arrive(&bars.tempty[buf]);
            }
            const int i = p0 + q * 32 + lane;
            const int c_base = q0 + h * Cfg::HALF;
            if (i < Pw) {
#pragma unroll
                for (int j = 0; j < Cfg::HALF; ++j)
                    if (c_base + j < Qw) atomicAdd(C + (size_t)(c_base + j) * ldc + i, acc[j]);
            }
#pragma unroll
            for (int j = 0; j < Cfg::HALF; ++j) acc[j] = 0.f;
        }
    } else if (warp != TC_MMA_WARP && warp != TC_TMA_WARP) {
        // ============================ converters (both operands) ============================
        reg_dec_other();
        const int ct = warp < TC_MMA_WARP ? threadIdx.x - TC_EPI_WARPS * 32 : threadIdx.x - (TC_TMA_WARP + 1) * 32 + 128;
        if (BF && PRE) {
            // both operands arrive in their final layout: converter warps 0..2 are TMA issuers (P lo, Q hi, Q lo; the TMA
            // warp issues P hi -- one issuing thread moves ~21 B/clk/SM, see tc_gemm_nt_pair_kernel), the MMA warp waits
            // on raw_full alone
            if ((ct >> 5) < 3) issue_loads((ct >> 5) + 1);
            __syncwarp();
        } else {
            int it = 0;
            for (int item = blockIdx.x; item < items; item += gridDim.x) {
                int64_t r0;
                int a0;
                const int nblk = item_nkb(item, r0) * item_anchors(item, a0);
                for (int kb = 0; kb < nblk; ++kb, ++it) {
                    const int s = it % STAGES;
                    mbar_wait_guard(&bars.raw_full[s], (it / STAGES) & 1);
                    if (BF) {
                        const uint32_t st = smem_base + s * Cfg::STAGE_BYTES;
                        // only P is converted here: Q (the narrow operand every P tile re-reads) was split to bf16
                        // hi/lo once in global memory and arrives in its final layout through TMA
                        for (int task = ct >> 5; task < KR / 4; task += TC_CONV_WARPS) convert_tn_rows4_bf16(st, task * 4, lane);
                        fence_proxy_async();
                    } else if (passes == 3) {
                        const uint32_t st = smem_base + s * Cfg::STAGE_BYTES;
                        convert_region(st, Cfg::A_BYTES, Cfg::A_BYTES / 16, ct);
                        convert_region(st + 2 * Cfg::A_BYTES, Cfg::B_BYTES, Cfg::B_BYTES / 16, ct);
                        fence_proxy_async();
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bars.full[s]);
                }
            }
        }
    } else if (warp == TC_MMA_WARP) {
        // ============================ MMA issuer ============================
        // whole warp in the loop, one elected lane issues (uniform-register operands: see tc_gemm_nt_pair_kernel)
        reg_dec_other();
        {
            const uint32_t idesc = BF ? make_idesc_bf16(TC_BM, BN, 1, 1) : make_idesc_tf32(TC_BM, BN, 1, 1);   // MN-major
            int it = 0, ci = 0;
            for (int item = blockIdx.x; item < items; item += gridDim.x) {
                int64_t r0;
                int a0;
                const int nkb = item_nkb(item, r0);
                const int nblk = nkb * item_anchors(item, a0);
                const int64_t rows = (r0 + rows_per_split < R ? r0 + rows_per_split : R) - r0;
                for (int kb0 = 0; kb0 < nblk; kb0 += chunk_kb, ++ci) {
                    const int buf = ci & 1;
                    mbar_wait_guard(&bars.tempty[buf], ((ci >> 1) & 1) ^ 1);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + (uint32_t)(buf * BN);
                    const int kb1 = kb0 + chunk_kb < nblk ? kb0 + chunk_kb : nblk;
                    for (int kb = kb0; kb < kb1; ++kb, ++it) {
                        const int s = it % STAGES;
                        mbar_wait_guard(&bars.raw_full[s], (it / STAGES) & 1);
                        if (!(BF && PRE)) mbar_wait_guard(&bars.full[s], (it / STAGES) & 1);
                        tc_fence_after();
                        const uint32_t p_hi = smem_base + s * Cfg::STAGE_BYTES;
                        const uint32_t p_lo = p_hi + Cfg::A_BYTES;
                        const uint32_t q_hi = p_lo + Cfg::A_BYTES;
                        const uint32_t q_lo = q_hi + Cfg::B_BYTES;
                        const int64_t rrem = rows - (int64_t)(kb % nkb) * KR;
                        const int ksteps = rrem >= KR ? KR / UK : (int)((rrem + UK - 1) / UK);
                        const uint64_t dp_hi0 = make_smem_desc(p_hi, MN_LBO, K_SBO, L32), dq_hi0 = make_smem_desc(q_hi, MN_LBO, K_SBO, L32);
                        const uint64_t dp_lo0 = make_smem_desc(p_lo, MN_LBO, K_SBO, L32), dq_lo0 = make_smem_desc(q_lo, MN_LBO, K_SBO, L32);
                        const bool first_kb = kb == kb0;
                        if (elect_one()) {
#pragma unroll
                            for (int ks = 0; ks < KR / UK; ++ks) {
                                if (ks < ksteps) {
                                    const uint64_t o = (uint64_t)(ks * (KSTEP_BYTES >> 4));   // one MMA's rows of R
                                    const uint32_t acc = (first_kb && ks == 0) ? 0u : 1u;
                                    if (BF && fast) {
                                        umma_bf16(d_tmem, dp_hi0 + o, dq_hi0 + o, idesc, acc);
                                    } else if (BF) {
                                        umma_bf16(d_tmem, dp_lo0 + o, dq_hi0 + o, idesc, acc);
                                        umma_bf16(d_tmem, dp_hi0 + o, dq_lo0 + o, idesc, 1);
                                        umma_bf16(d_tmem, dp_hi0 + o, dq_hi0 + o, idesc, 1);
                                    } else if (passes == 3) {
                                        umma_tf32(d_tmem, dp_lo0 + o, dq_hi0 + o, idesc, acc);
                                        umma_tf32(d_tmem, dp_hi0 + o, dq_lo0 + o, idesc, 1);
                                        umma_tf32(d_tmem, dp_hi0 + o, dq_hi0 + o, idesc, 1);
                                    } else {
                                        umma_tf32(d_tmem, dp_hi0 + o, dq_hi0 + o, idesc, acc);
                                    }
                                }
                            }
                            umma_commit(&bars.empty[s]);
                        }
                        __syncwarp();
                    }
                    if (elect_one()) umma_commit(&bars.tfull[buf]);
                    __syncwarp();
                }
            }
        }
        __syncwarp();
    } else {
        // ============================ TMA producer ============================
        reg_dec_other();
        issue_loads((BF && PRE) ? 0 : -1);
        __syncwarp();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == TC_MMA_WARP) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

// ---------------------------------------------------------------------------------- TN kernel, CTA pairs
// bf16x3 weight gradient with cta_group::2 (M = 256 columns of P per pair): CTA r converts its own 128 columns of P and
// fetches HALF of the pre-split Q k-block (BN/2 columns, hi and lo), which is two thirds of the single-CTA kernel's
// L2 -> SM traffic at BN = 256; a stage is 64 KB (three stages).  Work item = (R slice, anchor group, PAIR of P tiles,
// Q tile); barrier protocol as in tc_gemm_nt_pair_kernel.
template <int BN_>
struct TnPairCfg {
    static constexpr int BN = BN_;
    static constexpr int A_BYTES = 16384;                   // 64 k-rows x 128 columns of P as bf16 (hi or lo)
    static constexpr int QH_BYTES = (BN / 2) * 64 * 2;      // this CTA's half of the Q k-block, hi or lo
    static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * QH_BYTES;
    static constexpr int STAGES = TC_SMEM_BUDGET / STAGE_BYTES > 6 ? 6 : TC_SMEM_BUDGET / STAGE_BYTES;
    static constexpr int TMEM_COLS = 2 * BN;
    static constexpr int HALF = BN / 2;
    static_assert(STAGES >= 3 && BN % 128 == 0, "pair configuration");
};

template <int BN_, bool PRE = false>       // PRE: see tc_gemm_tn_kernel (wide operand P as bf16 planes)
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_gemm_tn_pair_kernel(const __grid_constant__ CUtensorMap map_p, const __grid_constant__ CUtensorMap map_q,
                       const __grid_constant__ CUtensorMap map_q2, int Pw, int Qw, float* __restrict__ C, int ldc, int64_t R,
                       int64_t rows_per_split, int splits, int chunk_kb, TcGather ga, int anchors_per_item,
                       const __grid_constant__ CUtensorMap map_p2, int fast) {
    using Cfg = TnPairCfg<BN_>;
    constexpr int STAGES = Cfg::STAGES, BN = Cfg::BN, KR = 64, UK = 16;
    constexpr uint32_t MN_LBO = 8192, K_SBO = 1024, KSTEP_BYTES = 2048;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
    __shared__ __align__(8) TcBarriers bars;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
    const int pp_tiles = ((Pw + TC_BM - 1) / TC_BM + 1) / 2;
    const int q_tiles = (Qw + BN - 1) / BN;
    const int tiles = pp_tiles * q_tiles;
    const int n_groups = ga.anchors > 0 ? (ga.anchors + anchors_per_item - 1) / anchors_per_item : 1;
    const int items = tiles * n_groups * splits;        // item -> (sp, grp, pair tile), tile fastest

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&bars.raw_full[s], PRE ? (fast ? 2 : 4) : 1);      // PRE: four issuing warps (see the NT pair kernel)
            mbar_init(&bars.full[s], PRE ? 2 : 2 * TC_CONV_WARPS);
            mbar_init(&bars.empty[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&bars.tfull[a], 1);
            mbar_init(&bars.tempty[a], 2 * TC_EPI_WARPS);
        }
        fence_barrier_init();
    }
    if (warp == TC_MMA_WARP) tmem_alloc_pair(&bars.tmem_base, Cfg::TMEM_COLS);
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = bars.tmem_base;

    auto item_nkb = [&](int item, int64_t& r0) {
        const int sp = item / (tiles * n_groups);
        r0 = (int64_t)sp * rows_per_split;
        const int64_t r1 = r0 + rows_per_split < R ? r0 + rows_per_split : R;
        return (int)((r1 - r0 + KR - 1) / KR);
    };
    auto item_anchors = [&](int item, int& a0) {
        if (ga.anchors <= 0) { a0 = 0; return 1; }
        const int grp = (item / tiles) % n_groups;
        a0 = grp * anchors_per_item;
        return min(anchors_per_item, ga.anchors - a0);
    };

    // TMA issue loop of one thread.  which = -1: every box of every stage (raw fp32 P, converter path); PRE: which = 0: P hi,
    // 1: P lo, 2: Q hi, 3: Q lo -- four issuing warps, see the NT pair kernel
    auto issue_loads = [&](int which) {
        if (which <= 0) tma_prefetch_desc(&map_p);
        if (PRE && which == 1) tma_prefetch_desc(&map_p2);
        if (which < 0 || which == 2) tma_prefetch_desc(&map_q);
        if (which < 0 || which == 3) tma_prefetch_desc(&map_q2);
        if (fast && (which == 1 || which == 3)) return;
        const uint32_t tx = which < 0 ? 2u * (uint32_t)Cfg::A_BYTES + (fast ? 1u : 2u) * (uint32_t)Cfg::QH_BYTES
                                      : (which < 2 ? (uint32_t)Cfg::A_BYTES : (uint32_t)Cfg::QH_BYTES);
        int it = 0;
        for (int item = pair; item < items; item += npairs) {
            const int tile = item % tiles;
            const int p0 = (2 * (tile / q_tiles) + (int)rank) * TC_BM;
            const int q0 = (tile % q_tiles) * BN + (int)rank * (BN / 2);
            int64_t r0;
            int a0;
            const int nkb = item_nkb(item, r0);
            const int nblk = nkb * item_anchors(item, a0);
            for (int kb = 0; kb < nblk; ++kb, ++it) {
                const int s = it % STAGES;
                mbar_wait_guard(&bars.empty[s], ((it / STAGES) & 1) ^ 1);
                if (!elect_one()) continue;      // the warp runs the loop, one lane issues (uniform-register operands)
                unsigned char* st = smem_al + (size_t)s * Cfg::STAGE_BYTES;
                unsigned char* sq = st + 2 * Cfg::A_BYTES;
                const int row = (int)(r0 + (int64_t)(kb % nkb) * KR);
                mbar_arrive_expect_tx(&bars.raw_full[s], tx);
                const int an = ga.anchors > 0 ? a0 + kb / nkb : 0;
                if (PRE) {
                    if (which < 2) {
                        const CUtensorMap* mp = which == 0 ? &map_p : &map_p2;
                        unsigned char* dst = st + which * Cfg::A_BYTES;
#pragma unroll
                        for (int a = 0; a < TC_BM / 64; ++a) {
                            const int cc = p0 + a * 64;
                            if (ga.anchors > 0) {
                                const int kk = cc < Pw ? cc / ga.c : 0;
                                const int c0 = cc < Pw ? cc - kk * ga.c : ga.c;                 // out of bounds: zeros
                                const int mid = cc < Pw ? __ldg(ga.table + an * ga.kk_n + kk) : 0;
                                tma_load_3d(dst + a * 8192, mp, c0, mid, row, &bars.raw_full[s]);
                            } else {
                                tma_load_2d(dst + a * 8192, mp, cc, row, &bars.raw_full[s]);
                            }
                        }
                    } else {
                        const CUtensorMap* mq = which == 2 ? &map_q : &map_q2;
                        unsigned char* dst = sq + (which - 2) * Cfg::QH_BYTES;
#pragma unroll
                        for (int a = 0; a < BN / 128; ++a) {
                            if (ga.anchors > 0) tma_load_3d(dst + a * 8192, mq, q0 + a * 64, an, row, &bars.raw_full[s]);
                            else tma_load_2d(dst + a * 8192, mq, q0 + a * 64, row, &bars.raw_full[s]);
                        }
                    }
                    continue;
                }
                if (ga.anchors > 0) {
#pragma unroll
                    for (int a = 0; a < TC_BM / 32; ++a) {
                        const int cc = p0 + a * 32;
                        if (cc < Pw) {
                            const int kk = cc / ga.c;
                            tma_load_3d(st + a * 8192, &map_p, cc - kk * ga.c, __ldg(ga.table + an * ga.kk_n + kk), row, &bars.raw_full[s]);
                        } else {
                            tma_load_3d(st + a * 8192, &map_p, ga.c, 0, row, &bars.raw_full[s]);   // out of bounds: zeros
                        }
                    }
#pragma unroll
                    for (int a = 0; a < BN / 128; ++a) {
                        tma_load_3d(sq + a * 8192, &map_q, q0 + a * 64, an, row, &bars.raw_full[s]);
                        if (!fast) tma_load_3d(sq + Cfg::QH_BYTES + a * 8192, &map_q2, q0 + a * 64, an, row, &bars.raw_full[s]);
                    }
                    continue;
                }
#pragma unroll
                for (int a = 0; a < TC_BM / 32; ++a) tma_load_2d(st + a * 8192, &map_p, p0 + a * 32, row, &bars.raw_full[s]);
#pragma unroll
                for (int a = 0; a < BN / 128; ++a) {
                    tma_load_2d(sq + a * 8192, &map_q, q0 + a * 64, row, &bars.raw_full[s]);
                    if (!fast) tma_load_2d(sq + Cfg::QH_BYTES + a * 8192, &map_q2, q0 + a * 64, row, &bars.raw_full[s]);
                }
            }
        }
    };

    if (warp < TC_EPI_WARPS) {
        // ============================ epilogue ============================
        reg_inc_epi();
        const int q = warp & 3, h = warp >> 2;
        float acc[Cfg::HALF];
#pragma unroll
        for (int i = 0; i < Cfg::HALF; ++i) acc[i] = 0.f;
        int ci = 0;
        for (int item = pair; item < items; item += npairs) {
            const int tile = item % tiles;
            const int p0 = (2 * (tile / q_tiles) + (int)rank) * TC_BM, q0 = (tile % q_tiles) * BN;
            int64_t r0;
            int a0;
            const int nblk = item_nkb(item, r0) * item_anchors(item, a0);
            for (int kb0 = 0; kb0 < nblk; kb0 += chunk_kb, ++ci) {
                const int buf = ci & 1;
                mbar_wait_guard(&bars.tfull[buf], (ci >> 1) & 1);
                tc_fence_after();
#pragma unroll
                for (int j = 0; j < Cfg::HALF / 32; ++j) {
                    float v[32];
                    tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN + h * Cfg::HALF + j * 32), v);
#pragma unroll
                    for (int i = 0; i < 32; ++i) acc[j * 32 + i] += v[i];
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_remote(&bars.tempty[buf], 0);
            }
            const int i = p0 + q * 32 + lane;
            const int c_base = q0 + h * Cfg::HALF;
            if (i < Pw) {
#pragma unroll
                for (int j = 0; j < Cfg::HALF; ++j)
                    if (c_base + j < Qw) atomicAdd(C + (size_t)(c_base + j) * ldc + i, acc[j]);
            }
#pragma unroll
            for (int j = 0; j < Cfg::HALF; ++j) acc[j] = 0.f;
        }
    } else if (warp != TC_MMA_WARP && warp != TC_TMA_WARP) {
        // ============================ converters (own P columns) ============================
        reg_dec_other();
        const int ct = warp < TC_MMA_WARP ? threadIdx.x - TC_EPI_WARPS * 32 : threadIdx.x - (TC_TMA_WARP + 1) * 32 + 128;
        if (PRE) {      // converter warps 0..2 issue boxes (P lo, Q hi, Q lo), warp 3 forwards the barrier, 4 and 5 idle
            const int cw = ct >> 5;
            if (cw < 3) {
                issue_loads(cw + 1);
            } else if (cw == 3) {
                int it = 0;
                for (int item = pair; item < items; item += npairs) {
                    int64_t r0;
                    int a0;
                    const int nblk = item_nkb(item, r0) * item_anchors(item, a0);
                    for (int kb = 0; kb < nblk; ++kb, ++it) {
                        const int s = it % STAGES;
                        mbar_wait_guard(&bars.raw_full[s], (it / STAGES) & 1);
                        if (lane == 0) mbar_arrive_remote(&bars.full[s], 0);
                    }
                }
            }
            __syncwarp();
        } else {
            int it = 0;
            for (int item = pair; item < items; item += npairs) {
                int64_t r0;
                int a0;
                const int nblk = item_nkb(item, r0) * item_anchors(item, a0);
                for (int kb = 0; kb < nblk; ++kb, ++it) {
                    const int s = it % STAGES;
                    mbar_wait_guard(&bars.raw_full[s], (it / STAGES) & 1);
                    const uint32_t st = smem_base + s * Cfg::STAGE_BYTES;
                    for (int task = ct >> 5; task < KR / 4; task += TC_CONV_WARPS) convert_tn_rows4_bf16(st, task * 4, lane);
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_remote(&bars.full[s], 0);
                }
            }
        }
    } else if (warp == TC_MMA_WARP) {
        // ============================ MMA issuer (leader CTA only) ============================
        // whole warp in the loop, one elected lane issues (uniform-register operands: see tc_gemm_nt_pair_kernel)
        reg_dec_other();
        if (rank == 0) {
            const uint32_t idesc = make_idesc_bf16(2 * TC_BM, BN, 1, 1);   // MN-major
            int it = 0, ci = 0;
            for (int item = pair; item < items; item += npairs) {
                int64_t r0;
                int a0;
                const int nkb = item_nkb(item, r0);
                const int nblk = nkb * item_anchors(item, a0);
                const int64_t rows = (r0 + rows_per_split < R ? r0 + rows_per_split : R) - r0;
                for (int kb0 = 0; kb0 < nblk; kb0 += chunk_kb, ++ci) {
                    const int buf = ci & 1;
                    mbar_wait_guard_cluster(&bars.tempty[buf], ((ci >> 1) & 1) ^ 1);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + (uint32_t)(buf * BN);
                    const int kb1 = kb0 + chunk_kb < nblk ? kb0 + chunk_kb : nblk;
                    for (int kb = kb0; kb < kb1; ++kb, ++it) {
                        const int s = it % STAGES;
                        mbar_wait_guard_cluster(&bars.full[s], (it / STAGES) & 1);
                        tc_fence_after();
                        const uint32_t p_hi = smem_base + s * Cfg::STAGE_BYTES;
                        const uint32_t p_lo = p_hi + Cfg::A_BYTES;
                        const uint32_t q_hi = p_lo + Cfg::A_BYTES;
                        const uint32_t q_lo = q_hi + Cfg::QH_BYTES;
                        const int64_t rrem = rows - (int64_t)(kb % nkb) * KR;
                        const int ksteps = rrem >= KR ? KR / UK : (int)((rrem + UK - 1) / UK);
                        const uint64_t dp_hi0 = make_smem_desc(p_hi, MN_LBO, K_SBO, 2), dq_hi0 = make_smem_desc(q_hi, MN_LBO, K_SBO, 2);
                        const uint64_t dp_lo0 = make_smem_desc(p_lo, MN_LBO, K_SBO, 2), dq_lo0 = make_smem_desc(q_lo, MN_LBO, K_SBO, 2);
                        const bool first_kb = kb == kb0;
                        if (elect_one()) {
#pragma unroll
                            for (int ks = 0; ks < KR / UK; ++ks) {
                                if (ks < ksteps) {
                                    const uint64_t o = (uint64_t)(ks * (KSTEP_BYTES >> 4));   // start-address field: address >> 4
                                    const uint32_t acc = (first_kb && ks == 0) ? 0u : 1u;
                                    if (fast) {
                                        umma_bf16_pair(d_tmem, dp_hi0 + o, dq_hi0 + o, idesc, acc);
                                    } else {
                                        umma_bf16_pair(d_tmem, dp_lo0 + o, dq_hi0 + o, idesc, acc);
                                        umma_bf16_pair(d_tmem, dp_hi0 + o, dq_lo0 + o, idesc, 1);
                                        umma_bf16_pair(d_tmem, dp_hi0 + o, dq_hi0 + o, idesc, 1);
                                    }
                                }
                            }
                            umma_commit_pair(&bars.empty[s]);
                        }
                        __syncwarp();
                    }
                    if (elect_one()) umma_commit_pair(&bars.tfull[buf]);
                    __syncwarp();
                }
            }
        }
        __syncwarp();
    } else {
        // ============================ TMA producer (both CTAs: own P columns, own half of Q) ============
        reg_dec_other();
        issue_loads(PRE ? 0 : -1);
        __syncwarp();
    }

    tc_fence_before();
    cluster_sync_all();
    if (warp == TC_MMA_WARP) {
        tc_fence_after();
        tmem_dealloc_pair(tmem_base, Cfg::TMEM_COLS);
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

// row-major matrix [rows, cols] of fp32 (or bf16); box = [box_rows, 128 bytes], 128-byte swizzle span, zero
// fill out of bounds
static int make_map_2d(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int box_rows, CUtensorMapSwizzle swz,
                       bool bf16 = false) {
    EncodeTiledFn enc = get_encoder();
    if (enc == nullptr) {
        set_error("cuTensorMapEncodeTiled not available from the driver");
        return VGTKB_EUNSUP;
    }
    const int esz = bf16 ? 2 : 4;
    const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t gstr[1] = {(cuuint64_t)cols * esz};
    const cuuint32_t box[2] = {(cuuint32_t)(128 / esz), (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(map, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                           const_cast<void*>(base), gdim, gstr, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d)", (int)r);
        return VGTKB_ECUDA;
    }
    return VGTKB_OK;
}

int tc::make_rows_map(TmaMap* out, const void* base, int64_t rows, int64_t cols, int box_rows) {
    static_assert(sizeof(CUtensorMap) == sizeof(TmaMap) && alignof(CUtensorMap) <= alignof(TmaMap), "CUtensorMap layout");
    return make_map_2d(reinterpret_cast<CUtensorMap*>(out), base, rows, cols, box_rows, CU_TENSOR_MAP_SWIZZLE_128B);
}

int tc::make_plane_map(TmaMap* out, const void* base, int64_t rows, int64_t cols, int box_rows) {
    EncodeTiledFn enc = get_encoder();
    if (enc == nullptr) {
        set_error("cuTensorMapEncodeTiled not available from the driver");
        return VGTKB_EUNSUP;
    }
    const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t gstr[1] = {(cuuint64_t)cols * 2};
    const cuuint32_t box[2] = {32u, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(reinterpret_cast<CUtensorMap*>(out), CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim,
                           gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(plane) failed (%d)", (int)r);
        return VGTKB_ECUDA;
    }
    return VGTKB_OK;
}

// X [points, anchors, c] fp32 (or bf16) as a 3-D tensor; box = [box_rows points, 1 anchor, 128 bytes of channels]
static int make_map_3d(CUtensorMap* map, const void* base, int64_t points, int anchors, int c, int box_rows,
                       CUtensorMapSwizzle swz, bool bf16 = false) {
    EncodeTiledFn enc = get_encoder();
    if (enc == nullptr) {
        set_error("cuTensorMapEncodeTiled not available from the driver");
        return VGTKB_EUNSUP;
    }
    const int esz = bf16 ? 2 : 4;
    const cuuint64_t gdim[3] = {(cuuint64_t)c, (cuuint64_t)anchors, (cuuint64_t)points};
    const cuuint64_t gstr[2] = {(cuuint64_t)c * esz, (cuuint64_t)c * esz * anchors};
    const cuuint32_t box[3] = {(cuuint32_t)(128 / esz), 1u, (cuuint32_t)box_rows};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = enc(map, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3,
                           const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(3d) failed (%d)", (int)r);
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

static int default_chunk(int passes, bool bf = false, bool fast = false) {
    if (fast) return 16;   // one MMA per k-step and no fp32-parity claim: long chunks, few epilogue drains
    // k-blocks the tensor core accumulates in TMEM before the epilogue folds the chunk into registers (experiments:
    // VGTKB_CHUNK_KB).  Draining a 128 x 256 fp32 chunk costs about as much as the MMAs of one k-block, so the chunk
    // must span several k-blocks; the round-toward-zero drift grows with the MMAs per chunk (12 per k-block).
    static const int env = getenv("VGTKB_CHUNK_KB") ? atoi(getenv("VGTKB_CHUNK_KB")) : 0;
    if (env > 0) return env;
    if (bf) return 4;      // 48 MMAs per chunk: drift ~1e-6, below the 5e-6 of the bf16x3 products
    return passes == 3 ? 2 : 4;
}

template <int BN, bool BF>
static int launch_nt(int64_t M, int N, int K, const float* A, const void* Bhi, const void* Blo, const float* bias,
                     float* C, int passes, cudaStream_t st, TcGather ga = TcGather{0, 0, 0, nullptr}) {
    using Cfg = TcCfg<BN>;
    CUtensorMap ma, mhi, mlo;
    int rc = ga.anchors > 0 ? make_map_3d(&ma, A, M, ga.anchors, ga.c, TC_BM, CU_TENSOR_MAP_SWIZZLE_128B)
                            : make_map_2d(&ma, A, M, K, TC_BM, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    rc = make_map_2d(&mhi, Bhi, N, K, BN, CU_TENSOR_MAP_SWIZZLE_128B, BF);
    if (rc) return rc;
    rc = make_map_2d(&mlo, Blo, N, K, BN, CU_TENSOR_MAP_SWIZZLE_128B, BF);
    if (rc) return rc;
    const size_t smem = (size_t)Cfg::STAGES * Cfg::STAGE_BYTES + 1024 + TC_EPI_WARPS * 4096;   // + epilogue staging
    auto kern = tc_gemm_nt_kernel<BN, BF>;
    VGTKB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t tiles = ceil_div64(M, TC_BM) * ceil_div(N, BN) * (ga.anchors > 0 ? ga.anchors : 1);
    const int grid = (int)(tiles < num_sms() ? tiles : num_sms());
    kern<<<grid, TC_THREADS, smem, st>>>(ma, mhi, mlo, bias, C, M, N, K, passes, default_chunk(passes, BF), ga);
    return check_launch("gemm_nt(tcgen05)");
}

static int tc_gemm_nt_impl(int64_t M, int N, int K, const float* A, const float* B, const float* bias, float* C, int passes,
                           float* workspace, cudaStream_t st, TcGather ga);

// CTA-pair launch (bf16x3, N > 128): clusters of two CTAs, one pair per two SMs
template <int BN_, bool PRE = false>
static int launch_nt_pair(int64_t M, int N, int K, const void* A, const void* Bhi, const void* Blo, const float* bias, float* C,
                          cudaStream_t st, TcGather ga, const void* A_lo = nullptr, int fast = 0) {
    using Cfg = PairCfg<BN_>;
    CUtensorMap ma, mhi, mlo, mc, ma2;
    int rc = ga.anchors > 0 ? make_map_3d(&ma, A, M, ga.anchors, ga.c, TC_BM, CU_TENSOR_MAP_SWIZZLE_128B, PRE)
                            : make_map_2d(&ma, A, M, K, TC_BM, CU_TENSOR_MAP_SWIZZLE_128B, PRE);
    if (rc) return rc;
    ma2 = ma;
    if (PRE && A_lo != nullptr) {
        rc = ga.anchors > 0 ? make_map_3d(&ma2, A_lo, M, ga.anchors, ga.c, TC_BM, CU_TENSOR_MAP_SWIZZLE_128B, true)
                            : make_map_2d(&ma2, A_lo, M, K, TC_BM, CU_TENSOR_MAP_SWIZZLE_128B, true);
        if (rc) return rc;
    }
    rc = make_map_2d(&mhi, Bhi, N, K, Cfg::BN / 2, CU_TENSOR_MAP_SWIZZLE_128B, true);
    if (rc) return rc;
    rc = make_map_2d(&mlo, Blo, N, K, Cfg::BN / 2, CU_TENSOR_MAP_SWIZZLE_128B, true);
    if (rc) return rc;
    // output through 32x32 tensor-map stores: plain (non-gather) GEMMs whose rows are 16-byte multiples; VGTKB_TMA_EPILOGUE=0
    // keeps the per-lane stores (comparison / fallback)
    static const bool tma_epi = []() { const char* e = getenv("VGTKB_TMA_EPILOGUE"); return e == nullptr || e[0] != '0'; }();
    const int tma_out = tma_epi && ga.anchors == 0 && N % 4 == 0 && (reinterpret_cast<uintptr_t>(C) & 15) == 0;
    mc = ma;
    if (tma_out) {
        rc = make_map_2d(&mc, C, M, N, 32, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc) return rc;
    }
    const size_t smem = (size_t)Cfg::STAGES * Cfg::STAGE_BYTES + 1024 + TC_EPI_WARPS * 4096;
    auto kern = tc_gemm_nt_pair_kernel<BN_, PRE>;
    VGTKB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t pair_tiles = ceil_div64(ceil_div64(M, TC_BM), 2) * ceil_div(N, Cfg::BN) * (ga.anchors > 0 ? ga.anchors : 1);
    const int max_pairs = num_sms() / 2;
    const int pairs = (int)(pair_tiles < max_pairs ? pair_tiles : max_pairs);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * pairs);
    cfg.blockDim = dim3(TC_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    static const int dbg = getenv("VGTKB_DBG") ? atoi(getenv("VGTKB_DBG")) : 0;      // timing experiments only
    VGTKB_CUDA(cudaLaunchKernelEx(&cfg, kern, ma, mhi, mlo, mc, tma_out, bias, C, M, N, K, default_chunk(3, true, fast != 0), ga, ma2,
                                  fast | (dbg << 4)));
    return check_launch("gemm_nt(tcgen05, cta pairs)");
}

int tc_gemm_nt(int64_t M, int N, int K, const float* A, const float* B, const float* bias, float* C, int passes,
               float* workspace, cudaStream_t st) {
    return tc_gemm_nt_impl(M, N, K, A, B, bias, C, passes, workspace, st, TcGather{0, 0, 0, nullptr});
}

// C[(pt, an), n] = sum_{kk, c} X[pt, table[an, kk], c] * B[n, kk*c_n + c]   (X: [points, anchors, c_n])
int tc_gemm_nt_gather(int64_t points, int anchors, int kk_n, int c_n, int N, const int32_t* table, const float* X,
                      const float* B, const float* bias, float* C, int passes, float* workspace, cudaStream_t st) {
    if (c_n % 64 != 0 || anchors < 1 || kk_n < 1) return VGTKB_EUNSUP;   // a k-block must not straddle two neighbours
    return tc_gemm_nt_impl(points, N, kk_n * c_n, X, B, bias, C, passes, workspace, st, TcGather{anchors, kk_n, c_n, table});
}

static int tc_gemm_nt_impl(int64_t M, int N, int K, const float* A, const float* B, const float* bias, float* C, int passes,
                           float* workspace, cudaStream_t st, TcGather ga) {
    // shapes the tensor-core path takes; everything else falls back to the FFMA kernel
    if (K % 4 != 0 || K < 8 || M < 1 || ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B)) & 15) != 0)
        return VGTKB_EUNSUP;
    if (bias != nullptr && (reinterpret_cast<uintptr_t>(bias) & 15) != 0) return VGTKB_EUNSUP;
    if (M >= (int64_t)1 << 31) return VGTKB_EUNSUP;
    // B == NULL: `workspace` already holds the bf16 hi | lo planes of B (vgtkb_weight_planes): bf16 modes only
    const bool presplit = B == nullptr;
    if (presplit && (workspace == nullptr || (reinterpret_cast<uintptr_t>(workspace) & 15) != 0 || K % 8 != 0 ||
                     (passes != 6 && passes != 7)))
        return VGTKB_EUNSUP;
    const int64_t nb = (int64_t)N * K;
    float* ws = nullptr;
    float* owned = nullptr;
    const float* Bhi = B;
    const float* Blo = B;
    const int fast = passes == 7;                 // 7 = bf16 single pass (mode 4)
    if (fast) passes = 6;
    const bool bf = passes == 6 && K % 8 == 0;   // 6 = bf16x3 (needs 16-byte aligned bf16 rows)
    if (passes == 6) passes = 3;                  // K % 8 != 0: 3xTF32 instead
    if (passes == 3) {
        ws = workspace;
        if (!presplit && (ws == nullptr || (reinterpret_cast<uintptr_t>(ws) & 15) != 0 || nb % 4 != 0)) {
            VGTKB_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&owned), sizeof(float) * 2 * (size_t)nb, st));
            ws = owned;
        }
        const int blocks = (int)(ceil_div64(nb, 256) < 1184 ? ceil_div64(nb, 256) : 1184);
        // bf16: hi | lo planes contiguous (N*K floats in all -- callers such as vgtkb_inter_conv_backward size their scratch
        // for that; the lo plane used to start at float offset N*K, which overran such a scratch by N*K floats whenever the
        // activation operand came without planes); TF32: two fp32 arrays, 2*N*K floats
        float* lo_at = bf ? reinterpret_cast<float*>(reinterpret_cast<uint16_t*>(ws) + nb) : ws + nb;
        if (presplit) {}
        else if (bf) split_bf16_kernel<<<blocks, 256, 0, st>>>(nb, B, reinterpret_cast<uint16_t*>(ws), reinterpret_cast<uint16_t*>(lo_at));
        else split_tf32_kernel<<<blocks, 256, 0, st>>>(nb, B, ws, lo_at);
        Bhi = ws;
        Blo = lo_at;
    }
    int rc;
    // CTA pairs (cta_group::2) are the default for the bf16x3 mode; VGTKB_CTA_PAIRS=0 selects the single-CTA kernels
    static const int use_pairs = getenv("VGTKB_CTA_PAIRS") ? atoi(getenv("VGTKB_CTA_PAIRS")) : 1;
    if (bf && (use_pairs || fast)) {
        if (N <= 64) rc = launch_nt_pair<64>(M, N, K, A, Bhi, Blo, bias, C, st, ga, nullptr, fast);
        else if (N <= 128) rc = launch_nt_pair<128>(M, N, K, A, Bhi, Blo, bias, C, st, ga, nullptr, fast);
        else rc = launch_nt_pair<256>(M, N, K, A, Bhi, Blo, bias, C, st, ga, nullptr, fast);
    } else if (bf) {
        if (N <= 64) rc = launch_nt<64, true>(M, N, K, A, Bhi, Blo, bias, C, passes, st, ga);
        else if (N <= 128) rc = launch_nt<128, true>(M, N, K, A, Bhi, Blo, bias, C, passes, st, ga);
        else rc = launch_nt<256, true>(M, N, K, A, Bhi, Blo, bias, C, passes, st, ga);
    } else if (N <= 64) rc = launch_nt<64, false>(M, N, K, A, Bhi, Blo, bias, C, passes, st, ga);
    else if (N <= 128) rc = launch_nt<128, false>(M, N, K, A, Bhi, Blo, bias, C, passes, st, ga);
    else rc = launch_nt<256, false>(M, N, K, A, Bhi, Blo, bias, C, passes, st, ga);
    if (owned != nullptr) cudaFreeAsync(owned, st);
    return rc;
}

// Work decomposition of the weight-gradient kernels: items = tiles x anchor groups x R splits are dealt round-robin to
// `units` persistent CTAs (or CTA pairs).  The first version took ceil(2 units / tiles) splits, which lands just ABOVE two
// full waves for the shapes of the backbone (e.g. 6 tiles x 25 splits = 150 items on 74 pairs: a third wave with two
// items, 68 % efficiency; ncu: tensor pipe 37-74 % on kernels whose operands stream at half the fabric rate).  Pick the
// (groups, splits) pair that fills whole waves best, at most ~3 waves, every split at least 512 rows long.
static void tn_decompose(int tiles, int units, int anchors, int64_t R, int KR, int& n_groups, int& ag, int64_t& splits, int64_t& rps) {
    const int64_t max_splits = ceil_div64(R, 512) > 0 ? ceil_div64(R, 512) : 1;
    const char* legacy = getenv("VGTKB_TN_DECOMP");
    if (legacy != nullptr && legacy[0] == '0') {      // the first version (comparison runs)
        n_groups = 1, ag = 1;
        if (anchors > 0) {
            n_groups = (int)ceil_div64((int64_t)2 * units, tiles);
            if (n_groups > anchors) n_groups = anchors;
            ag = ceil_div(anchors, n_groups);
            n_groups = ceil_div(anchors, ag);
        }
        splits = ceil_div64((int64_t)2 * units, (int64_t)tiles * n_groups);
        if (splits > max_splits) splits = max_splits;
        if (splits < 1) splits = 1;
        rps = ceil_div64(ceil_div64(R, splits), KR) * KR;
        splits = ceil_div64(R, rps);
        return;
    }
    double best = -1.0;
    n_groups = 1, ag = anchors > 0 ? anchors : 1, splits = 1, rps = ceil_div64(R, KR) * KR;
    const int ng_max = anchors > 0 ? anchors : 1;
    for (int ng = 1; ng <= ng_max; ++ng) {
        const int a_per = anchors > 0 ? ceil_div(anchors, ng) : 1;
        const int ng_eff = anchors > 0 ? ceil_div(anchors, a_per) : 1;
        if (ng_eff != ng) continue;                                   // same decomposition as a smaller ng
        for (int64_t sp = 1; sp <= max_splits && (int64_t)tiles * ng * sp <= (int64_t)3 * units + tiles; ++sp) {
            const int64_t r = ceil_div64(ceil_div64(R, sp), KR) * KR;
            const int64_t sp_eff = ceil_div64(R, r);
            if (sp_eff != sp) continue;
            const int64_t items = (int64_t)tiles * ng * sp;
            const int64_t waves = ceil_div64(items, units);
            // efficiency of the last wave, with a mild preference for more (shorter) items: better balance, shorter tails
            const double eff = (double)items / (double)(waves * units) - (items < units ? 0.0 : 0.02 / (double)waves);
            if (eff > best + 1e-9) {
                best = eff;
                n_groups = ng, ag = a_per, splits = sp, rps = r;
            }
        }
    }
}

template <int BN, bool BF, bool PRE = false>
static int launch_tn(const void* P, int Pw, const void* Q, const void* Q2, int Qw, float* C, int ldc, int64_t R, int passes,
                     cudaStream_t st, TcGather ga = TcGather{0, 0, 0, nullptr}, const void* P_lo = nullptr, int fast = 0) {
    using Cfg = TcCfg<BN>;
    constexpr int KR = BF ? 64 : TC_BK;
    CUtensorMap mp, mq, mq2, mp2;
    const CUtensorMapSwizzle swz = BF ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
    int rc;
    if (ga.anchors > 0) {   // R = points; P = X [points, anchors, c]; Q = Y [points, anchors, Qw]
        rc = make_map_3d(&mp, P, R, ga.anchors, ga.c, KR, swz, PRE);
        if (rc) return rc;
        if (PRE && P_lo != nullptr) {
            rc = make_map_3d(&mp2, P_lo, R, ga.anchors, ga.c, KR, swz, true);
            if (rc) return rc;
        }
        rc = make_map_3d(&mq, Q, R, ga.anchors, Qw, KR, swz, BF);
        if (rc) return rc;
        rc = make_map_3d(&mq2, Q2, R, ga.anchors, Qw, KR, swz, BF);
        if (rc) return rc;
    } else {
        rc = make_map_2d(&mp, P, R, Pw, KR, swz, PRE);
        if (rc) return rc;
        if (PRE && P_lo != nullptr) {
            rc = make_map_2d(&mp2, P_lo, R, Pw, KR, swz, true);
            if (rc) return rc;
        }
        rc = make_map_2d(&mq, Q, R, Qw, KR, swz, BF);      // BF: bf16 hi, box = 64 columns x 64 rows
        if (rc) return rc;
        rc = make_map_2d(&mq2, Q2, R, Qw, KR, swz, BF);    // BF: bf16 lo (tf32 path: unused duplicate)
        if (rc) return rc;
    }
    const int tiles = ceil_div(Pw, TC_BM) * ceil_div(Qw, BN);
    int n_groups = 1, ag = 1;
    int64_t splits = 1, rps = R;
    tn_decompose(tiles, num_sms(), ga.anchors, R, KR, n_groups, ag, splits, rps);
    const int64_t items = splits * tiles * n_groups;
    const size_t smem = (size_t)Cfg::STAGES * Cfg::STAGE_BYTES + 1024;
    if (!PRE || P_lo == nullptr) mp2 = mp;
    auto kern = tc_gemm_tn_kernel<BN, BF, PRE>;
    VGTKB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = (int)(items < num_sms() ? items : num_sms());
    kern<<<grid, TC_THREADS, smem, st>>>(mp, mq, mq2, Pw, Qw, C, ldc, R, rps, (int)splits, passes, default_chunk(passes, BF, fast != 0),
                                         ga, ag, mp2, fast);
    return check_launch("gemm_tn(tcgen05)");
}

// CTA-pair launch of the bf16x3 weight-gradient kernel (Q tile 128 or 256)
template <int BN_, bool PRE = false>
static int launch_tn_pair(const void* P, int Pw, const void* Q, const void* Q2, int Qw, float* C, int ldc, int64_t R,
                          cudaStream_t st, TcGather ga = TcGather{0, 0, 0, nullptr}, const void* P_lo = nullptr, int fast = 0) {
    using Cfg = TnPairCfg<BN_>;
    constexpr int KR = 64;
    CUtensorMap mp, mq, mq2, mp2;
    const CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B;
    int rc;
    if (ga.anchors > 0) {
        rc = make_map_3d(&mp, P, R, ga.anchors, ga.c, KR, swz, PRE);
        if (rc) return rc;
        if (PRE && P_lo != nullptr) {
            rc = make_map_3d(&mp2, P_lo, R, ga.anchors, ga.c, KR, swz, true);
            if (rc) return rc;
        }
        rc = make_map_3d(&mq, Q, R, ga.anchors, Qw, KR, swz, true);
        if (rc) return rc;
        rc = make_map_3d(&mq2, Q2, R, ga.anchors, Qw, KR, swz, true);
        if (rc) return rc;
    } else {
        rc = make_map_2d(&mp, P, R, Pw, KR, swz, PRE);
        if (rc) return rc;
        if (PRE && P_lo != nullptr) {
            rc = make_map_2d(&mp2, P_lo, R, Pw, KR, swz, true);
            if (rc) return rc;
        }
        rc = make_map_2d(&mq, Q, R, Qw, KR, swz, true);
        if (rc) return rc;
        rc = make_map_2d(&mq2, Q2, R, Qw, KR, swz, true);
        if (rc) return rc;
    }
    const int max_pairs = num_sms() / 2;
    const int tiles = ceil_div(ceil_div(Pw, TC_BM), 2) * ceil_div(Qw, Cfg::BN);
    int n_groups = 1, ag = 1;
    int64_t splits = 1, rps = R;
    tn_decompose(tiles, max_pairs, ga.anchors, R, KR, n_groups, ag, splits, rps);
    const int64_t items = splits * tiles * n_groups;
    const size_t smem = (size_t)Cfg::STAGES * Cfg::STAGE_BYTES + 1024;
    if (!PRE || P_lo == nullptr) mp2 = mp;
    auto kern = tc_gemm_tn_pair_kernel<BN_, PRE>;
    VGTKB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int pairs = (int)(items < max_pairs ? items : max_pairs);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * pairs);
    cfg.blockDim = dim3(TC_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    VGTKB_CUDA(cudaLaunchKernelEx(&cfg, kern, mp, mq, mq2, Pw, Qw, C, ldc, R, rps, (int)splits, default_chunk(3, true, fast != 0), ga, ag,
                                  mp2, fast));
    return check_launch("gemm_tn(tcgen05, cta pairs)");
}

static bool tn_pairs_enabled() {
    static const int v = getenv("VGTKB_CTA_PAIRS_TN") ? atoi(getenv("VGTKB_CTA_PAIRS_TN"))
                         : (getenv("VGTKB_CTA_PAIRS") ? atoi(getenv("VGTKB_CTA_PAIRS")) : 1);
    return v != 0;
}

// C[M,N] (+)= A[R,M]^T B[R,N]:  P = B (tiles of 128 over N), Q = A (tiles of <= 256 over M)
// `workspace` (bf16x3 only): R*M floats, holds the bf16 hi/lo split of A.
int tc_gemm_tn(int M, int N, int64_t R, const float* A, const float* B, float* C, int accumulate, int passes,
               float* workspace, cudaStream_t st) {
    if (M % 4 != 0 || N % 4 != 0 || R < 64 || R >= ((int64_t)1 << 31) ||
        ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B)) & 15) != 0)
        return VGTKB_EUNSUP;
    if ((int64_t)M * N < 64 * 64 / 4) return VGTKB_EUNSUP;   // tiny outputs: the FFMA split-R kernel is fine
    const int fast = passes == 7;                             // 7 = bf16 single pass (mode 4)
    if (fast) passes = 6;
    if (passes == 6 && (M % 8 != 0 || workspace == nullptr || (reinterpret_cast<uintptr_t>(workspace) & 15) != 0))
        passes = 3;                                           // bf16 rows must be 16-byte aligned: 3xTF32 instead
    if (!accumulate) VGTKB_CUDA(cudaMemsetAsync(C, 0, sizeof(float) * (size_t)M * N, st));
    if (passes == 6) {   // bf16x3 / bf16
        const int64_t na = R * (int64_t)M;
        uint16_t* hi = reinterpret_cast<uint16_t*>(workspace);
        uint16_t* lo = hi + na;
        const int blocks = (int)(ceil_div64(na, 256) < 2368 ? ceil_div64(na, 256) : 2368);
        split_bf16_kernel<<<blocks, 256, 0, st>>>(na, A, hi, lo);
        const TcGather none{0, 0, 0, nullptr};
        if (M <= 64) return launch_tn<64, true>(B, N, hi, lo, M, C, N, R, 3, st, none, nullptr, fast);
        if (tn_pairs_enabled() && N > TC_BM)
            return M <= 128 ? launch_tn_pair<128>(B, N, hi, lo, M, C, N, R, st, none, nullptr, fast)
                            : launch_tn_pair<256>(B, N, hi, lo, M, C, N, R, st, none, nullptr, fast);
        if (M <= 128) return launch_tn<128, true>(B, N, hi, lo, M, C, N, R, 3, st, none, nullptr, fast);
        return launch_tn<256, true>(B, N, hi, lo, M, C, N, R, 3, st, none, nullptr, fast);
    }
    if (M <= 64) return launch_tn<64, false>(B, N, A, A, M, C, N, R, passes, st);
    if (M <= 128) return launch_tn<128, false>(B, N, A, A, M, C, N, R, passes, st);
    return launch_tn<256, false>(B, N, A, A, M, C, N, R, passes, st);
}

// Weight gradient of the gather-GEMM:  C[m, kk*c_n + c] (+)= sum_{pt, an} Y[(pt, an), m] * X[pt, table[an, kk], c]
// X: [points, anchors, c_n], Y: [points, anchors, M].  `workspace` (bf16x3): points*anchors*M floats.
int tc_gemm_tn_gather(int64_t points, int anchors, int kk_n, int c_n, int M, const int32_t* table, const float* X,
                      const float* Y, float* C, int accumulate, int passes, float* workspace, cudaStream_t st) {
    const int N = kk_n * c_n;
    if (c_n % 32 != 0 || M % 4 != 0 || points < 64 || points >= ((int64_t)1 << 31) ||
        ((reinterpret_cast<uintptr_t>(X) | reinterpret_cast<uintptr_t>(Y)) & 15) != 0)
        return VGTKB_EUNSUP;
    const int fast = passes == 7;
    if (fast) passes = 6;
    if (passes == 6 && (M % 8 != 0 || workspace == nullptr || (reinterpret_cast<uintptr_t>(workspace) & 15) != 0)) passes = 3;
    if (!accumulate) VGTKB_CUDA(cudaMemsetAsync(C, 0, sizeof(float) * (size_t)M * N, st));
    const TcGather ga{anchors, kk_n, c_n, table};
    if (passes == 6) {
        const int64_t na = points * anchors * (int64_t)M;
        uint16_t* hi = reinterpret_cast<uint16_t*>(workspace);
        uint16_t* lo = hi + na;
        const int blocks = (int)(ceil_div64(na, 256) < 2368 ? ceil_div64(na, 256) : 2368);
        split_bf16_kernel<<<blocks, 256, 0, st>>>(na, Y, hi, lo);
        if (M <= 64) return launch_tn<64, true>(X, N, hi, lo, M, C, N, points, 3, st, ga, nullptr, fast);
        if (tn_pairs_enabled() && N > TC_BM)
            return M <= 128 ? launch_tn_pair<128>(X, N, hi, lo, M, C, N, points, st, ga, nullptr, fast)
                            : launch_tn_pair<256>(X, N, hi, lo, M, C, N, points, st, ga, nullptr, fast);
        if (M <= 128) return launch_tn<128, true>(X, N, hi, lo, M, C, N, points, 3, st, ga, nullptr, fast);
        return launch_tn<256, true>(X, N, hi, lo, M, C, N, points, 3, st, ga, nullptr, fast);
    }
    if (M <= 64) return launch_tn<64, false>(X, N, Y, Y, M, C, N, points, passes, st, ga);
    if (M <= 128) return launch_tn<128, false>(X, N, Y, Y, M, C, N, points, passes, st, ga);
    return launch_tn<256, false>(X, N, Y, Y, M, C, N, points, passes, st, ga);
}

}  // namespace vgtkb

// ---------------------------------------------------------------------------------- pre-split (plane) operands
// C[M,N] = (A_hi + A_lo)[M,K] * B[N,K]^T (+ bias) with A given as two bf16 planes (hi = bf16_rn(a), lo = bf16_rn(a - hi):
// vgtkb_split_bf16, or written directly by the producer -- vgtkb_inter_conv_forward's grouping kernel): the same bf16x3
// arithmetic as mode 3 of vgtkb_gemm_nt (bit-identical results), without the in-kernel operand conversion.

extern "C" int vgtkb_split_bf16(int64_t n, const float* x, void* hi, void* lo, void* stream) {
    using namespace vgtkb;
    VGTKB_REQUIRE(n >= 0, "split_bf16: bad size");
    if (n == 0) return VGTKB_OK;
    const int blocks = (int)(ceil_div64(n, 256) < 1184 ? ceil_div64(n, 256) : 1184);
    split_bf16_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(n, x, reinterpret_cast<uint16_t*>(hi), reinterpret_cast<uint16_t*>(lo));
    return check_launch("split_bf16");
}

namespace vgtkb {
// C[M,N] = (a_hi + a_lo)[M,K] * B[N,K]^T (+ bias); workspace: N*K floats (bf16 hi/lo split of B).  VGTKB_EUNSUP for
// shapes the CTA-pair kernel does not take.
int tc_gemm_nt_planes(int64_t M, int N, int K, const void* a_hi, const void* a_lo, const float* B, const float* bias, float* C,
                      float* workspace, cudaStream_t st, int fast) {
    if (M < 1 || N < 1 || K < 64 || K % 8 != 0 || M >= ((int64_t)1 << 31) || workspace == nullptr || a_hi == nullptr ||
        (a_lo == nullptr && !fast) ||
        ((reinterpret_cast<uintptr_t>(a_hi) | reinterpret_cast<uintptr_t>(a_lo) | reinterpret_cast<uintptr_t>(B) |
          reinterpret_cast<uintptr_t>(workspace)) & 15) != 0)
        return VGTKB_EUNSUP;
    const int64_t nb = (int64_t)N * K;
    const int blocks = (int)(ceil_div64(nb, 256) < 1184 ? ceil_div64(nb, 256) : 1184);
    uint16_t* bhi = reinterpret_cast<uint16_t*>(workspace);
    uint16_t* blo = bhi + nb;
    if (B != nullptr) split_bf16_kernel<<<blocks, 256, 0, st>>>(nb, B, bhi, blo);   // NULL: the planes are already there
    const TcGather none{0, 0, 0, nullptr};
    if (N <= 64) return launch_nt_pair<64, true>(M, N, K, a_hi, bhi, blo, bias, C, st, none, a_lo, fast);
    if (N <= 128) return launch_nt_pair<128, true>(M, N, K, a_hi, bhi, blo, bias, C, st, none, a_lo, fast);
    return launch_nt_pair<256, true>(M, N, K, a_hi, bhi, blo, bias, C, st, none, a_lo, fast);
}

// split `n` floats into bf16 hi / lo planes at the start of `workspace` unless the planes are given
static int planes_or_split(const float* x, const void* x_hi, const void* x_lo, int64_t n, float* workspace, const void** hi,
                           const void** lo, cudaStream_t st, int fast = 0) {
    if (x_hi != nullptr && (x_lo != nullptr || fast)) {
        *hi = x_hi;
        *lo = x_lo;
        return VGTKB_OK;
    }
    if (x == nullptr || workspace == nullptr || (reinterpret_cast<uintptr_t>(workspace) & 15) != 0) return VGTKB_EUNSUP;
    uint16_t* h = reinterpret_cast<uint16_t*>(workspace);
    uint16_t* l = h + n;
    const int blocks = (int)(ceil_div64(n, 256) < 2368 ? ceil_div64(n, 256) : 2368);
    split_bf16_kernel<<<blocks, 256, 0, st>>>(n, x, h, l);
    *hi = h;
    *lo = l;
    return VGTKB_OK;
}

// C[M,N] (+)= A[R,M]^T * B[R,N]; the narrow operand A as planes (a_hi / a_lo) or fp32 (split into workspace: R*M floats);
// the wide operand B as planes (b_hi / b_lo: pure TMA -> MMA stream) or fp32 (B: converted in the kernel)
int tc_gemm_tn_planes(int M, int N, int64_t R, const float* A, const void* a_hi, const void* a_lo, const float* B,
                      const void* b_hi, const void* b_lo, float* C, int accumulate, float* workspace, cudaStream_t st, int fast) {
    const bool bpl = b_hi != nullptr && (b_lo != nullptr || fast);
    if (M < 8 || M % 8 != 0 || N < 64 || N % 8 != 0 || R < 64 || R >= ((int64_t)1 << 31) ||
        ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(a_hi) | reinterpret_cast<uintptr_t>(a_lo) |
          reinterpret_cast<uintptr_t>(B) | reinterpret_cast<uintptr_t>(b_hi) | reinterpret_cast<uintptr_t>(b_lo)) & 15) != 0 ||
        (!bpl && B == nullptr))
        return VGTKB_EUNSUP;
    const void *hi, *lo;
    const int rc = planes_or_split(A, a_hi, a_lo, R * (int64_t)M, workspace, &hi, &lo, st, fast);
    if (rc) return rc;
    if (lo == nullptr) lo = hi;                       // fast: never loaded, only a valid tensor map is needed
    if (!accumulate) VGTKB_CUDA(cudaMemsetAsync(C, 0, sizeof(float) * (size_t)M * N, st));
    const TcGather none{0, 0, 0, nullptr};
    if (bpl) {
        if (M <= 64) return launch_tn<64, true, true>(b_hi, N, hi, lo, M, C, N, R, 3, st, none, b_lo, fast);
        if (N > TC_BM)
            return M <= 128 ? launch_tn_pair<128, true>(b_hi, N, hi, lo, M, C, N, R, st, none, b_lo, fast)
                            : launch_tn_pair<256, true>(b_hi, N, hi, lo, M, C, N, R, st, none, b_lo, fast);
        if (M <= 128) return launch_tn<128, true, true>(b_hi, N, hi, lo, M, C, N, R, 3, st, none, b_lo, fast);
        return launch_tn<256, true, true>(b_hi, N, hi, lo, M, C, N, R, 3, st, none, b_lo, fast);
    }
    if (M <= 64) return launch_tn<64, true>(B, N, hi, lo, M, C, N, R, 3, st, none, nullptr, fast);
    if (tn_pairs_enabled() && N > TC_BM)
        return M <= 128 ? launch_tn_pair<128>(B, N, hi, lo, M, C, N, R, st, none, nullptr, fast)
                        : launch_tn_pair<256>(B, N, hi, lo, M, C, N, R, st, none, nullptr, fast);
    if (M <= 128) return launch_tn<128, true>(B, N, hi, lo, M, C, N, R, 3, st, none, nullptr, fast);
    return launch_tn<256, true>(B, N, hi, lo, M, C, N, R, 3, st, none, nullptr, fast);
}

// gather-GEMM (intra conv) with the activation operand X [points, anchors, c_n] as bf16 planes; workspace: N*kk_n*c_n floats
int tc_gemm_nt_gather_planes(int64_t points, int anchors, int kk_n, int c_n, int N, const int32_t* table, const void* x_hi,
                             const void* x_lo, const float* B, const float* bias, float* C, float* workspace, cudaStream_t st) {
    const int K = kk_n * c_n;
    if (c_n % 64 != 0 || anchors < 1 || kk_n < 1 || points < 1 || points >= ((int64_t)1 << 31) || workspace == nullptr ||
        ((reinterpret_cast<uintptr_t>(x_hi) | reinterpret_cast<uintptr_t>(x_lo) | reinterpret_cast<uintptr_t>(B) |
          reinterpret_cast<uintptr_t>(workspace)) & 15) != 0 ||
        (bias != nullptr && (reinterpret_cast<uintptr_t>(bias) & 15) != 0))
        return VGTKB_EUNSUP;
    const int64_t nb = (int64_t)N * K;
    const int blocks = (int)(ceil_div64(nb, 256) < 1184 ? ceil_div64(nb, 256) : 1184);
    uint16_t* bhi = reinterpret_cast<uint16_t*>(workspace);
    uint16_t* blo = bhi + nb;
    if (B != nullptr) split_bf16_kernel<<<blocks, 256, 0, st>>>(nb, B, bhi, blo);   // NULL: the planes are already there
    const TcGather ga{anchors, kk_n, c_n, table};
    if (N <= 64) return launch_nt_pair<64, true>(points, N, K, x_hi, bhi, blo, bias, C, st, ga, x_lo);
    if (N <= 128) return launch_nt_pair<128, true>(points, N, K, x_hi, bhi, blo, bias, C, st, ga, x_lo);
    return launch_nt_pair<256, true>(points, N, K, x_hi, bhi, blo, bias, C, st, ga, x_lo);
}

// weight gradient of the gather-GEMM with X as planes; Y [points*anchors, M] as planes or fp32 (split into workspace)
int tc_gemm_tn_gather_planes(int64_t points, int anchors, int kk_n, int c_n, int M, const int32_t* table, const void* x_hi,
                             const void* x_lo, const float* Y, const void* y_hi, const void* y_lo, float* C, int accumulate,
                             float* workspace, cudaStream_t st) {
    const int N = kk_n * c_n;
    if (c_n % 64 != 0 || M % 8 != 0 || M < 8 || points < 64 || points >= ((int64_t)1 << 31) ||
        ((reinterpret_cast<uintptr_t>(x_hi) | reinterpret_cast<uintptr_t>(x_lo) | reinterpret_cast<uintptr_t>(Y) |
          reinterpret_cast<uintptr_t>(y_hi) | reinterpret_cast<uintptr_t>(y_lo)) & 15) != 0)
        return VGTKB_EUNSUP;
    const void *hi, *lo;
    const int rc = planes_or_split(Y, y_hi, y_lo, points * anchors * (int64_t)M, workspace, &hi, &lo, st);
    if (rc) return rc;
    if (!accumulate) VGTKB_CUDA(cudaMemsetAsync(C, 0, sizeof(float) * (size_t)M * N, st));
    const TcGather ga{anchors, kk_n, c_n, table};
    if (M <= 64) return launch_tn<64, true, true>(x_hi, N, hi, lo, M, C, N, points, 3, st, ga, x_lo);
    if (N > TC_BM)
        return M <= 128 ? launch_tn_pair<128, true>(x_hi, N, hi, lo, M, C, N, points, st, ga, x_lo)
                        : launch_tn_pair<256, true>(x_hi, N, hi, lo, M, C, N, points, st, ga, x_lo);
    if (M <= 128) return launch_tn<128, true, true>(x_hi, N, hi, lo, M, C, N, points, 3, st, ga, x_lo);
    return launch_tn<256, true, true>(x_hi, N, hi, lo, M, C, N, points, 3, st, ga, x_lo);
}
}  // namespace vgtkb

extern "C" int vgtkb_gemm_nt_presplit(int64_t M, int N, int K, const void* a_hi, const void* a_lo, const float* B,
                                      const float* bias, float* C, float* workspace, void* stream) {
    using namespace vgtkb;
    VGTKB_REQUIRE(M >= 1 && N >= 1 && K >= 64, "gemm_nt_presplit: bad size");
    const int rc = tc_gemm_nt_planes(M, N, K, a_hi, a_lo, B, bias, C, workspace, (cudaStream_t)stream, 0);
    if (rc == VGTKB_EUNSUP)
        set_error("gemm_nt_presplit: needs K %% 8 == 0, M < 2^31, 16-byte aligned operands and a workspace of N*K floats");
    return rc;
}

// weight-gradient counterpart: C [M, N] (+)= A [R, M]^T * (b_hi + b_lo) [R, N]; A (narrow, fp32) is split here as in
// vgtkb_gemm_tn, the wide operand arrives as bf16 planes.  workspace: R*M floats.
extern "C" int vgtkb_gemm_tn_presplit(int M, int N, int64_t R, const float* A, const void* b_hi, const void* b_lo, float* C,
                                      int accumulate, float* workspace, void* stream) {
    return vgtkb_gemm_tn_planes(M, N, R, A, nullptr, nullptr, nullptr, b_hi, b_lo, C, accumulate, workspace, stream);
}

extern "C" int vgtkb_gemm_tn_planes(int M, int N, int64_t R, const float* A, const void* a_hi, const void* a_lo, const float* B,
                                    const void* b_hi, const void* b_lo, float* C, int accumulate, float* workspace, void* stream) {
    using namespace vgtkb;
    const int rc = tc_gemm_tn_planes(M, N, R, A, a_hi, a_lo, B, b_hi, b_lo, C, accumulate, workspace, (cudaStream_t)stream, 0);
    if (rc == VGTKB_EUNSUP)
        set_error("gemm_tn_planes: needs M %% 8 == 0, M >= 8, N %% 8 == 0, N >= 64, R >= 64, 16-byte aligned operands, "
                  "and either planes or fp32 + a workspace of R*M floats for the narrow operand");
    return rc;
}

extern "C" int vgtkb_gather_gemm_nt_planes(int64_t points, int anchors, int kk, int c, int n, const int32_t* table,
                                           const void* x_hi, const void* x_lo, const float* w, const float* bias, float* out,
                                           float* workspace, void* stream) {
    using namespace vgtkb;
    VGTKB_REQUIRE(points >= 0 && anchors > 0 && kk > 0 && c > 0 && n > 0, "gather_gemm_nt_planes: bad size");
    if (points == 0) return VGTKB_OK;
    const int rc = tc_gemm_nt_gather_planes(points, anchors, kk, c, n, table, x_hi, x_lo, w, bias, out, workspace,
                                            (cudaStream_t)stream);
    if (rc == VGTKB_EUNSUP) set_error("gather_gemm_nt_planes: unsupported shape (needs c %% 64 == 0, 16-byte aligned operands, workspace)");
    return rc;
}

extern "C" int vgtkb_gather_gemm_tn_planes(int64_t points, int anchors, int kk, int c, int m, const int32_t* table,
                                           const void* x_hi, const void* x_lo, const float* y, const void* y_hi, const void* y_lo,
                                           float* out, int accumulate, float* workspace, void* stream) {
    using namespace vgtkb;
    VGTKB_REQUIRE(points >= 0 && anchors > 0 && kk > 0 && c > 0 && m > 0, "gather_gemm_tn_planes: bad size");
    const int rc = tc_gemm_tn_gather_planes(points, anchors, kk, c, m, table, x_hi, x_lo, y, y_hi, y_lo, out, accumulate, workspace,
                                            (cudaStream_t)stream);
    if (rc == VGTKB_EUNSUP)
        set_error("gather_gemm_tn_planes: unsupported shape (needs c %% 64 == 0, m %% 8 == 0, points >= 64, aligned operands)");
    return rc;
}

// ---------------------------------------------------------------------------------- weight planes, one launch per step
// Every contraction of the block takes its weight operand as two bf16 planes in its own index order: the forward conv
// W [co, (k, c)] (the stored parameter is [co, (c, k)]: so3conv/modules.py:31-36 of the reference), the data gradients the
// transposes [(k, c), co] (inter) and [c, (k, co)] (intra), the 1x1 skip conv [co, ci] and [ci, co].  All of them are 3-D
// index permutations of the parameter followed by the hi / lo split, so ONE kernel serves a whole model from a table of items
// (the weights change once per step: ~70 small permute / transpose / split launches per step become one).
//   item (10 x int64): src (fp32), hi, lo (bf16 planes, dense in destination order), n0, n1, n2 (destination extents),
//   s0, s1, s2 (source strides in elements of the three destination indices), first block of the item in the grid.
namespace vgtkb {
constexpr int WP_ELEMS_PER_BLOCK = 2048;
__global__ void __launch_bounds__(256) weight_planes_kernel(int n_items, const int64_t* __restrict__ items) {
    __shared__ int s_item;
    if (threadIdx.x == 0) {
        int lo = 0, hi = n_items - 1;                       // last item whose first block is <= blockIdx.x
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (items[(size_t)mid * 10 + 9] <= (int64_t)blockIdx.x) lo = mid;
            else hi = mid - 1;
        }
        s_item = lo;
    }
    __syncthreads();
    const int64_t* it = items + (size_t)s_item * 10;
    const float* src = reinterpret_cast<const float*>(it[0]);
    uint16_t* hi = reinterpret_cast<uint16_t*>(it[1]);
    uint16_t* lo = reinterpret_cast<uint16_t*>(it[2]);
    const int64_t n1 = it[4], n2 = it[5], s0 = it[6], s1 = it[7], s2 = it[8];
    const int64_t total = it[3] * n1 * n2;
    const int64_t base = ((int64_t)blockIdx.x - it[9]) * WP_ELEMS_PER_BLOCK;
#pragma unroll 4
    for (int j = 0; j < WP_ELEMS_PER_BLOCK / 256; ++j) {
        const int64_t i = base + j * 256 + threadIdx.x;
        if (i >= total) break;
        const int64_t i2 = i % n2, r = i / n2, i1 = r % n1, i0 = r / n1;
        const float v = __ldg(src + i0 * s0 + i1 * s1 + i2 * s2);
        const uint32_t h = pack_bf16x2(v, 0.f) & 0xFFFFu;
        hi[i] = (uint16_t)h;
        lo[i] = (uint16_t)(pack_bf16x2(v - __uint_as_float(h << 16), 0.f) & 0xFFFFu);
    }
}
}  // namespace vgtkb

extern "C" int vgtkb_weight_planes(int n_items, const int64_t* items, int total_blocks, void* stream) {
    using namespace vgtkb;
    VGTKB_REQUIRE(n_items >= 0 && total_blocks >= 0, "weight_planes: bad size");
    if (n_items == 0 || total_blocks == 0) return VGTKB_OK;
    VGTKB_REQUIRE(items != nullptr, "weight_planes: item table is NULL");
    weight_planes_kernel<<<total_blocks, 256, 0, (cudaStream_t)stream>>>(n_items, items);
    return check_launch("weight_planes");
}
