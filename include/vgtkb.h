/*
 * vgtkb.h -- C ABI of libvgtkb200.so: the B200 (sm_100a) implementation of the vgtk
 * SE(3)-equivariant point-convolution hot path of Meowuu7/equi-articulated-pose.
 *
 * Drop-in boundary.  Every entry point replaces one function of the reference's pybind
 * extension modules (or one torch expression of vgtk.{spconv,so3conv}.functional that the
 * reference evaluates in eager PyTorch); the reference interface each one stands for is cited
 * as file:line relative to the reference tree.  The binding a reference maintainer would add
 * is in INTEGRATION.md.
 *
 * Conventions
 *   - plain pointers and sizes; all pointers are DEVICE pointers unless the name ends in
 *     `_host`; no torch types.  The caller owns and allocates every buffer (the reference's
 *     pybind wrappers allocated outputs with torch::zeros -- the Python shim does that now).
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream, which is what
 *     the reference used: grouping_cuda_kernel.cu:478 etc.).
 *   - return value: 0 on success, negative VGTKB_E* on failure; vgtkb_last_error() returns a
 *     thread-local message.  The reference only printf'd launch errors
 *     (grouping_cuda_kernel.cu:485-487); here they are reported.
 *   - index tensors are int32 like the reference's (grouping_cuda.cpp:80-82).
 *   - feature tensors on the fused path are CHANNELS-LAST: X[B][N][A][C] (row = (b,n,a),
 *     C contiguous).  The literal 1:1 ops keep the reference layouts ([B,3,N], [B,C,N]).
 */
#ifndef VGTKB_H_
#define VGTKB_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VGTKB_OK 0
#define VGTKB_EINVAL (-1)   /* bad argument (size, alignment, unsupported shape) */
#define VGTKB_ECUDA (-2)    /* CUDA runtime / launch error */
#define VGTKB_EUNSUP (-3)   /* feature not available on this device (needs sm_100) */

const char* vgtkb_last_error(void);
int vgtkb_version(void);                 /* ABI version, bumped on signature changes */
int vgtkb_device_check(void);            /* VGTKB_OK iff current device is compute capability 10.x */

/* ---------------------------------------------------------------- literal 1:1 ops ---------- */

/* vgtk.cuda.grouping.ball_query  (vgtk/vgtk/cuda/grouping_cuda.cpp:71-86, kernel
 * grouping_cuda_kernel.cu:67-113).  new_xyz [b,3,m], xyz [b,3,n] fp32 -> idx [b,m,nsample].
 * First `nsample` support points with d2 < radius^2 in ascending index order; fewer than
 * nsample-1 hits are repeated cyclically; exactly nsample-1 hits leave the last slot 0;
 * zero hits give an all-zero row.  idx need not be zeroed by the caller. */
int vgtkb_ball_query(int b, int n, int m, float radius, int nsample,
                     const float* new_xyz, const float* xyz, int32_t* idx, void* stream);

/* vgtk.cuda.grouping.furthest_point_sampling  (grouping_cuda.cpp:160-174, kernel
 * grouping_cuda_kernel.cu:351-466).  xyz [b,3,n] fp32 -> idx [b,m].  Bit-exact including the
 * block-size dependent tie-break of the reference's shared-memory tree and the |p|^2 <= 1e-3
 * skip rule. */
int vgtkb_furthest_point_sampling(int b, int n, int m, const float* xyz, int32_t* idx, void* stream);
/* Plain farthest point sampling, same layouts: start at point 0, every point eligible (no |p|^2 <= 1e-3 exclusion),
 * lowest index on ties.  Backs the torch_cluster.fps(pos, batch, ratio, random_start=False) shim the reference's
 * wrappers call (SPConvNets/models/model_util.py:183-200, ...38_multi_stage.py:1739); torch_cluster==1.5.9 is not
 * under /root/reference and no reference test pins its tie-break: parity unpinned. */
int vgtkb_fps_plain(int b, int n, int m, const float* xyz, int32_t* idx, void* stream);

/* vgtk.cuda.gathering.gather_points_forward / backward (gathering_cuda.cpp:29-58, kernels
 * gathering_cuda_kernel.cu:43-98).  points [b,c,n], idx [b,m] -> out [b,c,m];
 * backward: grad_out [b,c,m] scatter-added into grad_points [b,c,n] (zeroed here). */
int vgtkb_gather_points_forward(int b, int c, int n, int m, const float* points, const int32_t* idx,
                                float* out, void* stream);
int vgtkb_gather_points_backward(int b, int c, int n, int m, const float* grad_out, const int32_t* idx,
                                 float* grad_points, void* stream);

/* chamfer.forward / chamfer.backward (extensions/chamfer_dist/chamfer_cuda.cpp:22-39, kernels
 * chamfer.cu:15-145,173-229).  xyz1 [b,n,3], xyz2 [b,m,3] fp32. */
int vgtkb_chamfer_forward(int b, int n, const float* xyz1, int m, const float* xyz2,
                          float* dist1, float* dist2, int32_t* idx1, int32_t* idx2, void* stream);
int vgtkb_chamfer_backward(int b, int n, const float* xyz1, int m, const float* xyz2,
                           const int32_t* idx1, const int32_t* idx2,
                           const float* grad_dist1, const float* grad_dist2,
                           float* grad_xyz1, float* grad_xyz2, void* stream);

/* Anchor chamfer: the reconstruction loss of model 38 (SPConvNets/models/unsup_seg_so3_pose_conv_pn_38_multi_stage.py:
 * 429-436) without its [B,A,M,3] transformed tensor and its A-fold replicated input cloud.
 *   Y[b,a,i] = rot[b,a] canon[b,i] + trans[b,a];  dist1/idx1 [b,a,m]: Y -> ori[b];  dist2/idx2 [b,a,n]: ori[b] -> Y[b,a]
 * canon [b,m,3], rot [b,a,3,3] row-major, trans [b,a,3], ori [b,n,3].  Backward: grad_y [b,a,m,3] (gradient w.r.t. the
 * transformed points; the caller folds it into rot / trans / canon) and grad_ori [b,n,3] (may be NULL). */
int vgtkb_anchor_chamfer_forward(int b, int a, int m, const float* canon, const float* rot, const float* trans, int n,
                                 const float* ori, float* dist1, float* dist2, int32_t* idx1, int32_t* idx2, void* stream);
int vgtkb_anchor_chamfer_backward(int b, int a, int m, const float* canon, const float* rot, const float* trans, int n,
                                  const float* ori, const int32_t* idx1, const int32_t* idx2, const float* grad_dist1,
                                  const float* grad_dist2, float* grad_y, float* grad_ori, void* stream);

/* ---------------------------------------------------------------- fused SO(3) conv path ---- */

/* Kernel-point correlation, materialised (API parity only; the fused path never stores it):
 * vgtk/vgtk/so3conv/functional.py:2508-2549.  xyz [b,3,n], sample_xyz [b,3,p], idx [b,p,nn],
 * rot_kernels [a,k,3] (= R_a kappa_k) -> w [b,p,a,k,nn]. */
int vgtkb_inter_weights(int b, int n, int p, int nn, int a, int k, const float* xyz, const float* sample_xyz,
                        const int32_t* idx, const float* rot_kernels, float sigma, float* w, void* stream);

/* Inter-anchor grouping = vgtk.cuda.zpconv.inter_zpconv_forward slot (zpconv_cuda.cpp:113-118)
 * with the semantics of the live torch path: inter_so3conv_grouping_anchor + add_shadow_feature +
 * inter_zpconv_grouping_naive (so3conv/functional.py:2508-2549, spconv/functional.py:375-406).
 *   feats X [b,n,a,ci] (channels-last) -> G [b,p,a,k,ci]   (row (b,p,a), column k*ci + c)
 *   G[b,p,a,k,c] = sum_j relu(1 - |xyz[b,:,idx[b,p,j]] - sample_xyz[b,:,p] - R_a kappa_k|^2 / sigma)
 *                        * X[b, idx[b,p,j], a, c]
 * The weights are recomputed on the fly, never stored.
 * mode: arithmetic of the small per-(point, anchor) product, matching the contraction that consumes G:
 *   0 = fp32 FFMA (exact fp32 accumulation)   3 = bf16x3 on the tensor cores (warp-level MMA, ~1e-5 of max|G|);
 *   any other value selects the FFMA kernel. */
int vgtkb_inter_group_forward(int b, int n, int p, int nn, int a, int k, int ci,
                              const float* xyz, const float* sample_xyz, const int32_t* idx,
                              const float* rot_kernels, float sigma,
                              const float* feats, float* grouped, int mode, void* stream);
/* inter_zpconv_backward slot: grad_feats [b,n,a,ci] += scatter of grad_grouped (grad_feats is
 * accumulated into; the caller zeroes it when needed). */
int vgtkb_inter_group_backward(int b, int n, int p, int nn, int a, int k, int ci,
                               const float* xyz, const float* sample_xyz, const int32_t* idx,
                               const float* rot_kernels, float sigma,
                               const float* grad_grouped, float* grad_feats, int mode, void* stream);

/* Intra-anchor grouping = intra_zpconv_forward slot with the semantics of
 * intra_so3conv_grouping (so3conv/functional.py:2553-2567):
 *   Y [rows, a, c] -> G [rows, a, kk, c],  G[r,a,k,c] = Y[r, intra_idx[a,k], c];  rows = b*p.
 * backward: grad_y[r,a',c] = sum_{(a,k): intra_idx[a,k]=a'} grad_g[r,a,k,c] (deterministic). */
int vgtkb_intra_group_forward(int64_t rows, int a, int kk, int c, const int32_t* intra_idx,
                              const float* y, float* grouped, void* stream);
int vgtkb_intra_group_backward(int64_t rows, int a, int kk, int c, const int32_t* intra_idx,
                               const float* grad_grouped, float* grad_y, void* stream);

/* Pose-aware inter grouping with arbitrary per-point rotations: the no-stride branch of
 * inter_so3poseconv_grouping_strided (so3conv/functional.py:1061-1261) behind InterSO3PoseConv (so3conv/modules.py:222-322).
 *   pose_neighbourhood: xyz [b,3,n], pose [b,n,4,4], idx [b,n,nn], anchors [a,3,3] ->
 *       rel_xyz [b,n,nn,3] = R_p R_j^T (x_j - x_p);  perm [b,n,nn,a] (uint8, may be NULL) = argmax_a' tr((R_rel^T R_a) R_a'^T)
 *   inter_pose_group_forward: G[b,p,a,k,c] = sum_n relu(1 - |rel_xyz[b,p,n] - R_a kappa_k|^2 / sigma) X[b, idx[b,p,n], perm[b,p,n,a], c]
 *       (perm NULL = identity: permute_modes 0);  backward accumulates into grad_feats (caller zeroes it).
 * X / G channels-last as in inter_group_forward.  The identity-pose case runs inter_group_forward instead. */
int vgtkb_pose_neighbourhood(int b, int n, int nn, int a, const float* xyz, const float* pose, const int32_t* idx,
                             const float* anchors, float* rel_xyz, uint8_t* perm, void* stream);
int vgtkb_inter_pose_group_forward(int b, int n, int nn, int a, int k, int ci, const int32_t* idx, const float* rel_xyz,
                                   const uint8_t* perm, const float* rot_kernels, float sigma, const float* feats,
                                   float* grouped, void* stream);
int vgtkb_inter_pose_group_backward(int b, int n, int nn, int a, int k, int ci, const int32_t* idx, const float* rel_xyz,
                                    const uint8_t* perm, const float* rot_kernels, float sigma, const float* grad_grouped,
                                    float* grad_feats, void* stream);
/* The strided branch of the same function (so3conv/functional.py:896-1060; sampling and ball query as in
 * spconv/functional.py:468-500): p = ceil(n / stride) centres sample_xyz [b,3,p] = xyz[sample_idx], centre rotations
 * pose[sample_idx] (sample_idx [b,p] int32; NULL = the identity map, which needs p == n), idx [b,p,nn] into the n points.
 * rel_xyz [b,p,nn,3], perm [b,p,nn,a], G [b,p,a,k,c]; X and grad_feats stay [b,n,a,c]. */
int vgtkb_pose_neighbourhood_strided(int b, int n, int p, int nn, int a, const float* xyz, const float* pose,
                                     const float* sample_xyz, const int32_t* sample_idx, const int32_t* idx,
                                     const float* anchors, float* rel_xyz, uint8_t* perm, void* stream);
int vgtkb_inter_pose_group_forward_strided(int b, int n, int p, int nn, int a, int k, int ci, const int32_t* idx,
                                           const float* rel_xyz, const uint8_t* perm, const float* rot_kernels, float sigma,
                                           const float* feats, float* grouped, void* stream);
int vgtkb_inter_pose_group_backward_strided(int b, int n, int p, int nn, int a, int k, int ci, const int32_t* idx,
                                            const float* rel_xyz, const uint8_t* perm, const float* rot_kernels, float sigma,
                                            const float* grad_grouped, float* grad_feats, void* stream);

/* vgtk.cuda.zpconv.{inter,intra}_zpconv_{forward,backward} (zpconv_cuda.cpp:113-118, kernels
 * zpconv_cuda_kernel.cu:33-195), reference layouts, explicit index/weight tensors:
 *   inter: idx, w [b,p,a,k,ann]; feats [b,c,nq,a]  -> out [b,c,k,p,a];  backward -> grad_feats [b,c,nq,a] (zeroed here)
 *   intra: idx [aout,ann], w [aout,k,ann]; feats [b,c,p,ain] -> out [b,c,k,p,aout]; backward -> grad_feats [b,c,p,ain]
 * Live grouping of the legacy S^2 ZPConv modules (vgtk/vgtk/spconv/modules.py:61-98, config 1b). */
int vgtkb_inter_zpconv_forward(int b, int c, int nq, int p, int a, int k, int ann, const int32_t* idx, const float* w,
                               const float* feats, float* out, void* stream);
int vgtkb_inter_zpconv_backward(int b, int c, int nq, int p, int a, int k, int ann, const int32_t* idx, const float* w,
                                const float* grad_out, float* grad_feats, void* stream);
int vgtkb_intra_zpconv_forward(int b, int c, int p, int ain, int aout, int k, int ann, const int32_t* idx, const float* w,
                               const float* feats, float* out, void* stream);
int vgtkb_intra_zpconv_backward(int b, int c, int p, int ain, int aout, int k, int ann, const int32_t* idx, const float* w,
                                const float* grad_out, float* grad_feats, void* stream);

/* Point-row gather used by the skip connection (zptk.functional.batched_index_select(feats, 2,
 * sample_idx), SPConvNets/utils/base_so3conv.py:212-213) on channels-last rows of `width`
 * floats: out[b,j,:] = x[b, idx[b,j], :]; backward accumulates (atomics) into grad_x. */
int vgtkb_row_gather_forward(int b, int n, int m, int width, const float* x, const int32_t* idx,
                             float* out, void* stream);
int vgtkb_row_gather_backward(int b, int n, int m, int width, const float* grad_out, const int32_t* idx,
                              float* grad_x, void* stream);

/* Anchor/kernel contraction = BasicSO3Conv.forward (so3conv/modules.py:48-55) and the 1x1 skip
 * conv, as row-major GEMMs.
 *   vgtkb_gemm_nt: C[M,N] = A[M,K] * B[N,K]^T (+ bias[N])
 *   vgtkb_gemm_tn: C[M,N] (+)= A[R,M]^T * B[R,N]   (weight gradients; reduction over rows R)
 * `mode`: 0 = fp32 FFMA (CUDA cores), 1 = tcgen05 3xTF32 (fp32-equivalent, sm_100a tensor cores),
 *         2 = tcgen05 single-pass TF32 (fast, ~1e-3 relative),
 *         3 = tcgen05 bf16x3 (operands split into two bf16, 16 significand bits, ~5e-6 relative; 2x the TF32 rate).
 * In modes 1-3 the tensor core only accumulates 64 reduction elements at a time in TMEM; the running sum
 * is kept in fp32 registers with round-to-nearest adds (the TMEM accumulator rounds toward zero).
 * `workspace`: device scratch, 16-byte aligned.  gemm_nt, modes 1/3: 2*N*K floats (hi/lo split of B); when
 * NULL the library allocates it stream-ordered (cudaMallocAsync), which is slower.  gemm_tn, mode 3: R*M
 * floats (bf16 hi/lo split of A); when NULL mode 3 runs as mode 1. */
int vgtkb_gemm_nt(int64_t M, int N, int K, const float* A, const float* B, const float* bias,
                  float* C, int mode, float* workspace, void* stream);
int vgtkb_gemm_tn(int M, int N, int64_t R, const float* A, const float* B, float* C,
                  int accumulate, int mode, float* workspace, void* stream);

/* Intra-anchor convolution as ONE gather-GEMM: intra_so3conv_grouping (so3conv/functional.py:2553-2567) fused
 * into BasicSO3Conv.forward (so3conv/modules.py:48-55); the 12x gathered tensor [b,c,12,p,a] is never built.
 * x [points, anchors, c] channels-last, table [anchors, kk] int32 (the 60x12 neighbour table, or its per-column
 * inverse for the data gradient), w [n, kk*c] (column = kk*c + channel):
 *   gather_gemm_nt: out[(pt,a), o]   = sum_{kk,ch} x[pt, table[a,kk], ch] * w[o, kk*c + ch] (+ bias[o])
 *   gather_gemm_tn: out[o, kk*c + ch] (+)= sum_{pt,a} y[(pt,a), o] * x[pt, table[a,kk], ch]      (y [points*anchors, m])
 * The operand tiles are regular boxes of x (3-D TMA loads with the anchor coordinate looked up in `table`).
 * Tensor-core modes only (1, 2, 3); c % 64 == 0.  workspace: nt 2*n*kk*c floats, tn (mode 3) points*anchors*m floats. */
int vgtkb_gather_gemm_nt(int64_t points, int anchors, int kk, int c, int n, const int32_t* table, const float* x,
                         const float* w, const float* bias, float* out, int mode, float* workspace, void* stream);
int vgtkb_gather_gemm_tn(int64_t points, int anchors, int kk, int c, int m, const int32_t* table, const float* x,
                         const float* y, float* out, int accumulate, int mode, float* workspace, void* stream);

/* Normalisation + leaky_relu on channels-last rows X[groups][rows_per_group][c].
 * groups == 1 : BatchNorm2d training statistics (base_so3conv.py:113,125-131)
 * groups == b : InstanceNorm2d(affine=False)            (base_so3conv.py:47-64)
 * stats [groups][2][c] = (mean, invstd); sums are accumulated in fp64 scratch [groups][2][c]. */
int vgtkb_norm_stats(int groups, int64_t rows_per_group, int c, const float* x, float eps,
                     double* scratch, float* stats, float* running_mean, float* running_var,
                     float momentum, void* stream);
/* y = leaky_relu((x-mean)*invstd*gamma+beta, slope) [+ residual]; gamma/beta may be NULL. */
int vgtkb_norm_act_forward(int groups, int64_t rows_per_group, int c, const float* x, const float* stats,
                           const float* gamma, const float* beta, float slope, const float* residual,
                           float* y, void* stream);
/* backward of the above w.r.t. x, gamma, beta.  grad_gamma/grad_beta [c] may be NULL (no affine).
 * scratch: fp64 [groups][2][c]. */
int vgtkb_norm_act_backward(int groups, int64_t rows_per_group, int c, const float* x, const float* stats,
                            const float* gamma, const float* beta, float slope, const float* grad_y,
                            double* scratch, float* grad_x, float* grad_gamma, float* grad_beta, void* stream);

/* The same two operations in two phases each, for SyncBatchNorm across data-parallel ranks (the reference converts its
 * BatchNorm layers with nn.SyncBatchNorm.convert_sync_batchnorm, SPConvNets/trainer_unsup_arti_align.py:430): the caller
 * all-reduces `scratch` (fp64 sums) between the phases and passes the global row count as total_rows.
 *   vgtkb_norm_stats        == vgtkb_norm_sums     + vgtkb_norm_finalize(total_rows = rows)
 *   vgtkb_norm_act_backward == vgtkb_norm_bwd_sums + vgtkb_norm_bwd_apply(total_rows = rows)
 * grad_gamma / grad_beta of bwd_sums are the LOCAL sums (reduced with the other parameter gradients). */
int vgtkb_norm_sums(int groups, int64_t rows_per_group, int c, const float* x, double* scratch, void* stream);
int vgtkb_norm_finalize(int groups, int64_t total_rows, int c, float eps, const double* scratch, float* stats,
                        float* running_mean, float* running_var, float momentum, void* stream);
int vgtkb_norm_bwd_sums(int groups, int64_t rows_per_group, int c, const float* x, const float* stats, const float* gamma,
                        const float* beta, float slope, const float* grad_y, double* scratch, float* grad_gamma,
                        float* grad_beta, void* stream);
int vgtkb_norm_bwd_apply(int groups, int64_t rows_per_group, int64_t total_rows, int c, const float* x, const float* stats,
                         const float* gamma, const float* beta, float slope, const float* grad_y, const double* scratch,
                         float* grad_x, void* stream);

/* SyncBatchNorm exchange over NVLink peer memory instead of an NCCL all-reduce (csrc/peer.cu): every rank owns a
 * mailbox in its HBM that its peers (other processes of the node, cudaIpc-mapped) store their fp64 sums into.
 *   vgtkb_peer_mailbox_bytes: size of one mailbox for `world` ranks (<= 16)
 *   vgtkb_peer_alloc:  cudaMalloc + zero a mailbox on the current device, export its IPC handle (64 bytes)
 *   vgtkb_peer_open / vgtkb_peer_close: map / unmap a peer's mailbox from its handle;  vgtkb_peer_free: own mailbox
 *   vgtkb_peer_allreduce_f64: buf[0..n) <- sum over ranks, n <= 2049; one single-CTA kernel; `mailboxes` = HOST array of
 *       `world` device pointers (entry `rank` = own mailbox); `seq` = 0: the kernel takes the next value of the mailbox's
 *       device-side exchange counter (replayable from a CUDA graph; what the Python host passes), or an explicit 1, 2, 3, ...
 *       identical on all ranks per call
 *   vgtkb_norm_finalize_peer: the exchange of scratch [2c sums | row count] fused with vgtkb_norm_finalize (groups = 1)
 *   vgtkb_peer_status: 0, or r + 1 when rank r did not arrive within the timeout in some exchange (synchronises `stream`)
 * Every rank must issue the same sequence of exchanges.  A peer that does not show up within VGTKB_PEER_TIMEOUT_S seconds
 * (default 600, the order of a process-group timeout) makes the kernel give up: it records the status word, prints a
 * message and returns with invalid sums -- the CUDA context survives (no trap).  VGTKB_PEER_SYNCBN=0 selects NCCL instead. */
#define VGTKB_IPC_HANDLE_BYTES 64
int vgtkb_peer_mailbox_bytes(int world, int64_t* bytes);
int vgtkb_peer_alloc(int64_t bytes, void** dev_ptr, void* ipc_handle);
int vgtkb_peer_open(const void* ipc_handle, void** dev_ptr);
int vgtkb_peer_close(void* dev_ptr);
int vgtkb_peer_free(void* dev_ptr);
int vgtkb_peer_allreduce_f64(int n, double* buf, int rank, int world, void* const* mailboxes, uint64_t seq, void* stream);
int vgtkb_norm_finalize_peer(int c, float eps, double* scratch, float* stats, float* running_mean, float* running_var,
                             float momentum, int rank, int world, void* const* mailboxes, uint64_t seq, void* stream);
int vgtkb_peer_status(const void* own_mailbox, int world, int64_t* status, void* stream);

/* PointnetSO3Conv pooling head (vgtk/vgtk/so3conv/modules.py:376-413).  e [b, n, a, co]: rows of the feature part of
 * the 1x1 conv (vgtkb_gemm_nt with the bias); v [a, co, 3] = sum_i W_x[o,i] anchors[a,j,i]; xc [b, 3, n] centred xyz.
 *   pool_forward:  out [b, co, a] = max_n (e + v . xc), arg [b, a, co] = first arg-max (int32)
 *   embed_xyz:     e += v . xc in place (the module's return_raw=True output)
 *   pool_backward: grad_e [b, n, a, co] = grad_out scattered to the arg-max rows (zero elsewhere) */
int vgtkb_pointnet_pool_forward(int b, int n, int a, int co, const float* e, const float* v, const float* xc, float* out,
                                int32_t* arg, void* stream);
int vgtkb_pointnet_embed_xyz(int b, int n, int a, int co, float* e, const float* v, const float* xc, void* stream);
int vgtkb_pointnet_pool_backward(int b, int n, int a, int co, const float* grad_out, const int32_t* arg, float* grad_e,
                                 void* stream);

/* PointNet++ set abstraction / feature propagation (SPConvNets/models/PointNet2.py:78-129, `PointnetPP`; csrc/pointnet2.cu).
 * Clouds are [b, n, 3] rows here (the layout PointnetPP works in), features [b, n, c] rows.
 *   knn_query (sample_and_group :85-87: sqrt(sum((centre - pos)^2)) + torch.topk(k, largest=False)):
 *       idx int32 [b, s, k] / dist [b, s, k] = the k nearest points of every centre, ascending distance, ties to the
 *       smaller index; squared distance evaluated as (dx*dx + dy*dy) + dz*dz (torch's CPU order), dist = sqrt of it
 *   sa_group_forward (:92-100): out [b, s, k, cpad] rows = [pos[idx] - centre | feat[idx] | zero padding to cpad];
 *       idx == NULL: identity neighbourhood (s = 1, k = n: the global level :151-156), centers == NULL: origin;
 *       feat == NULL with c = 0: coordinates only
 *   sa_group_backward: grad_feat [b, n, c] = scatter-add of grad_out[..., 3:3+c]
 *   sa_maxpool_forward (max_pooling_with_r :102-112): out [groups, c] = max_j (dist[g,j] <= radius ? y[g,j,:] : -1e8),
 *       arg int32 = first arg-max; dist == NULL: plain max
 *   sa_maxpool_backward: grad_y [groups, k, c] = grad_out at the arg-max row, zero elsewhere (written in one pass)
 *   three_nn (interpolate_features :114-123): for every p2 point the min(3, n1) nearest p1 points (torch.norm order:
 *       fma(dz,dz,fma(dy,dy,dx*dx))), weights (1/(d+1e-8)) / sum; unused slots have weight 0
 *   three_interpolate_forward/backward (:126-128): out [b, n2, c] = sum_j feat[b, idx_j, :] * w_j and its scatter-add */
int vgtkb_knn_query(int b, int n, int s, int k, const float* pos, const float* centers, int32_t* idx, float* dist, void* stream);
int vgtkb_sa_group_forward(int b, int n, int s, int k, int c, int cpad, const float* pos, const float* feat,
                           const float* centers, const int32_t* idx, float* out, void* stream);
int vgtkb_sa_group_backward(int b, int n, int s, int k, int c, int cpad, const float* grad_out, const int32_t* idx,
                            float* grad_feat, void* stream);
int vgtkb_sa_maxpool_forward(int64_t groups, int k, int c, const float* y, const float* dist, float radius, float* out,
                             int32_t* arg, void* stream);
int vgtkb_sa_maxpool_backward(int64_t groups, int k, int c, const float* grad_out, const int32_t* arg, float* grad_y,
                              void* stream);
int vgtkb_three_nn(int b, int n1, int n2, const float* p1, const float* p2, int32_t* idx, float* w, void* stream);
int vgtkb_three_interpolate_forward(int b, int n1, int n2, int c, const float* feat, const int32_t* idx, const float* w,
                                    float* out, void* stream);
int vgtkb_three_interpolate_backward(int b, int n1, int n2, int c, const float* grad_out, const int32_t* idx, const float* w,
                                     float* grad_feat, void* stream);

/* Plane operands (csrc/gemm_tc.cu): an activation operand stored once as two bf16 planes (hi = bf16_rn(x),
 * lo = bf16_rn(x - hi)) by its producer and consumed by the contraction without the in-kernel operand conversion that
 * bounds the narrow layers.  vgtkb_inter_conv_* below is built on them.
 *   split_bf16:        hi / lo planes (uint16 bf16 bit patterns, n elements each) of an fp32 array
 *   gemm_nt_presplit:  C [M, N] = (a_hi + a_lo) [M, K] * B [N, K]^T (+ bias), bf16x3 arithmetic of vgtkb_gemm_nt mode 3
 *                      (bit-identical results); K % 8 == 0, K >= 64; workspace: N*K floats (hi / lo planes of B), 16-byte aligned */
int vgtkb_split_bf16(int64_t n, const float* x, void* hi, void* lo, void* stream);
int vgtkb_gemm_nt_presplit(int64_t M, int N, int K, const void* a_hi, const void* a_lo, const float* B, const float* bias,
                           float* C, float* workspace, void* stream);
/*   gemm_tn_presplit:  C [M, N] (+)= A [R, M]^T * (b_hi + b_lo) [R, N] (weight gradient with the wide operand as planes);
 *                      M % 8 == 0, N % 8 == 0, R >= 64; workspace: R*M floats (split of the narrow operand) */
int vgtkb_gemm_tn_presplit(int M, int N, int64_t R, const float* A, const void* b_hi, const void* b_lo, float* C, int accumulate,
                           float* workspace, void* stream);

/*   gemm_tn_planes:    general form: narrow operand A as planes (a_hi / a_lo) or fp32 (split into workspace), wide operand B
 *                      as planes (b_hi / b_lo) or fp32 (B, converted in the kernel)
 *   gather_gemm_*_planes: the intra-conv gather-GEMMs (vgtkb_gather_gemm_nt / _tn) with X [points, a, c] given as planes
 *                      (c % 64 == 0); the weight gradient takes Y as planes too, or as fp32 (split into workspace: points*a*m floats) */
int vgtkb_gemm_tn_planes(int M, int N, int64_t R, const float* A, const void* a_hi, const void* a_lo, const float* B,
                         const void* b_hi, const void* b_lo, float* C, int accumulate, float* workspace, void* stream);
int vgtkb_gather_gemm_nt_planes(int64_t points, int anchors, int kk, int c, int n, const int32_t* table, const void* x_hi,
                                const void* x_lo, const float* w, const float* bias, float* out, float* workspace, void* stream);
int vgtkb_gather_gemm_tn_planes(int64_t points, int anchors, int kk, int c, int m, const int32_t* table, const void* x_hi,
                                const void* x_lo, const float* y, const void* y_hi, const void* y_lo, float* out, int accumulate,
                                float* workspace, void* stream);
/* norm kernels that also write the bf16 planes of their result (y_hi / y_lo resp. gx_hi / gx_lo, same shape as the fp32
 * output; both NULL = plain call): the consumer contraction then needs no operand conversion.  c % 4 == 0, 256 % (c/4) == 0. */
int vgtkb_norm_act_forward_planes(int groups, int64_t rows, int c, const float* x, const float* stats, const float* gamma,
                                  const float* beta, float slope, const float* residual, float* y, void* y_hi, void* y_lo,
                                  void* stream);
int vgtkb_norm_bwd_apply_planes(int groups, int64_t rows, int64_t total_rows, int c, const float* x, const float* stats,
                                const float* gamma, const float* beta, float slope, const float* grad_y, const double* scratch,
                                float* grad_x, void* gx_hi, void* gx_lo, void* stream);
int vgtkb_norm_act_backward_planes(int groups, int64_t rows, int c, const float* x, const float* stats, const float* gamma,
                                   const float* beta, float slope, const float* grad_y, double* scratch, float* grad_x,
                                   float* grad_gamma, float* grad_beta, void* gx_hi, void* gx_lo, void* stream);

/* InterSO3Conv as one call per direction: ball-neighbourhood grouping under the kernel-point correlation + contraction with W
 * (reference: inter_so3conv_grouping, vgtk/vgtk/so3conv/functional.py:144-203 -> inter_zpconv_grouping_naive,
 * vgtk/vgtk/spconv/functional.py:375-406 -> BasicSO3Conv.forward, vgtk/vgtk/so3conv/modules.py:48-55; backward = what autograd
 * derives from those three).  bf16x3 tensor-core arithmetic (mode 3).
 *   forward:  feats X [b,n,a,ci] (channels-last), w_kc [co, k*ci] (column k*ci + c) -> out [b*p*a, co].
 *             g_hi / g_lo [b*p*a, k*ci] bf16 each: the grouped tensor in the contraction's operand format, written once by
 *             the grouping kernel and read by TMA only (forward contraction here, weight gradient in backward); the caller
 *             keeps them for backward.  workspace: co*k*ci floats.
 *   backward: grad_out [b*p*a, co] -> grad_w [co, k*ci] (overwritten; NULL = skip) and grad_feats [b,n,a,ci] (overwritten;
 *             NULL = skip; needs the scratch grad_grouped [b*p*a, k*ci] fp32).  workspace: max(b*p*a*co, 2*k*ci*co) floats.
 *             grad_out_hi / grad_out_lo: optional bf16 planes of grad_out (vgtkb_norm_bwd_apply_planes writes them); NULL =
 *             grad_out is split / converted here.
 *   mode:     3 = bf16x3 (fp32-parity: both planes), 4 = single-pass bf16 (BASELINE config 3 "bf16": operands rounded to bf16
 *             once, one tensor-core pass, ~2e-3 of the output maximum per conv; g_lo / grad_out_lo unused, may be NULL).
 *             backward only: mode | 256 = grad_feats already holds a gradient (e.g. the skip branch's) and the scatter adds to it
 *             (no memset of grad_feats).
 *   supported: 1 iff the shape is taken (k <= 24, nn <= 64, ci % 32 == 0, co % 8 == 0, co <= 1024, b*p*a >= 64, 32-bit offsets);
 *             other shapes run vgtkb_inter_group_* + vgtkb_gemm_*. */
int vgtkb_inter_conv_supported(int b, int n, int p, int nn, int a, int k, int ci, int co);
int vgtkb_inter_conv_forward(int b, int n, int p, int nn, int a, int k, int ci, int co, const float* xyz,
                             const float* sample_xyz, const int32_t* idx, const float* rot_kernels, float sigma,
                             const float* feats, const float* w_kc, void* g_hi, void* g_lo, float* workspace, float* out,
                             int mode, void* stream);
int vgtkb_inter_conv_backward(int b, int n, int p, int nn, int a, int k, int ci, int co, const float* xyz,
                              const float* sample_xyz, const int32_t* idx, const float* rot_kernels, float sigma,
                              const float* w_kc, const void* g_hi, const void* g_lo, const float* grad_out,
                              const void* grad_out_hi, const void* grad_out_lo, float* grad_grouped, float* grad_feats,
                              float* grad_w, float* workspace, int mode, void* stream);

/* Weight operand planes for a whole model in ONE launch per step (csrc/gemm_tc.cu).  Every contraction of a block takes
 * its weight operand as two bf16 planes in its own index order -- forward W [co, (k, c)] (the reference stores the
 * parameter as [co, (c, k)], vgtk/vgtk/so3conv/modules.py:31-36), data gradients W^T [(k, c), co] (inter) and
 * [c, (k, co)] (intra), the 1x1 skip conv [co, ci] / [ci, co] -- each a 3-D index permutation of the parameter followed
 * by the hi / lo split.  `items` [n_items][10] int64 (device): src (const float*), hi, lo (bf16 planes, dense in
 * destination order), n0, n1, n2 (destination extents), s0, s1, s2 (source strides, in elements, of the three
 * destination indices), first grid block of the item (2048 elements per block; `total_blocks` = blocks of all items).
 * CONSUMERS: an entry point whose fp32 weight argument is NULL reads the prepared planes (hi | lo, contiguous) from its
 * `workspace` argument instead of splitting the weights itself: vgtkb_gemm_nt (B, modes 3 / 4, K % 8 == 0),
 * vgtkb_gemm_nt_presplit (B), vgtkb_gather_gemm_nt_planes (w), vgtkb_inter_conv_forward (w_kc; planes of W [co, k*ci]),
 * vgtkb_inter_conv_backward (w_kc; planes of W^T [k*ci, co]; needs grad_out_hi / _lo when grad_w is requested). */
int vgtkb_weight_planes(int n_items, const int64_t* items, int total_blocks, void* stream);

/* column sums of a row-major [rows, c] matrix (bias gradient of the skip conv) */
int vgtkb_col_sum(int64_t rows, int c, const float* x, double* scratch, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VGTKB_H_ */
