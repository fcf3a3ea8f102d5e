/*
 * oracle_ops.c -- CPU restatement of the reference's index/geometry CUDA kernels.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the checker for the sm_100a kernels in
 * equi_articulated_pose_b200/csrc; it is imported by tests/, by
 * __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs,
 * and by nothing else.  The product path never falls back to it.
 *
 * Each function follows one reference kernel (paths relative to /root/reference):
 *   oracle_ball_query   vgtk/vgtk/cuda/grouping_cuda_kernel.cu:67-113, grouping_cuda.cpp:71-86
 *   oracle_fps          vgtk/vgtk/cuda/grouping_cuda_kernel.cu:29-33,339-466, grouping_cuda.cpp:160-174
 *   oracle_gather_fwd   vgtk/vgtk/cuda/gathering_cuda_kernel.cu:43-68
 *   oracle_gather_bwd   vgtk/vgtk/cuda/gathering_cuda_kernel.cu:73-98
 *   oracle_chamfer_fwd  extensions/chamfer_dist/chamfer.cu:15-145
 *   oracle_chamfer_bwd  extensions/chamfer_dist/chamfer.cu:173-229
 *
 * Floating-point contraction is part of the semantics: nvcc
 * (default -fmad=true) turns every 3-term sum of squares of the reference into
 * FMUL + FFMA + FFMA.  We restate that with fmaf() so that the CPU result is
 * bit-identical to the recompiled reference on sm_100a; compile this file with
 * -ffp-contract=off so that gcc adds no contraction of its own.
 *
 *   sq3(a,b,c) = fma(c, c, fma(b, b, a*a))   (the order nvcc 12.9 emits, verified in SASS;
 *                                             see oracle/README.md)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline float sq3(float a, float b, float c) {
    return fmaf(c, c, fmaf(b, b, a * a));
}

/* grouping_cuda_kernel.cu:29-33 : largest power of two <= work_size, clamped to [1,1024] */
int oracle_opt_n_threads(int work_size) {
    int pow_2 = (int)(log((double)work_size) / log(2.0));
    int t = 1 << pow_2;
    if (t > 1024) t = 1024;
    if (t < 1) t = 1;
    return t;
}

/* new_xyz [b,3,m], xyz [b,3,n] -> idx [b,m,nsample] (zero-initialised by the caller in the
 * reference: grouping_cuda.cpp:80-82; we zero it here). */
void oracle_ball_query(int b, int n, int m, float radius, int nsample,
                       const float* new_xyz, const float* xyz, int32_t* idx) {
    memset(idx, 0, sizeof(int32_t) * (size_t)b * m * nsample);
    const float radius2 = radius * radius;
    for (int bi = 0; bi < b; ++bi) {
        const float* q = new_xyz + (size_t)bi * 3 * m;
        const float* s = xyz + (size_t)bi * 3 * n;
        int32_t* out = idx + (size_t)bi * m * nsample;
        for (int j = 0; j < m; ++j) {
            const float qx = q[j], qy = q[m + j], qz = q[2 * m + j];
            int cnt = 0;
            for (int k = 0; k < n && cnt < nsample; ++k) {
                const float d2 = sq3(qx - s[k], qy - s[n + k], qz - s[2 * n + k]);
                if (d2 < radius2) out[j * nsample + cnt++] = k;
            }
            if (cnt < nsample - 1) /* cyclic repetition; cnt == nsample-1 leaves the last slot 0 */
                for (int k = 0; k + cnt < nsample; ++k)
                    out[j * nsample + k + cnt] = out[j * nsample + k];
        }
    }
}

/* xyz [b,3,n] -> idxs [b,m].  Emulates the block of `bs` threads literally: per-thread
 * strided scan with first-strict-max, then the shared-memory tree whose ties keep the
 * left operand (grouping_cuda_kernel.cu:339-346). */
void oracle_fps(int b, int n, int m, const float* xyz, int32_t* idxs) {
    if (m <= 0) return;
    const int bs = oracle_opt_n_threads(n);
    float* temp = (float*)malloc(sizeof(float) * (size_t)n);
    float* dists = (float*)malloc(sizeof(float) * (size_t)bs);
    int* dists_i = (int*)malloc(sizeof(int) * (size_t)bs);
    for (int bi = 0; bi < b; ++bi) {
        const float* d = xyz + (size_t)bi * 3 * n;
        int32_t* out = idxs + (size_t)bi * m;
        for (int k = 0; k < n; ++k) temp[k] = 1e10f;
        int old = 0;
        out[0] = 0;
        for (int j = 1; j < m; ++j) {
            const float x1 = d[old], y1 = d[n + old], z1 = d[2 * n + old];
            for (int tid = 0; tid < bs; ++tid) {
                int besti = 0;
                float best = -1.f;
                for (int k = tid; k < n; k += bs) {
                    const float x2 = d[k], y2 = d[n + k], z2 = d[2 * n + k];
                    const float mag = sq3(x2, y2, z2);
                    if ((double)mag <= 1e-3) continue; /* double literal in the reference */
                    const float dd = sq3(x2 - x1, y2 - y1, z2 - z1);
                    const float d2 = fminf(dd, temp[k]);
                    temp[k] = d2;
                    if (d2 > best) { besti = k; best = d2; }
                }
                dists[tid] = best;
                dists_i[tid] = besti;
            }
            for (int s = bs / 2; s >= 1; s >>= 1)
                for (int tid = 0; tid < s; ++tid) {
                    const float v1 = dists[tid], v2 = dists[tid + s];
                    const int i1 = dists_i[tid], i2 = dists_i[tid + s];
                    dists[tid] = v1 > v2 ? v1 : v2;
                    dists_i[tid] = v2 > v1 ? i2 : i1;
                }
            old = dists_i[0];
            out[j] = old;
        }
    }
    free(temp); free(dists); free(dists_i);
}

/* plain farthest point sampling (the published algorithm of torch_cluster.fps with random_start=False, torch-cluster
 * 1.5.9, env.yaml:13 -- not vendored under /root/reference, tie-break unpinned): start at point 0, keep the running
 * minimum squared distance to the chosen set, take the first maximum.  Distances with the same FMUL,FFMA,FFMA
 * contraction as the other kernels. */
void oracle_fps_plain(int b, int n, int m, const float* xyz, int32_t* idxs) {
    if (m <= 0) return;
    float* temp = (float*)malloc(sizeof(float) * (size_t)n);
    for (int bi = 0; bi < b; ++bi) {
        const float* d = xyz + (size_t)bi * 3 * n;
        int32_t* out = idxs + (size_t)bi * m;
        for (int k = 0; k < n; ++k) temp[k] = 1e10f;
        int old = 0;
        out[0] = 0;
        for (int j = 1; j < m; ++j) {
            const float x1 = d[old], y1 = d[n + old], z1 = d[2 * n + old];
            int besti = 0;
            float best = -1.f;
            for (int k = 0; k < n; ++k) {
                const float dd = sq3(d[k] - x1, d[n + k] - y1, d[2 * n + k] - z1);
                const float d2 = fminf(dd, temp[k]);
                temp[k] = d2;
                if (d2 > best) { besti = k; best = d2; }
            }
            old = besti;
            out[j] = old;
        }
    }
    free(temp);
}

/* points [b,c,n], idx [b,m] -> out [b,c,m] */
void oracle_gather_fwd(int b, int c, int n, int m, const float* points, const int32_t* idx, float* out) {
    for (int bi = 0; bi < b; ++bi)
        for (int ci = 0; ci < c; ++ci)
            for (int j = 0; j < m; ++j)
                out[((size_t)bi * c + ci) * m + j] = points[((size_t)bi * c + ci) * n + idx[(size_t)bi * m + j]];
}

/* grad_out [b,c,m], idx [b,m] -> grad_points [b,c,n] (scatter-add; sequential order here) */
void oracle_gather_bwd(int b, int c, int n, int m, const float* grad_out, const int32_t* idx, float* grad_points) {
    memset(grad_points, 0, sizeof(float) * (size_t)b * c * n);
    for (int bi = 0; bi < b; ++bi)
        for (int ci = 0; ci < c; ++ci)
            for (int j = 0; j < m; ++j)
                grad_points[((size_t)bi * c + ci) * n + idx[(size_t)bi * m + j]] += grad_out[((size_t)bi * c + ci) * m + j];
}

/* one direction of the chamfer kernel: for every point of xyz1 [b,n,3] the nearest point of
 * xyz2 [b,m,3]; strict '<' everywhere => the lowest index wins ties, inside a 512 tile and
 * across tiles (chamfer.cu:139-142). */
static void chamfer_dir(int b, int n, const float* xyz1, int m, const float* xyz2, float* dist, int32_t* index) {
    for (int i = 0; i < b; ++i)
        for (int j = 0; j < n; ++j) {
            const float x1 = xyz1[((size_t)i * n + j) * 3 + 0];
            const float y1 = xyz1[((size_t)i * n + j) * 3 + 1];
            const float z1 = xyz1[((size_t)i * n + j) * 3 + 2];
            float best = 0.f;
            int besti = 0;
            for (int k = 0; k < m; ++k) {
                const float* p = xyz2 + ((size_t)i * m + k) * 3;
                /* recompiled reference (nvcc 12.9, sm_100a), every site of chamfer_dist_kernel:
                 * FMUL y*y, FFMA x, FFMA z  ->  fma(z,z,fma(x,x,y*y)) */
                const float dd = sq3(p[1] - y1, p[0] - x1, p[2] - z1);
                if (k == 0 || dd < best) { best = dd; besti = k; }
            }
            dist[(size_t)i * n + j] = m > 0 ? best : 0.f;
            index[(size_t)i * n + j] = besti;
        }
}

void oracle_chamfer_fwd(int b, int n, const float* xyz1, int m, const float* xyz2,
                        float* dist1, float* dist2, int32_t* idx1, int32_t* idx2) {
    chamfer_dir(b, n, xyz1, m, xyz2, dist1, idx1);
    chamfer_dir(b, m, xyz2, n, xyz1, dist2, idx2);
}

static void chamfer_grad_dir(int b, int n, const float* xyz1, int m, const float* xyz2,
                             const float* grad_dist1, const int32_t* idx1, float* g1, float* g2) {
    for (int i = 0; i < b; ++i)
        for (int j = 0; j < n; ++j) {
            const float* p1 = xyz1 + ((size_t)i * n + j) * 3;
            const int j2 = idx1[(size_t)i * n + j];
            const float* p2 = xyz2 + ((size_t)i * m + j2) * 3;
            const float g = grad_dist1[(size_t)i * n + j] * 2;
            for (int d = 0; d < 3; ++d) {
                const float v = g * (p1[d] - p2[d]);
                g1[((size_t)i * n + j) * 3 + d] += v;
                g2[((size_t)i * m + j2) * 3 + d] += -v;
            }
        }
}

void oracle_chamfer_bwd(int b, int n, const float* xyz1, int m, const float* xyz2,
                        const int32_t* idx1, const int32_t* idx2,
                        const float* grad_dist1, const float* grad_dist2,
                        float* grad_xyz1, float* grad_xyz2) {
    memset(grad_xyz1, 0, sizeof(float) * (size_t)b * n * 3);
    memset(grad_xyz2, 0, sizeof(float) * (size_t)b * m * 3);
    chamfer_grad_dir(b, n, xyz1, m, xyz2, grad_dist1, idx1, grad_xyz1, grad_xyz2);
    chamfer_grad_dir(b, m, xyz2, n, xyz1, grad_dist2, idx2, grad_xyz2, grad_xyz1);
}

/* ---- PointNet++ neighbourhoods (SPConvNets/models/PointNet2.py:85-87, 114-123) --------------------------------------
 * The reference evaluates these in torch: ppdist = sqrt(sum((centre - pos)^2, -1)) followed by topk(k, largest=False), and
 * dist = norm(p2 - p1) followed by topk(3).  Restated here with the evaluation order torch uses (verified on the host, see
 * tests/test_oracle_golden.py): sum over three elements = (dx*dx + dy*dy) + dz*dz with every product rounded (this file is
 * compiled with -ffp-contract=off); torch.norm = sqrt(fma(dz,dz,fma(dy,dy,dx*dx))); square roots correctly rounded (what
 * CUDA's sqrtf and numpy give; torch.sqrt on AVX-512 hosts is up to one ulp off).  Ties: smaller index first (torch.topk
 * leaves the order of equal keys unspecified). */
#include <math.h>

void oracle_knn(int b, int n, int s, int k, const float* pos, const float* centers, int32_t* idx, float* dist) {
    float* d2 = (float*)malloc(sizeof(float) * (size_t)n);
    int32_t* best = (int32_t*)malloc(sizeof(int32_t) * (size_t)k);
    for (int bi = 0; bi < b; ++bi)
        for (int si = 0; si < s; ++si) {
            const float* c = centers + ((size_t)bi * s + si) * 3;
            const float* p = pos + (size_t)bi * n * 3;
            for (int i = 0; i < n; ++i) {
                const float dx = c[0] - p[3 * i], dy = c[1] - p[3 * i + 1], dz = c[2] - p[3 * i + 2];
                const float xx = dx * dx, yy = dy * dy, zz = dz * dz;
                const float t = xx + yy;
                d2[i] = t + zz;
            }
            int cnt = 0;                                   /* insertion into the sorted list of the k best so far */
            for (int i = 0; i < n; ++i) {
                if (cnt == k && !(d2[i] < d2[best[k - 1]])) continue;      /* equal keys: the earlier index stays */
                int j = cnt < k ? cnt : k - 1;
                while (j > 0 && d2[i] < d2[best[j - 1]]) {
                    best[j] = best[j - 1];
                    --j;
                }
                best[j] = i;
                if (cnt < k) ++cnt;
            }
            for (int j = 0; j < k; ++j) {
                idx[((size_t)bi * s + si) * k + j] = best[j];
                dist[((size_t)bi * s + si) * k + j] = sqrtf(d2[best[j]]);
            }
        }
    free(d2);
    free(best);
}

void oracle_three_nn(int b, int n1, int n2, const float* p1, const float* p2, int32_t* idx, float* w) {
    const int kk = n1 < 3 ? n1 : 3;
    for (int bi = 0; bi < b; ++bi)
        for (int q = 0; q < n2; ++q) {
            const float* qq = p2 + ((size_t)bi * n2 + q) * 3;
            float d[3] = {INFINITY, INFINITY, INFINITY};
            int id[3] = {0, 0, 0};
            for (int i = 0; i < n1; ++i) {
                const float* pp = p1 + ((size_t)bi * n1 + i) * 3;
                const float dx = qq[0] - pp[0], dy = qq[1] - pp[1], dz = qq[2] - pp[2];
                const float v = sqrtf(fmaf(dz, dz, fmaf(dy, dy, dx * dx)));
                if (v < d[2]) {
                    int j = 2;
                    while (j > 0 && v < d[j - 1]) {
                        d[j] = d[j - 1];
                        id[j] = id[j - 1];
                        --j;
                    }
                    d[j] = v;
                    id[j] = i;
                }
            }
            float r[3] = {0.f, 0.f, 0.f};
            for (int j = 0; j < kk; ++j) r[j] = 1.0f / (d[j] + 1e-8f);
            const float t = r[0] + r[1];
            const float norm = t + r[2];
            for (int j = 0; j < 3; ++j) {
                idx[((size_t)bi * n2 + q) * 3 + j] = j < kk ? id[j] : 0;
                w[((size_t)bi * n2 + q) * 3 + j] = r[j] / norm;
            }
        }
}
