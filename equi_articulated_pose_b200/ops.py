"""Tensor-level operators over the C ABI (include/vgtkb.h) + their autograd wrappers.

Layouts: the literal 1:1 ops keep the reference layouts (xyz [B,3,N], points [B,C,N]); the
fused SO(3) path works on CHANNELS-LAST feature rows X[B,N,A,C].
Every function allocates its outputs with torch and launches on the current stream.
"""
import torch

from . import lib as _lib

call, ptr = _lib.call, _lib.ptr

# GEMM arithmetic: 0 = fp32 FFMA, 1 = tcgen05 3xTF32 (~1e-6 per GEMM), 2 = tcgen05 1xTF32 (~1e-3),
# 3 = tcgen05 bf16x3 (default: 16 significand bits per operand, ~5e-6 per GEMM, 3.5e-5 on the backbone output at
#     config-2 size against the 1e-4 bar; 2x the TF32 tensor rate)
# 4 = tcgen05 single-pass bf16 ("fast" mode of BASELINE config 3: operands rounded to bf16 once, fp32 accumulation, one
#     tensor-core pass instead of three; ~2e-3 of the tensor maximum per contraction, stated tolerance 2e-2 on the backbone
#     output; the reference has no reduced-precision mode, so this one is new and opt-in)
_GEMM_MODE = 3
LEAKY_SLOPE = 0.01  # F.leaky_relu default used by the reference blocks


def set_gemm_mode(mode):
    global _GEMM_MODE
    assert mode in (0, 1, 2, 3, 4)
    _GEMM_MODE = mode


def get_gemm_mode():
    return _GEMM_MODE


def _f32(t):
    if t.dtype != torch.float32:
        raise _lib.VgtkbError(f"expected float32, got {t.dtype}")
    return t.contiguous()


def _i32(t):
    return t.to(torch.int32).contiguous()


# ----------------------------------------------------------------------------- literal ops
def ball_query(new_xyz, xyz, radius, nsample):
    """vgtk.cuda.grouping.ball_query: new_xyz [B,3,M], xyz [B,3,N] -> int32 [B,M,nsample]."""
    new_xyz, xyz = _f32(new_xyz), _f32(xyz)
    b, _, m = new_xyz.shape
    n = xyz.shape[2]
    idx = torch.empty((b, m, nsample), dtype=torch.int32, device=xyz.device)
    call("vgtkb_ball_query", xyz.device, b, n, m, float(radius), int(nsample), ptr(new_xyz), ptr(xyz), ptr(idx))
    return idx


def furthest_point_sampling(xyz, m):
    """vgtk.cuda.grouping.furthest_point_sampling: xyz [B,3,N] -> int32 [B,m]."""
    xyz = _f32(xyz)
    b, _, n = xyz.shape
    idx = torch.zeros((b, m), dtype=torch.int32, device=xyz.device)
    call("vgtkb_furthest_point_sampling", xyz.device, b, n, int(m), ptr(xyz), ptr(idx))
    return idx


def fps_plain(xyz, m):
    """Plain farthest point sampling (torch_cluster.fps semantics, start at point 0): xyz [B,3,N] -> int32 [B,m]."""
    xyz = _f32(xyz)
    b, _, n = xyz.shape
    idx = torch.zeros((b, m), dtype=torch.int32, device=xyz.device)
    call("vgtkb_fps_plain", xyz.device, b, n, int(m), ptr(xyz), ptr(idx))
    return idx


def gather_points_forward(points, idx):
    """vgtk.cuda.gathering.gather_points_forward: points [B,C,N], idx [B,M] -> [B,C,M]."""
    points, idx = _f32(points), _i32(idx)
    b, c, n = points.shape
    m = idx.shape[1]
    out = torch.empty((b, c, m), dtype=torch.float32, device=points.device)
    call("vgtkb_gather_points_forward", points.device, b, c, n, m, ptr(points), ptr(idx), ptr(out))
    return out


def gather_points_backward(grad_out, idx, npoint):
    grad_out, idx = _f32(grad_out), _i32(idx)
    b, c, m = grad_out.shape
    out = torch.empty((b, c, npoint), dtype=torch.float32, device=grad_out.device)
    call("vgtkb_gather_points_backward", grad_out.device, b, c, int(npoint), m, ptr(grad_out), ptr(idx), ptr(out))
    return out


def chamfer_forward(xyz1, xyz2):
    """chamfer.forward: xyz1 [B,n,3], xyz2 [B,m,3] -> dist1 [B,n], dist2 [B,m], idx1, idx2 (int32)."""
    xyz1, xyz2 = _f32(xyz1), _f32(xyz2)
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    dev = xyz1.device
    d1 = torch.empty((b, n), dtype=torch.float32, device=dev)
    d2 = torch.empty((b, m), dtype=torch.float32, device=dev)
    i1 = torch.empty((b, n), dtype=torch.int32, device=dev)
    i2 = torch.empty((b, m), dtype=torch.int32, device=dev)
    call("vgtkb_chamfer_forward", dev, b, n, ptr(xyz1), m, ptr(xyz2), ptr(d1), ptr(d2), ptr(i1), ptr(i2))
    return d1, d2, i1, i2


def chamfer_backward(xyz1, xyz2, idx1, idx2, grad_dist1, grad_dist2):
    xyz1, xyz2, g1, g2 = _f32(xyz1), _f32(xyz2), _f32(grad_dist1), _f32(grad_dist2)
    idx1, idx2 = _i32(idx1), _i32(idx2)
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    gx1, gx2 = torch.empty_like(xyz1), torch.empty_like(xyz2)
    call("vgtkb_chamfer_backward", xyz1.device, b, n, ptr(xyz1), m, ptr(xyz2), ptr(idx1), ptr(idx2), ptr(g1), ptr(g2),
         ptr(gx1), ptr(gx2))
    return gx1, gx2


def inter_weights(xyz, sample_xyz, idx, rot_kernels, sigma):
    """Materialised kernel-point correlation w [B,P,A,K,nn] (API parity; fused path never stores it)."""
    xyz, sample_xyz, idx, rk = _f32(xyz), _f32(sample_xyz), _i32(idx), _f32(rot_kernels)
    b, _, n = xyz.shape
    p, nn = idx.shape[1], idx.shape[2]
    a, k = rk.shape[0], rk.shape[1]
    w = torch.empty((b, p, a, k, nn), dtype=torch.float32, device=xyz.device)
    call("vgtkb_inter_weights", xyz.device, b, n, p, nn, a, k, ptr(xyz), ptr(sample_xyz), ptr(idx), ptr(rk),
         float(sigma), ptr(w))
    return w


# ----------------------------------------------------------------------------- raw fused ops
def gemm_nt(a, b, bias=None, mode=None, b_planes=None):
    """C[M,N] = A[M,K] @ B[N,K]^T (+ bias).  b_planes: the prepared bf16 hi | lo planes of B (WeightPlanes; modes 3 / 4,
    K % 8 == 0) -- B itself is then not read and may be given as its shape (N, K)."""
    a = _f32(a)
    m, k = a.shape
    mode = _GEMM_MODE if mode is None else mode
    if b_planes is not None:
        n = b_planes.numel() // (2 * k)
        assert mode in (3, 4) and k % 8 == 0 and b_planes.numel() == 2 * n * k
        c = torch.empty((m, n), dtype=torch.float32, device=a.device)
        call("vgtkb_gemm_nt", a.device, m, n, k, ptr(a), None, ptr(bias.contiguous()) if bias is not None else None,
             ptr(c), mode, ptr(b_planes))
        return c
    b = _f32(b)
    n = b.shape[0]
    assert b.shape[1] == k
    c = torch.empty((m, n), dtype=torch.float32, device=a.device)
    ws = torch.empty(2 * n * k, dtype=torch.float32, device=a.device) if mode in (1, 3, 4) else None   # hi/lo split of b
    call("vgtkb_gemm_nt", a.device, m, n, k, ptr(a), ptr(b), ptr(bias.contiguous()) if bias is not None else None,
         ptr(c), mode, ptr(ws))
    return c


def gemm_tn(a, b, mode=None):
    """C[M,N] = A[R,M]^T @ B[R,N]."""
    a, b = _f32(a), _f32(b)
    r, m = a.shape
    n = b.shape[1]
    assert b.shape[0] == r
    c = torch.empty((m, n), dtype=torch.float32, device=a.device)
    mode = _GEMM_MODE if mode is None else mode
    ws = torch.empty(r * m, dtype=torch.float32, device=a.device) if mode in (3, 4) else None   # bf16 hi/lo split of a
    call("vgtkb_gemm_tn", a.device, m, n, r, ptr(a), ptr(b), ptr(c), 0, mode, ptr(ws))
    return c


def col_sum(x):
    x = _f32(x)
    rows, c = x.shape
    scratch = torch.empty(2 * c, dtype=torch.float64, device=x.device)
    out = torch.empty(c, dtype=torch.float32, device=x.device)
    call("vgtkb_col_sum", x.device, rows, c, ptr(x), ptr(scratch), ptr(out))
    return out


# ----------------------------------------------------------------------------- autograd wrappers
class InterGroupFn(torch.autograd.Function):
    """feats X [B,N,A,Ci] -> G [B,P,A,K*Ci]; weights recomputed in shared memory, never stored."""

    @staticmethod
    def forward(ctx, feats, xyz, sample_xyz, idx, rot_kernels, sigma):
        feats = _f32(feats)
        b, n, a, ci = feats.shape
        p, nn = idx.shape[1], idx.shape[2]
        k = rot_kernels.shape[1]
        g = torch.empty((b, p, a, k * ci), dtype=torch.float32, device=feats.device)
        mode = 3 if _GEMM_MODE == 3 else 0          # fixed here: backward uses the arithmetic the forward ran in
        call("vgtkb_inter_group_forward", feats.device, b, n, p, nn, a, k, ci, ptr(xyz), ptr(sample_xyz), ptr(idx),
             ptr(rot_kernels), float(sigma), ptr(feats), ptr(g), mode)
        ctx.save_for_backward(xyz, sample_xyz, idx, rot_kernels)
        ctx.meta = (b, n, p, nn, a, k, ci, float(sigma), mode)
        return g

    @staticmethod
    def backward(ctx, grad_g):
        if not ctx.needs_input_grad[0]:
            return (None,) * 6
        xyz, sample_xyz, idx, rot_kernels = ctx.saved_tensors
        b, n, p, nn, a, k, ci, sigma, mode = ctx.meta
        grad_g = _f32(grad_g)
        gx = torch.zeros((b, n, a, ci), dtype=torch.float32, device=grad_g.device)
        call("vgtkb_inter_group_backward", grad_g.device, b, n, p, nn, a, k, ci, ptr(xyz), ptr(sample_xyz), ptr(idx),
             ptr(rot_kernels), sigma, ptr(grad_g), ptr(gx), mode)
        return gx, None, None, None, None, None


# ----------------------------------------------------------------------------- operand planes
# A producer kernel (norm apply forward / backward) can write the bf16 hi / lo planes of its fp32 result; the contraction
# that consumes the tensor then loads its operand by TMA without converting it.  The planes travel beside the fp32 tensor
# in this table, keyed by the tensor's address (views keep it); an entry keeps the fp32 tensor alive, so the address cannot
# be reused while the entry exists, and is dropped by its (single) consumer.  A miss just means the fp32 operand is used.
_PLANES = {}
_PLANES_MAX = 64
PLANE_STATS = {"hit": 0, "miss": 0}


def register_planes(t, hi, lo):
    if len(_PLANES) >= _PLANES_MAX:
        _PLANES.pop(next(iter(_PLANES)))
    _PLANES[t.data_ptr()] = (t, t._version, hi, lo)


def take_planes(t, keep=False):
    """(hi, lo) planes registered for the storage `t` views (same start, same size, unmodified since), else (None, None)."""
    e = _PLANES.get(t.data_ptr()) if t.is_cuda else None
    if e is None or e[0].numel() != t.numel() or e[0]._version != e[1] or not t.is_contiguous():
        PLANE_STATS["miss"] += 1
        return None, None
    if not keep:
        del _PLANES[t.data_ptr()]
    PLANE_STATS["hit"] += 1
    return e[2], e[3]


def clear_planes():
    _PLANES.clear()


def planes_enabled():
    return _GEMM_MODE == 3 and _USE_PLANES


_USE_PLANES = True


def _alloc_planes(t):
    return (torch.empty(t.shape, dtype=torch.bfloat16, device=t.device), torch.empty(t.shape, dtype=torch.bfloat16, device=t.device))


def inter_conv_supported(b, n, p, nn, a, k, ci, co):
    """Shapes vgtkb_inter_conv_forward/backward take (modes 3 and 4; the others run InterGroupFn + LinearFn)."""
    return _GEMM_MODE in (3, 4) and bool(_lib.load().vgtkb_inter_conv_supported(b, n, p, nn, a, k, ci, co))


class PreparedWeight:
    """One conv weight of a model whose operand planes are produced by WeightPlanes.prepare(): `param` is the parameter in the
    reference's layout ([co, ci*k] with column c*k + kp, or the 1x1 conv's [co, ci, 1, 1]); `fwd` / `bwd` are the bf16 hi | lo
    planes (one bf16 tensor [2 * co*ci*k], hi then lo) of the forward operand and of the data-gradient operand.  valid(): the planes were
    computed from the parameter's current value (its version counter has not moved since prepare())."""
    __slots__ = ("param", "co", "ci", "k", "role", "fwd", "bwd", "version")

    def __init__(self, param, co, ci, k, role):
        self.param, self.co, self.ci, self.k, self.role = param, co, ci, k, role
        self.fwd = self.bwd = None
        self.version = -1

    def valid(self):
        return self.fwd is not None and self.param._version == self.version and _GEMM_MODE in (3, 4)

    def matrix(self):
        """The parameter as the [co, ci*k] matrix of the reference layout (a view)."""
        return self.param.view(self.co, self.ci * self.k)

    def kc(self):
        """[co, k*ci] with column kp*ci + c: the column order of the kernels (autograd-tracked copy)."""
        return self.param.view(self.co, self.ci, self.k).transpose(1, 2).reshape(self.co, -1)

    def grad_from_kc(self, gw_kc):
        """Gradient w.r.t. kc() -> gradient w.r.t. matrix() (the parameter's own column order)."""
        return gw_kc.view(self.co, self.k, self.ci).transpose(1, 2).reshape(self.co, self.ci * self.k)


class WeightPlanes:
    """bf16 hi / lo operand planes of ALL conv weights of a model, produced by ONE launch per step (vgtkb_weight_planes)
    instead of a permute / transpose copy plus a split launch per contraction (the weights change once per step).
    roles: 'inter' / 'intra' (BasicSO3Conv.W [co, ci*k], column c*k + kp; vgtk/vgtk/so3conv/modules.py:31-36 of the
    reference) and 'linear' (1x1 conv).  Destination orders: forward [co][k][ci]; data gradient [k][ci][co] (inter: W^T of the
    kernel-order matrix), [ci][k][co] (intra: IntraConvFn.backward's operand), [ci][co] (linear)."""

    ELEMS_PER_BLOCK = 2048

    def __init__(self):
        self.weights, self._table, self._ptrs, self._blocks, self._buf = [], None, None, 0, None

    def add(self, param, co, ci, k, role):
        assert role in ("inter", "intra", "linear") and param.numel() == co * ci * k and (co * ci * k) % 8 == 0
        assert param.is_contiguous() and param.dtype == torch.float32
        pw = PreparedWeight(param, co, ci, k, role)
        self.weights.append(pw)
        self._table = None
        return pw

    @staticmethod
    def _items(pw):
        co, ci, k = pw.co, pw.ci, pw.k
        fwd = ((co, k, ci), (ci * k, 1, k))
        if pw.role == "inter":
            bwd = ((k, ci, co), (1, k, ci * k))
        elif pw.role == "intra":
            bwd = ((ci, k, co), (k, 1, ci * k))
        else:
            bwd = ((1, ci, co), (0, 1, ci))
        return fwd, bwd

    def _build(self, dev):
        offs, total = [], 0
        for pw in self.weights:                          # two plane pairs (hi | lo) per weight, each 128-byte aligned
            n = pw.co * pw.ci * pw.k
            for _ in range(2):
                offs.append(total)
                total += (2 * n + 63) // 64 * 64
        self._buf = torch.empty(max(total, 64), dtype=torch.bfloat16, device=dev)
        base, rows, blocks = self._buf.data_ptr(), [], 0
        for i, pw in enumerate(self.weights):
            n = pw.co * pw.ci * pw.k
            for j, (sizes, strides) in enumerate(self._items(pw)):
                o = offs[2 * i + j]
                rows.append([pw.param.data_ptr(), base + 2 * o, base + 2 * (o + n), *sizes, *strides, blocks])
                blocks += (n + self.ELEMS_PER_BLOCK - 1) // self.ELEMS_PER_BLOCK
                if j == 0:
                    pw.fwd = self._buf[o:o + 2 * n]
                else:
                    pw.bwd = self._buf[o:o + 2 * n]
        self._table = torch.tensor(rows, dtype=torch.int64).to(dev)
        self._ptrs = [pw.param.data_ptr() for pw in self.weights]
        self._blocks = blocks

    def prepare(self):
        """Recompute every plane from the current parameter values: one launch on the current stream."""
        if not self.weights:
            return
        dev = self.weights[0].param.device
        if not dev.type == "cuda":
            raise _lib.VgtkbError("WeightPlanes.prepare: parameters are not on a CUDA device (there is no CPU path)")
        if self._table is None or self._table.device != dev or self._ptrs != [pw.param.data_ptr() for pw in self.weights]:
            self._build(dev)
        call("vgtkb_weight_planes", dev, len(self.weights) * 2, ptr(self._table), self._blocks)
        for pw in self.weights:
            pw.version = pw.param._version


class GradSlot:
    """Hand-over of a gradient between the two consumers of a block's input (the skip branch and the inter conv):
    backward runs the skip branch first (it was recorded later), which DEPOSITS its input gradient here and returns None;
    InterConvFn.backward then lets the scatter kernel add onto the deposited buffer (mode | 256) and returns the sum --
    instead of a memset of its own buffer plus autograd's separate add.  `closed` is set once the inter conv's backward has
    run: a later deposit attempt (other execution order, second backward over a retained graph) returns its gradient the
    normal way."""

    def __init__(self, shape):
        self.shape, self.buf, self.closed, self.armed = tuple(shape), None, False, False

    def deposit(self, g):
        """g: contiguous fp32 tensor holding the gradient in [B,N,A,Ci] order; True if it was taken."""
        if self.closed or not self.armed or self.buf is not None or g.numel() != self.shape[0] * self.shape[1] * self.shape[2] * self.shape[3]:
            return False
        self.buf = g.view(self.shape)
        return True


class InterConvFn(torch.autograd.Function):
    """InterSO3Conv in one call per direction (vgtkb_inter_conv_forward / _backward): feats X [B,N,A,Ci] channels-last,
    w_kc [Co, K*Ci] -> out [B*P*A, Co].  The grouped tensor exists only as the two bf16 operand planes the grouping
    kernel writes (kept for the weight gradient); reference: so3conv/functional.py:144-203, spconv/functional.py:375-406,
    so3conv/modules.py:48-55."""

    @staticmethod
    def forward(ctx, feats, w_kc, xyz, sample_xyz, idx, rot_kernels, sigma, slot=None, wp=None):
        # wp (PreparedWeight, valid): `w_kc` is the PARAMETER in its own layout; the kernels read the prepared planes and the
        # weight gradient is returned in the parameter's layout (no permute copy, no transpose, no split launches)
        feats = _f32(feats)
        ctx.wp = wp
        if wp is None:
            w_kc = _f32(w_kc)
        b, n, a, ci = feats.shape
        ctx.slot = slot if (slot is not None and slot.shape == (b, n, a, ci) and feats.requires_grad) else None
        if ctx.slot is not None:
            ctx.slot.armed = True
        p, nn = idx.shape[1], idx.shape[2]
        k, co = rot_kernels.shape[1], w_kc.shape[0]
        dev = feats.device
        rows = b * p * a
        mode = _GEMM_MODE                      # 3 = bf16x3 (both planes), 4 = single-pass bf16 (hi plane only)
        g_hi = torch.empty((rows, k * ci), dtype=torch.bfloat16, device=dev)
        g_lo = torch.empty((rows, k * ci), dtype=torch.bfloat16, device=dev) if mode == 3 else None
        out = torch.empty((rows, co), dtype=torch.float32, device=dev)
        if wp is not None:
            assert (wp.co, wp.ci, wp.k) == (co, ci, k) and wp.valid()
            call("vgtkb_inter_conv_forward", dev, b, n, p, nn, a, k, ci, co, ptr(xyz), ptr(sample_xyz), ptr(idx), ptr(rot_kernels),
                 float(sigma), ptr(feats), None, ptr(g_hi), ptr(g_lo), ptr(wp.fwd), ptr(out), mode)
            ctx.wp_bwd = wp.bwd
        else:
            ws = torch.empty(co * k * ci, dtype=torch.float32, device=dev)
            call("vgtkb_inter_conv_forward", dev, b, n, p, nn, a, k, ci, co, ptr(xyz), ptr(sample_xyz), ptr(idx), ptr(rot_kernels),
                 float(sigma), ptr(feats), ptr(w_kc), ptr(g_hi), ptr(g_lo), ptr(ws), ptr(out), mode)
        ctx.save_for_backward(xyz, sample_xyz, idx, rot_kernels, w_kc, g_hi, g_lo)
        ctx.meta = (b, n, p, nn, a, k, ci, co, float(sigma), mode)
        return out

    @staticmethod
    def backward(ctx, gy):
        xyz, sample_xyz, idx, rot_kernels, w_kc, g_hi, g_lo = ctx.saved_tensors
        b, n, p, nn, a, k, ci, co, sigma, mode = ctx.meta
        gy = _f32(gy)
        dev = gy.device
        rows, kc = b * p * a, k * ci
        need_x, need_w = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        acc = None
        if ctx.slot is not None:
            acc, ctx.slot.buf, ctx.slot.closed = ctx.slot.buf, None, True
        if acc is not None and need_x:
            gx, mode = acc, mode | 256                  # the skip branch's gradient: the scatter adds onto it
        else:
            gx = torch.empty((b, n, a, ci), dtype=torch.float32, device=dev) if need_x else None
        dg = torch.empty((rows, kc), dtype=torch.float32, device=dev) if need_x else None
        gw = torch.empty((co, kc), dtype=torch.float32, device=dev) if need_w else None
        gy_hi, gy_lo = take_planes(gy) if (mode & 255) == 3 else (None, None)
        wp = ctx.wp
        if wp is not None and (gy_hi is not None or not need_w):
            # prepared planes of W^T: the workspace argument carries them (w_kc = NULL)
            call("vgtkb_inter_conv_backward", dev, b, n, p, nn, a, k, ci, co, ptr(xyz), ptr(sample_xyz), ptr(idx),
                 ptr(rot_kernels), sigma, None, ptr(g_hi), ptr(g_lo), ptr(gy), ptr(gy_hi), ptr(gy_lo), ptr(dg), ptr(gx), ptr(gw),
                 ptr(ctx.wp_bwd), mode)
        else:
            if wp is not None:
                w_kc = wp.kc().detach().contiguous()
            ws = torch.empty(max(rows * co, 2 * kc * co), dtype=torch.float32, device=dev)
            call("vgtkb_inter_conv_backward", dev, b, n, p, nn, a, k, ci, co, ptr(xyz), ptr(sample_xyz), ptr(idx),
                 ptr(rot_kernels), sigma, ptr(w_kc), ptr(g_hi), ptr(g_lo), ptr(gy), ptr(gy_hi), ptr(gy_lo), ptr(dg), ptr(gx),
                 ptr(gw), ptr(ws), mode)
        if acc is not None and not need_x:
            raise _lib.VgtkbError("InterConvFn: a skip-branch gradient was deposited but the input gradient is not requested")
        if wp is not None and gw is not None:
            gw = wp.grad_from_kc(gw)                    # the parameter's own layout
        return gx, gw, None, None, None, None, None, None, None


def pose_neighbourhood(xyz, pose, idx, anchors, with_perm=True, sample_xyz=None, sample_idx=None):
    """xyz [B,3,N], pose [B,N,4,4], idx [B,P,nn] int32, anchors [A,3,3] -> rotated neighbour offsets [B,P,nn,3] and the
    anchor permutation table [B,P,nn,A] uint8 (None when with_perm is False).  Strided layers pass the P centres:
    sample_xyz [B,3,P] = xyz[sample_idx], sample_idx [B,P] (their rotations are pose[sample_idx]); default P = N."""
    xyz, pose, idx, anchors = _f32(xyz), _f32(pose), _i32(idx), _f32(anchors)
    b, _, n = xyz.shape
    p, nn, a = idx.shape[1], idx.shape[2], anchors.shape[0]
    if sample_idx is None:
        if p != n:
            raise _lib.VgtkbError("pose_neighbourhood: idx has P != N centres but no sample_idx / sample_xyz was given")
        sxyz, sidx = xyz, None
    else:
        sxyz, sidx = _f32(sample_xyz), _i32(sample_idx)
    rel = torch.empty((b, p, nn, 3), dtype=torch.float32, device=xyz.device)
    perm = torch.empty((b, p, nn, a), dtype=torch.uint8, device=xyz.device) if with_perm else None
    call("vgtkb_pose_neighbourhood_strided", xyz.device, b, n, p, nn, a, ptr(xyz), ptr(pose), ptr(sxyz), ptr(sidx), ptr(idx),
         ptr(anchors), ptr(rel), ptr(perm))
    return rel, perm


class PoseGroupFn(torch.autograd.Function):
    """Pose-aware inter grouping: feats X [B,N,A,Ci] -> G [B,P,A,K*Ci] with rotated offsets and anchor permutation
    (P = idx.shape[1] centres; P = N without stride)."""

    @staticmethod
    def forward(ctx, feats, idx, rel_xyz, perm, rot_kernels, sigma):
        feats = _f32(feats)
        b, n, a, ci = feats.shape
        p, nn, k = idx.shape[1], idx.shape[2], rot_kernels.shape[1]
        g = torch.empty((b, p, a, k * ci), dtype=torch.float32, device=feats.device)
        call("vgtkb_inter_pose_group_forward_strided", feats.device, b, n, p, nn, a, k, ci, ptr(idx), ptr(rel_xyz), ptr(perm),
             ptr(rot_kernels), float(sigma), ptr(feats), ptr(g))
        ctx.save_for_backward(idx, rel_xyz, rot_kernels)
        ctx.perm = perm
        ctx.meta = (b, n, p, nn, a, k, ci, float(sigma))
        return g

    @staticmethod
    def backward(ctx, grad_g):
        if not ctx.needs_input_grad[0]:
            return (None,) * 6
        idx, rel_xyz, rot_kernels = ctx.saved_tensors
        b, n, p, nn, a, k, ci, sigma = ctx.meta
        grad_g = _f32(grad_g)
        gx = torch.zeros((b, n, a, ci), dtype=torch.float32, device=grad_g.device)
        call("vgtkb_inter_pose_group_backward_strided", grad_g.device, b, n, p, nn, a, k, ci, ptr(idx), ptr(rel_xyz), ptr(ctx.perm),
             ptr(rot_kernels), sigma, ptr(grad_g), ptr(gx))
        return gx, None, None, None, None, None


class IntraGroupFn(torch.autograd.Function):
    """Y [R,A,C] -> G [R,A,KK*C] with G[r,a,k,:] = Y[r, intra_idx[a,k], :]."""

    @staticmethod
    def forward(ctx, y, intra_idx):
        y = _f32(y)
        r, a, c = y.shape
        kk = intra_idx.shape[1]
        g = torch.empty((r, a, kk * c), dtype=torch.float32, device=y.device)
        call("vgtkb_intra_group_forward", y.device, r, a, kk, c, ptr(intra_idx), ptr(y), ptr(g))
        ctx.save_for_backward(intra_idx)
        ctx.meta = (r, a, kk, c)
        return g

    @staticmethod
    def backward(ctx, grad_g):
        (intra_idx,) = ctx.saved_tensors
        r, a, kk, c = ctx.meta
        grad_g = _f32(grad_g)
        gy = torch.empty((r, a, c), dtype=torch.float32, device=grad_g.device)
        call("vgtkb_intra_group_backward", grad_g.device, r, a, kk, c, ptr(intra_idx), ptr(grad_g), ptr(gy))
        return gy, None


def gather_gemm_nt(x, table, w, bias=None, mode=None):
    """out[(pt,a), o] = sum_{kk,c} x[pt, table[a,kk], c] * w[o, kk*C + c];  x [points, A, C] -> [points*A, n]."""
    x, w = _f32(x), _f32(w)
    pts, a, c = x.shape
    kk, n = table.shape[1], w.shape[0]
    assert w.shape[1] == kk * c and table.shape[0] == a and table.dtype == torch.int32
    mode = _GEMM_MODE if mode is None else mode
    out = torch.empty((pts * a, n), dtype=torch.float32, device=x.device)
    ws = torch.empty(2 * n * kk * c, dtype=torch.float32, device=x.device)
    call("vgtkb_gather_gemm_nt", x.device, pts, a, kk, c, n, ptr(table), ptr(x), ptr(w),
         ptr(bias.contiguous()) if bias is not None else None, ptr(out), mode, ptr(ws))
    return out


def gather_gemm_tn(x, table, y, mode=None):
    """out[o, kk*C + c] = sum_{pt,a} y[(pt,a), o] * x[pt, table[a,kk], c];  -> [m, kk*C]."""
    x, y = _f32(x), _f32(y)
    pts, a, c = x.shape
    kk, m = table.shape[1], y.shape[1]
    assert y.shape[0] == pts * a
    mode = _GEMM_MODE if mode is None else mode
    out = torch.empty((m, kk * c), dtype=torch.float32, device=x.device)
    ws = torch.empty(pts * a * m, dtype=torch.float32, device=x.device) if mode in (3, 4) else None
    call("vgtkb_gather_gemm_tn", x.device, pts, a, kk, c, m, ptr(table), ptr(x), ptr(y), ptr(out), 0, mode, ptr(ws))
    return out


def gather_gemm_supported(c_in, c_out, points):
    """Shapes the fused intra-conv path takes (else: materialised gather + GEMM)."""
    return _GEMM_MODE != 0 and c_in % 64 == 0 and c_out % 64 == 0 and points >= 64


def gather_gemm_nt_planes(x_hi, x_lo, table, w, bias=None, w_planes=None):
    """gather_gemm_nt with x [points, A, C] given as bf16 planes (no operand conversion in the kernel).
    w_planes: prepared bf16 hi | lo planes of w [n, kk*C] (WeightPlanes); w is then not read (may be None)."""
    pts, a, c = x_hi.shape
    kk = table.shape[1]
    assert table.shape[0] == a and table.dtype == torch.int32
    dev = x_hi.device
    if w_planes is not None:
        n = w_planes.numel() // (2 * kk * c)
        assert w_planes.numel() == 2 * n * kk * c
        out = torch.empty((pts * a, n), dtype=torch.float32, device=dev)
        call("vgtkb_gather_gemm_nt_planes", dev, pts, a, kk, c, n, ptr(table), ptr(x_hi), ptr(x_lo), None,
             ptr(bias.contiguous()) if bias is not None else None, ptr(out), ptr(w_planes))
        return out
    w = _f32(w)
    n = w.shape[0]
    assert w.shape[1] == kk * c
    out = torch.empty((pts * a, n), dtype=torch.float32, device=dev)
    ws = torch.empty(n * kk * c, dtype=torch.float32, device=dev)
    call("vgtkb_gather_gemm_nt_planes", dev, pts, a, kk, c, n, ptr(table), ptr(x_hi), ptr(x_lo), ptr(w),
         ptr(bias.contiguous()) if bias is not None else None, ptr(out), ptr(ws))
    return out


def gather_gemm_tn_planes(x_hi, x_lo, table, y, y_hi=None, y_lo=None):
    """gather_gemm_tn with x as planes and y [points*A, m] as fp32 or planes -> [m, kk*C]."""
    pts, a, c = x_hi.shape
    kk, m = table.shape[1], y.shape[1]
    out = torch.empty((m, kk * c), dtype=torch.float32, device=y.device)
    ws = torch.empty(pts * a * m, dtype=torch.float32, device=y.device) if y_hi is None else None
    call("vgtkb_gather_gemm_tn_planes", y.device, pts, a, kk, c, m, ptr(table), ptr(x_hi), ptr(x_lo), ptr(y), ptr(y_hi), ptr(y_lo),
         ptr(out), 0, ptr(ws))
    return out


class IntraConvFn(torch.autograd.Function):
    """Intra-anchor group convolution as gather-GEMMs (forward, data gradient with the inverse table, weight
    gradient); x [points, A, C] channels-last, w_kc [Co, KK*C] -> [points*A, Co].  No gathered tensor exists.
    When the producers of x / of the output gradient registered operand planes (norm kernels), the contractions read
    them instead of converting the fp32 tensors k-block by k-block (the intra conv re-reads every element 12 times)."""

    @staticmethod
    def forward(ctx, x, w_kc, table, table_inv, wp=None):
        # wp (PreparedWeight, valid): `w_kc` is the PARAMETER in its own layout (column c*kk + k); see InterConvFn
        x = _f32(x)
        x_hi, x_lo = take_planes(x) if planes_enabled() and x.shape[2] % 64 == 0 else (None, None)
        ctx.planes = x_hi is not None
        ctx.mode = _GEMM_MODE                        # backward runs in the arithmetic of the forward
        ctx.wp = wp
        if wp is not None and not ctx.planes:        # prepared weights only serve the plane-fed kernels
            w_kc = wp.kc()
        if wp is None or not ctx.planes:
            w_kc = _f32(w_kc)
        if ctx.planes:
            x_hi, x_lo = x_hi.view(x.shape), x_lo.view(x.shape)
            ctx.save_for_backward(x_hi, x_lo, w_kc, table, table_inv)
            if wp is not None:
                assert wp.valid() and wp.ci == x.shape[2]
                ctx.wp_bwd = wp.bwd
                return gather_gemm_nt_planes(x_hi, x_lo, table, None, None, wp.fwd)
            return gather_gemm_nt_planes(x_hi, x_lo, table, w_kc)
        ctx.save_for_backward(x, w_kc, table, table_inv)
        return gather_gemm_nt(x, table, w_kc)

    @staticmethod
    def backward(ctx, gy):
        if ctx.planes:
            x_hi, x_lo, w_kc, table, table_inv = ctx.saved_tensors
            pts, a, c = x_hi.shape
        else:
            x, w_kc, table, table_inv = ctx.saved_tensors
            pts, a, c = x.shape
        kk, co = table.shape[1], w_kc.shape[0]
        gy = _f32(gy)
        gy_hi, gy_lo = take_planes(gy) if ctx.mode == 3 and planes_enabled() and co % 64 == 0 else (None, None)
        gx = gw = None
        wp = ctx.wp if ctx.planes else None          # (without planes the forward already went through wp.kc(): autograd permutes)
        if ctx.needs_input_grad[0]:
            # gx[(pt,a'), c] = sum_{kk,o} gy[pt, inv[a',kk], o] * w_kc[o, kk*C + c]
            if wp is not None and gy_hi is not None:
                gx = gather_gemm_nt_planes(gy_hi.view(pts, a, co), gy_lo.view(pts, a, co), table_inv, None, None,
                                           ctx.wp_bwd).view(pts, a, c)
            else:
                if wp is not None:
                    w_kc = wp.kc().detach()
                wt = w_kc.view(co, kk, c).permute(2, 1, 0).reshape(c, kk * co).contiguous()
                if gy_hi is not None:
                    gx = gather_gemm_nt_planes(gy_hi.view(pts, a, co), gy_lo.view(pts, a, co), table_inv, wt).view(pts, a, c)
                else:
                    gx = gather_gemm_nt(gy.view(pts, a, co), table_inv, wt, mode=ctx.mode).view(pts, a, c)
        if ctx.needs_input_grad[1]:
            if ctx.planes:
                gw = gather_gemm_tn_planes(x_hi, x_lo, table, gy, gy_hi, gy_lo)
            else:
                gw = gather_gemm_tn(x, table, gy, mode=ctx.mode)
            if ctx.wp is not None:
                gw = ctx.wp.grad_from_kc(gw)         # the parameter's own layout
        return gx, gw, None, None, None


class LinearFn(torch.autograd.Function):
    """y[M,N] = x[M,K] @ w[N,K]^T + bias  (the BasicSO3Conv contraction / 1x1 skip conv)."""

    @staticmethod
    def forward(ctx, x, w, bias, mode=None, slot=None, wp=None):
        # wp (PreparedWeight of a 1x1 conv, valid): the contractions read the prepared planes of w / w^T
        x, w = _f32(x), _f32(w)
        ctx.save_for_backward(x, w)
        ctx.slot = slot
        ctx.has_bias = bias is not None
        ctx.mode = _GEMM_MODE if mode is None else mode      # fixed here: backward uses the arithmetic of the forward
        ctx.wp_bwd = None
        if wp is not None and ctx.mode in (3, 4) and wp.valid() and (wp.co, wp.ci * wp.k) == tuple(w.shape):
            ctx.wp_bwd = wp.bwd
            return gemm_nt(x, None, bias, ctx.mode, wp.fwd)
        return gemm_nt(x, w, bias, ctx.mode)

    @staticmethod
    def backward(ctx, gy):
        x, w = ctx.saved_tensors
        gy = _f32(gy)
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            if ctx.wp_bwd is not None:
                gx = gemm_nt(gy, None, None, ctx.mode, ctx.wp_bwd)
            else:
                gx = gemm_nt(gy, w.t().contiguous(), None, ctx.mode)
            if ctx.slot is not None and ctx.slot.deposit(gx):
                gx = None                       # handed to the inter conv of the block (GradSlot)
        if ctx.needs_input_grad[1]:
            gw = gemm_tn(gy, x, ctx.mode)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            gb = col_sum(gy)
        return gx, gw, gb, None, None, None


class PointnetPoolFn(torch.autograd.Function):
    """out[b,o,a] = max_n (e[b,n,a,o] + v[a,o,:] . xc[b,:,n])  -- the pooled PointnetSO3Conv output
    (vgtk/vgtk/so3conv/modules.py:404-412); e [B,N,A,Co] rows of the feature part of the embedding, v [A,Co,3],
    xc [B,3,N] centred coordinates."""

    @staticmethod
    def forward(ctx, e, v, xc):
        e, v, xc = _f32(e), _f32(v), _f32(xc)
        b, n, a, co = e.shape
        out = torch.empty((b, co, a), dtype=torch.float32, device=e.device)
        arg = torch.empty((b, a, co), dtype=torch.int32, device=e.device)
        call("vgtkb_pointnet_pool_forward", e.device, b, n, a, co, ptr(e), ptr(v), ptr(xc), ptr(out), ptr(arg))
        ctx.save_for_backward(v, xc, arg)
        ctx.meta = (b, n, a, co)
        return out

    @staticmethod
    def backward(ctx, gout):
        v, xc, arg = ctx.saved_tensors
        b, n, a, co = ctx.meta
        gout = _f32(gout)
        ge = gv = gxc = None
        if ctx.needs_input_grad[0]:
            ge = torch.empty((b, n, a, co), dtype=torch.float32, device=gout.device)
            call("vgtkb_pointnet_pool_backward", gout.device, b, n, a, co, ptr(gout), ptr(arg), ptr(ge))
        if ctx.needs_input_grad[1] or ctx.needs_input_grad[2]:
            g = gout.permute(0, 2, 1)                                               # [B,A,Co]
            idx = arg.long().reshape(b, 1, a * co).expand(b, 3, a * co)
            if ctx.needs_input_grad[1]:
                xsel = torch.gather(xc, 2, idx).view(b, 3, a, co)                   # xc[b, :, arg[b,a,o]]
                gv = torch.einsum('bao,bjao->aoj', g, xsel)
            if ctx.needs_input_grad[2]:
                contrib = (g.unsqueeze(1) * v.permute(2, 0, 1).unsqueeze(0)).reshape(b, 3, a * co)
                gxc = torch.zeros_like(xc).scatter_add_(2, idx, contrib)
        return ge, gv, gxc


def pointnet_embed_xyz_(e, v, xc):
    """e [B,N,A,Co] += v[a,o,:] . xc[b,:,n] in place (PointnetSO3Conv with return_raw=True; inference only)."""
    b, n, a, co = e.shape
    call("vgtkb_pointnet_embed_xyz", e.device, b, n, a, co, ptr(e), ptr(_f32(v)), ptr(_f32(xc)))
    return e


class RowGatherFn(torch.autograd.Function):
    """x [B,N,W] , idx [B,M] -> [B,M,W] (skip-connection sub-sampling by sample_idx)."""

    @staticmethod
    def forward(ctx, x, idx, slot=None):
        x = _f32(x)
        b, n, w = x.shape
        m = idx.shape[1]
        out = torch.empty((b, m, w), dtype=torch.float32, device=x.device)
        call("vgtkb_row_gather_forward", x.device, b, n, m, w, ptr(x), ptr(idx), ptr(out))
        ctx.slot = slot
        ctx.save_for_backward(idx)
        ctx.meta = (b, n, m, w)
        return out

    @staticmethod
    def backward(ctx, gout):
        (idx,) = ctx.saved_tensors
        b, n, m, w = ctx.meta
        gout = _f32(gout)
        gx = torch.zeros((b, n, w), dtype=torch.float32, device=gout.device)
        call("vgtkb_row_gather_backward", gout.device, b, n, m, w, ptr(gout), ptr(idx), ptr(gx))
        if ctx.slot is not None and ctx.slot.deposit(gx):
            return None, None, None             # handed to the inter conv of the block (GradSlot)
        return gx, None, None


def _sync_world(group):
    """Number of ranks a SyncBatchNorm reduction spans (1 = plain BatchNorm).  `group`: False/None = off,
    True = the default process group, or a torch.distributed process group."""
    if group is None or group is False:
        return 1
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized():
        return 1
    return dist.get_world_size(None if group is True else group)


# dataparallel.PeerMailbox per process group: SyncBatchNorm exchanges over NVLink peer memory instead of NCCL.  A mailbox
# only serves the group it was built for (its rank / world / peer pointers / exchange counter are that group's); a
# BatchNorm synchronised over another group takes the NCCL path.
_PEER_MAILBOXES = {}
_PEER_MAILBOX = None     # mailbox of the default group (kept as a plain attribute for callers that test `is not None`)


def _group_key(group):
    return "default" if group is None or group is True else id(group)


def set_peer_mailbox(mailbox, group=None):
    """Register (or, with None, remove) the peer mailbox serving `group` (None / True = the default process group)."""
    global _PEER_MAILBOX
    key = _group_key(group if mailbox is None else getattr(mailbox, "group", group))
    if mailbox is None:
        _PEER_MAILBOXES.pop(key, None)
    else:
        _PEER_MAILBOXES[key] = mailbox
    _PEER_MAILBOX = _PEER_MAILBOXES.get("default")


def peer_mailbox_for(group):
    return _PEER_MAILBOXES.get(_group_key(group))


def _all_reduce_sums(scratch, group):
    mb = peer_mailbox_for(group)
    if mb is not None and scratch.is_cuda:
        mb.all_reduce_sums(scratch)
        return
    import torch.distributed as dist
    dist.all_reduce(scratch, op=dist.ReduceOp.SUM, group=None if group is True else group)


class NormActFn(torch.autograd.Function):
    """leaky_relu(norm(x)) [+ residual] on rows x [groups, rows, C].

    groups == 1: BatchNorm2d in training mode (batch statistics; running stats updated in place)
    groups == B: InstanceNorm2d(affine=False).  With `use_running` the running statistics are
    used instead (eval-mode BatchNorm).
    `sync_group` (groups == 1 only): SyncBatchNorm -- the fp64 channel sums (and the row count riding behind them) are
    all-reduced over the ranks between the two phases of the statistics and of the backward
    (nn.SyncBatchNorm semantics: SPConvNets/trainer_unsup_arti_align.py:430 converts every BatchNorm)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, residual, running_mean, running_var, momentum, eps, slope, use_running, sync_group=None,
                planes_fwd=False, planes_bwd=False):
        x = _f32(x)
        g, rows, c = x.shape
        dev = x.device
        stats = torch.empty((g, 2, c), dtype=torch.float32, device=dev)
        sync = (not use_running) and g == 1 and _sync_world(sync_group) > 1
        rm = ptr(running_mean) if running_mean is not None else None
        rv = ptr(running_var) if running_var is not None else None
        if use_running:
            stats[:, 0] = running_mean
            stats[:, 1] = torch.rsqrt(running_var + eps)
        elif sync:
            scratch = torch.empty(2 * c + 1, dtype=torch.float64, device=dev)
            scratch[2 * c:].fill_(float(rows))
            call("vgtkb_norm_sums", dev, 1, rows, c, ptr(x), ptr(scratch))
            mb = peer_mailbox_for(sync_group)
            if mb is not None:      # exchange + finalize in one kernel over peer memory
                call("vgtkb_norm_finalize_peer", dev, c, float(eps), ptr(scratch), ptr(stats), rm, rv, float(momentum),
                     mb.rank, mb.world, mb.ptrs, mb.next_seq())
            else:
                _all_reduce_sums(scratch, sync_group)
                call("vgtkb_norm_finalize", dev, 1, 0, c, float(eps), ptr(scratch), ptr(stats), rm, rv, float(momentum))
        else:
            scratch = torch.empty((g, 2, c), dtype=torch.float64, device=dev)
            call("vgtkb_norm_stats", dev, g, rows, c, ptr(x), float(eps), ptr(scratch), ptr(stats), rm, rv, float(momentum))
        y = torch.empty_like(x)
        gam = gamma.contiguous() if gamma is not None else None
        bet = beta.contiguous() if beta is not None else None
        res = _f32(residual) if residual is not None else None
        pl_ok = planes_enabled() and c % 4 == 0 and 256 % (c // 4) == 0
        if planes_fwd and pl_ok:
            y_hi, y_lo = _alloc_planes(y)
            call("vgtkb_norm_act_forward_planes", dev, g, rows, c, ptr(x), ptr(stats), ptr(gam), ptr(bet), float(slope), ptr(res),
                 ptr(y), ptr(y_hi), ptr(y_lo))
            register_planes(y, y_hi, y_lo)
        else:
            call("vgtkb_norm_act_forward", dev, g, rows, c, ptr(x), ptr(stats), ptr(gam), ptr(bet), float(slope), ptr(res),
                 ptr(y))
        ctx.save_for_backward(x, stats, gam, bet)
        ctx.meta = (g, rows, c, float(slope), use_running, residual is not None, sync_group if sync else None)
        ctx.planes_bwd = bool(planes_bwd and pl_ok)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, stats, gam, bet = ctx.saved_tensors
        g, rows, c, slope, use_running, has_res, sync_group = ctx.meta
        gy = _f32(gy)
        dev = gy.device
        gx = torch.empty_like(x)
        ggam = torch.empty(c, dtype=torch.float32, device=dev) if gam is not None else None
        gbet = torch.empty(c, dtype=torch.float32, device=dev) if bet is not None else None
        gx_hi, gx_lo = _alloc_planes(gx) if ctx.planes_bwd else (None, None)
        if use_running:
            # eval-mode BatchNorm (frozen statistics, e.g. fine-tuning with bn.eval()): the statistics do not depend on x, so
            # gx = gamma * invstd * dyp, ggamma = sum dyp * xhat, gbeta = sum dyp -- the sums kernel as is, and the apply
            # kernel with the two mean terms zeroed
            scratch = torch.empty((g, 2, c), dtype=torch.float64, device=dev)
            call("vgtkb_norm_bwd_sums", dev, g, rows, c, ptr(x), ptr(stats), ptr(gam), ptr(bet), slope, ptr(gy),
                 ptr(scratch), ptr(ggam), ptr(gbet))
            scratch.zero_()
            call("vgtkb_norm_bwd_apply_planes", dev, g, rows, rows, c, ptr(x), ptr(stats), ptr(gam), ptr(bet), slope, ptr(gy),
                 ptr(scratch), ptr(gx), ptr(gx_hi), ptr(gx_lo))
        elif sync_group is not None:
            scratch = torch.empty(2 * c + 1, dtype=torch.float64, device=dev)
            scratch[2 * c:].fill_(float(rows))
            call("vgtkb_norm_bwd_sums", dev, 1, rows, c, ptr(x), ptr(stats), ptr(gam), ptr(bet), slope, ptr(gy),
                 ptr(scratch), ptr(ggam), ptr(gbet))       # affine gradients: local sums, reduced with the other parameters
            _all_reduce_sums(scratch, sync_group)
            call("vgtkb_norm_bwd_apply_planes", dev, 1, rows, 0, c, ptr(x), ptr(stats), ptr(gam), ptr(bet), slope, ptr(gy),
                 ptr(scratch), ptr(gx), ptr(gx_hi), ptr(gx_lo))
        else:
            scratch = torch.empty((g, 2, c), dtype=torch.float64, device=dev)
            call("vgtkb_norm_act_backward_planes", dev, g, rows, c, ptr(x), ptr(stats), ptr(gam), ptr(bet), slope, ptr(gy),
                 ptr(scratch), ptr(gx), ptr(ggam), ptr(gbet), ptr(gx_hi), ptr(gx_lo))
        if gx_hi is not None:
            register_planes(gx, gx_hi, gx_lo)
        return gx, ggam, gbet, (gy if has_res else None), None, None, None, None, None, None, None, None, None


class SyncNormPairFn(torch.autograd.Function):
    """Two SyncBatchNorm + leaky_relu passes of one block -- the inter conv's norm and the skip branch's norm, same channel
    count and row count -- with ONE exchange per direction instead of two: the local fp64 sums of both tensors travel in one
    buffer ([sum, sum of squares | row count] twice).  Every exchange is a point where the ranks wait for the slowest one
    (measured ~30 us each at 2 and at 8 GPUs against ~5 us for the exchange itself), so the classic backbone goes from 28
    to 14 of them per step.  Arithmetic per norm: exactly NormActFn's sync path (same kernels, same sums, rank-ordered
    addition), so the results are those of two separate calls.
    x1, x2: [1, rows, C]; returns (y1, y2).  nn.SyncBatchNorm semantics as in NormActFn."""

    @staticmethod
    def forward(ctx, x1, gamma1, beta1, rm1, rv1, mom1, slope1, planes_fwd1, planes_bwd1,
                x2, gamma2, beta2, rm2, rv2, mom2, slope2, planes_fwd2, planes_bwd2, eps, sync_group):
        x1, x2 = _f32(x1), _f32(x2)
        g, rows, c = x1.shape
        assert g == 1 and x2.shape == x1.shape
        dev = x1.device
        blk = 2 * c + 1
        scratch = torch.empty(2 * blk, dtype=torch.float64, device=dev)
        scratch.view(2, blk)[:, 2 * c].fill_(float(rows))
        xs, outs, saved = (x1, x2), [], []
        for i in range(2):
            call("vgtkb_norm_sums", dev, 1, rows, c, ptr(xs[i]), ptr(scratch[i * blk:(i + 1) * blk]))
        _all_reduce_sums(scratch, sync_group)
        pl_ok = planes_enabled() and c % 4 == 0 and 256 % (c // 4) == 0
        pbs = []
        for i, (gamma, beta, rm, rv, mom, slope, pf, pb) in enumerate(((gamma1, beta1, rm1, rv1, mom1, slope1, planes_fwd1, planes_bwd1),
                                                                       (gamma2, beta2, rm2, rv2, mom2, slope2, planes_fwd2, planes_bwd2))):
            x = xs[i]
            stats = torch.empty((1, 2, c), dtype=torch.float32, device=dev)
            call("vgtkb_norm_finalize", dev, 1, 0, c, float(eps), ptr(scratch[i * blk:(i + 1) * blk]), ptr(stats),
                 ptr(rm) if rm is not None else None, ptr(rv) if rv is not None else None, float(mom))
            y = torch.empty_like(x)
            gam = gamma.contiguous() if gamma is not None else None
            bet = beta.contiguous() if beta is not None else None
            if pf and pl_ok:
                y_hi, y_lo = _alloc_planes(y)
                call("vgtkb_norm_act_forward_planes", dev, 1, rows, c, ptr(x), ptr(stats), ptr(gam), ptr(bet), float(slope), None,
                     ptr(y), ptr(y_hi), ptr(y_lo))
                register_planes(y, y_hi, y_lo)
            else:
                call("vgtkb_norm_act_forward", dev, 1, rows, c, ptr(x), ptr(stats), ptr(gam), ptr(bet), float(slope), None, ptr(y))
            outs.append(y)
            saved += [x, stats, gam, bet]
            pbs.append(bool(pb and pl_ok))
        ctx.save_for_backward(*saved)
        ctx.meta = (rows, c, float(slope1), float(slope2), sync_group, pbs)
        return outs[0], outs[1]

    @staticmethod
    def backward(ctx, gy1, gy2):
        saved = ctx.saved_tensors
        rows, c, slope1, slope2, sync_group, pbs = ctx.meta
        dev = saved[0].device
        blk = 2 * c + 1
        scratch = torch.empty(2 * blk, dtype=torch.float64, device=dev)
        scratch.view(2, blk)[:, 2 * c].fill_(float(rows))
        gys, slopes, res = [gy1, gy2], (slope1, slope2), []
        for i in range(2):
            x, stats, gam, bet = saved[4 * i:4 * i + 4]
            gys[i] = _f32(gys[i]) if gys[i] is not None else torch.zeros_like(x)
            ggam = torch.empty(c, dtype=torch.float32, device=dev) if gam is not None else None
            gbet = torch.empty(c, dtype=torch.float32, device=dev) if bet is not None else None
            call("vgtkb_norm_bwd_sums", dev, 1, rows, c, ptr(x), ptr(stats), ptr(gam), ptr(bet), slopes[i], ptr(gys[i]),
                 ptr(scratch[i * blk:(i + 1) * blk]), ptr(ggam), ptr(gbet))   # affine gradients: local sums (reduced with the parameters)
            res.append((ggam, gbet))
        _all_reduce_sums(scratch, sync_group)
        gxs = []
        for i in range(2):
            x, stats, gam, bet = saved[4 * i:4 * i + 4]
            gx = torch.empty_like(x)
            gx_hi, gx_lo = _alloc_planes(gx) if pbs[i] else (None, None)
            call("vgtkb_norm_bwd_apply_planes", dev, 1, rows, 0, c, ptr(x), ptr(stats), ptr(gam), ptr(bet), slopes[i], ptr(gys[i]),
                 ptr(scratch[i * blk:(i + 1) * blk]), ptr(gx), ptr(gx_hi), ptr(gx_lo))
            if gx_hi is not None:
                register_planes(gx, gx_hi, gx_lo)
            gxs.append(gx)
        n = (None,)
        return (gxs[0], res[0][0], res[0][1]) + n * 6 + (gxs[1], res[1][0], res[1][1]) + n * 6 + n * 2


def norm_act(x, gamma=None, beta=None, residual=None, running_mean=None, running_var=None, momentum=0.1, eps=1e-5,
             slope=LEAKY_SLOPE, use_running=False, sync_group=None, planes_fwd=False, planes_bwd=False):
    """planes_fwd: also write the bf16 operand planes of the result (its consumer is a tensor-core contraction);
    planes_bwd: likewise for the input gradient in backward (the producer of x is a contraction)."""
    return NormActFn.apply(x, gamma, beta, residual, running_mean, running_var, momentum, eps, slope, use_running, sync_group,
                           planes_fwd, planes_bwd)


class AnchorChamferFn(torch.autograd.Function):
    """Chamfer terms of A rigidly transformed copies of a reconstruction against one input cloud, fused
    (SPConvNets/models/unsup_seg_so3_pose_conv_pn_38_multi_stage.py:429-436):
        Y[b,a] = rot[b,a] canon[b] + trans[b,a];  dist1 [B,A,M] = Y -> ori,  dist2 [B,A,N] = ori -> Y.
    canon [B,M,3], rot [B,A,3,3], trans [B,A,3], ori [B,N,3]."""

    @staticmethod
    def forward(ctx, canon, rot, trans, ori):
        canon, rot, trans, ori = _f32(canon), _f32(rot), _f32(trans), _f32(ori)
        b, m, _ = canon.shape
        a, n, dev = rot.shape[1], ori.shape[1], canon.device
        d1 = torch.empty((b, a, m), dtype=torch.float32, device=dev)
        d2 = torch.empty((b, a, n), dtype=torch.float32, device=dev)
        i1 = torch.empty((b, a, m), dtype=torch.int32, device=dev)
        i2 = torch.empty((b, a, n), dtype=torch.int32, device=dev)
        call("vgtkb_anchor_chamfer_forward", dev, b, a, m, ptr(canon), ptr(rot), ptr(trans), n, ptr(ori), ptr(d1), ptr(d2),
             ptr(i1), ptr(i2))
        ctx.save_for_backward(canon, rot, trans, ori, i1, i2)
        ctx.mark_non_differentiable(i1, i2)
        return d1, d2, i1, i2

    @staticmethod
    def backward(ctx, g1, g2, _gi1, _gi2):
        canon, rot, trans, ori, i1, i2 = ctx.saved_tensors
        b, m, _ = canon.shape
        a, n, dev = rot.shape[1], ori.shape[1], canon.device
        gy = torch.empty((b, a, m, 3), dtype=torch.float32, device=dev)
        gori = torch.empty_like(ori) if ctx.needs_input_grad[3] else None
        call("vgtkb_anchor_chamfer_backward", dev, b, a, m, ptr(canon), ptr(rot), ptr(trans), n, ptr(ori), ptr(i1), ptr(i2),
             ptr(_f32(g1)), ptr(_f32(g2)), ptr(gy), ptr(gori))
        # fold dL/dY into the pose and the canonical points (B*A*M*3 values: tiny next to the matching)
        g_rot = torch.einsum('bamj,bmk->bajk', gy, canon) if ctx.needs_input_grad[1] else None
        g_trans = gy.sum(2) if ctx.needs_input_grad[2] else None
        g_canon = torch.einsum('bajk,bamj->bmk', rot, gy) if ctx.needs_input_grad[0] else None
        return g_canon, g_rot, g_trans, gori


def anchor_chamfer(canon, rot, trans, ori):
    """-> dist1 [B,A,M] (transformed reconstruction -> input), dist2 [B,A,N] (input -> reconstruction), idx1, idx2."""
    return AnchorChamferFn.apply(canon, rot, trans, ori)


class ChamferFn(torch.autograd.Function):
    """extensions/chamfer_dist/__init__.py:13-26 (ChamferFunction)."""

    @staticmethod
    def forward(ctx, xyz1, xyz2):
        d1, d2, i1, i2 = chamfer_forward(xyz1, xyz2)
        ctx.save_for_backward(xyz1, xyz2, i1, i2)
        ctx.mark_non_differentiable(i1, i2)
        return d1, d2, i1, i2

    @staticmethod
    def backward(ctx, g1, g2, _gi1, _gi2):
        xyz1, xyz2, i1, i2 = ctx.saved_tensors
        gx1, gx2 = chamfer_backward(xyz1, xyz2, i1, i2, g1.contiguous(), g2.contiguous())
        return gx1, gx2


# ----------------------------------------------------------------------------- PointNet++ set abstraction
def knn_query(pos, centers, k):
    """SPConvNets/models/PointNet2.py:85-87: pos [B,N,3], centers [B,S,3] -> idx int32 [B,S,k], dist [B,S,k]
    (ascending; dist = sqrt of the squared distance in torch's evaluation order)."""
    pos, centers = _f32(pos), _f32(centers)
    b, n, d = pos.shape
    if d != 3 or centers.shape[2] != 3:
        raise _lib.VgtkbError("knn_query: 3-D coordinates expected")
    s = centers.shape[1]
    idx = torch.empty((b, s, k), dtype=torch.int32, device=pos.device)
    dist = torch.empty((b, s, k), dtype=torch.float32, device=pos.device)
    call("vgtkb_knn_query", pos.device, b, n, s, int(k), ptr(pos), ptr(centers), ptr(idx), ptr(dist))
    return idx, dist


def _pad8(c):
    return (c + 7) // 8 * 8


class SaGroupFn(torch.autograd.Function):
    """rows [B,S,k,cpad] = [pos[idx] - centre | feat[idx] | 0]  (PointNet2.py:92-100; idx None: the identity
    neighbourhood of the global level :151-156).  Differentiable in feat; coordinates are data."""

    @staticmethod
    def forward(ctx, feat, pos, centers, idx, cpad):
        pos = _f32(pos)
        b, n, _ = pos.shape
        c = 0 if feat is None else feat.shape[2]
        feat = _f32(feat) if feat is not None else None
        if idx is None:
            s, k = 1, n
        else:
            s, k = idx.shape[1], idx.shape[2]
        out = torch.empty((b, s, k, cpad), dtype=torch.float32, device=pos.device)
        cen = _f32(centers) if centers is not None else None
        call("vgtkb_sa_group_forward", pos.device, b, n, s, k, c, int(cpad), ptr(pos), ptr(feat), ptr(cen), ptr(idx), ptr(out))
        ctx.idx = idx
        ctx.meta = (b, n, s, k, c, int(cpad))
        return out

    @staticmethod
    def backward(ctx, gout):
        b, n, s, k, c, cpad = ctx.meta
        if ctx.needs_input_grad[1] or ctx.needs_input_grad[2]:
            raise NotImplementedError("sa_group: gradients with respect to the coordinates are not part of this path")
        gfeat = None
        if c > 0 and ctx.needs_input_grad[0]:
            gout = _f32(gout)
            gfeat = torch.empty((b, n, c), dtype=torch.float32, device=gout.device)
            call("vgtkb_sa_group_backward", gout.device, b, n, s, k, c, cpad, ptr(gout), ptr(ctx.idx), ptr(gfeat))
        return gfeat, None, None, None, None


def sa_group(feat, pos, centers, idx, cpad=None):
    c = 0 if feat is None else feat.shape[2]
    return SaGroupFn.apply(feat, pos, centers, idx, _pad8(3 + c) if cpad is None else cpad)


class SaMaxPoolFn(torch.autograd.Function):
    """PointNet2.py:102-112: y [G,k,C], dist [G,k] or None, r -> [G,C]; entries beyond the radius count as -1e8."""

    @staticmethod
    def forward(ctx, y, dist, radius):
        y = _f32(y)
        g, k, c = y.shape
        out = torch.empty((g, c), dtype=torch.float32, device=y.device)
        arg = torch.empty((g, c), dtype=torch.int32, device=y.device)
        d = _f32(dist) if (dist is not None and radius is not None) else None
        call("vgtkb_sa_maxpool_forward", y.device, g, k, c, ptr(y), ptr(d), float(radius) if radius is not None else 0.0,
             ptr(out), ptr(arg))
        ctx.save_for_backward(arg)
        ctx.meta = (g, k, c)
        return out

    @staticmethod
    def backward(ctx, gout):
        (arg,) = ctx.saved_tensors
        g, k, c = ctx.meta
        gout = _f32(gout)
        gy = torch.empty((g, k, c), dtype=torch.float32, device=gout.device)
        call("vgtkb_sa_maxpool_backward", gout.device, g, k, c, ptr(gout), ptr(arg), ptr(gy))
        return gy, None, None


def sa_maxpool(y, dist=None, radius=None):
    return SaMaxPoolFn.apply(y, dist, radius)


def three_nn(p1, p2):
    """PointNet2.py:114-123: for every point of p2 [B,n2,3] the min(3,n1) nearest points of p1 [B,n1,3] and their
    normalised inverse-distance weights -> idx int32 [B,n2,3], w [B,n2,3]."""
    p1, p2 = _f32(p1), _f32(p2)
    b, n1, _ = p1.shape
    n2 = p2.shape[1]
    idx = torch.empty((b, n2, 3), dtype=torch.int32, device=p1.device)
    w = torch.empty((b, n2, 3), dtype=torch.float32, device=p1.device)
    call("vgtkb_three_nn", p1.device, b, n1, n2, ptr(p1), ptr(p2), ptr(idx), ptr(w))
    return idx, w


class ThreeInterpolateFn(torch.autograd.Function):
    """PointNet2.py:126-128: out [B,n2,C] = sum_j feat[b, idx[b,q,j], :] * w[b,q,j]."""

    @staticmethod
    def forward(ctx, feat, idx, w):
        feat, w = _f32(feat), _f32(w)
        b, n1, c = feat.shape
        n2 = idx.shape[1]
        out = torch.empty((b, n2, c), dtype=torch.float32, device=feat.device)
        call("vgtkb_three_interpolate_forward", feat.device, b, n1, n2, c, ptr(feat), ptr(idx), ptr(w), ptr(out))
        ctx.save_for_backward(idx, w)
        ctx.meta = (b, n1, n2, c)
        return out

    @staticmethod
    def backward(ctx, gout):
        idx, w = ctx.saved_tensors
        b, n1, n2, c = ctx.meta
        gout = _f32(gout)
        gfeat = torch.empty((b, n1, c), dtype=torch.float32, device=gout.device)
        call("vgtkb_three_interpolate_backward", gout.device, b, n1, n2, c, ptr(gout), ptr(idx), ptr(w), ptr(gfeat))
        return gfeat, None, None


def three_interpolate(feat, idx, w):
    return ThreeInterpolateFn.apply(feat, idx, w)


# ----------------------------------------------------------------------------- plane-operand contractions (raw form)
def split_bf16(x):
    """fp32 tensor -> (hi, lo) bf16 planes with hi = bf16_rn(x), lo = bf16_rn(x - hi) (the operand split of gemm mode 3)."""
    x = _f32(x)
    hi = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    lo = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    call("vgtkb_split_bf16", x.device, x.numel(), ptr(x), ptr(hi), ptr(lo))
    return hi, lo


def gemm_nt_presplit(a_hi, a_lo, b, bias=None):
    """C[M,N] = (a_hi + a_lo)[M,K] @ b[N,K]^T (+ bias): the bf16x3 contraction with its activation operand already stored
    as bf16 planes (no in-kernel conversion); the raw form of what InterConvFn / IntraConvFn use."""
    if a_hi.dtype != torch.bfloat16 or a_lo.dtype != torch.bfloat16 or a_hi.shape != a_lo.shape:
        raise _lib.VgtkbError("gemm_nt_presplit: two bf16 planes of the same shape expected")
    b = _f32(b)
    m, k = a_hi.shape
    n = b.shape[0]
    assert b.shape[1] == k
    c = torch.empty((m, n), dtype=torch.float32, device=b.device)
    ws = torch.empty(n * k, dtype=torch.float32, device=b.device)
    call("vgtkb_gemm_nt_presplit", b.device, m, n, k, ptr(a_hi.contiguous()), ptr(a_lo.contiguous()), ptr(b),
         ptr(bias.contiguous()) if bias is not None else None, ptr(c), ptr(ws))
    return c


def gemm_tn_presplit(a, b_hi, b_lo):
    """C[M,N] = a[R,M]^T @ (b_hi + b_lo)[R,N]: the bf16x3 weight-gradient contraction with its wide operand stored as bf16
    planes; the raw form of what InterConvFn's weight gradient uses."""
    if b_hi.dtype != torch.bfloat16 or b_lo.dtype != torch.bfloat16 or b_hi.shape != b_lo.shape:
        raise _lib.VgtkbError("gemm_tn_presplit: two bf16 planes of the same shape expected")
    a = _f32(a)
    r, m = a.shape
    n = b_hi.shape[1]
    assert b_hi.shape[0] == r
    c = torch.empty((m, n), dtype=torch.float32, device=a.device)
    ws = torch.empty(r * m, dtype=torch.float32, device=a.device)
    call("vgtkb_gemm_tn_presplit", a.device, m, n, r, ptr(a), ptr(b_hi.contiguous()), ptr(b_lo.contiguous()), ptr(c), 0, ptr(ws))
    return c
