"""vgtk.cuda.zpconv (reference: vgtk/vgtk/cuda/zpconv_cuda.cpp:113-118), served by libvgtkb200.so.

Reference tensor layouts, explicit index / weight tensors:
    inter_zpconv_forward(idx[B,P,A,K,ann] int32, w[B,P,A,K,ann], feats[B,C,Nq,A]) -> [B,C,K,P,A]
    intra_zpconv_forward(idx[Aout,ann] int32, w[Aout,K,ann], feats[B,C,P,Ain])    -> [B,C,K,P,Aout]
In the reference these are dead code on the SO(3) path; they are the grouping of the legacy S^2 ZPConv modules."""
import torch

from equi_articulated_pose_b200 import lib as _lib

call, ptr = _lib.call, _lib.ptr


def _prep(idx, w, x):
    for t, name in ((idx, "idx"), (w, "weights"), (x, "feats")):
        if not t.is_cuda:
            raise RuntimeError(f"{name} must be a CUDA tensor")
    return idx.to(torch.int32).contiguous(), w.float().contiguous(), x.float().contiguous()


def inter_zpconv_forward(idx, w, feats):
    idx, w, feats = _prep(idx, w, feats)
    b, p, a, k, ann = idx.shape
    c, nq = feats.shape[1], feats.shape[2]
    out = torch.empty((b, c, k, p, a), dtype=torch.float32, device=feats.device)
    call("vgtkb_inter_zpconv_forward", feats.device, b, c, nq, p, a, k, ann, ptr(idx), ptr(w), ptr(feats), ptr(out))
    return out


def inter_zpconv_backward(idx, w, grad_out, nq):
    idx, w, grad_out = _prep(idx, w, grad_out)
    b, p, a, k, ann = idx.shape
    c = grad_out.shape[1]
    gf = torch.empty((b, c, int(nq), a), dtype=torch.float32, device=grad_out.device)
    call("vgtkb_inter_zpconv_backward", grad_out.device, b, c, int(nq), p, a, k, ann, ptr(idx), ptr(w), ptr(grad_out), ptr(gf))
    return gf


def intra_zpconv_forward(idx, w, feats):
    idx, w, feats = _prep(idx, w, feats)
    aout, ann = idx.shape
    k = w.shape[1]
    b, c, p, ain = feats.shape
    out = torch.empty((b, c, k, p, aout), dtype=torch.float32, device=feats.device)
    call("vgtkb_intra_zpconv_forward", feats.device, b, c, p, ain, aout, k, ann, ptr(idx), ptr(w), ptr(feats), ptr(out))
    return out


def intra_zpconv_backward(idx, w, grad_out, ain):
    idx, w, grad_out = _prep(idx, w, grad_out)
    aout, ann = idx.shape
    b, c, k, p, _ = grad_out.shape
    gf = torch.empty((b, c, p, int(ain)), dtype=torch.float32, device=grad_out.device)
    call("vgtkb_intra_zpconv_backward", grad_out.device, b, c, p, int(ain), aout, k, ann, ptr(idx), ptr(w), ptr(grad_out), ptr(gf))
    return gf
