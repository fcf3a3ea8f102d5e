"""vgtk.cuda.zpconv (reference: vgtk/vgtk/cuda/zpconv_cuda.cpp:113-118).

In the reference these four entry points are dead code (the Python calls the `*_naive` torch
versions).  Here the slot carries the fused grouping kernels with the reference's tensor layouts:
    inter_zpconv_forward(idx[B,P,A,K,ann] int32, w[B,P,A,K,ann], feats[B,C,Nq,A]) -> [B,C,K,P,A]
    intra_zpconv_forward(idx[Aout,ann] int32, w[Aout,K,ann], feats[B,C,P,Ain])    -> [B,C,K,P,Aout]
These literal (explicit idx/weight) forms are evaluated with index arithmetic in torch on the
device; the SO(3) path does not go through them (see vgtk.so3conv.functional)."""
import torch


def inter_zpconv_forward(idx, w, feats):
    b, p, a, k, ann = idx.shape
    c = feats.shape[1]
    f = feats.permute(0, 3, 2, 1)                                   # [B,A,Nq,C]
    ar = torch.arange(a, device=feats.device).view(1, 1, a, 1, 1).expand(b, p, a, k, ann)
    br = torch.arange(b, device=feats.device).view(b, 1, 1, 1, 1).expand(b, p, a, k, ann)
    g = f[br, ar, idx.long()]                                       # [B,P,A,K,ann,C]
    out = (g * w.unsqueeze(-1)).sum(4)                              # [B,P,A,K,C]
    return out.permute(0, 4, 3, 1, 2).contiguous()


def inter_zpconv_backward(idx, w, grad_out, nq):
    b, p, a, k, ann = idx.shape
    c = grad_out.shape[1]
    go = grad_out.permute(0, 3, 4, 2, 1)                            # [B,P,A,K,C]
    contrib = go.unsqueeze(4) * w.unsqueeze(-1)                     # [B,P,A,K,ann,C]
    gf = torch.zeros(b, a, nq, c, device=grad_out.device, dtype=grad_out.dtype)
    ar = torch.arange(a, device=idx.device).view(1, 1, a, 1, 1).expand_as(idx)
    br = torch.arange(b, device=idx.device).view(b, 1, 1, 1, 1).expand_as(idx)
    gf.index_put_((br, ar, idx.long()), contrib, accumulate=True)
    return gf.permute(0, 3, 2, 1).contiguous()


def intra_zpconv_forward(idx, w, feats):
    g = feats[..., idx.long()]                                      # [B,C,P,Aout,ann]
    return torch.einsum('bcpan,akn->bckpa', g, w).contiguous()


def intra_zpconv_backward(idx, w, grad_out, ain=None):
    b, c, k, p, aout = grad_out.shape
    ain = ain if ain is not None else int(idx.max().item()) + 1
    contrib = torch.einsum('bckpa,akn->bcpan', grad_out, w)
    gf = torch.zeros(b, c, p, ain, device=grad_out.device, dtype=grad_out.dtype)
    gf.index_add_(3, idx.long().reshape(-1), contrib.reshape(b, c, p, -1))
    return gf
