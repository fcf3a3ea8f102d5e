"""vgtk.loss -- the two losses the reference trainer instantiates
(reference: vgtk/vgtk/loss.py; trainer_unsup_arti_align.py:455).  Head-side, outside the hot path."""
import torch
import torch.nn as nn
import torch.nn.functional as F


class CrossEntropyLoss(nn.Module):
    def forward(self, pred, label):
        loss = F.cross_entropy(pred, label.long().view(-1))
        acc = (pred.argmax(1) == label.view(-1)).float().mean()
        return loss, acc


class CrossEntropyLossPerP(nn.Module):
    """per-point cross entropy: pred [b,c,n], label [b,n]"""

    def forward(self, pred, label):
        loss = F.cross_entropy(pred, label.long())
        acc = (pred.argmax(1) == label).float().mean()
        return loss, acc


class AttentionCrossEntropyLoss(nn.Module):
    """classification loss + anchor-attention loss weighted by `beta`"""

    def __init__(self, loss_type, loss_margin):
        super().__init__()
        self.loss_type, self.loss_margin = loss_type, loss_margin

    def forward(self, pred, label, wts, rlabel, pretrain_step=2000):
        cls_loss = F.cross_entropy(pred, label.long().view(-1))
        r_loss = F.cross_entropy(wts, rlabel.long().view(-1)) if rlabel is not None else pred.new_zeros(())
        acc = (pred.argmax(1) == label.view(-1)).float().mean()
        r_acc = (wts.argmax(1) == rlabel.view(-1)).float().mean() if rlabel is not None else None
        return cls_loss + self.loss_margin * r_loss, cls_loss, r_loss, acc, r_acc
