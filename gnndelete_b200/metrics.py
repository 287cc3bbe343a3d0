"""Rank-based ROC-AUC and average precision in torch (device-resident), with sklearn's
tie handling (``roc_auc_score`` / ``average_precision_score`` as called at
``framework/trainer/base.py:247-248, 276-277``).  Evaluation-side helper — not part of the
timed hot path."""
from __future__ import annotations

import torch


def _group_ends(sorted_scores):
    """Indices of the last element of every run of equal values."""
    n = sorted_scores.numel()
    change = torch.ones(n, dtype=torch.bool, device=sorted_scores.device)
    change[:-1] = sorted_scores[1:] != sorted_scores[:-1]
    return change.nonzero().squeeze(1)


def roc_auc(label, score, as_tensor=False):
    label = label.double().flatten()
    score = score.double().flatten()
    order = torch.argsort(score)
    s, y = score[order], label[order]
    ends = _group_ends(s)
    starts = torch.cat([ends.new_zeros(1), ends[:-1] + 1])
    avg_rank = (starts + ends).double() / 2 + 1          # 1-based average rank of each tie group
    counts = (ends - starts + 1)
    ranks = torch.repeat_interleave(avg_rank, counts)
    n_pos = y.sum()
    n_neg = y.numel() - n_pos
    auc = ((ranks * y).sum() - n_pos * (n_pos + 1) / 2) / (n_pos * n_neg)
    return auc if as_tensor else auc.item()


def average_precision(label, score, as_tensor=False):
    label = label.double().flatten()
    score = score.double().flatten()
    order = torch.argsort(score, descending=True)
    s, y = score[order], label[order]
    ends = _group_ends(s)
    tps = torch.cumsum(y, 0)[ends]
    fps = (ends + 1).double() - tps
    precision = tps / (tps + fps)
    recall = tps / tps[-1]
    prev = torch.cat([recall.new_zeros(1), recall[:-1]])
    ap = ((recall - prev) * precision).sum()
    return ap if as_tensor else ap.item()
