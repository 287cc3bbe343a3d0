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


def resampled_auc_ap(neg_score, pos_pool, idx, chunk=32):
    """AUC and AP of ``B`` balanced problems at once: problem ``b`` scores the negatives ``neg_score [n]``
    (label 0: the deleted edges) against the positives ``pos_pool[idx[b]]`` (label 1: one random sample of
    retained edges), as ``Trainer.eval`` does 500 times per call (``base.py:256-280``).  Same tie handling
    as :func:`roc_auc` / :func:`average_precision` (= sklearn's), but two batched passes instead of
    ``2 B`` sorts:

    * AUC = (#{pos > neg} + 0.5 #{pos == neg}) / (n_pos n_neg): the per-positive credit against the fixed
      negatives is two ``searchsorted`` calls on the whole pool, a problem is then a gather + row sum
      (half-integers in float64: exact);
    * AP: ``chunk`` problems are sorted side by side (one 2-D sort), precision is taken at the end of every
      run of equal scores and weighted by the recall step since the previous run.

    Returns float64 tensors ``(auc [B], ap [B])`` on the scores' device; no host sync."""
    neg = neg_score.double().flatten()
    pool = pos_pool.double().flatten()
    n_neg, n_pos = neg.numel(), idx.shape[1]
    neg_sorted = torch.sort(neg)[0]
    less = torch.searchsorted(neg_sorted, pool, right=False)
    leq = torch.searchsorted(neg_sorted, pool, right=True)
    credit = less.double() + 0.5 * (leq - less).double()
    chunk = max(1, min(chunk, (1 << 25) // max(n_neg + n_pos, 1)))      # <= 256 MB per float64 temporary
    aucs, aps = [], []
    for b0 in range(0, idx.shape[0], chunk):
        ix = idx[b0:b0 + chunk]
        aucs.append(credit[ix].sum(1) / (n_pos * n_neg))
        rows = ix.shape[0]
        score = torch.cat([neg.unsqueeze(0).expand(rows, -1), pool[ix]], 1)
        s, order = torch.sort(score, dim=1, descending=True)
        y = (order >= n_neg).double()                                   # columns >= n_neg are the positives
        end = torch.ones_like(s, dtype=torch.bool)
        end[:, :-1] = s[:, 1:] != s[:, :-1]
        tps = torch.cumsum(y, 1)
        at_end = torch.where(end, tps, torch.zeros_like(tps))
        prev = torch.zeros_like(tps)                                    # true positives at the previous run end
        prev[:, 1:] = torch.cummax(at_end, 1)[0][:, :-1]
        rank = torch.arange(1, s.shape[1] + 1, dtype=torch.float64, device=s.device)
        step = at_end - torch.where(end, prev, torch.zeros_like(prev))  # recall step * n_pos at run ends, 0 elsewhere
        aps.append((step / n_pos * (tps / rank)).sum(1))
    return torch.cat(aucs), torch.cat(aps)
