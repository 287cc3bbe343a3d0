"""CPU: the rank-based AUC / AP helpers of `Trainer.eval` (`gnndelete_b200/metrics.py`, plain tensor ops) against
sklearn, which is what the reference calls (framework/trainer/base.py:247-248, 276-277)."""
import numpy as np
import pytest
import torch
from sklearn.metrics import average_precision_score, roc_auc_score

from gnndelete_b200 import metrics as M


@pytest.mark.parametrize('levels', [7, 50, 10 ** 6])            # heavy ties ... practically none
def test_auc_ap_match_sklearn(levels):
    g = torch.Generator().manual_seed(levels)
    label = (torch.rand(2000, generator=g) < 0.3).float()
    score = torch.randint(0, levels, (2000,), generator=g).float() / levels
    assert M.roc_auc(label, score) == pytest.approx(roc_auc_score(label.numpy(), score.numpy()), abs=1e-12)
    assert M.average_precision(label, score) == pytest.approx(average_precision_score(label.numpy(), score.numpy()), abs=1e-12)


@pytest.mark.parametrize('levels,chunk', [(40, 32), (10 ** 6, 7), (3, 1)])
def test_resampled_auc_ap_match_sklearn(levels, chunk):
    g = torch.Generator().manual_seed(1)
    n = 257
    neg = torch.randint(0, levels, (n,), generator=g).float() / levels
    pool = torch.randint(0, levels + levels // 4, (4000,), generator=g).float() / levels
    idx = torch.stack([torch.randperm(4000, generator=g)[:n] for _ in range(20)])
    auc, ap = M.resampled_auc_ap(neg, pool, idx, chunk=chunk)
    assert auc.shape == ap.shape == (20,) and auc.dtype == torch.float64
    lab = np.r_[np.zeros(n), np.ones(n)]
    for b in range(20):
        sc = np.r_[neg.numpy(), pool[idx[b]].numpy()]
        assert auc[b].item() == pytest.approx(roc_auc_score(lab, sc), abs=1e-12)
        assert ap[b].item() == pytest.approx(average_precision_score(lab, sc), abs=1e-12)
