"""ORACLE (test infrastructure, never on the product path): the loss of the reference's training step, restated.

`main.py:404-440` (inline in `train()`, so it cannot be imported): with `rec = nn.L1Loss()` (`utils.set_rec_loss`, default
`--loss_type L1`, utils.py:613-622),

    rec_D1 = lambda1/3 * sum_idx rec(GT_idx, S'_idx)                              Eq.(9)   idx = S0, S1, St
    rec_D2 = sum_{i=1..N_trn} lambda2/3 * sum_idx rec(GT_idx, S_final[i][idx])    Eq.(10)
    total  = rec_D1 + rec_D2

The reference adds the three terms and divides the running sum by 3 when idx == 2 (`:413-414`, `:432-433`), which is the same
number up to fp32 rounding of the running sum; this restatement follows that order.  Parity: the arithmetic is torch's own
`nn.L1Loss`; no reference fixture exists for it.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def rec_losses(sharps_prime, sharps_final, S0_GT, S1_GT, St_GT, rec_D1_lambda=1.0, rec_D2_lambda=1.0):
    gts = [S0_GT, S1_GT, St_GT]
    rec_D1 = 0.0
    rec_D2 = 0.0
    for idx in range(3):
        rec_D1 = rec_D1 + rec_D1_lambda * F.l1_loss(gts[idx], sharps_prime[idx])
        rec_D2 = rec_D2 + rec_D2_lambda * F.l1_loss(gts[idx], sharps_final[0][idx])
    rec_D1 = rec_D1 / 3
    rec_D2 = rec_D2 / 3
    for i in range(len(sharps_final) - 1):
        tmp = 0.0
        for idx in range(3):
            tmp = tmp + rec_D2_lambda * F.l1_loss(gts[idx], sharps_final[i + 1][idx])
        rec_D2 = rec_D2 + tmp / 3
    return rec_D1 + rec_D2, rec_D1, rec_D2
