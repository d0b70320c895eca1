"""Oracle (TEST INFRASTRUCTURE ONLY) of the reference's optimizer step -- radam.py:15-78 (``RAdam.step``) restated with
numpy in float64/float32, one tensor at a time.  Pinned against the REAL radam.RAdam by tests/golden/optim_radam.npz
(tests/golden/make_golden.py `optim_radam`); SGD and Adam need no restatement: the tests use torch.optim on the CPU.
Only tests/ may import this module."""
import math

import numpy as np


def radam_step(w, g, exp_avg, exp_avg_sq, step, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
    """One radam.py step on float32 numpy arrays (updated in place); ``step`` = number of steps taken before this one.
    radam.py:44-45 moment updates, :47-65 N_sma / step_size, :67-68 weight decay, :71-75 update."""
    beta1, beta2 = betas
    g = g.astype(np.float32)
    exp_avg_sq *= np.float32(beta2)
    exp_avg_sq += np.float32(1 - beta2) * g * g
    exp_avg *= np.float32(beta1)
    exp_avg += np.float32(1 - beta1) * g
    t = step + 1
    beta2_t = beta2 ** t
    n_sma_max = 2 / (1 - beta2) - 1
    n_sma = n_sma_max - 2 * t * beta2_t / (1 - beta2_t)
    if n_sma >= 5:
        step_size = lr * math.sqrt((1 - beta2_t) * (n_sma - 4) / (n_sma_max - 4) * (n_sma - 2) / n_sma * n_sma_max / (n_sma_max - 2)) / (1 - beta1 ** t)
    else:
        step_size = lr / (1 - beta1 ** t)
    if weight_decay != 0:
        w += np.float32(-weight_decay * lr) * w
    if n_sma >= 5:
        w += np.float32(-step_size) * exp_avg / (np.sqrt(exp_avg_sq) + np.float32(eps))
    else:
        w += np.float32(-step_size) * exp_avg
    return n_sma
