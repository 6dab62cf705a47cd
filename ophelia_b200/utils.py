"""Model-side helpers of the reference's utils.py (`utils.py:155-170`): attention guide and Noam learning rate.
(The DSP half of utils.py -- STFT features, Griffin-Lim, plotting -- is outside the hot path.)"""
import numpy as np


def get_attention_guide(xdim, ydim, g=0.2):
    '''Guided attention. Refer to page 3 on the paper.  W[n,t] = 1 - exp(-(t/T - n/N)^2 / (2 g^2)), float32.
    (Host copy for plotting/tests; the training kernels evaluate the same expression analytically.)'''
    n = np.arange(xdim, dtype=np.float64)[:, None] / float(xdim)
    t = np.arange(ydim, dtype=np.float64)[None, :] / float(ydim)
    return (1.0 - np.exp(-(t - n) ** 2 / (2 * g * g))).astype(np.float32)


def get_global_attention_guide(hp):
    return get_attention_guide(hp.max_N, hp.max_T, g=hp.g)


def learning_rate_decay(init_lr, global_step, warmup_steps=4000.0):
    '''Noam scheme from tensor2tensor (host mirror of adam_prepare_kernel).'''
    step = float(global_step + 1)
    return init_lr * warmup_steps ** 0.5 * min(step * warmup_steps ** -1.5, step ** -0.5)
