"""compute_truncation_residual (pyLOM/vmmath/truncation.py:17-39, src/truncation.c:46-74): O(n) host
scalar work on the n singular values -- stays on the host like in the reference's Cython path."""
import numpy as np
import torch


def compute_truncation_residual(S, r):
    """r > 0: first N with ||S[N:]||_2/||S||_2 < r ;  r < 0: first N whose cumulative sum(S)/sum(S) > |r|."""
    s = S.detach().cpu().numpy() if isinstance(S, torch.Tensor) else np.asarray(S)
    N = 0
    if r > 0:
        normS = np.linalg.norm(s, 2)
        for ii in range(s.shape[0]):
            if np.linalg.norm(s[ii:], 2) / normS < r:
                break
            N += 1
    else:
        r = abs(r)
        normS = np.sum(s)
        acc = 0
        for ii in range(s.shape[0]):
            acc += s[ii] / normS
            N += 1
            if acc > r:
                break
    return N
