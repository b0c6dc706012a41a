"""Oracle restatement of the merged max-mixture pose prior (test infrastructure only).

Follows reference utils/smplify/prior.py:102-174 (constant set-up from the GMM pickle) and
:181-196 (``merged_log_likelihood``).  Input: the GMM arrays (means (8,69), covars (8,69,69),
weights (8,), float64) as shipped in the reference's data/gmm_08.pkl (re-saved as npz).
"""
import numpy as np
import torch


def gmm_constants(gmm, dtype=torch.float32):
    """prior.py:126-160: fp32 means, fp32-cast inverse covariances, and the merged mixture
    weights ``w_m / ((2 pi)^(69/2) * sqrt(det C_m) / min_m sqrt(det C_m))``."""
    np_dtype = np.float32 if dtype == torch.float32 else np.float64
    means = gmm['means'].astype(np_dtype)
    covs = gmm['covars'].astype(np_dtype)
    precisions = np.stack([np.linalg.inv(c) for c in covs]).astype(np_dtype)
    sqrdets = np.array([np.sqrt(np.linalg.det(c)) for c in gmm['covars']])
    const = (2 * np.pi) ** (69 / 2.0)
    nll_weights = np.asarray(gmm['weights'] / (const * (sqrdets / sqrdets.min())))
    return dict(means=torch.tensor(means, dtype=dtype),
                precisions=torch.tensor(precisions, dtype=dtype),
                nll_weights=torch.tensor(nll_weights, dtype=dtype).unsqueeze(0))


def merged_nll(pose69, consts):
    """prior.py:181-196: min over components of 0.5 (x-mu)^T P (x-mu) - log(nll_weight)."""
    diff = pose69.unsqueeze(1) - consts['means']
    pd = torch.einsum('mij,bmj->bmi', consts['precisions'], diff)
    quad = (pd * diff).sum(-1)
    ll = 0.5 * quad - torch.log(consts['nll_weights'])
    return ll.min(dim=1)[0]
