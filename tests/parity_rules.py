"""The stated parity tolerances (BASELINE.json north_star: "within a stated epsilon"), in ONE place.

COEF_RTOL   DCT coefficients: max|dY| <= COEF_RTOL * max|Y| against the float64 definition, on the spectrally flat
            synthetic inputs of BASELINE configs 2-5.  The reference's own fp32 FFT path differs from the definition
            by ~8e-8 * max|Y|; ours measures 2-9e-8 there.
COEF_RTOL_NATURAL  the same bound on natural images (config 1).  Their energy sits in a few large low-frequency
            coefficients whose partial sums grow monotonically, and a direct-summation transform (K sequential fp32
            accumulations; the tensor core's accumulator additionally truncates toward zero: measured mean shrink of
            the 12 largest AC coefficients 6.8e-7 relative at 256^2) is less accurate on exactly those than the
            reference's FFT (log K depth).  Measured on the reference's 13 images at 256^2: worst 5.7e-7 * max|Y|
            with the tensor-core kernels, 5.4e-7 with the exact-fp32 FFMA kernels, 9e-8 for the reference's FFT.
EPS_SCORE   selection: the order may differ only inside runs of tokens whose reference scores differ by less
            than EPS_SCORE; at a top-k cut, tokens within EPS_SCORE of the k-th score may be swapped.
EPS_LFQ     LFQ sign bits, stated on the COEFFICIENT scale, where it is one number: a bit may differ from the
            reference's only where |Y_ref - median| < EPS_LFQ * max|Y|, EPS_LFQ = 2 * COEF_RTOL (one COEF_RTOL
            for the coefficient, one for the fitted median, which is itself a coefficient of the fitting
            batch).  SURVEY 8(c) proposed 1e-5 in NORMALISED units, which assumed a divisor b*sqrt(2) ~ 1;
            fitted tables have b ~ 0.02..0.3 per position, and (Y - median) / (b*sqrt(2) + eps) turns the
            same coefficient tolerance into 1e-5..1e-3 normalised, position by position.  Expressing the
            rule before the division keeps it a single stated constant; `lfq_bit_exempt` applies it.
EPS_VQ      VQ indices may differ only where the two best squared distances differ by < EPS_VQ relative.
"""
import numpy as np

COEF_RTOL = 4e-7
COEF_RTOL_NATURAL = 8e-7
EPS_SCORE = 1e-5
EPS_LFQ = 2 * COEF_RTOL
EPS_VQ = 1e-4


def lfq_bit_exempt(normed_ref: np.ndarray, std_tok: np.ndarray, ymax: float, eps: float = EPS_LFQ) -> np.ndarray:
    """Boolean mask of the entries of the reference's NORMALISED patches (rows, s, z) whose sign bit is allowed to
    differ: |normed| * std = |Y - median| < eps * max|Y|.  ``std_tok`` (rows, s, z) = b*sqrt(2) + eps_norm gathered
    at every token's (channel, h, w)."""
    return np.abs(normed_ref.astype(np.float64)) * std_tok < eps * ymax


def std_at_tokens(b_table: np.ndarray, channels: np.ndarray, positions: np.ndarray, eps_norm: float = 1e-6) -> np.ndarray:
    """(rows, s, z) divisor of PatchNorm (PN:157-165) at each token."""
    b = b_table[channels, positions[..., 0], positions[..., 1]]
    return b.astype(np.float64) * np.sqrt(2.0) + eps_norm


def order_equal_up_to_score_ties(keys_ours: np.ndarray, keys_ref: np.ndarray, scores_ref_sorted: np.ndarray,
                                 eps: float = EPS_SCORE):
    """Token order of ONE image: ``keys_*`` (k,) integer token ids in selection order, ``scores_ref_sorted`` the
    reference's scores in ITS order (descending, length >= k: include the tokens just past a top-k cut).  Returns
    (ok, n_moved): ok when every token we placed at rank i has a reference score within eps of the reference's
    rank-i score, i.e. permutations stay inside runs of near-tied scores and a cut only swaps near-tied tokens."""
    k = len(keys_ours)
    ref_rank = {int(t): i for i, t in enumerate(keys_ref)}
    moved = 0
    for i, t in enumerate(keys_ours.tolist()):
        j = ref_rank.get(int(t))
        if j == i:
            continue
        moved += 1
        if j is None:                 # not in the reference's top-k at all: must tie with the k-th score
            return False, moved
        if abs(float(scores_ref_sorted[j]) - float(scores_ref_sorted[i])) >= eps:
            return False, moved
    return True, moved
