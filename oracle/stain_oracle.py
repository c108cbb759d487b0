"""CPU (numpy, float64) restatement of the reference stain-normalisation path.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``): used by tests/, smoke() and
bench.py's cpu_baseline / ``--impl reference`` legs as the checker / timed CPU baseline.

Parity status: **pinned**.  Every function below is checked in
``tests/test_oracle_stain.py`` against (i) the reference's own golden vector
(``/root/reference/tests/test_utils.py:882-911``), (ii) the unmodified reference sources
run through ``oracle/ref_shim.py`` (when /root/reference is present), and (iii) the
fixtures under ``tests/golden/`` that ``oracle/make_golden.py`` generated from the
reference.

Each function cites the reference lines it follows (paths relative to
``/root/reference/tiatoolbox``).
"""
from __future__ import annotations

import numpy as np

RUIFROK = np.array([[0.65, 0.70, 0.29], [0.07, 0.99, 0.11]])  # tools/stainextract.py:127

# ---------------------------------------------------------------------------------------
# cv2.cvtColor(COLOR_RGB2LAB) 8-bit L channel — integer restatement (OpenCV 4.13 fixed-point
# path; exhaustively verified against cv2 over all 2^24 colours in the test-suite).
# utils/misc.py:281-283
# ---------------------------------------------------------------------------------------


def _srgb_gamma_tab() -> np.ndarray:
    x = np.arange(256, dtype=np.float64) / 255.0
    g = np.where(x <= 0.04045, x / 12.92, ((x + 0.055) / 1.055) ** 2.4)
    return np.rint(2040.0 * g).astype(np.int64)


def _lab_cbrt_tab() -> np.ndarray:
    t = np.arange(2041, dtype=np.float64) / 2040.0
    f = np.where(t < 216.0 / 24389.0, t * 841.0 / 108.0 + 16.0 / 116.0, np.cbrt(t))
    return np.rint(32768.0 * f).astype(np.int64)


TABG = _srgb_gamma_tab()
TABC = _lab_cbrt_tab()
Y_COEF = (871, 2929, 296)


def lab_y(img_u8: np.ndarray) -> np.ndarray:
    """Fixed-point luminance Y (0..2040) that cv2's 8-bit RGB2LAB derives L from."""
    r, g, b = (TABG[img_u8[..., i]] for i in range(3))
    return (Y_COEF[0] * r + Y_COEF[1] * g + Y_COEF[2] * b + 2048) >> 12


def lab_l(img_u8: np.ndarray) -> np.ndarray:
    """uint8 L channel of cv2.cvtColor(img, COLOR_RGB2LAB)."""
    y = lab_y(img_u8)
    return np.clip((296 * TABC[y] - 1336934 + 16384) >> 15, 0, 255).astype(np.uint8)


def y_threshold(threshold: float) -> int:
    """Largest Y with ``L(Y)/255.0 < threshold`` (-1 if none): tissue <=> Y <= y_threshold.

    L is monotone non-decreasing in Y, so the luminosity test of utils/misc.py:282-283
    collapses to one integer compare.  For the default 0.8 this is 1146.
    """
    y = np.arange(2041)
    l_of_y = np.clip((296 * TABC[y] - 1336934 + 16384) >> 15, 0, 255)
    ok = (l_of_y / 255.0) < threshold
    return int(y[ok].max()) if ok.any() else -1


# ---------------------------------------------------------------------------------------
# utils/misc.py:405-444 contrast_enhancer ; 261-290 get_luminosity_tissue_mask
# ---------------------------------------------------------------------------------------


def contrast_enhancer(img: np.ndarray, low_p: float = 2, high_p: float = 98) -> np.ndarray:
    if img.dtype != np.uint8:
        raise AssertionError("Image should be uint8.")
    p_low, p_high = np.percentile(img, (low_p, high_p))
    if p_low >= p_high:
        p_low, p_high = float(img.min()), float(img.max())
    if p_high > p_low:
        # skimage.exposure.rescale_intensity(in_range=(lo,hi), out_range=(0.,255.)) — float64
        x = np.clip(img.astype(np.float64), float(p_low), float(p_high))
        x = (x - float(p_low)) / (float(p_high) - float(p_low))
        return (x * 255.0 + 0.0).astype(np.uint8)
    return img.copy()


def get_luminosity_tissue_mask(img: np.ndarray, threshold: float) -> np.ndarray:
    img = img.astype("uint8")
    enh = contrast_enhancer(img, 2, 98)
    mask = (lab_l(enh) / 255.0) < threshold
    if mask.sum() == 0:
        raise ValueError("Empty tissue mask computed.")
    return mask


# ---------------------------------------------------------------------------------------
# utils/transforms.py:209-231 rgb2od (mutates its argument: zeros -> ones) ; 234-256 od2rgb
# ---------------------------------------------------------------------------------------


def rgb2od(img: np.ndarray) -> np.ndarray:
    img[img == 0] = 1
    return np.maximum(-1 * np.log(img / 255), 1e-6)


def od2rgb(od: np.ndarray) -> np.ndarray:
    od = np.maximum(od, 1e-6)
    return (255 * np.exp(-1 * od)).astype(np.uint8)


# ---------------------------------------------------------------------------------------
# tools/stainextract.py:177-227 MacenkoExtractor.get_stain_matrix
# ---------------------------------------------------------------------------------------


def macenko_stain_matrix(img, luminosity_threshold=0.8, angular_percentile=99, *, debug=None):
    img = img.astype("uint8")
    mask = get_luminosity_tissue_mask(img, luminosity_threshold).reshape(-1)
    od = rgb2od(img).reshape(-1, 3)[mask]
    _, vec = np.linalg.eigh(np.cov(od, rowvar=False))
    vec = vec[:, [2, 1]]
    if vec[0, 0] < 0:
        vec[:, 0] *= -1
    if vec[0, 1] < 0:
        vec[:, 1] *= -1
    proj = od @ vec
    phi = np.arctan2(proj[:, 1], proj[:, 0])
    min_phi = np.percentile(phi, 100 - angular_percentile)
    max_phi = np.percentile(phi, angular_percentile)
    v1 = vec @ np.array([np.cos(min_phi), np.sin(min_phi)])
    v2 = vec @ np.array([np.cos(max_phi), np.sin(max_phi)])
    he = np.array([v1, v2]) if v1[0] > v2[0] else np.array([v2, v1])
    if debug is not None:
        debug.update(mask=mask, eigvec=vec, min_phi=min_phi, max_phi=max_phi)
    return he / np.linalg.norm(he, axis=1)[:, None]


# ---------------------------------------------------------------------------------------
# tools/stainextract.py:281-322 VahadaneExtractor.get_stain_matrix — restates the
# scikit-learn (>=1.2; 1.9.0 installed) DictionaryLearning.fit_transform path
# (sklearn/decomposition/_dict_learning.py: _dict_learning, sparse_encode(lasso_lars),
# _update_dict) for n_components=2 on X = OD[tissue].T (3 x N).
# ---------------------------------------------------------------------------------------


_EPS32 = float(np.finfo(np.float32).eps)
_TINY32 = float(np.finfo(np.float32).tiny)


def _min_pos(vals) -> float:
    vals = [v for v in vals if v > 0]
    return min(vals) if vals else np.finfo(np.float64).max


def lasso2(gram: np.ndarray, c: np.ndarray, reg: float, n: int) -> np.ndarray:
    """2-variable LARS-lasso path exactly as scikit-learn's ``_lars_path_solver`` walks it.

    Follows sklearn/linear_model/_least_angle.py (method="lasso", precomputed Gram, Xy=c,
    alpha_min = reg / n, n_samples = n), including its early-stopping rule with the
    float32-eps ``equality_tolerance`` (for n > reg/eps32 ~ 8.4e5 tissue pixels the result
    degenerates to the unregularised least-squares fit) and the lasso drop step.
    """
    alpha_min = reg / n
    coef = np.zeros(2)
    prev_coef = np.zeros(2)
    prev_alpha = 0.0
    active: list[int] = []
    sign = np.zeros(2)
    n_iter = 0
    drop = False
    for _ in range(16):
        corr = c - gram @ coef
        inactive = [j for j in (0, 1) if j not in active]
        if inactive:
            j_star = max(inactive, key=lambda j: abs(corr[j]))
            c_max = abs(corr[j_star])
        else:
            j_star, c_max = -1, 0.0
        alpha = c_max / n
        if alpha <= alpha_min + _EPS32:
            if abs(alpha - alpha_min) > _EPS32 and n_iter > 0:
                ss = (prev_alpha - alpha_min) / (prev_alpha - alpha)
                coef = prev_coef + ss * (coef - prev_coef)
            break
        if len(active) >= 2:
            break
        if not drop:
            sign[j_star] = np.sign(corr[j_star])
            active.append(j_star)
        # equiangular direction of the active set
        if len(active) == 1:
            a = active[0]
            ls = np.zeros(2)
            ls[a] = sign[a] / gram[a, a]
            aa = 1.0 / np.sqrt(ls[a] * sign[a])
        else:
            det = gram[0, 0] * gram[1, 1] - gram[0, 1] * gram[1, 0]
            ls = np.array(
                [
                    (gram[1, 1] * sign[0] - gram[0, 1] * sign[1]) / det,
                    (gram[0, 0] * sign[1] - gram[1, 0] * sign[0]) / det,
                ]
            )
            aa = 1.0 / np.sqrt(ls[0] * sign[0] + ls[1] * sign[1])
        ls = ls * aa
        inactive = [j for j in (0, 1) if j not in active]
        ced = {j: float(np.round(gram[j] @ ls, 15)) for j in inactive}
        g1 = _min_pos([(c_max - corr[j]) / (aa - ced[j] + _TINY32) for j in inactive])
        g2 = _min_pos([(c_max + corr[j]) / (aa + ced[j] + _TINY32) for j in inactive])
        gamma = min(g1, g2, c_max / aa)
        drop = False
        z = {a: -coef[a] / (ls[a] + _TINY32) for a in active}
        z_pos = _min_pos(z.values())
        drop_idx = []
        if z_pos < gamma:
            drop_idx = [a for a in active if z[a] == z_pos]
            for a in drop_idx:
                sign[a] = -sign[a]
            gamma = z_pos
            drop = True
        n_iter += 1
        prev_coef, prev_alpha = coef, alpha
        coef = np.zeros(2)  # inactive (incl. previously dropped) coefficients are exactly 0
        for a in active:
            coef[a] = prev_coef[a] + gamma * ls[a]
        for a in drop_idx:
            active.remove(a)
            sign[a] = 0.0
    return coef


def vahadane_stain_matrix(img, luminosity_threshold=0.8, regularizer=0.1, *, max_iter=3, tol=1e-8):
    img = img.astype("uint8")
    mask = get_luminosity_tissue_mask(img, luminosity_threshold).reshape(-1)
    od = rgb2od(img).reshape(-1, 3)[mask]
    x = od.T  # 3 x N  (3 "samples", N "features")
    # SVD init through the 3x3 Gram (== scipy.linalg.svd(X) up to round-off) + svd_flip (u-based)
    evals, evecs = np.linalg.eigh(x @ x.T)
    order = np.argsort(evals)[::-1]
    u = evecs[:, order]
    for k in range(3):
        if u[np.argmax(np.abs(u[:, k])), k] < 0:
            u[:, k] *= -1
    code = u[:, :2].copy()  # 3 x 2
    d = code.T @ x  # 2 x N  == (S * Vt)[:2]
    errors = []
    for ii in range(max_iter):
        gram = d @ d.T
        cov = d @ x.T  # 2 x 3
        code = np.stack([lasso2(gram, cov[:, k], regularizer, x.shape[1]) for k in range(3)])  # 3 x 2
        a = code.T @ code
        b = x.T @ code  # N x 2
        for k in range(2):
            if a[k, k] > 1e-6:
                d[k] += (b[:, k] - a[k] @ d) / a[k, k]
            else:  # reference re-samples a random atom here (random_state=None): unspecified
                raise FloatingPointError("vahadane: unused atom (reference behaviour is random)")
            np.clip(d[k], 0, None, out=d[k])
            d[k] /= max(np.linalg.norm(d[k]), 1)
        cost = 0.5 * np.sum((x - code @ d) ** 2) + regularizer * np.sum(np.abs(code))
        errors.append(cost)
        if ii > 0 and errors[-2] - errors[-1] < tol * errors[-1]:
            break
    dictionary = code.T  # the reference uses the *code* (fit_transform output) transposed
    if dictionary[0, 0] < dictionary[1, 0]:
        dictionary = dictionary[[1, 0], :]
    return dictionary / np.linalg.norm(dictionary, axis=1)[:, None]


# ---------------------------------------------------------------------------------------
# tools/stainnorm.py:49-113 StainNormalizer.get_concentrations / fit / transform
# ---------------------------------------------------------------------------------------


def get_concentrations(img: np.ndarray, stain_matrix: np.ndarray) -> np.ndarray:
    od = rgb2od(img).reshape(-1, 3)
    # np.linalg.lstsq(M.T, od.T, rcond=-1)[0].T  ==  od @ pinv(M)  for a full-rank M
    return od @ np.linalg.pinv(stain_matrix)


def extract(method: str, img: np.ndarray, stain_matrix=None) -> np.ndarray:
    method = method.lower()
    if method == "macenko":
        return macenko_stain_matrix(img)
    if method == "vahadane":
        return vahadane_stain_matrix(img)
    if method == "ruifrok":
        return RUIFROK.copy()
    if method == "custom":
        return np.asarray(stain_matrix, dtype=np.float64)
    raise ValueError(method)


def fit(method: str, target: np.ndarray, stain_matrix=None):
    m_t = extract(method, target, stain_matrix)
    conc = get_concentrations(target, m_t)
    max_c = np.percentile(conc, 99, axis=0).reshape((1, 2))
    return m_t, max_c


def transform_float(method, img, m_target, max_c_target, stain_matrix=None, *, debug=None):
    """Pre-quantisation float64 result of StainNormalizer.transform (stainnorm.py:100-111)."""
    m_s = extract(method, img, stain_matrix)
    conc = get_concentrations(img, m_s)
    max_c_s = np.percentile(conc, 99, axis=0).reshape((1, 2))
    conc = conc * (max_c_target / max_c_s)
    trans = 255 * np.exp(-1 * np.dot(conc, m_target))
    trans[trans > 255] = 255
    trans[trans < 0] = 0
    if debug is not None:
        debug.update(stain_matrix_source=m_s, max_c_source=max_c_s)
    return trans.reshape(img.shape)


def transform(method, img, m_target, max_c_target, stain_matrix=None):
    return transform_float(method, img, m_target, max_c_target, stain_matrix).astype(np.uint8)


# ---------------------------------------------------------------------------------------
# tools/stainaugment.py:143-206 StainAugmentor.fit / augment with the random draws made explicit
# ---------------------------------------------------------------------------------------


def augment(img: np.ndarray, stain_matrix: np.ndarray, alphas, betas, *, threshold: float = 0.85, augment_background: bool = False) -> np.ndarray:
    """`fit(img, threshold)` followed by `augment()` for given per-stain (alpha_i, beta_i) (the reference draws them in `get_params`,
    stainaugment.py:231-236, once per stain inside the loop of :189-196)."""
    conc = get_concentrations(np.array(img, copy=True), stain_matrix).copy()
    if augment_background:
        for i in range(conc.shape[1]):
            conc[:, i] *= alphas[i]
            conc[:, i] += betas[i]
    else:
        mask = get_luminosity_tissue_mask(img, threshold).ravel()
        for i in range(conc.shape[1]):
            conc[mask, i] *= alphas[i]
            conc[mask, i] += betas[i]
    out = 255 * np.exp(-1 * np.dot(conc, stain_matrix))
    return np.clip(out.reshape(img.shape), 0, 255).astype(np.uint8)
