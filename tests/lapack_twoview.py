"""Test infrastructure: EpipolarGeometry::reconstruct restated in float64 on top of numpy.linalg.svd
(LAPACK gesdd — an Eigen-JacobiSVD-class singular value decomposition), following the reference
statement by statement (reference src/epipolar_geometry.cc:18-98, 114-283, 285-449, 451-733, 782-950).

It exists to QUANTIFY the distance between the repo's fp32 specification of the two-view path (the
oracle / the CUDA kernels: one-sided Jacobi + Householder null vectors, fixed summation orders) and an
independent implementation that uses a library SVD — the situation of the reference, whose Eigen
JacobiSVD cannot be reproduced bit for bit.  tests/test_twoview_vs_lapack.py runs both
over a few hundred seeded scenes and pins the agreement rates in a golden that is NOT regenerated when
the specification changes.  Nothing in the product imports this file."""
import numpy as np

TH_F, TH_SCORE, TH_H = 3.841, 5.991, 5.991


def _normalize(keys):
    k = np.asarray(keys, dtype=np.float64)
    mean = k.mean(axis=0)
    dev = np.abs(k - mean).mean(axis=0)
    s = 1.0 / dev
    T = np.array([[s[0], 0, -mean[0] * s[0]], [0, s[1], -mean[1] * s[1]], [0, 0, 1.0]])
    return (k - mean) * s, T


def _fit_F_all(pn1, pn2, T1, T2, sets):
    """sets (H, 8) indices into the matched normalised points -> F21 (H, 3, 3)."""
    a, b = pn1[sets], pn2[sets]                      # (H, 8, 2)
    u1, v1, u2, v2 = a[..., 0], a[..., 1], b[..., 0], b[..., 1]
    A = np.stack([u2 * u1, u2 * v1, u2, v2 * u1, v2 * v1, v2, u1, v1, np.ones_like(u1)], axis=-1)  # (H, 8, 9)
    _, _, Vt = np.linalg.svd(A, full_matrices=True)
    Fpre = Vt[:, 8, :].reshape(-1, 3, 3)
    U, w, Vt2 = np.linalg.svd(Fpre)
    w[:, 2] = 0.0
    Fn = (U * w[:, None, :]) @ Vt2
    return T2.T @ Fn @ T1


def _fit_H_all(pn1, pn2, T1, T2, sets):
    a, b = pn1[sets], pn2[sets]
    u1, v1, u2, v2 = a[..., 0], a[..., 1], b[..., 0], b[..., 1]
    z, o = np.zeros_like(u1), np.ones_like(u1)
    r0 = np.stack([z, z, z, -u1, -v1, -o, v2 * u1, v2 * v1, v2], axis=-1)
    r1 = np.stack([u1, v1, o, z, z, z, -u2 * u1, -u2 * v1, -u2], axis=-1)
    A = np.stack([r0, r1], axis=2).reshape(len(sets), 16, 9)
    _, _, Vt = np.linalg.svd(A, full_matrices=True)
    Hn = Vt[:, 8, :].reshape(-1, 3, 3)
    H21 = np.linalg.inv(T2) @ Hn @ T1
    return H21, np.linalg.inv(H21)


def _check_F_all(F, x1, x2, inv_s2):
    """F (H, 3, 3); x1, x2 (N, 2). Returns scores (H,), masks (H, N)."""
    h1 = np.c_[x1, np.ones(len(x1))]
    h2 = np.c_[x2, np.ones(len(x2))]
    l2 = np.einsum('hij,nj->hni', F, h1)             # F x1: line in image 2
    num2 = np.einsum('hni,ni->hn', l2, h2)
    chi1 = num2 ** 2 / (l2[..., 0] ** 2 + l2[..., 1] ** 2) * inv_s2
    l1 = np.einsum('hji,nj->hni', F, h2)             # F^T x2: line in image 1
    num1 = np.einsum('hni,ni->hn', l1, h1)
    chi2 = num1 ** 2 / (l1[..., 0] ** 2 + l1[..., 1] ** 2) * inv_s2
    ok1, ok2 = chi1 <= TH_F, chi2 <= TH_F
    score = np.where(ok1, TH_SCORE - chi1, 0.0).sum(1) + np.where(ok2, TH_SCORE - chi2, 0.0).sum(1)
    return score, ok1 & ok2


def _check_H_all(H21, H12, x1, x2, inv_s2):
    h1 = np.c_[x1, np.ones(len(x1))]
    h2 = np.c_[x2, np.ones(len(x2))]
    p = np.einsum('hij,nj->hni', H12, h2)
    d1 = ((x1[None] - p[..., :2] / p[..., 2:3]) ** 2).sum(-1) * inv_s2
    q = np.einsum('hij,nj->hni', H21, h1)
    d2 = ((x2[None] - q[..., :2] / q[..., 2:3]) ** 2).sum(-1) * inv_s2
    ok1, ok2 = d1 <= TH_H, d2 <= TH_H
    score = np.where(ok1, TH_H - d1, 0.0).sum(1) + np.where(ok2, TH_H - d2, 0.0).sum(1)
    return score, ok1 & ok2


def _best(score):
    """strictly greater wins, earliest on ties, score > 0 required (:153-157)."""
    i = int(np.argmax(score))
    return (i, float(score[i])) if score[i] > 0 else (-1, 0.0)


def _triangulate_all(x1, x2, P1, P2):
    A = np.stack([x1[:, 0:1] * P1[2] - P1[0], x1[:, 1:2] * P1[2] - P1[1],
                  x2[:, 0:1] * P2[2] - P2[0], x2[:, 1:2] * P2[2] - P2[1]], axis=1)   # (n, 4, 4)
    _, _, Vt = np.linalg.svd(A)
    v = Vt[:, 3, :]
    with np.errstate(divide='ignore', invalid='ignore'):
        return v[:, :3] / v[:, 3:4]


def _check_RT(R, t, x1, x2, inl, K, th2):
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    P1 = np.c_[K, np.zeros(3)]
    P2 = K @ np.c_[R, t]
    O2 = -R.T @ t
    idx = np.nonzero(inl)[0]
    p = _triangulate_all(x1[idx], x2[idx], P1, P2)
    fin = np.isfinite(p).all(1)
    with np.errstate(divide='ignore', invalid='ignore'):
        n2 = p - O2
        cosP = (p * n2).sum(1) / (np.linalg.norm(p, axis=1) * np.linalg.norm(n2, axis=1))
        ok = fin & ~((p[:, 2] <= 0) & (cosP < 0.99998))
        p2 = p @ R.T + t
        ok &= ~((p2[:, 2] <= 0) & (cosP < 0.99998))
        e1 = (fx * p[:, 0] / p[:, 2] + cx - x1[idx, 0]) ** 2 + (fy * p[:, 1] / p[:, 2] + cy - x1[idx, 1]) ** 2
        e2 = (fx * p2[:, 0] / p2[:, 2] + cx - x2[idx, 0]) ** 2 + (fy * p2[:, 1] / p2[:, 2] + cy - x2[idx, 1]) ** 2
        ok &= (e1 <= th2) & (e2 <= th2)
    n_good = int(ok.sum())
    if n_good > 0:
        c = np.sort(cosP[ok])
        parallax = np.degrees(np.arccos(min(1.0, c[min(50, len(c) - 1)])))
    else:
        parallax = 0.0
    good = np.zeros(len(x1), dtype=bool)
    good[idx[ok & (cosP < 0.99998)]] = True
    return n_good, parallax, good


def _reconstruct_F(F, x1, x2, inl, K, sigma2):
    N = int(inl.sum())
    E = K.T @ F @ K
    U, _, Vt = np.linalg.svd(E)
    t = U[:, 2] / np.linalg.norm(U[:, 2])
    W = np.array([[0, -1, 0], [1, 0, 0], [0, 0, 1.0]])
    R1 = U @ W @ Vt
    R1 = -R1 if np.linalg.det(R1) < 0 else R1
    R2 = U @ W.T @ Vt
    R2 = -R2 if np.linalg.det(R2) < 0 else R2
    hyps = [(R1, t), (R2, t), (R1, -t), (R2, -t)]
    res = [_check_RT(R, tt, x1, x2, inl, K, 4.0 * sigma2) for R, tt in hyps]
    n_good = [r[0] for r in res]
    mx = max(n_good)
    n_min = max(int(0.9 * N), 50)
    n_sim = sum(1 for g in n_good if g > 0.7 * mx)
    if mx < n_min or n_sim > 1:
        return False, None
    h = n_good.index(mx)
    if res[h][1] > 1.0:
        T = np.eye(4); T[:3, :3] = hyps[h][0]; T[:3, 3] = hyps[h][1]
        return True, T
    return False, None


def _reconstruct_H(H21, x1, x2, inl, K, sigma2):
    N = int(inl.sum())
    A = np.linalg.inv(K) @ H21 @ K
    U, w, Vt = np.linalg.svd(A)
    s = np.linalg.det(U) * np.linalg.det(Vt)
    d1, d2, d3 = w
    if d1 / d2 < 1.00001 or d2 / d3 < 1.00001:
        return False, None
    aux1 = np.sqrt((d1 * d1 - d2 * d2) / (d1 * d1 - d3 * d3))
    aux3 = np.sqrt((d2 * d2 - d3 * d3) / (d1 * d1 - d3 * d3))
    x1s = [aux1, aux1, -aux1, -aux1]
    x3s = [aux3, -aux3, aux3, -aux3]
    st = np.sqrt((d1 * d1 - d2 * d2) * (d2 * d2 - d3 * d3)) / ((d1 + d3) * d2)
    ct = (d2 * d2 + d1 * d3) / ((d1 + d3) * d2)
    hyps = []
    for i, sg in enumerate([st, -st, -st, st]):
        Rp = np.array([[ct, 0, -sg], [0, 1, 0], [sg, 0, ct]])
        R = s * U @ Rp @ Vt
        t = U @ (np.array([x1s[i], 0, -x3s[i]]) * (d1 - d3))
        hyps.append((R, t / np.linalg.norm(t)))
    sp = np.sqrt((d1 * d1 - d2 * d2) * (d2 * d2 - d3 * d3)) / ((d1 - d3) * d2)
    cp = (d1 * d3 - d2 * d2) / ((d1 - d3) * d2)
    for i, sg in enumerate([sp, -sp, -sp, sp]):
        Rp = np.array([[cp, 0, sg], [0, -1, 0], [sg, 0, -cp]])
        R = s * U @ Rp @ Vt
        t = U @ (np.array([x1s[i], 0, x3s[i]]) * (d1 + d3))
        hyps.append((R, t / np.linalg.norm(t)))
    best, second, bi, bpar = 0, 0, -1, -1.0
    for i, (R, t) in enumerate(hyps):
        g, par, _ = _check_RT(R, t, x1, x2, inl, K, 4.0 * sigma2)
        if g > best:
            second, best, bi, bpar = best, g, i, par
        elif g > second:
            second = g
    if second < 0.75 * best and bpar >= 1.0 and best > 50 and best > 0.9 * N:
        T = np.eye(4); T[:3, :3] = hyps[bi][0]; T[:3, 3] = hyps[bi][1]
        return True, T
    return False, None


def reconstruct(tv):
    """Returns dict(ok, used_H, best_F, best_H, SF, SH, mask_F, mask_H, T21) for a scene dict like synth.make_two_view + sets."""
    k1 = np.asarray(tv["keys1"], dtype=np.float64)
    k2 = np.asarray(tv["keys2"], dtype=np.float64)
    m = np.asarray(tv["matches12"])
    i1 = np.nonzero(m >= 0)[0]
    x1, x2 = k1[i1], k2[m[i1]]
    pn1a, T1 = _normalize(k1)
    pn2a, T2 = _normalize(k2)
    pn1, pn2 = pn1a[i1], pn2a[m[i1]]
    sets = np.asarray(tv["sets"])
    sigma = float(tv.get("sigma", 1.0))
    inv_s2 = 1.0 / (sigma * sigma)
    K = np.asarray(tv["K"], dtype=np.float64)
    F = _fit_F_all(pn1, pn2, T1, T2, sets)
    sF, mF = _check_F_all(F, x1, x2, inv_s2)
    H21, H12 = _fit_H_all(pn1, pn2, T1, T2, sets)
    sH, mH = _check_H_all(H21, H12, x1, x2, inv_s2)
    bF, SF = _best(sF)
    bH, SH = _best(sH)
    out = dict(best_F=bF, best_H=bH, SF=SF, SH=SH, mask_F=mF[bF] if bF >= 0 else np.zeros(len(x1), bool),
               mask_H=mH[bH] if bH >= 0 else np.zeros(len(x1), bool), ok=False, used_H=-1, T21=None)
    if SH + SF == 0:
        return out
    if SH / (SH + SF) > 0.5:
        out["used_H"] = 1
        out["ok"], out["T21"] = _reconstruct_H(H21[bH], x1, x2, out["mask_H"], K, sigma * sigma)
    else:
        out["used_H"] = 0
        out["ok"], out["T21"] = _reconstruct_F(F[bF], x1, x2, out["mask_F"], K, sigma * sigma)
    return out
