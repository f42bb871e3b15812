"""TEST INFRASTRUCTURE -- CPU restatement (numpy, float64 like MATLAB) of the reference's matching + RANSAC step.
Only tests/ may import this.  PARITY UNPINNED: the arithmetic lives in MATLAB (pdist2, svd, randsample), absent here; the
reference holds no test or golden vector for it.  Every function cites the script it follows.

    scripts/computeAndVisualizeMatches.m:43-52     match_descriptors, the call into ransacfitRt
    scripts/external/estimateRigidTransform.m      estimate_rigid_transform
    scripts/external/quat2rot.m                    quat2rot
    scripts/external/ransacfitRt.m, ransac.m       ransacfit_rt (sample triples are an input: MATLAB's stream is not reproducible)
"""
import numpy as np


def match_descriptors(desc1, desc2):
    """computeAndVisualizeMatches.m:43: [~, m] = pdist2(desc2, desc1, 'euclidean', 'smallest', 1) -> for every row of desc1 the
    index of the nearest row of desc2 (first index on ties)."""
    d1, d2 = np.asarray(desc1, np.float64), np.asarray(desc2, np.float64)
    dist2 = ((d1[:, None, :] - d2[None, :, :]) ** 2).sum(2)
    return dist2.argmin(1), dist2.min(1)


def quat2rot(q):
    """quat2rot.m"""
    q0, q1, q2, q3 = q
    return np.array([[q0 * q0 + q1 * q1 - q2 * q2 - q3 * q3, 2 * (q1 * q2 - q0 * q3), 2 * (q1 * q3 + q0 * q2)],
                     [2 * (q1 * q2 + q0 * q3), q0 * q0 - q1 * q1 + q2 * q2 - q3 * q3, 2 * (q2 * q3 - q0 * q1)],
                     [2 * (q1 * q3 - q0 * q2), 2 * (q2 * q3 + q0 * q1), q0 * q0 - q1 * q1 - q2 * q2 + q3 * q3]])


def estimate_rigid_transform(x, y):
    """estimateRigidTransform.m:44-86.  x, y: (3,N) with x ~ R y + t.  Returns (T (4,4), Eps)."""
    x, y = np.asarray(x, np.float64), np.asarray(y, np.float64)
    n = x.shape[1]
    xc, yc = x.sum(1) / n, y.sum(1) / n
    xz, yz = x - xc[:, None], y - yc[:, None]
    B = np.zeros((4, 4))
    for i in range(n):
        a = yz[:, i] - xz[:, i]          # R12 (row); R21 = -a (column)
        s = yz[:, i] + xz[:, i]          # R22_1 -> crossTimesMatrix
        K = np.array([[0, -s[2], s[1]], [s[2], 0, -s[0]], [-s[1], s[0], 0]])
        A = np.zeros((4, 4))
        A[0, 1:] = a
        A[1:, 0] = -a
        A[1:, 1:] = K
        B += A.T @ A
    _, S, Vt = np.linalg.svd(B)
    rot = quat2rot(Vt[3])                # V(:,4)
    T = np.eye(4)
    T[:3, :3] = rot
    T[:3, 3] = xc - rot @ yc             # T3 * T2 * T1
    return T, S[3]


def inliers_of(Rt, pts1, pts2, t):
    """ransacfitRt.m:79-99 (euc3Ddist): |x1 - (R x2 + t)| < t"""
    d = np.sqrt((((np.asarray(pts1, np.float64).T - (Rt[:, :3] @ np.asarray(pts2, np.float64).T + Rt[:, 3:4])) ** 2).sum(0)))
    return np.flatnonzero(np.abs(d) < t)


def ransacfit_rt(pts1, pts2, t, triples, max_trials=10000):
    """ransacfitRt.m:42-75 over ransac.m:103-214 with the trial samples given.  pts1, pts2: (N,3), pts1 ~ R pts2 + t.
    Returns (Rt (3,4) refit on the inliers or None, inlier indices, trialcount, chosen trial)."""
    pts1, pts2 = np.asarray(pts1, np.float64), np.asarray(pts2, np.float64)
    npts = pts1.shape[0]
    if npts < 3:
        return None, np.zeros(0, int), 0, -1
    if npts == 3:
        return estimate_rigid_transform(pts1.T, pts2.T)[0][:3], np.arange(3), 0, -1
    p, eps = 0.99, np.finfo(float).eps
    N, trialcount, bestscore, best, best_inl = 1.0, 0, 0, -1, np.zeros(0, int)
    while N > trialcount:
        ind = np.asarray(triples[trialcount])
        M = estimate_rigid_transform(pts1[ind].T, pts2[ind].T)[0][:3]
        inl = inliers_of(M, pts1, pts2, t)
        if len(inl) >= bestscore:                      # ransac.m:170
            bestscore, best, best_inl = len(inl), trialcount, inl
            frac = len(inl) / npts
            pno = min(1 - eps, max(eps, 1 - frac ** 3))
            N = max(np.log(1 - p) / np.log(pno), 10)
        trialcount += 1
        if trialcount > max_trials:
            break
    if len(best_inl) >= 3:
        return estimate_rigid_transform(pts1[best_inl].T, pts2[best_inl].T)[0][:3], best_inl, trialcount, best
    return None, np.zeros(0, int), trialcount, best
