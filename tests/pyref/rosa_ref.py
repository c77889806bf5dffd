"""Independent numpy/scipy restatement of Rosa::preprocess .. similarity_neighbor_extraction + surf_nbs -- TEST INFRASTRUCTURE.

Written from /root/reference/osep_skeleton_decomp/src/rosa.cpp:57-267,625-651 and include/rosa.hpp:99-128 plus the
third-party semantics listed in SURVEY.md Appendix A (PCL 1.12 VoxelGrid / NormalEstimation, FLANN L2_Simple, Eigen 3.4
SelfAdjointEigenSolver) -- NOT from oracle/ and NOT from csrc/.  Vectorised numpy in float32 (every operation is one
IEEE binary32 rounding, like the reference's -O3 build without FMA contraction); scipy's cKDTree only proposes
candidate sets, membership and order are decided here in float32 with the explicit (d2, index) tie-break.

It restates the DISCRETE stages the rest of the pipeline hangs on -- filter indices, kNN(20) sets, normals, the leaf
control law, voxel keys / counts / order / centroids, kNN(10) `surf_nbs`, radius candidates and `simi_nbs` -- so that
the C++ oracle is compared with a second, separately written implementation (tests/test_pyref_cpu.py).

Canonical choices shared with the oracle (where the reference itself is unspecified):
neighbour order (d2, index); normals by Eigen's iterative tridiagonal QR instead of PCL's trigonometric eigen33;
stable VoxelGrid sort; correctly rounded cube root.
"""
import numpy as np
from scipy.spatial import cKDTree

F = np.float32
FLT_MIN = F(np.finfo(np.float32).tiny)
FLT_EPS = F(np.finfo(np.float32).eps)


def d2_f32(a, b):
    """FLANN L2_Simple on float32 rows: ((dx*dx + dy*dy) + dz*dz)."""
    d = a - b
    return (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2]


# ------------------------------------------------------------------------------------------------------------------
# Eigen::SelfAdjointEigenSolver<Matrix3f>::compute (iterative path), vectorised over n matrices.
# Input: lower triangle a00,a10,a11,a20,a21,a22.  Output: eigenvalues ascending (n,3), eigenvectors (n,3,3) [row, col].
def _hypot(x, y):
    x, y = np.abs(x), np.abs(y)
    p = np.where(x < y, y, x)
    q = np.where(x < y, x, y)
    with np.errstate(all="ignore"):
        qp = q / p
        r = p * np.sqrt(F(1.0) + qp * qp)
    return np.where(p == 0, F(0.0), r)


def _givens(p, q):
    """JacobiRotation::makeGivens (real): returns c, s with the reference's case analysis."""
    with np.errstate(all="ignore"):
        big = np.abs(p) > np.abs(q)
        # |p| > |q|
        t1 = q / p
        u1 = np.sqrt(F(1.0) + t1 * t1)
        u1 = np.where(p < 0, -u1, u1)
        c1 = F(1.0) / u1
        s1 = -t1 * c1
        # else
        t2 = p / q
        u2 = np.sqrt(F(1.0) + t2 * t2)
        u2 = np.where(q < 0, -u2, u2)
        s2 = -F(1.0) / u2
        c2 = -t2 * s2
    c = np.where(big, c1, c2)
    s = np.where(big, s1, s2)
    pz = p == 0
    c = np.where(pz, F(0.0), c)
    s = np.where(pz, np.where(q < 0, F(1.0), F(-1.0)), s)
    qz = q == 0
    c = np.where(qz, np.where(p < 0, F(-1.0), F(1.0)), c)
    s = np.where(qz, F(0.0), s)
    return c.astype(F), s.astype(F)


def eig3_sym_batch(a00, a10, a11, a20, a21, a22):
    a = [np.asarray(x, dtype=F).copy() for x in (a00, a10, a11, a20, a21, a22)]
    n = a[0].shape[0]
    with np.errstate(all="ignore"):
        scale = np.abs(a[0])
        for x in (a[1], a[3], a[2], a[4], a[5]):      # coefficient-wise maxCoeff over the lower triangle (column-major order)
            ax = np.abs(x)
            scale = np.where(scale < ax, ax, scale)
        scale = np.where(scale == 0, F(1.0), scale)
        a00, a10, a11, a20, a21, a22 = [x / scale for x in a]
        # tridiagonalization_inplace_selector<MatrixType, 3, false>
        d = np.zeros((n, 3), F)
        e = np.zeros((n, 2), F)
        Q = np.zeros((n, 3, 3), F)
        Q[:, 0, 0] = 1
        d[:, 0] = a00
        v1n2 = a20 * a20
        small = v1n2 <= FLT_MIN
        beta = np.sqrt(a10 * a10 + v1n2)
        invb = F(1.0) / beta
        m01 = a10 * invb
        m02 = a20 * invb
        q = F(2.0) * m01 * a21 + m02 * (a22 - a11)
        d[:, 1] = np.where(small, a11, a11 + m02 * q)
        d[:, 2] = np.where(small, a22, a22 - m02 * q)
        e[:, 0] = np.where(small, a10, beta)
        e[:, 1] = np.where(small, a21, a21 - m01 * q)
        Q[:, 1, 1] = np.where(small, F(1.0), m01)
        Q[:, 1, 2] = np.where(small, F(0.0), m02)
        Q[:, 2, 1] = np.where(small, F(0.0), m02)
        Q[:, 2, 2] = np.where(small, F(1.0), -m01)
        # computeFromTridiagonal_impl
        end = np.full(n, 2, np.int64)
        start = np.zeros(n, np.int64)
        iters = np.zeros(n, np.int64)
        pinv = F(1.0) / FLT_EPS
        live = np.ones(n, bool)
        for _ in range(91):
            for i in (0, 1):
                inwin = live & (start <= i) & (i < end)
                tiny = np.abs(e[:, i]) < FLT_MIN
                sc = pinv * e[:, i]
                neg = (sc * sc) <= (np.abs(d[:, i]) + np.abs(d[:, i + 1]))
                e[:, i] = np.where(inwin & (tiny | neg), F(0.0), e[:, i])
            # while (end > 0 && subdiag[end-1] == 0) end--
            end = np.where(live & (end == 2) & (e[:, 1] == 0), 1, end)
            end = np.where(live & (end == 1) & (e[:, 0] == 0), 0, end)
            live = live & (end > 0)
            if not live.any():
                break
            iters = iters + live
            start = np.where(live, end - 1, start)
            start = np.where(live & (start == 1) & (e[:, 0] != 0), 0, start)
            # tridiagonal_qr_step
            e2 = end == 2
            dem1 = np.where(e2, d[:, 1], d[:, 0])
            de = np.where(e2, d[:, 2], d[:, 1])
            ee = np.where(e2, e[:, 1], e[:, 0])
            td = (dem1 - de) * F(0.5)
            esq = ee * ee
            h = _hypot(td, ee)
            sh = np.where(td > 0, h, -h)
            mu_a = de - np.abs(ee)
            mu_b = de - ee / ((td + sh) / ee)
            mu_c = de - esq / (td + sh)
            mu = np.where(td == 0, mu_a, np.where(ee == 0, de, np.where(esq == 0, mu_b, mu_c)))
            s0 = start == 0
            x = np.where(s0, d[:, 0], d[:, 1]) - mu
            z = np.where(s0, e[:, 0], e[:, 1])
            for k in (0, 1):
                act = live & (start <= k) & (k < end) & (z != 0)
                if not act.any():
                    continue
                c, s = _givens(x, z)
                dk, dk1, ek = d[:, k], d[:, k + 1], e[:, k]
                sdk = s * dk + c * ek
                dkp1 = s * ek + c * dk1
                ndk = c * (c * dk - s * ek) - s * (c * ek - s * dk1)
                ndk1 = s * sdk + c * dkp1
                nek = c * sdk - s * dkp1
                d[:, k] = np.where(act, ndk, dk)
                d[:, k + 1] = np.where(act, ndk1, dk1)
                e[:, k] = np.where(act, nek, ek)
                if k == 1:
                    notfirst = act & (start < 1)
                    e[:, 0] = np.where(notfirst, c * e[:, 0] - s * z, e[:, 0])
                x = np.where(act, e[:, k], x)
                if k == 0:
                    notlast = act & (end > 1)
                    z = np.where(notlast, -s * e[:, 1], z)
                    e[:, 1] = np.where(notlast, c * e[:, 1], e[:, 1])
                qk, qk1 = Q[:, :, k].copy(), Q[:, :, k + 1].copy()
                Q[:, :, k] = np.where(act[:, None], c[:, None] * qk - s[:, None] * qk1, qk)
                Q[:, :, k + 1] = np.where(act[:, None], s[:, None] * qk + c[:, None] * qk1, qk1)
        # ascending selection sort with column swaps (first minimum wins)
        for i in (0, 1):
            sub = d[:, i:]
            k = np.argmin(sub, axis=1) + i            # argmin returns the first minimum
            sw = k != i
            rows = np.nonzero(sw)[0]
            if rows.size:
                kk = k[rows]
                tmp = d[rows, i].copy(); d[rows, i] = d[rows, kk]; d[rows, kk] = tmp
                tq = Q[rows, :, i].copy(); Q[rows, :, i] = Q[rows, :, kk]; Q[rows, :, kk] = tq
        ev = d * scale[:, None]
    return ev, Q, iters


# ------------------------------------------------------------------------------------------------------------------
def exact_knn(P, k, chunk=20000):
    """kNN sets in the canonical (d2 float32, index) order: cKDTree proposes k + margin candidates (float64), the float32
    distances and the order are recomputed here; the margin is widened until the last candidate is strictly farther (in
    float64, with slack) than the k-th float32 distance, so no float32 tie can be missed."""
    n = P.shape[0]
    k_eff = min(k, n)
    tree = cKDTree(P.astype(np.float64))
    out = np.empty((n, k_eff), np.int32)
    for b in range(0, n, chunk):
        q = P[b:b + chunk]
        kk = min(n, k_eff + 12)
        while True:
            dd, ii = tree.query(q.astype(np.float64), k=kk)
            if kk == 1:
                dd, ii = dd[:, None], ii[:, None]
            d2 = d2_f32(q[:, None, :], P[ii]).astype(F)
            order = np.lexsort((ii, d2), axis=1)
            d2s = np.take_along_axis(d2, order, axis=1)
            iis = np.take_along_axis(ii, order, axis=1)
            kth = d2s[:, k_eff - 1].astype(np.float64)
            ok = (kk == n) | (dd[:, -1] ** 2 > kth * (1 + 1e-5) + 1e-30)
            if ok.all():
                out[b:b + chunk] = iis[:, :k_eff]
                break
            kk = min(n, kk * 2)
    return out


def normals_pcl(P, knn_idx):
    """pcl::NormalEstimation::computeFeature (dense path): computeMeanAndCovarianceMatrix shifted by the first neighbour,
    sums in neighbour order, smallest eigenvector (canonical solver), flipNormalTowardsViewpoint(0,0,0)."""
    n, k = knn_idx.shape
    with np.errstate(all="ignore"):
        K0 = P[knn_idx[:, 0]]
        acc = np.zeros((n, 9), F)
        for u in range(k):
            c = P[knn_idx[:, u]] - K0
            x, y, z = c[:, 0], c[:, 1], c[:, 2]
            acc[:, 0] += x * x; acc[:, 1] += x * y; acc[:, 2] += x * z
            acc[:, 3] += y * y; acc[:, 4] += y * z; acc[:, 5] += z * z
            acc[:, 6] += x; acc[:, 7] += y; acc[:, 8] += z
        acc = acc / F(k)
        c00 = acc[:, 0] - acc[:, 6] * acc[:, 6]; c01 = acc[:, 1] - acc[:, 6] * acc[:, 7]; c02 = acc[:, 2] - acc[:, 6] * acc[:, 8]
        c11 = acc[:, 3] - acc[:, 7] * acc[:, 7]; c12 = acc[:, 4] - acc[:, 7] * acc[:, 8]; c22 = acc[:, 5] - acc[:, 8] * acc[:, 8]
        ev, Q, _ = eig3_sym_batch(c00, c01, c11, c02, c12, c22)
        nv = Q[:, :, 0].copy()
        cos_theta = ((F(0.0) - P[:, 0]) * nv[:, 0] + (F(0.0) - P[:, 1]) * nv[:, 1]) + (F(0.0) - P[:, 2]) * nv[:, 2]
        nv = np.where((cos_theta < 0)[:, None], nv * F(-1.0), nv)
    if k < 3:
        nv[:] = np.nan
    return nv.astype(F)


def cbrt_f32(x):
    """correctly rounded binary32 cube root (canonical choice C5)"""
    return F(np.cbrt(np.float64(F(x))))


def voxel_keys(P, leaf, bmin, bmax):
    """PCL VoxelGrid::applyFilter steps 1-5 (float inverse leaf, float floor, float subtraction of the int min box)."""
    with np.errstate(all="ignore"):
        inv = F(1.0) / F(leaf)
        ext = [int(np.int64(F((bmax[d] - bmin[d]) * inv))) + 1 for d in range(3)]
        overflow = ext[0] * ext[1] * ext[2] > 2147483647
        minb = [int(np.floor(F(bmin[d] * inv))) for d in range(3)]
        maxb = [int(np.floor(F(bmax[d] * inv))) for d in range(3)]
        divb = [maxb[d] - minb[d] + 1 for d in range(3)]
        ijk = [(np.floor(P[:, d] * inv) - F(minb[d])).astype(np.int32) for d in range(3)]
        key = (ijk[0].astype(np.int64) + ijk[1].astype(np.int64) * divb[0] + ijk[2].astype(np.int64) * (divb[0] * divb[1]))
    return (key & 0xffffffff).astype(np.uint32), overflow


class RosaRef:
    def __init__(self, max_points=1000, min_points=100, pts_dist_lim=50.0, ne_knn=20, nb_knn=10):
        self.max_points, self.min_points, self.pts_dist_lim = int(max_points), int(min_points), F(pts_dist_lim)
        self.ne_knn, self.nb_knn = int(ne_knn), int(nb_knn)
        self.t = {}

    def run(self, xyz):
        xyz = np.ascontiguousarray(xyz, dtype=F)[:, :3]
        t = self.t = {}
        # ---- rosa.cpp:67-88
        with np.errstate(all="ignore"):
            x, y, z = xyz[:, 0], xyz[:, 1], xyz[:, 2]
            d2 = x * x + y * y + z * z
            keep = np.isfinite(x) & np.isfinite(y) & np.isfinite(z) & (d2 <= F(self.pts_dist_lim * self.pts_dist_lim))
        t["filt_idx"] = np.nonzero(keep)[0].astype(np.int32)
        P = xyz[keep]
        t["filtered"] = P
        if P.shape[0] < self.min_points:                                   # rosa.cpp:91-95
            return False
        # ---- rosa.cpp:97-101
        t["knn_full"] = exact_knn(P, self.ne_knn)
        t["normals_full"] = normals_pcl(P, t["knn_full"])
        # ---- rosa.hpp:99-110, rosa.cpp:118-148
        bmin, bmax = P.min(axis=0), P.max(axis=0)
        ext = [max(F(1e-6), F(bmax[d] - bmin[d])) for d in range(3)]
        vol = F(F(ext[0] * ext[1]) * ext[2])
        leaf = cbrt_f32(F(vol / max(F(1.0), F(self.max_points))))
        leaf_hist, nvox_hist = [], []
        tol, leaf_min, leaf_max = F(0.05), F(1e-4), F(1e4)
        keys = None
        for it in range(8):
            keys, overflow = voxel_keys(P, leaf, bmin, bmax)
            leaf_hist.append(F(leaf))
            leaf_used = F(leaf)
            nvox = P.shape[0] if overflow else int(np.unique(keys).size)
            nvox_hist.append(nvox)
            if nvox == 0:
                leaf = F(leaf * F(0.5))
                if leaf < leaf_min:
                    leaf = leaf_min
                    break
                continue
            ratio = F(F(nvox) / F(self.max_points))
            if abs(F(ratio - F(1.0))) <= tol:
                break
            leaf = F(leaf * cbrt_f32(ratio))
            if not np.isfinite(leaf):
                leaf = F(0.05)
            leaf = min(max(leaf, leaf_min), leaf_max)
        t["leaf_hist"] = np.array(leaf_hist, F); t["nvox_hist"] = np.array(nvox_hist, np.int32)
        self.leaf_size, self.leaf_used = F(leaf), leaf_used
        # ---- final VoxelGrid pass (PCL applyFilter steps 6-8), stable order inside a voxel
        order = np.lexsort((np.arange(P.shape[0]), keys))
        ks = keys[order]
        starts = np.nonzero(np.r_[True, ks[1:] != ks[:-1]])[0]
        ends = np.r_[starts[1:], ks.size]
        t["vox_keys"] = ks[starts]; t["vox_count"] = (ends - starts).astype(np.int32)
        N = t["normals_full"]
        M = starts.size
        pts = np.zeros((M, 3), F); nr = np.zeros((M, 3), F)
        with np.errstate(all="ignore"):
            for v in range(M):
                idx = order[starts[v]:ends[v]]
                z3 = np.zeros((1, 3), F)                                   # the accumulators start from +0 (Vector3f / Vector4f ::Zero())
                sp = np.cumsum(np.concatenate([z3, P[idx]]), axis=0, dtype=F)[-1]                # sequential float sums in ascending point index
                sn = np.cumsum(np.concatenate([z3, N[idx]]), axis=0, dtype=F)[-1]
                pts[v] = sp / F(idx.size)
                zz = (sn[0] * sn[0] + sn[2] * sn[2]) + (sn[1] * sn[1] + F(0.0))   # Vector4f::normalized(): packet reduction (a0+a2)+(a1+a3)
                if zz > 0:
                    sn = sn / np.sqrt(zz)
                # rosa.cpp:156-165
                L = np.sqrt(sn[0] * sn[0] + (sn[1] * sn[1] + sn[2] * sn[2]))
                nr[v] = sn / L if L > F(1e-6) else F(0.0)
        t["pts"], t["nrms"] = pts, nr
        # ---- rosa.cpp:259-267
        t["surf_idx"] = exact_knn(pts, self.nb_knn).reshape(-1)
        # ---- rosa.cpp:201-256, 625-651
        t["simi_off"], t["simi_idx"], t["radius_cand"] = self.similarity(pts, nr)
        return True

    def similarity(self, pts, nr):
        M = pts.shape[0]
        leaf = self.leaf_size
        radius_r = F(F(10.0) * leaf)
        th_sim = F(F(0.1) * leaf)
        r2 = F(radius_r * radius_r)
        off = [0]; idx_all = []; ncand = []
        with np.errstate(all="ignore"):
            sq = nr[:, 0] * nr[:, 0] + (nr[:, 1] * nr[:, 1] + nr[:, 2] * nr[:, 2])
            vn = np.where((sq > 0)[:, None], nr / np.sqrt(sq)[:, None], nr).astype(F)      # Eigen normalize(): only when squaredNorm > 0
            fin = np.isfinite(nr).all(axis=1)
            for i in range(M):
                if not fin[i]:
                    off.append(off[-1]); ncand.append(0)
                    continue
                d2 = d2_f32(pts[i][None, :], pts)
                cand = np.nonzero(d2 < r2)[0]
                cand = cand[np.lexsort((cand, d2[cand]))]
                ncand.append(int(cand.size))
                cand = cand[(cand != i) & fin[cand]]
                p1, v1 = pts[i], vn[i]
                p2, v2 = pts[cand], vn[cand]
                d = p1[None, :] - p2
                dot = v1[0] * v2[:, 0] + (v1[1] * v2[:, 1] + v1[2] * v2[:, 2])
                dot = np.maximum(F(0.0), np.minimum(F(1.0), dot))
                pr1 = d[:, 0] * v1[0] + (d[:, 1] * v1[1] + d[:, 2] * v1[2])
                pr2 = -(d[:, 0] * v2[:, 0] + (d[:, 1] * v2[:, 1] + d[:, 2] * v2[:, 2]))
                a = d + (F(5.0) * pr1)[:, None] * v1[None, :]
                b = d - (F(5.0) * pr2)[:, None] * v2
                na = np.sqrt(a[:, 0] * a[:, 0] + (a[:, 1] * a[:, 1] + a[:, 2] * a[:, 2])) / radius_r
                nb = np.sqrt(b[:, 0] * b[:, 0] + (b[:, 1] * b[:, 1] + b[:, 2] * b[:, 2])) / radius_r

                def bump(tt):
                    t2 = tt * tt
                    return np.where(tt > 1, F(0.0), F(2.0) * (t2 * tt) - F(3.0) * t2 + F(1.0))
                w = np.minimum(bump(na), bump(nb)) * (dot * dot)
                w = np.where(dot == 0, F(0.0), w)
                kept = cand[w > th_sim]
                idx_all.append(kept.astype(np.int32))
                off.append(off[-1] + kept.size)
        return np.array(off, np.int32), (np.concatenate(idx_all) if idx_all else np.zeros(0, np.int32)), np.array(ncand, np.int32)
