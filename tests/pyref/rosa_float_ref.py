"""Independent numpy restatement of the FLOAT stages of Rosa -- rosa_init, drosa, dcrosa, vertex_sampling,
vertex_smooth -- TEST INFRASTRUCTURE (continues tests/pyref/rosa_ref.py, which stops at the similarity lists).

Written from /root/reference/osep_skeleton_decomp/src/rosa.cpp:183-199, 258-620, 652-841 and include/rosa.hpp:112-128
plus the Eigen 3.4 evaluation rules below -- NOT from oracle/ and NOT from csrc/.  Every operation is one IEEE binary32
rounding (numpy float32 scalars / arrays; numpy never contracts a*b+c), like the reference's build without FMA.

Eigen rules used (Eigen 3.4 headers, as compiled for x86-64 SSE):
  E1  fixed-size 3-vector reductions (`sum`, `dot`, `squaredNorm`, `norm`, `trace`, the row-times-column sums of a
      coefficient-based 3x3 product) are completely unrolled by halving: a0 + (a1 + a2)     (Redux.h redux_novec_unroller);
  E2  the same reductions over a DYNAMIC-size strided block (`vset.row(p)` of a MatrixXf) run left to right:
      (a0 + a1) + a2                                                                         (Redux.h DefaultTraversal, NoUnrolling);
  E3  `normalize()` / `normalized()` divide by sqrt(squaredNorm) when squaredNorm > 0 (true division, not a reciprocal);
  E4  `dst += s * (a * b^T)` is evaluated as ((s * a) * b^T): element (i, j) = (s * a_i) * b_j
      (ProductEvaluators.h, the "Dense ?= scalar * Product" assignment);
  E5  `SelfAdjointEigenSolver<Matrix3f>` reads the lower triangle only (rosa_ref.eig3_sym_batch);
  E6  `LLT<Matrix3f>`: unblocked Cholesky (column by column, `A21 /= x` true divisions), `solve` = unrolled forward and
      backward substitution with E1 sums;
  E7  `double * Vector3f`: the scalar is narrowed to float first.

Canonical choices shared with the oracle and the CUDA path (the reference leaves them to the kd-tree
layout / the DFS visit order): neighbour lists in ascending (d2, index) order; every float sum over an ACTIVE SET runs in
lane-owner order -- lane l = (i >> 5) & 31 adds its members in ascending index, then the 32 lane sums are combined by the
xor butterfly 16, 8, 4, 2, 1.
"""
import numpy as np

from . import rosa_ref as rr

F = np.float32
I32 = np.arange(32)


def sum3(a0, a1, a2):                      # E1
    return a0 + (a1 + a2)


def dot3(a, b):                            # E1
    return a[..., 0] * b[..., 0] + (a[..., 1] * b[..., 1] + a[..., 2] * b[..., 2])


def d2_flann(a, b):                        # FLANN L2_Simple: ((dx*dx + dy*dy) + dz*dz)
    d = a - b
    return (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2]


def eig1(a00, a10, a11, a20, a21, a22):
    """one 3x3 solve -> (eigenvalues ascending (3,), eigenvectors (3,3) [row, col])"""
    ev, Q, _ = rr.eig3_sym_batch(*[np.array([x], F) for x in (a00, a10, a11, a20, a21, a22)])
    return ev[0], Q[0]


def seq0(X):
    """sequential float32 sum of the rows of X starting from +0 (`acc = Zero(); acc += x` ...): (-0) rows give +0 like the C++ loops"""
    return np.cumsum(np.concatenate([np.zeros((1,) + X.shape[1:], F), X.astype(F)], axis=0), axis=0, dtype=F)[-1]


def lane_sum(idx, X):
    """Canonical sum (C4) of the rows X[k] belonging to the ascending member indices idx[k]: per-lane sequential sums in
    ascending index, then the xor butterfly.  X: (m, c) float32 -> (c,) float32."""
    m, c = X.shape
    lane = (idx >> 5) & 31
    order = np.argsort(lane, kind="stable")          # members of a lane stay in ascending index
    ls = lane[order]
    cnt = np.bincount(ls, minlength=32)
    start = np.r_[0, np.cumsum(cnt)[:-1]]
    rank = np.arange(m) - start[ls]
    Z = np.zeros((32, int(cnt.max()) + 1, c), F)          # slot 0 stays +0: every lane's accumulator starts from zero
    Z[ls, rank + 1] = X[order]
    with np.errstate(all="ignore"):
        acc = np.cumsum(Z, axis=1, dtype=F)[:, -1, :]       # sequential; the zero padding after a lane's last member adds nothing
        for off in (16, 8, 4, 2, 1):
            acc = acc + acc[I32 ^ off]
    return acc[0]


def seq_sum_lists(vals, lens):
    """vals: (n, L, c) zero padded after lens[i] entries; sequential sums in list order -> (n, c)."""
    cs = np.cumsum(np.concatenate([np.zeros_like(vals[:, :1]), vals], axis=1), axis=1, dtype=F)      # accumulators start from +0
    return cs[np.arange(vals.shape[0]), lens]


def orthonormal_frame_x(nv):
    """rosa.hpp:112-128 create_orthonormal_frame(v).col(0)"""
    with np.errstate(all="ignore"):
        n = np.sqrt(sum3(nv[0] * nv[0], nv[1] * nv[1], nv[2] * nv[2]))
        if n == 0.0:
            return np.array([1, 0, 0], F)
        z = nv / n
        a = np.array([1, 0, 0], F) if abs(float(z[0])) < 0.9 else np.array([0, 1, 0], F)
        x = a - dot3(a, z) * z
        n2 = sum3(x[0] * x[0], x[1] * x[1], x[2] * x[2])
        if n2 > 0:
            x = x / np.sqrt(n2)
    return x.astype(F)


def llt_solve3(A, b):
    """E6.  A: (3,3) (lower triangle read), b: (3,) -> (ok, x)"""
    L = np.zeros((3, 3), F)
    with np.errstate(all="ignore"):
        x = A[0, 0]
        if not x > 0:
            return False, None
        L[0, 0] = np.sqrt(x)
        L[1, 0] = A[1, 0] / L[0, 0]
        L[2, 0] = A[2, 0] / L[0, 0]
        x = A[1, 1] - L[1, 0] * L[1, 0]
        if not x > 0:
            return False, None
        L[1, 1] = np.sqrt(x)
        L[2, 1] = (A[2, 1] - L[2, 0] * L[1, 0]) / L[1, 1]
        x = A[2, 2] - (L[2, 0] * L[2, 0] + L[2, 1] * L[2, 1])
        if not x > 0:
            return False, None
        L[2, 2] = np.sqrt(x)
        y0 = b[0] / L[0, 0]
        y1 = (b[1] - L[1, 0] * y0) / L[1, 1]
        y2 = (b[2] - (L[2, 0] * y0 + L[2, 1] * y1)) / L[2, 2]
        x2 = y2 / L[2, 2]
        x1 = (y1 - L[2, 1] * x2) / L[1, 1]
        x0 = (y0 - (L[1, 0] * x1 + L[2, 0] * x2)) / L[0, 0]
    return True, np.array([x0, x1, x2], F)


class RosaFloatRef:
    """Stages 3..7 of rosa_run on the outputs of rosa_ref.RosaRef (pts, nrms, surf_idx, simi lists, leaf_size)."""

    def __init__(self, pts, nrms, surf_idx, simi_off, simi_idx, leaf_size, nb_knn=10, niter_drosa=6, niter_dcrosa=5,
                 niter_smooth=3, alpha_recenter=0.3, radius_smooth=5.0, max_proj_range=10.0):
        self.pts = np.ascontiguousarray(pts, F)
        self.nrms = np.ascontiguousarray(nrms, F)
        self.M = M = self.pts.shape[0]
        self.surf = np.asarray(surf_idx).reshape(M, -1)
        self.simi = [np.asarray(simi_idx[simi_off[i]:simi_off[i + 1]], np.int64) for i in range(M)]
        self.leaf = F(leaf_size)
        self.niter_drosa, self.niter_dcrosa, self.niter_smooth = niter_drosa, niter_dcrosa, niter_smooth
        self.alpha, self.radius_smooth, self.max_proj_range = F(alpha_recenter), F(radius_smooth), F(max_proj_range)
        A = np.zeros((M, M), bool)
        for i, l in enumerate(self.simi):
            A[i, l] = True
        self.A = A
        self.t = {}

    # ---- rosa.cpp:183-199
    def rosa_init(self):
        self.pset = self.pts.copy()
        self.vset = np.stack([orthonormal_frame_x(self.nrms[i]) for i in range(self.M)]).astype(F)
        self.vvar = np.zeros(self.M, F)
        self.t["vset_init"] = self.vset.copy()

    # ---- rosa.cpp:652-683: the connected component of the seed in the similarity graph restricted to the slab
    def active(self, seed, p, n):
        Q = self.pts
        with np.errstate(all="ignore"):
            d = (n[0] * (p[0] - Q[:, 0]) + n[1] * (p[1] - Q[:, 1])) + n[2] * (p[2] - Q[:, 2])
            slab = np.abs(d) < self.leaf
        if not slab[seed]:
            return np.zeros(0, np.int64)
        vis = np.zeros(self.M, bool)
        vis[seed] = True
        front = np.array([seed])
        while front.size:
            new = self.A[front].any(axis=0) & slab & ~vis
            vis |= new
            front = np.nonzero(new)[0]
        return np.nonzero(vis)[0]

    def unit_normals(self, idx, thr_double, inv_double=False):
        """v * inv_len for the members (the three call sites differ in the comparison / reciprocal types)"""
        v = self.nrms[idx]
        with np.errstate(all="ignore"):
            s2 = sum3(v[:, 0] * v[:, 0], v[:, 1] * v[:, 1], v[:, 2] * v[:, 2])
            big = (s2.astype(np.float64) > 1e-20) if thr_double else (s2 > F(1e-20))
            if inv_double:
                inv = (1.0 / np.sqrt(s2).astype(np.float64)).astype(F)
            else:
                inv = F(1.0) / np.sqrt(s2)
            inv = np.where(big, inv, F(0.0)).astype(F)
        return v, inv, (v * inv[:, None]).astype(F)

    def normal_cov(self, idx):
        """sum_v, sum_outer -> covariance (lower triangle a00,a10,a11,a20,a21,a22): rosa.cpp:685-705 and 738-753"""
        _, _, vn = self.unit_normals(idx, thr_double=True)
        X = np.stack([vn[:, 0], vn[:, 1], vn[:, 2], vn[:, 0] * vn[:, 0], vn[:, 1] * vn[:, 0], vn[:, 1] * vn[:, 1],
                      vn[:, 2] * vn[:, 0], vn[:, 2] * vn[:, 1], vn[:, 2] * vn[:, 2]], axis=1).astype(F)
        s = lane_sum(idx, X)
        inv_m = F(1.0) / F(idx.size)
        mu = s[:3] * inv_m
        so = s[3:] * inv_m
        return (so[0] - mu[0] * mu[0], so[1] - mu[1] * mu[0], so[2] - mu[1] * mu[1],
                so[3] - mu[2] * mu[0], so[4] - mu[2] * mu[1], so[5] - mu[2] * mu[2])

    # ---- rosa.cpp:258-380
    def drosa(self):
        M = self.M
        vnew = np.zeros((M, 3), F)
        vset_iters = []
        with np.errstate(all="ignore"):
            for _ in range(self.niter_drosa):
                acts = [self.active(i, self.pset[i], self.vset[i]) for i in range(M)]
                has = np.array([a.size > 0 for a in acts])
                covs = np.zeros((M, 6), F)
                for i in np.nonzero(has)[0]:
                    covs[i] = self.normal_cov(acts[i])
                _, Q, _ = rr.eig3_sym_batch(*[covs[has, k] for k in range(6)])
                vnew[has] = Q[:, :, 0]                                             # rosa.cpp:705
                for i in range(M):                                                 # symmnormal_variance, rosa.cpp:708-735
                    if not has[i]:
                        self.vvar[i] = F(0.0)
                        continue
                    idx = acts[i]
                    v, inv, _ = self.unit_normals(idx, thr_double=False)
                    a = dot3(v, vnew[i][None, :]) * inv
                    s = lane_sum(idx, np.stack([a, a * a], axis=1).astype(F))
                    m = idx.size
                    inv_m = F(1.0) / F(m)
                    mu = s[0] * inv_m
                    var = s[1] * inv_m - mu * mu
                    if m > 1:
                        var = var * (F(m) / F(m - 1))
                    self.vvar[i] = var
                self.vset = vnew.copy()
                v2 = self.vvar * self.vvar
                self.vvar = (F(1.0) / (v2 * v2 + F(1e-5))).astype(F)
                for p in range(M):                                                 # in-place smoothing, rosa.cpp:295-308
                    snb = self.surf[p]
                    v = self.vset[snb]
                    wv = self.vvar[snb][:, None] * v                               # E4
                    terms = np.stack([wv[:, 0] * v[:, 0], wv[:, 1] * v[:, 0], wv[:, 1] * v[:, 1],
                                      wv[:, 2] * v[:, 0], wv[:, 2] * v[:, 1], wv[:, 2] * v[:, 2]], axis=1).astype(F)
                    Ms = seq0(terms)
                    _, Q = eig1(*Ms)
                    r = Q[:, 2].copy()
                    z = (r[0] * r[0] + r[1] * r[1]) + r[2] * r[2]                  # E2
                    if z > 0:
                        r = r / np.sqrt(z)
                    self.vset[p] = r
                vset_iters.append(self.vset.copy())
            self.t["vset_iters"] = np.concatenate(vset_iters) if vset_iters else np.zeros((0, 3), F)
            self.t["vvar"] = self.vvar.copy()
            # ---- position, rosa.cpp:311-365
            good = np.zeros(M, bool)
            centers = []
            poor = []
            active_size = np.zeros(M, np.int32)
            half_range = F(0.5) * self.max_proj_range
            for i in range(M):
                p = self.pset[i].copy()
                idx = self.active(i, p, self.vset[i])
                active_size[i] = idx.size
                if idx.size == 0:
                    poor.append(i)
                    continue
                ev, _ = eig1(*self.normal_cov(idx))
                planar = (ev[1] - ev[0]) / max(ev[2], F(1e-6))
                if planar < F(0.1):
                    _, _, vn = self.unit_normals(idx, thr_double=False)
                    s = lane_sum(idx, np.concatenate([self.pts[idx], vn], axis=1).astype(F))
                    inv_m = F(1.0) / F(idx.size)
                    center = s[:3] * inv_m - (s[3:] * inv_m) * half_range
                else:
                    center = self.closest_projection_point(idx)
                finite = np.isfinite(center).all()
                dlt = center - p
                in_range = np.sqrt(sum3(dlt[0] * dlt[0], dlt[1] * dlt[1], dlt[2] * dlt[2])) < self.max_proj_range
                if finite and in_range:
                    self.pset[i] = center
                    good[i] = True
                    centers.append(center.astype(F))
                else:
                    poor.append(i)
            # ---- fix-up, rosa.cpp:367-376: the tree holds tmp_pt_ = the good points at their place, (0,0,0) elsewhere
            tmp = np.where(good[:, None], self.pts, F(0.0)).astype(F)
            for i in poor:
                d2 = d2_flann(self.pts[i][None, :], tmp)
                gid = int(np.lexsort((np.arange(M), d2))[0])
                if gid < len(centers):                       # goodCenters is indexed by push order (the reference's mix-up)
                    self.pset[i] = centers[gid]
        self.t["active_size"] = active_size
        self.t["poor_idx"] = np.array(poor, np.int32)
        self.t["pset_drosa"] = self.pset.copy()

    # ---- rosa.cpp:755-790
    def closest_projection_point(self, idx):
        v, _, vn = self.unit_normals(idx, thr_double=False, inv_double=True)
        p = self.pts[idx]
        one, zero = F(1.0), F(0.0)
        P = {}
        for i in range(3):
            for j in range(3):
                P[i, j] = ((one if i == j else zero) - vn[:, i] * vn[:, j]).astype(F)
        cols = [P[i, j] for i in range(3) for j in range(3)]
        for i in range(3):
            cols.append(P[i, 0] * p[:, 0] + (P[i, 1] * p[:, 1] + P[i, 2] * p[:, 2]))       # E1
        s = lane_sum(idx, np.stack(cols, axis=1).astype(F))
        Mm = s[:9].reshape(3, 3).copy()
        B = s[9:].copy()
        tr = sum3(Mm[0, 0], Mm[1, 1], Mm[2, 2])
        reg = F(1e-6) * max(tr, F(1e-6))
        for i in range(3):
            Mm[i, i] = Mm[i, i] + reg
        ok, X = llt_solve3(Mm, B)
        if not ok:
            raise NotImplementedError("LDLT fall-back of closest_projection_point (rosa.cpp:781-788) is not restated here")
        return X

    # ---- rosa.cpp:382-422, 792-841
    def dcrosa(self):
        M = self.M
        lens = np.array([l.size for l in self.simi])
        L = max(int(lens.max()), 1)
        pad = np.zeros((M, L), np.int64)
        mask = np.zeros((M, L), bool)
        for i, l in enumerate(self.simi):
            pad[i, :l.size] = l
            mask[i, :l.size] = True
        lf = lens.astype(F)
        with np.errstate(all="ignore"):
            for _ in range(self.niter_dcrosa):
                vals = np.where(mask[:, :, None], self.pset[pad], F(0.0)).astype(F)
                acc = seq_sum_lists(vals, lens)
                new = np.where((lens > 0)[:, None], acc / lf[:, None], self.pset).astype(F)
                self.pset = new
                # local_line_fit on the averaged set
                vals = np.where(mask[:, :, None], self.pset[pad], F(0.0)).astype(F)
                mean = seq_sum_lists(vals, lens) / lf[:, None]
                q = np.where(mask[:, :, None], self.pset[pad] - mean[:, None, :], F(0.0)).astype(F)
                terms = np.stack([q[:, :, 0] * q[:, :, 0], q[:, :, 1] * q[:, :, 0], q[:, :, 1] * q[:, :, 1],
                                  q[:, :, 2] * q[:, :, 0], q[:, :, 2] * q[:, :, 1], q[:, :, 2] * q[:, :, 2]], axis=2).astype(F)
                C = seq_sum_lists(terms, lens) / lf[:, None]
                fit = lens >= 3
                ev, Q, _ = rr.eig3_sym_batch(*[C[fit, k] for k in range(6)])
                denom = np.maximum(F(1e-6), sum3(ev[:, 0], ev[:, 1], ev[:, 2]))
                dirv = Q[:, :, 2]
                nrm = np.sqrt(sum3(dirv[:, 0] * dirv[:, 0], dirv[:, 1] * dirv[:, 1], dirv[:, 2] * dirv[:, 2]))
                unitx = np.zeros_like(dirv); unitx[:, 0] = 1
                dirv = np.where((nrm > F(1e-6))[:, None], dirv / nrm[:, None], unitx).astype(F)
                conf = ev[:, 2] / denom
                x = self.pset[fit]
                mf = mean[fit]
                t = dot3(dirv, x - mf)
                proj = mf + t[:, None] * dirv
                out = self.pset.copy()
                out[fit] = np.where((conf > F(0.5))[:, None], proj, x)
                self.pset = out.astype(F)
        self.t["pset"] = self.pset.copy()

    # ---- rosa.cpp:424-532
    def vertex_sampling(self):
        fin = np.isfinite(self.pset).all(axis=1)
        pc = self.pset[fin]
        Mc = pc.shape[0]
        R = self.leaf
        r2 = F(R * R)
        covered = np.zeros(Mc, bool)
        corresp = np.full(Mc, -1, np.int32)
        mind2 = np.full(Mc, np.inf, F)
        skel, seeds = [], []
        with np.errstate(all="ignore"):
            while not covered.all():
                seed = int(np.argmin(covered))
                vid = len(skel)
                skel.append(self.pset[seed].copy())        # pset.row(seed): ORIGINAL row, compacted index (rosa.cpp:475)
                seeds.append(seed)
                d2 = d2_flann(pc[seed][None, :], pc)
                hit = d2 < r2
                if not hit.any():
                    covered[seed] = True; mind2[seed] = F(0.0); corresp[seed] = vid
                    continue
                upd = hit & (d2 < mind2)
                mind2[upd] = d2[upd]; corresp[upd] = vid
                covered |= hit
            V = len(skel)
            skel = np.array(skel, F).reshape(V, 3)
            c1 = F(1.0 - float(self.alpha))                # E7: (1.0 - alpha) in double, narrowed
            for v in range(V):
                idx = np.nonzero(corresp == v)[0]
                if idx.size == 0:
                    continue
                s = seq0(self.pts[idx])                               # RD.pts_ at the COMPACTED index (rosa.cpp:523)
                ec = s / F(idx.size)
                skel[v] = self.alpha * skel[v] + c1 * ec
        self.Mc = Mc
        self.t["seeds"] = np.array(seeds, np.int32)
        self.t["corresp"] = corresp
        self.t["skelver_sampled"] = skel.copy()
        self.skel = skel

    # ---- rosa.cpp:534-620
    def vertex_smooth(self):
        sk = self.skel
        fin = np.isfinite(sk).all(axis=1)
        tp = sk[fin]
        V = tp.shape[0]
        r2 = F(self.radius_smooth * self.radius_smooth)
        nbrs = []
        with np.errstate(all="ignore"):
            for i in range(V):
                d2 = d2_flann(tp[i][None, :], tp)
                cand = np.nonzero(d2 < r2)[0]
                cand = cand[np.lexsort((cand, d2[cand]))]
                if cand.size == 0:
                    nbrs.append(np.array([i]))
                elif cand[0] != i:
                    nbrs.append(np.r_[i, cand])
                else:
                    nbrs.append(cand)
            cur = sk.copy()
            one_m = F(1.0) - self.alpha
            for _ in range(self.niter_smooth):
                nxt = cur.copy()
                for i in range(V):
                    nb = nbrs[i]
                    n = nb.size
                    if n <= 1:
                        continue
                    mean = seq0(cur[nb]) / F(n)
                    d = cur[nb] - mean[None, :]
                    terms = np.stack([d[:, 0] * d[:, 0], d[:, 1] * d[:, 0], d[:, 1] * d[:, 1],
                                      d[:, 2] * d[:, 0], d[:, 2] * d[:, 1], d[:, 2] * d[:, 2]], axis=1).astype(F)
                    C = seq0(terms) / F(n)
                    _, Q = eig1(*C)
                    pd = Q[:, 2]
                    p = cur[i]
                    proj = mean + pd * dot3(pd, p - mean)
                    nxt[i] = self.alpha * p + one_m * proj
                cur = nxt
        self.t["skelver"] = cur.copy()
        self.skel = cur

    def run(self):
        self.rosa_init()
        self.drosa()
        self.dcrosa()
        self.vertex_sampling()
        self.vertex_smooth()
        return self.t
