"""Independent Python restatement of the reference's GSkel state machine -- TEST INFRASTRUCTURE.

Written from /root/reference/osep_skeleton_decomp/src/gskel.cpp, include/gskel.hpp and
include/lkf_vertex_fuse.hpp only (not from oracle/ and not from csrc/graph.cu): plain Python containers and
numpy.float32 scalars, one class per reference struct, every stage citing the lines it follows.  Its job is to be a
second, separately written statement of the same semantics, so that the C++ oracle (oracle/oracle_gskel.cpp) and the
library (csrc/gskel_dev.cu / graph.cu) are each compared with something that is not a copy of themselves.

Third-party semantics used (recalled; Eigen 3.4 / FLANN 1.9.1 as PCL 1.12 uses them):
  * Vector3f reductions (dot, squaredNorm, trace, 3x3 product coefficients) evaluate as x0 + (x1 + x2);
  * Matrix3f::inverse() = cofactors of column 0, det = c00 m00 + (c10 m10 + c20 m20), result(r,c) = cofactor(c,r) / det
    computed as cofactor * (1 / det);
  * kd-tree searches: squared float distance ((dx*dx + dy*dy) + dz*dz), radius search strict `<` on the squared
    radius, results ascending by (d2, index).
"""
import numpy as np

F = np.float32


def f32(x):
    return F(x)


def vsub(a, b):
    return (F(a[0] - b[0]), F(a[1] - b[1]), F(a[2] - b[2]))


def vadd(a, b):
    return (F(a[0] + b[0]), F(a[1] + b[1]), F(a[2] + b[2]))


def vscale(a, s):
    s = F(s)
    return (F(a[0] * s), F(a[1] * s), F(a[2] * s))


def vdiv(a, s):
    s = F(s)
    return (F(a[0] / s), F(a[1] / s), F(a[2] / s))


def red3(a, b, c):
    """Eigen's unrolled 3-element reduction."""
    return F(F(a) + F(F(b) + F(c)))


def sqnorm(a):
    return red3(F(a[0] * a[0]), F(a[1] * a[1]), F(a[2] * a[2]))


def norm(a):
    return F(np.sqrt(sqnorm(a)))


def finite3(a):
    return bool(np.isfinite(a[0]) and np.isfinite(a[1]) and np.isfinite(a[2]))


def flann_d2(a, b):
    dx, dy, dz = F(a[0] - b[0]), F(a[1] - b[1]), F(a[2] - b[2])
    return F(F(F(dx * dx) + F(dy * dy)) + F(dz * dz))


# ---- 3x3 float32 matrices as 3 rows of 3
def mat_add_diag(m, s):
    s = F(s)
    return [[F(m[r][c] + (s if r == c else F(0.0))) for c in range(3)] for r in range(3)]


def cof(m, i, j):
    i1, i2, j1, j2 = (i + 1) % 3, (i + 2) % 3, (j + 1) % 3, (j + 2) % 3
    return F(F(m[i1][j1] * m[i2][j2]) - F(m[i1][j2] * m[i2][j1]))


def mat_inverse(m):
    c0 = [cof(m, 0, 0), cof(m, 1, 0), cof(m, 2, 0)]
    det = red3(F(c0[0] * m[0][0]), F(c0[1] * m[1][0]), F(c0[2] * m[2][0]))
    invdet = F(F(1.0) / det)
    out = [[F(0.0)] * 3 for _ in range(3)]
    for c in range(3):
        out[0][c] = F(c0[c] * invdet)
    for r in (1, 2):
        for c in range(3):
            out[r][c] = F(cof(m, c, r) * invdet)
    return out


def mat_mul(a, b):
    return [[red3(F(a[r][0] * b[0][c]), F(a[r][1] * b[1][c]), F(a[r][2] * b[2][c])) for c in range(3)] for r in range(3)]


def mat_vec(a, v):
    return tuple(red3(F(a[r][0] * v[0]), F(a[r][1] * v[1]), F(a[r][2] * v[2])) for r in range(3))


class VertexLKF:
    """lkf_vertex_fuse.hpp:6-29"""

    def __init__(self, x0):
        self.x = tuple(F(v) for v in x0)
        self.P = [[F(1.0 if r == c else 0.0) for c in range(3)] for r in range(3)]     # gskel.cpp:135-136

    def update(self, z, q, r):
        P_pred = mat_add_diag(self.P, q)                # P + Q, Q = pn I (gskel.hpp:135)
        S = mat_add_diag(P_pred, r)                     # P_pred + R, R = mn I (gskel.hpp:136)
        K = mat_mul(P_pred, mat_inverse(S))
        self.x = vadd(self.x, mat_vec(K, vsub(z, self.x)))
        IK = [[F((F(1.0) if a == b else F(0.0)) - K[a][b]) for b in range(3)] for a in range(3)]
        self.P = mat_mul(IK, P_pred)

    def trace(self):
        return red3(self.P[0][0], self.P[1][1], self.P[2][2])


class Vertex:
    """gskel.hpp:70-86"""

    def __init__(self):
        self.position = (F(0), F(0), F(0))
        self.kf = None
        self.obs_count = 0
        self.unconf_check = 0
        self.type = 0
        self.smooth_iters = 0
        self.just_approved = False
        self.frozen = False
        self.conf_check = False
        self.marked_for_deletion = False
        self.updated = False

    def clone(self):
        v = Vertex()
        v.__dict__.update(self.__dict__)
        k = VertexLKF(self.kf.x)
        k.P = [row[:] for row in self.kf.P]
        v.kf = k
        return v


class GSkelRef:
    def __init__(self, gnd_th=60.0, fuse_dist_th=3.0, fuse_conf_th=0.1, lkf_pn=1e-4, lkf_mn=0.1, max_obs_wo_conf=5,
                 niter_smooth_vertex=3, vertex_smooth_coef=0.3, min_branch_length=5):
        self.gnd_th, self.fuse_dist_th, self.fuse_conf_th = F(gnd_th), F(fuse_dist_th), F(fuse_conf_th)
        self.lkf_pn, self.lkf_mn = F(lkf_pn), F(lkf_mn)
        self.max_obs_wo_conf, self.niter_smooth_vertex = int(max_obs_wo_conf), int(niter_smooth_vertex)
        self.vertex_smooth_coef, self.min_branch_length = F(vertex_smooth_coef), int(min_branch_length)
        self.new_cands = []
        self.prelim = []
        self.glob = []
        self.cloud = []               # global_vers_cloud
        self.new_idx = []
        self.joints, self.leafs, self.branches, self.adj = [], [], [], []
        self.size = 0
        self.failed_stage = 0
        self.events = {"merged": 0, "pruned": 0, "spawned": 0, "approved": 0, "deleted_prelim": 0}

    # ---- gskel.cpp:21-38
    def run(self, cands):
        self.new_cands = [tuple(F(v) for v in p[:3]) for p in np.asarray(cands, dtype=np.float32)]
        stages = [self.increment_skeleton, self.graph_adj, self.mst, self.vertex_merge, self.prune,
                  self.smooth_vertex_positions, self.extract_branches]
        with np.errstate(all="ignore"):
            for k, fn in enumerate(stages):
                if not fn():
                    self.failed_stage = k + 1
                    return False
        self.failed_stage = 0
        return True

    def size_assert(self):                                                        # gskel.cpp:591-599
        return len(self.glob) == len(self.cloud) == len(self.adj) == self.size

    # ---- gskel.cpp:40-199
    def increment_skeleton(self):
        if not self.new_cands:
            return False
        for v in self.prelim:
            v.just_approved = False
        tree = [v.position for v in self.prelim]                                  # snapshot (gskel.cpp:48-56)
        r2 = F(self.fuse_dist_th * self.fuse_dist_th)
        spawned = []
        for pt in self.new_cands:
            if not finite3(pt):
                continue
            chosen, chosen_d2 = -1, F(np.finfo(np.float32).max)
            any_in_radius = frozen_in_radius = False
            if tree:
                hits = sorted((flann_d2(pt, q), i) for i, q in enumerate(tree) if finite3(q) and flann_d2(pt, q) < r2)
                if hits:
                    any_in_radius = True
                    for d2, i in hits:
                        if self.prelim[i].frozen:
                            frozen_in_radius = True
                            continue
                        if d2 < chosen_d2:
                            chosen_d2, chosen = d2, i
            if chosen >= 0 and chosen_d2 <= r2:
                g = self.prelim[chosen]
                g.kf.update(pt, self.lkf_pn, self.lkf_mn)
                if not finite3(g.kf.x):
                    g.marked_for_deletion = True
                    continue
                g.position = g.kf.x
                g.obs_count += 1
                if (not g.conf_check) and g.kf.trace() < self.fuse_conf_th:
                    g.conf_check = g.just_approved = g.frozen = True
                    g.unconf_check = 0
                    self.events["approved"] += 1
                else:
                    g.unconf_check += 1
                if g.unconf_check > self.max_obs_wo_conf:
                    g.marked_for_deletion = True
            elif any_in_radius or frozen_in_radius:
                continue
            elif not any(sqnorm(vsub(s, pt)) <= r2 for s in spawned):
                nv = Vertex()
                nv.kf = VertexLKF(pt)
                nv.position = pt
                nv.obs_count += 1
                nv.smooth_iters = self.niter_smooth_vertex
                self.prelim.append(nv)
                spawned.append(pt)
                self.events["spawned"] += 1
        n0 = len(self.prelim)
        self.prelim = [v for v in self.prelim if not v.marked_for_deletion]
        self.events["deleted_prelim"] += n0 - len(self.prelim)
        self.new_idx = []
        for v in self.prelim:
            if v.just_approved and v.position[2] > self.gnd_th:
                self.glob.append(v.clone())          # std::move of a trivially copyable struct: the prelim entry stays (frozen)
                self.cloud.append(v.position)
                self.new_idx.append(len(self.glob) - 1)
                v.just_approved = False
                v.updated = True
        keep = [i for i, v in enumerate(self.glob) if finite3(v.position)]        # gskel.cpp:173-195
        self.glob = [self.glob[i] for i in keep]
        self.cloud = [v.position for v in self.glob]
        self.size = len(self.glob)
        return len(self.new_idx) != 0

    # ---- gskel.cpp:201-246
    def graph_adj(self):
        if not self.cloud:
            return False
        n = len(self.cloud)
        new_adj = [[] for _ in range(n)]
        K = 10
        max_dist_th = F(F(2.5) * self.fuse_dist_th)
        pos = [v.position for v in self.glob]
        for i in range(self.size):
            order = sorted((flann_d2(self.cloud[i], q), j) for j, q in enumerate(self.cloud) if finite3(q))[:K]
            idx = [j for _, j in order]
            for a in range(1, len(idx)):
                nb = idx[a]
                d_nb = norm(vsub(pos[i], pos[nb]))
                if d_nb > max_dist_th:
                    continue
                good = True
                for b in range(1, len(idx)):
                    if b == a:
                        continue
                    o = idx[b]
                    if norm(vsub(pos[nb], pos[o])) < d_nb and norm(vsub(pos[i], pos[o])) < d_nb:
                        good = False
                        break
                if good:
                    new_adj[i].append(nb)
                    new_adj[nb].append(i)
        self.adj = new_adj
        return self.size_assert()

    # ---- gskel.cpp:248-281 (ties of equal weight broken by (u, v): the reference's std::sort leaves them unspecified)
    def mst(self):
        if self.size == 0 or not self.adj:
            return False
        edges = []
        for i in range(self.size):
            for nb in self.adj[i]:
                if nb <= i:
                    continue
                edges.append((norm(vsub(self.glob[i].position, self.glob[nb].position)), i, nb))
        edges.sort()
        parent = list(range(self.size))

        def find(x):
            root = x
            while parent[root] != root:
                root = parent[root]
            while parent[x] != root:
                parent[x], x = root, parent[x]
            return root

        mst_adj = [[] for _ in range(self.size)]
        for w, u, v in edges:
            ru, rv = find(u), find(v)
            if ru == rv:
                continue
            parent[rv] = ru
            mst_adj[u].append(v)
            mst_adj[v].append(u)
        self.adj = mst_adj
        self.graph_decomp()
        return self.size_assert()

    # ---- gskel.cpp:539-567: `case 2` falls through into `default`, whose last statement resets the type to 0
    def graph_decomp(self):
        self.joints, self.leafs = [], []
        for i in range(len(self.glob)):
            deg = len(self.adj[i])
            t = 0
            if deg == 1:
                self.leafs.append(i)
                t = 1
            elif deg > 2:
                self.joints.append(i)
            self.glob[i].type = t

    # ---- gskel.cpp:569-589
    def merge_into(self, keep, dele):
        Vi, Vj = self.glob[keep], self.glob[dele]
        tot = Vi.obs_count + Vj.obs_count
        s = vadd(vscale(Vi.position, F(Vi.obs_count)), vscale(Vj.position, F(Vj.obs_count)))
        Vi.position = vdiv(s, F(tot))
        Vi.obs_count = tot
        for nb in list(self.adj[dele]):
            if nb == keep:
                continue
            if nb not in self.adj[keep]:
                self.adj[keep].append(nb)
            self.adj[nb] = [keep if x == dele else x for x in self.adj[nb]]

    # ---- gskel.cpp:283-353
    def vertex_merge(self):
        n_new = len(self.new_idx)
        if self.size == 0 or n_new == 0:
            return False
        if float(F(n_new) / F(self.size)) > 0.5:
            return False
        to_delete = set()
        for new_id in self.new_idx:
            if new_id < 0 or new_id >= self.size or new_id in to_delete:
                continue
            for nb_id in list(self.adj[new_id]):
                if nb_id < 0 or nb_id >= self.size:
                    continue
                if new_id == nb_id or nb_id in to_delete:
                    continue
                do_merge = False
                if new_id in self.joints and nb_id in self.joints:
                    do_merge = True
                    self.joints = [x for x in self.joints if x != nb_id]
                dist = norm(vsub(self.glob[new_id].position, self.glob[nb_id].position))
                if (not do_merge) and dist < F(F(0.5) * self.fuse_dist_th):
                    do_merge = True
                if not do_merge:
                    continue
                self.merge_into(nb_id, new_id)
                to_delete.add(new_id)
                self.events["merged"] += 1
                break
        if not to_delete:
            return True
        for dele in sorted(to_delete, reverse=True):
            if dele < 0 or dele > len(self.glob):          # `del > size` guard as written (gskel.cpp:329)
                continue
            del self.glob[dele]
            del self.cloud[dele]
            del self.adj[dele]
            for k in range(len(self.adj)):
                nb = [x for x in self.adj[k] if x != dele]
                nb = [x - 1 if x > dele else x for x in nb]
                self.adj[k] = sorted(set(nb))
            self.new_idx = [x for x in self.new_idx if x != dele]
            self.new_idx = [x - 1 if x > dele else x for x in self.new_idx]
        self.size = len(self.glob)
        self.graph_decomp()
        return self.size_assert()

    # ---- gskel.cpp:355-394
    def prune(self):
        n_new = len(self.new_idx)
        if self.size == 0 or n_new == 0:
            return False
        if float(F(n_new) / F(self.size)) > 0.5:
            return False
        joint_set, new_set = set(self.joints), set(self.new_idx)
        to_delete = []
        for leaf in self.leafs:
            if leaf not in new_set:
                continue
            if len(self.adj[leaf]) != 1:
                continue
            if self.adj[leaf][0] in joint_set:
                to_delete.append(leaf)
        for idx in sorted(to_delete, reverse=True):
            del self.glob[idx]
            del self.cloud[idx]
            del self.adj[idx]
            for k in range(len(self.adj)):
                self.adj[k] = [x - 1 if x > idx else x for x in self.adj[k] if x != idx]
            self.new_idx = [x - 1 if x > idx else x for x in self.new_idx if x != idx]
            self.events["pruned"] += 1
        self.size = len(self.glob)
        self.graph_decomp()
        return True

    # ---- gskel.cpp:396-434 (no guard on an empty neighbour list: 0/0 = NaN, dropped by the next tick's finite filter)
    def smooth_vertex_positions(self):
        if self.size == 0 or len(self.adj) == 0:
            return False
        new_pos = []
        for i in range(self.size):
            v = self.glob[i]
            if v.smooth_iters > 0:
                avg = (F(0), F(0), F(0))
                for j in self.adj[i]:
                    avg = vadd(avg, self.glob[j].position)
                avg = vdiv(avg, F(len(self.adj[i])))
                a = vscale(v.position, F(F(1.0) - self.vertex_smooth_coef))
                b = vscale(avg, self.vertex_smooth_coef)
                new_pos.append(vadd(a, b))
                v.smooth_iters -= 1
            else:
                new_pos.append(v.position)
        for i in range(self.size):
            self.glob[i].position = new_pos[i]
        self.cloud = [v.position for v in self.glob]
        return True

    # ---- gskel.cpp:436-536.  Range-for copies: `endp` is a per-leaf / per-joint copy that the swap at :448 / :485 mutates for
    # the REMAINING neighbours of the same vertex; the test at :486 ends in `;` and therefore filters nothing.
    def extract_branches(self):
        if self.size == 0:
            return False
        self.branches = []
        visited = set()

        def is_endpoint(v):
            return self.glob[v].type in (1, 3)

        def walk(endp, nb):
            branch = [endp]
            prev, curr = endp, nb
            guard = 0
            while guard <= self.size + 1:          # the reference loops `while (true)`: a cycle would hang it; a forest never needs the guard
                guard += 1
                branch.append(curr)
                if is_endpoint(curr) and curr != endp:
                    break
                nxt = -1
                for nn in self.adj[curr]:
                    if nn != prev:
                        nxt = nn
                        break
                if nxt == -1:
                    break
                prev, curr = curr, nxt
            for a, b in zip(branch[:-1], branch[1:]):
                visited.add((a, b))
            self.branches.append(branch)

        for leaf in self.leafs:
            endp = leaf
            for nb in list(self.adj[leaf]):
                if endp > nb:
                    endp, nb = nb, endp
                if (endp, nb) in visited:
                    continue
                walk(endp, nb)
        for joint in self.joints:
            endp = joint
            for nb in list(self.adj[joint]):
                if endp > nb:
                    endp, nb = nb, endp
                walk(endp, nb)
        self.branches.sort(key=len)                 # stable: equal lengths keep discovery order (std::sort leaves them unspecified)
        self.branches = [b for b in self.branches if len(b) >= self.min_branch_length]
        return True

    # ---- views used by the tests
    def positions(self):
        """output_gskel(): the published cloud (gskel.hpp:112).  NOT always the vertices' own positions: merge_into moves a
        vertex without touching the cloud, which is only rebuilt by smooth_vertex_positions / the next tick's finite filter."""
        return np.array([[float(c) for c in p] for p in self.cloud], dtype=np.float32).reshape(-1, 3)

    def vertex_positions(self):
        return np.array([[float(c) for c in v.position] for v in self.glob], dtype=np.float32).reshape(-1, 3)

    def prelim_positions(self):
        return np.array([[float(c) for c in v.position] for v in self.prelim], dtype=np.float32).reshape(-1, 3)

    def csr(self, lists):
        off = np.zeros(len(lists) + 1, np.int32)
        for i, l in enumerate(lists):
            off[i + 1] = off[i] + len(l)
        idx = np.array([x for l in lists for x in l], dtype=np.int32)
        return off, idx


def skeleton_graph(V, fuse_dist_th=3.0):
    """graph_adj + mst + graph_decomp (gskel.cpp:201-281, 539-567) of a bare vertex set -- what the frame's tail kernel and
    osep_rosa_graph / osep_graph_build compute.  Returns (rng lists, mst lists, mst edges in acceptance order, weights,
    types, joints, leafs)."""
    g = GSkelRef(fuse_dist_th=fuse_dist_th)
    for p in np.asarray(V, dtype=np.float32):
        v = Vertex()
        v.position = (F(p[0]), F(p[1]), F(p[2]))
        g.glob.append(v)
    g.cloud = [v.position for v in g.glob]
    g.size = len(g.glob)
    g.adj = [[] for _ in g.glob]
    with np.errstate(all="ignore"):
        if not g.graph_adj():
            return None
        rng = [list(a) for a in g.adj]
        if not g.mst():
            return None
        edges = sorted((norm(vsub(g.glob[u].position, g.glob[v].position)), u, v) for u in range(g.size) for v in g.adj[u] if v > u)
    return dict(rng=rng, mst=[list(a) for a in g.adj], edges=np.array([[u, v] for _, u, v in edges], np.int32).reshape(-1, 2),
                edge_w=np.array([w for w, _, _ in edges], np.float32), type=np.array([v.type for v in g.glob], np.int32),
                joints=np.array(g.joints, np.int32), leafs=np.array(g.leafs, np.int32))
