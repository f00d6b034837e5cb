"""Seeded synthetic SE2 pose + line-landmark graphs of the shapes named in BASELINE.json / SURVEY.md section 8d.

The reference stores no graphs (``*.g2o`` is git-ignored and the custom types' read/write are no-ops,
reference ``src/sparse_gslam/src/g2o_bindings/edge_se2_rhotheta.cpp:18-23``), so these are *shapes*:
vertex/edge counts of the public datasets, the reference's key-frame spacing (0.5 m, ``drone.cpp:111-112``),
sensor range (5 m, ``datasets/intel-lab/slam-11.yaml:11``), id scheme (poses 0.., landmarks from 10 000 000,
``include/drone.h:22``), first pose fixed (``drone.cpp:66``), landmark initialised from its first observation
(``drone.cpp:246``), and insertion order per key-frame: odometry edge, then that pose's line observations
(``drone.cpp:116-142``), then closures (``submap_loop_closer.cpp:272-285``).

Everything is float64 / int32 SoA, exactly the layout ``sgb_graph_soa`` (include/sgb_capi.h) takes.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import os

import numpy as np

LANDMARK_ID0 = 10_000_000  # reference include/drone.h:22
TWO_PI = 2.0 * np.pi

ODOM_SIGMA = (0.02, 0.02, 0.01)
LINE_SIGMA = (0.03, 0.02)
CLOSURE_SIGMA = (0.05, 0.05, 0.02)
RANGE_MAX = 5.0


def wrap(a):
    """g2o normalize_theta: floor-based wrap into [-pi, pi)."""
    a = np.asarray(a, dtype=np.float64)
    return a - TWO_PI * np.floor((a + np.pi) / TWO_PI)


@dataclass
class Graph:
    name: str
    pose_id: np.ndarray
    pose_est: np.ndarray
    pose_fixed: np.ndarray
    pose_gt: np.ndarray
    lm_id: np.ndarray
    lm_est: np.ndarray
    lm_fixed: np.ndarray
    lm_gt: np.ndarray
    pp_i: np.ndarray
    pp_j: np.ndarray
    pp_z: np.ndarray
    pp_info: np.ndarray
    pp_phi: np.ndarray
    pp_seq: np.ndarray
    pl_pose: np.ndarray
    pl_lm: np.ndarray
    pl_z: np.ndarray
    pl_info: np.ndarray
    pl_seq: np.ndarray
    meta: dict = field(default_factory=dict)

    @property
    def P(self):
        return int(self.pose_est.shape[0])

    @property
    def L(self):
        return int(self.lm_est.shape[0])

    @property
    def n_pp(self):
        return int(self.pp_i.shape[0])

    @property
    def n_pl(self):
        return int(self.pl_pose.shape[0])

    def chain_prefix(self, n: int) -> "Graph":
        """The landmark graph of the first n key-frames: poses < n, their odometry edges and line observations (what the
        reference re-optimises after key-frame n, drone.cpp:146-156; loop closures live in the other graph)."""
        import dataclasses
        mpp = (self.pp_j - self.pp_i == 1) & (self.pp_j < n)
        mpl = self.pl_pose < n
        return dataclasses.replace(
            self, name=f"{self.name}[:{n}]", pose_id=self.pose_id[:n], pose_est=self.pose_est[:n].copy(),
            pose_fixed=self.pose_fixed[:n], pose_gt=self.pose_gt[:n], pp_i=self.pp_i[mpp], pp_j=self.pp_j[mpp],
            pp_z=self.pp_z[mpp], pp_info=self.pp_info[mpp], pp_phi=self.pp_phi[mpp], pp_seq=self.pp_seq[mpp],
            pl_pose=self.pl_pose[mpl], pl_lm=self.pl_lm[mpl], pl_z=self.pl_z[mpl], pl_info=self.pl_info[mpl],
            pl_seq=self.pl_seq[mpl], meta=dict(self.meta))

    def pose_only(self, phi: float | None = None) -> "Graph":
        """Pose-graph view (what SubmapLoopCloser optimises with GN + DCS on closures)."""
        g = Graph(
            name=self.name + "-poses", pose_id=self.pose_id, pose_est=self.pose_est.copy(), pose_fixed=self.pose_fixed,
            pose_gt=self.pose_gt, lm_id=np.zeros(0, np.int32), lm_est=np.zeros((0, 2)), lm_fixed=np.zeros(0, np.uint8),
            lm_gt=np.zeros((0, 2)), pp_i=self.pp_i, pp_j=self.pp_j, pp_z=self.pp_z, pp_info=self.pp_info,
            pp_phi=self.pp_phi.copy(), pp_seq=self.pp_seq, pl_pose=np.zeros(0, np.int32), pl_lm=np.zeros(0, np.int32),
            pl_z=np.zeros((0, 2)), pl_info=np.zeros((0, 3)), pl_seq=np.zeros(0, np.int64), meta=dict(self.meta))
        if phi is not None:
            is_closure = (self.pp_j - self.pp_i) != 1
            g.pp_phi = np.where(is_closure, float(phi), 0.0)
        return g

    def copy(self) -> "Graph":
        kw = {k: (v.copy() if isinstance(v, np.ndarray) else (dict(v) if isinstance(v, dict) else v))
              for k, v in self.__dict__.items()}
        return Graph(**kw)


# ----------------------------------------------------------------------------- SE2 helpers (vectorised)
def se2_between(a, b):
    """inv(a) * b for arrays [...,3]."""
    c, s = np.cos(a[..., 2]), np.sin(a[..., 2])
    dx, dy = b[..., 0] - a[..., 0], b[..., 1] - a[..., 1]
    return np.stack([c * dx + s * dy, -s * dx + c * dy, wrap(b[..., 2] - a[..., 2])], axis=-1)


def line_in_pose_frame(pose, line):
    """(rho, alpha) of a world line seen from a pose; rho >= 0 (ls_extractor/utils.h:23-45 semantics)."""
    rho = line[..., 0] - pose[..., 0] * np.cos(line[..., 1]) - pose[..., 1] * np.sin(line[..., 1])
    al = line[..., 1] - pose[..., 2]
    flip = rho < 0
    rho = np.where(flip, -rho, rho)
    al = np.where(flip, al + np.pi, al)
    return np.stack([rho, wrap(al)], axis=-1)


def line_to_world(pose, z):
    """Inverse of line_in_pose_frame: world (rho, alpha) of an observation z made from pose."""
    al = z[..., 1] + pose[..., 2]
    rho = z[..., 0] + pose[..., 0] * np.cos(al) + pose[..., 1] * np.sin(al)
    flip = rho < 0
    rho = np.where(flip, -rho, rho)
    al = np.where(flip, al + np.pi, al)
    return np.stack([rho, wrap(al)], axis=-1)


def _corr_noise(rng, n, sigma, corr=0.25):
    """Per-edge SPD covariance D*C*D with random mild correlations; returns (noise[n,d], info_upper[n,d(d+1)/2])."""
    d = len(sigma)
    sig = np.asarray(sigma, dtype=np.float64) * rng.uniform(0.8, 1.25, size=(n, d))
    C = np.tile(np.eye(d), (n, 1, 1))
    iu = np.triu_indices(d, 1)
    r = rng.uniform(-corr, corr, size=(n, len(iu[0])))
    C[:, iu[0], iu[1]] = r
    C[:, iu[1], iu[0]] = r
    cov = sig[:, :, None] * C * sig[:, None, :]
    Lc = np.linalg.cholesky(cov)
    u = rng.standard_normal((n, d))
    noise = np.einsum("nij,nj->ni", Lc, u)
    info = np.linalg.inv(cov)
    info = 0.5 * (info + np.transpose(info, (0, 2, 1)))
    tu = np.triu_indices(d)
    return noise, np.ascontiguousarray(info[:, tu[0], tu[1]])


def _dead_reckon(gt0, odo_z, anchors=None, gt=None):
    """est_{i+1} = est_i (+) z_i, vectorised; optionally re-anchored to gt at the given sorted indices."""
    P = odo_z.shape[0] + 1
    est = np.zeros((P, 3))
    starts = np.array([0]) if anchors is None else np.unique(np.concatenate([[0], anchors]))
    starts = starts[starts < P]
    ends = np.concatenate([starts[1:], [P]])
    for s, e in zip(starts, ends):
        base = gt0 if s == 0 else gt[s]
        est[s] = base
        if e - s <= 1:
            continue
        z = odo_z[s:e - 1]
        th = base[2] + np.concatenate([[0.0], np.cumsum(z[:, 2])])
        c, sn = np.cos(th[:-1]), np.sin(th[:-1])
        dx = c * z[:, 0] - sn * z[:, 1]
        dy = sn * z[:, 0] + c * z[:, 1]
        est[s + 1:e, 0] = base[0] + np.cumsum(dx)
        est[s + 1:e, 1] = base[1] + np.cumsum(dy)
        est[s + 1:e, 2] = wrap(th[1:])
    return est


def _assemble(name, gt, lines_gt, pl_pairs, closures, rng, *, phi=0.0, reanchor=0, odom_sigma=ODOM_SIGMA,
              line_sigma=LINE_SIGMA, closure_sigma=CLOSURE_SIGMA, noise_scale=1.0, noise_free=False, meta=None):
    """Common tail: measurements, information matrices, initial guesses, insertion order."""
    P = gt.shape[0]
    L = lines_gt.shape[0]
    osig = tuple(noise_scale * s for s in odom_sigma)
    lsig = tuple(noise_scale * s for s in line_sigma)
    csig = tuple(noise_scale * s for s in closure_sigma)
    # odometry edges i -> i+1
    oi = np.arange(P - 1, dtype=np.int32)
    z_odo = se2_between(gt[:-1], gt[1:])
    n_o, info_o = _corr_noise(rng, P - 1, osig)
    if noise_free:
        n_o = np.zeros_like(n_o)
    z_odo = z_odo + n_o
    z_odo[:, 2] = wrap(z_odo[:, 2])
    # closures (i < j)
    closures = np.asarray(closures, dtype=np.int64).reshape(-1, 2)
    ci = np.minimum(closures[:, 0], closures[:, 1]).astype(np.int32)
    cj = np.maximum(closures[:, 0], closures[:, 1]).astype(np.int32)
    z_c = se2_between(gt[ci], gt[cj])
    n_c, info_c = _corr_noise(rng, len(ci), csig)
    if noise_free:
        n_c = np.zeros_like(n_c)
    z_c = z_c + n_c
    z_c[:, 2] = wrap(z_c[:, 2])
    # pose-line observations
    pl_pairs = np.asarray(pl_pairs, dtype=np.int64).reshape(-1, 2)
    pp_ = pl_pairs[:, 0].astype(np.int32)
    pl_ = pl_pairs[:, 1].astype(np.int32)
    z_l = line_in_pose_frame(gt[pp_], lines_gt[pl_])
    n_l, info_l = _corr_noise(rng, len(pp_), lsig)
    if noise_free:
        n_l = np.zeros_like(n_l)
    z_l = z_l + n_l
    neg = z_l[:, 0] < 0  # keep the measured rho non-negative like the extractor does
    z_l[neg, 0] = -z_l[neg, 0]
    z_l[neg, 1] += np.pi
    z_l[:, 1] = wrap(z_l[:, 1])

    # insertion order (g2o internalId): per key-frame k: odom edge (k-1 -> k), line observations of k, closures ending at k
    key_o = oi.astype(np.int64) + 1
    key_c = cj.astype(np.int64)
    key_l = pp_.astype(np.int64)
    keys = np.concatenate([key_o, key_l, key_c])
    rank = np.concatenate([np.zeros(len(key_o), np.int64), np.ones(len(key_l), np.int64), 2 * np.ones(len(key_c), np.int64)])
    sub = np.concatenate([np.zeros(len(key_o), np.int64), pl_.astype(np.int64), ci.astype(np.int64)])
    order = np.lexsort((sub, rank, keys))
    seq = np.empty(len(keys), np.int64)
    seq[order] = np.arange(len(keys))
    seq_o, seq_l, seq_c = np.split(seq, [len(key_o), len(key_o) + len(key_l)])

    pp_i = np.concatenate([oi, ci])
    pp_j = np.concatenate([oi + 1, cj])
    pp_z = np.concatenate([z_odo, z_c])
    pp_info = np.concatenate([info_o, info_c])
    pp_phi = np.concatenate([np.zeros(P - 1), np.full(len(ci), float(phi))])
    pp_seq = np.concatenate([seq_o, seq_c])
    so = np.argsort(pp_seq, kind="stable")
    pp_i, pp_j, pp_z, pp_info, pp_phi, pp_seq = pp_i[so], pp_j[so], pp_z[so], pp_info[so], pp_phi[so], pp_seq[so]
    sl = np.argsort(seq_l, kind="stable")
    pp_, pl_, z_l, info_l, seq_l = pp_[sl], pl_[sl], z_l[sl], info_l[sl], seq_l[sl]

    # initial guess: dead-reckoned odometry from the fixed first pose
    anchors = None
    if reanchor and reanchor > 0:
        anchors = np.arange(reanchor, P, reanchor)
    est = _dead_reckon(gt[0], z_odo, anchors, gt)
    # landmark initial estimate: first observation (lowest seq) mapped to the world through the observer's estimate
    lm_est = np.array(lines_gt, dtype=np.float64, copy=True)
    if len(pl_):
        first = np.full(L, -1, np.int64)
        # edges are sorted by seq, so the first occurrence of each landmark is its first observation
        uniq, idx = np.unique(pl_, return_index=True)
        first[uniq] = idx
        have = first >= 0
        lm_est[have] = line_to_world(est[pp_[first[have]]], z_l[first[have]])
    fixed = np.zeros(P, np.uint8)
    fixed[0] = 1
    m = dict(meta or {})
    m.update(P=P, L=L, E_o=int(len(pp_i)), E_l=int(len(pp_)), n_closures=int(len(ci)))
    return Graph(
        name=name, pose_id=np.arange(P, dtype=np.int32), pose_est=np.ascontiguousarray(est), pose_fixed=fixed,
        pose_gt=np.ascontiguousarray(gt), lm_id=(LANDMARK_ID0 + np.arange(L)).astype(np.int32),
        lm_est=np.ascontiguousarray(lm_est), lm_fixed=np.zeros(L, np.uint8), lm_gt=np.ascontiguousarray(lines_gt),
        pp_i=np.ascontiguousarray(pp_i, dtype=np.int32), pp_j=np.ascontiguousarray(pp_j, dtype=np.int32),
        pp_z=np.ascontiguousarray(pp_z), pp_info=np.ascontiguousarray(pp_info), pp_phi=np.ascontiguousarray(pp_phi),
        pp_seq=np.ascontiguousarray(pp_seq), pl_pose=np.ascontiguousarray(pp_, dtype=np.int32),
        pl_lm=np.ascontiguousarray(pl_, dtype=np.int32), pl_z=np.ascontiguousarray(z_l),
        pl_info=np.ascontiguousarray(info_l), pl_seq=np.ascontiguousarray(seq_l), meta=m)


# ----------------------------------------------------------------------------- trajectories on corridor networks
def _walk_polyline(waypoints, P, step):
    """Poses every `step` metres along a closed polyline, heading = direction of travel."""
    wp = np.asarray(waypoints, dtype=np.float64)
    out = []
    k = 0
    pos = wp[0].copy()
    nxt = 1
    while len(out) < P:
        tgt = wp[nxt % len(wp)]
        d = tgt - pos
        dist = np.hypot(*d)
        if dist < 1e-9:
            nxt += 1
            continue
        th = np.arctan2(d[1], d[0])
        n = int(np.floor(dist / step + 1e-9))
        for s in range(n + (1 if k == 0 else 0)):
            if len(out) >= P:
                break
            off = s if k == 0 else s + 1
            out.append([pos[0] + off * step * np.cos(th), pos[1] + off * step * np.sin(th), th])
        k += 1
        pos = np.array(out[-1][:2])
        if np.hypot(*(tgt - pos)) < step:
            nxt += 1
    g = np.array(out[:P])
    g[:, 2] = wrap(g[:, 2])
    return g


def _grid_walk(rng, P, step, nodes_x, nodes_y):
    """Random walk on a rectangular corridor network (nodes at the cross products), poses every `step`."""
    nx, ny = len(nodes_x), len(nodes_y)
    cur = (0, 0)
    prev = None
    out = []
    pos = np.array([nodes_x[0], nodes_y[0]], dtype=np.float64)
    while len(out) < P:
        nbrs = []
        for dxy in ((1, 0), (-1, 0), (0, 1), (0, -1)):
            c = (cur[0] + dxy[0], cur[1] + dxy[1])
            if 0 <= c[0] < nx and 0 <= c[1] < ny:
                nbrs.append(c)
        if prev is not None and len(nbrs) > 1 and rng.random() < 0.85:
            nbrs = [c for c in nbrs if c != prev]
        nx_ = nbrs[int(rng.integers(len(nbrs)))]
        tgt = np.array([nodes_x[nx_[0]], nodes_y[nx_[1]]], dtype=np.float64)
        d = tgt - pos
        dist = np.hypot(*d)
        th = np.arctan2(d[1], d[0])
        n = max(1, int(round(dist / step)))
        st = dist / n
        for s in range(1, n + 1):
            out.append([pos[0] + s * st * np.cos(th), pos[1] + s * st * np.sin(th), th])
            if len(out) >= P:
                break
        pos = tgt
        prev, cur = cur, nx_
    g = np.array(out[:P])
    g[:, 2] = wrap(g[:, 2])
    return g


def _walls_along(rng, gt, L, min_len=3.0, max_len=14.0, lateral=(0.8, 3.0)):
    """L axis-aligned wall segments placed beside the path: (rho, alpha, centre xy, half length, direction)."""
    P = gt.shape[0]
    idx = np.sort(rng.choice(P, size=L, replace=L > P))
    th = gt[idx, 2]
    horiz = np.abs(np.cos(th)) >= np.abs(np.sin(th))  # robot heading mostly along x => wall parallel to x
    side = rng.choice([-1.0, 1.0], size=L)
    lat = rng.uniform(lateral[0], lateral[1], size=L) * side
    half = 0.5 * rng.uniform(min_len, max_len, size=L)
    cx = np.where(horiz, gt[idx, 0] + rng.uniform(-2, 2, size=L), gt[idx, 0] + lat)
    cy = np.where(horiz, gt[idx, 1] + lat, gt[idx, 1] + rng.uniform(-2, 2, size=L))
    # wall parallel to x: y = cy => alpha = +-pi/2, rho = |cy| ; parallel to y: x = cx => alpha = 0 or pi
    jitter = rng.uniform(-0.03, 0.03, size=L)  # walls are not perfectly axis aligned
    alpha = np.where(horiz, np.where(cy >= 0, np.pi / 2, -np.pi / 2), np.where(cx >= 0, 0.0, np.pi)) + jitter
    rho = cx * np.cos(alpha) + cy * np.sin(alpha)
    neg = rho < 0
    rho = np.where(neg, -rho, rho)
    alpha = wrap(np.where(neg, alpha + np.pi, alpha))
    return np.stack([rho, alpha], 1), np.stack([cx, cy], 1), half


def _visible_pairs(gt, lines, centres, half, rmax=RANGE_MAX):
    """All (pose, line) with perpendicular distance <= rmax and foot point within the segment extent."""
    pairs = []
    ca, sa = np.cos(lines[:, 1]), np.sin(lines[:, 1])
    chunk = max(1, 4_000_000 // max(1, lines.shape[0]))
    for s in range(0, gt.shape[0], chunk):
        g = gt[s:s + chunk]
        dist = np.abs(lines[None, :, 0] - g[:, None, 0] * ca[None] - g[:, None, 1] * sa[None])
        along = -(g[:, None, 0] - centres[None, :, 0]) * sa[None] + (g[:, None, 1] - centres[None, :, 1]) * ca[None]
        ok = (dist <= rmax) & (dist > 0.6) & (np.abs(along) <= half[None])
        pi_, li_ = np.nonzero(ok)
        pairs.append(np.stack([pi_ + s, li_], 1))
    return np.concatenate(pairs) if pairs else np.zeros((0, 2), np.int64)


def _choose_observations(rng, cand, L, E_l):
    """Exactly E_l unique (pose, line) pairs, every line observed at least twice when possible."""
    if len(cand) < E_l:
        raise RuntimeError(f"only {len(cand)} visible pairs for E_l={E_l}")
    perm = rng.permutation(len(cand))
    cand = cand[perm]
    order = np.argsort(cand[:, 1], kind="stable")
    c2 = cand[order]
    _, start, cnt = np.unique(c2[:, 1], return_index=True, return_counts=True)
    take = np.zeros(len(c2), bool)
    for s, c in zip(start, cnt):
        take[s:s + min(2, c)] = True
    need = E_l - int(take.sum())
    if need < 0:
        raise RuntimeError("E_l too small for the landmark count")
    rest = np.nonzero(~take)[0]
    take[rng.choice(rest, size=need, replace=False)] = True
    return c2[take]


def _pick_closures(rng, gt, n, radius, min_gap):
    """n closure pairs between poses that revisit the same place (index gap >= min_gap)."""
    if n == 0:
        return np.zeros((0, 2), np.int64)
    P = gt.shape[0]
    cell = radius
    keys = np.floor(gt[:, :2] / cell).astype(np.int64)
    from collections import defaultdict
    buckets = defaultdict(list)
    for i, (kx, ky) in enumerate(keys):
        buckets[(int(kx), int(ky))].append(i)
    cand = []
    for (kx, ky), members in buckets.items():
        neigh = []
        for ax in (-1, 0, 1):
            for ay in (-1, 0, 1):
                neigh.extend(buckets.get((kx + ax, ky + ay), ()))
        neigh = np.array(neigh)
        for i in members:
            js = neigh[neigh >= i + min_gap]
            if len(js) == 0:
                continue
            d = np.hypot(gt[js, 0] - gt[i, 0], gt[js, 1] - gt[i, 1])
            for j in js[d <= radius]:
                cand.append((i, int(j)))
    cand = np.array(sorted(set(cand)), dtype=np.int64).reshape(-1, 2)
    if len(cand) < n:
        raise RuntimeError(f"only {len(cand)} closure candidates for n={n}")
    return cand[np.sort(rng.choice(len(cand), size=n, replace=False))]


def _world_graph(name, seed, gt, L, E_l, n_closures, *, closure_radius, min_gap, phi, noise_scale=1.0, reanchor=0,
                 noise_free=False, meta=None):
    rng = np.random.default_rng(seed)
    closures = _pick_closures(rng, gt, n_closures, closure_radius, min_gap)
    # oversample walls, keep L that are seen often enough
    lines, centres, half = _walls_along(rng, gt, int(L * 1.6) + 8)
    cand = _visible_pairs(gt, lines, centres, half)
    cnt = np.bincount(cand[:, 1], minlength=lines.shape[0])
    good = np.nonzero(cnt >= 2)[0]
    if len(good) < L:
        raise RuntimeError("not enough visible walls")
    keep = np.sort(rng.choice(good, size=L, replace=False))
    remap = -np.ones(lines.shape[0], np.int64)
    remap[keep] = np.arange(L)
    cand = cand[remap[cand[:, 1]] >= 0]
    cand[:, 1] = remap[cand[:, 1]]
    obs = _choose_observations(rng, cand, L, E_l)
    return _assemble(name, gt, lines[keep], obs, closures, rng, phi=phi, noise_scale=noise_scale, reanchor=reanchor,
                     noise_free=noise_free, meta=meta)


# Initial guesses: dead-reckoned noisy odometry, re-anchored to the (noisy-optimum-like) truth every `reanchor`
# poses. The reference never optimises from pure dead reckoning: it calls optimize() once per key-frame on a graph
# whose older part is already converged (drone.cpp:146-156), and its data association keeps a pose on the correct
# side of every wall it observes. Pure dead reckoning over hundreds of poses drifts by metres, crosses walls
# (rho sign flip of ls_extractor/utils.h:23-30) and lands LM in association-type local minima that the real
# pipeline cannot reach. reanchor=0 gives the pure dead-reckoned guess (used by the robustness tests).
def make_c1(seed: int = 1, noise_scale: float = 1.0, reanchor: int = 10) -> Graph:
    """C1 intel-lab-shaped: P=1228, 1227 odom + 256 closures, L=320, E_l=3700, 30 m x 30 m multi-loop floor, DCS phi=10."""
    wp = [(1, 1), (29, 1), (29, 29), (1, 29), (1, 15), (29, 15), (29, 1), (15, 1), (15, 29), (1, 29)]
    gt = _walk_polyline(wp, 1228, 0.5)
    return _world_graph("c1-intel-lab", seed, gt, 320, 3700, 256, closure_radius=1.5, min_gap=60, phi=10.0,
                        noise_scale=noise_scale, reanchor=reanchor, meta=dict(config="C1", dcs_phi=10.0))


def make_c2(seed: int = 2, noise_scale: float = 1.0, reanchor: int = 10) -> Graph:
    """C2 mit-killian-shaped: P=5489, 5488+2141 edges, L=1400, E_l=16000, 190 m x 240 m corridors, DCS phi=0.75."""
    rng = np.random.default_rng(seed + 1000)
    gt = _grid_walk(rng, 5489, 0.5, [0.0, 63.0, 127.0, 190.0], [0.0, 80.0, 160.0, 240.0])
    return _world_graph("c2-mit-killian", seed, gt, 1400, 16000, 2141, closure_radius=1.0, min_gap=200, phi=0.75,
                        noise_scale=noise_scale, reanchor=reanchor, meta=dict(config="C2", dcs_phi=0.75))


def make_c3(seed: int = 3, noise_scale: float = 1.0, reanchor: int = 10) -> Graph:
    """C3 M3500-style Manhattan world: P=3500 unit-grid random walk, 3499+1954 edges, L=900 grid-aligned lines, E_l=10500."""
    rng = np.random.default_rng(seed + 1000)
    P = 3500
    pos = np.zeros((P, 2))
    th = np.zeros(P)
    dirs = np.array([[1, 0], [0, 1], [-1, 0], [0, -1]], dtype=np.float64)
    d = 0
    run = 0
    for k in range(1, P):
        run += 1
        if run >= 4 and rng.random() < 0.45:
            d = (d + int(rng.choice([1, 3]))) % 4
            run = 0
        nxt = pos[k - 1] + dirs[d]
        if abs(nxt[0]) > 24 or abs(nxt[1]) > 24:  # stay inside a 48 m x 48 m block world
            d = (d + 1) % 4
            nxt = pos[k - 1] + dirs[d]
            run = 0
            if abs(nxt[0]) > 24 or abs(nxt[1]) > 24:
                d = (d + 1) % 4
                nxt = pos[k - 1] + dirs[d]
        pos[k] = nxt
        th[k] = d * np.pi / 2
    gt = np.column_stack([pos, wrap(th)])
    return _world_graph("c3-manhattan", seed, gt, 900, 10500, 1954, closure_radius=1.1, min_gap=30, phi=1.0,
                        noise_scale=noise_scale, reanchor=reanchor, meta=dict(config="C3", dcs_phi=1.0))


def make_c4_window(seed: int = 1000, noise_scale: float = 1.0, reanchor: int = 8) -> Graph:
    """One C4 aces-shaped sliding window: P=128, 127 odom edges, L=64, E_l=400 (landmark graph between prunes)."""
    rng = np.random.default_rng(seed + 7000)
    w = float(rng.uniform(18, 26))
    hgt = float(rng.uniform(8, 14))
    wp = [(0, 0), (w, 0), (w, hgt), (0, hgt)]
    gt = _walk_polyline(wp, 128, 0.5)
    return _world_graph(f"c4-window-{seed}", seed, gt, 64, 400, 0, closure_radius=1.0, min_gap=50, phi=0.0,
                        noise_scale=noise_scale, reanchor=reanchor, meta=dict(config="C4"))


def make_small(seed: int = 0, P: int = 80, L: int = 14, E_l: int = 180, n_closures: int = 6, noise_scale: float = 1.0,
               phi: float = 0.0, reanchor: int = 10, noise_free: bool = False) -> Graph:
    """Small graph for unit tests (a loop in a 9 m x 6 m room)."""
    wp = [(0, 0), (9, 0), (9, 6), (0, 6)]
    gt = _walk_polyline(wp, P, 0.5)
    return _world_graph(f"small-{seed}", seed, gt, L, E_l, n_closures, closure_radius=1.6, min_gap=max(8, P // 3),
                        phi=phi, noise_scale=noise_scale, reanchor=reanchor, noise_free=noise_free,
                        meta=dict(config="small"))


def make_c5(seed: int = 5, rows: int = 1000, cols: int = 1000, cell: float = 1.0, closure_p: float = 0.5,
            noise_scale: float = 1.0, reanchor: int = 4, phi: float = 0.0, wall_offset: float = 2.5) -> Graph:
    """C5 grid world: rows x cols boustrophedon path (P = rows*cols), odometry + closures to the vertically adjacent
    row with probability closure_p, one line landmark per 10-cell wall run (L = rows*cols/5), the 4 nearest lines
    observed from every pose (E_l = 4P). Fully vectorised. `reanchor` resets the dead-reckoning drift every that many
    poses: the reference always optimises a graph whose older part is already converged (one optimize() per
    key-frame, drone.cpp:146-156), which pure dead reckoning over 10^6 steps does not resemble. `wall_offset` shifts
    every wall run sideways by that many metres (the observation topology is unchanged) so that a 10 m run initialised
    from one noisy bearing does not swing across the poses that observe it."""
    assert rows % 10 == 0 and cols % 10 == 0
    rng = np.random.default_rng(seed)
    P = rows * cols
    r = np.repeat(np.arange(rows), cols)
    k = np.tile(np.arange(cols), rows)
    c = np.where(r % 2 == 0, k, cols - 1 - k)  # boustrophedon column
    x = c * cell
    y = r * cell
    th = np.where(r % 2 == 0, 0.0, -np.pi)  # heading along travel; -pi is in [-pi, pi)
    # at row ends the robot turns up towards the next row
    gt = np.column_stack([x, y, th]).astype(np.float64)
    pid = np.arange(P).reshape(rows, cols)  # pid[r, k]
    # pose index at (row, column)
    idx_rc = np.empty((rows, cols), np.int64)
    idx_rc[r, c] = np.arange(P)
    # closures: pose (r, c) -> pose (r-1, c) with prob closure_p
    rr, cc = np.meshgrid(np.arange(1, rows), np.arange(cols), indexing="ij")
    sel = rng.random(rr.shape) < closure_p
    ci = idx_rc[rr[sel] - 1, cc[sel]]
    cj = idx_rc[rr[sel], cc[sel]]
    closures = np.stack([ci, cj], 1)
    # lines: horizontal runs h(r, kk) above row r, 10 columns each; vertical runs v(c, m) right of column c, 10 rows each
    nh = rows * (cols // 10)
    nv = cols * (rows // 10)
    hr = np.repeat(np.arange(rows), cols // 10)
    h_y = (hr + 0.5) * cell + wall_offset + rng.uniform(-0.08, 0.08, size=nh)
    h_al = np.pi / 2 + rng.uniform(-0.02, 0.02, size=nh)
    hk = np.tile(np.arange(cols // 10), rows)
    h_cx = (hk * 10 + 4.5) * cell
    h_rho = h_cx * np.cos(h_al) + h_y * np.sin(h_al)
    vc = np.repeat(np.arange(cols), rows // 10)
    v_x = (vc + 0.5) * cell + wall_offset + rng.uniform(-0.08, 0.08, size=nv)
    v_al = rng.uniform(-0.02, 0.02, size=nv)
    vm = np.tile(np.arange(rows // 10), cols)
    v_cy = (vm * 10 + 4.5) * cell
    v_rho = v_x * np.cos(v_al) + v_cy * np.sin(v_al)
    lines = np.concatenate([np.stack([h_rho, h_al], 1), np.stack([v_rho, v_al], 1)])
    neg = lines[:, 0] < 0
    lines[neg, 0] *= -1
    lines[neg, 1] += np.pi
    lines[:, 1] = wrap(lines[:, 1])

    def hid(r_, c_):
        return r_ * (cols // 10) + c_ // 10

    def vid(c_, r_):
        return nh + c_ * (rows // 10) + r_ // 10

    p = np.arange(P)
    r_up = r
    r_dn = np.where(r > 0, r - 1, np.minimum(r + 1, rows - 1))
    c_rt = c
    c_lf = np.where(c > 0, c - 1, np.minimum(c + 1, cols - 1))
    obs = np.concatenate([
        np.stack([p, hid(r_up, c)], 1), np.stack([p, hid(r_dn, c)], 1),
        np.stack([p, vid(c_rt, r)], 1), np.stack([p, vid(c_lf, r)], 1)])
    return _assemble("c5-gridworld", gt, lines, obs, closures, rng, phi=phi, reanchor=reanchor, noise_scale=noise_scale,
                     meta=dict(config="C5", rows=rows, cols=cols, reanchor=reanchor))


def make(config: str, seed: int | None = None, **kw) -> Graph:
    config = config.lower()
    if config == "c1":
        return make_c1(1 if seed is None else seed, **kw)
    if config == "c2":
        return make_c2(2 if seed is None else seed, **kw)
    if config == "c3":
        return make_c3(3 if seed is None else seed, **kw)
    if config == "c4":
        return make_c4_window(1000 if seed is None else seed, **kw)
    if config == "c5":
        return make_c5(5 if seed is None else seed, **kw)
    if config == "small":
        return make_small(0 if seed is None else seed, **kw)
    raise ValueError(config)


# ----------------------------------------------------------------------------- g2o text files (C ABI: sgb_g2o_*)
def save_g2o(g: Graph, path: str) -> None:
    """Write `g` as a g2o text file (VERTEX_SE2 / VERTEX_RHOTHETA / EDGE_SE2 / EDGE_SE2_RHOTHETA / FIX) through
    sgb_g2o_save."""
    import ctypes as C

    from . import capi
    from .optimizer import pack_graph
    s, keep = pack_graph(g)
    st = capi.load().sgb_g2o_save(path.encode(), C.byref(s))
    if st != capi.OK:
        raise OSError(f"sgb_g2o_save({path}) failed with status {st}")


def load_g2o(path: str, name: str | None = None) -> Graph:
    """Read a g2o text file through sgb_g2o_load into a Graph (ground truth unknown: filled with NaN)."""
    import ctypes as C

    from . import capi
    L = capi.load()
    f = C.c_void_p()
    err = C.create_string_buffer(512)
    st = L.sgb_g2o_load(path.encode(), C.byref(f), err, len(err))
    if st != capi.OK:
        raise ValueError(err.value.decode() or f"sgb_g2o_load failed with status {st}")
    try:
        s = capi.GraphSoA()
        L.sgb_g2o_view(f, C.byref(s))

        def arr(ptr, n, ct, shape=None):
            if n == 0 or not ptr:
                a = np.zeros(0, dtype=np.dtype(ct))
            else:
                a = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ct)), shape=(n,)).copy()
            return a.reshape(shape) if shape is not None else a

        P, Lm, npp, npl = s.n_poses, s.n_landmarks, s.n_pp, s.n_pl
        return Graph(
            name=name or os.path.basename(path),
            pose_id=arr(s.pose_id, P, C.c_int32), pose_est=arr(s.pose_est, 3 * P, C.c_double, (P, 3)),
            pose_fixed=arr(s.pose_fixed, P, C.c_uint8), pose_gt=np.full((P, 3), np.nan),
            lm_id=arr(s.lm_id, Lm, C.c_int32), lm_est=arr(s.lm_est, 2 * Lm, C.c_double, (Lm, 2)),
            lm_fixed=arr(s.lm_fixed, Lm, C.c_uint8), lm_gt=np.full((Lm, 2), np.nan),
            pp_i=arr(s.pp_i, npp, C.c_int32), pp_j=arr(s.pp_j, npp, C.c_int32),
            pp_z=arr(s.pp_z, 3 * npp, C.c_double, (npp, 3)), pp_info=arr(s.pp_info, 6 * npp, C.c_double, (npp, 6)),
            pp_phi=arr(s.pp_phi, npp, C.c_double), pp_seq=arr(s.pp_seq, npp, C.c_int64),
            pl_pose=arr(s.pl_pose, npl, C.c_int32), pl_lm=arr(s.pl_lm, npl, C.c_int32),
            pl_z=arr(s.pl_z, 2 * npl, C.c_double, (npl, 2)), pl_info=arr(s.pl_info, 3 * npl, C.c_double, (npl, 3)),
            pl_seq=arr(s.pl_seq, npl, C.c_int64), meta=dict(source=path))
    finally:
        L.sgb_g2o_free(f)
