"""Synthetic two-epoch TLS tiles with known block motion (SURVEY 8d / BASELINE.md section 3).

No dataset of the reference is available offline, so benchmarks and parity tests run on seeded
synthetic tiles of the shape the reference processes: tile-local f32 coordinates, ~0.1 m point
spacing, rough terrain, 10 m checkerboard blocks of which half move rigidly, patch labels of a
supervoxel-like grid, point correspondences as the descriptor matcher would deliver them
(a fraction matched, a fraction of those wrong).  Pure torch so the same code runs on the CPU
(tests) and on the GPU (bench).
"""
import math

import torch


def _terrain(x, y, phases, amps, kx, ky):
    z = torch.zeros_like(x)
    for j in range(phases.numel()):
        z = z + amps[j] * torch.sin(kx[j] * x + ky[j] * y + phases[j])
    return z


def _axis_angle(axis, ang):
    axis = axis / axis.norm(dim=-1, keepdim=True)
    K = torch.zeros(axis.shape[0], 3, 3, dtype=axis.dtype, device=axis.device)
    K[:, 0, 1], K[:, 0, 2] = -axis[:, 2], axis[:, 1]
    K[:, 1, 0], K[:, 1, 2] = axis[:, 2], -axis[:, 0]
    K[:, 2, 0], K[:, 2, 1] = -axis[:, 1], axis[:, 0]
    eye = torch.eye(3, dtype=axis.dtype, device=axis.device).expand_as(K)
    s = torch.sin(ang)[:, None, None]
    c = torch.cos(ang)[:, None, None]
    return eye + s * K + (1 - c) * (K @ K)


def make_tile(n_pts, seed=0, device="cpu", spacing=0.1, block=10.0, patch_pts=256,
              matched_frac=0.6, outlier_frac=0.05, noise=0.005, jitter=0.04, desc_dim=0,
              origin=(0.0, 0.0)):
    """One synthetic tile.  Returns a dict of tensors on `device`:

      src (N,3) f32, tgt (N,3) f32 (permuted order), gt_tgt_of_src (N) i64 counterpart index,
      label_src (N) i64, label_tgt (N) i64 patch label (same id <=> same ground patch),
      corr3d (N,2) i64 [arange | tgt index or -1], block_of_src (N), R_gt (B,3,3), t_gt (B,3),
      [src_feat, tgt_feat (N,D) f32 unit rows when desc_dim > 0]
    """
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    f64 = torch.float64
    L = spacing * math.sqrt(n_pts)

    def U(*shape, lo=0.0, hi=1.0):
        return torch.rand(*shape, generator=g, device=device, dtype=f64) * (hi - lo) + lo

    def Nrm(*shape):
        return torch.randn(*shape, generator=g, device=device, dtype=f64)

    # terrain: 6 sinusoids, amplitude ~ 5 m * 2^-j
    J = 6
    amps = 5.0 * 0.5 ** torch.arange(1, J + 1, device=device, dtype=f64)
    wl = 40.0 * 0.6 ** torch.arange(J, device=device, dtype=f64)
    th = U(J, hi=2 * math.pi)
    kx = 2 * math.pi / wl * torch.cos(th)
    ky = 2 * math.pi / wl * torch.sin(th)
    ph = U(J, hi=2 * math.pi)

    x = U(n_pts, hi=L)
    y = U(n_pts, hi=L)
    z = _terrain(x, y, ph, amps, kx, ky)
    src = torch.stack([x, y, z], 1)

    # epoch 2: the same surface sampled at jittered positions (an independent sample of the
    # neighbourhood, not the identical point), block motion, then noise; order permuted.
    x2 = x + U(n_pts, lo=-jitter, hi=jitter)
    y2 = y + U(n_pts, lo=-jitter, hi=jitter)
    z2 = _terrain(x2, y2, ph, amps, kx, ky)
    tgt0 = torch.stack([x2, y2, z2], 1)

    nb = int(math.ceil(L / block))
    bx = torch.clamp((x / block).long(), max=nb - 1)
    by = torch.clamp((y / block).long(), max=nb - 1)
    blk = bx * nb + by
    B = nb * nb
    moving = ((torch.arange(B, device=device) // nb + torch.arange(B, device=device) % nb) % 2 == 1)
    ang = U(B, hi=math.radians(2.0)) * moving
    axis = Nrm(B, 3)
    tn = U(B, lo=0.05, hi=0.5) * moving
    tdir = Nrm(B, 3)
    tdir = tdir / tdir.norm(dim=1, keepdim=True)
    Rb = _axis_angle(axis, ang)
    cb = torch.stack([(torch.arange(B, device=device) // nb + 0.5) * block,
                      (torch.arange(B, device=device) % nb + 0.5) * block,
                      torch.zeros(B, device=device, dtype=f64)], 1).to(f64)
    tb_local = tdir * tn[:, None]
    # p' = R (p - c) + c + t  =>  global t = c + t - R c
    tb = cb + tb_local - torch.einsum("bij,bj->bi", Rb, cb)
    tgt_moved = torch.einsum("nij,nj->ni", Rb[blk], tgt0) + tb[blk]
    tgt_moved = tgt_moved + noise * Nrm(n_pts, 3)

    perm = torch.randperm(n_pts, generator=g, device=device)
    inv = torch.empty_like(perm)
    inv[perm] = torch.arange(n_pts, device=device)
    tgt = tgt_moved[perm]                      # tgt[j] = counterpart of src[perm[j]]
    gt_tgt_of_src = inv                        # src i  <-> tgt inv[i]

    # patches: xy grid sized for ~patch_pts points (supervoxel-like); the tgt label is taken on
    # the pre-motion position so equal ids denote the same piece of ground.
    side = spacing * math.sqrt(patch_pts)
    npx = int(math.ceil(L / side))
    lab_src = torch.clamp((x / side).long(), max=npx - 1) * npx + torch.clamp((y / side).long(), max=npx - 1)
    lab_tgt0 = torch.clamp((x2.clamp(0, L) / side).long(), max=npx - 1) * npx + \
        torch.clamp((y2.clamp(0, L) / side).long(), max=npx - 1)
    lab_tgt = lab_tgt0[perm]

    # correspondences: matched_frac of the src points have a match; outlier_frac of those point to
    # the counterpart of another point of the same patch (a wrong but nearby target).
    corr = torch.full((n_pts, 2), -1, dtype=torch.int64, device=device)
    corr[:, 0] = torch.arange(n_pts, device=device)
    has = torch.rand(n_pts, generator=g, device=device) < matched_frac
    wrong = has & (torch.rand(n_pts, generator=g, device=device) < outlier_frac)
    order = torch.argsort(lab_src, stable=True)
    sorted_lab = lab_src[order]
    uniq, counts = torch.unique_consecutive(sorted_lab, return_counts=True)
    starts = torch.cumsum(counts, 0) - counts
    seg_of_sorted = torch.repeat_interleave(torch.arange(uniq.numel(), device=device), counts)
    pos_in_sorted = torch.empty(n_pts, dtype=torch.int64, device=device)
    pos_in_sorted[order] = torch.arange(n_pts, device=device)
    seg = seg_of_sorted[pos_in_sorted]
    rnd = (torch.rand(n_pts, generator=g, device=device) * counts[seg].to(f64)).long().clamp(max=n_pts - 1)
    other = order[(starts[seg] + torch.minimum(rnd, counts[seg] - 1))]
    tgt_of = torch.where(wrong, gt_tgt_of_src[other], gt_tgt_of_src)
    corr[:, 1] = torch.where(has, tgt_of, torch.full_like(tgt_of, -1))

    ox, oy = origin
    off = torch.tensor([ox, oy, 0.0], dtype=f64, device=device)
    out = dict(src=(src + off).float().contiguous(), tgt=(tgt + off).float().contiguous(),
               gt_tgt_of_src=gt_tgt_of_src, label_src=lab_src, label_tgt=lab_tgt, corr3d=corr,
               block_of_src=blk, R_gt=Rb, t_gt=tb + off - torch.einsum("bij,j->bi", Rb, off),
               moving=moving, L=L, n_patches=npx * npx)
    if desc_dim > 0:
        fs = torch.randn(n_pts, desc_dim, generator=g, device=device)
        fs = fs / fs.norm(dim=1, keepdim=True)
        ft = fs + 0.15 * torch.randn(n_pts, desc_dim, generator=g, device=device)
        bad = torch.rand(n_pts, generator=g, device=device) < 0.10
        rndv = torch.randn(n_pts, desc_dim, generator=g, device=device)
        ft = torch.where(bad[:, None], rndv, ft)
        ft = ft / ft.norm(dim=1, keepdim=True)
        out["src_feat"] = fs.contiguous()
        out["tgt_feat"] = ft[perm].contiguous()
    return out


def patches_from_labels(labels, min_pts=10):
    """CSR of the reference's `prepare_pts2spt_dict` (base.py:1301-1351): patches with
    count > min_pts are kept, points of a patch in ascending index order, patches in ascending
    label order.  Returns (patch_labels (P) i64, ptr (P+1) i32, idx (sum) i32)."""
    order = torch.argsort(labels, stable=True)
    sl = labels[order]
    uniq, counts = torch.unique_consecutive(sl, return_counts=True)
    keep = counts > min_pts
    starts = torch.cumsum(counts, 0) - counts
    kept_counts = counts[keep]
    ptr = torch.zeros(kept_counts.numel() + 1, dtype=torch.int64, device=labels.device)
    ptr[1:] = torch.cumsum(kept_counts, 0)
    seg = torch.repeat_interleave(torch.arange(kept_counts.numel(), device=labels.device), kept_counts)
    within = torch.arange(int(ptr[-1]), device=labels.device) - ptr[:-1][seg]
    idx = order[starts[keep][seg] + within]
    return uniq[keep], ptr.to(torch.int32), idx.to(torch.int32)


def pair_patches(lab_s, lab_t):
    """Pairs (m, j) of kept src / tgt patches carrying the same label (the synthetic stand-in for
    the coarse matching result `spt_corres_src/tgt`, base.py:3156-3157)."""
    pos = torch.searchsorted(lab_t, lab_s).clamp(max=max(lab_t.numel() - 1, 0))
    ok = lab_t[pos] == lab_s if lab_t.numel() else torch.zeros_like(lab_s, dtype=torch.bool)
    m = torch.nonzero(ok).flatten()
    return m, pos[m]


def prepare_tile_host(src, tgt, label_src, label_tgt, corr3d, corr2d, min_pts, pairs, cls):
    """Torch-only twin of pipeline.prepare_tile for HOST tensors: builds the oracle / CPU-arm inputs (patch
    lists of the matched pairs) without touching the CUDA library.  `cls` = pipeline.TileInputs."""
    lab_s, ptr_s, idx_s = patches_from_labels(label_src, min_pts)
    lab_t, ptr_t, idx_t = patches_from_labels(label_tgt, min_pts)
    m, j = pairs if pairs is not None else pair_patches(lab_s, lab_t)
    dev = src.device

    def gather_csr(ptr, idx, sel):
        cnt = (ptr[1:] - ptr[:-1]).long()[sel]
        p = torch.zeros(sel.numel() + 1, dtype=torch.int64, device=dev)
        p[1:] = torch.cumsum(cnt, 0)
        seg = torch.repeat_interleave(torch.arange(sel.numel(), device=dev), cnt)
        within = torch.arange(int(p[-1]), device=dev) - p[:-1][seg]
        items = idx[(ptr[:-1].long()[sel])[seg] + within]
        return p.to(torch.int32), items.contiguous()

    t = cls()
    t.src, t.tgt, t.corr3d, t.corr2d = src.contiguous(), tgt.contiguous(), corr3d.contiguous(), corr2d
    t.sp_ptr, t.sp_idx = gather_csr(ptr_s, idx_s, m)
    t.tp_ptr, t.tp_idx = gather_csr(ptr_t, idx_t, j)
    tpo = torch.full((tgt.shape[0],), -1, dtype=torch.int32, device=dev)
    seg_t = torch.repeat_interleave(torch.arange(lab_t.numel(), device=dev, dtype=torch.int32),
                                    (ptr_t[1:] - ptr_t[:-1]).long())
    tpo[idx_t.long()] = seg_t
    t.tgt_patch_of_point = tpo
    t.pair_tgt_patch = j.to(torch.int32).contiguous()
    t.n_pairs = int(m.numel())
    t.n_src_items = int(t.sp_ptr[-1]) if t.n_pairs else 0
    t.n_tgt_items = int(t.tp_ptr[-1]) if t.n_pairs else 0
    return t


# ---------------------------------------------------------------------------------------------
# Scene generator v2 -- SURVEY 8(d) as written: epoch 2 is an INDEPENDENT resample of the surface (no 1:1
# counterpart), a nested 3-level superpoint-like hierarchy (~64 / 192 / 576 points), ~2 % of the points in
# patches of <= 10 points, descriptors tied to the nearest physical counterpart, 2D-lifted matches.
# ---------------------------------------------------------------------------------------------
def _grid_candidates_nn(q, r, cell, per_cell=4):
    """For every row of q (n,>=2) the nearest row of r (m,>=2) among the first `per_cell` points of each of the
    3x3 xy grid cells around it (pure torch; with ~1 point per cell this is the true nearest neighbour for all but
    a fraction of a percent of the queries).  Distances use all columns.  Returns (idx (n) i64 or -1, d2 (n))."""
    dev = q.device
    m = r.shape[0]
    nx = int(max(float(q[:, 0].max()), float(r[:, 0].max()), float(q[:, 1].max()), float(r[:, 1].max())) / cell) + 3

    def cid(p, dx=0, dy=0):
        ix = ((p[:, 0] / cell).floor().long() + 1 + dx).clamp(0, nx - 1)
        iy = ((p[:, 1] / cell).floor().long() + 1 + dy).clamp(0, nx - 1)
        return ix * nx + iy

    cr = cid(r)
    order = torch.argsort(cr, stable=True)
    sc = cr[order]
    best_d = torch.full((q.shape[0],), float("inf"), dtype=q.dtype, device=dev)
    best_i = torch.full((q.shape[0],), -1, dtype=torch.int64, device=dev)
    for dx in (-1, 0, 1):
        for dy in (-1, 0, 1):
            c = cid(q, dx, dy)
            start = torch.searchsorted(sc, c)
            for k in range(per_cell):
                pos = (start + k).clamp(max=m - 1)
                ok = sc[pos] == c
                j = order[pos]
                d = ((r[j] - q) ** 2).sum(1)
                d = torch.where(ok, d, torch.full_like(d, float("inf")))
                upd = (d < best_d) | ((d == best_d) & (j < best_i))
                best_d = torch.where(upd, d, best_d)
                best_i = torch.where(upd, j, best_i)
    return best_i, best_d


def hierarchy_labels(x, y, spacing, L, seed=0, small_frac=0.02):
    """Patch labels of a point set from its (pre-motion) xy position, int64 (4, n):
        row 0  supervoxel-like patches of ~256 points           (2 x 2 level-1 cells)
        row 1  superpoint level 1, ~64 points                   (warped 0.8 m grid at 0.1 m spacing)
        row 2  superpoint level 2, ~192 points                  (3 level-1 cells along the first axis)
        row 3  superpoint level 3, ~576 points                  (3 x 3 level-1 cells) -- the hierarchy is nested
    ~small_frac of the points sit in patches of <= 10 points (0.2 m cells picked by a hash of the cell id; they keep
    their tiny label on every row), which exercises the size gates (base.py:1309-1316, f2s3.py:222-225)."""
    s1 = spacing * 8.0
    u = x + 0.25 * s1 * torch.sin(2 * math.pi * y / (7.3 * s1))
    v = y + 0.25 * s1 * torch.sin(2 * math.pi * x / (5.9 * s1))
    n1 = int(math.ceil(L / s1)) + 4
    i = (u / s1).floor().long().clamp(-1, n1 - 3) + 1
    j = (v / s1).floor().long().clamp(-1, n1 - 3) + 1
    rows = [(i // 2) * n1 + (j // 2), i * n1 + j, (i // 3) * n1 + j, (i // 3) * n1 + (j // 3)]
    sf = 2.0 * spacing
    nf = int(math.ceil(L / sf)) + 2
    fid = (x / sf).floor().long().clamp(0, nf - 1) * nf + (y / sf).floor().long().clamp(0, nf - 1)
    h = (fid * 2654435761 + (seed + 1) * 40503) % 4294967296
    h = (h * 2246822519 + 3266489917) % 4294967296
    small = h < int(small_frac * 4294967296)
    off = n1 * n1
    return torch.stack([torch.where(small, off + fid, r) for r in rows], 0)


def make_scene(n_pts, seed=0, device="cpu", spacing=0.1, block=10.0, noise=0.005, desc_dim=0,
               desc_noise=0.15, desc_outliers=0.10, wrong_frac=0.05, frac_2d=0.0, outliers_2d=0.05,
               small_frac=0.02, origin=(0.0, 0.0)):
    """One synthetic tile, SURVEY 8(d).  Both epochs sample the same rough surface INDEPENDENTLY (n_pts points each);
    epoch 2 then moves block-wise (10 m checkerboard, half of the blocks rotate by U(0,2) degrees about a random axis
    through the block centre and shift by 0.05-0.5 m) and receives N(0,(5 mm)^2) noise.

    Returns a dict of tensors on `device`:
      src, tgt (n,3) f32; labels_src, labels_tgt (4,n) i64 (see hierarchy_labels; label_src/label_tgt = row 0);
      counterpart_of_tgt (n) i64: the source point nearest to the pre-motion position of each target point (-1: none
      within 1.5 spacings); corr3d (n,2) i64: what an exact descriptor matcher + magnitude gate delivers (target whose
      counterpart is the source point; `wrong_frac` of those point to the match of another point of the same patch);
      corr2d (n,2) i64 when frac_2d > 0: 2D-lifted matches of a frac_2d sample of the source points
      (nearest target of the same ground, `outliers_2d` wrong); src_feat/tgt_feat (n,D) unit rows when desc_dim > 0
      (target = normalize(feat[counterpart] + desc_noise * N(0,I)), `desc_outliers` replaced by random rows);
      block_of_src, R_gt (B,3,3), t_gt (B,3), moving (B) -- the known block motion."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    f64 = torch.float64
    L = spacing * math.sqrt(n_pts)
    dev = device

    def U(*shape, lo=0.0, hi=1.0):
        return torch.rand(*shape, generator=g, device=dev, dtype=f64) * (hi - lo) + lo

    def Nrm(*shape, dtype=f64):
        return torch.randn(*shape, generator=g, device=dev, dtype=dtype)

    J = 6
    amps = 5.0 * 0.5 ** torch.arange(1, J + 1, device=dev, dtype=f64)
    wl = 40.0 * 0.6 ** torch.arange(J, device=dev, dtype=f64)
    th = U(J, hi=2 * math.pi)
    kx = 2 * math.pi / wl * torch.cos(th)
    ky = 2 * math.pi / wl * torch.sin(th)
    ph = U(J, hi=2 * math.pi)

    x, y = U(n_pts, hi=L), U(n_pts, hi=L)
    src = torch.stack([x, y, _terrain(x, y, ph, amps, kx, ky)], 1)
    x2, y2 = U(n_pts, hi=L), U(n_pts, hi=L)                    # independent resample of the same surface
    tgt0 = torch.stack([x2, y2, _terrain(x2, y2, ph, amps, kx, ky)], 1)

    nb = int(math.ceil(L / block))
    ar = torch.arange(nb * nb, device=dev)

    def blk_of(px, py):
        return torch.clamp((px / block).long(), max=nb - 1) * nb + torch.clamp((py / block).long(), max=nb - 1)

    B = nb * nb
    moving = ((ar // nb + ar % nb) % 2 == 1)
    ang = U(B, hi=math.radians(2.0)) * moving
    axis = Nrm(B, 3)
    tn = U(B, lo=0.05, hi=0.5) * moving
    tdir = Nrm(B, 3)
    tdir = tdir / tdir.norm(dim=1, keepdim=True)
    Rb = _axis_angle(axis, ang)
    cb = torch.stack([(ar // nb + 0.5) * block, (ar % nb + 0.5) * block, torch.zeros(B, device=dev, dtype=f64)], 1).to(f64)
    tb = cb + tdir * tn[:, None] - torch.einsum("bij,bj->bi", Rb, cb)
    blk_t = blk_of(x2, y2)
    tgt = torch.einsum("nij,nj->ni", Rb[blk_t], tgt0) + tb[blk_t] + noise * Nrm(n_pts, 3)

    labels_src = hierarchy_labels(x, y, spacing, L, seed, small_frac)
    labels_tgt = hierarchy_labels(x2, y2, spacing, L, seed, small_frac)     # same ground <=> same id

    # physical counterpart of every target point (pre-motion position) among the source points, and the reverse
    cp, cp_d2 = _grid_candidates_nn(tgt0, src, spacing)
    cp = torch.where(cp_d2 < (1.5 * spacing) ** 2, cp, torch.full_like(cp, -1))
    ar_n = torch.arange(n_pts, device=dev)
    first_tgt = torch.full((n_pts,), n_pts, dtype=torch.int64, device=dev)
    okc = cp >= 0
    first_tgt.scatter_reduce_(0, cp[okc], ar_n[okc], reduce="amin", include_self=True)
    has = first_tgt < n_pts                                                  # a target point claims this source point

    # correspondences as an exact descriptor matcher would deliver them
    corr = torch.full((n_pts, 2), -1, dtype=torch.int64, device=dev)
    corr[:, 0] = ar_n
    good = has & (torch.rand(n_pts, generator=g, device=dev) >= desc_outliers)
    wrong = good & (torch.rand(n_pts, generator=g, device=dev) < wrong_frac)
    lab = labels_src[0]
    order = torch.argsort(lab, stable=True)
    _, counts = torch.unique_consecutive(lab[order], return_counts=True)
    starts = torch.cumsum(counts, 0) - counts
    seg_sorted = torch.repeat_interleave(torch.arange(counts.numel(), device=dev), counts)
    seg = torch.empty(n_pts, dtype=torch.int64, device=dev)
    seg[order] = seg_sorted
    rnd = (torch.rand(n_pts, generator=g, device=dev) * counts[seg].to(f64)).long()
    other = order[starts[seg] + torch.minimum(rnd, counts[seg] - 1)]
    tgt_of = torch.where(wrong & has[other], first_tgt[other], first_tgt)
    corr[:, 1] = torch.where(good, tgt_of, torch.full_like(tgt_of, -1))

    ox, oy = origin
    off = torch.tensor([ox, oy, 0.0], dtype=f64, device=dev)
    out = dict(src=(src + off).float().contiguous(), tgt=(tgt + off).float().contiguous(),
               labels_src=labels_src, labels_tgt=labels_tgt, label_src=labels_src[0], label_tgt=labels_tgt[0],
               counterpart_of_tgt=cp, corr3d=corr, block_of_src=blk_of(x, y), R_gt=Rb,
               t_gt=tb + off - torch.einsum("bij,j->bi", Rb, off), moving=moving, L=L)
    if frac_2d > 0:
        c2 = torch.full((n_pts, 2), -1, dtype=torch.int64, device=dev)
        c2[:, 0] = ar_n
        nn_t, nn_d2 = _grid_candidates_nn(src, tgt0, spacing)               # nearest target of the same ground
        pick = (torch.rand(n_pts, generator=g, device=dev) < frac_2d) & (nn_d2 < (1.5 * spacing) ** 2)
        bad2 = pick & (torch.rand(n_pts, generator=g, device=dev) < outliers_2d)
        rnd2 = (torch.rand(n_pts, generator=g, device=dev) * counts[seg].to(f64)).long()
        other2 = order[starts[seg] + torch.minimum(rnd2, counts[seg] - 1)]
        t2 = torch.where(bad2, nn_t[other2], nn_t)
        c2[:, 1] = torch.where(pick & (t2 >= 0), t2, torch.full_like(t2, -1))
        out["corr2d"] = c2
    if desc_dim > 0:
        fs = torch.randn(n_pts, desc_dim, generator=g, device=dev)
        fs = fs / fs.norm(dim=1, keepdim=True)
        rndv = torch.randn(n_pts, desc_dim, generator=g, device=dev)
        ft = fs[cp.clamp(min=0)] + desc_noise * torch.randn(n_pts, desc_dim, generator=g, device=dev)
        bad = (cp < 0) | (torch.rand(n_pts, generator=g, device=dev) < desc_outliers)
        ft = torch.where(bad[:, None], rndv, ft)
        out["src_feat"] = fs.contiguous()
        out["tgt_feat"] = (ft / ft.norm(dim=1, keepdim=True)).contiguous()
    return out
