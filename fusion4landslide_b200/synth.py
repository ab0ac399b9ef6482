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
