"""Deterministic synthetic SIFT-like frame chains (SURVEY.md section 8d recipe).

Descriptors: Gamma(0.6) -> L2-normalise -> clip 0.2 -> renormalise -> rint(512 g) clipped to
[0,255] (u8, sparse, heavy at small values).  Frame k+1 inherits `corr_frac` of frame k's
keypoints (descriptor + rint(N(0, desc_sigma)) noise, coordinates through a per-pair
ground-truth homography + N(0, coord_sigma) px); a fraction `outlier_frac` of the inherited
keypoints gets uniform-random coordinates instead (descriptor still matches -> RANSAC outlier);
the rest of the frame is fresh.  Rows are shuffled.

Runs on any torch device: CPU for the parity tests, the GPU for the full BASELINE shapes.
"""
import math
import torch

WIDTH, HEIGHT = 1920.0, 1080.0


def _sift_like(n, d, gen, device):
    g = torch._standard_gamma(torch.full((n, d), 0.6, device=device), generator=gen)
    g = g / g.norm(dim=1, keepdim=True).clamp_min(1e-12)
    g = g.clamp_max(0.2)
    g = g / g.norm(dim=1, keepdim=True).clamp_min(1e-12)
    return torch.clamp(torch.round(512.0 * g), 0, 255).to(torch.uint8)


def _rand_h(gen, device, dtype=torch.float64):
    r = torch.rand(7, generator=gen, device=device, dtype=dtype) * 2 - 1
    th = r[0] * math.radians(1.0)
    s = 1.0 + 0.01 * r[1]
    H = torch.eye(3, dtype=dtype, device=device)
    H[0, 0] = s * torch.cos(th); H[0, 1] = -s * torch.sin(th); H[0, 2] = 8.0 * r[2]
    H[1, 0] = s * torch.sin(th); H[1, 1] = s * torch.cos(th);  H[1, 2] = 8.0 * r[3]
    H[2, 0] = 1e-5 * r[4]; H[2, 1] = 1e-5 * r[5]
    return H


def make_chain(n_frames, n_kp, d=128, seed=0, device="cpu", corr_frac=0.6, outlier_frac=0.2,
               desc_sigma=6.0, coord_sigma=0.5, unmatched_frac=0.0):
    """Returns dict(desc u8 [F,N,D], coords f32 [F,N,2], H_gt f64 [F-1,3,3] mapping frame k+1 -> k,
    broken bool [F-1] (pairs built with no correspondences at all))."""
    device = torch.device(device)
    gen = torch.Generator(device=device)
    gen.manual_seed(1_000_003 * seed + 17)
    desc = torch.empty((n_frames, n_kp, d), dtype=torch.uint8, device=device)
    coords = torch.empty((n_frames, n_kp, 2), dtype=torch.float32, device=device)
    H_gt = torch.empty((max(n_frames - 1, 0), 3, 3), dtype=torch.float64, device=device)
    broken = torch.zeros(max(n_frames - 1, 0), dtype=torch.bool, device=device)
    size = torch.tensor([WIDTH, HEIGHT], device=device, dtype=torch.float32)
    desc[0] = _sift_like(n_kp, d, gen, device)
    coords[0] = torch.rand((n_kp, 2), generator=gen, device=device) * size
    n_corr = int(round(corr_frac * n_kp))
    for k in range(1, n_frames):
        Hinv = _rand_h(gen, device)                       # maps frame k-1 -> frame k
        H_gt[k - 1] = torch.linalg.inv(Hinv)
        H_gt[k - 1] /= H_gt[k - 1, 2, 2].clone()
        is_broken = bool(unmatched_frac > 0 and
                         torch.rand(1, generator=gen, device=device).item() < unmatched_frac)
        broken[k - 1] = is_broken
        nc = 0 if is_broken else n_corr
        new_d = _sift_like(n_kp, d, gen, device)
        new_c = torch.rand((n_kp, 2), generator=gen, device=device) * size
        if nc > 0:
            parents = torch.randperm(n_kp, generator=gen, device=device)[:nc]
            noise = torch.round(torch.randn((nc, d), generator=gen, device=device) * desc_sigma)
            new_d[:nc] = torch.clamp(desc[k - 1, parents].float() + noise, 0, 255).to(torch.uint8)
            p = torch.cat([coords[k - 1, parents].double(),
                           torch.ones((nc, 1), dtype=torch.float64, device=device)], 1) @ Hinv.T
            c = (p[:, :2] / p[:, 2:]).float() + torch.randn((nc, 2), generator=gen, device=device) * coord_sigma
            out = torch.rand(nc, generator=gen, device=device) < outlier_frac
            c = torch.where(out[:, None], new_c[:nc], c)
            new_c[:nc] = c
        perm = torch.randperm(n_kp, generator=gen, device=device)
        desc[k] = new_d[perm]
        coords[k] = new_c[perm]
    return dict(desc=desc, coords=coords, H_gt=H_gt, broken=broken)
