"""
Synthetic CT-like volume pairs, label masks and sphere phantoms (the sphere phantom
follows platipy/imaging/tests/test_cardiac.py:35-71).  Data generation only -- not part of the hot path.
torch is used as an array library so that the 512x512x256 cases are generated in seconds (on the GPU
when one is present); results are returned as numpy arrays.
"""
from __future__ import annotations

import numpy as np
import torch

from .sitk_compat import Image


def _dev():
    return torch.device("cuda") if torch.cuda.is_available() else torch.device("cpu")


def _grid(size_xyz, device):
    nx, ny, nz = size_xyz
    z = torch.arange(nz, dtype=torch.float32, device=device).view(nz, 1, 1)
    y = torch.arange(ny, dtype=torch.float32, device=device).view(1, ny, 1)
    x = torch.arange(nx, dtype=torch.float32, device=device).view(1, 1, nx)
    return x, y, z


def _phantom(x, y, z, size_xyz, blobs):
    """-1000 HU air, ellipsoidal body (semi-axes 0.40 * size) at 0 HU, Gaussian blobs inside."""
    nx, ny, nz = size_xyz
    cx, cy, cz = (nx - 1) / 2.0, (ny - 1) / 2.0, (nz - 1) / 2.0
    r = ((x - cx) / (0.40 * nx)) ** 2 + ((y - cy) / (0.40 * ny)) ** 2 + ((z - cz) / (0.40 * nz)) ** 2
    # smooth body edge (2-voxel-ish ramp) keeps image gradients finite
    body = torch.sigmoid((1.0 - r) * 40.0)
    v = -1000.0 + 1000.0 * body
    for (bx, by, bz, sg, amp) in blobs:
        v = v + body * amp * torch.exp(-((x - bx) ** 2 + (y - by) ** 2 + (z - bz) ** 2) / (2.0 * sg * sg))
    return v


def _truth_dvf(x, y, z, size_xyz, rng, peak_mm):
    nx, ny, nz = size_xyz
    comps = []
    for _ in range(3):
        u = torch.zeros(1, device=x.device)
        for _ in range(3):
            lam = rng.uniform(0.5, 1.25) * max(size_xyz)
            ph = rng.uniform(0, 2 * np.pi, size=3)
            kx, ky, kz = (2 * np.pi / lam) * rng.uniform(0.3, 1.0, size=3)
            u = u + torch.sin(kx * x + ph[0]) * torch.sin(ky * y + ph[1]) * torch.sin(kz * z + ph[2])
        comps.append(u * (peak_mm / 3.0))
    return comps


def synth_pair(size_xyz=(64, 64, 32), seed=0, spacing=(1.0, 1.0, 1.0), origin=(0.0, 0.0, 0.0), peak_mm=6.0, n_blobs=24,
               noise_hu=5.0, moving_seed=None):
    """(fixed, moving) Float32 ``Image`` pair: moving = fixed phantom evaluated at x + u(x) with a smooth
    ground-truth displacement u (peak |u| ~ ``peak_mm`` voxels), independent N(0, noise) noise on each."""
    rng = np.random.default_rng(seed)
    nx, ny, nz = size_xyz
    dev = _dev()
    x, y, z = _grid(size_xyz, dev)
    blobs = []
    for _ in range(n_blobs):
        d = rng.normal(size=3)
        d = d / np.linalg.norm(d) * rng.uniform(0, 0.8) ** (1 / 3)
        c = np.array([(nx - 1) / 2, (ny - 1) / 2, (nz - 1) / 2]) + d * 0.40 * np.array(size_xyz)
        sg = rng.uniform(4, 16) * min(size_xyz) / 256.0 + 2.0
        blobs.append((float(c[0]), float(c[1]), float(c[2]), float(sg), float(rng.uniform(-300, 600))))
    fixed = _phantom(x, y, z, size_xyz, blobs)
    rng_m = np.random.default_rng(seed if moving_seed is None else moving_seed)
    ux, uy, uz = _truth_dvf(x, y, z, size_xyz, rng_m, peak_mm)
    moving = _phantom(x + ux, y + uy, z + uz, size_xyz, blobs)
    gen = torch.Generator(device=dev)
    gen.manual_seed(int(seed) * 7919 + 1 + (0 if moving_seed is None else int(moving_seed)))
    fixed = fixed + noise_hu * torch.randn(fixed.shape, generator=gen, device=dev)
    moving = moving + noise_hu * torch.randn(moving.shape, generator=gen, device=dev)
    f = fixed.to(torch.float32).cpu().numpy()
    m = moving.to(torch.float32).cpu().numpy()
    direction = (1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.0)
    return Image(f, spacing, origin, direction), Image(m, spacing, origin, direction)


def synth_labels(size_xyz, n_labels, seed=200):
    """``n_labels`` UInt8 random-ellipsoid masks ``[z, y, x]`` (seeds seed + k)."""
    nx, ny, nz = size_xyz
    dev = _dev()
    x, y, z = _grid(size_xyz, dev)
    out = []
    for k in range(n_labels):
        rng = np.random.default_rng(seed + k)
        c = np.array([nx, ny, nz]) * rng.uniform(0.3, 0.7, size=3)
        r = np.array([nx, ny, nz]) * rng.uniform(0.05, 0.2, size=3)
        m = (((x - c[0]) / r[0]) ** 2 + ((y - c[1]) / r[1]) ** 2 + ((z - c[2]) / r[2]) ** 2) <= 1.0
        out.append(m.to(torch.uint8).cpu().numpy())
    return out


def smooth_random_dvf(size_xyz, seed=0, peak_mm=6.0):
    """Smooth random displacement field, AoS ``[z, y, x, 3]`` float64 (mm)."""
    dev = _dev()
    x, y, z = _grid(size_xyz, dev)
    rng = np.random.default_rng(seed)
    comps = _truth_dvf(x, y, z, size_xyz, rng, peak_mm)
    nx, ny, nz = size_xyz
    full = [c.expand(nz, ny, nx) for c in comps]
    return torch.stack(full, dim=-1).to(torch.float64).cpu().numpy()


def insert_sphere(arr, sp_radius=4, sp_centre=(0, 0, 0)):
    """Same phantom primitive as the reference's test fixture (generation/image.py:19-48): voxels within
    ``sp_radius`` of ``sp_centre`` (array index order) are set to 1."""
    out = arr.copy()
    radius = [sp_radius] * 3 if not hasattr(sp_radius, "__iter__") else list(sp_radius)
    idx = np.indices(arr.shape)
    d = sum(((idx[a] - sp_centre[a]) / radius[a]) ** 2.0 for a in range(3))
    out[d <= 1] = 1
    return out


def _structure_params(size_xyz, n_structures, seed):
    nx, ny, nz = size_xyz
    out = []
    for k in range(n_structures):
        rng = np.random.default_rng(seed + k)
        c = np.array([nx, ny, nz]) * rng.uniform(0.3, 0.7, size=3)
        r = np.array([nx, ny, nz]) * rng.uniform(0.05, 0.2, size=3)
        out.append((c, r))
    return out


def synth_atlas_case(size_xyz, n_structures, spacing=(1.0, 1.0, 1.5), seed=0, atlas_seed=None, peak_mm=6.0, n_blobs=24, noise_hu=5.0,
                     structure_seed=200, as_tensors=False, max_shift_mm=10.0, similarity=True):
    """One member of a synthetic atlas cohort: the common anatomy of ``seed`` (the phantom of ``synth_pair``)
    and ``n_structures`` ellipsoid structures, seen through a random similarity (rotation <= 5 degrees about a random axis, scale
    0.95-1.05, shift <= ``max_shift_mm``, about the image centre, in physical space) followed by a smooth displacement (peak ``peak_mm``):
    image(x) = phantom(T(x) + u(T(x))), label_k(x) = ellipsoid_k(T(x) + u(T(x))).  ``atlas_seed=None`` gives the target itself
    (identity transform, no displacement), whose labels are the ground truth of the fusion; ``similarity=False`` leaves the
    similarity out (atlases that are already on the target grid, i.e. the state after the linear step).  Returns ``(ct, [labels])`` as
    ``Image`` objects, or as torch tensors on the generating device with ``as_tensors`` (Float32 ``[z, y, x]`` and UInt8)."""
    nx, ny, nz = size_xyz
    dev = _dev()
    x, y, z = _grid(size_xyz, dev)
    rng = np.random.default_rng(seed)
    blobs = []
    for _ in range(n_blobs):
        d = rng.normal(size=3)
        d = d / np.linalg.norm(d) * rng.uniform(0, 0.8) ** (1 / 3)
        c = np.array([(nx - 1) / 2, (ny - 1) / 2, (nz - 1) / 2]) + d * 0.40 * np.array(size_xyz)
        sg = rng.uniform(4, 16) * min(size_xyz) / 256.0 + 2.0
        blobs.append((float(c[0]), float(c[1]), float(c[2]), float(sg), float(rng.uniform(-300, 600))))
    if atlas_seed is None:
        xs, ys, zs = x, y, z
    else:
        ra = np.random.default_rng(1000 + atlas_seed)
        axis = ra.normal(size=3)
        axis /= np.linalg.norm(axis)
        ang = np.deg2rad(ra.uniform(-5.0, 5.0))
        K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
        R = np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * (K @ K)
        A = ra.uniform(0.95, 1.05) * R
        shift = ra.uniform(-1.0, 1.0, size=3) * float(max_shift_mm)
        if not similarity:  # atlases already aligned with the target: only the smooth displacement remains
            A, shift = np.eye(3), np.zeros(3)
        sp = np.asarray(spacing, dtype=np.float64)
        cen = 0.5 * (np.array(size_xyz) - 1) * sp
        # physical p = x * spacing; p' = A (p - c) + c + shift; back to voxel coordinates
        px, py, pz = x * sp[0] - cen[0], y * sp[1] - cen[1], z * sp[2] - cen[2]
        qx = (A[0, 0] * px + A[0, 1] * py + A[0, 2] * pz + cen[0] + shift[0]) / sp[0]
        qy = (A[1, 0] * px + A[1, 1] * py + A[1, 2] * pz + cen[1] + shift[1]) / sp[1]
        qz = (A[2, 0] * px + A[2, 1] * py + A[2, 2] * pz + cen[2] + shift[2]) / sp[2]
        ux, uy, uz = _truth_dvf(qx, qy, qz, size_xyz, np.random.default_rng(100 + atlas_seed), peak_mm)
        xs, ys, zs = qx + ux / sp[0], qy + uy / sp[1], qz + uz / sp[2]
    ct = _phantom(xs, ys, zs, size_xyz, blobs)
    gen = torch.Generator(device=dev)
    gen.manual_seed(int(seed) * 7919 + 17 + (0 if atlas_seed is None else 31 * int(atlas_seed)))
    ct = (ct + noise_hu * torch.randn(ct.shape, generator=gen, device=dev)).to(torch.float32).expand(nz, ny, nx).contiguous()
    labels = []
    for c, r in _structure_params(size_xyz, n_structures, structure_seed):
        m = (((xs - c[0]) / r[0]) ** 2 + ((ys - c[1]) / r[1]) ** 2 + ((zs - c[2]) / r[2]) ** 2) <= 1.0
        labels.append(m.expand(nz, ny, nx).to(torch.uint8).contiguous())
    if as_tensors:
        return ct, labels
    direction = (1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.0)
    return (Image(ct.cpu().numpy(), spacing, (0.0, 0.0, 0.0), direction),
            [Image(l.cpu().numpy(), spacing, (0.0, 0.0, 0.0), direction) for l in labels])
