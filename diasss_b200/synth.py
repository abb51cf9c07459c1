"""Seeded synthetic side-scan surveys in the reference's data model.

One periodic seabed reflectivity field (band-limited noise + sparse bright "rocks") is sampled by every
waterfall image through its own ping poses and ground ranges, exactly as Frame::GetGeoImg (frame.cpp:126-165)
geo-references a bin; multiplicative speckle is added per image; the result goes through the reference's
normalisation (frame.cpp:57-81) to CV_8U.  Image ids alternate heading (lawn-mower pattern: the matcher assumes
odd/even ids run opposite ways, FEAmatcher.cpp:209-212).  Reported (dead-reckoning) poses differ from the true
ones by a bounded drift so that true matches fall inside the 8 m gate.

Used by tests/ and bench.py only (inputs, not compute).  torch ops so that the 8000 x 2000 bench images can be
generated on the GPU; everything is seeded and device-independent up to float rounding (the generated uint8
images are then the fixed input of BOTH the CUDA path and the oracle).
"""
import math

import numpy as np
import torch

PI_REF = 3.14159265359  # frame.cpp:16


def _gauss_kernel(sigma, device):
    r = max(1, int(math.ceil(3 * sigma)))
    x = torch.arange(-r, r + 1, dtype=torch.float32, device=device)
    k = torch.exp(-0.5 * (x / sigma) ** 2)
    return k / k.sum()


def _blur_periodic(t, sigma):
    k = _gauss_kernel(sigma, t.device)
    r = (len(k) - 1) // 2
    x = t[None, None]
    x = torch.nn.functional.pad(x, (r, r, 0, 0), mode="circular")
    x = torch.nn.functional.conv2d(x, k.view(1, 1, 1, -1))
    x = torch.nn.functional.pad(x, (0, 0, r, r), mode="circular")
    x = torch.nn.functional.conv2d(x, k.view(1, 1, -1, 1))
    return x[0, 0]


def seabed(size=2048, seed=1234, device="cpu", fine_amp=0.10):
    """Periodic reflectivity field, float32 [size, size], positive, mean ~1.  fine_amp scales the 2-px texture that
    makes most FAST corners (0.10: 10^4-10^5 candidates per level of a 2000 x 1000 swath; 0.02: 10^3-10^4)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    base = torch.randn(size, size, generator=g).to(device)
    fine = _blur_periodic(base, 2.0)
    fine = fine / fine.std()
    coarse = _blur_periodic(torch.randn(size, size, generator=g).to(device), 12.0)
    coarse = coarse / coarse.std()
    rocks = (torch.rand(size, size, generator=g) < 4e-4).float().to(device)
    rocks = _blur_periodic(rocks, 2.0)
    rocks = rocks / rocks.max()
    f = 1.0 + fine_amp * fine + 0.25 * coarse + 1.5 * rocks
    return f.clamp_min(0.05)


def _sample_periodic(field, u, v):
    """Bilinear sample of the periodic field at (u = x index, v = y index) float tensors."""
    size = field.shape[0]
    u0, v0 = torch.floor(u), torch.floor(v)
    fu, fv = (u - u0).float(), (v - v0).float()
    iu0 = torch.remainder(u0.long(), size); iv0 = torch.remainder(v0.long(), size)
    iu1 = torch.remainder(iu0 + 1, size); iv1 = torch.remainder(iv0 + 1, size)
    a = field[iv0, iu0] * (1 - fu) + field[iv0, iu1] * fu
    b = field[iv1, iu0] * (1 - fu) + field[iv1, iu1] * fu
    return a * (1 - fv) + b * fv


def geo_planes(rows, cols, pose, g_range, device="cpu"):
    """Frame::GetGeoImg (frame.cpp:126-165) in float64 torch ops (for sampling only; parity is checked elsewhere)."""
    pose = torch.as_tensor(pose, dtype=torch.float64, device=device)
    g = torch.as_tensor(g_range, dtype=torch.float64, device=device)
    half = cols // 2
    yaw = pose[:, 2:3]
    k_stb = torch.arange(0, cols - half, device=device)
    k_port = (cols - half) - torch.arange(0, half, device=device)
    gx = torch.empty(rows, cols, dtype=torch.float64, device=device)
    gy = torch.empty(rows, cols, dtype=torch.float64, device=device)
    gx[:, half:] = pose[:, 3:4] + g[k_stb][None, :] * torch.cos(yaw + PI_REF / 2)
    gy[:, half:] = pose[:, 4:5] + g[k_stb][None, :] * torch.sin(yaw + PI_REF / 2)
    gx[:, :half] = pose[:, 3:4] + g[k_port][None, :] * torch.cos(yaw - PI_REF / 2)
    gy[:, :half] = pose[:, 4:5] + g[k_port][None, :] * torch.sin(yaw - PI_REF / 2)
    return gx, gy


def make_mask(rows, cols, side=None, device="cpu"):
    """Frame::GetFilteredMask's fixed margins (frame.cpp:104-112) scaled to small test images: centre line +-10,
    first/last `side` pings, left/right 0.6*side bins.  255 = keep."""
    if side is None:
        side = 150 if rows >= 1500 else max(8, rows // 12)
    m = torch.full((rows, cols), 255, dtype=torch.uint8, device=device)
    width = 10
    m[:, cols // 2 - width + 1: cols // 2 + width] = 0
    m[:side, :] = 0
    m[rows - side + 1:, :] = 0
    e = int(math.ceil(side * 0.6))
    m[:, :e] = 0
    m[:, cols - e + 1:] = 0
    return m


def make_track(img_id, rows, cols, line_x, res=0.1, seed=1000, drift=(0.0, 0.0), yaw_sigma_deg=0.05):
    """Dead-reckoning data of one line: true pose, reported pose (true + drift), ground ranges.  Host, float64."""
    g = torch.Generator(device="cpu").manual_seed(seed + img_id)
    heading_up = (img_id % 2 == 0)
    t = torch.arange(rows, dtype=torch.float64)
    y = t * res if heading_up else (rows - 1 - t) * res
    yaw0 = PI_REF / 2 if heading_up else -PI_REF / 2
    yaw = yaw0 + torch.randn(rows, generator=g, dtype=torch.float64).cumsum(0) * math.radians(yaw_sigma_deg) / math.sqrt(rows)
    true_pose = torch.zeros(rows, 6, dtype=torch.float64)
    true_pose[:, 2] = yaw
    true_pose[:, 3] = line_x
    true_pose[:, 4] = y
    g_range = torch.arange(cols - cols // 2 + 1, dtype=torch.float64) * res      # B4: one more than the starboard bins
    pose = true_pose.clone()
    pose[:, 3] += drift[0]
    pose[:, 4] += drift[1]
    return dict(img_id=int(img_id), rows=rows, cols=cols, true_pose=true_pose, pose=pose.numpy(), g_range=g_range.numpy(),
                gen=g)


def render(field, track, res=0.1, speckle=0.04, device="cpu", side=None):
    """The waterfall image of a track: sample the seabed through the TRUE poses, add speckle, normalise to u8."""
    rows, cols = track["rows"], track["cols"]
    gx, gy = geo_planes(rows, cols, track["true_pose"], track["g_range"], device)
    img = _sample_periodic(field, (gx / res), (gy / res))
    del gx, gy
    noise = torch.randn(rows, cols, generator=track["gen"]).to(device)
    img = img * (1.0 + speckle * noise).clamp_min(0.05)
    # Frame::GetNormalizeSSS (frame.cpp:57-81) -- input preparation, not a parity subject here
    mean, mn = img.mean(), img.min()
    norm = ((img - mn) / (mean * 2.5 - mn) * 255.0).clamp(0, 255).round().to(torch.uint8)
    return norm, make_mask(rows, cols, side, device)


def make_frame(field, img_id, rows, cols, line_x, res=0.1, seed=1000, drift=(0.0, 0.0), speckle=0.04, yaw_sigma_deg=0.05,
               device="cpu", side=None):
    """One waterfall image + its dead-reckoning data (dict; norm_img / mask are torch tensors on `device`)."""
    tr = make_track(img_id, rows, cols, line_x, res, seed, drift, yaw_sigma_deg)
    norm, mask = render(field, tr, res, speckle, device, side)
    return dict(img_id=tr["img_id"], rows=rows, cols=cols, norm_img=norm, mask=mask, pose=tr["pose"], g_range=tr["g_range"])


def _to_numpy(f):
    out = dict(f)
    out["norm_img"] = f["norm_img"].cpu().numpy()
    out["mask"] = f["mask"].cpu().numpy()
    return out


def survey_tracks(n_images, rows, cols, seed=1234, spread=0.35, res=0.1, drift_m=1.5, ids=None):
    """Tracks of n_images overlapping lines over the same site: line k is offset by spread*swath*k/n across track so
    that EVERY pair overlaps (all-pairs matching does full work on every pair).  Headings alternate with the id."""
    rng = np.random.default_rng(seed + 17)
    swath = cols * res
    ids = list(range(n_images)) if ids is None else list(ids)
    out = []
    for k, img_id in enumerate(ids):
        line_x = 500.0 + spread * swath * k / max(n_images, 1)
        drift = rng.uniform(-drift_m, drift_m, 2)
        out.append(make_track(img_id, rows, cols, line_x, res, seed=seed + 1000, drift=drift))
    return out


def make_survey(n_images, rows, cols, seed=1234, spread=0.35, res=0.1, drift_m=1.5, device="cpu", field_size=2048,
                speckle=0.04, as_torch=False, side=None, ids=None):
    field = seabed(field_size, seed, device)
    frames = []
    for tr in survey_tracks(n_images, rows, cols, seed, spread, res, drift_m, ids):
        norm, mask = render(field, tr, res, speckle, device, side)
        f = dict(img_id=tr["img_id"], rows=rows, cols=cols, norm_img=norm, mask=mask, pose=tr["pose"], g_range=tr["g_range"])
        frames.append(f if as_torch else _to_numpy(f))
    return frames


def make_pair(rows=420, cols=360, seed=7, ids=(0, 1), **kw):
    return make_survey(2, rows, cols, seed=seed, ids=ids, **kw)


def noise_image(rows, cols, seed=0):
    """Pure-noise image (no structure shared with anything): edge case 'no true matches'."""
    rng = np.random.default_rng(seed)
    return rng.integers(0, 256, (rows, cols), dtype=np.uint8)
