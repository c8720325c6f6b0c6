"""Seeded synthetic Kinect-shaped turntable captures.

There is no sample data in the reference (the "horse" set is an external release,
README.md:161-163) and no network, so every test / bench input is rendered here:

* intrinsics and back-projection of the Kinect-v1 capture tool
  (capture/depth_capture/depth_capture.cpp:294-308: x = z(j-cx)/fx, y = -z(i-cy)/fy,
  z = -z, millimetres, rounded through "%.6f"),
* turntable geometry of rotate_align (rotate_align/rotate_align.cpp:136-138,224-235:
  rotation about +Y through the table centre, default centre (44, 60, 632.5) mm),
* Kinect-like axial noise sigma_z = 1.5 mm (z / 1 m)^2, depth quantised to uint16 mm.

Units of the returned clouds: metres (x0.001), float32.  This is data generation,
not part of the registration path.
"""
from __future__ import annotations

import numpy as np

# capture/depth_capture/depth_capture.cpp:297
KINECT_V1 = dict(width=640, height=480, fx=572.882768, fy=542.739980, cx=314.649173, cy=240.160459)
# capture/super_resolution_kv2/super_resolution_kv2.cpp:100
KINECT_V2 = dict(width=512, height=424, fx=365.531799, fy=365.531799, cx=256.136810, cy=206.013901)

# Table centre in the capture tool's output frame (mm): x right, y up, z = -depth.
TABLE_CENTRE_MM = np.array([44.0, -20.0, -632.5])


def _rot_y(a: float) -> np.ndarray:
    c, s = np.cos(a), np.sin(a)
    return np.array([[c, 0.0, s], [0.0, 1.0, 0.0], [-s, 0.0, c]])


def _bumpy_radius(ux: np.ndarray, uy: np.ndarray, uz: np.ndarray) -> np.ndarray:
    """Radius (mm) of the bumpy sphere along unit direction u: low-order polynomial
    lobes (no trigonometry, cheap to ray-march), registrable in all 6 dof."""
    return 150.0 * (1.0 + 0.10 * (4.0 * ux * ux * ux - 3.0 * ux) * (1.0 - uy * uy)
                    + 0.07 * uy * (uz * uz - ux * ux) * 3.0
                    + 0.06 * uz * ux * (1.0 + uy) + 0.05 * (2.0 * uy * uy - 1.0) * uz)


def _scene_sdf(p: np.ndarray, backdrop: str) -> np.ndarray:
    """Distance bound (mm) of the rigid scene in its own frame (origin = object centre,
    y up).  Lipschitz constant < 2; the ray marcher steps by 0.5 f."""
    x, y, z = p[..., 0], p[..., 1], p[..., 2]
    r = np.sqrt(x * x + y * y + z * z) + 1e-9
    f = r - _bumpy_radius(x / r, y / r, z / r)
    # turntable disc: radius 260 mm, top at y = -150 mm, 20 mm thick
    rho = np.sqrt(x * x + z * z)
    dy = np.abs(y + 160.0) - 10.0
    dr = rho - 260.0
    disc = np.minimum(np.maximum(dy, dr), 0.0) + np.sqrt(np.maximum(dy, 0.0) ** 2 + np.maximum(dr, 0.0) ** 2)
    f = np.minimum(f, disc)
    if backdrop != "none":
        # wavy backdrop 330 mm behind the object centre (toward -z)
        back = (z + 330.0) + 28.0 * np.sin(x / 55.0) * np.cos(y / 45.0) + 12.0 * np.sin(x / 17.0 + y / 23.0)
        if backdrop == "panel":  # finite panel: ~200k valid pixels at 640x480 (cfg1)
            back = np.maximum(back, np.maximum(np.abs(x) - 420.0, np.abs(y + 10.0) - 300.0))
        f = np.minimum(f, back)
    return f


def _rot_x(a: float) -> np.ndarray:
    c, s = np.cos(a), np.sin(a)
    return np.array([[1.0, 0.0, 0.0], [0.0, c, -s], [0.0, s, c]])


def render_depth(view_angle_rad: float, intr: dict = KINECT_V1, scale: float = 1.0,
                 backdrop: str = "none", seed: int | None = 0, noise: bool = True,
                 depth_max_mm: float = 2000.0, tilt_rad: float = 0.0) -> np.ndarray:
    """Depth image (uint16 mm, 0 = invalid) of the scene rotated by view_angle about +Y
    through the table centre.  backdrop: "none" (object + table), "panel" (finite
    backdrop, ~65 % fill: cfg1) or "full" (every pixel valid: cfg2).  `scale` < 1
    renders a proportionally smaller image with scaled intrinsics (CPU-sized tests)."""
    w, h = int(round(intr["width"] * scale)), int(round(intr["height"] * scale))
    fx, fy, cx, cy = intr["fx"] * scale, intr["fy"] * scale, intr["cx"] * scale, intr["cy"] * scale
    jj, ii = np.meshgrid(np.arange(w, dtype=np.float64), np.arange(h, dtype=np.float64))
    # ray through pixel, parameterised by depth z>0: p = z * d, d = ((j-cx)/fx, -(i-cy)/fy, -1)
    d = np.stack([(jj - cx) / fx, -(ii - cy) / fy, -np.ones_like(jj)], axis=-1).reshape(-1, 3)
    dn = np.linalg.norm(d, axis=-1)
    # world -> scene frame; tilt_rad != 0 additionally pitches the scene about X through the table
    # centre (top / bottom views of scripts/alignment.sh:118-119)
    rinv = _rot_y(-view_angle_rad) @ _rot_x(-tilt_rad)
    ds = d @ rinv.T
    o = -TABLE_CENTRE_MM @ rinv.T
    z = np.full(d.shape[0], 250.0)
    hit = np.zeros(d.shape[0], dtype=bool)
    active = np.arange(d.shape[0])
    for _ in range(400):
        if active.size == 0:
            break
        za = z[active]
        f = _scene_sdf(za[:, None] * ds[active] + o, backdrop)
        done = np.abs(f) < 0.01
        hit[active[done]] = True
        za = za + 0.5 * f / dn[active]
        z[active] = za
        keep = ~done & (za < depth_max_mm)
        active = active[keep]
    z = z.reshape(h, w)
    hit = hit.reshape(h, w) & (z > 0) & (z <= depth_max_mm)
    if noise and seed is not None:
        rng = np.random.default_rng(seed)
        z = z + rng.standard_normal(z.shape) * 1.5 * (z / 1000.0) ** 2
    zi = np.where(hit, np.rint(z), 0.0)
    return np.clip(zi, 0, 65535).astype(np.uint16)


def backproject(depth_mm: np.ndarray, intr: dict = KINECT_V1, scale: float = 1.0,
                unit_scale: float = 0.001) -> np.ndarray:
    """depth_capture.cpp:294-308: raster-order cloud of the valid pixels, coordinates
    rounded through '%.6f' (mm) as the capture tool's ASCII PLY does, then scaled."""
    h, w = depth_mm.shape
    fx, fy, cx, cy = intr["fx"] * scale, intr["fy"] * scale, intr["cx"] * scale, intr["cy"] * scale
    ii, jj = np.nonzero(depth_mm)
    z = depth_mm[ii, jj].astype(np.float64)
    x = z * (jj - cx) / fx
    y = -z * (ii - cy) / fy
    pts = np.stack([x, y, -z], axis=-1)
    pts = np.round(pts, 6)
    return (pts * unit_scale).astype(np.float32)


def kinect_view(view: int, step_deg: float = 5.0, scale: float = 1.0, backdrop: str = "none",
                intr: dict = KINECT_V1, noise: bool = True, tilt_deg: float = 0.0) -> np.ndarray:
    """Cloud (n,3) float32, metres, of turntable view `view` (angle = view*step); tilt_deg pitches
    the scene toward the camera (top view > 0, bottom view < 0)."""
    depth = render_depth(np.deg2rad(view * step_deg), intr=intr, scale=scale, backdrop=backdrop,
                         seed=1000 + view + (7000 if tilt_deg else 0), noise=noise, tilt_rad=np.deg2rad(tilt_deg))
    return backproject(depth, intr=intr, scale=scale)


# BASELINE configs[3]: Kinect-v2 frustum sampled on a 1344 x 1113 grid (~1.5 M valid pixels with the
# full backdrop): the "super-resolution-densified" stress size of SURVEY 8(d)
KV2_SR_SCALE = 1344.0 / 512.0


def kinect_v2_sr_view(view: int, step_deg: float = 15.0, backdrop: str = "full", tilt_deg: float = 0.0) -> np.ndarray:
    return kinect_view(view, step_deg=step_deg, scale=KV2_SR_SCALE, backdrop=backdrop, intr=KINECT_V2, tilt_deg=tilt_deg)


def tilt_motion(tilt_deg: float, unit_scale: float = 0.001) -> np.ndarray:
    """Ground-truth 4x4 taking a view rendered with `tilt_deg` onto the untilted frame: the scene
    was pitched by +tilt about X through the table centre, so the cloud has to be pitched back."""
    R = _rot_x(np.deg2rad(-tilt_deg))
    c = TABLE_CENTRE_MM * unit_scale
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = c - R @ c
    return T


def turntable_motion(step_deg: float, unit_scale: float = 0.001) -> np.ndarray:
    """Ground-truth 4x4 taking view v+1's cloud onto view v's frame: rotation by -step
    about +Y through the table centre."""
    R = _rot_y(np.deg2rad(-step_deg))
    c = TABLE_CENTRE_MM * unit_scale
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = c - R @ c
    return T


def turntable_prior(view: int, step_deg: float, extra_deg: float = 0.0, unit_scale: float = 0.001) -> np.ndarray:
    """rotate_align.cpp:224-235-style prior taking view `view` to view 0's frame, with
    an optional per-step angular error (the tool adds 1.17 deg per step)."""
    R = _rot_y(np.deg2rad(-(view * step_deg + view * extra_deg)))
    c = TABLE_CENTRE_MM * unit_scale
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = c - R @ c
    return T


def apply_transform(T: np.ndarray, xyz: np.ndarray) -> np.ndarray:
    return (xyz.astype(np.float64) @ T[:3, :3].T + T[:3, 3]).astype(np.float32)


def surface_samples(n: int, seed: int, noise_mm: float = 0.5, unit_scale: float = 0.001) -> np.ndarray:
    """MVS-scale cloud (cfg5): n points on the bumpy sphere + isotropic noise, shuffled
    order by construction (random directions)."""
    rng = np.random.default_rng(seed)
    v = rng.standard_normal((n, 3))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    rad = _bumpy_radius(v[:, 0], v[:, 1], v[:, 2])
    p = v * rad[:, None] + rng.standard_normal((n, 3)) * noise_mm + TABLE_CENTRE_MM
    return (p * unit_scale).astype(np.float32)


def rigid(rx_deg: float, ry_deg: float, rz_deg: float, t) -> np.ndarray:
    ax, ay, az = np.deg2rad([rx_deg, ry_deg, rz_deg])
    Rx = np.array([[1, 0, 0], [0, np.cos(ax), -np.sin(ax)], [0, np.sin(ax), np.cos(ax)]])
    Rz = np.array([[np.cos(az), -np.sin(az), 0], [np.sin(az), np.cos(az), 0], [0, 0, 1]])
    T = np.eye(4)
    T[:3, :3] = Rz @ _rot_y(ay) @ Rx
    T[:3, 3] = np.asarray(t, dtype=np.float64)
    return T
