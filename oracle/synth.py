"""Deterministic KITTI-shaped synthetic stereo sequences.  TEST INFRASTRUCTURE ONLY.

No dataset is available offline, so the harness ray-casts a textured,
NON-planar scene (ground plane, two side walls, boxes hashed along the road;
a single plane is degenerate for fundamental-matrix estimation) with the
reference's camera: K = (718.856, 718.856, 607.1928, 185.2157), baseline 0.54 m
(include/visualSLAM.h:68,82-87 of the reference), 1241x376, u8, 1 channel.
The trajectory moves ~0.85 m/frame forward (cf. the reference's
src/trajectory.txt) with a yaw sinusoid <= 0.5 deg/frame.  Ground truth is known.

The same scene description is rendered on the GPU by the harness kernel in
ros_stereo_slam_b200/csrc/synth.cu for long bench sequences; the two renderers
agree to within float rounding (occasional +-1 grey level) and are never mixed
inside one comparison.
"""
import numpy as np

FX = 718.856
FY = 718.856
CX = 607.1928
CY = 185.2157
BASELINE = 0.54
WIDTH = 1241
HEIGHT = 376

GROUND_Y = 1.65
WALL_L = -7.5
WALL_R = 8.5
CELL = 8.0
FOG = 160.0
OCT_FREQ = (0.9, 2.3, 6.1, 17.0)   # cycles per metre
OCT_AMP = (1.0, 0.8, 0.65, 0.5)


def _hash_u32(ix, iy, seed):
    """Integer lattice hash -> uint32 (same arithmetic as synth.cu)."""
    m = np.uint64(0xFFFFFFFF)
    ix = np.asarray(ix).astype(np.int64).astype(np.uint64) & m
    iy = np.asarray(iy).astype(np.int64).astype(np.uint64) & m
    sd = np.uint64(int(seed) & 0xFFFFFFFF)
    h = (ix * np.uint64(374761393) + iy * np.uint64(668265263) + sd * np.uint64(2246822519)) & m
    h = ((h ^ (h >> np.uint64(13))) * np.uint64(1274126177)) & m
    h = h ^ (h >> np.uint64(16))
    return h.astype(np.uint32)


def _hash01(ix, iy, seed):
    return (_hash_u32(ix, iy, seed) >> np.uint32(8)).astype(np.float64) * (1.0 / 16777216.0)


def _value_noise(u, v, seed):
    """Smooth value noise in [-1, 1]."""
    fu = np.floor(u)
    fv = np.floor(v)
    iu = fu.astype(np.int64)
    iv = fv.astype(np.int64)
    a = u - fu
    b = v - fv
    a = a * a * (3 - 2 * a)
    b = b * b * (3 - 2 * b)
    n00 = _hash01(iu, iv, seed)
    n10 = _hash01(iu + 1, iv, seed)
    n01 = _hash01(iu, iv + 1, seed)
    n11 = _hash01(iu + 1, iv + 1, seed)
    n0 = n00 + (n10 - n00) * a
    n1 = n01 + (n11 - n01) * a
    return 2.0 * (n0 + (n1 - n0) * b) - 1.0


def texture(u, v, footprint, seed):
    """Multi-octave value noise with per-octave anti-alias fade.  footprint =
    metres per pixel on the surface."""
    acc = np.zeros_like(u)
    wsum = 0.0
    for k, (f, amp) in enumerate(zip(OCT_FREQ, OCT_AMP)):
        period_px = 1.0 / (f * np.maximum(footprint, 1e-9))
        fade = np.clip((period_px - 2.0) * 0.5, 0.0, 1.0)
        acc += amp * fade * _value_noise(u * f, v * f, seed * 4 + k)
        wsum += amp
    return acc / wsum


class Scene:
    """Procedural road scene + trajectory."""

    def __init__(self, seed=0, speed=0.85, yaw_amp_deg=4.0, yaw_rate=0.1):
        self.seed = int(seed)
        self.speed = float(speed)
        self.yaw_amp = np.deg2rad(yaw_amp_deg)
        self.yaw_rate = float(yaw_rate)

    # -- trajectory -------------------------------------------------------
    def yaw(self, i):
        return self.yaw_amp * np.sin(self.yaw_rate * i)

    def position(self, i):
        """Closed-form-free integration so every renderer sees identical poses."""
        x = 0.0
        z = 0.0
        for k in range(int(i)):
            y = self.yaw(k)
            x += self.speed * np.sin(y)
            z += self.speed * np.cos(y)
        return np.array([x, 0.0, z])

    def pose(self, i):
        """(R_wc 3x3, c 3) camera->world of the LEFT camera at frame i."""
        y = self.yaw(i)
        c, s = np.cos(y), np.sin(y)
        R = np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])
        return R, self.position(i)

    def poses(self, n):
        out = []
        x = z = 0.0
        for k in range(n):
            y = self.yaw(k)
            c, s = np.cos(y), np.sin(y)
            out.append((np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]]), np.array([x, 0.0, z])))
            x += self.speed * s
            z += self.speed * c
        return out

    # -- boxes ------------------------------------------------------------
    def box(self, cell):
        """Axis-aligned box of z-cell ``cell`` or None: (xmin,xmax,ymin,ymax,zmin,zmax,id)."""
        ci = np.array([cell], np.int64)
        r = [float(_hash01(ci, np.array([k], np.int64), self.seed + 101)[0]) for k in range(6)]
        if r[0] > 0.8:
            return None
        w = 0.8 + 1.7 * r[1]
        h = 0.8 + 2.2 * r[2]
        d = 0.8 + 2.2 * r[3]
        if r[4] < 0.5:
            xc = -6.0 + 3.0 * r[5]
        else:
            xc = 3.0 + 4.0 * r[5]
        zc = (cell + 0.5) * CELL
        return (xc - w / 2, xc + w / 2, GROUND_Y - h, GROUND_Y, zc - d / 2, zc + d / 2, cell)

    # -- rendering --------------------------------------------------------
    def render(self, i, eye="L", width=WIDTH, height=HEIGHT, return_depth=False):
        R, c = self.pose(i)
        return self.render_pose(R, c, eye, width, height, return_depth)

    def render_pose(self, R, c, eye="L", width=WIDTH, height=HEIGHT, return_depth=False):
        if eye == "R":
            c = c + R @ np.array([BASELINE, 0.0, 0.0])
        u = np.arange(width, dtype=np.float64)
        v = np.arange(height, dtype=np.float64)
        uu, vv = np.meshgrid(u, v)
        dc = np.stack([(uu - CX) / FX, (vv - CY) / FY, np.ones_like(uu)], -1)
        d = dc @ R.T                       # world ray directions (not normalised; z_cam = 1)
        ox, oy, oz = c
        dx, dy, dz = d[..., 0], d[..., 1], d[..., 2]
        big = 1e30
        t_best = np.full(uu.shape, big)
        tex_u = np.zeros_like(uu)
        tex_v = np.zeros_like(uu)
        cosn = np.ones_like(uu)
        sid = np.zeros(uu.shape, np.int64)

        def consider(t, tu, tv, cn, surf_id, valid):
            nonlocal t_best, tex_u, tex_v, cosn, sid
            m = valid & (t > 1e-3) & (t < t_best)
            t_best = np.where(m, t, t_best)
            tex_u = np.where(m, tu, tex_u)
            tex_v = np.where(m, tv, tex_v)
            cosn = np.where(m, cn, cosn)
            sid = np.where(m, surf_id, sid)

        with np.errstate(divide="ignore", invalid="ignore"):
            # ground y = GROUND_Y
            t = (GROUND_Y - oy) / dy
            consider(t, ox + t * dx, oz + t * dz, np.abs(dy), 1, dy > 1e-9)
            # walls
            t = (WALL_L - ox) / dx
            consider(t, oz + t * dz, oy + t * dy, np.abs(dx), 2, dx < -1e-9)
            t = (WALL_R - ox) / dx
            consider(t, oz + t * dz, oy + t * dy, np.abs(dx), 3, dx > 1e-9)
            # ceiling far above so that every ray hits something
            t = (-12.0 - oy) / dy
            consider(t, ox + t * dx, oz + t * dz, np.abs(dy), 4, dy < -1e-9)
            # boxes in the next cells
            c0 = int(np.floor(oz / CELL)) - 1
            for cell in range(c0, c0 + 20):
                bx = self.box(cell)
                if bx is None:
                    continue
                x0, x1, y0, y1, z0, z1, bid = bx
                tx0 = (x0 - ox) / dx
                tx1 = (x1 - ox) / dx
                ty0 = (y0 - oy) / dy
                ty1 = (y1 - oy) / dy
                tz0 = (z0 - oz) / dz
                tz1 = (z1 - oz) / dz
                tnx, tfx = np.minimum(tx0, tx1), np.maximum(tx0, tx1)
                tny, tfy = np.minimum(ty0, ty1), np.maximum(ty0, ty1)
                tnz, tfz = np.minimum(tz0, tz1), np.maximum(tz0, tz1)
                tn = np.maximum(np.maximum(tnx, tny), tnz)
                tf = np.minimum(np.minimum(tfx, tfy), tfz)
                hit = (tn <= tf) & (tn > 1e-3)
                px = ox + tn * dx
                py = oy + tn * dy
                pz = oz + tn * dz
                face_x = (tnx >= tny) & (tnx >= tnz)
                face_y = (~face_x) & (tny >= tnz)
                tu = np.where(face_x, pz, px)
                tv = np.where(face_y, pz, py)
                cn = np.where(face_x, np.abs(dx), np.where(face_y, np.abs(dy), np.abs(dz)))
                consider(tn, tu + 13.7 * (bid % 97), tv + 7.3 * (bid % 89), cn,
                         5 + (bid % 1000) * 3 + np.where(face_x, 0, np.where(face_y, 1, 2)), hit)

        tt = np.where(t_best < big, t_best, 1e4)
        # metres per pixel on the surface: range/f divided by the incidence cosine
        dn = np.sqrt(dx * dx + dy * dy + dz * dz)
        footprint = tt * dn / FX / np.maximum(cosn / dn, 0.05)
        val = np.zeros_like(uu)
        for s in np.unique(sid):
            m = sid == s
            val[m] = texture(tex_u[m], tex_v[m], footprint[m], self.seed * 131 + int(s))
        fog = np.exp(-(tt * dn) / FOG)
        img = 128.0 + 118.0 * val * fog
        out = np.clip(np.rint(img), 0, 255).astype(np.uint8)
        if return_depth:
            return out, tt   # z_cam depth == t because rays have z_cam = 1
        return out

    def stereo_pair(self, i, width=WIDTH, height=HEIGHT):
        return self.render(i, "L", width, height), self.render(i, "R", width, height)


def pnp_stress_case(n=20000, outlier_frac=0.5, sigma=0.3, seed=3):
    """direct 3D-2D synthetic correspondences."""
    import cv2
    rng = np.random.default_rng(seed)
    X = np.stack([rng.uniform(-20, 20, n), rng.uniform(-3, 3, n), rng.uniform(4, 60, n)], 1)
    rvec = np.array([0.01, -0.02, 0.005])
    tvec = np.array([0.05, -0.02, -0.8])
    K = np.array([[FX, 0, CX], [0, FY, CY], [0, 0, 1]])
    proj, _ = cv2.projectPoints(X.reshape(-1, 1, 3), rvec, tvec, K, np.zeros((4, 1)))
    xy = proj.reshape(-1, 2) + rng.normal(0, sigma, (n, 2))
    n_out = int(n * outlier_frac)
    out_idx = rng.permutation(n)[:n_out]
    xy[out_idx] = np.stack([rng.uniform(0, WIDTH, n_out), rng.uniform(0, HEIGHT, n_out)], 1)
    return X.astype(np.float32), xy.astype(np.float32), rvec, tvec, np.sort(out_idx)
