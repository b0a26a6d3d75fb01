"""Synthetic Gaussians and cameras of the shapes named in BASELINE.json / SURVEY.md section 8d.

Camera matrices follow the reference conventions: ``world_view_transform`` and
``full_proj_transform`` are the TRANSPOSED (row-vector) 4x4 tensors built in
scene/cameras.py:71-78 from utils/graphics_utils.py:38-74 (znear 0.01, zfar 100).
Everything is generated on the CPU from ``torch.Generator().manual_seed(seed)``.
"""
import math
from dataclasses import dataclass

import torch

ZNEAR, ZFAR = 0.01, 100.0


@dataclass
class SynthCamera:
    image_width: int
    image_height: int
    FoVx: float
    FoVy: float
    world_view_transform: torch.Tensor   # [4,4], transposed
    full_proj_transform: torch.Tensor    # [4,4], transposed
    camera_center: torch.Tensor          # [3]
    bClusterOccur = None

    @property
    def tanfovx(self):
        return math.tan(self.FoVx * 0.5)

    @property
    def tanfovy(self):
        return math.tan(self.FoVy * 0.5)

    def to(self, device):
        return SynthCamera(self.image_width, self.image_height, self.FoVx, self.FoVy,
                           self.world_view_transform.to(device), self.full_proj_transform.to(device),
                           self.camera_center.to(device))


def projection_matrix(fovx: float, fovy: float, znear: float = ZNEAR, zfar: float = ZFAR) -> torch.Tensor:
    """Perspective matrix with z_sign = +1 (utils/graphics_utils.py:54-74), NOT transposed."""
    tx, ty = math.tan(fovx / 2), math.tan(fovy / 2)
    P = torch.zeros(4, 4, dtype=torch.float64)
    P[0, 0] = 1.0 / tx
    P[1, 1] = 1.0 / ty
    P[3, 2] = 1.0
    P[2, 2] = zfar / (zfar - znear)
    P[2, 3] = -(zfar * znear) / (zfar - znear)
    return P


def camera_from_RT(R, T, FoVx: float, FoVy: float, W: int, H: int) -> SynthCamera:
    """Camera from the (R, T) pair the reference's dataset readers store (R = camera-to-world rotation,
    T = world-to-camera translation), composed as scene/cameras.py:71-78 does: transposed world-view
    (utils/graphics_utils.py getWorld2View2 with trans = 0, scale = 1), transposed projection, their
    product and the camera centre."""
    R = torch.as_tensor(R, dtype=torch.float64)
    T = torch.as_tensor(T, dtype=torch.float64)
    V = torch.eye(4, dtype=torch.float64)
    V[:3, :3] = R.t()
    V[:3, 3] = T
    wvt = V.t().contiguous()
    full = (wvt @ projection_matrix(FoVx, FoVy).t()).contiguous()
    center = torch.linalg.inv(wvt)[3, :3]
    return SynthCamera(W, H, FoVx, FoVy, wvt.float(), full.float(), center.float())


def look_at(eye, target, W: int, H: int, fovx: float, up=(0.0, 0.0, 1.0)) -> SynthCamera:
    """COLMAP-style camera (x right, y down, z forward) at `eye` looking at `target`."""
    eye = torch.as_tensor(eye, dtype=torch.float64)
    target = torch.as_tensor(target, dtype=torch.float64)
    up = torch.as_tensor(up, dtype=torch.float64)
    fwd = target - eye
    fwd = fwd / fwd.norm()
    right = torch.linalg.cross(fwd, up)
    if right.norm() < 1e-8:
        right = torch.linalg.cross(fwd, torch.tensor([0.0, 1.0, 0.0], dtype=torch.float64))
    right = right / right.norm()
    down = torch.linalg.cross(fwd, right)
    Rwc = torch.stack([right, down, fwd], 0)          # world -> camera rotation
    V = torch.eye(4, dtype=torch.float64)
    V[:3, :3] = Rwc
    V[:3, 3] = -Rwc @ eye
    fovy = 2.0 * math.atan(math.tan(fovx / 2) * H / W)
    Pm = projection_matrix(fovx, fovy)
    wvt = V.t().contiguous()
    full = (wvt @ Pm.t()).contiguous()
    return SynthCamera(W, H, fovx, fovy, wvt.float(), full.float(), eye.float())


def orbit_cameras(n: int, radius: float, W: int, H: int, fovx: float, height: float = 0.0,
                  target=(0.0, 0.0, 0.0), phase: float = 0.0):
    cams = []
    for i in range(n):
        a = phase + 2 * math.pi * i / max(n, 1)
        eye = (target[0] + radius * math.cos(a), target[1] + radius * math.sin(a), target[2] + height)
        cams.append(look_at(eye, target, W, H, fovx))
    return cams


def make_gaussians(P: int, kind: str = "blender", seed: int = 0, sh_degree: int = 3, feat_dim: int = 6,
                   scale_mult: float = 0.6):
    """Returns a dict of ACTIVATED parameters in the layouts the rasterizer consumes
    (what GaussianModel's getters return: scene/gaussian_model.py:122-169)."""
    g = torch.Generator().manual_seed(seed)
    if kind == "blender":            # U(-1.3, 1.3)^3, scene/dataset_readers.py:346
        xyz = (torch.rand(P, 3, generator=g) * 2 - 1) * 1.3
        vol = 2.6 ** 3
    elif kind == "room":             # ScanNet-like 8 x 6 x 3 m room centred at the origin
        ext = torch.tensor([8.0, 6.0, 3.0])
        xyz = (torch.rand(P, 3, generator=g) - 0.5) * ext
        vol = 8.0 * 6.0 * 3.0
    elif kind == "lerf":             # 70 % N(0, 1.5^2) core + 30 % shell out to radius 20
        n_core = int(P * 0.7)
        core = torch.randn(n_core, 3, generator=g) * 1.5
        d = torch.randn(P - n_core, 3, generator=g)
        d = d / d.norm(dim=1, keepdim=True)
        r = 4.0 + torch.rand(P - n_core, 1, generator=g) * 16.0
        xyz = torch.cat([core, d * r], 0)
        vol = 4.0 / 3.0 * math.pi * 4.5 ** 3
    else:
        raise ValueError(kind)
    mean_nn = (vol / P) ** (1.0 / 3.0)
    log_s = math.log(scale_mult * mean_nn) + 0.4 * torch.randn(P, 3, generator=g)
    scales = torch.exp(log_s)
    if kind == "lerf":               # shell Gaussians are sparser -> proportionally larger
        n_core = int(P * 0.7)
        scales[n_core:] *= 6.0
    rot = torch.randn(P, 4, generator=g)
    rot = rot / rot.norm(dim=1, keepdim=True)
    opacity = torch.sigmoid(2.0 * torch.randn(P, 1, generator=g))
    M = (3 + 1) ** 2
    shs = torch.zeros(P, M, 3)
    shs[:, 0] = torch.randn(P, 3, generator=g) * 0.25 / 0.2821
    shs[:, 1:] = torch.randn(P, M - 1, 3, generator=g) * 0.05
    ins_feat = torch.rand(P, feat_dim, generator=g)
    return dict(means3D=xyz.contiguous(), scales=scales.contiguous(), rotations=rot.contiguous(),
                opacities=opacity.contiguous(), shs=shs.contiguous(), ins_feat=ins_feat.contiguous(),
                sh_degree=sh_degree)


SCENES = {
    # name: (kind, P, W, H, fovx, camera radius, camera height, scale_mult)
    # scale_mult (sigma as a fraction of the mean nearest-neighbour distance) is tuned so that the
    # realised duplicates/visible-Gaussian ratio is what trained 3DGS scenes show (7-17 tiles per
    # Gaussian): N ~ 2.1 M (blender 300k), 5.9 M (scannet 1M), 8.2 M (lerf 1M @1080p = SURVEY 8d's
    # worked example).
    "plumbing_10k_256": ("blender", 10_000, 256, 256, 0.69, 4.0, 1.0, 0.6),
    "blender_300k_800": ("blender", 300_000, 800, 800, 0.69, 4.0, 1.0, 0.3),
    "scannet_1m_1296x968": ("room", 1_000_000, 1296, 968, 2 * math.atan(1296 / (2 * 1170.0)), 2.5, 0.3, 0.3),
    "lerf_1m_1080p": ("lerf", 1_000_000, 1920, 1080, 1.0, 5.0, 1.0, 0.2),
    "lerf_3m_1080p": ("lerf", 3_000_000, 1920, 1080, 1.0, 5.0, 1.0, 0.2),
}


def make_scene(name: str, n_views: int = 8, seed: int = 0, P: int = None):
    kind, P0, W, H, fovx, rad, height, scale_mult = SCENES[name]
    gs = make_gaussians(P or P0, kind, seed, scale_mult=scale_mult)
    cams = orbit_cameras(n_views, rad, W, H, fovx, height)
    return gs, cams


class SynthModel:
    """The part of GaussianModel that render() consumes, on synthetic parameters: the PARAMETER tensors as
    scene/gaussian_model.py:66-74 holds them (`_xyz, _scaling` (log), `_rotation` (unnormalised), `_opacity`
    (logit), `_features_dc`, `_features_rest`, `_ins_feat`) and the getters of :122-169.  Used by the tests
    and the Stage-1 leg of bench.py; stage0=True makes every parameter trainable, otherwise only `_ins_feat`
    (OpenGaussian detaches the geometry from stage 1 on, train.py:431-436)."""

    def __init__(self, gs, device, stage0: bool = False):
        t = lambda x: x.to(device)  # noqa: E731
        self._xyz = t(gs["means3D"]).requires_grad_(stage0)
        self._scaling = t(torch.log(gs["scales"])).requires_grad_(stage0)
        self._rotation = t(gs["rotations"] * 1.3).requires_grad_(stage0)
        self._opacity = t(torch.logit(gs["opacities"].clamp(1e-4, 1 - 1e-4))).requires_grad_(stage0)
        self._features_dc = t(gs["shs"][:, :1].contiguous()).requires_grad_(stage0)
        self._features_rest = t(gs["shs"][:, 1:].contiguous()).requires_grad_(stage0)
        self._ins_feat = t(gs["ins_feat"] * 2 - 1).requires_grad_(True)
        self._ins_feat_q = None
        self.active_sh_degree = int(gs.get("sh_degree", 3))
        self.max_sh_degree = 3

    def parameters(self):
        return [self._xyz, self._scaling, self._rotation, self._opacity, self._features_dc, self._features_rest,
                self._ins_feat]

    get_xyz = property(lambda s: s._xyz)
    get_scaling = property(lambda s: torch.exp(s._scaling))
    get_rotation = property(lambda s: torch.nn.functional.normalize(s._rotation))
    get_opacity = property(lambda s: torch.sigmoid(s._opacity))
    get_features = property(lambda s: torch.cat((s._features_dc, s._features_rest), dim=1))

    def get_ins_feat(self, origin=False):
        f = self._ins_feat if (origin or self._ins_feat_q is None) else self._ins_feat_q
        return torch.nn.functional.normalize(f, dim=1)


def sam_like_ids(M: int, H: int, W: int, seed: int = 0) -> torch.Tensor:
    """[H,W] int64 in [0, M): compact regions (rectangles painted in order), id 0 = the rest of the image."""
    g = torch.Generator().manual_seed(seed)
    ids = torch.zeros(H, W, dtype=torch.int64)
    for m in range(1, M):
        y0, x0 = int(torch.randint(0, H - 8, (1,), generator=g)), int(torch.randint(0, W - 8, (1,), generator=g))
        h, w = int(torch.randint(8, H // 3, (1,), generator=g)), int(torch.randint(8, W // 3, (1,), generator=g))
        ids[y0:y0 + h, x0:x0 + w] = m
    return ids


def sam_like_masks(M: int, H: int, W: int, seed: int = 0) -> torch.Tensor:
    """[M,H,W] bool: the one-hot expansion of sam_like_ids (a partition of the image)."""
    ids = sam_like_ids(M, H, W, seed)
    return torch.stack([(ids == m) for m in range(M)])


def sam_like_id_map(M: int, H: int, W: int, seed: int = 0) -> torch.Tensor:
    """[4,H,W] int32 in the dataset's layout (the `original_sam_mask` of a view, utils/opengs_utlis.py:125-148): level 0
    holds sam_like_ids (ids 0..M-1), the other levels coarser partitions whose ids continue after the previous level's
    maximum.  get_SAM_mask_and_feat(level=0) of it returns exactly sam_like_masks(M, H, W, seed)."""
    l0 = sam_like_ids(M, H, W, seed)
    levels, base = [l0], M
    for k in (2, 3, 4):
        lv = sam_like_ids(max(M // k, 2), H, W, seed + k) + base
        levels.append(lv)
        base = int(lv.max()) + 1
    return torch.stack(levels).to(torch.int32)


def blob_labels(gs, n_blobs: int, seed: int = 0) -> torch.Tensor:
    """A spatial partition of the Gaussians into n_blobs Voronoi cells of randomly chosen Gaussians: [P] int64."""
    g = torch.Generator().manual_seed(seed)
    centres = gs["means3D"][torch.randperm(gs["means3D"].shape[0], generator=g)[:n_blobs]]
    return torch.cdist(gs["means3D"], centres).argmin(1)


def blob_view_masks(cam, pc, labels: torch.Tensor, n_blobs: int, min_pixels: int = 50) -> torch.Tensor:
    """View-consistent synthetic "SAM" masks for a camera: [M,H,W] bool, mask m = pixels whose accumulated blending
    weight is dominated by blob m (the blobs are rendered one-hot, 13 channels per pass, through the rasterizer);
    masks smaller than min_pixels are dropped (get_SAM_mask_and_feat's filter_th, utils/opengs_utlis.py:125-182)."""
    from .rasterizer import GaussianRasterizationSettings, GaussianRasterizer
    dev = labels.device
    H, W = cam.image_height, cam.image_width
    rs = GaussianRasterizationSettings(H, W, cam.tanfovx, cam.tanfovy, torch.zeros(3, device=dev), 1.0,
                                       cam.world_view_transform, cam.full_proj_transform, 0, cam.camera_center, False, False)
    rast = GaussianRasterizer(rs)
    onehot = torch.nn.functional.one_hot(labels, n_blobs).float()
    maps = []
    with torch.no_grad():
        for c0 in range(0, n_blobs, 13):
            extra = onehot[:, c0:c0 + 13].contiguous()
            out = rast(means3D=pc.get_xyz, means2D=torch.zeros_like(pc.get_xyz), opacities=pc.get_opacity,
                       colors_precomp=torch.zeros(labels.shape[0], 3, device=dev), scales=pc.get_scaling,
                       rotations=pc.get_rotation, extra_feats=extra)
            maps.append(out[4])
    votes = torch.cat(maps, 0)
    ids = votes.argmax(0)
    covered = votes.sum(0) > 0.5
    masks = torch.stack([(ids == k) & covered for k in range(n_blobs)])
    return masks[masks.flatten(1).sum(1) > min_pixels]
