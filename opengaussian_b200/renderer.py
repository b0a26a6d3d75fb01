"""``render()`` of OpenGaussian (reference gaussian_renderer/__init__.py:22-373) on the fused
B200 rasterizer: same signature, same 14-key return dict, same control flow for the coarse /
fine cluster passes -- but every group of reference passes that shares geometry is ONE launch of
the C-channel rasterizer:

  reference (Stage 1 / 2.1, :104-163)              here
  -------------------------------------------      -----------------------------------------
  pass 1  RGB (SH)                                 rescale_factor == 1 (always when
  pass 2  ins_feat[:, :3] as colours               rescale=False: Stage 1, render.py,
  pass 3  ins_feat[:, 3:6] as colours              pseudo-label construction):
  pass 4  SH again, only alpha kept (silhouette)       1 pass: RGB + 6 feature channels + depth + alpha
                                                   rescale_factor != 1 (Stage 2, p = 0.5):
                                                       RGB pass + 1 pass (6 feature channels + alpha)
  cluster / leaf passes: 2 x 3 channels (:203-225, :327-345)   1 pass with 6 channels

The feature channels get the same background as the reference gives them (bg_color applied to
each 3-channel group, because the reference reuses one rasterizer for all passes).  The CPU RNG
draws (`torch.rand(1)` at :121,124) are reproduced so that training stays in lock-step.
``fused=False`` falls back to the reference's literal pass structure (used by the parity test).
"""
import math

import torch

from .rasterizer import GaussianRasterizationSettings, GaussianRasterizer, pin_for_capture
from .rasterizer import view_cache as _view_cache


def _knn_mean_dists(x: torch.Tensor, K: int, chunk: int = 4096) -> torch.Tensor:
    """Squared distances to the K nearest neighbours (self included), like
    pytorch3d.ops.knn_points(x, x, K).dists[0] -- used only by post_process (:293-309)."""
    out = []
    for i in range(0, x.shape[0], chunk):
        d = torch.cdist(x[i:i + chunk], x) ** 2
        out.append(torch.topk(d, K, dim=1, largest=False).values)
    return torch.cat(out, 0)


_ZERO_SINKS = {}     # (P, device) -> (zeros [P,3], zeros [P,3]): the frozen-geometry stand-in for the screen-space gradient sink
_TILED_BG = {}       # device -> (signature of bg3, nb, tiled background, bg3)


def _zero_sink(xyz):
    """A zero `viewspace_points` tensor whose .grad is zeros, shared by every frozen-geometry render() of this model size
    (two 12 MB fills per call at 1 M Gaussians otherwise).  Nothing writes to it: the densification statistics that read
    it (train.py:598) only run while the positions train, and then render() builds a real gradient sink."""
    key = (xyz.shape[0], xyz.device, xyz.dtype)
    hit = _ZERO_SINKS.get(key)
    if hit is None:
        _ZERO_SINKS.clear()                      # one model size at a time
        hit = (torch.zeros_like(xyz), torch.zeros_like(xyz))
        hit[0].grad = hit[1]
        _ZERO_SINKS[key] = hit
    pin_for_capture(xyz.device, *hit)
    return hit[0]


def _tiled_bg(bg3, nb):
    """bg_color repeated over the feature channels (the reference reuses one rasterizer, hence one background, for all
    its 3-channel passes); cached per device on the background tensor's storage and version counter."""
    sig = (bg3.data_ptr(), bg3._version, nb)
    hit = _TILED_BG.get(bg3.device)
    if hit is None or hit[0] != sig:
        hit = (sig, torch.cat([bg3] * ((nb + 2) // 3))[:nb].contiguous(), bg3)     # keeps bg3 alive: the address stays its own
        _TILED_BG[bg3.device] = hit
    pin_for_capture(bg3.device, hit[1])
    return hit[1]


def _feat_pass(rasterizer, bg3, means3D, means2D, opacity, scales, rotations, cov3D_precomp, feat, shs=None,
               fused=True):
    """One pass compositing all channels of `feat` ([n,3] or [n,6]) -> (image [F,H,W], alpha [1,H,W]).
    With shs given (seg_rgb) the RGB image is rendered and duplicated like the reference does.
    fused=False: the reference's literal two 3-channel passes."""
    if shs is None and not fused and feat.shape[-1] > 3:
        img, _, _, alpha = rasterizer(means3D=means3D, means2D=means2D, shs=None, colors_precomp=feat[:, :3],
                                      opacities=opacity, scales=scales, rotations=rotations,
                                      cov3D_precomp=cov3D_precomp)
        img2, _, _, alpha = rasterizer(means3D=means3D, means2D=means2D, shs=None, colors_precomp=feat[:, 3:],
                                       opacities=opacity, scales=scales, rotations=rotations,
                                       cov3D_precomp=cov3D_precomp)
        return torch.cat((img, img2), dim=0), alpha
    if shs is not None:
        img, _, _, alpha = rasterizer(means3D=means3D, means2D=means2D, shs=shs, colors_precomp=None,
                                      opacities=opacity, scales=scales, rotations=rotations,
                                      cov3D_precomp=cov3D_precomp)
        return torch.cat((img, img), dim=0), alpha
    if feat.shape[-1] > 3:
        nb = feat.shape[-1] - 3
        ebg = torch.cat([bg3] * ((nb + 2) // 3))[:nb]
        img, _, _, alpha, img2 = rasterizer(means3D=means3D, means2D=means2D, shs=None, colors_precomp=feat[:, :3],
                                            opacities=opacity, scales=scales, rotations=rotations,
                                            cov3D_precomp=cov3D_precomp, extra_feats=feat[:, 3:], extra_bg=ebg)
        return torch.cat((img, img2), dim=0), alpha
    img, _, _, alpha = rasterizer(means3D=means3D, means2D=means2D, shs=None, colors_precomp=feat,
                                  opacities=opacity, scales=scales, rotations=rotations, cov3D_precomp=cov3D_precomp)
    return img, alpha


def render(viewpoint_camera, pc, pipe, bg_color: torch.Tensor, iteration,
           scaling_modifier=1.0, override_color=None, visible_mask=None, mask_num=0,
           cluster_idx=None,       # per-point cluster id (coarse-level)
           leaf_cluster_idx=None,  # per-point cluster id (fine-level)
           rescale=True,           # re-scale (for enhance ins_feat)
           origin_feat=False,      # origin ins_feat (not quantized)
           render_feat_map=True,   # render image-level feat map
           render_color=True,      # render rgb image
           render_cluster=False,   # render cluster, stage 2.2
           better_vis=False,       # filter some points
           selected_root_id=None,  # coarse-level cluster id
           selected_leaf_id=None,  # fine-level cluster id (possibly more than one)
           pre_mask=None,
           seg_rgb=False,          # render cluster rgb, not feat
           post_process=False,     # post
           root_num=64, leaf_num=10,
           fused=True):
    """Render the scene.  Background tensor (bg_color) must be on GPU!
    ``pipe.view_cache = False`` keeps this call out of the frozen-geometry view cache (rasterizer.ViewCache)."""
    if getattr(pipe, "view_cache", True) is False and _view_cache.enabled:
        with _view_cache.disabled():
            return render(viewpoint_camera, pc, pipe, bg_color, iteration, scaling_modifier, override_color, visible_mask,
                          mask_num, cluster_idx, leaf_cluster_idx, rescale, origin_feat, render_feat_map, render_color,
                          render_cluster, better_vis, selected_root_id, selected_leaf_id, pre_mask, seg_rgb, post_process,
                          root_num, leaf_num, fused)
    xyz = pc.get_xyz
    # Screen-space gradient sink (reference :45): its .grad feeds the densification statistics (train.py:598,
    # scene/gaussian_model.py:512-514), which only exist while the geometry trains.  From stage 1 on OpenGaussian
    # detaches the geometry (train.py:431-436); asking the rasterizer for dL/dmeans2D there forces the FULL geometry
    # backward (blend moments + preprocess chain, ~0.45 ms per step at 1 M Gaussians) for a tensor nobody reads.  So the
    # sink tracks gradients only when the positions do (or when pipe.viewspace_grad = True); otherwise it is a plain
    # zero tensor whose .grad is zeros, so that code reading viewspace_points.grad keeps working.
    track_vs = bool(xyz.requires_grad or getattr(pipe, "viewspace_grad", False))
    if track_vs:
        screenspace_points = torch.zeros_like(xyz, dtype=xyz.dtype, requires_grad=True, device=xyz.device) + 0
        try:
            screenspace_points.retain_grad()
        except Exception:
            pass
    else:
        screenspace_points = _zero_sink(xyz)

    tanfovx = math.tan(viewpoint_camera.FoVx * 0.5)
    tanfovy = math.tan(viewpoint_camera.FoVy * 0.5)
    raster_settings = GaussianRasterizationSettings(
        image_height=int(viewpoint_camera.image_height),
        image_width=int(viewpoint_camera.image_width),
        tanfovx=tanfovx, tanfovy=tanfovy, bg=bg_color, scale_modifier=scaling_modifier,
        viewmatrix=viewpoint_camera.world_view_transform, projmatrix=viewpoint_camera.full_proj_transform,
        sh_degree=pc.active_sh_degree, campos=viewpoint_camera.camera_center, prefiltered=False,
        debug=pipe.debug)
    rasterizer = GaussianRasterizer(raster_settings=raster_settings)

    means3D = xyz
    means2D = screenspace_points

    # probabilistically rescale (same RNG draws as the reference, :121-124)
    prob = torch.rand(1)
    rescale_factor = None         # 1.0: the reference builds a device scalar here every call (:122, a host->device copy)
    rescaled = False
    if prob > 0.5 and rescale:
        rescale_factor = torch.rand(1).to(xyz.device)
        rescaled = True
    one_pass = fused and render_color and render_feat_map and not rescaled

    # Raw-parameter path: when this call is exactly one fused pass over the whole model,
    # hand the PARAMETERS to the rasterizer and let preprocess apply the getters of
    # scene/gaussian_model.py:122-169 (exp / normalize / sigmoid, no get_features cat, no ins_feat
    # normalize) -- forward and backward -- instead of ~10 elementwise kernels and a 192 B/Gaussian copy.
    needs_activated = (render_cluster and cluster_idx is not None) or \
        (leaf_cluster_idx is not None and leaf_cluster_idx.numel() > 0)
    raw_names = ("_xyz", "_opacity", "_scaling", "_rotation", "_features_dc", "_features_rest", "_ins_feat")
    use_raw = (one_pass and not needs_activated and override_color is None and not pipe.compute_cov3D_python
               and not pipe.convert_SHs_python and all(hasattr(pc, n) for n in raw_names)
               and getattr(pipe, "raw_parameter_path", True))

    opacity = scales = rotations = cov3D_precomp = shs = colors_precomp = None
    if not use_raw:
        opacity = pc.get_opacity
        if pipe.compute_cov3D_python:
            cov3D_precomp = pc.get_covariance(scaling_modifier)
        else:
            scales = pc.get_scaling
            rotations = pc.get_rotation
        if override_color is None:
            if pipe.convert_SHs_python:
                from .sh import eval_sh
                shs_view = pc.get_features.transpose(1, 2).view(-1, 3, (pc.max_sh_degree + 1) ** 2)
                dir_pp = (pc.get_xyz - viewpoint_camera.camera_center.repeat(pc.get_features.shape[0], 1))
                dir_pp_normalized = dir_pp / dir_pp.norm(dim=1, keepdim=True)
                sh2rgb = eval_sh(pc.active_sh_degree, shs_view, dir_pp_normalized)
                colors_precomp = torch.clamp_min(sh2rgb + 0.5, 0.0)
            else:
                shs = pc.get_features
        else:
            colors_precomp = override_color

    def sc(s):
        return None if s is None else (s * rescale_factor if rescaled else s)

    rendered_image = radii = rendered_depth = rendered_alpha = None
    rendered_ins_feat = silhouette = None
    bg3 = bg_color.reshape(-1).float()
    if use_raw:
        q = getattr(pc, "_ins_feat_q", None)
        raw_feat = pc._ins_feat if (origin_feat or q is None or len(q) == 0) else q
        nb = raw_feat.shape[-1]
        ebg = _tiled_bg(bg3, nb)
        rendered_image, radii, rendered_depth, rendered_alpha, rendered_ins_feat = rasterizer.forward_raw(
            means3D, means2D, pc._opacity, pc._features_dc, pc._features_rest, pc._scaling, pc._rotation,
            raw_ins_feat=raw_feat, extra_bg=ebg)
        silhouette = rendered_alpha
    elif one_pass:
        # ONE launch: RGB + feature channels + depth + alpha; silhouette == alpha (same geometry)
        ins_feat = (pc.get_ins_feat(origin=origin_feat) + 1) / 2
        nb = ins_feat.shape[-1]
        ebg = _tiled_bg(bg3, nb)
        rendered_image, radii, rendered_depth, rendered_alpha, rendered_ins_feat = rasterizer(
            means3D=means3D, means2D=means2D, shs=shs, colors_precomp=colors_precomp, opacities=opacity,
            scales=scales, rotations=rotations, cov3D_precomp=cov3D_precomp, extra_feats=ins_feat, extra_bg=ebg)
        silhouette = rendered_alpha
    else:
        if render_color:
            rendered_image, radii, rendered_depth, rendered_alpha = rasterizer(
                means3D=means3D, means2D=means2D, shs=shs, colors_precomp=colors_precomp, opacities=opacity,
                scales=scales, rotations=rotations, cov3D_precomp=cov3D_precomp)
        if render_feat_map:
            ins_feat = (pc.get_ins_feat(origin=origin_feat) + 1) / 2
            if fused:
                rendered_ins_feat, silhouette = _feat_pass(rasterizer, bg3, means3D, means2D, opacity, sc(scales),
                                                           rotations, cov3D_precomp, ins_feat)
            else:   # the reference's literal 3 passes (:129-163)
                rendered_ins_feat, _, _, _ = rasterizer(
                    means3D=means3D, means2D=means2D, shs=None, colors_precomp=ins_feat[:, :3], opacities=opacity,
                    scales=sc(scales), rotations=rotations, cov3D_precomp=cov3D_precomp)
                if ins_feat.shape[-1] > 3:
                    rendered_ins_feat2, _, _, _ = rasterizer(
                        means3D=means3D, means2D=means2D, shs=None, colors_precomp=ins_feat[:, 3:6],
                        opacities=opacity, scales=sc(scales), rotations=rotations, cov3D_precomp=cov3D_precomp)
                    rendered_ins_feat = torch.cat((rendered_ins_feat, rendered_ins_feat2), dim=0)
                _, _, _, silhouette = rasterizer(
                    means3D=means3D, means2D=means2D, shs=shs, colors_precomp=colors_precomp, opacities=opacity,
                    scales=sc(scales), rotations=rotations, cov3D_precomp=cov3D_precomp)

    def filt(t, idx):
        return None if t is None else t[idx]

    # ---- [Stage 2.2] coarse cluster-level feature maps (:173-236) ----
    viewed_pts = radii > 0
    if cluster_idx is not None:
        num_cluster = cluster_idx.max() + 1
        cluster_occur = torch.zeros(num_cluster).to(torch.bool)
    else:
        cluster_occur = None
    if render_cluster and cluster_idx is not None and viewed_pts.sum() != 0:
        ins_feat = (pc.get_ins_feat(origin=origin_feat) + 1) / 2
        rendered_clusters = []
        rendered_cluster_silhouettes = []
        scale_filter = (scales < 0.5).all(dim=1)
        for idx in range(num_cluster):
            if not better_vis and idx != selected_root_id:
                continue
            if viewpoint_camera.bClusterOccur is not None and viewpoint_camera.bClusterOccur[idx] == False:  # noqa: E712
                continue
            filter_idx = cluster_idx == idx
            filter_idx = filter_idx & viewed_pts
            if better_vis:
                filter_idx = filter_idx & scale_filter
                if filter_idx.sum() < 100:
                    continue
            rendered_cluster, cluster_silhouette = _feat_pass(
                rasterizer, bg3, means3D[filter_idx], means2D[filter_idx], opacity[filter_idx],
                sc(scales[filter_idx]), rotations[filter_idx], filt(cov3D_precomp, filter_idx), ins_feat[filter_idx],
                fused=fused)
            if cluster_silhouette.max() > 0.8:
                cluster_occur[idx] = True
                rendered_clusters.append(rendered_cluster)
                rendered_cluster_silhouettes.append(cluster_silhouette)
        if len(rendered_cluster_silhouettes) != 0:
            rendered_cluster_silhouettes = torch.vstack(rendered_cluster_silhouettes)
    else:
        rendered_clusters, rendered_cluster_silhouettes = None, None

    # ---- [Stage 2.2 & 3] fine cluster-level feature maps (:245-356) ----
    if leaf_cluster_idx is not None and leaf_cluster_idx.numel() > 0:
        ins_feat = (pc.get_ins_feat(origin=origin_feat) + 1) / 2
        scale_filter = (scales < 0.1).all(dim=1)
        rendered_leaf_clusters = []
        rendered_leaf_cluster_silhouettes = []
        occured_leaf_id = []
        if selected_leaf_id is None:
            if selected_root_id is not None:
                start_leaf = selected_root_id * leaf_num
                end_leaf = start_leaf + leaf_num
            else:
                start_leaf = 0
                end_leaf = root_num * leaf_num
            lerf_range = range(start_leaf, end_leaf)
        else:
            lerf_range = selected_leaf_id.tolist()
        for _, leaf_idx in enumerate(lerf_range):
            if viewpoint_camera.bClusterOccur is not None and \
                    viewpoint_camera.bClusterOccur[selected_root_id] == False:  # noqa: E712
                continue
            if selected_leaf_id is None:
                filter_idx = leaf_cluster_idx == leaf_idx
            else:
                filter_idx = (leaf_cluster_idx.unsqueeze(1) == selected_leaf_id).any(dim=1)
            if pre_mask is not None:
                filter_idx = filter_idx & pre_mask
            filter_idx = filter_idx & viewed_pts
            if better_vis:
                filter_idx = filter_idx & scale_filter
                if filter_idx.sum() < 100:
                    continue
            max_time = 5
            if post_process and max_time > 0:
                nearest_k_distance = _knn_mean_dists(means3D[filter_idx].detach(), int(filter_idx.sum() ** 0.5))
                mean_d, std_d = nearest_k_distance.mean(), nearest_k_distance.std()
                mask = nearest_k_distance.mean(dim=-1) < mean_d + std_d
                mask = mask.squeeze()
                filter_idx[filter_idx != 0] = mask
                max_time -= 1
            if filter_idx.sum() < 10:
                continue
            occured_leaf_id.append(leaf_idx)
            leaf_img, leaf_sil = _feat_pass(
                rasterizer, bg3, means3D[filter_idx], means2D[filter_idx], opacity[filter_idx], scales[filter_idx],
                rotations[filter_idx], filt(cov3D_precomp, filter_idx), ins_feat[filter_idx],
                shs=shs[filter_idx] if seg_rgb else None, fused=fused)
            rendered_leaf_clusters.append(leaf_img)
            rendered_leaf_cluster_silhouettes.append(leaf_sil)
            if selected_leaf_id is not None and len(rendered_leaf_clusters) > 0:
                break
        if len(rendered_leaf_cluster_silhouettes) != 0:
            rendered_leaf_cluster_silhouettes = torch.vstack(rendered_leaf_cluster_silhouettes)
    else:
        rendered_leaf_clusters = None
        rendered_leaf_cluster_silhouettes = None
        occured_leaf_id = None

    return {"render": rendered_image,
            "alpha": rendered_alpha,
            "depth": rendered_depth,
            "silhouette": silhouette,
            "ins_feat": rendered_ins_feat,
            "cluster_imgs": rendered_clusters,
            "cluster_silhouettes": rendered_cluster_silhouettes,
            "leaf_clusters_imgs": rendered_leaf_clusters,
            "leaf_cluster_silhouettes": rendered_leaf_cluster_silhouettes,
            "occured_leaf_id": occured_leaf_id,
            "cluster_occur": cluster_occur,
            "viewspace_points": screenspace_points,
            "visibility_filter": viewed_pts,
            "radii": radii}
