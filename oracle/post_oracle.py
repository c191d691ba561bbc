"""TEST INFRASTRUCTURE — torch restatement (any device, CPU in the tests) of what the reference's render() does
with the rasteriser's allmap: renderer/gaussian_renderer/__init__.py:110-147 and utils/point_utils.py:4-33, line
for line, minus the hard-coded `.cuda()` calls.  Checker only (tests/, bench.py); pinned against golden vectors
produced by the reference code itself (tools/make_golden_post.py -> tests/golden/surface_maps_*.npz)."""
import torch


def depths_to_points(view, depthmap):  # utils/point_utils.py:4-20
    dev = depthmap.device
    c2w = (view.world_view_transform.T).inverse()
    W, H = view.image_width, view.image_height
    dt = depthmap.dtype  # float32 in the reference (`.float()`); float64 when the tests want a yardstick
    ndc2pix = torch.tensor([[W / 2, 0, 0, (W) / 2], [0, H / 2, 0, (H) / 2], [0, 0, 0, 1]]).to(dt).to(dev).T
    projection_matrix = c2w.T @ view.full_proj_transform
    intrins = (projection_matrix @ ndc2pix)[:3, :3].T
    grid_x, grid_y = torch.meshgrid(torch.arange(W, device=dev).to(dt), torch.arange(H, device=dev).to(dt),
                                    indexing='xy')
    points = torch.stack([grid_x, grid_y, torch.ones_like(grid_x)], dim=-1).reshape(-1, 3)
    rays_d = points @ intrins.inverse().T @ c2w[:3, :3].T
    rays_o = c2w[:3, 3]
    return depthmap.reshape(-1, 1) * rays_d + rays_o


def depth_to_normal(view, depth):  # utils/point_utils.py:22-33
    points = depths_to_points(view, depth).reshape(*depth.shape[1:], 3)
    output = torch.zeros_like(points)
    dx = torch.cat([points[2:, 1:-1] - points[:-2, 1:-1]], dim=0)
    dy = torch.cat([points[1:-1, 2:] - points[1:-1, :-2]], dim=1)
    normal_map = torch.nn.functional.normalize(torch.cross(dx, dy, dim=-1), dim=-1)
    output[1:-1, 1:-1, :] = normal_map
    return output


def surface_maps(allmap, view, depth_ratio):  # renderer/gaussian_renderer/__init__.py:110-147
    render_alpha = allmap[1:2]
    render_normal = allmap[2:5]
    render_normal = (render_normal.permute(1, 2, 0) @ (view.world_view_transform[:3, :3].T)).permute(2, 0, 1)
    render_depth_median = torch.nan_to_num(allmap[5:6], 0, 0)
    render_depth_expected = torch.nan_to_num(allmap[0:1] / render_alpha, 0, 0)
    render_dist = allmap[6:7]
    surf_depth = render_depth_expected * (1 - depth_ratio) + depth_ratio * render_depth_median
    surf_normal = depth_to_normal(view, surf_depth).permute(2, 0, 1)
    surf_normal = surf_normal * render_alpha.detach()
    return {"rend_alpha": render_alpha, "rend_normal": render_normal, "rend_dist": render_dist,
            "surf_depth": surf_depth, "surf_normal": surf_normal}
