"""Restatement of kornia.geometry.camera.perspective.{project_points, unproject_points}.

kornia is a third-party dependency of the reference that is absent here and unpinned
(gecco-torch/pyproject.toml:25).  Published semantics (kornia >= 0.6):
  project_points(p, K)  = denormalize_points_with_intrinsics(convert_points_from_homogeneous(p), K)
  unproject_points(uv, d, K, normalize) = [normalize](to_homogeneous(normalize_with_intrinsics(uv, K))) * d
with convert_points_from_homogeneous using scale = where(|z| > 1e-8, 1 / (z + 1e-8), 1).
Call sites in the reference: gecco_torch/reparam.py:126,136 and gecco_torch/models/ray.py:74.
"""
import torch
import torch.nn.functional as F


def project_points(point_3d: torch.Tensor, camera_matrix: torch.Tensor) -> torch.Tensor:
    eps = 1e-8
    z = point_3d[..., -1:]
    scale = torch.where(z.abs() > eps, 1.0 / (z + eps), torch.ones_like(z))
    xy = scale * point_3d[..., :-1]
    fx = camera_matrix[..., 0, 0]
    fy = camera_matrix[..., 1, 1]
    cx = camera_matrix[..., 0, 2]
    cy = camera_matrix[..., 1, 2]
    u = xy[..., 0] * fx + cx
    v = xy[..., 1] * fy + cy
    return torch.stack([u, v], dim=-1)


def unproject_points(point_2d: torch.Tensor, depth: torch.Tensor, camera_matrix: torch.Tensor,
                     normalize: bool = False) -> torch.Tensor:
    fx = camera_matrix[..., 0, 0]
    fy = camera_matrix[..., 1, 1]
    cx = camera_matrix[..., 0, 2]
    cy = camera_matrix[..., 1, 2]
    x = (point_2d[..., 0] - cx) / fx
    y = (point_2d[..., 1] - cy) / fy
    xyz = torch.stack([x, y, torch.ones_like(x)], dim=-1)
    if normalize:
        xyz = F.normalize(xyz, dim=-1, p=2.0)
    return xyz * depth
