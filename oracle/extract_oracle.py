"""TEST INFRASTRUCTURE — torch restatement of the per-view epilogue of PartGS's extraction loop
(utils/mesh_utils.py:77-89 partmap_to_rgbmap, :113 F.normalize).  Only tests/ may import it.

Pinned by golden vectors of the reference's own GaussianExtractor.partmap_to_rgbmap / estimate_bounding_sphere run on
CPU with a given palette (tools/make_golden_extract.py -> tests/golden/extract_*.npz)."""
import torch


def partmap_to_rgbmap(part: torch.Tensor, palette: torch.Tensor) -> torch.Tensor:
    """part [S,H,W], palette [S+1, >=3] (what get_fancy_color(S+1) returns) -> [3,H,W]."""
    part = torch.clamp(part, 0.0, 1.0)
    background = torch.sum(part, dim=0)
    predicted = torch.argmax(part, dim=0)
    img = torch.zeros((predicted.shape[0], predicted.shape[1], 3), dtype=torch.float32, device=part.device)
    for cls in range(part.shape[0]):
        img[predicted == cls] = palette[cls, :3].to(part.device)
    img[background < 1e-1, :] = 1.
    return img.permute(2, 0, 1)


def unit_normals(rend_normal: torch.Tensor) -> torch.Tensor:
    return torch.nn.functional.normalize(rend_normal, dim=0)
