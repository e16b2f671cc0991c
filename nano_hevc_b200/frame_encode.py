"""CLI-faithful frame coder on the GPU: ``encode_frame_intra`` of
``nano_hevc/__main__.py:142-189`` -- per block DC vs planar by residual energy (DC wins ties,
``:173``), neighbours from the SOURCE plane (N samples per side, 128 at frame edges,
``block.py:38-55``), top_right = top[-1], bottom_left = left[-1] (``:166-167``), the clipped
*prediction* is written as the reconstruction (no transform / quant), chroma block size =
block_size // 2 with a minimum of 4 (``:156-158``).  Composition of the library's batched kernels;
every step runs on the device."""
from __future__ import annotations

import torch

from . import batched


def encode_plane_intra(plane: torch.Tensor, bs: int):
    """One plane.  Returns (recon_plane int16 (H, W), n_dc, n_planar)."""
    H, W = plane.shape
    top, left, _ = batched.gather_refs(plane, bs, bs, bs)          # K1, CLI convention: N samples per side
    t, l = top[:, 1:bs + 1].contiguous(), left[:, 1:bs + 1].contiguous()
    tr, bl = top[:, bs].contiguous(), left[:, bs].contiguous()      # top[-1], left[-1]
    orig = batched.plane_to_blocks(plane, bs)
    dc = batched.intra_dc_predict_batched(t, l, bs)
    pl = batched.intra_planar_predict_batched(t, l, tr, bl, bs)
    e_dc = batched.block_costs(orig, dc, outputs=("energy",))[2]    # residual_energy, metrics.py:46-48
    e_pl = batched.block_costs(orig, pl, outputs=("energy",))[2]
    use_dc = e_dc <= e_pl                                           # __main__.py:173
    best = torch.where(use_dc.view(-1, 1, 1), dc, pl)
    recon = batched.blocks_to_plane(batched.clip_to_pixel_range_batched(best), H, W)
    n_dc = int(use_dc.sum().item())
    return recon, n_dc, int(use_dc.numel()) - n_dc


def encode_frame_intra(y: torch.Tensor, u: torch.Tensor, v: torch.Tensor, block_size: int):
    """Y / U / V int16 CUDA planes -> ((recon_y, recon_u, recon_v), stats) with
    stats = {"dc", "planar", "blocks"} summed over the three planes, like the reference."""
    stats = {"dc": 0, "planar": 0, "blocks": 0}
    out = []
    for name, plane in (("Y", y), ("U", u), ("V", v)):
        bs = block_size if name == "Y" else max(block_size // 2, 4)
        recon, n_dc, n_pl = encode_plane_intra(plane, bs)
        stats["dc"] += n_dc
        stats["planar"] += n_pl
        stats["blocks"] += n_dc + n_pl
        out.append(recon)
    return tuple(out), stats


def y_psnr(orig_y: torch.Tensor, recon_y: torch.Tensor) -> float:
    """__main__.py:206-208: PSNR of the uint8 views of the luma planes."""
    o8 = (orig_y.to(torch.int32) & 0xff).to(torch.int16)   # astype(np.uint8) wrap-around
    r8 = (recon_y.to(torch.int32) & 0xff).to(torch.int16)
    sse = int(batched.sse_sad(o8, r8)[0].item())
    return batched.psnr_from_sse(sse, o8.numel())
