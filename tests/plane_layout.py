"""Test helper: the plane-record chunk layout of the training kernels in plain torch.

A chunk is 64 columns of a 128-row tile stored as [k-group of 8 columns][128 rows][8 fp16] (16 KiB,
csrc/es_program.h LayerProg::dump).  Geometry tiles hold 32 points x 4 streams: stream s of point (Q, p) of the tile
is tile row 32 Q + 8 s + p (csrc/es_mlp.cu, tangent mode)."""
import torch

CHUNK_BYTES = 16384


def to_chunks(m: torch.Tensor) -> torch.Tensor:
    """[tiles*128, 64*c] fp16 -> uint8 record [tiles][c][16 KiB]."""
    rows, cols = m.shape
    assert rows % 128 == 0 and cols % 64 == 0 and m.dtype == torch.float16
    t, c = rows // 128, cols // 64
    v = m.reshape(t, 128, c, 8, 8)            # tile, row, chunk, kgroup, elt
    v = v.permute(0, 2, 3, 1, 4).contiguous()  # tile, chunk, kgroup, row, elt
    return v.view(torch.uint8).reshape(t, c, CHUNK_BYTES)


def from_chunks(rec: torch.Tensor) -> torch.Tensor:
    """uint8 record [tiles][c][16 KiB] -> [tiles*128, 64*c] fp16."""
    t, c, _ = rec.shape
    v = rec.contiguous().view(torch.float16).reshape(t, c, 8, 128, 8)
    return v.permute(0, 3, 1, 2, 4).reshape(t * 128, c * 64)


def geom_row_of(point: torch.Tensor, stream: int) -> torch.Tensor:
    """global plane row of (point, stream) in a geometry record."""
    tile, r = point // 32, point % 32
    return tile * 128 + 32 * (r // 8) + 8 * stream + (r % 8)
