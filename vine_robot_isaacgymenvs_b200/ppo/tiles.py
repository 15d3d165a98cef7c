"""Conversions between row-major torch tensors and the 128-row UMMA row-blocked bf16 tiles the recurrent kernels keep
in HBM (csrc/vine_umma.cuh: offset(row, k) = (row/8)*(K/8)*128 + (k/8)*128 + (row%8)*16 + (k%8)*2 bytes, K = 128)."""
import torch


def to_tiles(x, width=128):
    """x [n, C] (C a multiple of ``width``) -> bf16 tensor [tiles, C/width, 128*width] laid out tile by tile."""
    n, C = x.shape
    tiles = (n + 127) // 128
    pad = torch.zeros(tiles * 128, C, device=x.device, dtype=torch.bfloat16)
    pad[:n] = x.to(torch.bfloat16)
    t = pad.view(tiles, 16, 8, C // width, width // 8, 8)          # tile, row-group, row8, part, col-group, col8
    return t.permute(0, 3, 1, 4, 2, 5).contiguous().view(tiles, C // width, 128 * width)


def from_tiles(t, n, width=128):
    """Inverse of to_tiles: -> f32 [n, C]."""
    tiles, parts, _ = t.shape
    x = t.view(tiles, parts, 16, width // 8, 8, 8).permute(0, 2, 4, 1, 3, 5).contiguous().view(tiles * 128, parts * width)
    return x[:n].float()
