"""Oracle: CogVideoX 3D causal VAE (restatement of diffusers `AutoencoderKLCogVideoX`).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Follows, by upstream module path,
`diffusers/models/autoencoders/autoencoder_kl_cogvideox.py` (CogVideoXSafeConv3d,
CogVideoXCausalConv3d, CogVideoXSpatialNorm3D, CogVideoXResnetBlock3D, CogVideoXDownBlock3D,
CogVideoXMidBlock3D, CogVideoXUpBlock3D, CogVideoXEncoder3D, CogVideoXDecoder3D,
AutoencoderKLCogVideoX._encode/_decode/tiled_encode/tiled_decode/blend_v/blend_h),
`diffusers/models/downsampling.py::CogVideoXDownsample3D`,
`diffusers/models/upsampling.py::CogVideoXUpsample3D`,
`diffusers/models/autoencoders/vae.py::DiagonalGaussianDistribution`.
Reference call sites: /root/reference/inference_script.py:408-409 (encode + sample),
:500 (decode_latents), :643-645 (enable_slicing / enable_tiling).
Parameter names match the diffusers state-dict keys so real checkpoints load.
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import torch
import torch.nn as nn
import torch.nn.functional as F

VAE_CONFIG = dict(
    in_channels=3, out_channels=3, latent_channels=16,
    block_out_channels=(128, 256, 256, 512), layers_per_block=3,
    norm_eps=1e-6, norm_num_groups=32, temporal_compression_ratio=4,
    sample_height=480, sample_width=720, scaling_factor=0.7,
)


class CausalConv3d(nn.Module):
    """k=3: spatial zero pad 1 inside the conv, temporal pad = 2 cached frames or frame 0 twice."""

    def __init__(self, cin, cout, kernel_size):
        super().__init__()
        self.kt = kernel_size
        p = (kernel_size - 1) // 2
        self.conv = nn.Conv3d(cin, cout, kernel_size, stride=1, padding=(0, p, p))

    def forward(self, x, conv_cache=None):
        if self.kt > 1:
            cached = [conv_cache] if conv_cache is not None else [x[:, :, :1]] * (self.kt - 1)
            x = torch.cat(cached + [x], dim=2)
        new_cache = x[:, :, -self.kt + 1:].clone() if self.kt > 1 else None
        return self.conv(x), new_cache


class SpatialNorm3D(nn.Module):
    def __init__(self, f_channels, zq_channels, groups=32):
        super().__init__()
        self.norm_layer = nn.GroupNorm(groups, f_channels, eps=1e-6, affine=True)
        self.conv_y = CausalConv3d(zq_channels, f_channels, 1)
        self.conv_b = CausalConv3d(zq_channels, f_channels, 1)

    def forward(self, f, zq):
        if f.shape[2] > 1 and f.shape[2] % 2 == 1:
            z_first = F.interpolate(zq[:, :, :1], size=f[:, :, :1].shape[-3:])
            z_rest = F.interpolate(zq[:, :, 1:], size=f[:, :, 1:].shape[-3:])
            zq = torch.cat([z_first, z_rest], dim=2)
        else:
            zq = F.interpolate(zq, size=f.shape[-3:])
        y, _ = self.conv_y(zq)
        b, _ = self.conv_b(zq)
        return self.norm_layer(f) * y + b


class ResnetBlock3D(nn.Module):
    def __init__(self, cin, cout, groups=32, eps=1e-6, spatial_norm_dim=None):
        super().__init__()
        self.cin, self.cout = cin, cout
        self.spatial = spatial_norm_dim is not None
        if not self.spatial:
            self.norm1 = nn.GroupNorm(groups, cin, eps=eps)
            self.norm2 = nn.GroupNorm(groups, cout, eps=eps)
        else:
            self.norm1 = SpatialNorm3D(cin, spatial_norm_dim, groups)
            self.norm2 = SpatialNorm3D(cout, spatial_norm_dim, groups)
        self.conv1 = CausalConv3d(cin, cout, 3)
        self.conv2 = CausalConv3d(cout, cout, 3)
        if cin != cout:
            self.conv_shortcut = nn.Conv3d(cin, cout, 1)

    def forward(self, x, zq=None, conv_cache=None):
        cc = conv_cache or {}
        new = {}
        h = self.norm1(x, zq) if self.spatial else self.norm1(x)
        h = F.silu(h)
        h, new["conv1"] = self.conv1(h, cc.get("conv1"))
        h = self.norm2(h, zq) if self.spatial else self.norm2(h)
        h = F.silu(h)
        h, new["conv2"] = self.conv2(h, cc.get("conv2"))
        if self.cin != self.cout:
            x = self.conv_shortcut(x)
        return h + x, new


class Downsample3D(nn.Module):
    def __init__(self, c, compress_time):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, stride=2, padding=0)
        self.compress_time = compress_time

    def forward(self, x):
        if self.compress_time:
            b, c, f, h, w = x.shape
            x = x.permute(0, 3, 4, 1, 2).reshape(b * h * w, c, f)
            if x.shape[-1] % 2 == 1:
                first, rest = x[..., 0], x[..., 1:]
                if rest.shape[-1] > 0:
                    rest = F.avg_pool1d(rest, kernel_size=2, stride=2)
                x = torch.cat([first[..., None], rest], dim=-1)
            else:
                x = F.avg_pool1d(x, kernel_size=2, stride=2)
            x = x.reshape(b, h, w, c, x.shape[-1]).permute(0, 3, 4, 1, 2)
        x = F.pad(x, (0, 1, 0, 1), mode="constant", value=0)
        b, c, f, h, w = x.shape
        x = x.permute(0, 2, 1, 3, 4).reshape(b * f, c, h, w)
        x = self.conv(x)
        return x.reshape(b, f, x.shape[1], x.shape[2], x.shape[3]).permute(0, 2, 1, 3, 4)


class Upsample3D(nn.Module):
    def __init__(self, c, compress_time):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, stride=1, padding=1)
        self.compress_time = compress_time

    def forward(self, x):
        if self.compress_time:
            if x.shape[2] > 1 and x.shape[2] % 2 == 1:
                first = F.interpolate(x[:, :, 0], scale_factor=2.0)
                rest = F.interpolate(x[:, :, 1:], scale_factor=2.0)
                x = torch.cat([first[:, :, None], rest], dim=2)
            elif x.shape[2] > 1:
                x = F.interpolate(x, scale_factor=2.0)
            else:
                x = F.interpolate(x.squeeze(2), scale_factor=2.0)[:, :, None]
        else:
            b, c, f, h, w = x.shape
            x = x.permute(0, 2, 1, 3, 4).reshape(b * f, c, h, w)
            x = F.interpolate(x, scale_factor=2.0)
            x = x.reshape(b, f, c, *x.shape[2:]).permute(0, 2, 1, 3, 4)
        b, c, f, h, w = x.shape
        x = x.permute(0, 2, 1, 3, 4).reshape(b * f, c, h, w)
        x = self.conv(x)
        return x.reshape(b, f, *x.shape[1:]).permute(0, 2, 1, 3, 4)


class DownBlock3D(nn.Module):
    def __init__(self, cin, cout, num_layers, eps, groups, add_downsample, compress_time):
        super().__init__()
        self.resnets = nn.ModuleList(
            [ResnetBlock3D(cin if i == 0 else cout, cout, groups, eps) for i in range(num_layers)])
        self.downsamplers = nn.ModuleList([Downsample3D(cout, compress_time)]) if add_downsample else None

    def forward(self, x, zq=None, conv_cache=None):
        cc = conv_cache or {}
        new = {}
        for i, r in enumerate(self.resnets):
            x, new[f"resnet_{i}"] = r(x, zq, cc.get(f"resnet_{i}"))
        if self.downsamplers is not None:
            for d in self.downsamplers:
                x = d(x)
        return x, new


class MidBlock3D(nn.Module):
    def __init__(self, c, num_layers, eps, groups, spatial_norm_dim=None):
        super().__init__()
        self.resnets = nn.ModuleList(
            [ResnetBlock3D(c, c, groups, eps, spatial_norm_dim) for _ in range(num_layers)])

    def forward(self, x, zq=None, conv_cache=None):
        cc = conv_cache or {}
        new = {}
        for i, r in enumerate(self.resnets):
            x, new[f"resnet_{i}"] = r(x, zq, cc.get(f"resnet_{i}"))
        return x, new


class UpBlock3D(nn.Module):
    def __init__(self, cin, cout, num_layers, eps, groups, spatial_norm_dim, add_upsample, compress_time):
        super().__init__()
        self.resnets = nn.ModuleList(
            [ResnetBlock3D(cin if i == 0 else cout, cout, groups, eps, spatial_norm_dim)
             for i in range(num_layers)])
        self.upsamplers = nn.ModuleList([Upsample3D(cout, compress_time)]) if add_upsample else None

    def forward(self, x, zq=None, conv_cache=None):
        cc = conv_cache or {}
        new = {}
        for i, r in enumerate(self.resnets):
            x, new[f"resnet_{i}"] = r(x, zq, cc.get(f"resnet_{i}"))
        if self.upsamplers is not None:
            for u in self.upsamplers:
                x = u(x)
        return x, new


class Encoder3D(nn.Module):
    def __init__(self, in_channels, out_channels, block_out_channels, layers_per_block, eps, groups, tcr):
        super().__init__()
        tlevel = int(math.log2(tcr))
        self.conv_in = CausalConv3d(in_channels, block_out_channels[0], 3)
        self.down_blocks = nn.ModuleList()
        oc = block_out_channels[0]
        for i, c in enumerate(block_out_channels):
            ic, oc = oc, c
            final = i == len(block_out_channels) - 1
            self.down_blocks.append(DownBlock3D(ic, oc, layers_per_block, eps, groups,
                                                add_downsample=not final, compress_time=i < tlevel))
        self.mid_block = MidBlock3D(block_out_channels[-1], 2, eps, groups)
        self.norm_out = nn.GroupNorm(groups, block_out_channels[-1], eps=1e-6)
        self.conv_out = CausalConv3d(block_out_channels[-1], 2 * out_channels, 3)

    def forward(self, x, conv_cache=None):
        cc = conv_cache or {}
        new = {}
        h, new["conv_in"] = self.conv_in(x, cc.get("conv_in"))
        for i, blk in enumerate(self.down_blocks):
            h, new[f"down_block_{i}"] = blk(h, None, cc.get(f"down_block_{i}"))
        h, new["mid_block"] = self.mid_block(h, None, cc.get("mid_block"))
        h = F.silu(self.norm_out(h))
        h, new["conv_out"] = self.conv_out(h, cc.get("conv_out"))
        return h, new


class Decoder3D(nn.Module):
    def __init__(self, in_channels, out_channels, block_out_channels, layers_per_block, eps, groups, tcr):
        super().__init__()
        rev = list(reversed(block_out_channels))
        tlevel = int(math.log2(tcr))
        self.conv_in = CausalConv3d(in_channels, rev[0], 3)
        self.mid_block = MidBlock3D(rev[0], 2, eps, groups, spatial_norm_dim=in_channels)
        self.up_blocks = nn.ModuleList()
        oc = rev[0]
        for i, c in enumerate(rev):
            pc, oc = oc, c
            final = i == len(rev) - 1
            self.up_blocks.append(UpBlock3D(pc, oc, layers_per_block + 1, eps, groups, in_channels,
                                            add_upsample=not final, compress_time=i < tlevel))
        self.norm_out = SpatialNorm3D(rev[-1], in_channels, groups)
        self.conv_out = CausalConv3d(rev[-1], out_channels, 3)

    def forward(self, z, conv_cache=None):
        cc = conv_cache or {}
        new = {}
        h, new["conv_in"] = self.conv_in(z, cc.get("conv_in"))
        h, new["mid_block"] = self.mid_block(h, z, cc.get("mid_block"))
        for i, blk in enumerate(self.up_blocks):
            h, new[f"up_block_{i}"] = blk(h, z, cc.get(f"up_block_{i}"))
        h = F.silu(self.norm_out(h, z))
        h, new["conv_out"] = self.conv_out(h, cc.get("conv_out"))
        return h, new


class DiagonalGaussianDistribution:
    def __init__(self, parameters):
        self.parameters = parameters
        self.mean, self.logvar = torch.chunk(parameters, 2, dim=1)
        self.logvar = torch.clamp(self.logvar, -30.0, 20.0)
        self.std = torch.exp(0.5 * self.logvar)

    def sample(self, generator=None, noise=None):
        if noise is None:
            # global RNG of the latent's device, one draw, same shape/dtype
            noise = torch.randn(self.mean.shape, generator=generator, device=self.parameters.device,
                                dtype=self.parameters.dtype)
        return self.mean + self.std * noise

    def mode(self):
        return self.mean


class OracleAutoencoderKLCogVideoX(nn.Module):
    def __init__(self, **kw):
        super().__init__()
        cfg = dict(VAE_CONFIG)
        cfg.update(kw)
        self.config = SimpleNamespace(**cfg)
        c = self.config
        self.encoder = Encoder3D(c.in_channels, c.latent_channels, c.block_out_channels, c.layers_per_block,
                                 c.norm_eps, c.norm_num_groups, c.temporal_compression_ratio)
        self.decoder = Decoder3D(c.latent_channels, c.out_channels, c.block_out_channels, c.layers_per_block,
                                 c.norm_eps, c.norm_num_groups, c.temporal_compression_ratio)
        self.use_tiling = False
        self.use_slicing = False
        self.num_latent_frames_batch_size = 2
        self.num_sample_frames_batch_size = 8
        self.tile_sample_min_height = c.sample_height // 2
        self.tile_sample_min_width = c.sample_width // 2
        sf = 2 ** (len(c.block_out_channels) - 1)
        self.tile_latent_min_height = int(self.tile_sample_min_height / sf)
        self.tile_latent_min_width = int(self.tile_sample_min_width / sf)
        self.tile_overlap_factor_height = 1 / 6
        self.tile_overlap_factor_width = 1 / 5

    @property
    def device(self):
        return next(self.parameters()).device

    @property
    def dtype(self):
        return next(self.parameters()).dtype

    def enable_tiling(self):
        self.use_tiling = True

    def enable_slicing(self):
        self.use_slicing = True

    # -- frame-batched passes -------------------------------------------------------------
    @staticmethod
    def frame_batches(num_frames, bs):
        nb = max(num_frames // bs, 1)
        rem = num_frames % bs
        return [(bs * i + (0 if i == 0 else rem), bs * (i + 1) + rem) for i in range(nb)]

    def _encode(self, x):
        _, _, f, h, w = x.shape
        if self.use_tiling and (w > self.tile_sample_min_width or h > self.tile_sample_min_height):
            return self.tiled_encode(x)
        cache, outs = None, []
        for s, e in self.frame_batches(f, self.num_sample_frames_batch_size):
            o, cache = self.encoder(x[:, :, s:e], cache)
            outs.append(o)
        return torch.cat(outs, dim=2)

    def encode(self, x):
        return SimpleNamespace(latent_dist=DiagonalGaussianDistribution(self._encode(x)))

    def _decode(self, z):
        _, _, f, h, w = z.shape
        if self.use_tiling and (w > self.tile_latent_min_width or h > self.tile_latent_min_height):
            return self.tiled_decode(z)
        cache, outs = None, []
        for s, e in self.frame_batches(f, self.num_latent_frames_batch_size):
            o, cache = self.decoder(z[:, :, s:e], cache)
            outs.append(o)
        return torch.cat(outs, dim=2)

    def decode(self, z):
        return SimpleNamespace(sample=self._decode(z))

    # -- spatial tiling (`--is_vae_st`) ------------------------------------------------------
    @staticmethod
    def blend_v(a, b, extent):
        extent = min(a.shape[3], b.shape[3], extent)
        for y in range(extent):
            b[:, :, :, y, :] = a[:, :, :, -extent + y, :] * (1 - y / extent) + b[:, :, :, y, :] * (y / extent)
        return b

    @staticmethod
    def blend_h(a, b, extent):
        extent = min(a.shape[4], b.shape[4], extent)
        for x in range(extent):
            b[:, :, :, :, x] = a[:, :, :, :, -extent + x] * (1 - x / extent) + b[:, :, :, :, x] * (x / extent)
        return b

    def tile_ints(self):
        """Integers of the tiled paths (pinned in tests: 200/288, 5/9, 25/36; 25/36, 40/72, 200/288)."""
        enc = dict(
            stride_h=int(self.tile_sample_min_height * (1 - self.tile_overlap_factor_height)),
            stride_w=int(self.tile_sample_min_width * (1 - self.tile_overlap_factor_width)),
            blend_h=int(self.tile_latent_min_height * self.tile_overlap_factor_height),
            blend_w=int(self.tile_latent_min_width * self.tile_overlap_factor_width),
        )
        enc["limit_h"] = self.tile_latent_min_height - enc["blend_h"]
        enc["limit_w"] = self.tile_latent_min_width - enc["blend_w"]
        dec = dict(
            stride_h=int(self.tile_latent_min_height * (1 - self.tile_overlap_factor_height)),
            stride_w=int(self.tile_latent_min_width * (1 - self.tile_overlap_factor_width)),
            blend_h=int(self.tile_sample_min_height * self.tile_overlap_factor_height),
            blend_w=int(self.tile_sample_min_width * self.tile_overlap_factor_width),
        )
        dec["limit_h"] = self.tile_sample_min_height - dec["blend_h"]
        dec["limit_w"] = self.tile_sample_min_width - dec["blend_w"]
        return enc, dec

    def _tiled(self, x, fn, bs, tile_h, tile_w, ti):
        _, _, f, h, w = x.shape
        rows = []
        for i in range(0, h, ti["stride_h"]):
            row = []
            for j in range(0, w, ti["stride_w"]):
                cache, outs = None, []
                for s, e in self.frame_batches(f, bs):
                    o, cache = fn(x[:, :, s:e, i:i + tile_h, j:j + tile_w], cache)
                    outs.append(o)
                row.append(torch.cat(outs, dim=2))
            rows.append(row)
        result_rows = []
        for i, row in enumerate(rows):
            result_row = []
            for j, tile in enumerate(row):
                if i > 0:
                    tile = self.blend_v(rows[i - 1][j], tile, ti["blend_h"])
                if j > 0:
                    tile = self.blend_h(row[j - 1], tile, ti["blend_w"])
                result_row.append(tile[:, :, :, :ti["limit_h"], :ti["limit_w"]])
            result_rows.append(torch.cat(result_row, dim=4))
        return torch.cat(result_rows, dim=3)

    def tiled_encode(self, x):
        enc, _ = self.tile_ints()
        return self._tiled(x, self.encoder, self.num_sample_frames_batch_size,
                           self.tile_sample_min_height, self.tile_sample_min_width, enc)

    def tiled_decode(self, z):
        _, dec = self.tile_ints()
        return self._tiled(z, self.decoder, self.num_latent_frames_batch_size,
                           self.tile_latent_min_height, self.tile_latent_min_width, dec)
