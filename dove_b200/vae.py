"""B200-native CogVideoX 3-D causal VAE (host orchestration over libdove_b200 kernels).

Mirror of the diffusers `AutoencoderKLCogVideoX` surface used by the reference:
  pipe.vae.encode(x).latent_dist.sample()   ref: /root/reference/inference_script.py:408-409
  pipe.vae.decode(z).sample (via pipe.decode_latents)               :500
  pipe.vae.config.scaling_factor / block_out_channels                :409, :467
  pipe.vae.enable_slicing() / enable_tiling()                        :643-645
  pipe.vae.device / dtype                                            :407

Data layout in HBM: every activation is CHANNELS-LAST bf16 [T, H, W, C] (C contiguous) so a convolution tap is
a plain 4-D TMA box [th, tw, 64ch] of the input; causal 3x3x3 convs take their 2 preceding frames from the conv cache (a
view of the previous frame batch's input) through a second tensor map — no padded copy (CogVideoXCausalConv3d semantics).  Frame batching
(8 pixel frames / 2 latent frames, first batch takes the remainder) and the per-conv cache follow
`AutoencoderKLCogVideoX._encode/_decode`.

Kernel sequence per ResnetBlock3D:  gn_stats -> gn_apply(+SiLU) -> causal conv (tcgen05 implicit GEMM, cache frames
through a second tensor map) -> gn_stats -> gn_apply ->
[1x1x1 shortcut GEMM] -> conv with the residual add fused in the epilogue.
"""
from __future__ import annotations

from types import SimpleNamespace

import torch

from . import _lib as L
from .weights import VAE_CONFIG

BF = torch.bfloat16


def _pad_to(n, m):
    return (n + m - 1) // m * m


class _TrackedStateDict:
    """Strict checkpoint access: a missing key raises with the component name, and any tensor the build did not
    consume is an error (a silently dropped parameter would make every output wrong with released weights)."""

    def __init__(self, sd, what):
        self._sd, self._what, self._used = sd, what, set()

    def __getitem__(self, k):
        if k not in self._sd:
            raise KeyError(f"{self._what} checkpoint is missing '{k}' (strict loading: no silent defaults)")
        self._used.add(k)
        return self._sd[k]

    def __contains__(self, k):
        return k in self._sd

    def check_all_consumed(self):
        extra = sorted(set(self._sd) - self._used)
        if extra:
            raise KeyError(f"{self._what} checkpoint has {len(extra)} unconsumed tensors, e.g. {extra[:4]}")


class _Conv:
    """Weights of one conv in implicit-GEMM layout [Cout_pad, kt*kh*kw*Cin_pad], K index = tap*Cin_pad + c."""

    def __init__(self, w, b, device):
        if w.dim() == 4:      # Conv2d [Cout, Cin, kh, kw]
            w = w[:, :, None]
        cout, cin, kt, kh, kw = w.shape
        self.cin, self.cout, self.kt, self.kh, self.kw = cin, cout, kt, kh, kw
        self.cin_pad = _pad_to(cin, 64)
        self.cout_pad = _pad_to(cout, 16)
        wp = torch.zeros(self.cout_pad, kt, kh, kw, self.cin_pad, dtype=BF, device=device)
        wp[:cout, :, :, :, :cin] = w.to(device=device, dtype=BF).permute(0, 2, 3, 4, 1)
        self.w = wp.reshape(self.cout_pad, kt * kh * kw * self.cin_pad).contiguous()
        bp = torch.zeros(self.cout_pad, dtype=BF, device=device)
        bp[:cout] = b.to(device=device, dtype=BF)
        self.b = bp


class _Norm:
    def __init__(self, sd, name, device, spatial):
        pre = f"{name}.norm_layer" if spatial else name
        self.gamma = sd[f"{pre}.weight"].to(device=device, dtype=BF).contiguous()
        self.beta = sd[f"{pre}.bias"].to(device=device, dtype=BF).contiguous()
        self.spatial = spatial
        if spatial:
            self.conv_y = _Conv(sd[f"{name}.conv_y.conv.weight"], sd[f"{name}.conv_y.conv.bias"], device)
            self.conv_b = _Conv(sd[f"{name}.conv_b.conv.weight"], sd[f"{name}.conv_b.conv.bias"], device)


class _Resnet:
    def __init__(self, sd, name, device, spatial):
        self.name = name
        self.norm1 = _Norm(sd, f"{name}.norm1", device, spatial)
        self.norm2 = _Norm(sd, f"{name}.norm2", device, spatial)
        self.conv1 = _Conv(sd[f"{name}.conv1.conv.weight"], sd[f"{name}.conv1.conv.bias"], device)
        self.conv2 = _Conv(sd[f"{name}.conv2.conv.weight"], sd[f"{name}.conv2.conv.bias"], device)
        self.shortcut = None
        if f"{name}.conv_shortcut.weight" in sd:
            self.shortcut = _Conv(sd[f"{name}.conv_shortcut.weight"], sd[f"{name}.conv_shortcut.bias"], device)


class LatentDistribution:
    """DiagonalGaussianDistribution over channels-last moments [T, h, w, 32] (mean | logvar)."""

    def __init__(self, moments_cl, shape):
        self._m = moments_cl
        self._shape = shape      # (T, h, w)

    @property
    def parameters(self):
        T, h, w = self._shape
        out = torch.empty(1, 32, T, h, w, dtype=BF, device=self._m.device)
        L.cl_to_ncthw(self._m, out, 32, T, h, w, 32)
        return out

    def sample(self, generator=None, noise=None, scaling=1.0):
        """mean + std * randn.  Exactly one `torch.randn` of shape [1,16,T,h,w] is drawn from the global CUDA
        generator (same stream semantics as diffusers' randn_tensor)."""
        T, h, w = self._shape
        if noise is None:
            noise = torch.randn((1, 16, T, h, w), generator=generator, device=self._m.device, dtype=BF)
        z = torch.empty(1, 16, T, h, w, dtype=BF, device=self._m.device)
        L.gaussian_sample(self._m, noise.contiguous(), z, T * h * w, scaling)
        return z

    def mode(self):
        return self.parameters[:, :16].contiguous()


class AutoencoderKLCogVideoX:
    def __init__(self, state_dict, config=None, device="cuda"):
        cfg = dict(VAE_CONFIG)
        cfg.update(config or {})
        self.config = SimpleNamespace(**cfg)
        self._device = torch.device(device)
        if self._device.type != "cuda":
            raise L.DoveError("dove_b200.AutoencoderKLCogVideoX runs on CUDA (sm_100a) only")
        L.init(self._device.index if self._device.index is not None else torch.cuda.current_device())
        self.use_tiling = False
        self.use_slicing = False
        self.num_latent_frames_batch_size = 2
        self.num_sample_frames_batch_size = 8
        self._build(state_dict)
        self._partial = torch.empty(L.gn_partial_floats(0), dtype=torch.float32, device=self._device)

    # ---- surface ----------------------------------------------------------------------------------
    @property
    def device(self):
        return self._device

    @property
    def dtype(self):
        return BF

    def to(self, *a, **k):
        return self

    def enable_slicing(self):
        self.use_slicing = True      # no-op at batch 1 (diffusers slices over the batch dim)

    def enable_tiling(self):
        self.use_tiling = True

    # ---- weights ----------------------------------------------------------------------------------
    def _build(self, sd):
        sd = _TrackedStateDict(sd, "vae")
        self._build_from(sd)
        sd.check_all_consumed()

    def _build_from(self, sd):
        dev = self._device
        c = self.config
        boc = list(c.block_out_channels)
        lpb = c.layers_per_block
        conv = lambda n: _Conv(sd[f"{n}.weight"], sd[f"{n}.bias"], dev)
        self.enc_conv_in = conv("encoder.conv_in.conv")
        self.enc_down = []
        for i in range(len(boc)):
            res = [_Resnet(sd, f"encoder.down_blocks.{i}.resnets.{j}", dev, False) for j in range(lpb)]
            down = conv(f"encoder.down_blocks.{i}.downsamplers.0.conv") if i != len(boc) - 1 else None
            self.enc_down.append((res, down, i < 2))
        self.enc_mid = [_Resnet(sd, f"encoder.mid_block.resnets.{j}", dev, False) for j in range(2)]
        self.enc_norm_out = _Norm(sd, "encoder.norm_out", dev, False)
        self.enc_conv_out = conv("encoder.conv_out.conv")
        self.dec_conv_in = conv("decoder.conv_in.conv")
        self.dec_mid = [_Resnet(sd, f"decoder.mid_block.resnets.{j}", dev, True) for j in range(2)]
        self.dec_up = []
        for i in range(len(boc)):
            res = [_Resnet(sd, f"decoder.up_blocks.{i}.resnets.{j}", dev, True) for j in range(lpb + 1)]
            up = conv(f"decoder.up_blocks.{i}.upsamplers.0.conv") if i != len(boc) - 1 else None
            self.dec_up.append((res, up, i < 2))
        self.dec_norm_out = _Norm(sd, "decoder.norm_out", dev, True)
        self.dec_conv_out = conv("decoder.conv_out.conv")

    # ---- building blocks (channels-last) ---------------------------------------------------------------
    def _empty(self, *shape):
        return torch.empty(*shape, dtype=BF, device=self._device)

    def _norm_into_padded(self, norm, x, T, H, W, C, zq, silu=True, fused_stats=False):
        """GroupNorm / SpatialNorm3D (+SiLU) of x [T,H,W,C] -> fresh conv input [T,H,W,C] (no temporal padding:
        the causal conv reads its two preceding frames straight from the cache view, see _causal_conv)."""
        stats = torch.empty(64, dtype=torch.float32, device=self._device)
        if fused_stats:      # the producing conv's epilogue already left per-CTA partial sums in self._partial
            L.gn_finalize(self._partial, T * H * W, C, 32, self.config.norm_eps, stats)
        else:
            L.gn_stats(x, C, 32, self.config.norm_eps, self._partial, stats)
        xin = self._empty(T, H, W, C)
        zy = zb = None
        if norm.spatial:
            zq_cl, (Tz, hz, wz) = zq
            nvz = Tz * hz * wz
            zy = self._empty(Tz, hz, wz, C)
            zb = self._empty(Tz, hz, wz, C)
            L.gemm(zq_cl.view(nvz, 64), norm.conv_y.w, zy.view(nvz, C), norm.conv_y.b)
            L.gemm(zq_cl.view(nvz, 64), norm.conv_b.w, zb.view(nvz, C), norm.conv_b.b)
        L.gn_apply(x, xin, T, H, W, C, 32, stats, norm.gamma, norm.beta, silu, zy, zb)
        return xin

    def _causal_conv(self, conv, xin, T, H, W, cache, key, out=None, aux=None, out_mode=0, plane_stride=0,
                     want_stats=False):
        """xin: the conv's input frames [T,H,W,Cin_pad] of this frame batch.  Temporal padding is ZERO-COPY: the two
        preceding frames are the cache = a view of the last two frames of the previous batch's xin (read by the kernel
        through a second tensor map), or frame 0 replicated for the first batch (CogVideoXCausalConv3d semantics)."""
        prev = cache.get(key)
        if out is None:
            out = self._empty(T, H, W, conv.cout_pad)
        res = L.conv3d_causal(xin, prev, conv.w, conv.b, out, conv.cout,
                              epilogue=L.EPI_ADD if aux is not None else L.EPI_BIAS, aux=aux, out_mode=out_mode,
                              plane_stride=plane_stride, gn_partial=self._partial if want_stats else None)
        has_stats = res[1] if want_stats else False
        if T >= 2:
            cache[key] = xin[T - 2:]                      # view: keeps xin alive until the next batch has used it
        else:                                             # single-frame batch: [previous last frame, this frame]
            first = prev[1:2] if prev is not None else xin[0:1]
            cache[key] = torch.cat([first, xin[0:1]], dim=0).contiguous()
        return (out, has_stats) if want_stats else out

    def _resnet(self, r, x, T, H, W, zq, cache, x_stats=False):
        """-> (out, out_has_fused_stats).  x_stats: self._partial holds the GroupNorm partials of x (left there by the
        conv that produced x); every conv here asks for the statistics of ITS output for the next norm."""
        cin, cout = r.conv1.cin, r.conv1.cout
        xin = self._norm_into_padded(r.norm1, x, T, H, W, cin, zq, fused_stats=x_stats)
        h, h_stats = self._causal_conv(r.conv1, xin, T, H, W, cache, r.name + ".conv1", want_stats=True)
        del xin
        xin2 = self._norm_into_padded(r.norm2, h, T, H, W, cout, zq, fused_stats=h_stats)
        del h
        if r.shortcut is not None:
            nv = T * H * W
            res = self._empty(T, H, W, cout)
            L.gemm(x.view(nv, cin), r.shortcut.w, res.view(nv, cout), r.shortcut.b)
        else:
            res = x
        return self._causal_conv(r.conv2, xin2, T, H, W, cache, r.name + ".conv2", aux=res, want_stats=True)

    # ---- encoder ----------------------------------------------------------------------------------
    def _encoder_batch(self, pix, t0, t1, F, H, W, cache):
        """pix: [3, F, H, W] (fp32 or bf16) contiguous; frames t0..t1 -> moments [T', H/8, W/8, 32]."""
        T = t1 - t0
        xin = self._empty(T, H, W, 64)
        # per-channel planes of this frame batch are strided inside pix: gather through a contiguous view
        src = pix[:, t0:t1].contiguous()
        L.pixels_to_cl(src, xin, T, H, W, 64)
        x, st = self._causal_conv(self.enc_conv_in, xin, T, H, W, cache, "conv_in", want_stats=True)
        del xin
        for bi, (res, down, compress_time) in enumerate(self.enc_down):
            for r in res:
                x, st = self._resnet(r, x, T, H, W, None, cache, st)
            if down is not None:
                C = down.cin
                if compress_time:
                    To = 1 + (T - 1) // 2 if T % 2 else T // 2
                    y = self._empty(To, H, W, C)
                    L.time_pool(x, y, T, H * W * C)
                    x, T = y, To
                Ho, Wo = H // 2, W // 2
                y = self._empty(T, Ho, Wo, C)
                _, st = L.conv_cl(x, down.w, down.b, y, T, 1, 3, 3, 2, 0, Ho, Wo, down.cout, gn_partial=self._partial)
                x, H, W = y, Ho, Wo
        for r in self.enc_mid:
            x, st = self._resnet(r, x, T, H, W, None, cache, st)
        C = self.enc_conv_out.cin
        xin = self._norm_into_padded(self.enc_norm_out, x, T, H, W, C, None, fused_stats=st)
        return self._causal_conv(self.enc_conv_out, xin, T, H, W, cache, "conv_out"), (T, H, W)

    @staticmethod
    def frame_batches(num_frames, bs):
        nb = max(num_frames // bs, 1)
        rem = num_frames % bs
        # diffusers slices x[:, :, start:end]; for clips shorter than one batch (e.g. the per-frame path, F = 1) the
        # nominal end bs + rem overshoots and the slice clamps it
        return [(bs * i + (0 if i == 0 else rem), min(bs * (i + 1) + rem, num_frames)) for i in range(nb)]

    @classmethod
    def latent_frames(cls, num_frames, bs=8, levels=2):
        """Latent frame count of `_encode` for a clip of num_frames (frame batching + 2 temporal poolings)."""
        total = 0
        for s, e in cls.frame_batches(num_frames, bs):
            t = e - s
            for _ in range(levels):
                t = 1 + (t - 1) // 2 if t % 2 else t // 2
            total += t
        return total

    def tile_ints(self):
        """Integers of diffusers' tiled_encode / tiled_decode (pinned: 200/288, 5/9, 25/36 | 25/36, 40/72, 200/288)."""
        c = self.config
        sf = 2 ** (len(c.block_out_channels) - 1)
        smin_h, smin_w = c.sample_height // 2, c.sample_width // 2
        lmin_h, lmin_w = int(smin_h / sf), int(smin_w / sf)
        oh, ow = 1 / 6, 1 / 5
        enc = dict(tile_h=smin_h, tile_w=smin_w, stride_h=int(smin_h * (1 - oh)), stride_w=int(smin_w * (1 - ow)),
                   blend_h=int(lmin_h * oh), blend_w=int(lmin_w * ow))
        enc["limit_h"], enc["limit_w"] = lmin_h - enc["blend_h"], lmin_w - enc["blend_w"]
        dec = dict(tile_h=lmin_h, tile_w=lmin_w, stride_h=int(lmin_h * (1 - oh)), stride_w=int(lmin_w * (1 - ow)),
                   blend_h=int(smin_h * oh), blend_w=int(smin_w * ow))
        dec["limit_h"], dec["limit_w"] = smin_h - dec["blend_h"], smin_w - dec["blend_w"]
        return enc, dec

    def _encode_untiled(self, pix):
        """pix [3,F,H,W] contiguous -> moments channels-last [T',H/8,W/8,32] (frame-batched, fresh conv cache)."""
        _, F, H, W = pix.shape
        cache, outs = {}, []
        for (s, e) in self.frame_batches(F, self.num_sample_frames_batch_size):
            m, _ = self._encoder_batch(pix, s, e, F, H, W, cache)
            outs.append(m)
        return torch.cat(outs, dim=0) if len(outs) > 1 else outs[0]

    def _blend_and_stitch(self, rows, ti, layout, out):
        """diffusers blend_v / blend_h in raster order, in place on the un-cropped tiles, then crop + place."""
        yo = 0
        for i, row in enumerate(rows):
            xo = 0
            for j, tile in enumerate(row):
                if i > 0:
                    L.blend(rows[i - 1][j], tile, "v", ti["blend_h"], layout)
                if j > 0:
                    L.blend(row[j - 1], tile, "h", ti["blend_w"], layout)
                if layout == "cl":
                    crop = tile[:, :ti["limit_h"], :ti["limit_w"]]
                    out[:, yo:yo + crop.shape[1], xo:xo + crop.shape[2]] = crop
                    dy, dx = crop.shape[1], crop.shape[2]
                else:
                    crop = tile[:, :, :ti["limit_h"], :ti["limit_w"]]
                    out[:, :, yo:yo + crop.shape[2], xo:xo + crop.shape[3]] = crop
                    dy, dx = crop.shape[2], crop.shape[3]
                xo += dx
            yo += dy
        return out

    def _encode_tiled(self, pix):
        """AutoencoderKLCogVideoX.tiled_encode (`--is_vae_st`): 240x360 px tiles at stride 200x288, each with its own
        frame-batched pass (so GroupNorm statistics are per tile), linear blend of 5/9 latent px, crop 25x36."""
        enc, _ = self.tile_ints()
        _, F, H, W = pix.shape
        rows = []
        for i in range(0, H, enc["stride_h"]):
            row = []
            for j in range(0, W, enc["stride_w"]):
                row.append(self._encode_untiled(pix[:, :, i:i + enc["tile_h"], j:j + enc["tile_w"]].contiguous()))
            rows.append(row)
        Tl = rows[0][0].shape[0]
        out = self._empty(Tl, H // 8, W // 8, 32)
        return self._blend_and_stitch(rows, enc, "cl", out)

    def encode_cl(self, x):
        """x: [1,3,F,H,W] on device (fp32 or bf16) -> (moments channels-last [T',h,w,32], (T',h,w))."""
        assert x.dim() == 5 and x.shape[0] == 1 and x.shape[1] == 3, "batch 1, 3 channels"
        _, _, F, H, W = x.shape
        assert H % 8 == 0 and W % 8 == 0, "H, W must be multiples of 8"
        pix = x[0].contiguous()
        enc, _ = self.tile_ints()
        if self.use_tiling and (W > enc["tile_w"] or H > enc["tile_h"]):
            mom = self._encode_tiled(pix)
        else:
            mom = self._encode_untiled(pix)
        return mom, (mom.shape[0], mom.shape[1], mom.shape[2])

    def encode(self, x):
        x = x.to(self._device)
        if x.dtype not in (torch.float32, BF):
            x = x.to(BF)
        mom, shape = self.encode_cl(x)
        return SimpleNamespace(latent_dist=LatentDistribution(mom, shape))

    # ---- decoder ----------------------------------------------------------------------------------
    def _decoder_batch(self, z, t0, t1, h, w, scale, cache, out, f_off, F_out, out_mode=L.OUT_PLANAR):
        """z: [16, Tz_total, h, w] bf16; latent frames t0..t1 -> pixels written into out[3, F_out, 8h, 8w].
        out_mode: L.OUT_PLANAR (raw decoder output), L.OUT_PLANAR_POST (`*0.5+0.5` clamp of ref :501 fused into the
        last conv's epilogue) or L.OUT_PLANAR_U8 (that, quantised like the reference's savers; `out` is uint8)."""
        T = t1 - t0
        zsrc = z[:, t0:t1].contiguous()
        xin = self._empty(T, h, w, 64)
        L.ncthw_to_cl(zsrc, xin, 16, T, h, w, 64, scale)
        zq = (xin, (T, h, w))                     # the (scaled) latent batch conditions every SpatialNorm3D
        x, st = self._causal_conv(self.dec_conv_in, xin, T, h, w, cache, "conv_in", want_stats=True)
        H, W = h, w
        for r in self.dec_mid:
            x, st = self._resnet(r, x, T, H, W, zq, cache, st)
        for (res, up, compress_time) in self.dec_up:
            for r in res:
                x, st = self._resnet(r, x, T, H, W, zq, cache, st)
            if up is not None:
                C = up.cin
                if compress_time and T > 1:
                    To = 1 + 2 * (T - 1) if T % 2 else 2 * T
                else:
                    To = T
                y = self._empty(To, 2 * H, 2 * W, C)
                L.upsample_nearest(x, y, T, H, W, C, compress_time)
                x, T, H, W = y, To, 2 * H, 2 * W
                y = self._empty(T, H, W, C)
                _, st = L.conv_cl(x, up.w, up.b, y, T, 1, 3, 3, 1, 1, H, W, up.cout, gn_partial=self._partial)
                x = y
        C = self.dec_conv_out.cin
        xin2 = self._norm_into_padded(self.dec_norm_out, x, T, H, W, C, zq, fused_stats=st)
        del x
        # conv_out writes planar NCDHW straight into the output clip at frame offset f_off
        self._causal_conv(self.dec_conv_out, xin2, T, H, W, cache, "conv_out",
                          out=out.view(3, -1)[:, f_off * H * W:], out_mode=out_mode, plane_stride=F_out * H * W)
        return T

    def _decode_untiled(self, zc, scale, out_mode=L.OUT_PLANAR):
        """zc [16,Tz,h,w] bf16 contiguous -> [1,3,F,8h,8w] bf16 / uint8 (frame-batched, fresh conv cache)."""
        _, Tz, h, w = zc.shape
        sf = 2 ** (len(self.config.block_out_channels) - 1)
        batches = self.frame_batches(Tz, self.num_latent_frames_batch_size)
        F_out = 0
        for (s, e) in batches:     # frames produced: the first batch keeps frame 0 single
            T = e - s
            for _ in range(2):
                T = (1 + 2 * (T - 1) if T % 2 else 2 * T) if T > 1 else T
            F_out += T
        out = torch.empty(1, 3, F_out, h * sf, w * sf, device=self._device,
                          dtype=torch.uint8 if out_mode == L.OUT_PLANAR_U8 else BF)
        cache = {}
        f_off = 0
        for (s, e) in batches:
            f_off += self._decoder_batch(zc, s, e, h, w, scale, cache, out[0], f_off, F_out, out_mode)
        return out

    def _decode_tiled(self, zc, scale):
        """AutoencoderKLCogVideoX.tiled_decode: 30x45 latent tiles at stride 25x36, blend 40/72 px, crop 200x288."""
        _, dec = self.tile_ints()
        _, Tz, h, w = zc.shape
        sf = 2 ** (len(self.config.block_out_channels) - 1)
        rows = []
        for i in range(0, h, dec["stride_h"]):
            row = []
            for j in range(0, w, dec["stride_w"]):
                tile = zc[:, :, i:i + dec["tile_h"], j:j + dec["tile_w"]].contiguous()
                row.append(self._decode_untiled(tile, scale)[0])
            rows.append(row)
        F_out = rows[0][0].shape[1]
        out = self._empty(3, F_out, h * sf, w * sf)
        return self._blend_and_stitch(rows, dec, "planar", out)[None]

    def decode_scaled(self, z, scale=1.0, post=None):
        """z: [1,16,Tz,h,w] bf16 -> [1,3,F,8h,8w]; `scale` is applied to z first (decode_latents' 1/0.7).
        post=None: raw decoder output (bf16, about [-1,1]); "unit": `(x*0.5+0.5).clamp(0,1)` (ref :501) bf16;
        "uint8": that value quantised as the reference's savers do, trunc(x*255) (ref :124, :143, :168).
        Untiled decodes fuse the post-processing into the last conv's epilogue; VAE-tiled decodes blend the raw
        tiles first (diffusers order) and post-process afterwards."""
        assert z.dim() == 5 and z.shape[0] == 1 and z.shape[1] == 16
        assert post in (None, "unit", "uint8")
        _, _, Tz, h, w = z.shape
        zc = z[0].to(BF).contiguous()
        _, dec = self.tile_ints()
        if self.use_tiling and (w > dec["tile_w"] or h > dec["tile_h"]):
            raw = self._decode_tiled(zc, scale)
            if post is None:
                return raw
            unit = torch.empty_like(raw)
            L.post_scale(raw, unit)
            return unit if post == "unit" else (unit.float() * 255.0).to(torch.uint8)
        mode = {None: L.OUT_PLANAR, "unit": L.OUT_PLANAR_POST, "uint8": L.OUT_PLANAR_U8}[post]
        return self._decode_untiled(zc, scale, mode)

    def decode(self, z):
        return SimpleNamespace(sample=self.decode_scaled(z.to(self._device), 1.0))
