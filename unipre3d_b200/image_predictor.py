"""Frozen image branch with the reference's interface: `ImageFeaturePredictor(cfg, out_channels, pretrained_path)`
returning {"decoder_block_i": tensor} for the four decoder up-blocks of the Stable-Diffusion VAE
(/root/reference/model/image_predictor.py:10-81, which wraps `diffusers.AutoencoderKL.from_pretrained(weights/)` and
hooks `decoder.up_blocks[i]`; GaussianSplatPredictor consumes `decoder_block_3`: 128 channels at image resolution).

`diffusers` is not a dependency here: `AutoencoderKL` below restates the architecture of the sd-vae-ft-mse configuration
(block_out_channels (128,256,512,512), 2 layers per block, 4 latent channels, GroupNorm(32), SiLU, single-head
mid-block attention) with the SAME module / state-dict names as diffusers, so the reference's
`weights/diffusion_pytorch_model.bin` loads with `load_state_dict(strict=True)`; without a checkpoint (none is shipped
with the reference, and there is no network here) the weights are random-initialised, as BASELINE.json's synthetic
setting prescribes.  Parity of this file against diffusers is UNPINNED (the package is absent from the image): the
checks are the parameter count of the public checkpoint (83,653,863) and shape / determinism tests.

Selected with `cfg.model.image_branch = "sdvae"`; the default object-level benchmark keeps the weight-free analytic stem
(`gaussian_predictor.FrozenImageStem`), see DESIGN.md §2.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F


class ResnetBlock2D(nn.Module):
    def __init__(self, cin: int, cout: int, groups: int = 32, eps: float = 1e-6):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=eps)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.norm2 = nn.GroupNorm(groups, cout, eps=eps)
        self.dropout = nn.Dropout(0.0)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x):
        h = self.conv1(F.silu(self.norm1(x)))
        h = self.conv2(self.dropout(F.silu(self.norm2(h))))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return x + h


class Attention(nn.Module):
    """Single-head spatial self-attention of the VAE mid block (diffusers `Attention` with residual_connection=True)."""

    def __init__(self, channels: int, groups: int = 32, eps: float = 1e-6):
        super().__init__()
        self.group_norm = nn.GroupNorm(groups, channels, eps=eps)
        self.to_q = nn.Linear(channels, channels)
        self.to_k = nn.Linear(channels, channels)
        self.to_v = nn.Linear(channels, channels)
        self.to_out = nn.ModuleList([nn.Linear(channels, channels), nn.Dropout(0.0)])

    def forward(self, x):
        B, C, H, W = x.shape
        h = self.group_norm(x).reshape(B, C, H * W).transpose(1, 2)              # (B, HW, C)
        q, k, v = self.to_q(h), self.to_k(h), self.to_v(h)
        h = F.scaled_dot_product_attention(q.unsqueeze(1), k.unsqueeze(1), v.unsqueeze(1)).squeeze(1)
        h = self.to_out[1](self.to_out[0](h))
        return x + h.transpose(1, 2).reshape(B, C, H, W)


class _Conv(nn.Module):
    """`downsamplers.0` / `upsamplers.0`: a module holding `.conv` (diffusers Downsample2D / Upsample2D)."""

    def __init__(self, channels: int, down: bool):
        super().__init__()
        self.down = down
        self.conv = nn.Conv2d(channels, channels, 3, stride=2 if down else 1, padding=0 if down else 1)

    def forward(self, x):
        if self.down:
            return self.conv(F.pad(x, (0, 1, 0, 1)))                              # asymmetric padding, as diffusers
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class DownEncoderBlock2D(nn.Module):
    def __init__(self, cin, cout, layers, add_downsample):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(cin if i == 0 else cout, cout) for i in range(layers)])
        self.downsamplers = nn.ModuleList([_Conv(cout, down=True)]) if add_downsample else None

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        if self.downsamplers is not None:
            x = self.downsamplers[0](x)
        return x


class UpDecoderBlock2D(nn.Module):
    def __init__(self, cin, cout, layers, add_upsample):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(cin if i == 0 else cout, cout) for i in range(layers)])
        self.upsamplers = nn.ModuleList([_Conv(cout, down=False)]) if add_upsample else None

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        if self.upsamplers is not None:
            x = self.upsamplers[0](x)
        return x


class UNetMidBlock2D(nn.Module):
    def __init__(self, channels):
        super().__init__()
        self.attentions = nn.ModuleList([Attention(channels)])
        self.resnets = nn.ModuleList([ResnetBlock2D(channels, channels), ResnetBlock2D(channels, channels)])

    def forward(self, x):
        return self.resnets[1](self.attentions[0](self.resnets[0](x)))


class Encoder(nn.Module):
    def __init__(self, in_channels, latent_channels, block_out_channels: Sequence[int], layers_per_block):
        super().__init__()
        ch = list(block_out_channels)
        self.conv_in = nn.Conv2d(in_channels, ch[0], 3, padding=1)
        self.down_blocks = nn.ModuleList([
            DownEncoderBlock2D(ch[max(i - 1, 0)], ch[i], layers_per_block, add_downsample=i < len(ch) - 1)
            for i in range(len(ch))])
        self.mid_block = UNetMidBlock2D(ch[-1])
        self.conv_norm_out = nn.GroupNorm(32, ch[-1], eps=1e-6)
        self.conv_out = nn.Conv2d(ch[-1], 2 * latent_channels, 3, padding=1)

    def forward(self, x):
        x = self.conv_in(x)
        for b in self.down_blocks:
            x = b(x)
        x = self.mid_block(x)
        return self.conv_out(F.silu(self.conv_norm_out(x)))


class Decoder(nn.Module):
    def __init__(self, out_channels, latent_channels, block_out_channels: Sequence[int], layers_per_block):
        super().__init__()
        ch = list(reversed(block_out_channels))
        self.conv_in = nn.Conv2d(latent_channels, ch[0], 3, padding=1)
        self.mid_block = UNetMidBlock2D(ch[0])
        self.up_blocks = nn.ModuleList([
            UpDecoderBlock2D(ch[max(i - 1, 0)], ch[i], layers_per_block + 1, add_upsample=i < len(ch) - 1)
            for i in range(len(ch))])
        self.conv_norm_out = nn.GroupNorm(32, ch[-1], eps=1e-6)
        self.conv_out = nn.Conv2d(ch[-1], out_channels, 3, padding=1)

    def forward(self, z, collect: Optional[Dict[str, torch.Tensor]] = None, stop_after: Optional[int] = None):
        x = self.mid_block(self.conv_in(z))
        for i, b in enumerate(self.up_blocks):
            x = b(x)
            if collect is not None:
                collect[f"decoder_block_{i}"] = x
            if stop_after is not None and i == stop_after:
                return x
        return self.conv_out(F.silu(self.conv_norm_out(x)))


class AutoencoderKL(nn.Module):
    """sd-vae-ft-mse configuration of diffusers.AutoencoderKL (same state-dict keys)."""

    def __init__(self, in_channels=3, out_channels=3, latent_channels=4, block_out_channels=(128, 256, 512, 512),
                 layers_per_block=2):
        super().__init__()
        self.config = {"block_out_channels": list(block_out_channels), "latent_channels": latent_channels,
                       "in_channels": in_channels, "out_channels": out_channels, "layers_per_block": layers_per_block}
        self.encoder = Encoder(in_channels, latent_channels, block_out_channels, layers_per_block)
        self.decoder = Decoder(out_channels, latent_channels, block_out_channels, layers_per_block)
        self.quant_conv = nn.Conv2d(2 * latent_channels, 2 * latent_channels, 1)
        self.post_quant_conv = nn.Conv2d(latent_channels, latent_channels, 1)

    def encode_mode(self, x):
        """Mode (= mean) of the diagonal-Gaussian posterior: what `AutoencoderKL.forward(sample_posterior=False)` decodes."""
        mean, _ = self.quant_conv(self.encoder(x)).chunk(2, dim=1)
        return mean

    def forward(self, x, collect: Optional[Dict[str, torch.Tensor]] = None, stop_after: Optional[int] = None):
        return self.decoder(self.post_quant_conv(self.encode_mode(x)), collect=collect, stop_after=stop_after)


class ImageFeaturePredictor(nn.Module):
    """image_predictor.py:10-81: frozen VAE, `forward(x)` -> {"decoder_block_0..3"}.  The reference decodes all the way
    to the RGB reconstruction and discards it; here the decoder stops after the last hooked block (same features)."""

    def __init__(self, cfg, out_channels, pretrained_path: Optional[str] = None):
        super().__init__()
        self.out_channels, self.cfg = out_channels, cfg
        self.encoder = AutoencoderKL()
        if pretrained_path is not None:
            sd = torch.load(pretrained_path, map_location="cpu")
            sd = sd.get("model", sd) if isinstance(sd, dict) else sd
            self.encoder.load_state_dict(self.remap_legacy_attention_keys(sd), strict=True)
        self.encoder.eval()
        for p in self.encoder.parameters():
            p.requires_grad = False
        self.encoder_config = self.encoder.config
        # None: follow the caller -- fp32 unless a torch.autocast region is active (the reference runs the VAE in fp32,
        # its config sets force_upcast); a dtype here forces that autocast dtype on CUDA
        self.compute_dtype = None

    @staticmethod
    def remap_legacy_attention_keys(sd):
        """Older sd-vae checkpoints name the mid-block attention projections query / key / value / proj_attn; diffusers
        remaps them to to_q / to_k / to_v / to_out.0 on load (and reshapes their 1x1-conv weights to Linear).  Same here,
        so that such a file loads strictly.  (Untested against the real file: no weights are shipped or downloadable.)"""
        ren = {"query": "to_q", "key": "to_k", "value": "to_v", "proj_attn": "to_out.0"}
        out = {}
        for k, v in sd.items():
            parts = k.split(".")
            if "attentions" in parts and len(parts) >= 2 and parts[-2] in ren:
                parts[-2:-1] = ren[parts[-2]].split(".")
                if v.ndim == 4 and parts[-1] == "weight":
                    v = v.reshape(v.shape[0], v.shape[1])
            out[".".join(parts)] = v
        return out

    def train(self, mode: bool = True):
        super().train(mode)
        self.encoder.eval()                        # the reference keeps the VAE in eval mode
        return self

    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> Dict[str, torch.Tensor]:
        feats: Dict[str, torch.Tensor] = {}
        if x.is_cuda:
            dt = self.compute_dtype
            if dt is None and torch.is_autocast_enabled("cuda"):
                dt = torch.get_autocast_dtype("cuda")
            with torch.autocast("cuda", dtype=dt or torch.bfloat16, enabled=dt is not None):
                self.encoder(x.contiguous(memory_format=torch.channels_last), collect=feats, stop_after=3)
            return {k: v.float() for k, v in feats.items()}
        self.encoder(x.float(), collect=feats, stop_after=3)
        return feats


def count_parameters(m: nn.Module) -> int:
    return sum(p.numel() for p in m.parameters())
