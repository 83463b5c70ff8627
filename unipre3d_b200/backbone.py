"""Object-level point-cloud backbone of the render-loss path: FPS/ball-query tokenizer + mini-PointNet +
16-block ViT encoder.

Mirrors the operator surface and the state-dict layout of
/root/reference/openpoints/models/backbone/transformer.py:10-327 (Mlp, Attention, Block, TransformerEncoder,
Encoder, PointTransformerEncoder) and /root/reference/openpoints/models/layers/group_embed.py:14-57
(SubsampleGroup), so a reference checkpoint's `point_network.encoder.*` keys load unchanged.

B200 notes
  - tokenizer: FPS -> (gather centres + ball query + grouping + centring) are 2 hand-written kernels
    (unipre3d_b200/csrc/pointops.cu) instead of the reference's 4 kernels + 3 torch ops; the (B,3,N) transposed
    copy of the cloud is never materialised.
  - attention: 129 tokens x 6 heads x 64 -> one fused SDPA call per block (library kernel); the dense
    projections / MLPs are library GEMMs (cuBLASLt), optionally in bf16 via torch.autocast (reference: fp32).
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import pointops
from .fusion import FeatureFusion


class DropPath(nn.Module):
    """Stochastic depth per sample (timm.models.layers.DropPath semantics: scale_by_keep=True)."""

    def __init__(self, drop_prob: float = 0.0):
        super().__init__()
        self.drop_prob = float(drop_prob)

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = 1.0 - self.drop_prob
        mask = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep)
        if keep > 0.0:
            mask.div_(keep)
        return x * mask


class Mlp(nn.Module):
    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.0):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = act_layer()
        self.fc2 = nn.Linear(hidden_features, out_features)
        self.drop = nn.Dropout(drop)

    def forward(self, x):
        return self.drop(self.fc2(self.drop(self.act(self.fc1(x)))))


class Attention(nn.Module):
    """transformer.py:36-77: qkv without bias, scale head_dim**-0.5, softmax, proj with bias."""

    def __init__(self, dim, num_heads=8, qkv_bias=False, qk_scale=None, attn_drop=0.0, proj_drop=0.0):
        super().__init__()
        self.num_heads = num_heads
        head_dim = dim // num_heads
        self.scale = qk_scale or head_dim ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)

    def forward(self, x):
        B, N, C = x.shape
        qkv = self.qkv(x).reshape(B, N, 3, self.num_heads, C // self.num_heads).permute(2, 0, 3, 1, 4)
        q, k, v = qkv[0], qkv[1], qkv[2]
        p = self.attn_drop.p if self.training else 0.0
        x = F.scaled_dot_product_attention(q, k, v, dropout_p=p, scale=self.scale)
        x = x.transpose(1, 2).reshape(B, N, C)
        return self.proj_drop(self.proj(x))


class Block(nn.Module):
    def __init__(self, dim, num_heads, mlp_ratio=4.0, qkv_bias=False, qk_scale=None, drop=0.0, attn_drop=0.0,
                 drop_path=0.0, act_layer=nn.GELU, norm_layer=nn.LayerNorm):
        super().__init__()
        self.norm1 = norm_layer(dim)
        self.drop_path = DropPath(drop_path) if drop_path > 0.0 else nn.Identity()
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer, drop=drop)
        self.attn = Attention(dim, num_heads=num_heads, qkv_bias=qkv_bias, qk_scale=qk_scale, attn_drop=attn_drop,
                              proj_drop=drop)

    def forward(self, x):
        x = x + self.drop_path(self.attn(self.norm1(x)))
        x = x + self.drop_path(self.mlp(self.norm2(x)))
        return x


class TransformerEncoder(nn.Module):
    """transformer.py:123-207: `x = block(x + pos)` on EVERY layer; feature fusion after the last block."""

    def __init__(self, embed_dim=768, depth=4, num_heads=12, mlp_ratio=4.0, qkv_bias=False, qk_scale=None,
                 drop_rate=0.0, attn_drop_rate=0.0, drop_path_rate=0.0):
        super().__init__()
        self.blocks = nn.ModuleList([
            Block(dim=embed_dim, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, qk_scale=qk_scale,
                  drop=drop_rate, attn_drop=attn_drop_rate,
                  drop_path=drop_path_rate[i] if isinstance(drop_path_rate, list) else drop_path_rate)
            for i in range(depth)])

    def forward(self, x, pos, center, image_features, c2w_projection_matrix, intrinsic, feature_fusion):
        if x.is_cuda and not getattr(self, "force_module_path", False):
            # B200 path: all blocks as one autograd node over the fused kernels of csrc/backbone.cu (the module
            # loop below is the same arithmetic, kept for host-side parity tests against the reference fixtures)
            from . import fused_encoder
            if fused_encoder.supports(self.blocks):
                x = fused_encoder.run_encoder_stack(self.blocks, x, pos.expand_as(x), self.training)
                if feature_fusion is not None:
                    x = feature_fusion(x, center, image_features, c2w_projection_matrix, intrinsic)
                return x
        last = len(self.blocks) - 1
        for idx, block in enumerate(self.blocks):
            x = block(x + pos)
            if feature_fusion is not None and idx == last:
                x = feature_fusion(x, center, image_features, c2w_projection_matrix, intrinsic)
        return x


class Encoder(nn.Module):
    """Mini-PointNet per group (transformer.py:210-243).  Conv1d(k=1) layers kept as Conv1d so checkpoint keys and
    BatchNorm1d statistics (over B*G*K samples) are the reference's."""

    def __init__(self, encoder_channel):
        super().__init__()
        self.encoder_channel = encoder_channel
        self.first_conv = nn.Sequential(nn.Conv1d(3, 128, 1), nn.BatchNorm1d(128), nn.ReLU(inplace=True),
                                        nn.Conv1d(128, 256, 1))
        self.second_conv = nn.Sequential(nn.Conv1d(512, 512, 1), nn.BatchNorm1d(512), nn.ReLU(inplace=True),
                                         nn.Conv1d(512, self.encoder_channel, 1))

    @staticmethod
    def _pointwise(conv: nn.Conv1d, x2d: torch.Tensor) -> torch.Tensor:
        """Conv1d(kernel 1) on channels-last rows == one GEMM (rows x C_in) @ W^T; parameters keep Conv1d's shape."""
        return F.linear(x2d, conv.weight.squeeze(-1), conv.bias)

    def forward(self, point_groups):
        """point_groups: (B, G, K, 3) -> (B, G, C).  Same arithmetic as transformer.py:227-243, laid out
        channels-last so the four 1x1 convolutions are plain GEMMs over B*G*K rows (BatchNorm1d on (rows, C) uses
        the same per-channel statistics over all B*G*K samples as on (B*G, C, K))."""
        bs, g, n, _ = point_groups.shape
        rows = point_groups.reshape(bs * g * n, 3)
        f = self._pointwise(self.first_conv[0], rows)
        f = self.first_conv[2](self.first_conv[1](f))
        f = self._pointwise(self.first_conv[3], f)                                   # (BGK, 256)
        f3 = f.reshape(bs * g, n, -1)
        fg = torch.max(f3, dim=1, keepdim=True)[0]                                   # (BG, 1, 256)
        f = torch.cat([fg.expand(-1, n, -1), f3], dim=2).reshape(bs * g * n, -1)     # (BGK, 512) global || local
        f = self._pointwise(self.second_conv[0], f)
        f = self.second_conv[2](self.second_conv[1](f))
        f = self._pointwise(self.second_conv[3], f)                                  # (BGK, C)
        fg = torch.max(f.reshape(bs * g, n, -1), dim=1)[0]
        return fg.reshape(bs, g, self.encoder_channel)


    def forward_grouped(self, neighborhood):
        """neighborhood (B,3,G,K), the tokenizer's own layout -> (B,G,C).  On CUDA in train mode this is ONE autograd
        node over the fused kernels of csrc/pointnet.cu (fused_pointnet.py); `forward` is the same arithmetic as
        module calls (eval mode, CPU parity tests)."""
        if neighborhood.is_cuda and self.training and not getattr(self, "force_module_path", False):
            from . import fused_pointnet
            if fused_pointnet.supports(self):
                return fused_pointnet.run_mini_pointnet(self, neighborhood)
        return self.forward(neighborhood.permute(0, 2, 3, 1))


class SubsampleGroup(nn.Module):
    """group_embed.py:14-57 restricted to what the path uses: FPS subsample + ball-query grouping."""

    def __init__(self, num_groups=256, group_size=32, subsample="fps", group="ballquery", radius=0.1, **kwargs):
        super().__init__()
        self.num_groups, self.group_size, self.radius = num_groups, group_size, radius
        if not any(s in subsample.lower() for s in ("fps", "furthest", "farthest")):
            raise NotImplementedError(f"{subsample.lower()} is not implemented. Only support fps")
        if not ("ball" in group.lower() or "query" in group.lower()):
            raise NotImplementedError(f"{group.lower()} is not implemented. Only support ballquery")

        # Set by trainer.Trainer ONLY while it captures its step graph: (neighborhood, center) static tensors that hold
        # the grouping of the step's batch, computed ahead of the step (FPS + ball query depend on the input coordinates
        # alone, so the trainer runs them for batch i+1 next to step i).  None everywhere else: forward() then groups inline.
        self.lookahead = None

    def group(self, p, out=None):
        """FPS + ball-query grouping of p (B,N,3) -> (neighborhood (B,3,G,K), center (B,G,3))."""
        return pointops.subsample_group(p, self.num_groups, self.group_size, self.radius, out=out)

    def forward(self, p, x=None):
        if x is not None:
            raise NotImplementedError("feature grouping is not on the render-loss path (transformer.py:305)")
        if self.lookahead is not None:
            neigh, center = self.lookahead
            if neigh.shape[0] != p.shape[0] or neigh.device != p.device:
                raise RuntimeError("SubsampleGroup.lookahead does not belong to this batch")
            return neigh, center
        return self.group(p)


class PointTransformerEncoder(nn.Module):
    """transformer.py:246-327.  point_predictor.py:62-64 builds it with
    (in_channels=3, num_groups=128, encoder_dims=384, depth=16)."""

    def __init__(self, num_groups=256, group_size=32, subsample="fps", group="ballquery", radius=0.1,
                 encoder_dims=256, trans_dim=384, drop_path_rate=0.1, depth=12, num_heads=6, **kwargs):
        super().__init__()
        self.group_divider = SubsampleGroup(num_groups, group_size, subsample, group, radius)
        self.trans_dim = trans_dim
        self.encoder = Encoder(encoder_channel=encoder_dims)
        self.reduce_dim = nn.Linear(encoder_dims, trans_dim)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, trans_dim))
        self.cls_pos = nn.Parameter(torch.randn(1, 1, trans_dim))
        self.pos_embed = nn.Sequential(nn.Linear(3, 128), nn.GELU(), nn.Linear(128, trans_dim))
        dpr = [x.item() for x in torch.linspace(0, drop_path_rate, depth)]
        self.blocks = TransformerEncoder(embed_dim=trans_dim, depth=depth, drop_path_rate=dpr, num_heads=num_heads)
        self.norm = nn.LayerNorm(trans_dim)
        self.use_fusion = kwargs.get("use_fusion", True)

    def forward(self, pts, image_features, c2w_projection_matrix, feature_mlps, intricsic):
        feature_fusion = FeatureFusion(feature_mlps) if (self.use_fusion and feature_mlps is not None) else None
        if isinstance(pts, dict):
            pts = pts["pos"]
        pts = pts[:, :, :3].contiguous()
        neighborhood, center = self.group_divider(pts)                       # (B,3,G,K), (B,G,3)
        group_input_tokens = self.encoder.forward_grouped(neighborhood)      # (B,3,G,K) -> (B,G,C)
        group_input_tokens = self.reduce_dim(group_input_tokens)
        cls_tokens = self.cls_token.expand(group_input_tokens.size(0), -1, -1)
        cls_pos = self.cls_pos.expand(group_input_tokens.size(0), -1, -1)
        pos = self.pos_embed(center)
        x = torch.cat((cls_tokens, group_input_tokens), dim=1)
        pos = torch.cat((cls_pos, pos), dim=1)
        x = self.blocks.forward(x, pos, center, image_features, c2w_projection_matrix, intricsic, feature_fusion)
        if x.is_cuda and not getattr(self.blocks, "force_module_path", False):
            from .fused_encoder import fused_layer_norm
            x = fused_layer_norm(self.norm, x)
        else:
            x = self.norm(x)
        return x[:, 1:, :], center
