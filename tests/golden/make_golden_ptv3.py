"""Generates tests/golden/ptv3_small.npz by running the REFERENCE's own PointTransformerV3
(/root/reference/pointcept/models/point_transformer_v3/point_transformer_v3m1_base.py, with pointcept/models/modules.py,
utils/structure.py, utils/misc.py and utils/serialization/* loaded from /root/reference unmodified) on the CPU.

Its un-vendored dependencies are stubbed: spconv.pytorch (SubMConv3d = the dense-grid oracle oracle/sparse_oracle.py with the
kernels' bf16 operand rounding; SparseConvTensor), torch_scatter.segment_csr (loop restatement), addict.Dict, timm DropPath,
flash_attn absent (=> the reference's exact non-flash attention branch), PDNorm / MODELS registry / PointFusion placeholders.
Run where /root/reference exists:  python tests/golden/make_golden_ptv3.py
"""
import importlib
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(OUT))
sys.path.insert(0, ROOT)


def stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


class Dict(dict):                      # addict.Dict, the part Point uses
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v


class SparseConvTensor:
    def __init__(self, features, indices, spatial_shape, batch_size):
        self.features, self.indices, self.spatial_shape, self.batch_size = features, indices, spatial_shape, batch_size

    def replace_feature(self, f):
        return SparseConvTensor(f, self.indices, self.spatial_shape, self.batch_size)


class SubMConv3d(torch.nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, bias=True, indice_key=None):
        super().__init__()
        k = kernel_size
        self.weight = torch.nn.Parameter(torch.randn(out_channels, k, k, k, in_channels) * 0.05)     # spconv v2 layout
        self.bias = torch.nn.Parameter(torch.zeros(out_channels)) if bias else None

    def forward(self, x):
        from oracle import sparse_oracle as so
        out = so.subm_conv(x.features, x.indices, self.weight, self.bias).float()
        return x.replace_feature(out)


def segment_csr(src, indptr, reduce="sum"):
    outs = []
    for a, b in zip(indptr[:-1].tolist(), indptr[1:].tolist()):
        seg = src[a:b]
        outs.append({"sum": seg.sum(0), "mean": seg.mean(0), "max": seg.max(0).values, "min": seg.min(0).values}[reduce])
    return torch.stack(outs)


class DropPath(torch.nn.Module):
    def __init__(self, p=0.0):
        super().__init__()
        self.drop_prob = p

    def forward(self, x):
        assert not (self.training and self.drop_prob > 0)
        return x


class _Registry:
    def register_module(self, *a, **k):
        return lambda cls: cls


def main():
    stub("addict", Dict=Dict)
    sp = stub("spconv")
    spp = stub("spconv.pytorch", SubMConv3d=SubMConv3d, SparseConvTensor=SparseConvTensor)
    spp.modules = types.SimpleNamespace(is_spconv_module=lambda m: isinstance(m, SubMConv3d))
    sp.pytorch = spp
    stub("torch_scatter", segment_csr=segment_csr)
    stub("timm"); stub("timm.models"); stub("timm.models.layers", DropPath=DropPath)
    sys.modules["flash_attn"] = None                                           # => ImportError => non-flash branch
    stub("fusion"); stub("fusion.point_fusion", PointFusion=object)
    pc = stub("pointcept"); pc.__path__ = [os.path.join(REF, "pointcept")]
    pm = stub("pointcept.models"); pm.__path__ = [os.path.join(REF, "pointcept/models")]
    stub("pointcept.models.point_prompt_training", PDNorm=object)
    stub("pointcept.models.builder", MODELS=_Registry())
    ck = stub("pointcept.models.utils.checkpoint", checkpoint=None)            # utils/__init__ imports it (torch.utils.checkpoint wrapper)
    mod = importlib.import_module("pointcept.models.point_transformer_v3.point_transformer_v3m1_base")

    torch.manual_seed(0)
    kw = dict(in_channels=6, order=("z", "z-trans"), stride=(2, 2), enc_depths=(1, 2, 1), enc_channels=(32, 64, 128),
              enc_num_head=(2, 4, 8), enc_patch_size=(16, 16, 16), dec_depths=(1, 1), dec_channels=(32, 64),
              dec_num_head=(2, 4), dec_patch_size=(16, 16), drop_path=0.0, shuffle_orders=False, enable_flash=False,
              upcast_attention=False, upcast_softmax=False)
    net = mod.PointTransformerV3(cfg=types.SimpleNamespace(opt=types.SimpleNamespace(use_fusion=False)), **kw)
    with torch.no_grad():
        for m in net.modules():
            if isinstance(m, (torch.nn.BatchNorm1d, torch.nn.LayerNorm)):
                m.weight.uniform_(0.5, 1.5); m.bias.normal_(0, 0.2)
    net.train()
    g = torch.Generator().manual_seed(1)
    n_per, grid = (330, 290), 20
    coords, batches = [], []
    for b, n in enumerate(n_per):
        cells = torch.randperm(grid ** 3, generator=g)[:n]
        c = torch.stack([cells // grid ** 2, (cells // grid) % grid, cells % grid], 1)
        c[:, 2] = c[:, 2] % 5 + c[:, 0] // 6                                     # surface-like occupancy
        coords.append(torch.unique(c, dim=0))
    grid_coord = torch.cat(coords).int()
    offset = torch.cumsum(torch.tensor([c.shape[0] for c in coords]), 0)
    n = grid_coord.shape[0]
    feat = torch.randn(n, 6, generator=g)
    coord = grid_coord.float() * 0.02 + torch.rand(n, 3, generator=g) * 0.01
    data = {"coord": coord.clone(), "grid_coord": grid_coord.clone(), "feat": feat.clone(), "offset": offset.clone()}
    inter = {}
    hooks = [net.embedding.register_forward_hook(lambda m, i, o: inter.__setitem__("emb_feat", o.feat.detach().clone())),
             net.enc.enc0.register_forward_hook(lambda m, i, o: inter.__setitem__("enc0_feat", o.feat.detach().clone())),
             net.enc.enc1.down.register_forward_hook(lambda m, i, o: inter.__setitem__("down1_feat", o.feat.detach().clone())),
             net.enc.enc0.block0.cpe.register_forward_hook(lambda m, i, o: inter.__setitem__("cpe0_feat", o.feat.detach().clone())),
             net.enc.enc0.block0.attn.register_forward_hook(lambda m, i, o: inter.__setitem__("attn0_feat", o.feat.detach().clone()))]
    # SerializedPooling is built with its default shuffle_orders=True (the model-level flag is not forwarded to it,
    # point_transformer_v3m1_base.py:633-641): every forward draws torch.randperm(2) per pooling layer from the global CPU
    # generator.  Seed it so that the consumer of this fixture can replay the same draws.
    torch.manual_seed(1234)
    with torch.no_grad():
        point = net(data, None, None, None)
    for h in hooks:
        h.remove()
    out = {"cfg_json": np.array(repr(kw)), "coord": coord.numpy(), "grid_coord": grid_coord.numpy(), "feat": feat.numpy(),
           "offset": offset.numpy(), "out_feat": point.feat.numpy(), "out_coord": point.coord.numpy(),
           "out_batch": point.batch.numpy()}
    out.update({"inter." + k: v.numpy() for k, v in inter.items()})
    out.update({"sd." + k: v.numpy() for k, v in net.state_dict().items()})
    np.savez_compressed(os.path.join(OUT, "ptv3_small.npz"), **out)
    print("wrote ptv3_small.npz", point.feat.shape, float(point.feat.abs().max()))


if __name__ == "__main__":
    main()
