"""Point serialization keys (first stage of PTv3): the numpy oracle against golden vectors generated
from the reference's own `pointcept/models/utils/serialization` (CPU), and the CUDA kernel against both (bit-exact)."""
import os

import numpy as np
import pytest
import torch

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ("small", "deep", "bigbatch")


@pytest.mark.parametrize("name", CASES)
def test_serialization_oracle_matches_reference_golden(name):
    from oracle import serialization_oracle as so
    z = np.load(os.path.join(G, "serialization.npz"))
    coord, batch, depth = z[f"{name}.coord"], z[f"{name}.batch"], int(z[f"{name}.depth"])
    for o in ("z", "z-trans", "hilbert", "hilbert-trans"):
        assert np.array_equal(so.encode(coord, batch, depth, o), z[f"{name}.code.{o}"]), o
    assert np.array_equal(so.encode(coord, None, depth, "z"), z[f"{name}.code_nobatch.z"])
    d, code, order, inverse = so.serialization(coord, batch, ("z", "z-trans"), depth)
    for k in range(2):
        assert np.all(np.diff(code[k][order[k]]) >= 0)                       # sortedness
        assert np.array_equal(order[k][inverse[k]], np.arange(code.shape[1]))  # inverse really inverts


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_zorder_kernel_bit_exact_vs_reference_golden_and_oracle(name):
    from oracle import serialization_oracle as so
    from unipre3d_b200 import serialization as ser
    z = np.load(os.path.join(G, "serialization.npz"))
    coord, batch, depth = z[f"{name}.coord"], z[f"{name}.batch"], int(z[f"{name}.depth"])
    c, b = torch.tensor(coord, device="cuda"), torch.tensor(batch, device="cuda")
    for o in ("z", "z-trans", "hilbert", "hilbert-trans"):
        got = ser.encode(c, b, depth, o).cpu().numpy()
        assert np.array_equal(got, z[f"{name}.code.{o}"]), o
    assert np.array_equal(ser.encode(c, None, depth, "z").cpu().numpy(), z[f"{name}.code_nobatch.z"])
    d, code, order, inverse = ser.serialization(c, b, ("z", "z-trans"), depth)
    rd, rcode, rorder, rinv = so.serialization(coord, batch, ("z", "z-trans"), depth)
    assert d == rd and np.array_equal(code.cpu().numpy(), rcode)
    assert np.array_equal(order.cpu().numpy(), rorder) and np.array_equal(inverse.cpu().numpy(), rinv)


@pytest.mark.gpu
def test_zorder_kernel_large_and_edge_cases():
    from oracle import serialization_oracle as so
    from unipre3d_b200 import serialization as ser
    g = torch.Generator().manual_seed(5)
    n = 200_003                                                            # the scene-level configs' point counts
    coord = torch.randint(0, 1 << 9, (n, 3), generator=g, dtype=torch.int32)
    batch = torch.sort(torch.randint(0, 4, (n,), generator=g))[0]
    got = ser.encode(coord.cuda(), batch.cuda(), 9, "z-trans").cpu().numpy()
    assert np.array_equal(got, so.encode(coord.numpy(), batch.numpy(), 9, "z-trans"))
    assert ser.encode(torch.zeros((0, 3), dtype=torch.int32, device="cuda"), None, 4, "z").numel() == 0
    with pytest.raises(RuntimeError, match="depth"):
        ser.encode(coord[:4].cuda(), None, 17, "z")
    for o in ("hilbert", "hilbert-trans"):
        got = ser.encode(coord.cuda(), batch.cuda(), 9, o).cpu().numpy()
        assert np.array_equal(got, so.encode(coord.numpy(), batch.numpy(), 9, o)), o
        # a Hilbert curve visits face-adjacent cells consecutively: sorted unique cells differ by exactly one step
    cells = torch.stack(torch.meshgrid(*[torch.arange(8)] * 3, indexing="ij"), -1).reshape(-1, 3).int()
    hk = ser.encode(cells.cuda(), None, 3, "hilbert").cpu()
    path = cells[torch.argsort(hk)].long()
    assert sorted(hk.tolist()) == list(range(512)) and bool(((path[1:] - path[:-1]).abs().sum(1) == 1).all())
    with pytest.raises(RuntimeError, match="CUDA device"):
        ser.encode(coord[:4], None, 9, "z")
