"""CPU (numpy) restatement of the point-cloud serialization PTv3 starts with.

TEST INFRASTRUCTURE ONLY -- imported by tests/, never by the product package `unipre3d_b200`.

Follows /root/reference/pointcept/models/utils/serialization/z_order.py:41-51 (`KeyLUT.xyz2key`: bit i of x, y, z goes to
bits 3i+2, 3i+1, 3i of the key), serialization/default.py:8-25 (`encode`: "z-trans" swaps x and y; the batch index is OR-ed
in above bit 3*depth) and pointcept/models/utils/structure.py:47-107 (`Point.serialization`: order = argsort(code),
inverse = scatter of arange).  Pinned against the reference's own functions by tests/golden/serialization.npz
(tests/golden/make_golden.py `serialization_fixture`).
"""
from __future__ import annotations

import numpy as np


def z_order_key(x, y, z, depth: int = 16) -> np.ndarray:
    x, y, z = (np.asarray(a).astype(np.int64) for a in (x, y, z))
    key = np.zeros_like(x)
    for i in range(depth):
        m = np.int64(1) << i
        key |= ((x & m) << (2 * i + 2)) | ((y & m) << (2 * i + 1)) | ((z & m) << (2 * i))
    return key


def hilbert_key(x, y, z, depth: int) -> np.ndarray:
    """serialization/hilbert.py:96-191 (Skilling's transpose on integers instead of bit planes): per bit from the top and
    per dimension, invert the lower bits of dimension 0 (bit set) or exchange the differing lower bits with dimension 0
    (bit clear); interleave with dimension 0 most significant; Gray-decode."""
    m = (1 << depth) - 1
    X = [np.asarray(v).astype(np.int64) & m for v in (x, y, z)]
    Q = 1 << (depth - 1)
    while Q > 0:
        P = Q - 1
        for d in range(3):
            on = (X[d] & Q) != 0
            t = np.where(on, 0, (X[0] ^ X[d]) & P)
            X[0] = np.where(on, X[0] ^ P, X[0] ^ t)
            if d != 0:
                X[d] = X[d] ^ t
        Q >>= 1
    g = np.zeros_like(X[0])
    for i in range(depth):
        for d in range(3):
            g |= ((X[d] >> i) & 1) << (3 * i + (2 - d))
    s = 1
    while s < 3 * depth:
        g ^= g >> s
        s <<= 1
    return g


def encode(grid_coord, batch=None, depth: int = 16, order: str = "z") -> np.ndarray:
    g = np.asarray(grid_coord).astype(np.int64)
    if order == "z":
        code = z_order_key(g[:, 0], g[:, 1], g[:, 2], depth)
    elif order == "z-trans":
        code = z_order_key(g[:, 1], g[:, 0], g[:, 2], depth)
    elif order == "hilbert":
        code = hilbert_key(g[:, 0], g[:, 1], g[:, 2], depth)
    elif order == "hilbert-trans":
        code = hilbert_key(g[:, 1], g[:, 0], g[:, 2], depth)
    else:
        raise NotImplementedError(order)
    if batch is not None:
        code = (np.asarray(batch).astype(np.int64) << (depth * 3)) | code
    return code


def serialization(grid_coord, batch, orders=("z", "z-trans"), depth=None):
    """-> depth, code (k,n), order (k,n), inverse (k,n)   (structure.py:47-107 without shuffle_orders)."""
    g = np.asarray(grid_coord)
    if depth is None:
        depth = int(g.max()).bit_length()
    code = np.stack([encode(g, batch, depth, o) for o in orders])
    order = np.argsort(code, axis=1, kind="stable")
    inverse = np.zeros_like(order)
    for k in range(code.shape[0]):
        inverse[k, order[k]] = np.arange(code.shape[1])
    return depth, code, order, inverse
