"""Synthetic, valid instances of the shielder update-note relation, built with the library's own
GPU Poseidon (no oracle involved): random note fields, random Merkle path, a 2-token account and a
deposit/withdraw amount that keeps balances in range (SURVEY.md section 8d "Concrete synthetic
inputs").  Row layout = the input of b200zk_update_note_witness_batch (include/b200zk.h)."""
from __future__ import annotations

import numpy as np

from .ffi import Context, DEPOSIT, WITHDRAW, poseidon_hash_batch


def _rand_fr(rng, *shape) -> np.ndarray:
    """Uniform 254-bit values (< r): valid Montgomery words of some field element."""
    a = rng.integers(0, 256, size=(*shape, 32), dtype=np.uint8)
    a[..., 31] &= 0x3F
    return a


def _to_mont(ctx: Context, ints) -> np.ndarray:
    canon = np.frombuffer(b"".join(int(v).to_bytes(32, "little") for v in ints), dtype=np.uint8)
    return ctx.field_op(0, 5, canon).reshape(len(ints), 32)          # B200ZK_OP_TO_MONT on the GPU


def make_update_note_instances(ctx: Context, n: int, seed: int, kind: int = WITHDRAW, tree_height: int = 10) -> np.ndarray:
    """-> uint8 array of shape (n, 18 + 2*tree_height, 32)."""
    H = tree_height
    rng = np.random.default_rng(seed)
    tokens = _rand_fr(rng, n, 2)
    which = rng.integers(0, 2, size=n)
    bal = [[int(rng.integers(1, 1 << 62)) << 30 | int(rng.integers(0, 1 << 30)), int(rng.integers(0, 1 << 62))] for _ in range(n)]
    amount, new_bal = [], []
    for i in range(n):
        b = bal[i][which[i]]
        a = int(rng.integers(0, 1 << 62)) % (b + 1) if kind == WITHDRAW else int(rng.integers(0, 1 << 62))
        amount.append(a)
        nb = list(bal[i])
        nb[which[i]] = b - a if kind == WITHDRAW else b + a
        new_bal.append(nb)
    m_amount = _to_mont(ctx, amount)
    m_bal = _to_mont(ctx, [b for row in bal for b in row]).reshape(n, 2, 32)
    m_nbal = _to_mont(ctx, [b for row in new_bal for b in row]).reshape(n, 2, 32)
    old_acc = np.stack([tokens[:, 0], m_bal[:, 0], tokens[:, 1], m_bal[:, 1]], axis=1)       # (n, 4, 32)
    new_acc = np.stack([tokens[:, 0], m_nbal[:, 0], tokens[:, 1], m_nbal[:, 1]], axis=1)
    h_old_acc = poseidon_hash_batch(ctx, old_acc.reshape(-1), 4).reshape(n, 32)
    h_new_acc = poseidon_hash_batch(ctx, new_acc.reshape(-1), 4).reshape(n, 32)
    zk_id = _rand_fr(rng, n)
    old_note = np.stack([zk_id, _rand_fr(rng, n), _rand_fr(rng, n), h_old_acc], axis=1)
    new_note = np.stack([zk_id, _rand_fr(rng, n), _rand_fr(rng, n), h_new_acc], axis=1)
    leaf = poseidon_hash_batch(ctx, old_note.reshape(-1), 4).reshape(n, 32)
    new_note_hash = poseidon_hash_batch(ctx, new_note.reshape(-1), 4).reshape(n, 32)
    shape_bits = rng.integers(0, 2, size=(n, H))
    path = _rand_fr(rng, n, H)
    one = _to_mont(ctx, [1])[0]
    cur = leaf
    for lvl in range(H):
        left_is_cur = shape_bits[:, lvl].astype(bool)[:, None]                                # shape 1: current is the left child
        left = np.where(left_is_cur, cur, path[:, lvl])
        right = np.where(left_is_cur, path[:, lvl], cur)
        cur = poseidon_hash_batch(ctx, np.stack([left, right], axis=1).reshape(-1), 2).reshape(n, 32)
    root = cur
    user = _rand_fr(rng, n)
    token = tokens[np.arange(n), which]
    shape_fr = np.where(shape_bits[..., None].astype(bool), one[None, None, :], np.zeros(32, dtype=np.uint8)[None, None, :])
    rows = np.concatenate([m_amount[:, None], token[:, None], user[:, None], new_note_hash[:, None], root[:, None],
                           new_note, old_note, shape_fr.astype(np.uint8), path, user[:, None], old_acc], axis=1)
    assert rows.shape == (n, 18 + 2 * H, 32)
    return np.ascontiguousarray(rows.astype(np.uint8))
