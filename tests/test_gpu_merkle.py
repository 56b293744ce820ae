"""GPU: the device-resident note tree (csrc/merkle.cu) through the C ABI against the tree oracle
(oracle/pyref/merkle.py, itself pinned to the reference's merkle.rs tests in test_oracle_merkle.py).
Bit-exact: all arithmetic is integer."""
import numpy as np
import pytest

import zk_apps_b200 as z
from oracle import corac
from oracle.pyref import bls12_381 as bls, merkle as om, relations as rel
from tests.util import fr_mont_array, fr_from_mont_array, rand_fr

pytestmark = pytest.mark.gpu


def mont(v):
    return bls.fr_to_mont_bytes(v)


def test_add_two_leaves_and_root(ctx):
    # merkle.rs:115-132 with the circuit's hash
    t = z.MerkleTree(ctx, 10)
    assert t.add_leaf(mont(1)) == 0
    assert t.add_leaf(mont(2)) == 1
    h = om.poseidon_hash(1, 2)
    for _ in range(1, 10):
        h = om.poseidon_hash(h, 0)
    assert t.root() == mont(h)
    t.free()


def test_size_limit_and_full_tree_errors(ctx):
    # merkle.rs:134-142 and :91-93
    t = z.MerkleTree(ctx, 10)
    t.add_leaves(fr_mont_array(range(1 << 10)))
    with pytest.raises(z.B200zkError) as e:
        t.add_leaf(mont(0))
    assert e.value.code == -8
    with pytest.raises(z.B200zkError) as e:
        t.gen_proof(0)
    assert e.value.code == -9
    assert t.next_leaf_idx == 1 << 10
    t.free()
    t = z.MerkleTree(ctx, 3)
    with pytest.raises(z.B200zkError) as e:       # root of an empty tree: MerkleTreeNonExistingNode
        t.root()
    assert e.value.code == -10
    with pytest.raises(z.B200zkError) as e:       # all or nothing
        t.add_leaves(fr_mont_array(range(9)))
    assert e.value.code == -8 and t.next_leaf_idx == 0
    t.free()


def test_historical_root(ctx):
    # merkle.rs:144-168
    t = z.MerkleTree(ctx, 10)
    roots = []
    for i in range(10):
        t.add_leaf(mont(i))
        roots.append(t.root())
    t.free()
    t = z.MerkleTree(ctx, 10)
    for i in range(10):
        assert all(t.is_historical_root(r) for r in roots[:i])
        assert not any(t.is_historical_root(r) for r in roots[i:])
        t.add_leaf(mont(i))
    t.free()


@pytest.mark.parametrize("depth,chunks", [(1, [2]), (4, [1, 2, 5, 3]), (7, [100, 1, 26]), (10, [3, 600, 1, 200])])
def test_tree_vs_oracle_ragged_appends(ctx, depth, chunks):
    """Every node, every historical root and every proof of a tree grown by ragged batches == the oracle grown
    leaf by leaf."""
    total = sum(chunks)
    leaves = rand_fr(0xB2000700 + depth, total)
    o = om.MerkleTree(depth)
    want_roots = []
    for v in leaves:
        o.add_leaf(v)
        want_roots.append(o.root())
    t = z.MerkleTree(ctx, depth)
    got_roots, pos = [], 0
    for c in chunks:
        first, roots = t.add_leaves(fr_mont_array(leaves[pos:pos + c]), want_roots=True)
        assert first == pos
        got_roots += fr_from_mont_array(roots)
        pos += c
    assert got_roots == want_roots
    for i in range(1, 2 << depth):
        assert bls.fr_from_mont_bytes(t.node(i)) == o.node_value(i), "node %d" % i
    assert all(t.is_historical_root(mont(r)) for r in want_roots)
    assert not t.is_historical_root(mont(12345))
    if total < (1 << depth):
        ids = sorted(set([0, total - 1, total // 2, min(total, (1 << depth) - 1)]))
        path, shape = t.gen_proofs(ids)
        for k, i in enumerate(ids):
            assert fr_from_mont_array(path[k].reshape(-1)) == o.gen_proof(i)
            assert [bool(b) for b in shape[k]] == o.path_shape(i)
    t.free()


def test_large_tree_vs_cpp_port_and_path_property(ctx):
    """2^16 leaves: the root against the C++ port hashing level by level, and -- a size-independent property --
    every sampled proof walks back to the root under the circuit's rule (merkle_proof.rs:49-57)."""
    depth, n = 17, 1 << 16
    rng = np.random.default_rng(7)
    leaves = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    leaves[:, 31] &= 0x3F
    t = z.MerkleTree(ctx, depth, log_roots=False)
    t.add_leaves(leaves.reshape(-1))
    level = leaves.reshape(-1)
    for _ in range(16):
        level = corac.poseidon_hash_batch(level, 2)
    zero = np.zeros(32, dtype=np.uint8)
    root = corac.poseidon_hash_batch(np.concatenate([level, zero]), 2)     # the right half of the tree is empty: 0
    assert t.root() == root.tobytes()
    ids = [0, 1, 2, 12345, n - 1]
    path, shape = t.gen_proofs(ids)
    for k, i in enumerate(ids):
        cur = leaves[i]
        for lvl in range(depth):
            pair = np.concatenate([cur, path[k, lvl]] if shape[k, lvl] else [path[k, lvl], cur])
            cur = corac.poseidon_hash_batch(pair, 2)
        assert cur.tobytes() == t.root()
    t.free()


def test_tree_feeds_the_prover(ctx):
    """Notes inserted into the tree, paths written straight into device-resident input rows, rows proven:
    the witness generator accepts them (status 0) and the assignment equals the oracle's for the same path."""
    H = rel.TREE_HEIGHT
    relation = z.UpdateNoteRelation(z.WITHDRAW, H)
    ws = [rel.make_witness(s, rel.WITHDRAW) for s in (21, 22, 23)]
    t = z.MerkleTree(ctx, H)
    filler = rand_fr(99, 40)
    hashes = [w.old_note.hash() for w in ws]
    leaves = filler[:7] + [hashes[0]] + filler[7:20] + [hashes[1], hashes[2]] + filler[20:]
    ids = [7, 21, 22]
    t.add_leaves(fr_mont_array(leaves))
    rows = np.concatenate([fr_mont_array(rel.witness_to_inputs(w)) for w in ws])
    n_in = relation.n_inputs_per_proof
    d_rows = ctx.alloc(rows.nbytes)
    d_ids = ctx.alloc(8 * len(ids))
    ctx.upload(d_rows, rows)
    ctx.upload(d_ids, np.asarray(ids, dtype=np.uint64).view(np.uint8))
    t.fill_update_note_inputs_device(d_ids, len(ids), d_rows)
    ctx.sync()
    got = ctx.download(d_rows, rows.nbytes)
    o = om.MerkleTree(H)
    for v in leaves:
        o.add_leaf(v)
    for k, i in enumerate(ids):
        row = fr_from_mont_array(got[k * n_in * 32:(k + 1) * n_in * 32])
        assert row[4] == o.root()
        assert row[13:13 + H] == [1 if s else 0 for s in o.path_shape(i)]
        assert row[13 + H:13 + 2 * H] == o.gen_proof(i)
    z_all, status = relation.witness_batch(ctx, got, len(ids))
    assert list(status) == [0, 0, 0]
    ctx.free(d_rows)
    ctx.free(d_ids)
    t.free()
