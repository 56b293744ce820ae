"""CPU: the tree oracle (oracle/pyref/merkle.py) against the reference's own tree tests
(shielder/contract/merkle.rs:110-170), first with the contract's SHA-256 node hash -- which pins the
indexing, the "missing node = 0" rule and the error behaviour against the reference's expectations --
then with the circuit's Poseidon-2, which is what the GPU tree computes."""
import pytest

from oracle.pyref import merkle as om
from oracle.pyref import relations as rel


def u128_scalar(v: int) -> bytes:
    """`1_u128.into()` for the mock Scalar: 32 bytes; only zero/non-zero distinctness matters below."""
    return v.to_bytes(32, "little")


ZERO = u128_scalar(0)


def test_add_two_leaves_and_root_sha256():
    # merkle.rs:115-132, literally
    t = om.MerkleTree(10, om.sha256_hash, ZERO)
    assert t.add_leaf(u128_scalar(1)) == 0
    assert t.add_leaf(u128_scalar(2)) == 1
    hash_left = om.sha256_hash(u128_scalar(1), u128_scalar(2))
    for _ in range(1, 10):
        hash_left = om.sha256_hash(hash_left, ZERO)
    assert hash_left == t.root()


def test_size_limit():
    # merkle.rs:134-142 (depth 5 keeps the Poseidon variant fast; the rule does not depend on depth)
    t = om.MerkleTree(10, om.sha256_hash, ZERO)
    for i in range(1 << 10):
        t.add_leaf(u128_scalar(i))
    with pytest.raises(om.LimitExceeded):
        t.add_leaf(ZERO)
    with pytest.raises(om.ProofGenFail):          # merkle.rs:91-93
        t.gen_proof(0)


def test_historical_root():
    # merkle.rs:144-168
    t = om.MerkleTree(10, om.sha256_hash, ZERO)
    roots = []
    for i in range(10):
        t.add_leaf(u128_scalar(i))
        roots.append(t.root())
    t = om.MerkleTree(10, om.sha256_hash, ZERO)
    for i in range(10):
        assert all(t.is_historical_root(r) for r in roots[:i])
        assert not any(t.is_historical_root(r) for r in roots[i:])
        t.add_leaf(u128_scalar(i))


def test_empty_root_is_an_error():
    with pytest.raises(om.MerkleError):
        om.MerkleTree(4).root()


def test_poseidon_tree_matches_circuit_walk():
    """gen_proof + path_shape of the Poseidon tree feed merkle_proof.rs's walk back to the root, and the
    relation oracle's own walk agrees."""
    t = om.MerkleTree(4)
    leaves = [1000 + i for i in range(11)]
    for v in leaves:
        t.add_leaf(v)
    for i in (0, 5, 10):
        path, shape = t.gen_proof(i), t.path_shape(i)
        assert om.root_from_path(leaves[i], shape, path) == t.root()
        assert rel.merkle_root_from_path(leaves[i], shape, path) == t.root()
    # two leaves: root = H(...H(H(l0, l1), 0)..., 0)   (the Poseidon form of merkle.rs:115-132)
    t2 = om.MerkleTree(4)
    t2.add_leaf(1); t2.add_leaf(2)
    h = om.poseidon_hash(1, 2)
    for _ in range(3):
        h = om.poseidon_hash(h, 0)
    assert h == t2.root()
