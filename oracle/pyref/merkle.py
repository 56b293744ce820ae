"""ORACLE (test infrastructure; never imported by zk-apps_b200/) -- the note tree.

Line-by-line restatement of the reference's Merkle tree, the only *executable* tree in it:
  MerkleTree<DEPTH>            shielder/contract/merkle.rs:11-22   (heap layout: root = node 1,
                                                                    leaves at size + idx, size = 2^DEPTH)
  add_leaf                     merkle.rs:48-80    (missing nodes read as 0 -- an empty subtree is 0, NOT H(0,0))
  is_historical_root           merkle.rs:82-87    (roots_log = set of every root the tree ever had)
  gen_proof                    merkle.rs:89-102   (sibling = node[id ^ 1], missing -> 0; fails when the tree is full)
  root                         merkle.rs:104-106
with the hash left pluggable: the contract hashes with SHA-256 (`compute_hash`, merkle.rs:24-28), the
circuit walks the same tree with Poseidon-2 (`hash_fix_len_array(&[left, right])`,
shielder/relations/src/merkle_proof.rs:49-57).  `sha256_hash` pins the indexing against the reference's
own test `add_two_leaves_and_root` (merkle.rs:115-132); `poseidon_hash` is what the GPU tree uses.

path_shape convention (merkle_proof.rs:53-55): selector = is_zero(shape); left = select(sibling, current,
selector) -- so shape[i] = True means the current node is the LEFT child at level i, i.e. its heap id is even.
"""
from __future__ import annotations

import hashlib

from . import poseidon


class MerkleError(Exception):
    pass


class LimitExceeded(MerkleError):       # ShielderError::MerkleTreeLimitExceeded
    pass


class ProofGenFail(MerkleError):        # ShielderError::MerkleTreeProofGenFail
    pass


def poseidon_hash(left: int, right: int) -> int:
    return poseidon.hash_fix_len_array([left, right])


def sha256_hash(left: bytes, right: bytes) -> bytes:
    """compute_hash (merkle.rs:24-28) on 32-byte Scalars."""
    return hashlib.sha256(left + right).digest()


class MerkleTree:
    def __init__(self, depth: int, hash2=poseidon_hash, zero=0):
        self.depth, self.size = depth, 1 << depth
        self.nodes = {}
        self.roots_log = set()
        self.next_leaf_idx = 0
        self.hash2, self.zero = hash2, zero

    def node_value(self, i):
        return self.nodes.get(i, self.zero)

    def add_leaf(self, leaf_value) -> int:                                       # merkle.rs:48-80
        if self.next_leaf_idx == self.size:
            raise LimitExceeded()
        i = self.next_leaf_idx + self.size
        cur = self.next_leaf_idx
        self.nodes[i] = leaf_value
        i //= 2
        while i > 0:
            self.nodes[i] = self.hash2(self.node_value(2 * i), self.node_value(2 * i + 1))
            i //= 2
        self.next_leaf_idx += 1
        self.roots_log.add(self.root())
        return cur

    def is_historical_root(self, r) -> bool:                                     # merkle.rs:82-87
        return r in self.roots_log

    def gen_proof(self, leaf_id: int) -> list:                                   # merkle.rs:89-102
        if self.next_leaf_idx == self.size:
            raise ProofGenFail()
        i = leaf_id + self.size
        res = []
        for _ in range(self.depth):
            res.append(self.node_value(i ^ 1))
            i //= 2
        return res

    def path_shape(self, leaf_id: int) -> list:
        """The [bool; H] that goes with gen_proof(leaf_id) into MerkleProof::new (merkle_proof.rs:22-25)."""
        i = leaf_id + self.size
        out = []
        for _ in range(self.depth):
            out.append((i & 1) == 0)
            i //= 2
        return out

    def root(self):                                                              # merkle.rs:104-106
        if 1 not in self.nodes:
            raise MerkleError("MerkleTreeNonExistingNode")
        return self.nodes[1]


def root_from_path(leaf, shape, path, hash2=poseidon_hash):
    """CircuitMerkleProof::verify's walk (merkle_proof.rs:49-57), off-circuit."""
    cur = leaf
    for s, sib in zip(shape, path):
        cur = hash2(cur, sib) if s else hash2(sib, cur)
    return cur
