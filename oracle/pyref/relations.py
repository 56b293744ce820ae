"""ORACLE (test infrastructure) -- the shielder "update note" statement as an R1CS.

Functional spec = the reference's halo2 circuit builders (nothing in the reference executes them;
they are the definition of WHAT is proven):
  update_note_circuit        /root/reference/shielder/relations/src/relations/update_note.rs:106-149
  UpdateNoteInput::new       update_note.rs:47-88   (witness loading order)
  verify_note_circuit        update_note.rs:91-103
  CircuitMerkleProof::verify relations/src/merkle_proof.rs:38-61
  update_account_circuit     relations/src/relations/update_account.rs:68-95
  Note field order           relations/src/note.rs:33-37
The reference leaves Account / Operation abstract (account.rs:8-21, operation.rs:3-23).  The
concrete shapes are borrowed from the mock:
  Account = TOKENS_NUMBER = 2 (token, balance) pairs   shielder/mocked_zk/src/account.rs:12-14, lib.rs:17
  OpPub   = {amount: u128, token, user}; Deposit adds, Withdraw subtracts   mocked_zk/src/ops.rs:6-25
  OpPriv  = {user};  combine requires op_pub.user == op_priv.user          ops.rs:30-32,47-62
  update  = checked add/sub of amount on the balance whose token matches;  error if none matches or
            on u128 overflow/underflow                                      account.rs:36-79
The R1CS encoding (not the halo2 gates) is new design (BASELINE.json fixes Groth16): PARITY UNPINNED.
Deposit and withdraw are the same statement with the sign of the balance update flipped; each is
its own circuit / proving key.  TREE_HEIGHT is a const generic without default in the reference
(merkle_proof.rs:11); the only concrete depth is mocked_zk::MERKLE_TREE_DEPTH = 10 (lib.rs:16).

Public inputs, in order (update_note.rs:121,127): op_pub = (amount, token, user), new_note_hash,
merkle_root, old_note.nullifier.
"""
from __future__ import annotations
from dataclasses import dataclass, field
from .bls12_381 import R
from . import poseidon as pos
from .r1cs import (ConstraintSystem, LC, assert_equal, is_equal, is_zero, mul, poseidon_hash, range_bits, select)

TREE_HEIGHT = 10
TOKENS_NUMBER = 2
BALANCE_BITS = 128
DEPOSIT, WITHDRAW = 0, 1


@dataclass
class Note:                                  # relations/src/note.rs:6-11
    zk_id: int
    trapdoor: int
    nullifier: int
    account_hash: int
    def to_vec(self): return [self.zk_id, self.trapdoor, self.nullifier, self.account_hash]
    def hash(self): return pos.hash_fix_len_array(self.to_vec())


@dataclass
class Account:                               # mocked_zk/src/account.rs:12-14
    tokens: list
    balances: list
    def to_vec(self):
        out = []
        for t, b in zip(self.tokens, self.balances): out += [t, b]
        return out
    def hash(self): return pos.hash_fix_len_array(self.to_vec())
    def update(self, kind: int, amount: int, token: int) -> "Account":
        """Checked update (mocked_zk/src/account.rs:36-79); raises like the mock returns Err."""
        matches = [i for i, t in enumerate(self.tokens) if t == token]
        if len(matches) != 1: raise ValueError("AccountUpdateError: token")
        i = matches[0]
        nb = self.balances[i] + amount if kind == DEPOSIT else self.balances[i] - amount
        if not 0 <= nb < (1 << BALANCE_BITS): raise ValueError("AccountUpdateError: range")
        bal = list(self.balances); bal[i] = nb
        return Account(list(self.tokens), bal)


@dataclass
class UpdateNoteWitness:
    """Everything UpdateNoteInput::new loads (update_note.rs:47-88)."""
    kind: int
    amount: int
    token: int
    user: int                    # op_pub.user
    new_note_hash: int
    merkle_root: int
    new_note: Note
    old_note: Note
    path_shape: list             # TREE_HEIGHT bools; True = current node is the LEFT child (merkle_proof.rs:30,53-55)
    path: list                   # TREE_HEIGHT siblings
    op_priv_user: int
    old_account: Account

    def public_inputs(self):
        return [self.amount % R, self.token % R, self.user % R, self.new_note_hash % R,
                self.merkle_root % R, self.old_note.nullifier % R]


def merkle_root_from_path(leaf: int, path_shape, path) -> int:
    """Off-circuit walk with the circuit's convention: shape True -> (current, sibling)."""
    cur = leaf
    for shape, sib in zip(path_shape, path):
        cur = pos.hash_fix_len_array([cur, sib] if shape else [sib, cur])
    return cur


def make_witness(seed: int, kind: int = WITHDRAW, tree_height: int = TREE_HEIGHT) -> UpdateNoteWitness:
    """A valid, seeded instance (random note fields, random leaf position)."""
    from .algos import SplitMix64
    g = SplitMix64(0xB2000000 + seed)
    tokens = [g.fr(), g.fr()]
    balances = [(g.next() | (g.next() << 64)) % (1 << 126), g.next()]
    old_acc = Account(tokens, balances)
    which = g.next() & 1
    amount = (g.next() % (old_acc.balances[which] + 1)) if kind == WITHDRAW else g.next()
    new_acc = old_acc.update(kind, amount, tokens[which])
    zk_id = g.fr()
    old_note = Note(zk_id, g.fr(), g.fr(), old_acc.hash())
    new_note = Note(zk_id, g.fr(), g.fr(), new_acc.hash())
    shape = [bool(g.next() & 1) for _ in range(tree_height)]
    path = [g.fr() for _ in range(tree_height)]
    root = merkle_root_from_path(old_note.hash(), shape, path)
    user = g.fr()
    return UpdateNoteWitness(kind, amount, tokens[which], user, new_note.hash(), root, new_note, old_note,
                             shape, path, user, old_acc)


def synthesize_update_note(w: UpdateNoteWitness, tree_height: int = TREE_HEIGHT) -> ConstraintSystem:
    """update_note_circuit as R1CS; returns the constraint system with its full assignment."""
    cs = ConstraintSystem()
    # ---- instance variables (make_public order, update_note.rs:121,127)
    amount = cs.alloc_input(w.amount)
    token = cs.alloc_input(w.token)
    user = cs.alloc_input(w.user)
    new_note_hash = cs.alloc_input(w.new_note_hash)
    merkle_root = cs.alloc_input(w.merkle_root)
    old_nullifier = cs.alloc_input(w.old_note.nullifier)
    # ---- witnesses, in UpdateNoteInput::new order (update_note.rs:58-76)
    new_note = [cs.alloc_witness(v) for v in w.new_note.to_vec()]
    old_zk_id = cs.alloc_witness(w.old_note.zk_id)
    old_trapdoor = cs.alloc_witness(w.old_note.trapdoor)
    old_account_hash = cs.alloc_witness(w.old_note.account_hash)
    old_note = [old_zk_id, old_trapdoor, old_nullifier, old_account_hash]
    path_shape = [cs.alloc_witness(1 if s else 0) for s in w.path_shape]      # merkle_proof.rs:27-34
    path = [cs.alloc_witness(v) for v in w.path]
    op_priv_user = cs.alloc_witness(w.op_priv_user)
    acc_tokens, acc_balances = [], []
    for t, b in zip(w.old_account.tokens, w.old_account.balances):
        acc_tokens.append(cs.alloc_witness(t)); acc_balances.append(cs.alloc_witness(b))

    # ---- verify_note_circuit(new_note, new_note_hash)            update_note.rs:129 -> :91-103
    assert_equal(cs, poseidon_hash(cs, new_note), new_note_hash)
    # ---- old_note_hash                                            update_note.rs:131
    current = poseidon_hash(cs, old_note)
    # ---- merkle_proof.verify                                      merkle_proof.rs:38-61
    for i in range(tree_height):
        selector = is_zero(cs, path_shape[i])                       # :53
        left = select(cs, path[i], current, selector)               # :54
        right = select(cs, current, path[i], selector)              # :55
        current = poseidon_hash(cs, [left, right])                  # :56
    assert_equal(cs, current, merkle_root)                          # :59-60
    # ---- CircuitOperation::combine(op_priv, op_pub).unwrap()      update_note.rs:139; ops.rs:47-62
    assert_equal(cs, user, op_priv_user)
    # ---- update_account_circuit                                   update_account.rs:68-95
    old_vec = []
    for t, b in zip(acc_tokens, acc_balances): old_vec += [t, b]
    assert_equal(cs, poseidon_hash(cs, old_vec), old_account_hash)  # :79-85
    # new_account = old_account.update(operation)                   :87 ; account.rs:36-79
    range_bits(cs, amount, BALANCE_BITS)
    matches = LC()
    new_vec = []
    for t, b in zip(acc_tokens, acc_balances):
        eq = is_equal(cs, t, token)
        delta = mul(cs, eq, amount)
        nb = b + delta if w.kind == DEPOSIT else b - delta
        range_bits(cs, nb, BALANCE_BITS)                            # checked_add / checked_sub
        matches = matches + eq
        new_vec += [t, nb]
    assert_equal(cs, matches, LC.const(1))                          # exactly one token matches
    assert_equal(cs, poseidon_hash(cs, new_vec), new_note[3])       # :88-94 (new_note.account_hash)
    return cs


# ------------------------------------------------------------------------ update-account as a relation of its own
@dataclass
class UpdateAccountWitness:
    """UpdateAccountInput (relations/src/relations/update_account.rs:18-30): public old/new account hash and the
    operation, witness = the old account."""
    kind: int
    old_account_hash: int
    new_account_hash: int
    amount: int
    token: int
    user: int
    old_account: Account

    def public_inputs(self):
        return [self.old_account_hash % R, self.new_account_hash % R, self.amount % R, self.token % R, self.user % R]


def make_account_witness(seed: int, kind: int = WITHDRAW) -> UpdateAccountWitness:
    from .algos import SplitMix64
    g = SplitMix64(0xACC00000 + seed)
    tokens = [g.fr(), g.fr()]
    balances = [(g.next() | (g.next() << 64)) % (1 << 126), g.next()]
    old = Account(tokens, balances)
    which = g.next() & 1
    amount = (g.next() % (old.balances[which] + 1)) if kind == WITHDRAW else g.next()
    new = old.update(kind, amount, tokens[which])
    return UpdateAccountWitness(kind, old.hash(), new.hash(), amount, tokens[which], g.fr(), old)


def _update_account_gadget(cs, kind, amount, token, acc_tokens, acc_balances, old_account_hash, new_account_hash):
    """update_account_circuit (update_account.rs:68-95) over already allocated variables; shared by both relations."""
    old_vec = []
    for t, b in zip(acc_tokens, acc_balances): old_vec += [t, b]
    assert_equal(cs, poseidon_hash(cs, old_vec), old_account_hash)  # :79-85
    range_bits(cs, amount, BALANCE_BITS)                            # new_account = old_account.update(operation)  :87
    matches = LC()
    new_vec = []
    for t, b in zip(acc_tokens, acc_balances):
        eq = is_equal(cs, t, token)
        delta = mul(cs, eq, amount)
        nb = b + delta if kind == DEPOSIT else b - delta
        range_bits(cs, nb, BALANCE_BITS)                            # checked_add / checked_sub
        matches = matches + eq
        new_vec += [t, nb]
    assert_equal(cs, matches, LC.const(1))                          # exactly one token matches
    assert_equal(cs, poseidon_hash(cs, new_vec), new_account_hash)  # :88-94


def synthesize_update_account(w: UpdateAccountWitness) -> ConstraintSystem:
    """update_account_circuit as an R1CS of its own: instance = (old_account_hash, new_account_hash, amount, token,
    user) in the field order of UpdateAccountInput's "public inputs" (:23-26), witness = old_account (:29)."""
    cs = ConstraintSystem()
    old_h = cs.alloc_input(w.old_account_hash)
    new_h = cs.alloc_input(w.new_account_hash)
    amount = cs.alloc_input(w.amount)
    token = cs.alloc_input(w.token)
    cs.alloc_input(w.user)
    acc_tokens, acc_balances = [], []
    for t, b in zip(w.old_account.tokens, w.old_account.balances):
        acc_tokens.append(cs.alloc_witness(t)); acc_balances.append(cs.alloc_witness(b))
    _update_account_gadget(cs, w.kind, amount, token, acc_tokens, acc_balances, old_h, new_h)
    return cs


def account_witness_to_inputs(w: UpdateAccountWitness) -> list:
    """The 9 field elements of one instance in UpdateAccountInput::new argument order (update_account.rs:37-42)."""
    return [w.old_account_hash, w.new_account_hash, w.amount, w.token, w.user] + w.old_account.to_vec()


def witness_to_inputs(w: UpdateNoteWitness) -> list:
    """The (18 + 2H) field elements of one instance in UpdateNoteInput::new argument order -- the
    input row of b200zk_update_note_witness_batch (include/b200zk.h)."""
    return ([w.amount, w.token, w.user, w.new_note_hash, w.merkle_root] + w.new_note.to_vec() + w.old_note.to_vec()
            + [1 if s else 0 for s in w.path_shape] + list(w.path) + [w.op_priv_user] + w.old_account.to_vec())
