"""zk-apps_b200 -- B200-native Groth16 / BLS12-381 prover backend (hot path only).

Host-side Python mirror of the arkworks call sites BASELINE.json names, over the C ABI in
include/b200zk.h (libb200zk.so, built in-tree by build.py).  There is no CPU fallback: loading
fails loudly if the CUDA library is missing, and creating a Context fails without a GPU.
"""
from .ffi import (B200zkError, Context, lib, lib_path, Radix2EvaluationDomain, VariableBaseMSM,  # noqa: F401
                  FR_BYTES, G1_BYTES, G2_BYTES, fr_to_mont, fr_from_mont, STATUS, DEPOSIT, WITHDRAW,
                  poseidon_constants, poseidon_hash_batch, UpdateNoteRelation, UpdateAccountRelation, ProvingKey, VerifyingKey, Groth16, MerkleTree,
                  points_compress, points_decompress, PROOF_STATUS)
from . import ffi, sharded  # noqa: F401,E402
