//! arkworks-shaped front end of libb200zk (include/b200zk.h).  NOT COMPILED in the repository's image (no Rust
//! toolchain there); kept in step with the C ABI by construction: it only calls `b200zk_sys`, which is generated
//! from the header.  Each item names the arkworks 0.4 item it stands in for.
//!
//! Layout facts relied on: `Fp<MontBackend<_, N>>` is `BigInt<N>([u64; N])` in Montgomery form, so `&[Fr]` is
//! `n x 32` bytes as is; `Affine { x, y, infinity }` is not ABI-stable, so points are packed explicitly.
use ark_bls12_381::{Bls12_381, Fq, Fq2, Fr, G1Affine, G1Projective, G2Affine, G2Projective};
use ark_ff::{BigInt, PrimeField};
use ark_groth16::{Proof, ProvingKey, VerifyingKey};
use ark_relations::r1cs::SynthesisError;
use ark_serialize::{CanonicalDeserialize, CanonicalSerialize};
use ark_std::rand::RngCore;
use ark_std::UniformRand;
use b200zk_sys as sys;
use core::ffi::CStr;

/// One GPU, one stream set.  `!Sync`: a ctx is single-threaded (create one per thread).
pub struct Gpu {
    ctx: *mut sys::b200zk_ctx,
}

#[derive(Debug)]
pub struct Error {
    pub code: i32,
    pub message: String,
}

#[derive(Clone, Copy)]
pub enum OpKind {
    Deposit = 0,
    Withdraw = 1,
}

fn fq_bytes(x: &Fq, out: &mut [u8]) {
    for (j, l) in x.0 .0.iter().enumerate() {
        out[j * 8..][..8].copy_from_slice(&l.to_le_bytes());
    }
}

fn fq_from(b: &[u8]) -> Fq {
    let mut l = [0u64; 6];
    for j in 0..6 {
        l[j] = u64::from_le_bytes(b[j * 8..][..8].try_into().unwrap());
    }
    ark_ff::Fp(BigInt(l), core::marker::PhantomData)
}

/// x||y Montgomery limbs (96 B per point) + one infinity flag per point
pub fn pack_g1(bases: &[G1Affine]) -> (Vec<u8>, Vec<u8>) {
    let mut out = vec![0u8; bases.len() * 96];
    let mut inf = vec![0u8; bases.len()];
    for (i, p) in bases.iter().enumerate() {
        if p.infinity {
            inf[i] = 1;
            continue;
        }
        fq_bytes(&p.x, &mut out[i * 96..][..48]);
        fq_bytes(&p.y, &mut out[i * 96 + 48..][..48]);
    }
    (out, inf)
}

/// x.c0||x.c1||y.c0||y.c1 (192 B per point)
pub fn pack_g2(bases: &[G2Affine]) -> (Vec<u8>, Vec<u8>) {
    let mut out = vec![0u8; bases.len() * 192];
    let mut inf = vec![0u8; bases.len()];
    for (i, p) in bases.iter().enumerate() {
        if p.infinity {
            inf[i] = 1;
            continue;
        }
        for (k, c) in [&p.x.c0, &p.x.c1, &p.y.c0, &p.y.c1].into_iter().enumerate() {
            fq_bytes(c, &mut out[i * 192 + k * 48..][..48]);
        }
    }
    (out, inf)
}

fn unpack_g1(b: &[u8; 96], inf: bool) -> G1Affine {
    if inf {
        return G1Affine::identity();
    }
    G1Affine::new_unchecked(fq_from(&b[..48]), fq_from(&b[48..]))
}

fn unpack_g2(b: &[u8; 192], inf: bool) -> G2Affine {
    if inf {
        return G2Affine::identity();
    }
    G2Affine::new_unchecked(
        Fq2::new(fq_from(&b[..48]), fq_from(&b[48..96])),
        Fq2::new(fq_from(&b[96..144]), fq_from(&b[144..])),
    )
}

fn bigint_bytes(xs: &[Fr]) -> Vec<u8> {
    // canonical BigInt<4>, as `into_bigint()` -- what `msm_bigint` and (r, s) expect
    xs.iter().flat_map(|x| x.into_bigint().0.into_iter().flat_map(|l| l.to_le_bytes())).collect()
}

impl Gpu {
    /// Fails with B200ZK_ERR_NO_DEVICE when there is no GPU: there is no CPU path to fall back to.
    pub fn new(device: i32) -> Result<Self, Error> {
        let mut ctx = core::ptr::null_mut();
        let rc = unsafe { sys::b200zk_init(device, &mut ctx) };
        if rc != sys::B200ZK_OK {
            return Err(Error { code: rc, message: "b200zk_init failed".into() });
        }
        Ok(Gpu { ctx })
    }

    fn check(&self, rc: i32) -> Result<(), Error> {
        if rc == sys::B200ZK_OK {
            return Ok(());
        }
        let message = unsafe { CStr::from_ptr(sys::b200zk_last_error(self.ctx)) }.to_string_lossy().into_owned();
        Err(Error { code: rc, message })
    }

    /// `<G1Projective as VariableBaseMSM>::msm_bigint(bases, bigints)` (truncates to the shorter input, like
    /// `msm_unchecked`; `VariableBaseMSM::msm` callers check the lengths first and return `Err(len)`).
    pub fn msm_bigint_g1(&self, bases: &[G1Affine], scalars: &[BigInt<4>]) -> Result<G1Projective, Error> {
        let n = bases.len().min(scalars.len());
        let (b, inf) = pack_g1(&bases[..n]);
        let s = unsafe { core::slice::from_raw_parts(scalars.as_ptr() as *const u8, n * 32) };
        let (mut out, mut is_inf) = ([0u8; 96], 0u8);
        self.check(unsafe {
            sys::b200zk_msm_g1(self.ctx, b.as_ptr(), inf.as_ptr(), s.as_ptr(), n, out.as_mut_ptr(), &mut is_inf)
        })?;
        Ok(unpack_g1(&out, is_inf != 0).into())
    }

    /// `<G2Projective as VariableBaseMSM>::msm_bigint`
    pub fn msm_bigint_g2(&self, bases: &[G2Affine], scalars: &[BigInt<4>]) -> Result<G2Projective, Error> {
        let n = bases.len().min(scalars.len());
        let (b, inf) = pack_g2(&bases[..n]);
        let s = unsafe { core::slice::from_raw_parts(scalars.as_ptr() as *const u8, n * 32) };
        let (mut out, mut is_inf) = ([0u8; 192], 0u8);
        self.check(unsafe {
            sys::b200zk_msm_g2(self.ctx, b.as_ptr(), inf.as_ptr(), s.as_ptr(), n, out.as_mut_ptr(), &mut is_inf)
        })?;
        Ok(unpack_g2(&out, is_inf != 0).into())
    }

    /// `Radix2EvaluationDomain::<Fr>::{fft,ifft}_in_place` and the `get_coset(offset)` forms.  arkworks zero-pads
    /// to the domain size; `Radix2EvaluationDomain::new` returning `None` maps to B200ZK_ERR_DOMAIN_TOO_LARGE.
    pub fn fft_in_place(&self, log_size: u32, offset: Option<Fr>, inverse: bool, v: &mut Vec<Fr>) -> Result<(), Error> {
        v.resize(1usize << log_size, Fr::from(0u64));
        let off = offset.map(|o| o.0 .0);
        self.check(unsafe {
            sys::b200zk_ntt_fr(
                self.ctx,
                v.as_mut_ptr() as *mut u8,
                log_size,
                inverse as i32,
                off.as_ref().map_or(core::ptr::null(), |o| o.as_ptr() as *const u8),
                1,
            )
        })
    }
}

/// Device-resident bases of one rank's slice of a large MSM (`b200zk_bases_upload`, plain bases).
pub struct GpuBasesG1<'a> {
    gpu: &'a Gpu,
    raw: *mut sys::b200zk_bases,
    n: usize,
}

impl Drop for GpuBasesG1<'_> {
    fn drop(&mut self) {
        unsafe { sys::b200zk_bases_free(self.gpu.ctx, self.raw) }
    }
}

impl Gpu {
    /// The 128-byte NCCL id rank 0 creates and hands to the other ranks (any transport of the host's choosing).
    pub fn comm_unique_id() -> Result<[u8; 128], Error> {
        let mut id = [0u8; 128];
        let rc = unsafe { sys::b200zk_comm_unique_id(id.as_mut_ptr()) };
        if rc != sys::B200ZK_OK {
            return Err(Error { code: rc, message: "NCCL not loadable (libnccl.so.2)".into() });
        }
        Ok(id)
    }

    /// Collective over all ranks: attaches an NCCL communicator to this context (`ncclCommInitRank`).
    pub fn comm_init(&self, id: &[u8; 128], rank: i32, world: i32) -> Result<(), Error> {
        self.check(unsafe { sys::b200zk_comm_init(self.ctx, id.as_ptr(), rank, world) })
    }

    /// This rank's contiguous slice of the bases of a sharded MSM, uploaded once.
    pub fn upload_bases_g1(&self, bases: &[G1Affine]) -> Result<GpuBasesG1<'_>, Error> {
        let (b, inf) = pack_g1(bases);
        let mut raw = core::ptr::null_mut();
        self.check(unsafe { sys::b200zk_bases_upload(self.ctx, 1, b.as_ptr(), inf.as_ptr(), bases.len(), 0, &mut raw) })?;
        Ok(GpuBasesG1 { gpu: self, raw, n: bases.len() })
    }

    /// `VariableBaseMSM::msm_bigint` over all GPUs of the communicator: every rank passes ITS slice of the scalars
    /// (same range as its bases) and gets the same result: local bucket pipeline -> in-place all_gather of the
    /// 96-byte affine partials -> sum, one stream of work inside the library.
    pub fn msm_sharded_g1(&self, bases: &GpuBasesG1, scalars_local: &[BigInt<4>]) -> Result<G1Projective, Error> {
        if scalars_local.len() > bases.n {
            return Err(Error { code: sys::B200ZK_ERR_BAD_LEN, message: "more scalars than bases in this rank's slice".into() });
        }
        let (mut out, mut is_inf) = ([0u8; 96], 0u8);
        self.check(unsafe {
            sys::b200zk_msm_sharded(self.ctx, bases.raw, scalars_local.as_ptr() as *const _, 0, scalars_local.len(), out.as_mut_ptr(), &mut is_inf)
        })?;
        Ok(unpack_g1(&out, is_inf != 0).into())
    }
}

impl Drop for Gpu {
    fn drop(&mut self) {
        unsafe { sys::b200zk_destroy(self.ctx) }
    }
}

/// Device-resident `ark_groth16::ProvingKey` of the shielder update-note relation (deposit / withdraw).
pub struct GpuProvingKey<'a> {
    gpu: &'a Gpu,
    raw: *mut sys::b200zk_pk,
    r1cs: *mut sys::b200zk_r1cs,
    tree_height: u32,
}

/// Upload once; `precompute` stores the window multiples of every query (all windows then share one bucket set).
pub fn upload_pk<'a>(gpu: &'a Gpu, kind: OpKind, tree_height: u32, pk: &ProvingKey<Bls12_381>, precompute: bool) -> Result<GpuProvingKey<'a>, Error> {
    let mut r1cs = core::ptr::null_mut();
    gpu.check(unsafe { sys::b200zk_update_note_r1cs(kind as i32, tree_height, &mut r1cs) })?;
    let (alpha, _) = pack_g1(&[pk.vk.alpha_g1]);
    let (beta1, _) = pack_g1(&[pk.beta_g1]);
    let (delta1, _) = pack_g1(&[pk.delta_g1]);
    let (beta2, _) = pack_g2(&[pk.vk.beta_g2]);
    let (delta2, _) = pack_g2(&[pk.vk.delta_g2]);
    // infinity = all-zero encoding inside the queries (b queries are ~30 % infinity)
    let (a, _) = pack_g1(&pk.a_query);
    let (b1, _) = pack_g1(&pk.b_g1_query);
    let (b2, _) = pack_g2(&pk.b_g2_query);
    let (l, _) = pack_g1(&pk.l_query);
    let (h, _) = pack_g1(&pk.h_query);
    let mut raw = core::ptr::null_mut();
    gpu.check(unsafe {
        sys::b200zk_pk_upload(
            gpu.ctx, r1cs, alpha.as_ptr(), beta1.as_ptr(), beta2.as_ptr(), delta1.as_ptr(), delta2.as_ptr(), a.as_ptr(),
            b1.as_ptr(), b2.as_ptr(), l.as_ptr(), h.as_ptr(), precompute as i32, &mut raw,
        )
    })?;
    Ok(GpuProvingKey { gpu, raw, r1cs, tree_height })
}

impl Drop for GpuProvingKey<'_> {
    fn drop(&mut self) {
        unsafe {
            sys::b200zk_pk_free(self.gpu.ctx, self.raw);
            sys::b200zk_r1cs_free(self.r1cs);
        }
    }
}

/// `Groth16::<Bls12_381>::create_proof_with_reduction(circuit, &pk, r, s)` for a batch of update-note instances.
/// `inputs`: rows of `18 + 2 * tree_height` field elements in `UpdateNoteInput::new` argument order
/// (shielder/relations/src/relations/update_note.rs:47-57).  An unsatisfied witness -> `SynthesisError::Unsatisfiable`.
pub fn create_proofs(pk: &GpuProvingKey, inputs: &[Fr], batch: usize, r: &[Fr], s: &[Fr]) -> Result<Vec<Proof<Bls12_381>>, SynthesisError> {
    assert_eq!(r.len(), batch);
    assert_eq!(s.len(), batch);
    // the C side reads batch * (18 + 2 * tree_height) elements: a short slice must never reach it
    assert_eq!(inputs.len(), batch * (18 + 2 * pk.tree_height as usize), "inputs: batch rows of 18 + 2 * tree_height elements");
    let (rb, sb) = (bigint_bytes(r), bigint_bytes(s));
    let mut out = vec![0u8; batch * 192];
    let rc = unsafe {
        sys::b200zk_update_note_prove_batch(
            pk.gpu.ctx, pk.raw, inputs.as_ptr() as *const u8, batch, rb.as_ptr(), sb.as_ptr(), out.as_mut_ptr(),
            core::ptr::null_mut(),
        )
    };
    match rc {
        sys::B200ZK_OK => Ok(out.chunks(192).map(|c| Proof::deserialize_compressed(c).expect("proof bytes")).collect()),
        sys::B200ZK_ERR_UNSATISFIED => Err(SynthesisError::Unsatisfiable),
        _ => panic!("{:?}", pk.gpu.check(rc)),
    }
}

/// A batch submitted with [`submit_proofs`]: `wait` blocks until the GPU is done and yields the proofs.  The output
/// buffer lives inside the ticket, so the pointer the library keeps until the wait stays valid.
pub struct PendingProofs<'a, 'b> {
    pk: &'a GpuProvingKey<'b>,
    ticket: u64,
    out: Vec<u8>,
}

/// Asynchronous form of [`create_proofs`] (`b200zk_update_note_prove_submit`): returns without waiting for the GPU;
/// up to two batches may be in flight per context, so that the witness generation of one batch and the assembly of
/// the other run under bucket accumulation.  The input slices are free again on return.
pub fn submit_proofs<'a, 'b>(pk: &'a GpuProvingKey<'b>, inputs: &[Fr], batch: usize, r: &[Fr], s: &[Fr]) -> Result<PendingProofs<'a, 'b>, Error> {
    if r.len() != batch || s.len() != batch || inputs.len() != batch * (18 + 2 * pk.tree_height as usize) {
        return Err(Error { code: sys::B200ZK_ERR_BAD_LEN, message: "inputs / r / s do not match the batch size".into() });
    }
    let (rb, sb) = (bigint_bytes(r), bigint_bytes(s));
    let mut out = vec![0u8; batch * 192];
    let mut ticket = 0u64;
    pk.gpu.check(unsafe {
        sys::b200zk_update_note_prove_submit(
            pk.gpu.ctx, pk.raw, inputs.as_ptr() as *const _, 0, batch, rb.as_ptr(), sb.as_ptr(), out.as_mut_ptr(),
            core::ptr::null_mut(), &mut ticket,
        )
    })?;
    Ok(PendingProofs { pk, ticket, out })
}

impl PendingProofs<'_, '_> {
    pub fn wait(self) -> Result<Vec<Proof<Bls12_381>>, SynthesisError> {
        match unsafe { sys::b200zk_prove_wait(self.pk.gpu.ctx, self.ticket) } {
            sys::B200ZK_OK => Ok(self.out.chunks(192).map(|c| Proof::deserialize_compressed(c).expect("proof bytes")).collect()),
            sys::B200ZK_ERR_UNSATISFIED => Err(SynthesisError::Unsatisfiable),
            rc => panic!("{:?}", self.pk.gpu.check(rc)),
        }
    }
}

/// `create_random_proof_with_reduction(circuit, &pk, rng)`: arkworks draws `r`, then `s`.
pub fn create_random_proofs<R: RngCore>(pk: &GpuProvingKey, inputs: &[Fr], batch: usize, rng: &mut R) -> Result<Vec<Proof<Bls12_381>>, SynthesisError> {
    let (mut r, mut s) = (Vec::with_capacity(batch), Vec::with_capacity(batch));
    for _ in 0..batch {
        r.push(Fr::rand(rng));
        s.push(Fr::rand(rng));
    }
    create_proofs(pk, inputs, batch, &r, &s)
}

/// `Groth16::verify_proof` for a batch: one verdict per proof (`true` = accepted); the place of the mock
/// `ZkProof::verify_update` (shielder/mocked_zk/src/relations.rs:127-155).
pub fn verify_proofs(gpu: &Gpu, vk: &VerifyingKey<Bls12_381>, proofs: &[Proof<Bls12_381>], public_inputs: &[Fr]) -> Result<Vec<bool>, Error> {
    // the C side reads proofs.len() * (gamma_abc_g1.len() - 1) elements
    if public_inputs.len() != proofs.len() * (vk.gamma_abc_g1.len() - 1) {
        return Err(Error { code: sys::B200ZK_ERR_BAD_LEN, message: "public_inputs: one row of num_inputs - 1 elements per proof".into() });
    }
    let mut vkb = Vec::new();
    vk.serialize_compressed(&mut vkb).expect("vk bytes");
    let mut raw = core::ptr::null_mut();
    gpu.check(unsafe { sys::b200zk_vk_deserialize(gpu.ctx, vkb.as_ptr(), vkb.len(), 1, &mut raw) })?;
    let mut pb = Vec::with_capacity(proofs.len() * 192);
    for p in proofs {
        p.serialize_compressed(&mut pb).expect("proof bytes");
    }
    let mut st = vec![0i32; proofs.len()];
    let rc = unsafe {
        sys::b200zk_groth16_verify_batch(
            gpu.ctx, raw, pb.as_ptr() as *const _, public_inputs.as_ptr() as *const _, 0, proofs.len(), 1, st.as_mut_ptr(),
        )
    };
    unsafe { sys::b200zk_vk_free(gpu.ctx, raw) };
    gpu.check(rc)?;
    Ok(st.into_iter().map(|v| v == sys::B200ZK_PROOF_ACCEPTED).collect())
}
