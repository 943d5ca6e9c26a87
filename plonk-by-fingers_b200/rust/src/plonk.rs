//! Drop-in replacement for the reference's `src/plonk.rs`: the same public items with the same names, generics,
//! fields and signatures, with `SRS::create`, `SRS::eval_at_s`, `Plonk::prove` and `Plonk::verify` running on the GPU
//! through the C ABI of `include/pbh_b200.h` (`crate::ffi`).  Every other module of the crate (`ec`, `poly`, `matrix`,
//! `fft`, `constraints`, `pbh::*`, `utils`) stays as it is, so `src/pbh/mod.rs::test_plonk_gen_proof` compiles and runs
//! unchanged against this file.
//!
//! SOURCE ONLY: this repository's image has no cargo / rustc, so the file is never compiled here (tests/test_rust_shim.py
//! checks its surface textually against the reference's signatures).
//!
//! | item here                               | reference                       |
//! |-----------------------------------------|---------------------------------|
//! | `PlonkTypes`                            | src/plonk.rs:15-26              |
//! | `SRS<P>` + pub fields, `create`, `eval_at_s` | src/plonk.rs:28-58         |
//! | `Proof<P>` (Debug, PartialEq), `Challange<P>` | src/plonk.rs:60-108       |
//! | `Plonk<P>`, `new`, `prove`, `verify`    | src/plonk.rs:110-175, 191-650   |
//!
//! Additions (no counterpart upstream): `B200Wire` (how a `PlonkTypes` instance is encoded at the C ABI, implemented for
//! `PlonkByHandTypes`), `Plonk::prove_batch` / `verify_batch` over byte planes, `Plonk::on_device`.
//!
//! Threading: the reference's `Plonk<P>` is `Send + Sync` because it is plain data.  Here the per-circuit GPU contexts
//! live behind a `Mutex`, a context is only ever used while the lock is held (the C ABI asks for one host thread per
//! context at a time), and the raw handle is wrapped in a `Send` newtype: `Plonk<P>` and `SRS<P>` stay `Send + Sync`.
#![allow(clippy::many_single_char_names)]

use crate::{
    constraints::*,
    ec::{Field, G1Point, G2Point, GTPoint, Pairing},
    ffi::*,
    poly::Poly,
};
use std::{collections::HashMap, ptr, sync::Mutex};

/// src/plonk.rs:15-26, unchanged: the type-level configuration of a PLONK instance.
pub trait PlonkTypes: PartialEq {
    type GF: Field;
    type HF: Field;
    type G1: G1Point<S = Self::GF>;
    type G2: G2Point<S = Self::GF>;
    type GT: GTPoint;
    type E: Pairing<G1 = Self::G1, G2 = Self::G2, GT = Self::GT>;
    const K1: Self::HF;
    const K2: Self::HF;
    const OMEGA: Self::HF;
    fn gf(sf: Self::HF) -> Self::GF;
}

/// Encoding of a `PlonkTypes` instance at the C ABI: one byte per field element (include/pbh_b200.h "wire layout").
/// The CUDA library implements exactly one instance, the reference's `PlonkByHandTypes` (F_17 / F_101, y^2 = x^3 + 3).
pub trait B200Wire: PlonkTypes {
    fn hf_to_byte(x: &Self::HF) -> u8;
    fn hf_from_byte(b: u8) -> Self::HF;
    fn gf_to_byte(x: &Self::GF) -> u8;
    fn g1_to_bytes(p: &Self::G1) -> (u8, u8, bool);
    fn g1_from_bytes(x: u8, y: u8, infinite: bool) -> Self::G1;
    fn g2_from_bytes(a: u8, b: u8) -> Self::G2;
}

impl B200Wire for crate::pbh::PlonkByHandTypes {
    fn hf_to_byte(x: &Self::HF) -> u8 { x.as_u64() as u8 }
    fn hf_from_byte(b: u8) -> Self::HF { crate::pbh::f17(b as u64) }
    fn gf_to_byte(x: &Self::GF) -> u8 { x.as_u64() as u8 }
    fn g1_to_bytes(p: &Self::G1) -> (u8, u8, bool) { (p.x.as_u64() as u8, p.y.as_u64() as u8, p.infinite) }
    fn g1_from_bytes(x: u8, y: u8, infinite: bool) -> Self::G1 {
        crate::pbh::g1::G1P { x: crate::pbh::f101(x as u64), y: crate::pbh::f101(y as u64), infinite }
    }
    fn g2_from_bytes(a: u8, b: u8) -> Self::G2 { crate::pbh::g2::G2P::new(crate::pbh::f101(a as u64), crate::pbh::f101(b as u64)) }
}

/// An owned `pbh_ctx`.  Only dereferenced while the `Mutex` that holds it is locked.
struct Ctx(*mut pbh_ctx);
unsafe impl Send for Ctx {}
impl Drop for Ctx {
    fn drop(&mut self) { unsafe { pbh_ctx_destroy(self.0) } }
}

fn create_ctx(circuit: &pbh_circuit, secret: u8, n: u32, omega_pows: u8, device: i32) -> Ctx {
    let mut h = ptr::null_mut();
    let rc = unsafe { pbh_ctx_create(circuit, secret, n, omega_pows, device, &mut h) };
    // PBH_ERR_SETUP_PANIC: the reference's SRS::create / Plonk::new panics on these parameters (s = 0, 17 * G2: Q12)
    assert!(rc != PBH_ERR_SETUP_PANIC, "called `Option::unwrap()` on a `None` value");
    assert_eq!(rc, PBH_OK, "pbh_ctx_create failed ({}): there is no CPU fallback", rc);
    Ctx(h)
}

/// src/plonk.rs:28-32.  The three public fields are the reference's; the private ones keep what the GPU needs.
pub struct SRS<P: PlonkTypes> {
    pub g1s: Vec<P::G1>,
    pub g2_1: P::G2,
    pub g2_s: P::G2,
    secret: u8,
    ctx: Mutex<Ctx>,
}

impl<P: PlonkTypes + B200Wire> SRS<P> {
    /// src/plonk.rs:35-48 (Q11 and Q12 of SURVEY.md included: the library reproduces the reference's powers mod 101 and
    /// its panics).  The points are computed by the library's setup and read back with `pbh_ctx_get_srs`.
    pub fn create(s: P::GF, n: usize) -> Self {
        let mut circuit = pbh_circuit::default();
        unsafe { pbh_circuit_pbh_test(&mut circuit) };
        let secret = P::gf_to_byte(&s);
        let ctx = create_ctx(&circuit, secret, n as u32, 4, 0);
        let mut raw = vec![0u8; 3 * (n + 1)];
        let (mut count, mut g2) = (0u32, [0u8; 4]);
        let rc = unsafe { pbh_ctx_get_srs(ctx.0, raw.as_mut_ptr(), n + 1, &mut count, g2.as_mut_ptr()) };
        assert_eq!(rc, PBH_OK);
        let g1s = raw.chunks(3).take(count as usize).map(|p| P::g1_from_bytes(p[0], p[1], p[2] != 0)).collect();
        SRS { g1s, g2_1: P::g2_from_bytes(g2[0], g2[1]), g2_s: P::g2_from_bytes(g2[2], g2[3]), secret, ctx: Mutex::new(ctx) }
    }

    /// src/plonk.rs:51-58: the KZG commitment sum_n g1s[n] * gf(v_n), a batch of one through `pbh_kzg_commit_batch`.
    /// Panics like the reference (index out of bounds) when the polynomial has more coefficients than the SRS has points.
    pub fn eval_at_s(&self, vs: &Poly<P::HF>) -> P::G1 {
        let coeffs = vs.coeffs();
        assert!(coeffs.len() <= self.g1s.len(), "index out of bounds: the len is {} but the index is {}", self.g1s.len(), self.g1s.len());
        assert!(coeffs.len() <= 7, "pbh_kzg_commit_batch takes the 7 coefficient planes the pbh circuit needs");
        let mut planes = [0u8; 7];
        for (k, c) in coeffs.iter().enumerate() { planes[k] = P::hf_to_byte(c); }
        let mut out = [0u8; 3];
        let ctx = self.ctx.lock().unwrap();
        let rc = unsafe { pbh_kzg_commit_batch(ctx.0, 1, planes.as_ptr(), 1, out.as_mut_ptr(), 1, 0) };
        assert_eq!(rc, PBH_OK);
        P::g1_from_bytes(out[0], out[1], out[2] != 0)
    }
}

/// src/plonk.rs:60-95: nine commitments, then seven evaluations.
#[derive(Debug, PartialEq)]
pub struct Proof<P: PlonkTypes> {
    pub a_s: P::G1,
    pub b_s: P::G1,
    pub c_s: P::G1,
    pub z_s: P::G1,
    pub t_lo_s: P::G1,
    pub t_mid_s: P::G1,
    pub t_hi_s: P::G1,
    pub w_z_s: P::G1,
    pub w_z_omega_s: P::G1,
    pub a_z: P::HF,
    pub b_z: P::HF,
    pub c_z: P::HF,
    pub s_sigma_1_z: P::HF,
    pub s_sigma_2_z: P::HF,
    pub r_z: P::HF,
    pub z_omega_z: P::HF,
}

/// src/plonk.rs:97-108: the verifier's five challenges.
pub struct Challange<P: PlonkTypes> {
    pub alpha: P::HF,
    pub beta: P::HF,
    pub gamma: P::HF,
    pub z: P::HF,
    pub v: P::HF,
}

/// src/plonk.rs:110-117.  One GPU context per circuit: the circuit-constant work the reference redoes on every call
/// (src/plonk.rs:222-243, 328-333, 506-517, 557-562) is hoisted into it at first use.
pub struct Plonk<P: PlonkTypes> {
    srs: SRS<P>,
    omega_pows: u8,
    device: i32,
    ctxs: Mutex<HashMap<[u8; 44], Ctx>>,
}

fn circuit_of<F: Field>(c: &Constrains<F>, byte: impl Fn(&F) -> u8) -> pbh_circuit {
    assert_eq!(c.q_l.len(), 4, "the reference's prove() is hard-wired to 4 gates (src/plonk.rs:376-378)");
    let mut k = pbh_circuit::default();
    let enc = |x: &CopyOf| match x {
        CopyOf::A(n) => (0u8, *n as u8),
        CopyOf::B(n) => (1, *n as u8),
        CopyOf::C(n) => (2, *n as u8),
    };
    for i in 0..4 {
        k.q_l[i] = byte(&c.q_l[i]); k.q_r[i] = byte(&c.q_r[i]); k.q_o[i] = byte(&c.q_o[i]);
        k.q_m[i] = byte(&c.q_m[i]); k.q_c[i] = byte(&c.q_c[i]);
        let (w, n) = enc(&c.c_a[i]); k.c_a_wire[i] = w; k.c_a_index[i] = n;
        let (w, n) = enc(&c.c_b[i]); k.c_b_wire[i] = w; k.c_b_index[i] = n;
        let (w, n) = enc(&c.c_c[i]); k.c_c_wire[i] = w; k.c_c_index[i] = n;
    }
    k
}

impl<P: PlonkTypes + B200Wire> Plonk<P> {
    /// src/plonk.rs:120-175.  `omega_pows` must be the reference's OMEGA = 4 (the library checks it at context creation).
    pub fn new(srs: SRS<P>, omega_pows: P::HF) -> Self {
        Plonk { srs, omega_pows: P::hf_to_byte(&omega_pows), device: 0, ctxs: Mutex::new(HashMap::new()) }
    }

    /// Addition: the CUDA device the contexts of this instance are created on (default 0).
    pub fn on_device(mut self, device: i32) -> Self { self.device = device; self }

    fn with_ctx<R>(&self, c: &Constrains<P::HF>, f: impl FnOnce(*mut pbh_ctx) -> R) -> R {
        let k = circuit_of(c, P::hf_to_byte);
        let key: [u8; 44] = unsafe { std::mem::transmute(k) };
        let mut map = self.ctxs.lock().unwrap();
        let ctx = map.entry(key).or_insert_with(|| create_ctx(&k, self.srs.secret, (self.srs.g1s.len() - 1) as u32, self.omega_pows, self.device));
        f(ctx.0)      // the lock is held for the whole call: one host thread per context at a time
    }

    /// Addition: `n` proofs at once over column-major byte planes (include/pbh_b200.h "wire layout"): `wit` 12 x n,
    /// `rand` 9 x n, `chal` 5 x n  ->  (proof 27 x n, status n).  status != 0 names the reference's panic site.
    pub fn prove_batch(&self, c: &Constrains<P::HF>, wit: &[u8], rand: &[u8], chal: &[u8], n: usize) -> (Vec<u8>, Vec<u8>) {
        assert!(wit.len() >= 12 * n && rand.len() >= 9 * n && chal.len() >= 5 * n);
        let (mut proof, mut status) = (vec![0u8; 27 * n], vec![0u8; n]);
        let rc = self.with_ctx(c, |h| unsafe {
            pbh_prove_batch(h, n, wit.as_ptr(), n, rand.as_ptr(), n, chal.as_ptr(), n, proof.as_mut_ptr(), n, status.as_mut_ptr())
        });
        assert_eq!(rc, PBH_OK);
        (proof, status)
    }

    /// Addition: `n` verifications at once: proof 27 x n, chal 5 x n, u n  ->  result bytes (PBH_VR_*).
    pub fn verify_batch(&self, c: &Constrains<P::HF>, proof: &[u8], chal: &[u8], u: &[u8], n: usize) -> Vec<u8> {
        assert!(proof.len() >= 27 * n && chal.len() >= 5 * n && u.len() >= n);
        let mut result = vec![0u8; n];
        let rc = self.with_ctx(c, |h| unsafe {
            pbh_verify_batch(h, n, proof.as_ptr(), n, chal.as_ptr(), n, u.as_ptr(), result.as_mut_ptr(), ptr::null_mut(), 0)
        });
        assert_eq!(rc, PBH_OK);
        result
    }

    /// Addition: the same two calls over PACKED records (include/pbh_b200.h "packed wire format": 16-byte inputs, 12-byte
    /// proofs, 4-byte challenge words), the form that crosses PCIe with 32 bytes up and 13 bytes down per proof +
    /// verification.  `PackedWitness::pack` / `PackedProof::unpack` below convert to and from byte planes on the host.
    pub fn prove_packed(&self, c: &Constrains<P::HF>, input: &[PackedWitness]) -> Vec<PackedProof> {
        let mut out = vec![PackedProof::default(); input.len()];
        let rc = self.with_ctx(c, |h| unsafe { pbh_prove_packed(h, input.len(), input.as_ptr(), out.as_mut_ptr()) });
        assert_eq!(rc, PBH_OK);
        out
    }
    pub fn verify_packed(&self, c: &Constrains<P::HF>, proofs: &[PackedProof], chal_u: &[u32]) -> Vec<u8> {
        assert_eq!(proofs.len(), chal_u.len());
        let mut result = vec![0u8; proofs.len()];
        let rc = self.with_ctx(c, |h| unsafe { pbh_verify_packed(h, proofs.len(), proofs.as_ptr(), chal_u.as_ptr(), result.as_mut_ptr()) });
        assert_eq!(rc, PBH_OK);
        result
    }

    /// src/plonk.rs:191-466: a batch of one; panics where (and with what) the reference panics.
    pub fn prove(
        &self,
        constraints: &Constrains<P::HF>,
        assigments: &Assigments<P::HF>,
        challange: &Challange<P>,
        rand: [P::HF; 9],
    ) -> Proof<P> {
        let b = P::hf_to_byte;
        let wit: Vec<u8> = assigments.a.iter().chain(assigments.b.iter()).chain(assigments.c.iter()).map(b).collect();
        let rnd: Vec<u8> = rand.iter().map(b).collect();
        let chal = [b(&challange.alpha), b(&challange.beta), b(&challange.gamma), b(&challange.z), b(&challange.v)];
        let (p, st) = self.prove_batch(constraints, &wit, &rnd, &chal, 1);
        match st[0] {
            PBH_ST_OK => {}
            PBH_ST_UNSATISFIED => panic!("assertion failed: constraints.satisfies(assigments)"),   // src/plonk.rs:199
            PBH_ST_ACC_DIV0 => panic!("called `Option::unwrap()` on a `None` value"),              // src/plonk.rs:297
            PBH_ST_T_REMAINDER => panic!("assertion `left == right` failed"),                      // src/plonk.rs:370
            PBH_ST_T_SLICE => panic!("range end index 18 out of range for slice"),                 // src/plonk.rs:376
            PBH_ST_SRS_OOB => panic!("index out of bounds"),                                       // src/plonk.rs:56
            s => panic!("pbh status {}", s),
        }
        let pt = |k: usize| {
            let inf = if k < 8 { (p[18] >> k) & 1 } else { p[19] & 1 } != 0;
            P::g1_from_bytes(p[2 * k], p[2 * k + 1], inf)
        };
        let ev = |k: usize| P::hf_from_byte(p[20 + k]);
        Proof {
            a_s: pt(0), b_s: pt(1), c_s: pt(2), z_s: pt(3), t_lo_s: pt(4), t_mid_s: pt(5), t_hi_s: pt(6), w_z_s: pt(7), w_z_omega_s: pt(8),
            a_z: ev(0), b_z: ev(1), c_z: ev(2), s_sigma_1_z: ev(3), s_sigma_2_z: ev(4), r_z: ev(5), z_omega_z: ev(6),
        }
    }

    /// src/plonk.rs:468-650: a batch of one; `false` for the three reject classes, a panic where the reference panics.
    pub fn verify(
        &self,
        constraints: &Constrains<P::HF>,
        proof: &Proof<P>,
        challange: &Challange<P>,
        rand: [P::HF; 1],
    ) -> bool {
        let b = P::hf_to_byte;
        let mut p = [0u8; 27];
        let points = [&proof.a_s, &proof.b_s, &proof.c_s, &proof.z_s, &proof.t_lo_s, &proof.t_mid_s, &proof.t_hi_s, &proof.w_z_s, &proof.w_z_omega_s];
        for (k, g) in points.iter().enumerate() {
            let (x, y, inf) = P::g1_to_bytes(g);
            p[2 * k] = x; p[2 * k + 1] = y;
            if inf { if k < 8 { p[18] |= 1 << k } else { p[19] |= 1 } }
        }
        let evals = [&proof.a_z, &proof.b_z, &proof.c_z, &proof.s_sigma_1_z, &proof.s_sigma_2_z, &proof.r_z, &proof.z_omega_z];
        for (k, e) in evals.iter().enumerate() { p[20 + k] = b(e); }
        let chal = [b(&challange.alpha), b(&challange.beta), b(&challange.gamma), b(&challange.z), b(&challange.v)];
        let r = self.verify_batch(constraints, &p, &chal, &[b(&rand[0])], 1)[0];
        if r == PBH_VR_PANIC_ZH0 { panic!("called `Option::unwrap()` on a `None` value") }      // src/plonk.rs:579
        r & 1 == 1
    }
}

impl PackedWitness {
    /// byte planes (12 x n, 9 x n, 5 x n, n) -> packed records; `None` when a value is >= 17 (it has no packed form)
    pub fn pack(wit: &[u8], rand: &[u8], chal: &[u8], u: &[u8], n: usize) -> Option<Vec<PackedWitness>> {
        assert!(wit.len() >= 12 * n && rand.len() >= 9 * n && chal.len() >= 5 * n && u.len() >= n);
        let mut out = vec![PackedWitness::default(); n];
        let rc = unsafe { pbh_pack_witness_host(n, wit.as_ptr(), n, rand.as_ptr(), n, chal.as_ptr(), n, u.as_ptr(), out.as_mut_ptr()) };
        if rc == PBH_OK { Some(out) } else { None }
    }
}
impl PackedProof {
    /// packed proofs -> (proof planes 27 x n, status n)
    pub fn unpack(packed: &[PackedProof]) -> (Vec<u8>, Vec<u8>) {
        let n = packed.len();
        let (mut proof, mut status) = (vec![0u8; 27 * n], vec![0u8; n]);
        let rc = unsafe { pbh_unpack_proofs_host(n, packed.as_ptr(), proof.as_mut_ptr(), n, status.as_mut_ptr()) };
        assert_eq!(rc, PBH_OK);
        (proof, status)
    }
}
