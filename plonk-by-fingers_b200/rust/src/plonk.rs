//! The reference's `plonk.rs` surface over the CUDA library.  SOURCE ONLY: never compiled in this repository.
//! Replaces: SRS::create (src/plonk.rs:35-48), Plonk::new (:120-175), Plonk::prove (:191-466), Plonk::verify (:468-650).
use crate::constraints::{Assigments, Constrains, CopyOf};
use crate::ffi::*;
use crate::pbh::{g1::G1P, F17, PlonkByHandTypes};
use std::{cell::RefCell, collections::HashMap, ptr};

pub struct SRS { pub s: u8, pub n: u32 }
impl SRS {
    pub fn create(s: crate::pbh::F101, n: usize) -> Self { SRS { s: s.as_u64() as u8, n: n as u32 } }
}

#[derive(Debug, PartialEq)]
pub struct Proof {
    pub a_s: G1P, pub b_s: G1P, pub c_s: G1P, pub z_s: G1P, pub t_lo_s: G1P, pub t_mid_s: G1P, pub t_hi_s: G1P,
    pub w_z_s: G1P, pub w_z_omega_s: G1P,
    pub a_z: F17, pub b_z: F17, pub c_z: F17, pub s_sigma_1_z: F17, pub s_sigma_2_z: F17, pub r_z: F17, pub z_omega_z: F17,
}
pub struct Challange { pub alpha: F17, pub beta: F17, pub gamma: F17, pub z: F17, pub v: F17 }

/// One `pbh_ctx` per (circuit, SRS): the circuit-constant work the reference redoes per call is hoisted into it.
pub struct Plonk { srs: SRS, omega_pows: u8, device: i32, ctxs: RefCell<HashMap<[u8; 44], *mut pbh_ctx>> }

fn circuit_of(c: &Constrains<F17>) -> pbh_circuit {
    assert_eq!(c.q_l.len(), 4, "this build mirrors the reference's prove(), hard-wired to 4 gates (src/plonk.rs:376-378)");
    let mut k = pbh_circuit::default();
    for i in 0..4 {
        k.q_l[i] = c.q_l[i].as_u64() as u8; k.q_r[i] = c.q_r[i].as_u64() as u8; k.q_o[i] = c.q_o[i].as_u64() as u8;
        k.q_m[i] = c.q_m[i].as_u64() as u8; k.q_c[i] = c.q_c[i].as_u64() as u8;
        let enc = |x: &CopyOf| match x { CopyOf::A(n) => (0u8, *n as u8), CopyOf::B(n) => (1, *n as u8), CopyOf::C(n) => (2, *n as u8) };
        let (w, n) = enc(&c.c_a[i]); k.c_a_wire[i] = w; k.c_a_index[i] = n;
        let (w, n) = enc(&c.c_b[i]); k.c_b_wire[i] = w; k.c_b_index[i] = n;
        let (w, n) = enc(&c.c_c[i]); k.c_c_wire[i] = w; k.c_c_index[i] = n;
    }
    k
}

impl Plonk {
    pub fn new(srs: SRS, omega_pows: F17) -> Self {
        Plonk { srs, omega_pows: omega_pows.as_u64() as u8, device: 0, ctxs: RefCell::new(HashMap::new()) }
    }
    fn ctx(&self, c: &Constrains<F17>) -> *mut pbh_ctx {
        let k = circuit_of(c);
        let key: [u8; 44] = unsafe { std::mem::transmute(k) };
        *self.ctxs.borrow_mut().entry(key).or_insert_with(|| {
            let mut h = ptr::null_mut();
            let rc = unsafe { pbh_ctx_create(&k, self.srs.s, self.srs.n, self.omega_pows, self.device, &mut h) };
            // PBH_ERR_SETUP_PANIC: the reference's SRS::create / Plonk::new panics on these parameters (e.g. s = 0, Q12)
            assert_eq!(rc, PBH_OK, "pbh_ctx_create failed: {}", rc);
            h
        })
    }

    /// Column-major byte planes, one item per column (include/pbh_b200.h "wire layout").
    pub fn prove_batch(&self, c: &Constrains<F17>, wit: &[u8], rand: &[u8], chal: &[u8], n: usize) -> (Vec<u8>, Vec<u8>) {
        let (mut proof, mut status) = (vec![0u8; 27 * n], vec![0u8; n]);
        let rc = unsafe { pbh_prove_batch(self.ctx(c), n, wit.as_ptr(), n, rand.as_ptr(), n, chal.as_ptr(), n, proof.as_mut_ptr(), n, status.as_mut_ptr()) };
        assert_eq!(rc, PBH_OK);
        (proof, status)
    }
    pub fn verify_batch(&self, c: &Constrains<F17>, proof: &[u8], chal: &[u8], u: &[u8], n: usize) -> Vec<u8> {
        let mut result = vec![0u8; n];
        let rc = unsafe { pbh_verify_batch(self.ctx(c), n, proof.as_ptr(), n, chal.as_ptr(), n, u.as_ptr(), result.as_mut_ptr(), ptr::null_mut(), 0) };
        assert_eq!(rc, PBH_OK);
        result
    }

    /// src/plonk.rs:191-197 — batch of one; panics where the reference panics.
    pub fn prove(&self, c: &Constrains<F17>, a: &Assigments<F17>, ch: &Challange, rand: [F17; 9]) -> Proof {
        let b = |x: &F17| x.as_u64() as u8;
        let wit: Vec<u8> = a.a.iter().chain(a.b.iter()).chain(a.c.iter()).map(b).collect();
        let rnd: Vec<u8> = rand.iter().map(b).collect();
        let chal = [b(&ch.alpha), b(&ch.beta), b(&ch.gamma), b(&ch.z), b(&ch.v)];
        let (p, st) = self.prove_batch(c, &wit, &rnd, &chal, 1);
        match st[0] {
            0 => {}
            1 => panic!("assertion failed: constraints.satisfies(assigments)"),            // src/plonk.rs:199
            2 => panic!("called `Option::unwrap()` on a `None` value"),                     // src/plonk.rs:297
            3 => panic!("assertion failed: `(left == right)`"),                             // src/plonk.rs:370
            4 => panic!("range end index 18 out of range for slice"),                       // src/plonk.rs:376
            5 => panic!("index out of bounds"),                                             // src/plonk.rs:56
            s => panic!("pbh status {}", s),
        }
        let pt = |k: usize| { let inf = if k < 8 { (p[18] >> k) & 1 } else { p[19] & 1 } != 0;
                              G1P { x: crate::pbh::f101(p[2 * k] as u64), y: crate::pbh::f101(p[2 * k + 1] as u64), infinite: inf } };
        let ev = |k: usize| crate::pbh::f17(p[20 + k] as u64);
        Proof { a_s: pt(0), b_s: pt(1), c_s: pt(2), z_s: pt(3), t_lo_s: pt(4), t_mid_s: pt(5), t_hi_s: pt(6), w_z_s: pt(7), w_z_omega_s: pt(8),
                a_z: ev(0), b_z: ev(1), c_z: ev(2), s_sigma_1_z: ev(3), s_sigma_2_z: ev(4), r_z: ev(5), z_omega_z: ev(6) }
    }

    /// src/plonk.rs:468-474 — batch of one.
    pub fn verify(&self, c: &Constrains<F17>, proof: &Proof, ch: &Challange, rand: [F17; 1]) -> bool {
        let b = |x: &F17| x.as_u64() as u8;
        let mut p = [0u8; 27];
        for (k, g) in [&proof.a_s, &proof.b_s, &proof.c_s, &proof.z_s, &proof.t_lo_s, &proof.t_mid_s, &proof.t_hi_s, &proof.w_z_s, &proof.w_z_omega_s].iter().enumerate() {
            p[2 * k] = g.x.as_u64() as u8; p[2 * k + 1] = g.y.as_u64() as u8;
            if g.infinite { if k < 8 { p[18] |= 1 << k } else { p[19] |= 1 } }
        }
        for (k, e) in [&proof.a_z, &proof.b_z, &proof.c_z, &proof.s_sigma_1_z, &proof.s_sigma_2_z, &proof.r_z, &proof.z_omega_z].iter().enumerate() { p[20 + k] = b(e); }
        let chal = [b(&ch.alpha), b(&ch.beta), b(&ch.gamma), b(&ch.z), b(&ch.v)];
        let r = self.verify_batch(c, &p, &chal, &[b(&rand[0])], 1)[0];
        if r == 0x10 { panic!("called `Option::unwrap()` on a `None` value") }              // src/plonk.rs:579
        r & 1 == 1
    }
}

impl Drop for Plonk {
    fn drop(&mut self) { for (_, h) in self.ctxs.borrow_mut().drain() { unsafe { pbh_ctx_destroy(h) } } }
}
