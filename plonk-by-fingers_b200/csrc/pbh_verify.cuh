// pbh_verify.cuh — one Plonk-by-hand verification per thread.
//
// Reproduces Plonk::verify (src/plonk.rs:468-650) bit for bit: the on-curve test that ignores the infinity
// flag (Q9), the missing alpha in t_z (Q3), the panic on Z_H(z) = 0 (Q4), the identity pairing value 0+0u (Q10)
// and acceptance of any of the 101 affine curve points, not only the order-17 subgroup (Q17).  The eight
// selector / permutation commitments the reference recomputes on every call (src/plonk.rs:506-517) are context
// constants.
//
// Group work, two bit-exact algorithms:
//   ARITH  real affine arithmetic per item: one Straus (shared-doubling) multi-scalar multiplication over the
//          nine proof points with 5-bit scalars, fixed-base multiples for the constant commitments, two Miller
//          loops over r = 17 and two final exponentiations in F_101^2.
//   TABLE  E(F_101) is cyclic of order 102, so a point is its discrete log to a fixed generator; G1 arithmetic
//          becomes arithmetic mod 102 and the two pairings against the fixed G2 points become 102-entry
//          tables.  All tables fit in <= 32 shared-memory banks, so divergent lookups are conflict free.
#pragma once
#include "pbh_prove.cuh"

namespace pbh {

enum : uint32_t {
  VR_ACCEPT = 0x01, VR_REJECT_PAIRING = 0x00, VR_NOT_ON_CURVE = 0x02, VR_NOT_IN_FIELD = 0x04, VR_PANIC_ZH0 = 0x10,
  VR_BAD_ENCODING = 0x20
};

// px, py: coordinates of a_s b_s c_s z_s t_lo_s t_mid_s t_hi_s w_z_s w_z_omega_s; infbits: bit k = point k's
// `infinite` flag; ev: a_z b_z c_z s_sigma_1_z s_sigma_2_z r_z z_omega_z; ch: alpha beta gamma z v; u = rand[0].
// All raw bytes.  Returns the PBH_VR_* result byte; e1/e2 are defined when the pairing check was reached.
template <int ALGO>
PBH_HD uint32_t verify_one(const uint32_t (&px_in)[9], const uint32_t (&py_in)[9], uint32_t infbits, const uint32_t (&ev)[7],
                           const uint32_t (&ch_in)[5], uint32_t u, const Consts& K, const Tables& T, GT& e1, GT& e2) {
  e1.a = e1.b = e2.a = e2.b = 0;
  // ---- encoding (see include/pbh_b200.h).  Out-of-range bytes are replaced by 0 so that no table is indexed out
  // of bounds; the verdict of such an item is PBH_VR_BAD_ENCODING whatever is computed below.
  bool bad = (infbits >> 9) != 0u || u >= 17u;
#pragma unroll
  for (int k = 0; k < 9; k++) bad = bad || px_in[k] >= 101u || py_in[k] >= 101u;
#pragma unroll
  for (int k = 0; k < 5; k++) bad = bad || ch_in[k] >= 17u;
  uint32_t px[9], py[9], ch[5];
#pragma unroll
  for (int k = 0; k < 9; k++) { px[k] = bad ? 0u : px_in[k]; py[k] = bad ? 0u : py_in[k]; }
#pragma unroll
  for (int k = 0; k < 5; k++) ch[k] = bad ? 0u : ch_in[k];
  if (bad) u = 0u;

  // ---- Step 1: in_curve on the raw coordinates (flag ignored)              src/plonk.rs:523-534
  bool off_curve = false;
  uint32_t idx[9];   // TABLE: discrete logs
  G1 pt[9];          // ARITH: points
#pragma unroll
  for (int k = 0; k < 9; k++) {
    bool inf = (infbits >> k) & 1u;
    if (ALGO == ALGO_TABLE) {
      uint32_t xi = px[k] < 127u ? px[k] : 127u;
      uint32_t ys = T.y_of_x[xi], i0 = T.idx_of_x[xi];
      bool lo = py[k] == ys, hi = (py[k] + ys == 101u);
      off_curve = off_curve || ys == 0xFFu || !(lo || hi);
      uint32_t i = lo ? i0 : 102u - i0;
      idx[k] = inf ? 0u : i;                 // a flagged point acts as the identity (g1.rs:121, 148-150)
    } else {
      off_curve = off_curve || !g1_in_curve(px[k], py[k]);
      pt[k] = inf ? g1_identity() : g1_make(px[k], py[k]);
    }
  }
  // ---- Step 2: in_field                                                     src/plonk.rs:538-547
  bool off_field = false;
#pragma unroll
  for (int k = 0; k < 7; k++) off_field = off_field || ev[k] >= 17u;

  const uint32_t alpha = ch[0], beta = ch[1], gamma = ch[2], z = ch[3], v = ch[4];
  // evaluations >= 17 are representable inputs (verdict false at Step 2); reduce them so that everything computed
  // below stays inside the table ranges
  const uint32_t a_z = mod17(ev[0]), b_z = mod17(ev[1]), c_z = mod17(ev[2]), s1_z = mod17(ev[3]), s2_z = mod17(ev[4]),
                 r_z = mod17(ev[5]), zw_z = mod17(ev[6]);

  // ---- Steps 4-7                                                            src/plonk.rs:553-579
  uint32_t z2 = mul17(z, z), z3 = mul17(z2, z), z4 = mul17(z2, z2);
  uint32_t zh_z = sub17(z4, 1u);
  uint32_t l1_z = mod17(K.L1[0] + K.L1[1] * z + K.L1[2] * z2 + K.L1[3] * z3);
  uint32_t a2 = mul17(alpha, alpha);
  uint32_t l1a2 = mul17(l1_z, a2);
  uint32_t perm_a = mod17(beta * s1_z + gamma + a_z), perm_b = mod17(beta * s2_z + gamma + b_z);
  uint32_t perm = mul17(mul17(mul17(perm_a, perm_b), mod17(c_z + gamma)), zw_z);   // Q3: no alpha here
  bool zh0 = zh_z == 0u;                                                     // Q4
  uint32_t t_z = mul17(mod17(r_z + 34u - perm - l1a2), T.inv17[zh_z]);

  // ---- scalars of Steps 8-11                                                src/plonk.rs:583-644
  uint32_t v2 = mul17(v, v), v3 = mul17(v2, v), v4 = mul17(v3, v), v5 = mul17(v4, v), v6 = mul17(v5, v);
  uint32_t z6 = mul17(z4, z2), z12 = mul17(z6, z6);
  uint32_t bz = mul17(beta, z);
  uint32_t s_qm = mul17(mul17(a_z, b_z), v), s_ql = mul17(a_z, v), s_qr = mul17(b_z, v), s_qo = mul17(c_z, v), s_qc = v;
  uint32_t s_zs = mod17(mul17(mul17(mul17(mul17(mod17(a_z + bz + gamma), mod17(b_z + 2u * bz + gamma)),
                                          mod17(c_z + 3u * bz + gamma)), alpha), v) +
                        mul17(l1a2, v) + u);
  uint32_t s_s3 = mul17(mul17(mul17(mul17(mul17(perm_a, perm_b), alpha), v), beta), zw_z);
  uint32_t s_e = mod17(t_z + v * r_z + v2 * a_z + v3 * b_z + v4 * c_z + v5 * s1_z + v6 * s2_z + u * zw_z);
  uint32_t s_wzw = mul17(mul17(u, z), 4u);   // u z omega

  if (ALGO == ALGO_TABLE) {
    // fixed-base part in the exponent of G: d_1 - d_3 + v^5 sigma_1 + v^6 sigma_2 - e
    uint32_t efix = mod17(s_qm * K.vdlog[0] + s_ql * K.vdlog[1] + s_qr * K.vdlog[2] + s_qo * K.vdlog[3] + s_qc * K.vdlog[4] +
                          v5 * K.vdlog[5] + v6 * K.vdlog[6] + neg17(s_s3) * K.vdlog[7] + neg17(s_e));
    uint32_t i1 = mod102(idx[7] + u * idx[8]);
    uint32_t i2 = mod102(z * idx[7] + s_wzw * idx[8] + idx[4] + z6 * idx[5] + z12 * idx[6] + s_zs * idx[3] + v2 * idx[0] +
                         v3 * idx[1] + v4 * idx[2] + 6u * efix);
    e1.a = T.pair_s_a[i1]; e1.b = T.pair_s_b[i1];
    e2.a = T.pair_1_a[i2]; e2.b = T.pair_1_b[i2];
  } else {
    // fixed-base multiples (the constant commitments lie in <G>, of order 17, so -[s]P = [17-s]P)
    G1 acc = g1_unpack(T.vfix_mult[0][s_qm]);
    acc = g1_add(acc, g1_unpack(T.vfix_mult[1][s_ql]), T.inv101);
    acc = g1_add(acc, g1_unpack(T.vfix_mult[2][s_qr]), T.inv101);
    acc = g1_add(acc, g1_unpack(T.vfix_mult[3][s_qo]), T.inv101);
    acc = g1_add(acc, g1_unpack(T.vfix_mult[4][s_qc]), T.inv101);
    acc = g1_add(acc, g1_unpack(T.vfix_mult[5][v5]), T.inv101);
    acc = g1_add(acc, g1_unpack(T.vfix_mult[6][v6]), T.inv101);
    acc = g1_add(acc, g1_unpack(T.vfix_mult[7][neg17(s_s3)]), T.inv101);
    acc = g1_add(acc, g1_unpack(T.vfix_mult[8][neg17(s_e)]), T.inv101);
    // Straus: sum_k [sc_k] pt_k with shared doublings, scalars < 17 (5 bits)
    const uint32_t sc[9] = {v2, v3, v4, s_zs, 1u, z6, z12, z, s_wzw};
    G1 r = g1_identity();
#pragma unroll
    for (int j = 4; j >= 0; j--) {
      if (j != 4) r = g1_add(r, r, T.inv101);
#pragma unroll
      for (int k = 0; k < 9; k++) {
        if (j == 4 && k == 4) continue;   // the scalar of t_lo_s is 1
        G1 s = g1_add(r, pt[k], T.inv101);
        if ((sc[k] >> j) & 1u) r = s;
      }
    }
    G1 q2 = g1_add(r, acc, T.inv101);                                        // e_2_q1
    G1 q1 = g1_add(pt[7], g1_smul<5>(pt[8], u, T.inv101), T.inv101);         // e_1_q1
    e1 = pairing(q1, K.g2_s[0], K.g2_s[1], T.inv101);                        // src/plonk.rs:646
    e2 = pairing(q2, K.g2_1[0], K.g2_1[1], T.inv101);                        // src/plonk.rs:647
  }

  uint32_t res = (e1.a == e2.a && e1.b == e2.b) ? VR_ACCEPT : VR_REJECT_PAIRING;
  if (zh0) res = VR_PANIC_ZH0;
  if (off_field) res = VR_NOT_IN_FIELD;
  if (off_curve) res = VR_NOT_ON_CURVE;
  if (bad) res = VR_BAD_ENCODING;
  if (res != VR_ACCEPT && res != VR_REJECT_PAIRING) e1.a = e1.b = e2.a = e2.b = 0;
  return res;
}

}  // namespace pbh
