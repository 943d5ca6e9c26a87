// pbh_arith.cuh — F_17 / F_101 / G1 / GT arithmetic for the B200 kernels.
//
// Everything is `__host__ __device__` so that context creation (host) and the kernels (device) run
// the same routines.  Values live in 32-bit registers; reductions are a low multiply by a
// reciprocal and a shift (IMAD, SHF, IMAD: IMAD.HI is half-rate on sm_100), applied lazily: sums of
// products are accumulated unreduced and reduced once per output.  Inverses come from 17- and 101-byte tables that are
// packed into <= 32 shared-memory banks, so a divergent lookup is bank-conflict free.
//
// Reference semantics reproduced here (file:line relative to the reference repository):
//   U64Field ops            src/utils/u64field.rs:107-228   (results are canonical residues)
//   G1P add / neg / mul     src/pbh/g1.rs:108-168           (affine, explicit identity (0,0,inf))
//   GTP mul / conj / pow    src/pbh/gt.rs:21-69
//   pairing_f / pairing     src/pbh/pairing.rs:12-47
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define PBH_HD __host__ __device__ __forceinline__
#else
#define PBH_HD inline
#endif

namespace pbh {

PBH_HD uint32_t umulhi32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
  return __umulhi(a, b);
#else
  return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}

// Reductions by a low multiply and a shift: floor(x/m) = (x * M) >> k with M = ceil(2^k/m) is exact while
// x * (M*m - 2^k) < 2^k, and x * M must fit 32 bits.  Not a multiply-high: IMAD.HI issues at half the IMAD rate on sm_100
// (profiles/r01_pipe_rates.txt) and the integer paths (PBH_ALGO_ARITH curve arithmetic, the int32 prover, the sweeps)
// are bound by that pipe.  The ranges below cover every argument in this code base; test builds record the largest
// argument seen (PBH_RANGE_TRACK, host compilation only) and the test suite checks it against the range.
//   mod17:  M = 61681 = ceil(2^20/17),  61681*17  - 2^20 = 1:  exact for x < 69631 (x*M < 2^32)
//   mod101: M = 41528 = ceil(2^22/101), 41528*101 - 2^22 = 24: exact for x < 103000
//   mod102: M = 41121 = ceil(2^22/102), 41121*102 - 2^22 = 38: exact for x < 104000
#if defined(PBH_RANGE_TRACK) && !defined(__CUDA_ARCH__)
static uint32_t g_mod_max[3] = {0, 0, 0};   // test builds only
#define PBH_TRACK(i, x) do { if ((x) > g_mod_max[i]) g_mod_max[i] = (x); } while (0)
#else
#define PBH_TRACK(i, x) do { } while (0)
#endif
PBH_HD uint32_t mod17(uint32_t x) { PBH_TRACK(0, x); return x - 17u * ((x * 61681u) >> 20); }
PBH_HD uint32_t mod101(uint32_t x) { PBH_TRACK(1, x); return x - 101u * ((x * 41528u) >> 22); }
PBH_HD uint32_t mod102(uint32_t x) { PBH_TRACK(2, x); return x - 102u * ((x * 41121u) >> 22); }

PBH_HD uint32_t mul17(uint32_t a, uint32_t b) { return mod17(a * b); }
PBH_HD uint32_t mul101(uint32_t a, uint32_t b) { return mod101(a * b); }
PBH_HD uint32_t neg17(uint32_t a) { return a ? 17u - a : 0u; }
PBH_HD uint32_t neg101(uint32_t a) { return a ? 101u - a : 0u; }
PBH_HD uint32_t add101(uint32_t a, uint32_t b) { uint32_t s = a + b; return s >= 101u ? s - 101u : s; }
PBH_HD uint32_t sub101(uint32_t a, uint32_t b) { return a >= b ? a - b : a + 101u - b; }
PBH_HD uint32_t add17(uint32_t a, uint32_t b) { uint32_t s = a + b; return s >= 17u ? s - 17u : s; }
PBH_HD uint32_t sub17(uint32_t a, uint32_t b) { return a >= b ? a - b : a + 17u - b; }

// Three-point fixed-base tables for PBH_ALGO_ARITH (a width-3 comb over the fixed points): entry a + 17 b + 289 c of
// triple j is [a]P_3j + [b]P_3j+1 + [c]P_3j+2, packed like pt17 (the SRS tables take CENTRED digits: entry
// (a+8) + 17 (b+8) + 289 (c+8) for a, b, c in [-8, 8], so that the FP32 prover indexes them with two FMAs on its centred
// residues instead of making each coefficient canonical first), so a commitment to L coefficients costs ceil(L/3) lookups
// and one addition fewer than that: 11 additions per proof for the nine commitments, 2 for the verifier's nine fixed
// terms.  7 x 19.2 KB: kept in global memory (L1 / L2 resident), not staged into shared memory with `Tables`.
struct FixedBaseTables {
  uint32_t srs_tri[4][4913];    // P_i = g1s[i] (identity beyond the SRS; triple 3 holds g1s[9] alone)   src/plonk.rs:51-58
  uint32_t vfix_tri[3][4913];   // P_i = q_m_s q_l_s q_r_s q_o_s q_c_s sigma_1_s sigma_2_s sigma_3_s, G    src/plonk.rs:510-517
};

// ---- tables shared by all kernels (one copy in global memory per context, staged into shared memory) ----
struct Tables {
  uint8_t inv17[32];     // inv17[a] = a^-1 mod 17, inv17[0] = 0
  uint8_t inv101[128];   // inv101[a] = a^-1 mod 101, inv101[0] = 0
  uint32_t pt17[32];     // [e]G for e < 17 as x | y<<8 | inf<<16     (G = (1,2), src/pbh/g1.rs:71-77)
  float inv17c[32];      // centred inverse mod 17 as a float, indexed by the canonical residue (pbh_prove_f32.cuh)
  float inv101c[408];    // entry i: centred inverse mod 101 of (i - 202), 0 where that is 0 mod 101 (pbh_g1f.cuh f_inv101)
  // Group-structure tables (PBH_ALGO_TABLE).  E(F_101): y^2 = x^3 + 3 is cyclic of order 102; a point's
  // index is its discrete log to a generator g102 chosen so that G = [6]g102; index 0 is the identity.
  uint8_t y_of_x[128];   // the root y <= 50 of x^3 + 3, 0xFF when x^3 + 3 is a non-residue
  uint8_t idx_of_x[128]; // index of (x, y_of_x[x])
  uint8_t x_of_idx[128], y_of_idx[128];
  uint8_t pair_s_a[128], pair_s_b[128];   // pairing(point idx, g2_s)   (src/plonk.rs:642, 646)
  uint8_t pair_1_a[128], pair_1_b[128];   // pairing(point idx, g2_1)   (src/plonk.rs:644, 647)
  // Fixed-base multiples (PBH_ALGO_ARITH): [k]P for k < 17, packed like pt17.
  uint32_t srs_mult[10][17];              // P = g1s[i]                 (src/plonk.rs:51-58)
  uint32_t vfix_mult[9][17];              // q_m_s q_l_s q_r_s q_o_s q_c_s sigma_1_s sigma_2_s sigma_3_s, G
  const FixedBaseTables* fixed;           // device (or, in host builds, host) address of the three-point tables
};

// read-only lookup in a global-memory table (through the read-only data cache on the device)
PBH_HD uint32_t pair_lookup(const uint32_t* table, uint32_t index) {
#if defined(__CUDA_ARCH__)
  return __ldg(table + index);
#else
  return table[index];
#endif
}

// ---- G1: y^2 = x^3 + 3 over F_101 ---------------------------------------------------------------
struct G1 {
  uint32_t x, y, inf;   // canonical identity: (0, 0, 1)
};
PBH_HD G1 g1_identity() { G1 r; r.x = 0; r.y = 0; r.inf = 1; return r; }
PBH_HD G1 g1_make(uint32_t x, uint32_t y) { G1 r; r.x = x; r.y = y; r.inf = 0; return r; }
PBH_HD uint32_t g1_pack(const G1& p) { return p.x | (p.y << 8) | (p.inf << 16); }
PBH_HD G1 g1_unpack(uint32_t w) { G1 r; r.x = w & 0xFF; r.y = (w >> 8) & 0xFF; r.inf = (w >> 16) & 1; return r; }
// src/pbh/g1.rs:63-65 — ignores the infinity flag (Q9)
PBH_HD bool g1_in_curve(uint32_t x, uint32_t y) { return mod101(y * y) == mod101(mod101(x * x) * x + 3u); }
// src/pbh/g1.rs:108-117
PBH_HD G1 g1_neg(const G1& p) { G1 r = p; r.y = neg101(p.y); return r; }

// Complete affine addition, src/pbh/g1.rs:119-144 (Q16), branch-free.  `bad` is set when the reference
// would panic with "cannot add" (equal x, y neither equal nor opposite: impossible for curve points).
PBH_HD G1 g1_add(const G1& p, const G1& q, const uint8_t* inv101, bool* bad = nullptr) {
  bool same_x = p.x == q.x;
  bool opposite = same_x && (add101(p.y, q.y) == 0u);   // self == -rhs  (also doubling a y = 0 point)
  bool same = same_x && (p.y == q.y);
  // slope: 3x^2 / 2y when doubling, (y2 - y1)/(x2 - x1) otherwise
  uint32_t num = same ? mod101(3u * p.x * p.x) : sub101(q.y, p.y);
  uint32_t den = same ? mod101(2u * p.y) : sub101(q.x, p.x);
  uint32_t lambda = mul101(num, inv101[den]);
  uint32_t x3 = mod101(lambda * lambda + 202u - p.x - q.x);
  uint32_t y3 = mod101(lambda * (p.x + 101u - x3) + 101u - p.y);
  G1 r;
  r.x = x3; r.y = y3; r.inf = 0;
  if (opposite) r = g1_identity();
  if (q.inf) r = p;
  if (p.inf) r = q;
  if (bad) *bad = !p.inf && !q.inf && same_x && !opposite && !same;
  return r;
}

// [k]P for k < 2^BITS by MSB-first double-and-add, branch-free.  Equals src/pbh/g1.rs:146-168 for every
// input because both compute the group multiple and emit the canonical identity.
template <int BITS>
PBH_HD G1 g1_smul(const G1& p, uint32_t k, const uint8_t* inv101) {
  G1 r = g1_identity();
#pragma unroll
  for (int j = BITS - 1; j >= 0; j--) {
    r = g1_add(r, r, inv101);
    G1 t = g1_add(r, p, inv101);
    if ((k >> j) & 1u) r = t;
  }
  if (p.inf) r = g1_identity();
  return r;
}

// ---- GT = F_101[u]/(u^2 + 2) ---------------------------------------------------------------------
struct GT { uint32_t a, b; };
// src/pbh/gt.rs:61-69:  (a + bu)(c + du) = (ac - 2bd) + (ad + bc)u;  -2 = 99 mod 101
PBH_HD GT gt_mul(const GT& p, const GT& q) {
  GT r;
  r.a = mod101(p.a * q.a + 99u * mod101(p.b * q.b));
  r.b = mod101(p.a * q.b + p.b * q.a);
  return r;
}
PBH_HD GT gt_conj(const GT& p) { GT r; r.a = p.a; r.b = neg101(p.b); return r; }   // src/pbh/gt.rs:21-29 (Q13)

// x^600 = (x^100)^6 with x^100 = conj(x)/x = conj(x)^2 / norm(x), norm = a^2 + 2b^2.  Equals GTP::pow(600)
// (src/pbh/gt.rs:33-59) on all of F_101^2: conj is the Frobenius x -> x^101, and 0 maps to 0 (inv101[0] = 0).
PBH_HD GT gt_final_exp(const GT& f, const uint8_t* inv101) {
  uint32_t norm = mod101(f.a * f.a + 2u * mod101(f.b * f.b));
  uint32_t ninv = inv101[norm];
  GT c = gt_conj(f);
  GT c2 = gt_mul(c, c);
  GT x; x.a = mul101(c2.a, ninv); x.b = mul101(c2.b, ninv);   // x = f^100
  GT x2 = gt_mul(x, x);
  GT x3 = gt_mul(x2, x);
  return gt_mul(x3, x3);
}

// The line through a and b evaluated at Q = (qa, qb*u): src/pbh/pairing.rs:25-34, 41, 45.  Uses the raw
// coordinates, so an identity operand contributes (0, 0) exactly like the reference (Q10).
PBH_HD GT miller_line(const G1& a, const G1& b, uint32_t qa, uint32_t qb) {
  uint32_t m = sub101(b.x, a.x);
  uint32_t n = sub101(b.y, a.y);
  // x = n, y = -m, c = m*a.y - n*a.x
  uint32_t c = mod101(m * a.y + (101u * 101u) - n * a.x);
  GT r;
  r.a = mod101(qa * n + c);
  r.b = mod101(qb * neg101(m));
  return r;
}

// pairing_f(17, P, Q): the recursion of src/pbh/pairing.rs:23-47 unrolled for r = 17:
//   f = ((((L(P,-2P))^2 L(2P,-4P))^2 L(4P,-8P))^2 L(8P,-16P)) L(16P,P)
// Only P, 2P, 4P, 8P, 16P are needed (4 doublings) where the reference performs 13 scalar multiplications.
PBH_HD GT miller_f17(const G1& p_in, uint32_t qa, uint32_t qb, const uint8_t* inv101) {
  G1 p = p_in;
  if (p.inf) p = g1_identity();   // G1P * k returns the canonical identity for a flagged point (g1.rs:148-150)
  G1 p2 = g1_add(p, p, inv101);
  G1 p4 = g1_add(p2, p2, inv101);
  G1 p8 = g1_add(p4, p4, inv101);
  G1 p16 = g1_add(p8, p8, inv101);
  GT f = miller_line(p, g1_neg(p2), qa, qb);
  f = gt_mul(gt_mul(f, f), miller_line(p2, g1_neg(p4), qa, qb));
  f = gt_mul(gt_mul(f, f), miller_line(p4, g1_neg(p8), qa, qb));
  f = gt_mul(gt_mul(f, f), miller_line(p8, g1_neg(p16), qa, qb));
  f = gt_mul(f, miller_line(p16, p_in, qa, qb));   // the reference passes the caller's P itself here (pairing.rs:40)
  return f;
}
// src/pbh/pairing.rs:12-20
PBH_HD GT pairing(const G1& p, uint32_t qa, uint32_t qb, const uint8_t* inv101) {
  return gt_final_exp(miller_f17(p, qa, qb, inv101), inv101);
}

}  // namespace pbh
