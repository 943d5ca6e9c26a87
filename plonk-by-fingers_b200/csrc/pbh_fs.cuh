// pbh_fs.cuh — challenge sources of the fused prover.
//
// The reference's prove receives its Challange from the caller (src/plonk.rs:195, 201-206).  The per-item routines ask
// a "challenge source" for each challenge at the point where the reference first uses it:
//   beta, gamma  after the round-1 commitments            (first use src/plonk.rs:283)
//   alpha        after the commitment to z                (first use src/plonk.rs:343)
//   z            after the commitments to t_lo, t_mid, t_hi (first use src/plonk.rs:393)
//   v            after the seven evaluations              (first use src/plonk.rs:430)
//   u            after the two opening commitments        (verifier's rand[0], src/plonk.rs:519)
// FixedChal returns the caller's values (the reference's behaviour, and the same instructions as before this header
// existed); FsChal derives them from the SHA-256 transcript specified in include/pbh_b200.h.
#pragma once
#include "pbh_sha256.cuh"

namespace pbh {

template <class T>
struct FixedChal {
  static constexpr bool kNeedsPoints = false;
  const T (&ch)[5];   // alpha beta gamma z v
  PBH_HD explicit FixedChal(const T (&c)[5]) : ch(c) {}
  PBH_HD void beta_gamma(uint32_t, uint32_t, uint32_t, T& beta, T& gamma) { beta = ch[1]; gamma = ch[2]; }
  PBH_HD T alpha(uint32_t) { return ch[0]; }
  PBH_HD T zeta(uint32_t, uint32_t, uint32_t) { return ch[3]; }
  PBH_HD T v(const uint32_t (&)[7]) { return ch[4]; }
  PBH_HD void u(uint32_t, uint32_t) {}
};

// transcript word of a packed point (x | y << 8 | inf << 16): the bytes x, y, infinite, 0 read big-endian
PBH_HD uint32_t fs_point_word(uint32_t packed) {
  return ((packed & 0xFFu) << 24) | ((packed & 0xFF00u) << 8) | ((packed >> 8) & 0x100u);
}

// Conv turns a canonical residue (0..16) into the routine's scalar type; OUTLINE_MASK selects, per transcript step, the
// shared out-of-line copy of the compression (pbh_sha256.cuh)
template <class T, class Conv, int OUTLINE_MASK = 0>
struct FsChal {
  // bit k of OUTLINE_MASK: step k (beta/gamma, alpha, z, v, u) calls the shared out-of-line compression
  template <int STEP> static constexpr bool outlined() { return (OUTLINE_MASK >> STEP) & 1; }
  static constexpr bool kNeedsPoints = true;
  uint32_t st[8];
  uint32_t derived[6];   // alpha beta gamma z v u
  bool want_u;           // the prover does not use u: the fifth compression is skipped when nobody asks for it
  PBH_HD explicit FsChal(const uint32_t (&seed)[8], bool want_u_ = true) : want_u(want_u_) {
#pragma unroll
    for (int i = 0; i < 8; i++) st[i] = seed[i];
#pragma unroll
    for (int i = 0; i < 6; i++) derived[i] = 0u;
  }
  PBH_HD void beta_gamma(uint32_t pa, uint32_t pb, uint32_t pc, T& beta, T& gamma) {
    const uint32_t m[6] = {fs_point_word(pa), fs_point_word(pb), fs_point_word(pc), 0u, 0u, 0u};
    sha256_absorb<outlined<0>()>(st, m, 12);
    derived[1] = sha256_squeeze17(st, 0); derived[2] = sha256_squeeze17(st, 1);
    beta = Conv()(derived[1]); gamma = Conv()(derived[2]);
  }
  PBH_HD T alpha(uint32_t pz) {
    const uint32_t m[6] = {fs_point_word(pz), 0u, 0u, 0u, 0u, 0u};
    sha256_absorb<outlined<1>()>(st, m, 4);
    derived[0] = sha256_squeeze17(st, 0);
    return Conv()(derived[0]);
  }
  PBH_HD T zeta(uint32_t lo, uint32_t mid, uint32_t hi) {
    const uint32_t m[6] = {fs_point_word(lo), fs_point_word(mid), fs_point_word(hi), 0u, 0u, 0u};
    sha256_absorb<outlined<2>()>(st, m, 12);
    derived[3] = sha256_squeeze17(st, 0);
    return Conv()(derived[3]);
  }
  PBH_HD T v(const uint32_t (&ev)[7]) {
    const uint32_t m[6] = {(ev[0] << 24) | (ev[1] << 16) | (ev[2] << 8) | ev[3], (ev[4] << 24) | (ev[5] << 16) | (ev[6] << 8), 0u, 0u, 0u, 0u};
    sha256_absorb<outlined<3>()>(st, m, 7);
    derived[4] = sha256_squeeze17(st, 0);
    return Conv()(derived[4]);
  }
  PBH_HD void u(uint32_t wz, uint32_t wzw) {
    if (!want_u) return;
    const uint32_t m[6] = {fs_point_word(wz), fs_point_word(wzw), 0u, 0u, 0u, 0u};
    sha256_absorb<outlined<4>()>(st, m, 8);
    derived[5] = sha256_squeeze17(st, 0);
  }
};

struct ConvU32 { PBH_HD uint32_t operator()(uint32_t x) const { return x; } };

}  // namespace pbh
