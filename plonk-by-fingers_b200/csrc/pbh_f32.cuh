// pbh_f32.cuh — the plain-float scalar policy of the FP32-pipe arithmetic (pbh_prove_f32.cuh, pbh_g1f.cuh): exact small
// integers held in floats.  The CPU test suite supplies a second policy with the same function names that propagates
// worst-case magnitudes instead of values; the arithmetic templates are instantiated with both.
#pragma once
#include <math.h>
#include <stdint.h>

#include "pbh_arith.cuh"

namespace pbh {

// ---- scalar policy: plain float -------------------------------------------------------------------------------------
struct F32 {
  float v;
  PBH_HD F32() : v(0.f) {}
  PBH_HD explicit F32(float x) : v(x) {}
};
PBH_HD F32 f_const(float c, F32*) { return F32(c); }
PBH_HD F32 f_fma(F32 a, F32 b, F32 c) { return F32(fmaf(a.v, b.v, c.v)); }
PBH_HD F32 f_mul(F32 a, F32 b) { return F32(a.v * b.v); }
PBH_HD F32 f_add(F32 a, F32 b) { return F32(a.v + b.v); }
PBH_HD F32 f_sub(F32 a, F32 b) { return F32(a.v - b.v); }
PBH_HD F32 f_red(F32 x) {
  float t = fmaf(x.v, 0.058823529411764705f, 12582912.0f);
  float q = t - 12582912.0f;
  return F32(fmaf(q, -17.0f, x.v));
}
PBH_HD bool f_is_zero(F32 x) { return x.v == 0.0f; }               // for reduced values
// true when the predicate holds for any item handled by this warp (a warp vote on the device, the item itself on the host)
PBH_HD bool f_any(bool b, F32*) {
#if defined(__CUDA_ARCH__)
  return __any_sync(__activemask(), b) != 0;
#else
  return b;
#endif
}
// canonical residue 0..16 of a centred one, as an integer (full-rate ops: compare/select, FADD, LOP3)
PBH_HD uint32_t f_canon(F32 x) {
  float c = x.v < 0.0f ? x.v + 17.0f : x.v;
#if defined(__CUDA_ARCH__)
  return (uint32_t)__float_as_int(c + 12582912.0f) & 0xFFu;
#else
  return (uint32_t)(int)c;
#endif
}
// rint(x / d) for an exact integer x (|x| < 2^21) that is never half-way between two multiples of d
PBH_HD F32 f_rint_div(F32 x, float d, float inv_d, F32*) {
  (void)d;
  float t = fmaf(x.v, inv_d, 12582912.0f);
  return F32(t - 12582912.0f);
}
// canonical residue 0..16 of a centred one, kept as a float
PBH_HD F32 f_canon_f(F32 x, F32*) { return F32(x.v < 0.0f ? x.v + 17.0f : x.v); }
// table index 0..101 of a centred residue mod 102
PBH_HD uint32_t f_index102(F32 x) {
  float c = x.v < 0.0f ? x.v + 102.0f : x.v;
#if defined(__CUDA_ARCH__)
  return (uint32_t)__float_as_int(c + 12582912.0f) & 0xFFu;
#else
  return (uint32_t)(int)c;
#endif
}
PBH_HD F32 f_from_u32(uint32_t b, F32*) {                          // exact for b < 2^23
#if defined(__CUDA_ARCH__)
  return F32(__int_as_float(0x4B000000 | (int)b) - 8388608.0f);
#else
  return F32((float)b);
#endif
}

}  // namespace pbh
