// pbh_packed.cuh — the PACKED wire format of include/pbh_b200.h ("packed records"): codec and transposition kernels.
//
// The byte-plane layout spends one byte per field element; across PCIe that is 59 bytes up and 29 bytes down per
// proof + verification, and the host link - not the GPU - bounds the end-to-end rate.  The packed format carries the same
// information in radix form: 27 F_17 values in four 32-bit words (7 base-17 digits per word), the nine G1 points of a proof
// as nine base-102 digits of one 64-bit word (E(F_101): y^2 = x^3 + 3 has exactly 101 finite points, so a point that is on
// the curve - every point a prover emits - is one of 102 codes), the seven evaluations as seven base-17 digits plus a 3-bit
// status in one 32-bit word.  16 bytes per prover input, 12 bytes per proof, 4 bytes of challenges + u per verification.
//
// The codec functions are __host__ __device__: the kernels below run them on the device next to the plane kernels; the
// pbh_*_host helpers of the C ABI run the very same functions on the CPU so that a host program can convert between its own
// structures and packed records (a format conversion, not a compute path: nothing here proves or verifies anything).
#pragma once
#include <cstddef>
#include <cstdint>

#include "../../include/pbh_b200.h"

#if defined(__CUDACC__)
#define PBH_PK_HD __host__ __device__ __forceinline__
#else
#define PBH_PK_HD inline
#endif

namespace pbh {

// Point codes: 0 = the identity (0, 0, infinite); code c >= 1 = the c-th finite point of y^2 = x^3 + 3 over F_101 in
// increasing (x, y) order.  Built at compile time; 102 codes in all (the group order).
struct PackedTables {
  uint8_t pt_x[104], pt_y[104];   // code -> coordinates
  uint8_t first[104];             // x -> code of (x, smaller y); 0 when x^3 + 3 is not a square
  uint8_t ysm[104];               // x -> the smaller of the two y (<= 50); 0xFF when x^3 + 3 is not a square
  int count;
};
constexpr PackedTables make_packed_tables() {
  PackedTables t{};
  int code = 1;
  for (int x = 0; x < 104; x++) { t.first[x] = 0; t.ysm[x] = 0xFF; t.pt_x[x] = 0; t.pt_y[x] = 0; }
  for (int x = 0; x < 101; x++) {
    const int rhs = (x * x * x + 3) % 101;
    int y0 = -1;
    for (int y = 0; y <= 50; y++)
      if ((y * y) % 101 == rhs) { y0 = y; break; }
    if (y0 < 0) continue;
    t.first[x] = (uint8_t)code;
    t.ysm[x] = (uint8_t)y0;
    t.pt_x[code] = (uint8_t)x; t.pt_y[code] = (uint8_t)y0; code++;
    if (y0 != 0) { t.pt_x[code] = (uint8_t)x; t.pt_y[code] = (uint8_t)(101 - y0); code++; }
  }
  t.count = code;
  return t;
}
static constexpr PackedTables kPackedTablesHost = make_packed_tables();
static_assert(kPackedTablesHost.count == 102, "E(F_101) has 101 finite points");
static_assert(kPackedTablesHost.pt_x[kPackedTablesHost.first[1]] == 1 && kPackedTablesHost.ysm[1] == 2, "the generator (1, 2) is on the curve");

constexpr uint32_t kPow17_5 = 1419857u, kPow17_6 = 24137569u, kPow17_7 = 410338673u;
constexpr uint32_t kStatusShift = 29, kEvalMask = (1u << 29) - 1;
constexpr uint32_t kPow102_4 = 108243216u;   // 102^4

PBH_PK_HD uint32_t div17(uint32_t x) {
#if defined(__CUDA_ARCH__)
  return __umulhi(x, 0xF0F0F0F1u) >> 4;
#else
  return x / 17u;
#endif
}

// word -> `digits` base-17 digits, least significant first; whatever is left above the last digit is folded into it and
// saturates at 255, so that a word outside the format decodes to a byte >= 17 (which the plane kernels report as
// PBH_ST_BAD_ENCODING / PBH_VR_BAD_ENCODING / PBH_VR_NOT_IN_FIELD exactly as for a byte-plane input)
template <int DIGITS>
PBH_PK_HD void digits17(uint32_t x, uint8_t* v) {
#pragma unroll
  for (int k = 0; k < DIGITS - 1; k++) {
    const uint32_t q = div17(x);
    v[k] = (uint8_t)(x - 17u * q);
    x = q;
  }
  v[DIGITS - 1] = (uint8_t)(x > 255u ? 255u : x);
}

// pbh_packed_witness -> wit[12] rand[9] chal[5] u (27 values)
PBH_PK_HD void unpack_witness_item(const uint32_t w[4], uint8_t v[27]) {
  digits17<7>(w[0], v);
  digits17<7>(w[1], v + 7);
  digits17<7>(w[2], v + 14);
  digits17<6>(w[3], v + 21);
}
// 27 values (each < 17) -> four words; false when a value is not representable
PBH_PK_HD bool pack_witness_item(const uint8_t v[27], uint32_t w[4]) {
  bool ok = true;
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const int nd = j < 3 ? 7 : 6;
    uint32_t x = 0;
    for (int k = nd - 1; k >= 0; k--) {
      ok = ok && v[7 * j + k] < 17;
      x = x * 17u + v[7 * j + k];
    }
    w[j] = x;
  }
  return ok;
}
// the challenges + u word of pbh_verify_packed: digits alpha beta gamma z v u
PBH_PK_HD void unpack_chal_u(uint32_t w, uint8_t v[6]) { digits17<6>(w, v); }
PBH_PK_HD bool pack_chal_u(const uint8_t v[6], uint32_t* w) {
  bool ok = true;
  uint32_t x = 0;
  for (int k = 5; k >= 0; k--) { ok = ok && v[k] < 17; x = x * 17u + v[k]; }
  *w = x;
  return ok;
}

PBH_PK_HD uint32_t status_to_code(uint8_t status) {
  return status <= 5 ? status : status == PBH_ST_BAD_ENCODING ? 7u : 6u;
}
PBH_PK_HD uint8_t code_to_status(uint32_t c) { return c <= 5 ? (uint8_t)c : c == 7 ? (uint8_t)PBH_ST_BAD_ENCODING : (uint8_t)PBH_ST_UNREPRESENTABLE; }

// 27 proof bytes (plane order) + status -> 12-byte packed proof (points_lo, points_hi, evals_status).  A proof that the
// format cannot express (a point off the curve, a flagged identity with coordinates, an evaluation >= 17, undefined flag
// bits) gets status code 6; whenever the status code is not 0 the two payload words are zero.
PBH_PK_HD void pack_proof_item(const PackedTables& T, const uint8_t p[27], uint8_t status, uint32_t out[3]) {
  bool ok = (p[PBH_PROOF_INF_PLANE + 1] & 0xFEu) == 0;
  uint32_t code_of[9];
#pragma unroll
  for (int k = 8; k >= 0; k--) {
    const uint32_t x = p[2 * k], y = p[2 * k + 1];
    const bool inf = k < 8 ? ((p[PBH_PROOF_INF_PLANE] >> k) & 1u) : (p[PBH_PROOF_INF_PLANE + 1] & 1u);
    uint32_t code = 0;
    if (inf) {
      ok = ok && x == 0 && y == 0;
    } else {
      const uint32_t xs = x < 101 ? x : 101;          // first[101..103] = 0, ysm = 0xFF
      const uint32_t y0 = T.ysm[xs];
      const bool lo = y == y0, hi = y0 != 0 && y0 != 0xFF && y == 101 - y0;
      ok = ok && y0 != 0xFF && (lo || hi);
      code = T.first[xs] + (hi ? 1u : 0u);
    }
    code_of[k] = code;
  }
  // digits 0..3 and 4..7 by 32-bit Horner, then P = a + 102^4 (b + 102^4 c)
  const uint32_t a = ((code_of[3] * 102u + code_of[2]) * 102u + code_of[1]) * 102u + code_of[0];
  const uint32_t b = ((code_of[7] * 102u + code_of[6]) * 102u + code_of[5]) * 102u + code_of[4];
  uint64_t P = ((uint64_t)code_of[8] * kPow102_4 + b) * kPow102_4 + a;
  uint32_t E = 0;
#pragma unroll
  for (int k = 6; k >= 0; k--) {
    const uint32_t e = p[PBH_PROOF_EVAL_PLANE + k];
    ok = ok && e < 17;
    E = E * 17u + e;
  }
  uint32_t sc = status_to_code(status);
  if (sc == 0 && !ok) sc = 6;
  if (sc != 0) { P = 0; E = 0; }
  out[0] = (uint32_t)P;
  out[1] = (uint32_t)(P >> 32);
  out[2] = E | (sc << kStatusShift);
}
// 12-byte packed proof -> 27 proof bytes + status.  Status code != 0: the all-zero proof (what pbh_prove_batch leaves for an
// item that did not produce one; the verifier answers PBH_VR_NOT_ON_CURVE for it).  A ninth digit >= 102 decodes to the
// point (0, 0) without the flag, which is not on the curve either.
PBH_PK_HD void unpack_proof_item(const PackedTables& T, const uint32_t in[3], uint8_t p[27], uint8_t* status) {
  const uint32_t sc = in[2] >> kStatusShift;
  *status = code_to_status(sc);
#pragma unroll
  for (int k = 0; k < 27; k++) p[k] = 0;
  if (sc != 0) return;
  // P = a + 102^4 (b + 102^4 c): two 64-bit divisions by 102^4, then 32-bit digit extraction (a, b < 102^4 < 2^27);
  // c = floor(P / 102^8) is the ninth digit, whatever its size
  const uint64_t P = (uint64_t)in[0] | ((uint64_t)in[1] << 32);
  const uint64_t q1 = P / kPow102_4;
  uint32_t a = (uint32_t)(P - q1 * kPow102_4);
  const uint64_t q2 = q1 / kPow102_4;
  uint32_t b = (uint32_t)(q1 - q2 * kPow102_4);
  uint32_t flags = 0;
#pragma unroll
  for (int k = 0; k < 9; k++) {
    uint32_t d;
    if (k < 4) { const uint32_t q = a / 102u; d = a - q * 102u; a = q; }
    else if (k < 8) { const uint32_t q = b / 102u; d = b - q * 102u; b = q; }
    else d = q2 >= 102u ? 102u : (uint32_t)q2;
    if (d == 0) flags |= 1u << k;
    const uint32_t c = d < 102 ? d : 0;
    p[2 * k] = T.pt_x[c];
    p[2 * k + 1] = T.pt_y[c];
  }
  p[PBH_PROOF_INF_PLANE] = (uint8_t)(flags & 0xFFu);
  p[PBH_PROOF_INF_PLANE + 1] = (uint8_t)(flags >> 8);
  digits17<7>(in[2] & kEvalMask, p + PBH_PROOF_EVAL_PLANE);
}

#if defined(__CUDACC__)
__constant__ PackedTables kPackedTablesDev = make_packed_tables();

__device__ __forceinline__ void stage_packed_tables(PackedTables* sT) {
  static_assert(sizeof(PackedTables) % 4 == 0, "word copy");
  const uint32_t* src = reinterpret_cast<const uint32_t*>(&kPackedTablesDev);
  uint32_t* dst = reinterpret_cast<uint32_t*>(sT);
  for (int i = threadIdx.x; i < (int)(sizeof(PackedTables) / 4); i += blockDim.x) dst[i] = src[i];
  __syncthreads();
}

// FOUR consecutive items per thread: the packed side moves as 128-bit words, the plane side as one 32-bit word per plane
// (byte lane j = item 4q + j; a warp writes 128 contiguous bytes per plane).  `vec` = every plane base and pitch is a multiple
// of 4 (the host decides); otherwise, and for the ragged last group, the bytes are stored one by one.  Null plane pointers are
// skipped.
__device__ __forceinline__ void store_plane_group(uint8_t* plane, size_t i0, size_t n, uint32_t word, bool vec) {
  if (vec && i0 + 4 <= n) {
    *reinterpret_cast<uint32_t*>(plane + i0) = word;
  } else {
#pragma unroll
    for (int j = 0; j < 4; j++)
      if (i0 + j < n) plane[i0 + j] = (uint8_t)(word >> (8 * j));
  }
}
__device__ __forceinline__ uint32_t load_plane_group(const uint8_t* plane, size_t i0, size_t n, bool vec) {
  if (vec && i0 + 4 <= n) return *reinterpret_cast<const uint32_t*>(plane + i0);
  uint32_t word = 0;
#pragma unroll
  for (int j = 0; j < 4; j++)
    if (i0 + j < n) word |= (uint32_t)plane[i0 + j] << (8 * j);
  return word;
}

__global__ void __launch_bounds__(256) unpack_witness_kernel(size_t n, const pbh_packed_witness* __restrict__ in,
                                                             uint8_t* __restrict__ wit, size_t wit_pitch, uint8_t* __restrict__ rnd,
                                                             size_t rand_pitch, uint8_t* __restrict__ chal, size_t chal_pitch,
                                                             uint8_t* __restrict__ u, const bool vec) {
  const size_t groups = (n + 3) / 4;
  for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += (size_t)gridDim.x * blockDim.x) {
    const size_t i0 = 4 * g;
    uint32_t acc[27];
#pragma unroll
    for (int k = 0; k < 27; k++) acc[k] = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      if (i0 + j < n) {
        const uint4 q = reinterpret_cast<const uint4*>(in)[i0 + j];
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
        uint8_t v[27];
        unpack_witness_item(w, v);
#pragma unroll
        for (int k = 0; k < 27; k++) acc[k] |= (uint32_t)v[k] << (8 * j);
      }
    }
    if (wit) {
#pragma unroll
      for (int k = 0; k < 12; k++) store_plane_group(wit + (size_t)k * wit_pitch, i0, n, acc[k], vec);
    }
    if (rnd) {
#pragma unroll
      for (int k = 0; k < 9; k++) store_plane_group(rnd + (size_t)k * rand_pitch, i0, n, acc[12 + k], vec);
    }
    if (chal) {
#pragma unroll
      for (int k = 0; k < 5; k++) store_plane_group(chal + (size_t)k * chal_pitch, i0, n, acc[21 + k], vec);
    }
    if (u) store_plane_group(u, i0, n, acc[26], vec);
  }
}

// proof planes + status -> 12-byte packed proofs: four items = 48 contiguous bytes = three 128-bit stores (`vec` also requires
// `out` to be 16-byte aligned)
__global__ void __launch_bounds__(256) pack_proof_kernel(size_t n, const uint8_t* __restrict__ proof, size_t proof_pitch,
                                                         const uint8_t* __restrict__ status, pbh_packed_proof* __restrict__ out, const bool vec) {
  __shared__ PackedTables sT;
  stage_packed_tables(&sT);
  uint32_t* o = reinterpret_cast<uint32_t*>(out);
  const size_t groups = (n + 3) / 4;
  for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += (size_t)gridDim.x * blockDim.x) {
    const size_t i0 = 4 * g;
    uint32_t pl[27];
#pragma unroll
    for (int k = 0; k < 27; k++) pl[k] = load_plane_group(proof + (size_t)k * proof_pitch, i0, n, vec);
    const uint32_t st = status ? load_plane_group(status, i0, n, vec) : 0u;
    uint32_t w[12];
#pragma unroll
    for (int j = 0; j < 4; j++) {
      uint8_t p[27];
#pragma unroll
      for (int k = 0; k < 27; k++) p[k] = (uint8_t)(pl[k] >> (8 * j));
      pack_proof_item(sT, p, (uint8_t)(st >> (8 * j)), &w[3 * j]);
    }
    if (vec && i0 + 4 <= n) {
      uint4* dst = reinterpret_cast<uint4*>(o + 3 * i0);
      dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
      dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
      dst[2] = make_uint4(w[8], w[9], w[10], w[11]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; j++)
        if (i0 + j < n) { o[3 * (i0 + j)] = w[3 * j]; o[3 * (i0 + j) + 1] = w[3 * j + 1]; o[3 * (i0 + j) + 2] = w[3 * j + 2]; }
    }
  }
}

// packed proofs (+ the challenges-and-u words) -> the verifier's planes.  Null pointers are skipped.
__global__ void __launch_bounds__(256) unpack_proof_kernel(size_t n, const pbh_packed_proof* __restrict__ in, const uint32_t* __restrict__ chal_u,
                                                           uint8_t* __restrict__ proof, size_t proof_pitch, uint8_t* __restrict__ status,
                                                           uint8_t* __restrict__ chal, size_t chal_pitch, uint8_t* __restrict__ u, const bool vec) {
  __shared__ PackedTables sT;
  stage_packed_tables(&sT);
  const uint32_t* src = reinterpret_cast<const uint32_t*>(in);
  const size_t groups = (n + 3) / 4;
  for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += (size_t)gridDim.x * blockDim.x) {
    const size_t i0 = 4 * g;
    if (in && proof) {
      uint32_t w[12];
      if (vec && i0 + 4 <= n) {
        const uint4* s4 = reinterpret_cast<const uint4*>(src + 3 * i0);
        const uint4 a = s4[0], b = s4[1], c = s4[2];
        w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w; w[8] = c.x; w[9] = c.y; w[10] = c.z; w[11] = c.w;
      } else {
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const bool in_range = i0 + j < n;
          w[3 * j] = in_range ? src[3 * (i0 + j)] : 0u;
          w[3 * j + 1] = in_range ? src[3 * (i0 + j) + 1] : 0u;
          w[3 * j + 2] = in_range ? src[3 * (i0 + j) + 2] : 0u;
        }
      }
      uint32_t acc[27], st_acc = 0;
#pragma unroll
      for (int k = 0; k < 27; k++) acc[k] = 0;
#pragma unroll
      for (int j = 0; j < 4; j++) {
        uint8_t p[27], st;
        unpack_proof_item(sT, &w[3 * j], p, &st);
#pragma unroll
        for (int k = 0; k < 27; k++) acc[k] |= (uint32_t)p[k] << (8 * j);
        st_acc |= (uint32_t)st << (8 * j);
      }
#pragma unroll
      for (int k = 0; k < 27; k++) store_plane_group(proof + (size_t)k * proof_pitch, i0, n, acc[k], vec);
      if (status) store_plane_group(status, i0, n, st_acc, vec);
    }
    if (chal_u) {
      uint32_t acc[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
      for (int j = 0; j < 4; j++) {
        if (i0 + j < n) {
          uint8_t v[6];
          unpack_chal_u(chal_u[i0 + j], v);
#pragma unroll
          for (int k = 0; k < 6; k++) acc[k] |= (uint32_t)v[k] << (8 * j);
        }
      }
      if (chal) {
#pragma unroll
        for (int k = 0; k < 5; k++) store_plane_group(chal + (size_t)k * chal_pitch, i0, n, acc[k], vec);
      }
      if (u) store_plane_group(u, i0, n, acc[5], vec);
    }
  }
}
#endif  // __CUDACC__

}  // namespace pbh
