// Counter-based Philox4x32-10 draws, keyed by a path-derived particle key.
// The addressing scheme (streams, counter layout) is specified in DESIGN.md section 4 and mirrored by the
// CPU oracle (oracle/philox.py) so that both consume identical uniforms.
#pragma once
#include <stdint.h>
#include <string.h>

namespace pb {

constexpr uint32_t PHILOX_M0 = 0xD2511F53u, PHILOX_M1 = 0xCD9E8D57u;
constexpr uint32_t PHILOX_W0 = 0x9E3779B9u, PHILOX_W1 = 0xBB67AE85u;

enum Stream : uint32_t {
  ST_SUBSTEP = 1, ST_FINAL = 2, ST_MCS = 3, ST_CHOICE = 4, ST_VEGAS = 5, ST_KIN = 6, ST_DECAY = 7,
  ST_DBIN = 8, ST_PE = 11, ST_C0 = 12, ST_KEY_ROOT = 0xA0, ST_KEY_CHILD = 0xA1
};
constexpr uint32_t MCS_FINAL_INDEX = 0xFFFFFFFFu;

struct U4 { uint32_t x, y, z, w; };

__host__ __device__ __forceinline__ uint32_t mulhi32(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
  return __umulhi(a, b);
#else
  return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}

#ifndef PB_PHILOX_ROUNDS
#define PB_PHILOX_ROUNDS 10          // timing experiments only: the oracle (oracle/philox.py) implements 10
#endif
template <int ROUNDS = 10>
__host__ __device__ __forceinline__ U4 philox4x32_10(U4 c, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < ROUNDS; ++r) {
    uint32_t hi0 = mulhi32(PHILOX_M0, c.x), lo0 = PHILOX_M0 * c.x;
    uint32_t hi1 = mulhi32(PHILOX_M1, c.z), lo1 = PHILOX_M1 * c.z;
    c = U4{hi1 ^ c.y ^ k0, lo1, hi0 ^ c.w ^ k1, lo0};
    k0 += PHILOX_W0;
    k1 += PHILOX_W1;
  }
  return c;
}

// 52-bit uniform in [0,1) from two words, built in the mantissa of a double in [1,2) (three integer instructions and one
// DADD; no 64-bit integer -> double conversion): mantissa = hi (32 bits) : lo >> 12 (20 bits).  The low 12 bits of `lo`
// are left over ("spare" bits): a call's two spare fields feed the draws that need only a few bits.
__host__ __device__ __forceinline__ double u52(uint32_t hi, uint32_t lo) {
#ifdef __CUDA_ARCH__
  return __hiloint2double((int)(0x3FF00000u | (hi >> 12)), (int)((hi << 20) | (lo >> 12))) - 1.0;
#else
  uint64_t bits = ((uint64_t)(0x3FF00000u | (hi >> 12)) << 32) | (uint64_t)((hi << 20) | (lo >> 12));
  double d;
  memcpy(&d, &bits, sizeof(d));
  return d - 1.0;
#endif
}

struct D2 { double a, b; };

// The same uniforms shifted by one: doubles in [1, 2), i.e. before the "- 1.0" of u52 / u48.  For a consumer that scales the
// uniform, fma(v, s, -s) with v = u + 1 is bit-identical to u * s (v - 1 is exact, so both round the same real number once)
// and saves the subtraction.
__device__ __forceinline__ double u52p1(uint32_t hi, uint32_t lo) {
  return __hiloint2double((int)(0x3FF00000u | (hi >> 12)), (int)((hi << 20) | (lo >> 12)));
}
__device__ __forceinline__ double u48p1(uint32_t s0, uint32_t s1) {
  return __hiloint2double((int)(0x3FF00000u | (s0 >> 4)), (int)(((s0 & 0xFu) << 28) | (s1 << 4)));
}

__host__ __device__ __forceinline__ D2 draw2(uint2 key, uint32_t c0, uint32_t stream, uint32_t c2 = 0, uint32_t c3 = 0) {
  U4 o = philox4x32_10<PB_PHILOX_ROUNDS>(U4{c0, stream, c2, c3}, key.x, key.y);
  return D2{u52(o.x, o.y), u52(o.z, o.w)};
}
// same two doubles plus the call's 24 spare bits ((o1 & 0xFFF) << 12 | (o3 & 0xFFF))
__host__ __device__ __forceinline__ D2 draw2s(uint2 key, uint32_t c0, uint32_t stream, uint32_t c2, uint32_t c3, uint32_t& spare) {
  U4 o = philox4x32_10<PB_PHILOX_ROUNDS>(U4{c0, stream, c2, c3}, key.x, key.y);
  spare = ((o.y & 0xFFFu) << 12) | (o.w & 0xFFFu);
  return D2{u52(o.x, o.y), u52(o.z, o.w)};
}
__device__ __forceinline__ D2 draw2s_p1(uint2 key, uint32_t c0, uint32_t stream, uint32_t c2, uint32_t c3, uint32_t& spare);
// 48-bit uniform in [0,1) from the spare bits of two calls (the accept/reject uniform of a 4-D trial)
__host__ __device__ __forceinline__ double u48(uint32_t s0, uint32_t s1) {
#ifdef __CUDA_ARCH__
  return __hiloint2double((int)(0x3FF00000u | (s0 >> 4)), (int)(((s0 & 0xFu) << 28) | (s1 << 4))) - 1.0;
#else
  uint64_t bits = ((uint64_t)(0x3FF00000u | (s0 >> 4)) << 32) | (uint64_t)(((s0 & 0xFu) << 28) | (s1 << 4));
  double d;
  memcpy(&d, &bits, sizeof(d));
  return d - 1.0;
#endif
}

__device__ __forceinline__ D2 draw2s_p1(uint2 key, uint32_t c0, uint32_t stream, uint32_t c2, uint32_t c3, uint32_t& spare) {
  U4 o = philox4x32_10<PB_PHILOX_ROUNDS>(U4{c0, stream, c2, c3}, key.x, key.y);
  spare = ((o.y & 0xFFFu) << 12) | (o.w & 0xFFFu);
  return D2{u52p1(o.x, o.y), u52p1(o.z, o.w)};
}

__host__ __device__ __forceinline__ uint2 root_key(uint64_t seed, uint64_t shower) {
  U4 o = philox4x32_10(U4{(uint32_t)shower, (uint32_t)(shower >> 32), 0u, (uint32_t)ST_KEY_ROOT},
                       (uint32_t)seed, (uint32_t)(seed >> 32));
  return make_uint2(o.x, o.y);
}

__host__ __device__ __forceinline__ uint2 child_key(uint2 key, uint32_t bit) {
  U4 o = philox4x32_10(U4{bit, (uint32_t)ST_KEY_CHILD, 0u, 0u}, key.x, key.y);
  return make_uint2(o.x, o.y);
}

}  // namespace pb
