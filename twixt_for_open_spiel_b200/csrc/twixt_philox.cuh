// twixt_philox.cuh -- Philox4x32-10 (Salmon, Moraes, Dror, Shaw, SC'11) and the
// playout policy's random-word -> legal-list-index mapping.
//
// Stream definition (include/twixt_b200.h, twixt_playout): key = 64-bit batch
// seed, counter = (stream id lo, hi, step >> 2, 0); one 128-bit block serves
// four consecutive moves of one env.
#pragma once
#include <stdint.h>

#include "twixt_engine.cuh"

namespace twixt {

TW_HD void mulhilo32(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
#if defined(__CUDA_ARCH__)
  lo = a * b;
  hi = __umulhi(a, b);
#else
  uint64_t p = static_cast<uint64_t>(a) * b;
  lo = static_cast<uint32_t>(p);
  hi = static_cast<uint32_t>(p >> 32);
#endif
}

TW_HD void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                         uint32_t out[4]) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int round = 0; round < 10; ++round) {
    uint32_t hi0, lo0, hi1, lo1;
    mulhilo32(0xD2511F53u, c0, hi0, lo0);
    mulhilo32(0xCD9E8D57u, c2, hi1, lo1);
    const uint32_t n0 = hi1 ^ c1 ^ k0;
    const uint32_t n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// index into a list of length L >= 1
TW_HD uint32_t playout_index(uint32_t word, uint32_t L) {
#if defined(__CUDA_ARCH__)
  return __umulhi(word, L);
#else
  return static_cast<uint32_t>((static_cast<uint64_t>(word) * L) >> 32);
#endif
}

}  // namespace twixt
