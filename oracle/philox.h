/* TEST INFRASTRUCTURE ONLY (oracle/).
 * Philox4x32-10 counter-based generator (Salmon et al., "Parallel random
 * numbers: as easy as 1, 2, 3", SC'11) and the playout policy that maps a
 * random word to an index into the ascending legal-action list.  Restated on
 * the CPU so the oracle can regenerate exactly the games the fused CUDA
 * playout kernel plays.  The reference itself contains no RNG: upstream
 * example.cc draws uniformly from LegalActions(); the
 * stream definition below is this repo's own (DESIGN.md "RNG"):
 *
 *   key     = (seed_lo, seed_hi)
 *   counter = (stream_lo, stream_hi, step >> 2, 0)      step = 0,1,2,... of
 *   word    = out[step & 3]                              this playout call
 *   index   = (word * L) >> 32                           L = |LegalActions()|
 */
#ifndef ORACLE_PHILOX_H_
#define ORACLE_PHILOX_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

static inline void oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2],
                                        uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
  uint32_t k0 = key[0], k1 = key[1];
  for (int round = 0; round < 10; ++round) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* Random word for (seed, stream, step). */
static inline uint32_t oracle_playout_word(uint64_t seed, uint64_t stream, uint32_t step) {
  uint32_t ctr[4] = {(uint32_t)stream, (uint32_t)(stream >> 32), step >> 2, 0u};
  uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  uint32_t out[4];
  oracle_philox4x32_10(ctr, key, out);
  return out[step & 3u];
}

/* Index into a list of length L (L >= 1). */
static inline uint32_t oracle_playout_index(uint32_t word, uint32_t L) {
  return (uint32_t)(((uint64_t)word * (uint64_t)L) >> 32);
}

#ifdef __cplusplus
}
#endif
#endif /* ORACLE_PHILOX_H_ */
