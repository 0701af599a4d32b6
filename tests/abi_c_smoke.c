/* TEST-ONLY: include/twixt_b200.h must be usable from plain C (it is the FFI boundary).  Compiled with
 * gcc -std=c99 -pedantic and linked against libtwixt_b200.so; only GPU-free entry points are called here,
 * plus twixt_create to see that it reports the missing GPU instead of computing on the host. */
#include <stdio.h>
#include <string.h>

#include "twixt_b200.h"

int main(void) {
  twixt_game_info info;
  twixt_batch* b = NULL;
  int rc;
  if (twixt_game_info_for(24, &info) != TWIXT_OK) return 1;
  if (info.num_distinct_actions != 576 || info.obs_size != 12 * 24 * 22 || info.record_words != 220) return 2;
  if (info.max_game_length != 573 || info.max_legal_actions != 528) return 3;
  if (twixt_game_info_for(30, &info) != TWIXT_EINVAL) return 4;
  if (strcmp(twixt_last_error(), "board_size out of range [5..24]: 30") != 0) return 5;
  if (twixt_reset(NULL, 0, 1) != TWIXT_EINVAL) return 6;
  {
    /* multi-GPU plumbing needs no GPU: shares of 10 envs over 4 ranks are 3,3,2,2 and contiguous */
    int64_t first = -1, count = -1, next = 0;
    int r;
    twixt_stats total, part;
    for (r = 0; r < 4; ++r) {
      if (twixt_shard_range(10, 4, r, &first, &count) != TWIXT_OK || first != next || count != (r < 2 ? 3 : 2)) return 9;
      next = first + count;
    }
    if (next != 10 || twixt_shard_range(10, 4, 4, &first, &count) != TWIXT_EINVAL) return 10;
    memset(&total, 0, sizeof(total));
    memset(&part, 0, sizeof(part));
    part.plies = 7; part.games = 2; part.max_length = 5;
    twixt_stats_accumulate(&total, &part);
    part.max_length = 3;
    twixt_stats_accumulate(&total, &part);
    if (total.plies != 14 || total.games != 4 || total.max_length != 5) return 11;
  }
  rc = twixt_create(8, 16, 0, 1u, &b);
  if (rc == TWIXT_OK) {
    /* a GPU is present: exercise one call and clean up */
    int8_t player[16];
    if (twixt_current_player(b, 0, 16, player) != TWIXT_OK || player[0] != 0) return 7;
    twixt_destroy(b);
    printf("OK gpu\n");
  } else {
    if (rc != TWIXT_ECUDA || b != NULL || strstr(twixt_last_error(), "no CPU fallback") == NULL) return 8;
    printf("OK nogpu\n");
  }
  return 0;
}
