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
