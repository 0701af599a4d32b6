/* TEST INFRASTRUCTURE ONLY (oracle/).
 *
 * Plain-C CPU restatement of the reference TwixT engine
 * (/root/reference/open_spiel/games/twixt/: twixtboard.cc, twixtcell.h,
 * twixt.h, twixt.cc).  It is the checker the CUDA path is compared with; it
 * is never linked into, imported by or executed from the product path
 * (twixt_for_open_spiel_b200/).  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / reference arm may use it.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_reference.py checks this
 * restatement against (a) every known-answer test in the reference's
 * twixt_test.cc and the golden playthrough.txt, and (b) the unmodified
 * reference compiled into oracle/_ref/libtwixt_ref.so, move by move, on
 * seeded random games at every board size 5..24 (legal lists, player,
 * terminal flag, returns, observation tensor and all cell internals).
 */
#ifndef ORACLE_TWIXT_ORACLE_H_
#define ORACLE_TWIXT_ORACLE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORACLE_MIN_BOARD 5   /* twixtboard.h:32 */
#define ORACLE_MAX_BOARD 24  /* twixtboard.h:33 */
#define ORACLE_NUM_PLANES 12 /* twixtboard.h:46 */
#define ORACLE_TERMINAL_PLAYER (-4)

/* Packed-state record shared with the CUDA engine (include/twixt_b200.h,
 * "State record"): 4 header words followed by 9 bit-planes of n column
 * words each (bit y of word x <=> cell (x,y)). */
#define ORACLE_HEADER_WORDS 4
#define ORACLE_NUM_STATE_PLANES 9

typedef struct oracle_game oracle_game;
typedef struct oracle_state oracle_state;

oracle_game* oracle_game_new(int board_size, char* err, int errcap);
void oracle_game_free(oracle_game* g);
int oracle_game_board_size(const oracle_game* g);
int oracle_num_distinct_actions(const oracle_game* g);
int oracle_max_game_length(const oracle_game* g);
int oracle_observation_size(const oracle_game* g);
/* number of crossing links stored for the directed link (x,y,dir); the links
 * themselves go to out (triples x,y,dir) when out != NULL */
int oracle_blockers(const oracle_game* g, int x, int y, int dir, int* out);

oracle_state* oracle_state_new(const oracle_game* g);
oracle_state* oracle_state_clone(const oracle_state* s);
void oracle_state_copy(oracle_state* dst, const oracle_state* src);
void oracle_state_free(oracle_state* s);

int oracle_legal_actions(const oracle_game* g, const oracle_state* s, int64_t* out);
int oracle_legal_list_of(const oracle_game* g, const oracle_state* s, int player, int64_t* out);
int oracle_apply(const oracle_game* g, oracle_state* s, int64_t action, char* err, int errcap);
int oracle_current_player(const oracle_state* s);
int oracle_is_terminal(const oracle_state* s);
void oracle_returns(const oracle_state* s, double* out2);
void oracle_observation(const oracle_game* g, const oracle_state* s, float* out);
void oracle_board_header(const oracle_state* s, int* out5);
void oracle_export_cells(const oracle_game* g, const oracle_state* s, int* out);
int oracle_record_words(const oracle_game* g);
void oracle_export_record(const oracle_game* g, const oracle_state* s, uint32_t* out);

int oracle_replay(const oracle_game* g, oracle_state* s, const int64_t* actions, int len);
int oracle_playout_philox(const oracle_game* g, oracle_state* s, uint64_t seed, uint64_t stream,
                          int max_plies, int64_t* actions_out);
/* Time-bounded mt-free random playouts for bench.py's cpu_baseline ("port"
 * kind): plays Philox-policy games from the initial state for about
 * `seconds`; returns plies, writes games to *games_out. */
int64_t oracle_bench_playouts(const oracle_game* g, double seconds, uint64_t seed,
                              int64_t* games_out, double* elapsed_out);

void oracle_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);

#ifdef __cplusplus
}
#endif
#endif /* ORACLE_TWIXT_ORACLE_H_ */
