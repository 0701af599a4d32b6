/* twixt_b200.h -- C ABI of libtwixt_b200.so, the B200-native batched TwixT engine.
 *
 * This is the drop-in boundary for the hot path of
 * stevens68/TwixT_for_open_spiel: one batched entry point per method of the
 * open_spiel Game/State surface that the reference overrides in
 *   open_spiel/games/twixt/twixt.h:31-146  and  twixt.cc:35-145.
 * A binding for open_spiel (C++), pyspiel-style Python (ctypes) or any other
 * FFI only needs this header; no CUDA, torch or C++ types appear in it.
 *
 * Conventions
 *  - Every call returns TWIXT_OK (0) or a negative TWIXT_E* code and never
 *    aborts the process (the reference calls SpielFatalError, twixt.h:96,
 *    twixt.cc:140-143).  twixt_last_error() returns the message of the last
 *    failing call on the calling thread; messages that the reference pins are
 *    reproduced byte for byte ("board_size out of range [5..24]: 30",
 *    "Not a legal action: 11").
 *  - A batch holds `num_envs` independent games ("envs") of one board size in
 *    device memory.  Batched calls address the contiguous env range
 *    [first, first+count).
 *  - Every data pointer may be a DEVICE pointer (cudaMalloc / torch CUDA
 *    tensor: the kernel reads/writes it directly, asynchronously on the
 *    batch's stream) or a HOST pointer (pageable or pinned: the library stages
 *    it -- small transfers through a pinned, device-mapped arena the kernel
 *    accesses directly, large ones through device scratch -- and the call
 *    returns after the data arrived).  The kind is detected with
 *    cudaPointerGetAttributes.
 *  - A batch is not thread-safe; distinct batches are independent.
 *  - There is no CPU fallback: every entry point that touches game state
 *    launches sm_100a kernels and fails with TWIXT_ECUDA when no such device
 *    is present.
 *
 * State record (what twixt_export_state / twixt_import_state move, and what
 * lives in HBM): per env `record_words` 32-bit words, 16-byte aligned:
 *    word 0  ply            (Board::move_counter_, twixtboard.h:75)
 *    word 1  bits 0-1 result (0 open 1 red won 2 blue won 3 draw, twixtboard.h:48)
 *            bit 2 swapped   (twixtboard.h:76)
 *    word 2  action of the first move (Board::move_one_), 0xFFFFFFFF before it
 *    word 3  bits 0-15 / 16-31: number of empty cells red / blue may play on
 *    then 9 bit-planes of n column words each; bit y of word x <=> cell (x,y),
 *    x = column, y = row counted from the bottom (twixtboard.h:153-213):
 *      0 red pegs                 1 blue pegs
 *      2..5 links NNE, ENE, ESE, SSE stored at their WEST endpoint
 *            (Cell::links_ bits 0..3 there, bits 4..7 at the other end,
 *             twixtcell.h:58-78)
 *      6 / 7 peg is linked to its owner's start / end border line
 *            (Cell::linked_to_border_, twixtcell.h:89-95,107)
 *      8 peg has a blocked neighbour in an east direction
 *            (Cell::HasBlockedNeighborsEast, twixtcell.h:82-84)
 *    zero padding up to a multiple of 4 words.
 */
#ifndef TWIXT_B200_H_
#define TWIXT_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TWIXT_MIN_BOARD_SIZE 5      /* twixtboard.h:32 */
#define TWIXT_MAX_BOARD_SIZE 24     /* twixtboard.h:33 */
#define TWIXT_DEFAULT_BOARD_SIZE 8  /* twixtboard.h:34 */
#define TWIXT_NUM_OBS_PLANES 12     /* twixtboard.h:46 */
#define TWIXT_TERMINAL_PLAYER (-4)  /* open_spiel kTerminalPlayerId, playthrough.txt:678 */
#define TWIXT_HEADER_WORDS 4
#define TWIXT_NUM_STATE_PLANES 9

enum {
  TWIXT_OK = 0,
  TWIXT_EINVAL = -1,   /* bad argument (range, null pointer, size) */
  TWIXT_ECUDA = -2,    /* CUDA runtime error / no sm_100 device */
  TWIXT_EILLEGAL = -3, /* twixt_apply: at least one action was not legal */
  TWIXT_ENOMEM = -4
};

typedef struct twixt_batch twixt_batch;

/* Game-level constants.  Mirrors TwixTGame's accessors, twixt.h:122-141. */
typedef struct twixt_game_info {
  int32_t board_size;
  int32_t num_distinct_actions; /* n*n            twixt.h:122-124 */
  int32_t num_players;          /* 2              twixt.h:126 */
  int32_t max_game_length;      /* n*n-4+1        twixt.h:136-139 */
  int32_t obs_shape[3];         /* {12, n, n-2}   twixt.h:131-134 (without its static-shape bug) */
  int32_t obs_size;             /* 12*n*(n-2) */
  int32_t max_legal_actions;    /* n*(n-2)        twixtboard.cc:253 */
  int32_t record_words;         /* words per env state record */
  double min_utility;           /* -1  twixt.h:127 */
  double max_utility;           /* +1  twixt.h:129 */
  double utility_sum;           /*  0  twixt.h:128 */
} twixt_game_info;

/* Aggregate counters of the playouts run on a batch since the last
 * twixt_stats_reset (the only quantity ever reduced across GPUs). */
typedef struct twixt_stats {
  int64_t plies;      /* moves applied inside twixt_playout */
  int64_t games;      /* envs that reached a terminal state inside twixt_playout */
  int64_t red_wins;
  int64_t blue_wins;
  int64_t draws;
  int64_t swaps;      /* playouts in which the swap rule was used */
  int64_t max_length; /* longest finished game (plies) */
  int64_t kernel_launches; /* kernels launched by this batch since creation */
  int64_t debug_violations; /* always 0 in the product build; the bounds-instrumented test variant of the
                               playout kernel counts accesses outside an env's own storage here */
} twixt_stats;

const char* twixt_last_error(void);
const char* twixt_version(void);

/* TwixTGame::TwixTGame range check, twixt.cc:134-145.  No GPU needed. */
int twixt_game_info_for(int board_size, twixt_game_info* out);

/* Replaces TwixTGame construction + NewInitialState() (twixt.h:118-120,
 * twixt.cc:62-65) for num_envs games at once: allocates the records on
 * `device` and resets every env to the initial position.  `seed` keys the
 * Philox streams of twixt_playout. */
int twixt_create(int board_size, int64_t num_envs, int device, uint64_t seed, twixt_batch** out);
void twixt_destroy(twixt_batch* b);
int twixt_get_info(const twixt_batch* b, twixt_game_info* out);
int64_t twixt_num_envs(const twixt_batch* b);

/* Use an existing CUDA stream (e.g. torch's current stream) instead of the
 * batch's own; pass the cudaStream_t as an integer. */
int twixt_set_stream(twixt_batch* b, uintptr_t cuda_stream);
uintptr_t twixt_get_stream(const twixt_batch* b);
int twixt_synchronize(twixt_batch* b);
int twixt_set_seed(twixt_batch* b, uint64_t seed);
/* Philox stream id of env 0 of this batch when twixt_playout gets no
 * stream_ids (env i uses base + i): lets shards on several GPUs draw from
 * disjoint streams so results do not depend on the GPU count. */
int twixt_set_stream_base(twixt_batch* b, uint64_t base);

/* NewInitialState(), twixt.h:118-120 / Board::Board twixtboard.cc:168-174. */
int twixt_reset(twixt_batch* b, int64_t first, int64_t count);

/* State::Clone(), twixt.h:80-82.  dst env i <- src env i (ranges may not
 * overlap), or dst env first+i <- env src_ids[i] for the gather form
 * (src_ids: int64, host or device).  Every id must lie in [0, num_envs) and outside the destination range;
 * an offending id copies nothing for its env and fails the call with TWIXT_EINVAL (ids on the device are
 * checked by the kernel itself).  twixt_clone_from rejects overlapping ranges within one batch. */
int twixt_clone(twixt_batch* b, int64_t src_first, int64_t dst_first, int64_t count);
int twixt_clone_gather(twixt_batch* b, const int64_t* src_ids, int64_t dst_first, int64_t count);
/* Clone between two batches of the same board size (possibly on one device). */
int twixt_clone_from(twixt_batch* dst, int64_t dst_first, const twixt_batch* src, int64_t src_first,
                     int64_t count);

/* TwixTState::LegalActions(), twixt.h:86-90: ascending actions of the player
 * to move, empty when terminal.  out_actions is [count, stride] with
 * stride >= max_legal_actions elements of elem_bytes (2: uint16, 4: int32,
 * 8: int64 = open_spiel::Action); entries past the count are left untouched.
 * out_counts is [count] int32.  Either pointer may be null. */
int twixt_legal_actions(twixt_batch* b, int64_t first, int64_t count, void* out_actions,
                        int32_t elem_bytes, int64_t stride, int32_t* out_counts);

/* Upstream State::LegalActionsMask(): [count, n*n] uint8, 1 = legal. */
int twixt_legal_mask(twixt_batch* b, int64_t first, int64_t count, uint8_t* out_mask);

/* State::ApplyAction -> TwixTState::DoApplyAction, twixt.h:93-104 ->
 * Board::ApplyAction, twixtboard.cc:457-499.  actions is [count] int32, one
 * per env; a negative action skips that env.  out_status (nullable) is
 * [count] int32: 0 applied, 1 "Not a legal action" (env left unchanged),
 * 2 skipped.  Returns TWIXT_EILLEGAL if any env reported 1 (only checked when
 * out_status is a host pointer or null); the message names the first such
 * action like the reference does. */
int twixt_apply(twixt_batch* b, int64_t first, int64_t count, const int32_t* actions,
                int32_t* out_status);

/* CurrentPlayer() twixt.h:38 (0 red, 1 blue, -4 terminal), IsTerminal()
 * twixt.h:45-48, Returns() twixt.h:50-63 ([count,2] float32: +-1 or 0). */
int twixt_current_player(twixt_batch* b, int64_t first, int64_t count, int8_t* out);
int twixt_is_terminal(twixt_batch* b, int64_t first, int64_t count, uint8_t* out);
int twixt_returns(twixt_batch* b, int64_t first, int64_t count, float* out);

/* TwixTState::ObservationTensor, twixt.cc:76-132: [count, 12, n, n-2]
 * float32, zero-filled then 1.0s; identical for both players. */
int twixt_observation(twixt_batch* b, int64_t first, int64_t count, float* out);

/* ObservationTensor + LegalActionsMask of the same envs in ONE pass over the records (the producer of an
 * AlphaZero-style inference batch, BASELINE config C5): out_obs as twixt_observation, out_mask as
 * twixt_legal_mask; bit-identical to the two separate calls. */
int twixt_observation_and_mask(twixt_batch* b, int64_t first, int64_t count, float* out_obs,
                               uint8_t* out_mask);

/* Replays a whole action history per env in one launch: upstream serialises a state as its action history
 * (State::Serialize) and Game::DeserializeState re-applies it move by move; the reference has no UndoAction
 * (twixt.h:84), so replay is also how a search returns to an earlier position.  actions is [count, stride]
 * int32; env first+i applies actions[i*stride + k] for k = 0 .. len_i-1 from its CURRENT state, with the
 * legality test of DoApplyAction (twixt.h:93-104) before every move, where len_i = lengths[i] (nullable:
 * then a row ends at its first negative entry or at stride).  An env stops at its first illegal action,
 * keeping the state reached; out_applied (nullable) is [count] int32 = moves made.  Returns TWIXT_EILLEGAL
 * with the reference's message for the lowest such env (only checked when out_applied is null or any
 * pointer is a host pointer; with device pointers throughout the call is asynchronous). */
int twixt_replay(twixt_batch* b, int64_t first, int64_t count, const int32_t* actions, int64_t stride,
                 const int32_t* lengths, int32_t* out_applied);

/* One State step for an UNBATCHED caller -- what upstream example.cc / mcts_example.cc do per move:
 * ApplyAction, then IsTerminal / CurrentPlayer / Returns / LegalActions of the new state (twixt.h:38-104).
 * One kernel launch does all of it for env `env`: `action` >= 0 is applied with the legality test of
 * DoApplyAction (status 1 and TWIXT_EILLEGAL "Not a legal action: N" if illegal, env unchanged),
 * TWIXT_STEP_QUERY applies nothing, TWIXT_STEP_RESET resets the env to the initial position first; then
 * `out` and the ascending `out_legal` ([max_legal_actions] int64 = open_spiel::Action, nullable) describe the
 * resulting state.  An adapter keeps the answer with its State object, so the four query methods cost no
 * further GPU call until the next ApplyAction (twixt_for_open_spiel_b200/adapter). */
#define TWIXT_STEP_QUERY (-1)
#define TWIXT_STEP_RESET (-2)
typedef struct twixt_step_result {
  int32_t status;         /* 0 applied / nothing to apply, 1 illegal action */
  int32_t current_player; /* 0, 1 or TWIXT_TERMINAL_PLAYER */
  int32_t is_terminal;
  int32_t num_legal;      /* entries written to out_legal (0 when terminal) */
  float returns[2];
} twixt_step_result;
int twixt_step(twixt_batch* b, int64_t env, int32_t action, twixt_step_result* out, int64_t* out_legal);

/* The random-playout loop of upstream example.cc / RandomRolloutEvaluator
 * (LegalActions -> uniform pick -> ApplyAction until IsTerminal), fused into
 * one kernel with the env state held on chip.  Every env in the range is
 * played from its CURRENT state until it is terminal or max_plies moves were
 * made, and is left in the state reached.  Move `step` (0-based within this
 * call) of an env picks index (word*L)>>32 of its ascending legal list of
 * length L, where word = Philox4x32-10(key=seed, counter=(stream, step>>2,0))
 * [step&3] and stream = stream_ids[i] if given, else the env index.
 * Outputs (all nullable): out_returns [count,2] float32, out_lengths [count]
 * int32 = moves made by THIS call, out_actions [max_plies_traced, count]
 * uint16 step-major action trace (0xFFFF where no move was made) with
 * trace_plies rows. */
int twixt_playout(twixt_batch* b, int64_t first, int64_t count, int32_t max_plies,
                  const uint64_t* stream_ids, float* out_returns, int32_t* out_lengths,
                  uint16_t* out_actions, int32_t trace_plies);

/* Raw state records, [count, record_words] uint32 (host or device).
 * The reference can only reach a state through DoApplyAction (twixt.h:93-104); records handed to
 * twixt_import_state are therefore VALIDATED on the device before any env is overwritten (header ranges,
 * peg placement and counts, link endpoints, flag/blocked bits on pegs only, recounted legal-cell counters,
 * zero padding).  A bad record fails the whole call with TWIXT_EINVAL, "invalid state record at index I:
 * <reason>", and leaves the batch untouched.  twixt_set_validation(b, 0) skips the check for trusted
 * records (a plain copy, asynchronous for device pointers); the default is on. */
int twixt_export_state(twixt_batch* b, int64_t first, int64_t count, uint32_t* out_records);
int twixt_import_state(twixt_batch* b, int64_t first, int64_t count, const uint32_t* records);
int twixt_set_validation(twixt_batch* b, int enabled);

int twixt_get_stats(twixt_batch* b, twixt_stats* out);
int twixt_stats_reset(twixt_batch* b);

/* Multi-GPU plumbing for hosts that are not Python (envs shard by contiguous global id ranges,
 * no data-path collective, one reduction of the counters).  twixt_shard_range: rank's share [first,
 * first+count) of global_envs envs over `world` ranks (shares differ by at most one env); create the rank's
 * batch with `count` envs and call twixt_set_stream_base(b, first) so games do not depend on the GPU
 * count.  twixt_stats_accumulate: acc += part (sums; max for max_length) -- apply it to the ranks'
 * twixt_get_stats results, or to the output of an MPI/NCCL all-gather.  Neither needs a GPU. */
int twixt_shard_range(int64_t global_envs, int32_t world, int32_t rank, int64_t* out_first,
                      int64_t* out_count);
int twixt_stats_accumulate(twixt_stats* acc, const twixt_stats* part);

#ifdef __cplusplus
}
#endif
#endif /* TWIXT_B200_H_ */
