/* TEST INFRASTRUCTURE ONLY (oracle/) -- see twixt_oracle.h for the rules on
 * who may use this file and for its parity status (PINNED against the
 * reference's own tests and against the compiled reference).
 *
 * A cell-array restatement of the reference engine.  Each function names the
 * reference lines it follows.  The one deliberate difference in *method*: the
 * reference spells the link-crossing table out as data
 * (twixtboard.cc:38-144) and expands it into a process-global map per board
 * (twixtboard.cc:176-190); here the same relation is derived from geometry
 * (two knight-move segments block each other iff they properly intersect),
 * which tests/test_oracle_vs_reference.py shows to be identical link by link.
 */
#define _POSIX_C_SOURCE 200809L
#include "twixt_oracle.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "philox.h"

enum { RED = 0, BLUE = 1, EMPTY = 2, OFFBOARD = 3 };     /* twixtboard.h:50 */
enum { OPEN = 0, RED_WIN = 1, BLUE_WIN = 2, DRAW = 3 };  /* twixtboard.h:48 */
enum { START = 0, END = 1 };                             /* twixtcell.h:50 */
#define NB ORACLE_MAX_BOARD
#define NDIR 8

/* Compass offsets, twixtcell.h:58-68 / twixtboard.cc:40,54,67,80,93,106,119,132 */
static const int kDx[NDIR] = {1, 2, 2, 1, -1, -2, -2, -1};
static const int kDy[NDIR] = {2, 1, -1, -2, -2, -1, 1, 2};

typedef struct { int8_t x, y, d; } link_ref;

struct oracle_game {
  int n;
  /* for the directed link (x,y,d): crossing links, each named once by its
   * west endpoint and east direction */
  uint8_t nblock[NB][NB][NDIR];
  link_ref block[NB][NB][NDIR][9];
};

struct oracle_state {
  int n;
  int move_counter;   /* twixtboard.h:75 */
  int swapped;        /* :76 */
  int move_one_x;     /* :77 */
  int move_one_y;
  int result;         /* :78 */
  int current_player; /* twixt.h:107 */
  uint8_t color[NB][NB];
  uint8_t links[NB][NB];    /* twixtcell.h:100 */
  uint8_t blocked[NB][NB];  /* twixtcell.h:102 */
  uint8_t border[NB][NB];   /* twixtcell.h:107: bit (2*player + border) */
  uint8_t in_list[2][NB * NB]; /* membership form of legal_actions_[p], twixtboard.h:82 */
  int list_count[2];
};

static int opp_dir(int d) { return (d + NDIR / 2) % NDIR; } /* twixtboard.cc:28-30 */

/* twixtboard.cc:625-631 */
static int off_board(int n, int x, int y) {
  return y < 0 || y > n - 1 || x < 0 || x > n - 1 ||
         ((x == 0 || x == n - 1) && (y == 0 || y == n - 1));
}

/* twixtboard.cc:615-623 */
static int on_border(int n, int player, int x, int y) {
  if (player == RED) return (y == 0 || y == n - 1) && (x > 0 && x < n - 1);
  return (x == 0 || x == n - 1) && (y > 0 && y < n - 1);
}

static long cross(long ax, long ay, long bx, long by) { return ax * by - ay * bx; }

/* proper intersection of segments p1-p2 and q1-q2 (no shared endpoint; a
 * knight-move segment contains no lattice point besides its ends) */
static int segments_cross(int p1x, int p1y, int p2x, int p2y, int q1x, int q1y, int q2x, int q2y) {
  long d1 = cross(q2x - q1x, q2y - q1y, p1x - q1x, p1y - q1y);
  long d2 = cross(q2x - q1x, q2y - q1y, p2x - q1x, p2y - q1y);
  long d3 = cross(p2x - p1x, p2y - p1y, q1x - p1x, q1y - p1y);
  long d4 = cross(p2x - p1x, p2y - p1y, q2x - p1x, q2y - p1y);
  return ((d1 > 0 && d2 < 0) || (d1 < 0 && d2 > 0)) && ((d3 > 0 && d4 < 0) || (d3 < 0 && d4 > 0));
}

/* The relation built by InitializeBlockerMap (twixtboard.cc:176-190) from
 * kLinkDescriptorTable (38-144): only links with both ends on the board. */
static void build_blockers(oracle_game* g) {
  int n = g->n;
  memset(g->nblock, 0, sizeof(g->nblock));
  for (int x = 0; x < n; ++x)
    for (int y = 0; y < n; ++y) {
      if (off_board(n, x, y)) continue;
      for (int d = 0; d < NDIR; ++d) {
        int tx = x + kDx[d], ty = y + kDy[d];
        if (off_board(n, tx, ty)) continue;
        int cnt = 0;
        for (int wx = x - 3; wx <= x + 3; ++wx)
          for (int wy = y - 3; wy <= y + 3; ++wy) {
            if (off_board(n, wx, wy)) continue;
            for (int e = 0; e < 4; ++e) {
              int ex = wx + kDx[e], ey = wy + kDy[e];
              if (off_board(n, ex, ey)) continue;
              if ((wx == x && wy == y) || (wx == tx && wy == ty) || (ex == x && ey == y) ||
                  (ex == tx && ey == ty))
                continue;
              if (!segments_cross(x, y, tx, ty, wx, wy, ex, ey)) continue;
              if (cnt < 9) {
                g->block[x][y][d][cnt].x = (int8_t)wx;
                g->block[x][y][d][cnt].y = (int8_t)wy;
                g->block[x][y][d][cnt].d = (int8_t)e;
              }
              ++cnt;
            }
          }
        if (cnt > 9) { fprintf(stderr, "oracle: >9 crossers\n"); abort(); }
        g->nblock[x][y][d] = (uint8_t)cnt;
      }
    }
}

/* twixt.cc:134-145 */
oracle_game* oracle_game_new(int board_size, char* err, int errcap) {
  if (board_size < ORACLE_MIN_BOARD || board_size > ORACLE_MAX_BOARD) {
    if (err && errcap > 0)
      snprintf(err, (size_t)errcap, "board_size out of range [%d..%d]: %d", ORACLE_MIN_BOARD,
               ORACLE_MAX_BOARD, board_size);
    return NULL;
  }
  oracle_game* g = (oracle_game*)calloc(1, sizeof(oracle_game));
  g->n = board_size;
  build_blockers(g);
  return g;
}

void oracle_game_free(oracle_game* g) { free(g); }
int oracle_game_board_size(const oracle_game* g) { return g->n; }
int oracle_num_distinct_actions(const oracle_game* g) { return g->n * g->n; }          /* twixt.h:122-124 */
int oracle_max_game_length(const oracle_game* g) { return g->n * g->n - 4 + 1; }        /* twixt.h:136-139 */
int oracle_observation_size(const oracle_game* g) { return ORACLE_NUM_PLANES * g->n * (g->n - 2); }

int oracle_blockers(const oracle_game* g, int x, int y, int dir, int* out) {
  int c = g->nblock[x][y][dir];
  if (out)
    for (int i = 0; i < c; ++i) {
      out[3 * i + 0] = g->block[x][y][dir][i].x;
      out[3 * i + 1] = g->block[x][y][dir][i].y;
      out[3 * i + 2] = g->block[x][y][dir][i].d;
    }
  return c;
}

/* twixtboard.cc:252-276 */
static void init_legal_actions(oracle_state* s) {
  int n = s->n;
  memset(s->in_list, 0, sizeof(s->in_list));
  s->list_count[RED] = s->list_count[BLUE] = 0;
  for (int col = 0; col < n; ++col)
    for (int row = 0; row < n; ++row) {
      int a = col * n + row;
      if (off_board(n, col, row)) continue;
      if (on_border(n, RED, col, row)) {
        s->in_list[RED][a] = 1; s->list_count[RED]++;
      } else if (on_border(n, BLUE, col, row)) {
        s->in_list[BLUE][a] = 1; s->list_count[BLUE]++;
      } else {
        s->in_list[RED][a] = 1; s->list_count[RED]++;
        s->in_list[BLUE][a] = 1; s->list_count[BLUE]++;
      }
    }
}

/* twixtboard.cc:168-174, 209-236 */
static void state_init(const oracle_game* g, oracle_state* s) {
  int n = g->n;
  memset(s, 0, sizeof(*s));
  s->n = n;
  s->result = OPEN;
  s->current_player = RED;
  s->move_one_x = s->move_one_y = -1;
  for (int x = 0; x < n; ++x)
    for (int y = 0; y < n; ++y) {
      if (off_board(n, x, y)) { s->color[x][y] = OFFBOARD; continue; }
      s->color[x][y] = EMPTY;
      if (x == 0) s->border[x][y] |= 1u << (2 * BLUE + START);
      else if (x == n - 1) s->border[x][y] |= 1u << (2 * BLUE + END);
      else if (y == 0) s->border[x][y] |= 1u << (2 * RED + START);
      else if (y == n - 1) s->border[x][y] |= 1u << (2 * RED + END);
    }
  init_legal_actions(s);
}

oracle_state* oracle_state_new(const oracle_game* g) {
  oracle_state* s = (oracle_state*)malloc(sizeof(oracle_state));
  state_init(g, s);
  return s;
}
oracle_state* oracle_state_clone(const oracle_state* s) { /* twixt.h:80-82 */
  oracle_state* c = (oracle_state*)malloc(sizeof(oracle_state));
  memcpy(c, s, sizeof(*c));
  return c;
}
void oracle_state_copy(oracle_state* dst, const oracle_state* src) { memcpy(dst, src, sizeof(*dst)); }
void oracle_state_free(oracle_state* s) { free(s); }

int oracle_is_terminal(const oracle_state* s) { /* twixt.h:45-48 */
  return s->result == RED_WIN || s->result == BLUE_WIN || s->result == DRAW;
}
int oracle_current_player(const oracle_state* s) { return s->current_player; } /* twixt.h:38 */

void oracle_returns(const oracle_state* s, double* out2) { /* twixt.h:50-63 */
  if (s->result == RED_WIN) { out2[0] = 1.0; out2[1] = -1.0; }
  else if (s->result == BLUE_WIN) { out2[0] = -1.0; out2[1] = 1.0; }
  else { out2[0] = 0.0; out2[1] = 0.0; }
}

int oracle_legal_list_of(const oracle_game* g, const oracle_state* s, int player, int64_t* out) {
  int n2 = g->n * g->n, c = 0;
  for (int a = 0; a < n2; ++a)
    if (s->in_list[player][a]) { if (out) out[c] = a; ++c; }
  return c;
}

/* twixt.h:86-90 (the lists only shrink by erase or are rebuilt in ascending
 * order, so enumerating the membership set ascending gives the same vector) */
int oracle_legal_actions(const oracle_game* g, const oracle_state* s, int64_t* out) {
  if (oracle_is_terminal(s)) return 0;
  return oracle_legal_list_of(g, s, s->current_player, out);
}

/* twixtboard.cc:633-640 */
static void remove_legal_action(oracle_state* s, int player, int x, int y) {
  int a = x * s->n + y;
  if (s->in_list[player][a]) { s->in_list[player][a] = 0; s->list_count[player]--; }
}

static int has_flag(const oracle_state* s, int x, int y, int player, int border) {
  return (s->border[x][y] >> (2 * player + border)) & 1;
}
static void set_flag(oracle_state* s, int x, int y, int player, int border) {
  s->border[x][y] |= (uint8_t)(1u << (2 * player + border));
}

/* twixtboard.cc:573-588.  The reference's visited set only ever holds cells
 * that already carry the flag, so the flag test alone decides. */
static void explore_local_graph(oracle_state* s, int player, int x, int y, int border) {
  for (int d = 0; d < NDIR; ++d) {
    if (!((s->links[x][y] >> d) & 1)) continue;
    int tx = x + kDx[d], ty = y + kDy[d];
    if (!has_flag(s, tx, ty, player, border)) {
      set_flag(s, tx, ty, player, border);
      explore_local_graph(s, player, tx, ty, border);
    }
  }
}

/* twixtboard.cc:501-571 */
static void set_peg_and_links(const oracle_game* g, oracle_state* s, int player, int x, int y) {
  int n = s->n;
  int linked_to_neutral = 0, new_links = 0;
  s->color[x][y] = (uint8_t)player;
  for (int d = 0; d < NDIR; ++d) {
    int tx = x + kDx[d], ty = y + kDy[d];
    if (off_board(n, tx, ty)) continue;
    if (s->color[tx][ty] != s->color[x][y]) continue;
    int blocked = 0;
    for (int i = 0; i < g->nblock[x][y][d]; ++i) {
      link_ref b = g->block[x][y][d][i];
      if ((s->links[b.x][b.y] >> b.d) & 1) { blocked = 1; break; } /* any colour: :523 */
    }
    if (!blocked) {
      s->links[x][y] |= (uint8_t)(1u << d);
      s->links[tx][ty] |= (uint8_t)(1u << opp_dir(d));
      new_links = 1;
      if (has_flag(s, tx, ty, player, START)) set_flag(s, x, y, player, START);
      else if (has_flag(s, tx, ty, player, END)) set_flag(s, x, y, player, END);
      else linked_to_neutral = 1;
    } else {
      s->blocked[x][y] |= (uint8_t)(1u << d);
      s->blocked[tx][ty] |= (uint8_t)(1u << opp_dir(d));
    }
  }
  if (new_links) {
    if (has_flag(s, x, y, player, START) && linked_to_neutral) explore_local_graph(s, player, x, y, START);
    if (has_flag(s, x, y, player, END) && linked_to_neutral) explore_local_graph(s, player, x, y, END);
  }
}

/* twixtboard.cc:192-207 */
static void update_result(oracle_state* s, int player, int x, int y) {
  if (has_flag(s, x, y, player, START) && has_flag(s, x, y, player, END)) {
    s->result = player == RED ? RED_WIN : BLUE_WIN;
    return;
  }
  if (s->list_count[1 - player] == 0) s->result = DRAW;
}

/* twixtboard.cc:457-499 */
static void board_apply(const oracle_game* g, oracle_state* s, int player, int64_t action) {
  int n = s->n;
  int x = (int)action / n, y = (int)action % n; /* :599-601 */
  if (s->move_counter == 1) {
    if (x == s->move_one_x && y == s->move_one_y) {
      s->swapped = 1;
      s->color[x][y] = EMPTY;  /* UndoFirstMove :450-455 */
      init_legal_actions(s);
      int rx = y, ry = n - x - 1; /* :471-473 */
      x = rx; y = ry;
    } else {
      remove_legal_action(s, RED, s->move_one_x, s->move_one_y);
      remove_legal_action(s, BLUE, s->move_one_x, s->move_one_y);
    }
  }
  set_peg_and_links(g, s, player, x, y);
  if (s->move_counter == 0) {
    s->move_one_x = x; s->move_one_y = y;
  } else {
    remove_legal_action(s, RED, x, y);
    remove_legal_action(s, BLUE, x, y);
  }
  s->move_counter++;
  update_result(s, player, x, y);
}

/* twixt.h:93-104 */
int oracle_apply(const oracle_game* g, oracle_state* s, int64_t action, char* err, int errcap) {
  int legal = 0;
  if (!oracle_is_terminal(s) && action >= 0 && action < (int64_t)g->n * g->n)
    legal = s->in_list[s->current_player][action];
  if (!legal) {
    if (err && errcap > 0) snprintf(err, (size_t)errcap, "Not a legal action: %lld", (long long)action);
    return 1;
  }
  board_apply(g, s, s->current_player, action);
  if (s->result == OPEN) s->current_player = 1 - s->current_player;
  else s->current_player = ORACLE_TERMINAL_PLAYER;
  return 0;
}

/* twixt.cc:76-132 with GetTensorPosition twixtboard.cc:590-597 */
void oracle_observation(const oracle_game* g, const oracle_state* s, float* out) {
  int n = g->n, w = n - 2;
  memset(out, 0, sizeof(float) * (size_t)oracle_observation_size(g));
  for (int c = 0; c < n; ++c)
    for (int r = 0; r < n; ++r) {
      int color = s->color[c][r];
      int offset, tx, ty;
      if (color == RED) { offset = 0; tx = n - r - 1; ty = c - 1; }
      else if (color == BLUE) { offset = ORACLE_NUM_PLANES / 2; tx = n - c - 1; ty = n - r - 2; }
      else continue;
      if (s->links[c][r] > 0) {
        for (int d = 0; d < 4; ++d)
          if ((s->links[c][r] >> d) & 1) out[((offset + 1 + d) * n + tx) * w + ty] = 1.0f;
      } else {
        out[((offset + 0) * n + tx) * w + ty] = 1.0f;
      }
      if ((s->blocked[c][r] & 15u) > 0) out[((offset + 5) * n + tx) * w + ty] = 1.0f;
    }
}

void oracle_board_header(const oracle_state* s, int* out5) {
  out5[0] = s->move_counter; out5[1] = s->swapped; out5[2] = s->result;
  out5[3] = s->move_one_x; out5[4] = s->move_one_y;
}

void oracle_export_cells(const oracle_game* g, const oracle_state* s, int* out) {
  int n = g->n;
  for (int x = 0; x < n; ++x)
    for (int y = 0; y < n; ++y) {
      int i = x * n + y;
      out[4 * i + 0] = s->color[x][y];
      out[4 * i + 1] = s->links[x][y];
      out[4 * i + 2] = s->blocked[x][y];
      out[4 * i + 3] = s->border[x][y];
    }
}

int oracle_record_words(const oracle_game* g) {
  int w = ORACLE_HEADER_WORDS + ORACLE_NUM_STATE_PLANES * g->n;
  return (w + 3) & ~3;
}

/* The packed record of include/twixt_b200.h, derived from the cell arrays.
 * planes: 0 red pegs, 1 blue pegs, 2..5 links NNE/ENE/ESE/SSE at the west
 * endpoint, 6/7 peg is linked to its owner's start/end border line, 8 peg has
 * a blocked neighbour in an east direction. */
void oracle_export_record(const oracle_game* g, const oracle_state* s, uint32_t* out) {
  int n = g->n;
  memset(out, 0, sizeof(uint32_t) * (size_t)oracle_record_words(g));
  int cnt[2] = {0, 0};
  uint32_t* pl = out + ORACLE_HEADER_WORDS;
  for (int x = 0; x < n; ++x)
    for (int y = 0; y < n; ++y) {
      int c = s->color[x][y];
      if (c == OFFBOARD) continue;
      if (c == EMPTY) {
        if (x >= 1 && x <= n - 2) cnt[RED]++;
        if (y >= 1 && y <= n - 2) cnt[BLUE]++;
        continue;
      }
      uint32_t bit = 1u << y;
      pl[c * n + x] |= bit;
      for (int d = 0; d < 4; ++d)
        if ((s->links[x][y] >> d) & 1) pl[(2 + d) * n + x] |= bit;
      if (has_flag(s, x, y, c, START)) pl[6 * n + x] |= bit;
      if (has_flag(s, x, y, c, END)) pl[7 * n + x] |= bit;
      if (s->blocked[x][y] & 15u) pl[8 * n + x] |= bit;
    }
  out[0] = (uint32_t)s->move_counter;
  out[1] = (uint32_t)s->result | ((uint32_t)(s->swapped ? 1 : 0) << 2);
  out[2] = s->move_counter == 0 ? 0xFFFFFFFFu : (uint32_t)(s->move_one_x * n + s->move_one_y);
  out[3] = (uint32_t)cnt[RED] | ((uint32_t)cnt[BLUE] << 16);
}

int oracle_replay(const oracle_game* g, oracle_state* s, const int64_t* actions, int len) {
  int i = 0;
  for (; i < len; ++i)
    if (oracle_apply(g, s, actions[i], NULL, 0) != 0) break;
  return i;
}

int oracle_playout_philox(const oracle_game* g, oracle_state* s, uint64_t seed, uint64_t stream,
                          int max_plies, int64_t* actions_out) {
  int64_t legal[NB * NB];
  int step = 0;
  while (!oracle_is_terminal(s) && step < max_plies) {
    int L = oracle_legal_actions(g, s, legal);
    uint32_t word = oracle_playout_word(seed, stream, (uint32_t)step);
    int64_t a = legal[oracle_playout_index(word, (uint32_t)L)];
    if (actions_out) actions_out[step] = a;
    oracle_apply(g, s, a, NULL, 0);
    ++step;
  }
  return step;
}

int64_t oracle_bench_playouts(const oracle_game* g, double seconds, uint64_t seed, int64_t* games_out,
                              double* elapsed_out) {
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  oracle_state* proto = oracle_state_new(g);
  oracle_state* s = oracle_state_clone(proto);
  int64_t plies = 0, games = 0;
  double el = 0.0;
  for (;;) {
    oracle_state_copy(s, proto);
    plies += oracle_playout_philox(g, s, seed, (uint64_t)games, 1 << 30, NULL);
    ++games;
    clock_gettime(CLOCK_MONOTONIC, &t1);
    el = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
    if (el >= seconds) break;
  }
  oracle_state_free(s);
  oracle_state_free(proto);
  if (games_out) *games_out = games;
  if (elapsed_out) *elapsed_out = el;
  return plies;
}

void oracle_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  oracle_philox4x32_10(ctr, key, out);
}
