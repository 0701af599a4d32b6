// TEST INFRASTRUCTURE ONLY (oracle/): a C interface over the UNMODIFIED
// reference implementation, compiled from where it lies under /root/reference
// against the header shim in oracle/shim (see oracle/Makefile).  Nothing in
// the product path may call this.  It exists so that tests can (a) pin the C
// restatement in oracle/twixt_oracle.c against the reference's own behaviour,
// (b) generate the golden fixtures under tests/golden/, and (c) time the
// reference's CPU path for bench.py's reference arm.
//
// The reference keeps its crossing table in a process-global static map that
// every Board constructor clears and rebuilds (twixtboard.h:142-151,
// twixtboard.cc:148-165, 211): never keep live states of two different board
// sizes, and never construct states from two threads.
#include <algorithm>
#include <array>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <iostream>
#include <map>
#include <memory>
#include <set>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

#include "open_spiel/spiel.h"
#include "open_spiel/spiel_utils.h"
#include "open_spiel/utils/tensor_view.h"

// The board internals (cells, move counter) are private in the reference; the
// oracle needs them to compare complete states, not only the public surface.
// Access control does not change object layout, so the driver alone is
// compiled with the keyword neutralised.
#define private public
#define protected public
#include "open_spiel/games/twixt/twixt.h"
#undef private
#undef protected

#include "philox.h"

namespace {

using open_spiel::Action;
using open_spiel::twixt::TwixTGame;
using open_spiel::twixt::TwixTState;

struct RefGame {
  std::shared_ptr<const open_spiel::Game> game;
  int n;
};

void CopyMessage(const std::string& msg, char* err, int cap) {
  if (err == nullptr || cap <= 0) return;
  std::snprintf(err, static_cast<size_t>(cap), "%s", msg.c_str());
}

TwixTState* AsState(void* s) { return static_cast<TwixTState*>(s); }

}  // namespace

extern "C" {

// TwixTGame construction incl. the board_size range check (twixt.cc:134-145).
void* ref_game_new(int board_size, int ansi, char* err, int errcap) {
  try {
    open_spiel::GameParameters params;
    params.insert({"board_size", open_spiel::GameParameter(board_size, false)});
    params.insert({"ansi_color_output", open_spiel::GameParameter(ansi != 0, false)});
    auto* g = new RefGame();
    g->game = std::shared_ptr<const open_spiel::Game>(new TwixTGame(params));
    g->n = board_size;
    return g;
  } catch (const std::exception& e) {
    CopyMessage(e.what(), err, errcap);
    return nullptr;
  }
}

// Unknown-parameter path (twixt_test.cc:85-91).
int ref_game_new_with_param(const char* name, int value, char* err, int errcap) {
  try {
    open_spiel::GameParameters params;
    params.insert({name, open_spiel::GameParameter(value, false)});
    TwixTGame g(params);
    return 0;
  } catch (const std::exception& e) {
    CopyMessage(e.what(), err, errcap);
    return 1;
  }
}

void ref_game_free(void* g) { delete static_cast<RefGame*>(g); }

// out[0]=NumDistinctActions out[1]=MaxGameLength out[2]=NumPlayers
// out[3..5]= {kNumPlanes, n, n-2} computed here, NOT via the reference's
// ObservationTensorShape(), whose function-static caches the first game's
// shape (twixt.h:131-134).
void ref_game_info(void* gp, int* out) {
  auto* g = static_cast<RefGame*>(gp);
  const auto& tg = static_cast<const TwixTGame&>(*g->game);
  out[0] = tg.NumDistinctActions();
  out[1] = tg.MaxGameLength();
  out[2] = tg.NumPlayers();
  out[3] = open_spiel::twixt::kNumPlanes;
  out[4] = g->n;
  out[5] = g->n - 2;
}

void ref_game_utils(void* gp, double* out) {
  auto* g = static_cast<RefGame*>(gp);
  out[0] = g->game->MinUtility();
  out[1] = g->game->MaxUtility();
  out[2] = g->game->UtilitySum().value();
}

void* ref_state_new(void* gp) {
  auto* g = static_cast<RefGame*>(gp);
  return g->game->NewInitialState().release();
}
void* ref_state_clone(void* s) { return AsState(s)->Clone().release(); }
void ref_state_free(void* s) { delete AsState(s); }

int ref_legal_actions(void* s, int64_t* out, int cap) {
  std::vector<Action> v = AsState(s)->LegalActions();
  int n = static_cast<int>(v.size());
  for (int i = 0; i < n && i < cap; ++i) out[i] = v[i];
  return n;
}

int ref_apply(void* s, int64_t action, char* err, int errcap) {
  try {
    AsState(s)->ApplyAction(action);
    return 0;
  } catch (const std::exception& e) {
    CopyMessage(e.what(), err, errcap);
    return 1;
  }
}

int ref_current_player(void* s) { return AsState(s)->CurrentPlayer(); }
int ref_is_terminal(void* s) { return AsState(s)->IsTerminal() ? 1 : 0; }
void ref_returns(void* s, double* out2) {
  std::vector<double> r = AsState(s)->Returns();
  out2[0] = r[0];
  out2[1] = r[1];
}

int ref_observation(void* s, int player, float* out, int len, char* err, int errcap) {
  try {
    AsState(s)->ObservationTensor(player, absl::Span<float>(out, static_cast<size_t>(len)));
    return 0;
  } catch (const std::exception& e) {
    CopyMessage(e.what(), err, errcap);
    return 1;
  }
}

int ref_to_string(void* s, char* buf, int cap) {
  std::string str = AsState(s)->ToString();
  int n = static_cast<int>(str.size());
  if (buf != nullptr && cap > 0) {
    int c = std::min(n, cap - 1);
    std::memcpy(buf, str.data(), static_cast<size_t>(c));
    buf[c] = 0;
  }
  return n;
}

int ref_action_to_string(void* s, int player, int64_t action, char* buf, int cap) {
  std::string str = AsState(s)->ActionToString(player, action);
  std::snprintf(buf, static_cast<size_t>(cap), "%s", str.c_str());
  return static_cast<int>(str.size());
}

// Board internals: out[0]=move_counter out[1]=swapped out[2]=result
// out[3]=move_one.x out[4]=move_one.y (garbage before the first move).
void ref_board_header(void* s, int* out) {
  const auto& b = AsState(s)->board_;
  out[0] = b.move_counter_;
  out[1] = b.swapped_ ? 1 : 0;
  out[2] = b.result_;
  out[3] = b.move_one_.x;
  out[4] = b.move_one_.y;
}

// Per cell, index x*n+y: out[4*i+0]=color out[4*i+1]=links (8-bit, Compass
// order) out[4*i+2]=blocked_neighbors out[4*i+3]= border flags, bit0
// red/start bit1 red/end bit2 blue/start bit3 blue/end (twixtcell.h:70-109).
void ref_export_cells(void* s, int* out) {
  const auto& b = AsState(s)->board_;
  int n = b.size();
  for (int x = 0; x < n; ++x) {
    for (int y = 0; y < n; ++y) {
      const auto& c = b.cell_[x][y];
      int i = x * n + y;
      out[4 * i + 0] = c.color_;
      out[4 * i + 1] = c.links_;
      out[4 * i + 2] = c.blocked_neighbors_;
      out[4 * i + 3] = (c.linked_to_border_[0][0] ? 1 : 0) | (c.linked_to_border_[0][1] ? 2 : 0) |
                       (c.linked_to_border_[1][0] ? 4 : 0) | (c.linked_to_border_[1][1] ? 8 : 0);
    }
  }
}

// The reference's crossing relation for the directed link (x,y,dir) of the board
// size most recently constructed (BlockerMap is process-global,
// twixtboard.h:142-151): triples (x,y,dir), both encodings of every crosser.
int ref_blockers(int x, int y, int dir, int* out, int cap) {
  Link link;
  link.position = {x, y};
  link.direction = dir;
  const std::set<Link>& bl = open_spiel::twixt::BlockerMap::GetBlockers(link);
  int i = 0;
  for (const Link& b : bl) {
    if (i < cap) {
      out[3 * i + 0] = b.position.x;
      out[3 * i + 1] = b.position.y;
      out[3 * i + 2] = b.direction;
    }
    ++i;
  }
  return i;
}

// Both maintained legal lists (not only the mover's): player p's list.
int ref_legal_list_of(void* s, int player, int64_t* out, int cap) {
  std::vector<Action> v = AsState(s)->board_.GetLegalActions(player);
  int n = static_cast<int>(v.size());
  for (int i = 0; i < n && i < cap; ++i) out[i] = v[i];
  return n;
}

// Apply a whole action sequence; returns the number applied before an error
// (== len when all were legal).
int ref_replay(void* s, const int64_t* actions, int len) {
  int i = 0;
  try {
    for (; i < len; ++i) AsState(s)->ApplyAction(actions[i]);
  } catch (const std::exception&) {
  }
  return i;
}

// Play the Philox policy of oracle/philox.h from state s (mutated) until the
// game ends or max_plies moves were made.  actions_out may be null.  Returns
// the number of moves made.
int ref_playout_philox(void* s, uint64_t seed, uint64_t stream, int max_plies,
                       int64_t* actions_out) {
  TwixTState* st = AsState(s);
  int step = 0;
  while (!st->IsTerminal() && step < max_plies) {
    std::vector<Action> legal = st->LegalActions();
    uint32_t word = oracle_playout_word(seed, stream, static_cast<uint32_t>(step));
    uint32_t idx = oracle_playout_index(word, static_cast<uint32_t>(legal.size()));
    Action a = legal[idx];
    if (actions_out != nullptr) actions_out[step] = a;
    st->ApplyAction(a);
    ++step;
  }
  return step;
}

}  // extern "C"
