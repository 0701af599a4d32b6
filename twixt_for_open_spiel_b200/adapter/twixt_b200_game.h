// twixt_b200_game.h -- open_spiel plug-in adapter over libtwixt_b200 (C ABI in
// include/twixt_b200.h).  Drop-in replacement for the reference's
//   class TwixTState : public State   (open_spiel/games/twixt/twixt.h:31-112)
//   class TwixTGame  : public Game    (twixt.h:114-146)
// with the same short name, parameters and error texts (twixt.cc:35-58,
// 134-145), so `LoadGame("twixt(board_size=12)")`, upstream example.cc,
// mcts_example.cc and the game's own tests keep working while every rule
// evaluation runs on the GPU.  One State = one env slot of a device pool owned
// by the Game; Clone() is a device-side record copy.
//
// Per-call cost.  A State keeps the answer of the last twixt_step -- current player, terminal flag, returns
// and the legal-action list of its position -- so ApplyAction is ONE kernel launch + one synchronisation and
// CurrentPlayer / IsTerminal / Returns / LegalActions are served from the host without touching the GPU
// (the loop of example.cc / mcts_example.cc calls all of them between two moves).  Clone() is an
// asynchronous device copy plus a copy of that cached answer.  ObservationTensor and ToString read the env.
//
// To build inside an open_spiel checkout: copy this directory to
// open_spiel/games/twixt_b200/, add twixt_b200_game.cc to GAME_SOURCES and link
// libtwixt_b200.so.  Here it is compile- and run-tested
// against the minimal open_spiel header shim used by the oracle.
#ifndef TWIXT_B200_ADAPTER_GAME_H_
#define TWIXT_B200_ADAPTER_GAME_H_

#include <memory>
#include <mutex>
#include <string>
#include <utility>
#include <vector>

#include "open_spiel/spiel.h"
#include "twixt_b200.h"

namespace open_spiel {
namespace twixt_b200 {

inline constexpr int kNumPlayers = 2;
inline constexpr int kDefaultBoardSize = TWIXT_DEFAULT_BOARD_SIZE;  // twixtboard.h:34
inline constexpr bool kDefaultAnsiColorOutput = true;               // twixtboard.h:36
inline constexpr int kDefaultPoolSize = 256;

class TwixTB200Game;

// Env slots of one game, grown a batch at a time.
//
// THREAD SAFETY.  open_spiel callers (AlphaZero actors, parallel MCTS) use States of one shared Game from
// several threads, while a twixt_batch is not thread-safe (include/twixt_b200.h: it owns one stream, one
// set of staging buffers and one error slot).  Up to pool_size States share a batch, so every twixt_* call a
// State makes is serialised on that batch's mutex (BatchLock below); the free list has its own.  Guarantee:
// distinct States may be used concurrently from different threads, one State from one thread at a time --
// the same contract as the reference's State objects.  Calls on States of different batches do not contend.
class EnvPool {
 public:
  EnvPool(int board_size, int device, int pool_size);
  ~EnvPool();
  struct Slot {
    twixt_batch* batch;
    int64_t index;
    std::mutex* mu;  // serialises all calls into `batch`
  };
  Slot Take();
  void Give(Slot s);

 private:
  int board_size_, device_, pool_size_;
  std::mutex mu_;
  std::vector<twixt_batch*> batches_;
  std::vector<std::unique_ptr<std::mutex>> batch_mu_;
  std::vector<Slot> free_;
};

class TwixTB200State : public State {
 public:
  explicit TwixTB200State(std::shared_ptr<const Game> game);
  TwixTB200State(const TwixTB200State& other);
  TwixTB200State& operator=(const TwixTB200State&) = delete;
  ~TwixTB200State() override;

  Player CurrentPlayer() const override;
  std::string ActionToString(Player player, Action action) const override;
  std::string ToString() const override;
  bool IsTerminal() const override;
  std::vector<double> Returns() const override;
  std::string InformationStateString(Player player) const override;
  std::string ObservationString(Player player) const override;
  void ObservationTensor(Player player, absl::Span<float> values) const override;
  std::unique_ptr<State> Clone() const override;
  void UndoAction(Player, Action) override {}  // a stub in the reference too (twixt.h:84)
  std::vector<Action> LegalActions() const override;

 protected:
  void DoApplyAction(Action action) override;

 private:
  const TwixTB200Game& parent() const;
  void Step(int32_t action);  // twixt_step: (reset | apply | nothing), then refresh the cached answer
  EnvPool::Slot slot_;
  twixt_step_result now_;       // of the current position
  std::vector<Action> legal_;   // ascending; empty when terminal (twixt.h:86-90)
};

class TwixTB200Game : public Game {
 public:
  explicit TwixTB200Game(const GameParameters& params);

  std::unique_ptr<State> NewInitialState() const override {
    return std::unique_ptr<State>(new TwixTB200State(shared_from_this()));
  }
  int NumDistinctActions() const override { return info_.num_distinct_actions; }
  int NumPlayers() const override { return kNumPlayers; }
  double MinUtility() const override { return info_.min_utility; }
  absl::optional<double> UtilitySum() const override { return info_.utility_sum; }
  double MaxUtility() const override { return info_.max_utility; }
  std::vector<int> ObservationTensorShape() const override {
    return {info_.obs_shape[0], info_.obs_shape[1], info_.obs_shape[2]};  // per game, not a function-static
  }
  int MaxGameLength() const { return info_.max_game_length; }
  bool ansi_color_output() const { return ansi_color_output_; }
  int board_size() const { return board_size_; }
  EnvPool& pool() const { return *pool_; }

 private:
  bool ansi_color_output_;
  int board_size_;
  twixt_game_info info_;
  std::unique_ptr<EnvPool> pool_;
};

// The picture of Board::ToString (twixtboard.cc:278-448) from a state record.
std::string RenderRecord(const uint32_t* record, int board_size, bool ansi_color_output);

}  // namespace twixt_b200
}  // namespace open_spiel

#endif  // TWIXT_B200_ADAPTER_GAME_H_
