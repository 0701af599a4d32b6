// TEST INFRASTRUCTURE ONLY (oracle/): minimal stand-in for the part of
// deepmind/open_spiel's public API that the reference plug-in
// (/root/reference/open_spiel/games/twixt/twixt.h:31-146, twixt.cc:35-145)
// touches.  Upstream open_spiel is un-vendored and un-pinned by the reference
// (README.md:10), is not installed in this image and cannot be fetched, so the
// reference sources are compiled UNMODIFIED against this shim instead.  The
// shim contains no game arithmetic: it only supplies base classes, parameter
// plumbing and an error hook.  Written from scratch for this repo.
#ifndef ORACLE_SHIM_OPEN_SPIEL_SPIEL_H_
#define ORACLE_SHIM_OPEN_SPIEL_SPIEL_H_

#include <algorithm>
#include <array>
#include <cstddef>
#include <cstdint>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace absl {

// Just enough of absl::Span<T> for ObservationTensor(player, Span<float>).
template <typename T>
class Span {
 public:
  Span() : ptr_(nullptr), len_(0) {}
  Span(T* ptr, std::size_t len) : ptr_(ptr), len_(len) {}
  template <typename A>
  Span(std::vector<A>& v) : ptr_(v.data()), len_(v.size()) {}  // NOLINT
  T* data() const { return ptr_; }
  std::size_t size() const { return len_; }
  T& operator[](std::size_t i) const { return ptr_[i]; }
  T* begin() const { return ptr_; }
  T* end() const { return ptr_ + len_; }

 private:
  T* ptr_;
  std::size_t len_;
};

// Just enough of absl::optional<T> for Game::UtilitySum().
template <typename T>
class optional {
 public:
  optional() : has_(false), value_() {}
  optional(const T& v) : has_(true), value_(v) {}  // NOLINT
  bool has_value() const { return has_; }
  const T& value() const { return value_; }
  const T& operator*() const { return value_; }

 private:
  bool has_;
  T value_;
};

}  // namespace absl

namespace open_spiel {

using Action = int64_t;
using Player = int;

constexpr Player kTerminalPlayerId = -4;  // playthrough.txt:678

// The error hook.  Upstream aborts unless a handler is installed
// (twixt_test.cc:43-46); the shim always throws so a driver can catch it.
class SpielError : public std::runtime_error {
 public:
  explicit SpielError(const std::string& m) : std::runtime_error(m) {}
};

[[noreturn]] inline void SpielFatalError(const std::string& msg) {
  throw SpielError(msg);
}

#define SPIEL_SHIM_CHECK_OP(a, op, b)                                        \
  do {                                                                       \
    if (!((a)op(b)))                                                         \
      ::open_spiel::SpielFatalError(std::string("CHECK failed: ") + #a +     \
                                    " " #op " " + #b);                       \
  } while (0)
#define SPIEL_CHECK_GE(a, b) SPIEL_SHIM_CHECK_OP(a, >=, b)
#define SPIEL_CHECK_GT(a, b) SPIEL_SHIM_CHECK_OP(a, >, b)
#define SPIEL_CHECK_LE(a, b) SPIEL_SHIM_CHECK_OP(a, <=, b)
#define SPIEL_CHECK_LT(a, b) SPIEL_SHIM_CHECK_OP(a, <, b)
#define SPIEL_CHECK_EQ(a, b) SPIEL_SHIM_CHECK_OP(a, ==, b)
#define SPIEL_CHECK_NE(a, b) SPIEL_SHIM_CHECK_OP(a, !=, b)
#define SPIEL_CHECK_TRUE(a) SPIEL_SHIM_CHECK_OP(static_cast<bool>(a), ==, true)
#define SPIEL_CHECK_FALSE(a) SPIEL_SHIM_CHECK_OP(static_cast<bool>(a), ==, false)

// A game parameter holding either an int or a bool (the only kinds the
// reference declares, twixt.cc:50-51).
class GameParameter {
 public:
  enum class Kind { kUnset, kInt, kBool };
  GameParameter() : kind_(Kind::kUnset), int_(0), bool_(false) {}
  explicit GameParameter(int v, bool mandatory = false)
      : kind_(Kind::kInt), int_(v), bool_(false) { (void)mandatory; }
  explicit GameParameter(bool v, bool mandatory = false)
      : kind_(Kind::kBool), int_(0), bool_(v) { (void)mandatory; }
  Kind kind() const { return kind_; }
  int int_value() const { return int_; }
  bool bool_value() const { return bool_; }

 private:
  Kind kind_;
  int int_;
  bool bool_;
};

using GameParameters = std::map<std::string, GameParameter>;

struct GameType {
  enum class Dynamics { kSimultaneous, kSequential };
  enum class ChanceMode { kDeterministic, kExplicitStochastic, kSampledStochastic };
  enum class Information { kOneShot, kPerfectInformation, kImperfectInformation };
  enum class Utility { kZeroSum, kConstantSum, kGeneralSum, kIdentical };
  enum class RewardModel { kRewards, kTerminal };

  std::string short_name;
  std::string long_name;
  Dynamics dynamics;
  ChanceMode chance_mode;
  Information information;
  Utility utility;
  RewardModel reward_model;
  int max_num_players;
  int min_num_players;
  bool provides_information_state_string;
  bool provides_information_state_tensor;
  bool provides_observation_string;
  bool provides_observation_tensor;
  GameParameters parameter_specification;
};

class State;

class Game : public std::enable_shared_from_this<Game> {
 public:
  virtual ~Game() = default;
  virtual std::unique_ptr<State> NewInitialState() const = 0;
  virtual int NumDistinctActions() const = 0;
  virtual int NumPlayers() const = 0;
  virtual double MinUtility() const = 0;
  virtual double MaxUtility() const = 0;
  virtual absl::optional<double> UtilitySum() const { return absl::optional<double>(); }
  virtual std::vector<int> ObservationTensorShape() const { return {}; }
  const GameType& GetType() const { return type_; }
  const GameParameters& GetParameters() const { return params_; }

 protected:
  Game(const GameType& type, const GameParameters& params)
      : type_(type), params_(params) {
    for (const auto& kv : params_) {
      if (type_.parameter_specification.count(kv.first) == 0) {
        std::string names;
        for (const auto& spec : type_.parameter_specification) {
          if (!names.empty()) names += ", ";
          names += spec.first;
        }
        SpielFatalError("Unknown parameter '" + kv.first +
                        "'. Available parameters are: " + names);
      }
    }
  }

  template <typename T>
  T ParameterValue(const std::string& name, T default_value) const;

 private:
  GameType type_;
  GameParameters params_;
};

template <>
inline int Game::ParameterValue<int>(const std::string& name, int dflt) const {
  auto it = params_.find(name);
  return it == params_.end() ? dflt : it->second.int_value();
}
template <>
inline bool Game::ParameterValue<bool>(const std::string& name, bool dflt) const {
  auto it = params_.find(name);
  return it == params_.end() ? dflt : it->second.bool_value();
}

class State {
 public:
  explicit State(std::shared_ptr<const Game> game) : game_(std::move(game)) {}
  State(const State&) = default;
  State& operator=(const State&) = default;
  virtual ~State() = default;

  virtual Player CurrentPlayer() const = 0;
  virtual std::string ActionToString(Player player, Action action) const = 0;
  virtual std::string ToString() const = 0;
  virtual bool IsTerminal() const = 0;
  virtual std::vector<double> Returns() const = 0;
  virtual std::string InformationStateString(Player) const { return ""; }
  virtual std::string ObservationString(Player) const { return ""; }
  virtual void ObservationTensor(Player, absl::Span<float>) const {}
  virtual std::unique_ptr<State> Clone() const = 0;
  virtual void UndoAction(Player, Action) {}
  virtual std::vector<Action> LegalActions() const = 0;

  // Upstream State::ApplyAction records the move and forwards to the game.
  void ApplyAction(Action action) {
    Player p = CurrentPlayer();
    DoApplyAction(action);
    history_.push_back({p, action});
  }
  double PlayerReturn(Player p) const { return Returns()[p]; }
  std::vector<Action> History() const {
    std::vector<Action> h;
    for (const auto& pa : history_) h.push_back(pa.second);
    return h;
  }
  std::shared_ptr<const Game> GetGame() const { return game_; }

 protected:
  virtual void DoApplyAction(Action action) = 0;
  std::shared_ptr<const Game> game_;
  std::vector<std::pair<Player, Action>> history_;
};

// The reference registers itself with upstream's game registry
// (twixt.cc:58); the oracle constructs TwixTGame directly, so registration is
// reduced to keeping the factory referenced.
#define REGISTER_SPIEL_GAME(type, factory)                                 \
  static const void* const spiel_shim_registered_factory_ [[maybe_unused]] = \
      reinterpret_cast<const void*>(&factory)

}  // namespace open_spiel

#endif  // ORACLE_SHIM_OPEN_SPIEL_SPIEL_H_
