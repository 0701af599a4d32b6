// twixt_b200_game.cc -- see twixt_b200_game.h.
#include "twixt_b200_game.h"

#include <algorithm>
#include <array>
#include <cstdint>

namespace open_spiel {
namespace twixt_b200 {
namespace {

// Facts about the game: identical to the reference's kGameType (twixt.cc:35-52).
const GameType kGameType{
    /*short_name=*/"twixt",
    /*long_name=*/"TwixT",
    GameType::Dynamics::kSequential,
    GameType::ChanceMode::kDeterministic,
    GameType::Information::kPerfectInformation,
    GameType::Utility::kZeroSum,
    GameType::RewardModel::kTerminal,
    /*max_num_players=*/2,
    /*min_num_players=*/2,
    /*provides_information_state_string=*/true,
    /*provides_information_state_tensor=*/false,
    /*provides_observation_string=*/true,
    /*provides_observation_tensor=*/true,
    /*parameter_specification=*/
    {{"board_size", GameParameter(kDefaultBoardSize)},
     {"ansi_color_output", GameParameter(kDefaultAnsiColorOutput)}},
};

std::unique_ptr<Game> Factory(const GameParameters& params) {
  return std::unique_ptr<Game>(new TwixTB200Game(params));
}

REGISTER_SPIEL_GAME(kGameType, Factory);

void Check(int rc) {
  if (rc != TWIXT_OK) SpielFatalError(twixt_last_error());  // (thread-local text: read before the lock is dropped)
}

// Holds the mutex of a slot's batch for the duration of one twixt_* call (see EnvPool in the header).
using BatchLock = std::lock_guard<std::mutex>;

}  // namespace

// ---------------------------------------------------------------- pool ----
EnvPool::EnvPool(int board_size, int device, int pool_size)
    : board_size_(board_size), device_(device), pool_size_(pool_size) {}

EnvPool::~EnvPool() {
  for (twixt_batch* b : batches_) twixt_destroy(b);
}

EnvPool::Slot EnvPool::Take() {
  std::lock_guard<std::mutex> lock(mu_);
  if (free_.empty()) {
    twixt_batch* b = nullptr;
    Check(twixt_create(board_size_, pool_size_, device_, /*seed=*/0, &b));
    batches_.push_back(b);
    batch_mu_.push_back(std::make_unique<std::mutex>());
    for (int64_t i = pool_size_ - 1; i >= 0; --i) free_.push_back({b, i, batch_mu_.back().get()});
  }
  Slot s = free_.back();
  free_.pop_back();
  return s;
}

void EnvPool::Give(Slot s) {
  std::lock_guard<std::mutex> lock(mu_);
  free_.push_back(s);
}

// ---------------------------------------------------------------- game ----
TwixTB200Game::TwixTB200Game(const GameParameters& params)
    : Game(kGameType, params),
      ansi_color_output_(ParameterValue<bool>("ansi_color_output", kDefaultAnsiColorOutput)),
      board_size_(ParameterValue<int>("board_size", kDefaultBoardSize)) {
  // "board_size out of range [5..24]: N", same text as twixt.cc:139-144
  Check(twixt_game_info_for(board_size_, &info_));
  pool_ = std::make_unique<EnvPool>(board_size_, /*device=*/0, kDefaultPoolSize);
}

// --------------------------------------------------------------- state ----
const TwixTB200Game& TwixTB200State::parent() const { return static_cast<const TwixTB200Game&>(*game_); }

TwixTB200State::TwixTB200State(std::shared_ptr<const Game> game) : State(std::move(game)) {
  slot_ = parent().pool().Take();
  Step(TWIXT_STEP_RESET);  // NewInitialState (twixt.h:118-120): reset + the initial position's answers, one launch
}

TwixTB200State::TwixTB200State(const TwixTB200State& other)
    : State(other), now_(other.now_), legal_(other.legal_) {
  slot_ = parent().pool().Take();
  if (slot_.batch == other.slot_.batch) {
    BatchLock lock(*slot_.mu);
    Check(twixt_clone(slot_.batch, other.slot_.index, slot_.index, 1));  // asynchronous: ordered on the batch's stream
  } else {
    // two batches: both streams are involved (the copy waits for the source's pending work); std::lock
    // takes the two mutexes without a lock-order deadlock
    std::unique_lock<std::mutex> a(*slot_.mu, std::defer_lock), b(*other.slot_.mu, std::defer_lock);
    std::lock(a, b);
    Check(twixt_clone_from(slot_.batch, slot_.index, other.slot_.batch, other.slot_.index, 1));
  }
}

TwixTB200State::~TwixTB200State() { parent().pool().Give(slot_); }

void TwixTB200State::Step(int32_t action) {
  const int n = parent().board_size();
  std::vector<Action> legal(static_cast<size_t>(n * (n - 2)));
  twixt_step_result res;
  static_assert(sizeof(Action) == 8, "open_spiel::Action is int64");
  {
    BatchLock lock(*slot_.mu);
    Check(twixt_step(slot_.batch, slot_.index, action, &res, legal.data()));  // "Not a legal action: N" (twixt.h:96)
  }
  legal.resize(static_cast<size_t>(res.num_legal));
  now_ = res;
  legal_ = std::move(legal);
}

Player TwixTB200State::CurrentPlayer() const { return now_.current_player; }  // -4 == kTerminalPlayerId when over

bool TwixTB200State::IsTerminal() const { return now_.is_terminal != 0; }

std::vector<double> TwixTB200State::Returns() const {
  return {static_cast<double>(now_.returns[0]), static_cast<double>(now_.returns[1])};
}

std::vector<Action> TwixTB200State::LegalActions() const { return legal_; }

void TwixTB200State::DoApplyAction(Action action) {
  if (action < 0 || action > INT32_MAX) SpielFatalError("Not a legal action: " + std::to_string(action));
  Step(static_cast<int32_t>(action));
}

void TwixTB200State::ObservationTensor(Player player, absl::Span<float> values) const {
  SPIEL_CHECK_GE(player, 0);
  SPIEL_CHECK_LT(player, kNumPlayers);
  const int n = parent().board_size();
  SPIEL_CHECK_EQ(static_cast<int>(values.size()), TWIXT_NUM_OBS_PLANES * n * (n - 2));
  BatchLock lock(*slot_.mu);
  Check(twixt_observation(slot_.batch, slot_.index, 1, values.data()));
}

std::unique_ptr<State> TwixTB200State::Clone() const {
  return std::unique_ptr<State>(new TwixTB200State(*this));
}

std::string TwixTB200State::ActionToString(Player player, Action action) const {  // twixt.cc:67-74
  const int n = parent().board_size();
  std::string s = (player == 0) ? "x" : "o";
  s += static_cast<char>('a' + static_cast<int>(action) / n);
  s.append(std::to_string(n - static_cast<int>(action) % n));
  return s;
}

std::string TwixTB200State::ToString() const {
  twixt_game_info info;
  Check(twixt_get_info(slot_.batch, &info));
  std::vector<uint32_t> rec(static_cast<size_t>(info.record_words));
  {
    BatchLock lock(*slot_.mu);
    Check(twixt_export_state(slot_.batch, slot_.index, 1, rec.data()));
  }
  return RenderRecord(rec.data(), parent().board_size(), parent().ansi_color_output());
}

std::string TwixTB200State::InformationStateString(Player player) const {
  SPIEL_CHECK_GE(player, 0);
  SPIEL_CHECK_LT(player, kNumPlayers);
  return ToString();
}

std::string TwixTB200State::ObservationString(Player player) const {
  SPIEL_CHECK_GE(player, 0);
  SPIEL_CHECK_LT(player, kNumPlayers);
  return ToString();
}

// ------------------------------------------------------------ renderer ----
namespace {

constexpr char kAnsiRed[] = "\x1b[91m";
constexpr char kAnsiBlue[] = "\x1b[94m";
constexpr char kAnsiDefault[] = "\x1b[0m";
enum Dir { kNNE, kENE, kESE, kSSE, kSSW, kWSW, kWNW, kNNW };
constexpr int kDx[8] = {1, 2, 2, 1, -1, -2, -2, -1};
constexpr int kDy[8] = {2, 1, -1, -2, -2, -1, 1, 2};

// A slot of the picture: glyphs of the links passing through it.  `always`
// entries are all drawn, `fallback` entries only while the slot is empty.
struct Glyph {
  int dx, dy, dir;
  char ch;
};
struct SlotSpec {
  std::array<Glyph, 3> always;
  int n_always;
  Glyph fallback;
  bool has_fallback;
};
constexpr Glyph kNone{0, 0, 0, ' '};
constexpr SlotSpec kBefore[3] = {
    {{{{-1, 0, kENE, '/'}, {-1, -1, kNNE, '/'}, {0, 0, kWNW, '_'}}}, 3, kNone, false},
    {{{{0, 0, kNNE, '|'}, kNone, kNone}}, 1, {0, 0, kNNW, '|'}, true},
    {{{{1, 0, kWNW, '\\'}, {1, -1, kNNW, '\\'}, {0, 0, kENE, '_'}}}, 3, kNone, false},
};
constexpr SlotSpec kPegLeft = {{{{-1, -1, kNNE, '|'}, {0, 0, kWSW, '_'}, kNone}}, 2, kNone, false};
constexpr SlotSpec kPegRight = {{{{1, -1, kNNW, '|'}, {0, 0, kESE, '_'}, kNone}}, 2, kNone, false};
constexpr SlotSpec kAfter[3] = {
    {{{{1, -1, kWNW, '\\'}, {0, -1, kNNW, '\\'}, kNone}}, 2, kNone, false},
    {{{{-1, -1, kENE, '_'}, {1, -1, kWNW, '_'}, {0, 0, kSSW, '|'}}}, 3, {0, 0, kSSE, '|'}, true},
    {{{{-1, -1, kENE, '/'}, {0, -1, kNNE, '/'}, kNone}}, 2, kNone, false},
};

struct Picture {
  int n;
  bool ansi;
  std::vector<int> color;  // 0 red 1 blue 2 empty
  std::vector<int> links;  // 8-bit Compass mask per cell

  bool OffBoard(int x, int y) const {
    return x < 0 || y < 0 || x >= n || y >= n || ((x == 0 || x == n - 1) && (y == 0 || y == n - 1));
  }
  void Paint(std::string* s, const char* color_code, const std::string& text) const {
    if (ansi) s->append(color_code);
    s->append(text);
    if (ansi) s->append(kAnsiDefault);
  }
  void LinkGlyph(std::string* s, int x, int y, int dir, char ch) const {
    if (OffBoard(x, y) || !((links[x * n + y] >> dir) & 1)) return;
    const int c = color[x * n + y];
    if (c == 0) Paint(s, kAnsiRed, std::string(1, ch));
    else if (c == 1) Paint(s, kAnsiBlue, std::string(1, ch));
    else s->push_back(ch);
  }
  void Slot(std::string* s, int x, int y, const SlotSpec& spec) const {
    const size_t before = s->size();
    for (int i = 0; i < spec.n_always; ++i)
      LinkGlyph(s, x + spec.always[i].dx, y + spec.always[i].dy, spec.always[i].dir, spec.always[i].ch);
    if (spec.has_fallback && s->size() == before)
      LinkGlyph(s, x + spec.fallback.dx, y + spec.fallback.dy, spec.fallback.dir, spec.fallback.ch);
    if (s->size() == before) s->push_back(' ');
  }
  void Peg(std::string* s, int x, int y) const {
    const int c = color[x * n + y];
    if (c == 0) Paint(s, kAnsiRed, "x");
    else if (c == 1) Paint(s, kAnsiBlue, "o");
    else if (OffBoard(x, y)) s->push_back(' ');
    else if (x == 0 || x == n - 1) Paint(s, kAnsiBlue, ".");
    else if (y == 0 || y == n - 1) Paint(s, kAnsiRed, ".");
    else s->push_back('.');
  }
};

}  // namespace

std::string RenderRecord(const uint32_t* record, int n, bool ansi) {
  Picture p{n, ansi, std::vector<int>(static_cast<size_t>(n * n), 2), std::vector<int>(static_cast<size_t>(n * n), 0)};
  const uint32_t* planes = record + TWIXT_HEADER_WORDS;
  for (int x = 0; x < n; ++x)
    for (int y = 0; y < n; ++y) {
      if ((planes[0 * n + x] >> y) & 1u) p.color[x * n + y] = 0;
      else if ((planes[1 * n + x] >> y) & 1u) p.color[x * n + y] = 1;
      for (int d = 0; d < 4; ++d)
        if ((planes[(2 + d) * n + x] >> y) & 1u) {  // stored at the west endpoint; mirror to the east one
          p.links[x * n + y] |= 1 << d;
          p.links[(x + kDx[d]) * n + (y + kDy[d])] |= 1 << (d + 4);
        }
    }
  std::string s = "     ";
  for (int x = 0; x < n; ++x) p.Paint(&s, kAnsiRed, std::string(1, static_cast<char>('a' + x)) + "  ");
  s.push_back('\n');
  for (int y = n - 1; y >= 0; --y) {
    s.append("    ");
    for (int x = 0; x < n; ++x)
      for (const SlotSpec& spec : kBefore) p.Slot(&s, x, y, spec);
    s.push_back('\n');
    s.append(n - y < 10 ? "  " : " ");
    p.Paint(&s, kAnsiBlue, std::to_string(n - y) + " ");
    for (int x = 0; x < n; ++x) {
      p.Slot(&s, x, y, kPegLeft);
      p.Peg(&s, x, y);
      p.Slot(&s, x, y, kPegRight);
    }
    s.push_back('\n');
    s.append("    ");
    for (int x = 0; x < n; ++x)
      for (const SlotSpec& spec : kAfter) p.Slot(&s, x, y, spec);
    s.push_back('\n');
  }
  s.push_back('\n');
  if ((record[1] >> 2) & 1u) s.append("[swapped]");
  switch (record[1] & 3u) {
    case 1: s.append("[x has won]"); break;
    case 2: s.append("[o has won]"); break;
    case 3: s.append("[draw]"); break;
    default: break;
  }
  return s;
}

}  // namespace twixt_b200
}  // namespace open_spiel
