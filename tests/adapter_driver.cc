// TEST-ONLY driver for the C++ open_spiel adapter (twixt_for_open_spiel_b200/adapter): re-hosts the
// assertions of the reference's twixt_test.cc:50-199 on TwixTB200Game / TwixTB200State, compiled against
// the open_spiel header shim.  `adapter_driver cpu` runs what needs no GPU (parameter errors, renderer);
// `adapter_driver gpu` runs the state tests on cuda:0.
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <iostream>
#include <string>
#include <thread>
#include <vector>

#include "twixt_b200_game.h"

using open_spiel::Action;
using open_spiel::GameParameter;
using open_spiel::GameParameters;
using open_spiel::twixt_b200::TwixTB200Game;

#define EXPECT(cond)                                                          \
  do {                                                                        \
    if (!(cond)) {                                                            \
      std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); \
      return 1;                                                               \
    }                                                                         \
  } while (0)

static std::shared_ptr<const open_spiel::Game> Load(int board_size) {
  GameParameters params;
  if (board_size > 0) params.insert({"board_size", GameParameter(board_size, false)});
  return std::shared_ptr<const open_spiel::Game>(new TwixTB200Game(params));
}

static bool Has(const std::vector<Action>& v, Action a) { return std::find(v.begin(), v.end(), a) != v.end(); }

static int ParameterTest() {  // twixt_test.cc:50-92
  Load(10);
  for (int bad : {30, 3}) {
    try {
      Load(bad);
      EXPECT(false);
    } catch (const open_spiel::SpielError& e) {
      EXPECT(std::string(e.what()) == "board_size out of range [5..24]: " + std::to_string(bad));
    }
  }
  try {
    GameParameters params;
    params.insert({"bad_param", GameParameter(3, false)});
    TwixTB200Game g(params);
    EXPECT(false);
  } catch (const open_spiel::SpielError& e) {
    EXPECT(std::string(e.what()) ==
           "Unknown parameter 'bad_param'. Available parameters are: ansi_color_output, board_size");
  }
  auto g = Load(12);
  EXPECT(g->NumDistinctActions() == 144);
  EXPECT((g->ObservationTensorShape() == std::vector<int>{12, 12, 10}));
  auto g5 = Load(5);  // no stale function-static shape (twixt.h:131-134)
  EXPECT((g5->ObservationTensorShape() == std::vector<int>{12, 5, 3}));
  EXPECT(static_cast<const TwixTB200Game&>(*g).MaxGameLength() == 141);
  return 0;
}

static int RenderTest(const char* expected_path) {
  // the initial 8x8 record rendered on the host equals the playthrough's State 0 string
  std::vector<uint32_t> rec(76, 0u);
  rec[2] = 0xFFFFFFFFu;
  rec[3] = 48u | (48u << 16);
  std::string got = open_spiel::twixt_b200::RenderRecord(rec.data(), 8, true);
  FILE* f = std::fopen(expected_path, "rb");
  EXPECT(f != nullptr);
  std::string want;
  char buf[4096];
  size_t k;
  while ((k = std::fread(buf, 1, sizeof(buf), f)) > 0) want.append(buf, k);
  std::fclose(f);
  EXPECT(got == want);
  return 0;
}

static int SwapTest() {  // twixt_test.cc:108-131
  auto game = Load(0);
  auto state = game->NewInitialState();
  EXPECT(state->CurrentPlayer() == 0);
  EXPECT(Has(state->LegalActions(), 11));
  state->ApplyAction(19);
  EXPECT(state->CurrentPlayer() == 1);
  state->ApplyAction(19);
  EXPECT(Has(state->LegalActions(), 19));
  EXPECT(!Has(state->LegalActions(), 29));
  EXPECT(state->CurrentPlayer() == 0);
  state->ApplyAction(36);
  EXPECT(Has(state->LegalActions(), 19));
  EXPECT(!Has(state->LegalActions(), 29));
  EXPECT(!Has(state->LegalActions(), 36));
  EXPECT(state->ToString().find("[swapped]") != std::string::npos);
  return 0;
}

static int LegalActionsTest() {  // twixt_test.cc:133-183
  auto game = Load(0);
  auto state = game->NewInitialState();
  EXPECT(!state->IsTerminal());
  EXPECT(state->LegalActions().size() == 48);
  const int moves[] = {21, 38, 15, 11};
  const size_t sizes[] = {48, 46, 46, 44};
  for (int i = 0; i < 4; ++i) {
    state->ApplyAction(moves[i]);
    EXPECT(state->LegalActions().size() == sizes[i]);
  }
  try {
    state->ApplyAction(11);
    EXPECT(false);
  } catch (const open_spiel::SpielError& e) {
    EXPECT(std::string(e.what()) == "Not a legal action: 11");
  }
  auto clone = state->Clone();
  const int moves2[] = {27, 17, 42, 45};
  const size_t sizes2[] = {44, 42, 42, 40};
  for (int i = 0; i < 4; ++i) {
    state->ApplyAction(moves2[i]);
    EXPECT(state->LegalActions().size() == sizes2[i]);
  }
  state->ApplyAction(48);
  EXPECT(state->IsTerminal());
  EXPECT(state->PlayerReturn(0) == 1.0);
  EXPECT(state->PlayerReturn(1) == -1.0);
  EXPECT(state->CurrentPlayer() == open_spiel::kTerminalPlayerId);
  EXPECT(state->LegalActions().empty());
  EXPECT(state->ToString().find("[x has won]") != std::string::npos);
  EXPECT(!clone->IsTerminal() && clone->LegalActions().size() == 44);  // the clone did not move
  EXPECT(state->ActionToString(0, 19) == "xc5" && state->ActionToString(1, 43) == "of5");  // twixtboard.h:166-167
  std::vector<float> obs(576);
  state->ObservationTensor(0, absl::Span<float>(obs.data(), obs.size()));
  // the oracle's tensor for this position: exactly these seven 1.0s
  const int want_ones[] = {170, 192, 205, 226, 301, 306, 519};
  size_t k = 0;
  for (size_t i = 0; i < obs.size(); ++i) {
    if (obs[i] == 1.0f) {
      EXPECT(k < 7 && static_cast<int>(i) == want_ones[k]);
      ++k;
    } else {
      EXPECT(obs[i] == 0.0f);
    }
  }
  EXPECT(k == 7);
  return 0;
}

static int DrawTest() {  // twixt_test.cc:185-199
  auto game = Load(5);
  auto state = game->NewInitialState();
  while (!state->IsTerminal()) {
    state->ApplyAction(state->LegalActions().at(0));
    state->ApplyAction(state->LegalActions().at(1));
  }
  EXPECT(state->PlayerReturn(0) == 0.0 && state->PlayerReturn(1) == 0.0);
  EXPECT(state->History().size() == 18);
  EXPECT(state->ToString().find("[draw]") != std::string::npos);
  return 0;
}

// States of ONE shared game used from several threads at once (what AlphaZero actors / parallel MCTS do):
// every thread replays the same fixed games on its own States and must see exactly the legal lists, tensors
// and strings the single-threaded run saw.  All these States share one device batch -- the adapter serialises
// the calls into it (EnvPool, twixt_b200_game.h).
static int ThreadsTest() {
  auto game = Load(0);
  const std::vector<std::vector<Action>> games = {{21, 38, 15, 11, 27, 17, 42, 45, 48}, {19, 19, 36, 21, 50, 13, 9}};
  struct Seen {
    std::vector<std::vector<Action>> legal;
    std::vector<std::vector<float>> obs;
    std::string last;
  };
  auto play = [&](const std::vector<Action>& acts) {
    Seen s;
    auto state = game->NewInitialState();
    for (Action a : acts) {
      s.legal.push_back(state->LegalActions());
      std::vector<float> o(576);
      state->ObservationTensor(0, absl::Span<float>(o.data(), o.size()));
      s.obs.push_back(o);
      auto copy = state->Clone();  // clones are made and dropped concurrently too
      copy->ApplyAction(a);
      state->ApplyAction(a);
      if (copy->ToString() != state->ToString()) s.last = "clone differs";
    }
    if (s.last.empty()) s.last = state->ToString();
    return s;
  };
  std::vector<Seen> want;
  for (const auto& g : games) want.push_back(play(g));
  const int kThreads = 8, kRounds = 6;
  std::vector<int> bad(kThreads, 0);
  std::vector<std::thread> threads;
  for (int t = 0; t < kThreads; ++t)
    threads.emplace_back([&, t]() {
      for (int r = 0; r < kRounds; ++r) {
        const size_t gi = static_cast<size_t>((t + r) % games.size());
        Seen got = play(games[gi]);
        if (got.legal != want[gi].legal || got.obs != want[gi].obs || got.last != want[gi].last) ++bad[t];
      }
    });
  for (auto& th : threads) th.join();
  for (int t = 0; t < kThreads; ++t) EXPECT(bad[t] == 0);
  return 0;
}

// CRC-32 (IEEE, as zlib.crc32) of a byte string
static uint32_t Crc32(const std::string& s) {
  uint32_t c = 0xFFFFFFFFu;
  for (unsigned char ch : s) {
    c ^= ch;
    for (int k = 0; k < 8; ++k) c = (c >> 1) ^ (0xEDB88320u & (0u - (c & 1u)));
  }
  return c ^ 0xFFFFFFFFu;
}

// ToString() of the adapter at EVERY ply of reference-generated games (tests/golden/ref_strings.json, written
// out by the test as plain text): the CRC of the picture at every ply, the full text at the plies the
// fixture keeps.  The games hold links of all eight directions in both colours, the swap, both wins and a
// draw, with and without ANSI colour codes -- so RenderRecord (and the record it is fed by the CUDA path)
// is compared with Board::ToString (twixtboard.cc:278-448) beyond the empty board.
static int StringsTest(const char* path) {
  FILE* f = std::fopen(path, "rb");
  EXPECT(f != nullptr);
  int games = 0;
  char tag[16];
  while (std::fscanf(f, "%15s", tag) == 1 && std::string(tag) == "GAME") {
    int n = 0, ansi = 0, num = 0;
    EXPECT(std::fscanf(f, "%d %d %d", &n, &ansi, &num) == 3);
    std::vector<long long> actions(static_cast<size_t>(num));
    for (auto& a : actions) EXPECT(std::fscanf(f, "%lld", &a) == 1);
    std::vector<uint32_t> crcs(static_cast<size_t>(num) + 1);
    for (auto& c : crcs) EXPECT(std::fscanf(f, "%u", &c) == 1);
    int kept = 0;
    EXPECT(std::fscanf(f, "%d", &kept) == 1);
    std::vector<std::pair<int, std::string>> texts;
    for (int k = 0; k < kept; ++k) {
      int ply = 0;
      long bytes = 0;
      EXPECT(std::fscanf(f, "%d %ld", &ply, &bytes) == 2);
      EXPECT(std::fgetc(f) == '\n');
      std::string t(static_cast<size_t>(bytes), '\0');
      EXPECT(std::fread(&t[0], 1, t.size(), f) == t.size());
      texts.emplace_back(ply, std::move(t));
    }
    GameParameters params;
    params.insert({"board_size", GameParameter(n, false)});
    params.insert({"ansi_color_output", GameParameter(ansi != 0, false)});
    std::shared_ptr<const open_spiel::Game> game(new TwixTB200Game(params));
    auto state = game->NewInitialState();
    size_t next_text = 0;
    for (int ply = 0; ply <= num; ++ply) {
      const std::string got = state->ToString();
      if (Crc32(got) != crcs[static_cast<size_t>(ply)]) {
        std::fprintf(stderr, "ToString differs: game %d n=%d ansi=%d ply %d\n", games, n, ansi, ply);
        return 1;
      }
      if (next_text < texts.size() && texts[next_text].first == ply) {
        EXPECT(got == texts[next_text].second);
        EXPECT(state->ObservationString(0) == got && state->InformationStateString(1) == got);  // twixt.h:65-75
        ++next_text;
      }
      if (ply < num) state->ApplyAction(actions[static_cast<size_t>(ply)]);
    }
    EXPECT(next_text == texts.size());
    ++games;
  }
  std::fclose(f);
  EXPECT(games >= 10);
  return 0;
}

// RenderRecord alone (no GPU): records written by the test from the oracle at plies of the reference-generated
// string fixture, each with the reference's picture of that position.
static int RenderRecordsTest(const char* path) {
  FILE* f = std::fopen(path, "rb");
  EXPECT(f != nullptr);
  int seen = 0;
  char tag[16];
  while (std::fscanf(f, "%15s", tag) == 1 && std::string(tag) == "REC") {
    int n = 0, ansi = 0, words = 0;
    long bytes = 0;
    EXPECT(std::fscanf(f, "%d %d %d %ld", &n, &ansi, &words, &bytes) == 4);
    std::vector<uint32_t> rec(static_cast<size_t>(words));
    for (auto& w : rec) EXPECT(std::fscanf(f, "%u", &w) == 1);
    EXPECT(std::fgetc(f) == '\n');
    std::string want(static_cast<size_t>(bytes), '\0');
    EXPECT(std::fread(&want[0], 1, want.size(), f) == want.size());
    if (open_spiel::twixt_b200::RenderRecord(rec.data(), n, ansi != 0) != want) {
      std::fprintf(stderr, "RenderRecord differs: record %d (n=%d ansi=%d)\n", seen, n, ansi);
      return 1;
    }
    ++seen;
  }
  std::fclose(f);
  EXPECT(seen >= 50);
  return 0;
}

int main(int argc, char** argv) {
  const std::string mode = argc > 1 ? argv[1] : "cpu";
  try {
    if (mode == "cpu") {
      if (ParameterTest() != 0) return 1;
      if (argc > 2 && RenderTest(argv[2]) != 0) return 1;
      if (argc > 3 && RenderRecordsTest(argv[3]) != 0) return 1;
    } else {
      if (SwapTest() != 0 || LegalActionsTest() != 0 || DrawTest() != 0 || ThreadsTest() != 0) return 1;
      if (argc > 2 && StringsTest(argv[2]) != 0) return 1;
    }
  } catch (const std::exception& e) {
    std::fprintf(stderr, "unexpected exception: %s\n", e.what());
    return 1;
  }
  std::puts("OK");
  return 0;
}
