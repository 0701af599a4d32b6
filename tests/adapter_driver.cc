// TEST-ONLY driver for the C++ open_spiel adapter (twixt_for_open_spiel_b200/adapter): re-hosts the
// assertions of the reference's twixt_test.cc:50-199 on TwixTB200Game / TwixTB200State, compiled against
// the open_spiel header shim.  `adapter_driver cpu` runs what needs no GPU (parameter errors, renderer);
// `adapter_driver gpu` runs the state tests on cuda:0.
#include <cstdio>
#include <cstring>
#include <iostream>
#include <string>
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

int main(int argc, char** argv) {
  const std::string mode = argc > 1 ? argv[1] : "cpu";
  try {
    if (mode == "cpu") {
      if (ParameterTest() != 0) return 1;
      if (argc > 2 && RenderTest(argv[2]) != 0) return 1;
    } else {
      if (SwapTest() != 0 || LegalActionsTest() != 0 || DrawTest() != 0) return 1;
    }
  } catch (const std::exception& e) {
    std::fprintf(stderr, "unexpected exception: %s\n", e.what());
    return 1;
  }
  std::puts("OK");
  return 0;
}
