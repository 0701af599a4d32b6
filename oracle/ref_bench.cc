// TEST/BENCH INFRASTRUCTURE ONLY (oracle/): times the UNMODIFIED reference's
// CPU random-playout path (the loop upstream example.cc runs: LegalActions ->
// uniform pick -> ApplyAction until IsTerminal) on all
// host cores.  One PROCESS per worker, because the reference's static
// BlockerMap (twixtboard.h:142-151) makes threaded NewInitialState a data
// race.
//
//   ref_bench <board_size> <workers> <seconds> <mode> [seed]
//     mode "faithful": NewInitialState() per game (what the reference does;
//                      the Board ctor rebuilds the crossing map every time)
//     mode "clone":    one prebuilt initial state per worker, Clone() per game
//                      (steel-man: removes the constructor cost)
//     mode "rollout":  BASELINE config C3 (mcts_example --rollout_count=4 shape): leaves at a random ply in
//                      [0, 60]; per leaf 4 x { Clone(); random moves to the end; Returns() } as upstream's
//                      RandomRolloutEvaluator::Evaluate does; only the evaluations are timed -> leaves/s
//     mode "latency":  one process; mean ns per call of LegalActions / ApplyAction / ObservationTensor /
//                      Clone / IsTerminal+Returns over random games (what an unbatched caller pays per call)
// Prints one JSON line: aggregate plies / longest worker wall time.
#include <sys/wait.h>
#include <unistd.h>

#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <random>
#include <string>
#include <vector>

#include "open_spiel/spiel.h"
#include "open_spiel/games/twixt/twixt.h"

namespace {

struct WorkerResult {
  int64_t plies;
  int64_t games;
  int64_t red;
  int64_t blue;
  int64_t draws;
  double seconds;
};

WorkerResult RunWorker(int n, double budget_s, bool faithful, uint64_t seed) {
  using Clock = std::chrono::steady_clock;
  open_spiel::GameParameters params;
  params.insert({"board_size", open_spiel::GameParameter(n, false)});
  params.insert({"ansi_color_output", open_spiel::GameParameter(false, false)});
  std::shared_ptr<const open_spiel::Game> game(new open_spiel::twixt::TwixTGame(params));
  std::unique_ptr<open_spiel::State> proto = game->NewInitialState();
  std::mt19937 rng(static_cast<uint32_t>(seed));
  WorkerResult r{0, 0, 0, 0, 0, 0.0};
  auto t0 = Clock::now();
  for (;;) {
    std::unique_ptr<open_spiel::State> st = faithful ? game->NewInitialState() : proto->Clone();
    while (!st->IsTerminal()) {
      std::vector<open_spiel::Action> legal = st->LegalActions();
      std::uniform_int_distribution<size_t> pick(0, legal.size() - 1);
      st->ApplyAction(legal[pick(rng)]);
      ++r.plies;
    }
    ++r.games;
    std::vector<double> ret = st->Returns();
    if (ret[0] > 0) ++r.red; else if (ret[1] > 0) ++r.blue; else ++r.draws;
    r.seconds = std::chrono::duration<double>(Clock::now() - t0).count();
    if (r.seconds >= budget_s) break;
  }
  return r;
}

// config C3: leaves/s of the reference's rollout evaluation (steel-man: the leaf itself is made by Clone)
WorkerResult RunRolloutWorker(int n, double budget_s, uint64_t seed) {
  using Clock = std::chrono::steady_clock;
  open_spiel::GameParameters params;
  params.insert({"board_size", open_spiel::GameParameter(n, false)});
  params.insert({"ansi_color_output", open_spiel::GameParameter(false, false)});
  std::shared_ptr<const open_spiel::Game> game(new open_spiel::twixt::TwixTGame(params));
  std::unique_ptr<open_spiel::State> proto = game->NewInitialState();
  std::mt19937 rng(static_cast<uint32_t>(seed));
  WorkerResult r{0, 0, 0, 0, 0, 0.0};
  double timed = 0.0;
  auto random_move = [&](open_spiel::State* st) {
    std::vector<open_spiel::Action> legal = st->LegalActions();
    std::uniform_int_distribution<size_t> pick(0, legal.size() - 1);
    st->ApplyAction(legal[pick(rng)]);
  };
  while (timed < budget_s) {
    std::unique_ptr<open_spiel::State> leaf = proto->Clone();
    const int depth = static_cast<int>(rng() % 61u);
    for (int d = 0; d < depth && !leaf->IsTerminal(); ++d) random_move(leaf.get());
    auto t0 = Clock::now();
    double sum = 0.0;
    for (int k = 0; k < 4; ++k) {
      std::unique_ptr<open_spiel::State> st = leaf->Clone();
      while (!st->IsTerminal()) {
        random_move(st.get());
        ++r.plies;
      }
      sum += st->Returns()[0];
    }
    timed += std::chrono::duration<double>(Clock::now() - t0).count();
    r.red += sum > 0;
    ++r.games;  // leaves evaluated
  }
  r.seconds = timed;
  return r;
}

int RunLatency(int n, double budget_s, uint64_t seed) {
  using Clock = std::chrono::steady_clock;
  open_spiel::GameParameters params;
  params.insert({"board_size", open_spiel::GameParameter(n, false)});
  params.insert({"ansi_color_output", open_spiel::GameParameter(false, false)});
  std::shared_ptr<const open_spiel::Game> game(new open_spiel::twixt::TwixTGame(params));
  std::unique_ptr<open_spiel::State> proto = game->NewInitialState();
  std::mt19937 rng(static_cast<uint32_t>(seed));
  std::vector<float> obs(static_cast<size_t>(12 * n * (n - 2)));
  double t_legal = 0, t_apply = 0, t_obs = 0, t_clone = 0, t_query = 0;
  int64_t calls = 0, clones = 0;
  size_t sink = 0;
  auto since = [](Clock::time_point t0) { return std::chrono::duration<double>(Clock::now() - t0).count(); };
  auto start = Clock::now();
  while (since(start) < budget_s) {
    auto t0 = Clock::now();
    std::unique_ptr<open_spiel::State> st = proto->Clone();
    t_clone += since(t0);
    ++clones;
    for (;;) {
      t0 = Clock::now();
      const bool term = st->IsTerminal();
      std::vector<double> ret = st->Returns();
      t_query += since(t0);
      sink += ret.size();
      if (term) break;
      t0 = Clock::now();
      std::vector<open_spiel::Action> legal = st->LegalActions();
      t_legal += since(t0);
      t0 = Clock::now();
      st->ObservationTensor(0, absl::Span<float>(obs.data(), obs.size()));
      t_obs += since(t0);
      std::uniform_int_distribution<size_t> pick(0, legal.size() - 1);
      const open_spiel::Action a = legal[pick(rng)];
      t0 = Clock::now();
      st->ApplyAction(a);
      t_apply += since(t0);
      ++calls;
    }
  }
  std::printf(
      "{\"board_size\": %d, \"mode\": \"latency\", \"calls\": %lld, \"legal_actions_ns\": %.1f, \"apply_action_ns\": %.1f, "
      "\"observation_tensor_ns\": %.1f, \"clone_ns\": %.1f, \"is_terminal_returns_ns\": %.1f, \"sink\": %zu}\n",
      n, static_cast<long long>(calls), 1e9 * t_legal / calls, 1e9 * t_apply / calls, 1e9 * t_obs / calls,
      1e9 * t_clone / clones, 1e9 * t_query / (calls + clones), sink);
  return 0;
}

}  // namespace

int main(int argc, char** argv) {
  if (argc < 5) {
    std::fprintf(stderr, "usage: %s <board_size> <workers> <seconds> <faithful|clone> [seed]\n", argv[0]);
    return 2;
  }
  int n = std::atoi(argv[1]);
  int workers = std::atoi(argv[2]);
  double seconds = std::atof(argv[3]);
  bool faithful = std::strcmp(argv[4], "faithful") == 0;
  const bool rollout = std::strcmp(argv[4], "rollout") == 0;
  uint64_t seed = argc > 5 ? std::strtoull(argv[5], nullptr, 0) : 1;
  if (workers < 1) workers = 1;
  if (std::strcmp(argv[4], "latency") == 0) return RunLatency(n, seconds, seed);

  std::vector<int> fds(workers);
  std::vector<pid_t> pids(workers);
  for (int w = 0; w < workers; ++w) {
    int p[2];
    if (pipe(p) != 0) { std::perror("pipe"); return 1; }
    pid_t pid = fork();
    if (pid < 0) { std::perror("fork"); return 1; }
    if (pid == 0) {
      close(p[0]);
      const uint64_t wseed = seed * 1000003ull + static_cast<uint64_t>(w);
      WorkerResult r = rollout ? RunRolloutWorker(n, seconds, wseed) : RunWorker(n, seconds, faithful, wseed);
      ssize_t ignored = write(p[1], &r, sizeof(r));
      (void)ignored;
      close(p[1]);
      _exit(0);
    }
    close(p[1]);
    fds[w] = p[0];
    pids[w] = pid;
  }
  WorkerResult total{0, 0, 0, 0, 0, 0.0};
  for (int w = 0; w < workers; ++w) {
    WorkerResult r{};
    ssize_t got = read(fds[w], &r, sizeof(r));
    close(fds[w]);
    int status = 0;
    waitpid(pids[w], &status, 0);
    if (got != static_cast<ssize_t>(sizeof(r))) { std::fprintf(stderr, "worker %d failed\n", w); return 1; }
    total.plies += r.plies;
    total.games += r.games;
    total.red += r.red;
    total.blue += r.blue;
    total.draws += r.draws;
    if (r.seconds > total.seconds) total.seconds = r.seconds;
  }
  if (rollout) {
    std::printf("{\"board_size\": %d, \"workers\": %d, \"mode\": \"rollout\", \"rollouts_per_leaf\": 4, \"leaves\": %lld, "
                "\"plies\": %lld, \"seconds\": %.6f, \"leaves_per_sec\": %.1f, \"steps_per_sec\": %.1f}\n",
                n, workers, static_cast<long long>(total.games), static_cast<long long>(total.plies), total.seconds,
                total.seconds > 0 ? static_cast<double>(total.games) / total.seconds : 0.0,
                total.seconds > 0 ? static_cast<double>(total.plies) / total.seconds : 0.0);
    return 0;
  }
  std::printf(
      "{\"board_size\": %d, \"workers\": %d, \"mode\": \"%s\", \"plies\": %lld, \"games\": %lld, "
      "\"red\": %lld, \"blue\": %lld, \"draws\": %lld, \"seconds\": %.6f, \"steps_per_sec\": %.1f}\n",
      n, workers, faithful ? "faithful" : "clone", static_cast<long long>(total.plies),
      static_cast<long long>(total.games), static_cast<long long>(total.red),
      static_cast<long long>(total.blue), static_cast<long long>(total.draws), total.seconds,
      total.seconds > 0 ? static_cast<double>(total.plies) / total.seconds : 0.0);
  return 0;
}
