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
  uint64_t seed = argc > 5 ? std::strtoull(argv[5], nullptr, 0) : 1;
  if (workers < 1) workers = 1;

  std::vector<int> fds(workers);
  std::vector<pid_t> pids(workers);
  for (int w = 0; w < workers; ++w) {
    int p[2];
    if (pipe(p) != 0) { std::perror("pipe"); return 1; }
    pid_t pid = fork();
    if (pid < 0) { std::perror("fork"); return 1; }
    if (pid == 0) {
      close(p[0]);
      WorkerResult r = RunWorker(n, seconds, faithful, seed * 1000003ull + static_cast<uint64_t>(w));
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
  std::printf(
      "{\"board_size\": %d, \"workers\": %d, \"mode\": \"%s\", \"plies\": %lld, \"games\": %lld, "
      "\"red\": %lld, \"blue\": %lld, \"draws\": %lld, \"seconds\": %.6f, \"steps_per_sec\": %.1f}\n",
      n, workers, faithful ? "faithful" : "clone", static_cast<long long>(total.plies),
      static_cast<long long>(total.games), static_cast<long long>(total.red),
      static_cast<long long>(total.blue), static_cast<long long>(total.draws), total.seconds,
      total.seconds > 0 ? static_cast<double>(total.plies) / total.seconds : 0.0);
  return 0;
}
