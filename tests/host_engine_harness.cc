// TEST-ONLY: compiles the product's per-env rules header
// (twixt_for_open_spiel_b200/csrc/twixt_engine.cuh) for the HOST so that
// `pytest -m "not gpu"` can check the rules, the crossing masks, the legal
// selection and the Philox stream against the oracle without a GPU.  This
// file is never part of libtwixt_b200.so: the product has no CPU path.
// TW_TEST_STACK (default 48, tests also build with 2) sizes the flood stack
// so the overflow branch is exercised.
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cstring>

#include "twixt_engine.cuh"
#include "twixt_philox.cuh"

#ifndef TW_TEST_STACK
#define TW_TEST_STACK 48
#endif

using namespace twixt;
using Rec = RecordRef<1>;

// Test-only accessor with the per-column count cache the fused playout kernel keeps in shared memory.
struct CachedRec : public RecordRef<1> {
  uint32_t cache[6];
  static constexpr bool kCountCache = true;
  uint32_t cache_ld(int i) const { return cache[i]; }
  void cache_st(int i, uint32_t v) { cache[i] = v; }
  void note_peg(int x, int y, int delta) {
    const uint32_t inc = (1u | ((y == 0 || y == n() - 1) ? 32u : 0u)) << (8 * (x & 3));
    cache[x >> 2] = delta > 0 ? cache[x >> 2] + inc : cache[x >> 2] - inc;
  }
};

// An accessor with the fused kernel's shared-memory layout (twixt_kernel_playout.cu, PlayoutRef): planes in
// the order BLUE, RED, links, START, END followed by the flood-stack words, and link-window / flag reads
// WITHOUT bounds tests.  The stack words are poisoned so that a read off the board that mattered would show.
// select_bit_lut's table, as the kernel builds it in shared memory
static const uint8_t* host_select_lut() {
  static uint8_t lut[8 * 256];
  static bool ready = false;
  if (!ready) {
    for (int e = 0; e < 8 * 256; ++e) fill_select_lut(lut, e);
    ready = true;
  }
  return lut;
}

struct SmemRec {
  static constexpr bool kSelectLut = true;
  const uint8_t* select_lut() const { return host_select_lut(); }
  std::vector<uint32_t> w;
  uint32_t* blocked;  // the record's blocked plane (global memory in the kernel)
  int n_rt;
  uint32_t cache[6];
  static constexpr bool kCountCache = true;
  int n() const { return n_rt; }
  int smem_word(int i) const { return i < n_rt ? i + n_rt : (i < 2 * n_rt ? i - n_rt : i); }
  void take(const uint32_t* rec, int n) {
    n_rt = n;
    w.assign(8 * n + 25, 0xDEADBEEFu);
    for (int i = 0; i < 8 * n; ++i) w[smem_word(i)] = rec[kHeaderWords + i];
  }
  void give(uint32_t* rec) const {
    for (int i = 0; i < 8 * n_rt; ++i) rec[kHeaderWords + i] = w[smem_word(i)];
  }
  uint32_t ld(int plane, int col) const { return w.at(plane * n_rt + col); }
  void st(int plane, int col, uint32_t v) { w.at(plane * n_rt + col) = v; }
  void st_if(bool c, int plane, int col, uint32_t v) {
    if (c) {
      if (col < 0 || col >= n_rt) std::abort();
      st(plane, col, v);
    }
  }
  uint32_t ld_guard(int plane, int col) const {
    return (static_cast<unsigned>(col) < static_cast<unsigned>(n_rt)) ? ld(plane, col) : 0u;
  }
  uint32_t ld_link(int plane, int col) const { return ld(plane, col); }
  uint32_t ld_any(int plane, int col) const { return ld(plane, col); }
  uint32_t ld_pegs(int plane, int col) const { return ld(plane ^ 1, col); }
  void st_pegs(int plane, int col, uint32_t v) { st(plane ^ 1, col, v); }
  uint32_t ld_pegs_guard(int plane, int col) const { return ld_guard(plane ^ 1, col); }
  void or_blocked3(int x, const uint32_t blk[3]) {
    blocked[x] |= blk[0];
    blocked[x - 1 > 0 ? x - 1 : 0] |= blk[1];
    blocked[x - 2 > 0 ? x - 2 : 0] |= blk[2];
  }
  uint32_t cache_ld(int i) const { return cache[i]; }
  void cache_st(int i, uint32_t v) { cache[i] = v; }
  void note_peg(int x, int y, int delta) {
    const uint32_t inc = (1u | ((y == 0 || y == n() - 1) ? 32u : 0u)) << (8 * (x & 3));
    cache[x >> 2] = delta > 0 ? cache[x >> 2] + inc : cache[x >> 2] - inc;
  }
};

// The same playout with the structure of the CUDA kernel: per-column count cache, moves interleaved with
// single flood visits (a visit takes a whole stack entry; a new flood starts from the peg itself; the
// opponent keeps moving while a colour's flood is still running), two
// buffered Philox blocks refreshed on a fixed 4-iteration schedule, the selection of move i+1 issued
// (speculatively) between the placement and the link evaluation of move i, and the swap's first half run
// ahead of the move in the rare-events step.
template <class B>
int playout_like_kernel(B& b, Header& h, int n, uint64_t seed, uint64_t stream, int max_plies, int64_t* actions_out) {
  count_cache_build(b);
  const uint32_t s_lo = static_cast<uint32_t>(stream), s_hi = static_cast<uint32_t>(stream >> 32);
  const uint32_t k_lo = static_cast<uint32_t>(seed), k_hi = static_cast<uint32_t>(seed >> 32);
  int step = 0;
  uint32_t ra[4], rb[4], rq = 0;
  philox4x32_10(s_lo, s_hi, 0u, 0u, k_lo, k_hi, ra);
  philox4x32_10(s_lo, s_hi, 1u, 0u, k_lo, k_hi, rb);
  auto word_at = [&](uint32_t index) {
    const uint32_t rel = index - 4u * rq;
    return rel < 4u ? ra[rel] : rb[rel - 4u];
  };
  uint32_t pend = 0, origin_r = 0, origin_b = 0, fcol = 0;  // per-colour flood bookkeeping, as in the kernel
  int fplane = P_START;
  LocalStack<TW_TEST_STACK> stk;
  bool playing = h.result == kOpen && max_plies > 0;
  int sx = 0, sy = 0;
  if (playing) select_legal(b, h, static_cast<int>(playout_index(word_at(0), static_cast<uint32_t>(legal_count(h, n)))), sx, sy);
  int sact = sx * n + sy;
  bool swap_next = playing && is_swap(h, static_cast<uint32_t>(sact));
  for (uint32_t it = 1; playing || pend != 0u || !stk.empty(); ++it) {
    if ((it & 3u) == 0u && static_cast<uint32_t>(step) + 1u >= 4u * (rq + 1u)) {
      rq += 1u;
      for (int i = 0; i < 4; ++i) ra[i] = rb[i];
      philox4x32_10(s_lo, s_hi, rq + 1u, 0u, k_lo, k_hi, rb);
    }
    if (swap_next) {
      swap_first_move(b, h, sx, sy);
      swap_next = false;
    }
    const uint32_t mover = h.ply & 1u;
    const bool flood_blocks = ((pend >> (2u * mover)) & 3u) != 0u || (!stk.empty() && fcol == mover);
    if (playing && !flood_blocks) {
      if (actions_out) actions_out[step] = sact;
      const Placement pl = begin_move<true>(b, h, sx, sy);
      Header hn = h;  // the position the following move is chosen in, if this move does not end the game
      hn.ply = h.ply + 1u;
      const int ln = legal_count(hn, n);
      int nx, ny;
      select_legal(b, hn, static_cast<int>(playout_index(word_at(static_cast<uint32_t>(step) + 1u), static_cast<uint32_t>(ln))), nx, ny);
      uint32_t owed;
      const bool win = link_move<true>(b, pl, owed);
      finish_move(h, pl, win);
      pend |= owed << (2 * pl.player);
      (pl.player == kRed ? origin_r : origin_b) = flood_entry(pl.x, 1u << pl.y);
      ++step;
      playing = h.result == kOpen && step < max_plies;
      sx = nx;
      sy = ny;
      sact = nx * n + ny;
      swap_next = playing && is_swap(h, static_cast<uint32_t>(sact));
    }
    if (!stk.empty() || pend != 0u) {
      const bool begin = stk.empty();
      const uint32_t lsb = pend & (0u - pend);  // the lowest owed flood first, as in the kernel
      const bool blue = (lsb & 12u) != 0u;
      fplane = begin ? ((lsb & 5u) != 0u ? P_START : P_END) : fplane;
      fcol = begin ? (blue ? 1u : 0u) : fcol;
      pend ^= begin ? lsb : 0u;
      const uint32_t e = stk.top_or(blue ? origin_b : origin_r);
      flood_visit_entry(b, fplane, stk, e);
      if (stk.empty() && stk.overflow) {
        flood_closure(b, fcol == kRed ? P_RED : P_BLUE, fplane);
        stk.overflow = false;
      }
    }
  }
  return step;
}

extern "C" {

int he_record_words(int n) { return record_words(n); }

void he_init(uint32_t* rec, int n) {
  Rec b{rec, n};
  init_record(b);
}

int he_legal_actions(uint32_t* rec, int n, int64_t* out) {
  Rec b{rec, n};
  Header h;
  load_header(b, h);
  if (h.result != kOpen) return 0;
  int c = 0;
  for (int x = 0; x < n; ++x) {
    uint32_t w = legal_word(b, h, x);
    while (w) {
      int y = tw_ctz(w);
      w &= w - 1u;
      out[c++] = x * n + y;
    }
  }
  return c;
}

int he_legal_count(uint32_t* rec, int n) {
  Rec b{rec, n};
  Header h;
  load_header(b, h);
  return legal_count(h, n);
}

// 0 applied, 1 illegal (record untouched)
int he_apply(uint32_t* rec, int n, int action) {
  Rec b{rec, n};
  Header h;
  load_header(b, h);
  if (!is_legal(b, h, action)) return 1;
  apply_legal_cell<TW_TEST_STACK>(b, h, action / n, action % n);
  store_header(b, h);
  return 0;
}

int he_current_player(uint32_t* rec, int n) {
  Rec b{rec, n};
  Header h;
  load_header(b, h);
  return current_player(h);
}

void he_observation(uint32_t* rec, int n, float* out) {
  Rec b{rec, n};
  int w = n - 2;
  for (int p = 0; p < 12; ++p)
    for (int r = 0; r < n; ++r)
      for (int c = 0; c < w; ++c) {
        int x, y;
        obs_cell(n, p, r, c, x, y);
        out[(p * n + r) * w + c] = ((obs_plane_word(b, p, x) >> y) & 1u) ? 1.0f : 0.0f;
      }
}

int he_playout(uint32_t* rec, int n, uint64_t seed, uint64_t stream, int max_plies, int64_t* actions_out) {
  Rec b{rec, n};
  Header h;
  load_header(b, h);
  int step = 0;
  uint32_t r[4] = {0, 0, 0, 0};
  while (h.result == kOpen && step < max_plies) {
    if ((step & 3) == 0)
      philox4x32_10(static_cast<uint32_t>(stream), static_cast<uint32_t>(stream >> 32),
                    static_cast<uint32_t>(step) >> 2, 0u, static_cast<uint32_t>(seed),
                    static_cast<uint32_t>(seed >> 32), r);
    int L = legal_count(h, n);
    int k = static_cast<int>(playout_index(r[step & 3], static_cast<uint32_t>(L)));
    int x, y;
    select_legal(b, h, k, x, y);
    if (actions_out) actions_out[step] = x * n + y;
    apply_legal_cell<TW_TEST_STACK>(b, h, x, y);
    ++step;
  }
  store_header(b, h);
  return step;
}

int he_playout_cached(uint32_t* rec, int n, uint64_t seed, uint64_t stream, int max_plies, int64_t* actions_out) {
  CachedRec b;
  b.p = rec;
  b.n_rt = n;
  Header h;
  load_header(b, h);
  const int steps = playout_like_kernel(b, h, n, seed, stream, max_plies, actions_out);
  store_header(b, h);
  return steps;
}

// ... on the kernel's shared-memory layout with its unguarded reads
int he_playout_smem(uint32_t* rec, int n, uint64_t seed, uint64_t stream, int max_plies, int64_t* actions_out) {
  Rec r{rec, n};
  Header h;
  load_header(r, h);
  SmemRec b;
  b.take(rec, n);
  b.blocked = rec + kHeaderWords + P_BLOCKED * n;
  const int steps = playout_like_kernel(b, h, n, seed, stream, max_plies, actions_out);
  b.give(rec);
  store_header(r, h);
  return steps;
}

void he_philox(const uint32_t* ctr, const uint32_t* key, uint32_t* out) {
  philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1], out);
}

int he_select_bit(uint32_t w, int k) { return select_bit(w, k); }

}  // extern "C"
