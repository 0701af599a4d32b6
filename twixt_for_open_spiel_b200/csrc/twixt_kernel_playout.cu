// twixt_kernel_playout.cu -- K5, the fused random-playout kernel.
//
// One THREAD per env.  Each lane copies the bit-planes of its env from HBM into
// shared memory, transposed so that word w of lane l sits at smem[w*32 + l]:
// whatever word each lane indexes, lane l always hits bank l, so every access
// of the scalar rules in twixt_engine.cuh is conflict-free.  The whole game is
// then played out of shared memory -- select a uniformly random legal action
// (Philox4x32-10, one block per four moves), apply it (peg, links, crossing
// test, border flags, result) -- and the final planes are written back once.
// HBM sees 2 x record bytes per GAME, not per move; the limiter is issue slots
// and shared-memory latency, which is why the thread-per-env mapping is used:
// it spends ~1/10 of the warp-instructions per move of a warp-per-env mapping
// because nothing is computed redundantly across lanes ("Mapping
// one game onto the machine").
//
// Divergence control (profiles/r1_playout_v1_*: 9.5 of 32 lanes active):
//  * the border-flag flood (ExploreLocalGraph) is cut into single visits that
//    are interleaved with the other lanes' moves: every loop iteration runs
//    one MOVE for the lanes without flood work and one flood VISIT for the
//    lanes with it, instead of 31 lanes idling through one lane's whole flood;
//    the flood stack lives in shared memory next to the planes;
//  * candidate links are handled direction-major (place_peg).
// Shared memory per env: planes 0..7 (pegs, links, border flags) + the flood
// stack + a per-column count cache that replaces the 24-word legal scan by 6
// bytewise words (twixt_engine.cuh, count_cache_*).  The "blocked neighbour" plane is write-only for the rules, so it
// stays in HBM and is OR-ed in place on the rare blocked link.
//
// Reference loop reproduced: upstream example.cc / RandomRolloutEvaluator
// (LegalActions -> uniform pick -> ApplyAction until IsTerminal), with the
// per-move semantics of twixtboard.cc:457-499.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "twixt_engine.cuh"
#include "twixt_kernels.cuh"
#include "twixt_philox.cuh"

namespace twixt {

namespace {

constexpr int kPlayoutThreads = 128;  // 4 warps per block
constexpr int kCacheWords = 6;        // per-column count cache, four columns per word
constexpr unsigned kFullMask = 0xFFFFFFFFu;

// Two placements of the env state during a playout (template parameter SP =
// number of leading planes staged in shared memory):
//   SP = 8  all planes the rules read (pegs, links, border flags) on chip;
//           888 B per env at n=24 -> 256 envs = 8 warps per SM
//   SP = 2  only the two peg planes (touched by every move) on chip, links and
//           border flags are read/written in place in HBM/L2 (a move fetches its
//           whole 5-column link window with independent loads, one round trip);
//           264 B per env at n=24 -> 768 envs = 24 warps per SM to hide latency
template <int SP>
struct PlayoutCfg {
  static constexpr int kStackWords = SP >= 8 ? 24 : 12;  // flood stack entries (one per word)
  __host__ __device__ static constexpr int words(int n) { return SP * n + kStackWords + kCacheWords; }
};

// The env's planes: the first SP in shared memory (stride 32 words), the rest in its HBM record.
template <int NT, int SP>
struct PlayoutRef {
  uint32_t* p;     // smem, this lane's column
  uint32_t* gpl;   // global: the record's plane words (record + kHeaderWords)
  int n_rt;
  __device__ __forceinline__ int n() const { return NT > 0 ? NT : n_rt; }
  __device__ __forceinline__ uint32_t ld(int plane, int col) const {
    return plane < SP ? p[(plane * n() + col) * 32] : gpl[plane * n() + col];
  }
  __device__ __forceinline__ void st(int plane, int col, uint32_t v) {
    if (plane < SP) p[(plane * n() + col) * 32] = v;
    else gpl[plane * n() + col] = v;
  }
  __device__ __forceinline__ uint32_t ld_guard(int plane, int col) const {
    return (static_cast<unsigned>(col) < static_cast<unsigned>(n())) ? ld(plane, col) : 0u;
  }
  // peg planes are always on chip
  __device__ __forceinline__ uint32_t ld_pegs(int plane, int col) const { return p[(plane * n() + col) * 32]; }
  __device__ __forceinline__ void st_pegs(int plane, int col, uint32_t v) { p[(plane * n() + col) * 32] = v; }
  __device__ __forceinline__ uint32_t ld_pegs_guard(int plane, int col) const {
    return (static_cast<unsigned>(col) < static_cast<unsigned>(n())) ? ld_pegs(plane, col) : 0u;
  }
  // fire-and-forget reduction (RED.OR): no load to wait for; only this thread touches the word
  __device__ __forceinline__ void or_blocked(int col, uint32_t bits) { atomicOr(gpl + P_BLOCKED * n() + col, bits); }
  // per-column count cache (twixt_engine.cuh, count_cache_*), after the planes and the stack
  static constexpr bool kCountCache = true;
  __device__ __forceinline__ uint32_t* cache_word(int i) const {
    return p + (SP * n() + PlayoutCfg<SP>::kStackWords + i) * 32;
  }
  __device__ __forceinline__ uint32_t cache_ld(int i) const { return *cache_word(i); }
  __device__ __forceinline__ void cache_st(int i, uint32_t v) { *cache_word(i) = v; }
  __device__ __forceinline__ void note_peg(int x, int y, int delta) {
    const uint32_t inc = (1u | ((y == 0 || y == n() - 1) ? 32u : 0u)) << (8 * (x & 3));
    uint32_t* w = cache_word(x >> 2);
    *w = delta > 0 ? *w + inc : *w - inc;
  }
};

template <int CAP>
struct SmemStack {
  uint32_t* base;  // smem, this lane's column of the stack words
  int sp;
  bool overflow;
  __device__ __forceinline__ bool empty() const { return sp == 0; }
  __device__ __forceinline__ void push(uint32_t c) {
    if (sp < CAP) base[(sp++) * 32] = c;
    else overflow = true;
  }
  __device__ __forceinline__ uint32_t pop() { return base[(--sp) * 32]; }
};

template <int NT, int SP>
__global__ void __launch_bounds__(kPlayoutThreads) playout_kernel(const PlayoutArgs a) {
  extern __shared__ uint4 smem_raw[];
  uint32_t* smem = reinterpret_cast<uint32_t*>(smem_raw);
  const int n = NT > 0 ? NT : a.n;
  const int rw = record_words(n);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t idx = blockIdx.x * static_cast<int64_t>(kPlayoutThreads) + threadIdx.x;
  const bool active = idx < a.count;
  uint32_t* mine = smem + warp * (PlayoutCfg<SP>::words(n) * 32) + lane;
  uint32_t* grec = a.records + idx * rw;
  const int staged_pairs = (SP * n) / 2;  // staged plane words, as 8-byte pieces (planes start 16-byte aligned)

  Header h;
  h.ply = 0; h.result = kDraw; h.swapped = 0; h.move_one = kNoMove; h.cnt[0] = h.cnt[1] = 0;
  if (active) {
    const uint4 hw = *reinterpret_cast<const uint4*>(grec);
    unpack_header(hw.x, hw.y, hw.z, hw.w, h);
    const uint2* src = reinterpret_cast<const uint2*>(grec + kHeaderWords);
    for (int q = 0; q < staged_pairs; ++q) {
      const uint2 v = src[q];
      mine[(2 * q + 0) * 32] = v.x;
      mine[(2 * q + 1) * 32] = v.y;
    }
  }
  // every thread only ever touches its own column of the staging buffer: no barrier needed

  PlayoutRef<NT, SP> b{mine, grec + kHeaderWords, n};
  SmemStack<PlayoutCfg<SP>::kStackWords> stk{mine + SP * n * 32, 0, false};
  if (active) count_cache_build(b);
  const uint32_t swapped_before = h.swapped;
  const bool open_at_start = active && h.result == kOpen;

  const uint64_t stream =
      !active ? 0ull : (a.stream_ids != nullptr ? a.stream_ids[idx] : a.stream_base + static_cast<uint64_t>(idx));
  const uint32_t s_lo = static_cast<uint32_t>(stream), s_hi = static_cast<uint32_t>(stream >> 32);
  const uint32_t k_lo = static_cast<uint32_t>(a.seed), k_hi = static_cast<uint32_t>(a.seed >> 32);

  uint32_t r0 = 0, r1 = 0, r2 = 0, r3 = 0;
  int step = 0;
  uint32_t pend = 0, origin = 0;
  int fplane = P_START;
  bool playing = open_at_start && a.max_plies > 0;
  while (__any_sync(kFullMask, playing || pend != 0u || !stk.empty())) {
    // ---- MOVE: lanes with no flood work left make their next move -----------
    if (playing && pend == 0u && stk.empty()) {
      if ((step & 3) == 0) {
        uint32_t r[4];
        philox4x32_10(s_lo, s_hi, static_cast<uint32_t>(step) >> 2, 0u, k_lo, k_hi, r);
        r0 = r[0]; r1 = r[1]; r2 = r[2]; r3 = r[3];
      }
      const uint32_t word = (step & 2) ? ((step & 1) ? r3 : r2) : ((step & 1) ? r1 : r0);
      const int L = legal_count(h, n);
      const int k = static_cast<int>(playout_index(word, static_cast<uint32_t>(L)));
      int x, y;
      select_legal(b, h, k, x, y);
      if (a.out_actions != nullptr && step < a.trace_plies)
        a.out_actions[static_cast<int64_t>(step) * a.count + idx] = static_cast<uint16_t>(x * n + y);
      apply_begin(b, h, x, y, pend);
      origin = static_cast<uint32_t>((x << 8) | y);
      ++step;
      playing = h.result == kOpen && step < a.max_plies;
    }
    // ---- FLOOD: one visit for the lanes that owe border-flag propagation -----
    if (stk.empty() && pend != 0u) {
      const bool start = (pend & kFloodStart) != 0u;
      fplane = start ? P_START : P_END;
      pend &= start ? ~kFloodStart : ~kFloodEnd;
      stk.push(origin);
    }
    if (!stk.empty()) {
      flood_visit(b, fplane, stk);
      if (stk.empty() && stk.overflow) {
        flood_closure(b, ((h.ply - 1u) & 1u) == kRed ? P_RED : P_BLUE, fplane);
        stk.overflow = false;
      }
    }
  }

  if (active) {
    uint4 hw;
    pack_header(h, hw.x, hw.y, hw.z, hw.w);
    *reinterpret_cast<uint4*>(grec) = hw;
    uint2* dst = reinterpret_cast<uint2*>(grec + kHeaderWords);
    for (int q = 0; q < staged_pairs; ++q) dst[q] = make_uint2(mine[(2 * q + 0) * 32], mine[(2 * q + 1) * 32]);
    if (a.out_returns != nullptr) {
      const float r = h.result == kRedWin ? 1.0f : (h.result == kBlueWin ? -1.0f : 0.0f);
      reinterpret_cast<float2*>(a.out_returns)[idx] = make_float2(r, r == 0.0f ? 0.0f : -r);
    }
    if (a.out_lengths != nullptr) a.out_lengths[idx] = step;
  }

  // batch statistics: warp-reduce, one atomic per warp and counter
  if (a.stats != nullptr) {
    const bool finished = open_at_start && h.result != kOpen;
    unsigned long long plies = static_cast<unsigned long long>(step);
    for (int o = 16; o > 0; o >>= 1) plies += __shfl_xor_sync(kFullMask, plies, o);
    const unsigned games = __popc(__ballot_sync(kFullMask, finished));
    const unsigned red = __popc(__ballot_sync(kFullMask, finished && h.result == kRedWin));
    const unsigned blue = __popc(__ballot_sync(kFullMask, finished && h.result == kBlueWin));
    const unsigned draws = __popc(__ballot_sync(kFullMask, finished && h.result == kDraw));
    const unsigned swaps = __popc(__ballot_sync(kFullMask, active && h.swapped != swapped_before));
    unsigned maxlen = finished ? h.ply : 0u;
    for (int o = 16; o > 0; o >>= 1) maxlen = max(maxlen, __shfl_xor_sync(kFullMask, maxlen, o));
    if (lane == 0) {
      if (plies) atomicAdd(&a.stats->plies, plies);
      if (games) atomicAdd(&a.stats->games, static_cast<unsigned long long>(games));
      if (red) atomicAdd(&a.stats->red_wins, static_cast<unsigned long long>(red));
      if (blue) atomicAdd(&a.stats->blue_wins, static_cast<unsigned long long>(blue));
      if (draws) atomicAdd(&a.stats->draws, static_cast<unsigned long long>(draws));
      if (swaps) atomicAdd(&a.stats->swaps, static_cast<unsigned long long>(swaps));
      if (maxlen) atomicMax(&a.stats->max_length, static_cast<unsigned long long>(maxlen));
    }
  }
}

int g_smem_planes = 8;  // TWIXT_PLAYOUT_SMEM_PLANES=2|8 (experiments); see PlayoutCfg

template <int NT, int SP>
cudaError_t launch_nt(const PlayoutArgs& a, cudaStream_t s) {
  const size_t smem = static_cast<size_t>(kPlayoutThreads) * PlayoutCfg<SP>::words(a.n) * sizeof(uint32_t);
  const int64_t blocks = (a.count + kPlayoutThreads - 1) / kPlayoutThreads;
  playout_kernel<NT, SP><<<static_cast<unsigned>(blocks), kPlayoutThreads, smem, s>>>(a);
  return cudaGetLastError();
}

template <int NT, int SP>
cudaError_t setup_nt(int n_for_size) {
  const size_t smem = static_cast<size_t>(kPlayoutThreads) * PlayoutCfg<SP>::words(n_for_size) * sizeof(uint32_t);
  cudaError_t e = cudaFuncSetAttribute(playout_kernel<NT, SP>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(smem));
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(playout_kernel<NT, SP>, cudaFuncAttributePreferredSharedMemoryCarveout,
                              cudaSharedmemCarveoutMaxShared);
}

template <int SP>
cudaError_t launch_sp(const PlayoutArgs& a, cudaStream_t s) {
  switch (a.n) {
    case 8: return launch_nt<8, SP>(a, s);
    case 12: return launch_nt<12, SP>(a, s);
    case 24: return launch_nt<24, SP>(a, s);
    default: return launch_nt<0, SP>(a, s);
  }
}

template <int SP>
cudaError_t setup_sp() {
  cudaError_t e;
  if ((e = setup_nt<0, SP>(24)) != cudaSuccess) return e;
  if ((e = setup_nt<8, SP>(8)) != cudaSuccess) return e;
  if ((e = setup_nt<12, SP>(12)) != cudaSuccess) return e;
  return setup_nt<24, SP>(24);
}

}  // namespace

cudaError_t playout_setup() {
  const char* v = getenv("TWIXT_PLAYOUT_SMEM_PLANES");
  if (v != nullptr && (v[0] == '2' || v[0] == '8')) g_smem_planes = v[0] - '0';
  cudaError_t e = setup_sp<8>();
  if (e != cudaSuccess) return e;
  return setup_sp<2>();
}

cudaError_t launch_playout(const PlayoutArgs& a, cudaStream_t s) {
  if (a.count <= 0) return cudaSuccess;
  return g_smem_planes == 2 ? launch_sp<2>(a, s) : launch_sp<8>(a, s);
}

}  // namespace twixt
