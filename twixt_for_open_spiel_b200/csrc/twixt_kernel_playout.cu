// twixt_kernel_playout.cu -- K5, the fused random-playout kernel.
//
// One THREAD per env.  Each warp copies the records of its 32 consecutive envs
// from HBM into shared memory, transposed so that word w of lane l sits at
// smem[w*32 + l]: whatever word each lane indexes, lane l always hits bank l,
// so every access of the scalar rules in twixt_engine.cuh is conflict-free.
// The whole game is then played out of shared memory -- select a uniformly
// random legal action (Philox4x32-10, one block per four moves), apply it
// (peg, links, crossing test, border flags, result) -- and the final records
// are written back once.  HBM sees 2 x record bytes per GAME, not per move;
// the limiter is issue slots and shared-memory latency, which is why the
// thread-per-env mapping is used: it spends ~1/10 of the warp-instructions per
// move of a warp-per-env mapping because nothing is computed redundantly
// across lanes ("Mapping one game onto the machine").
//
// Reference loop reproduced: upstream example.cc / RandomRolloutEvaluator
// (LegalActions -> uniform pick -> ApplyAction until IsTerminal), with the
// per-move semantics of twixtboard.cc:457-499.
#include <cuda_runtime.h>
#include <stdint.h>

#include "twixt_engine.cuh"
#include "twixt_kernels.cuh"
#include "twixt_philox.cuh"

namespace twixt {

namespace {

constexpr int kPlayoutThreads = 128;  // 4 warps; n=24: 110 KB of records per block, 2 blocks per SM
constexpr int kFloodStack = 48;
constexpr unsigned kFullMask = 0xFFFFFFFFu;

template <int NT>
__global__ void __launch_bounds__(kPlayoutThreads) playout_kernel(const PlayoutArgs a) {
  extern __shared__ uint4 smem_raw[];
  uint32_t* smem = reinterpret_cast<uint32_t*>(smem_raw);
  const int n = NT > 0 ? NT : a.n;
  const int rw = record_words(n);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t idx = blockIdx.x * static_cast<int64_t>(kPlayoutThreads) + threadIdx.x;
  const bool active = idx < a.count;
  uint32_t* mine = smem + warp * (rw * 32) + lane;
  uint32_t* grec = a.records + idx * rw;

  if (active) {
    const uint4* src = reinterpret_cast<const uint4*>(grec);
    for (int q = 0; q < rw / 4; ++q) {
      const uint4 v = src[q];
      mine[(4 * q + 0) * 32] = v.x;
      mine[(4 * q + 1) * 32] = v.y;
      mine[(4 * q + 2) * 32] = v.z;
      mine[(4 * q + 3) * 32] = v.w;
    }
  }
  // every thread only ever touches its own column of the staging buffer

  RecordRef<32, NT> b{mine, n};
  Header h;
  h.ply = 0; h.result = kDraw; h.swapped = 0; h.move_one = kNoMove; h.cnt[0] = h.cnt[1] = 0;
  if (active) load_header(b, h);
  const uint32_t swapped_before = h.swapped;
  const bool open_at_start = active && h.result == kOpen;

  const uint64_t stream = !active ? 0ull : (a.stream_ids != nullptr ? a.stream_ids[idx] : a.stream_base + static_cast<uint64_t>(idx));
  const uint32_t s_lo = static_cast<uint32_t>(stream), s_hi = static_cast<uint32_t>(stream >> 32);
  const uint32_t k_lo = static_cast<uint32_t>(a.seed), k_hi = static_cast<uint32_t>(a.seed >> 32);

  uint32_t r0 = 0, r1 = 0, r2 = 0, r3 = 0;
  int step = 0;
  while (h.result == kOpen && step < a.max_plies) {
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
    apply_legal_cell<kFloodStack>(b, h, x, y);
    ++step;
  }

  if (active) {
    store_header(b, h);
    uint4* dst = reinterpret_cast<uint4*>(grec);
    for (int q = 0; q < rw / 4; ++q) {
      uint4 v;
      v.x = mine[(4 * q + 0) * 32];
      v.y = mine[(4 * q + 1) * 32];
      v.z = mine[(4 * q + 2) * 32];
      v.w = mine[(4 * q + 3) * 32];
      dst[q] = v;
    }
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

template <int NT>
cudaError_t launch_nt(const PlayoutArgs& a, cudaStream_t s) {
  const int rw = record_words(a.n);
  const size_t smem = static_cast<size_t>(kPlayoutThreads) * rw * sizeof(uint32_t);
  const int64_t blocks = (a.count + kPlayoutThreads - 1) / kPlayoutThreads;
  playout_kernel<NT><<<static_cast<unsigned>(blocks), kPlayoutThreads, smem, s>>>(a);
  return cudaGetLastError();
}

template <int NT>
cudaError_t setup_nt(int n_for_size) {
  const size_t smem = static_cast<size_t>(kPlayoutThreads) * record_words(n_for_size) * sizeof(uint32_t);
  return cudaFuncSetAttribute(playout_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
}

}  // namespace

cudaError_t playout_setup() {
  cudaError_t e;
  if ((e = setup_nt<0>(24)) != cudaSuccess) return e;
  if ((e = setup_nt<8>(8)) != cudaSuccess) return e;
  if ((e = setup_nt<12>(12)) != cudaSuccess) return e;
  if ((e = setup_nt<24>(24)) != cudaSuccess) return e;
  return cudaSuccess;
}

cudaError_t launch_playout(const PlayoutArgs& a, cudaStream_t s) {
  if (a.count <= 0) return cudaSuccess;
  switch (a.n) {
    case 8: return launch_nt<8>(a, s);
    case 12: return launch_nt<12>(a, s);
    case 24: return launch_nt<24>(a, s);
    default: return launch_nt<0>(a, s);
  }
}

}  // namespace twixt
