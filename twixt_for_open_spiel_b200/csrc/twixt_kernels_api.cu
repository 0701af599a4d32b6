// twixt_kernels_api.cu -- the per-call kernels behind the State surface:
// reset (NewInitialState), clone, LegalActions (list + mask), ApplyAction,
// CurrentPlayer / IsTerminal / Returns and ObservationTensor.
//
// All of them are HBM-bound byte/bit work (no tensor cores: nothing here is a
// contraction).  Mapping:
//   reset / clone       one thread per 16-byte piece of a record (128-bit stores)
//   legal list / mask   one warp per env, lane = board column; popc + warp
//                       prefix sum compacts the ascending action list
//   apply               one thread per env working in place on its record
//                       (touches only the few words a move needs)
//   observation         one block per env, planes staged in shared memory,
//                       float4 stores
#include <cuda_runtime.h>
#include <stdint.h>

#include "twixt_engine.cuh"
#include "twixt_kernels.cuh"

namespace twixt {

namespace {

constexpr unsigned kFullMask = 0xFFFFFFFFu;
constexpr int kFloodStack = 48;

__device__ __forceinline__ uint4 ldg128(const uint32_t* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }

// ---------------------------------------------------------------- reset ---
// Board::Board (twixtboard.cc:168-174): empty board, both legal lists full.
__global__ void reset_kernel(uint32_t* __restrict__ records, int64_t count, int n, int quads_per_record) {
  const int64_t total = count * quads_per_record;
  const uint32_t cnt = static_cast<uint32_t>(n * (n - 2));
  for (int64_t q = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; q < total;
       q += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int in_rec = static_cast<int>(q % quads_per_record);
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (in_rec == 0) v = make_uint4(0u, 0u, kNoMove, cnt | (cnt << 16));
    reinterpret_cast<uint4*>(records)[q] = v;
  }
}

// ---------------------------------------------------------------- clone ---
// State::Clone (twixt.h:80-82): dst env i <- src env (src_ids ? src_ids[i] : i)
__global__ void clone_kernel(uint32_t* __restrict__ dst, const uint32_t* __restrict__ src,
                             const int64_t* __restrict__ src_ids, int64_t count, int quads_per_record) {
  const int64_t total = count * quads_per_record;
  for (int64_t q = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; q < total;
       q += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t env = q / quads_per_record;
    const int in_rec = static_cast<int>(q - env * quads_per_record);
    const int64_t from = src_ids ? src_ids[env] : env;
    reinterpret_cast<uint4*>(dst)[q] = __ldg(reinterpret_cast<const uint4*>(src) + from * quads_per_record + in_rec);
  }
}

// ------------------------------------------------------ legal list / mask ---
// Legal word of column `lane` for the warp's env (0 for lanes >= n and for
// terminal envs): TwixTState::LegalActions twixt.h:86-90.
__device__ __forceinline__ uint32_t warp_legal_word(const uint32_t* rec, int n, int lane) {
  const uint4 hw = ldg128(rec);
  Header h;
  unpack_header(hw.x, hw.y, hw.z, hw.w, h);
  if (h.result != kOpen || lane >= n) return 0u;
  const int player = static_cast<int>(h.ply & 1u);
  const uint32_t play = playable_word(n, player, lane);
  if (h.ply == 1u) return play;
  const uint32_t occ = __ldg(rec + kHeaderWords + lane) | __ldg(rec + kHeaderWords + n + lane);
  return play & ~occ;
}

template <typename T>
__global__ void legal_actions_kernel(const uint32_t* __restrict__ records, int64_t count, int n, int rw,
                                     T* __restrict__ out_actions, int64_t stride, int32_t* __restrict__ out_counts) {
  const int lane = threadIdx.x & 31;
  const int64_t env = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
  if (env >= count) return;  // whole warp leaves together
  const uint32_t w = warp_legal_word(records + env * rw, n, lane);
  const int c = __popc(w);
  int incl = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(kFullMask, incl, o);
    if (lane >= o) incl += t;
  }
  const int total = __shfl_sync(kFullMask, incl, 31);
  if (out_counts != nullptr && lane == 0) out_counts[env] = total;
  if (out_actions != nullptr) {
    T* row = out_actions + env * stride + (incl - c);
    uint32_t rest = w;
    int r = 0;
    while (rest) {  // ascending rows of this column: action = x*n + y
      const int y = __ffs(static_cast<int>(rest)) - 1;
      rest &= rest - 1u;
      row[r++] = static_cast<T>(lane * n + y);
    }
  }
}

__global__ void legal_mask_kernel(const uint32_t* __restrict__ records, int64_t count, int n, int rw,
                                  uint8_t* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t env = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
  if (env >= count) return;
  const uint32_t w = warp_legal_word(records + env * rw, n, lane);
  const int cells = n * n;
  uint8_t* row = out + env * cells;
  for (int base = 0; base < cells; base += 32) {
    const int c = base + lane;
    const int x = min(c / n, n - 1);
    const int y = c - x * n;
    const uint32_t col = __shfl_sync(kFullMask, w, x);
    if (c < cells) row[c] = static_cast<uint8_t>((col >> y) & 1u);
  }
}

// ---------------------------------------------------------------- apply ---
// TwixTState::DoApplyAction (twixt.h:93-104): legality, Board::ApplyAction,
// turn hand-over (implicit in ply parity / result).
__global__ void apply_kernel(uint32_t* __restrict__ records, int64_t count, int n, int rw,
                             const int32_t* __restrict__ actions, int32_t* __restrict__ out_status,
                             DeviceStats* __restrict__ stats) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= count) return;
  const int action = actions[i];
  int status = 2;
  if (action >= 0) {
    RecordRef<1> b{records + i * rw, n};
    const uint4 hw = *reinterpret_cast<const uint4*>(b.p);
    Header h;
    unpack_header(hw.x, hw.y, hw.z, hw.w, h);
    if (!is_legal(b, h, action)) {
      status = 1;
      atomicMin(&stats->illegal_index, static_cast<unsigned int>(i));  // host names this action in its message
    } else {
      const int x = action / n;
      apply_legal_cell<kFloodStack>(b, h, x, action - x * n);
      uint4 o;
      pack_header(h, o.x, o.y, o.z, o.w);
      *reinterpret_cast<uint4*>(b.p) = o;
      status = 0;
    }
  }
  if (out_status != nullptr) out_status[i] = status;
}

// ---------------------------------------------------------------- query ---
// CurrentPlayer twixt.h:38, IsTerminal twixt.h:45-48, Returns twixt.h:50-63.
__global__ void query_kernel(const uint32_t* __restrict__ records, int64_t count, int rw,
                             int8_t* __restrict__ out_player, uint8_t* __restrict__ out_terminal,
                             float* __restrict__ out_returns) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= count) return;
  const uint4 hw = ldg128(records + i * rw);
  Header h;
  unpack_header(hw.x, hw.y, hw.z, hw.w, h);
  if (out_player != nullptr) out_player[i] = static_cast<int8_t>(current_player(h));
  if (out_terminal != nullptr) out_terminal[i] = h.result != kOpen ? 1 : 0;
  if (out_returns != nullptr) {
    const float r = h.result == kRedWin ? 1.0f : (h.result == kBlueWin ? -1.0f : 0.0f);
    reinterpret_cast<float2*>(out_returns)[i] = make_float2(r, r == 0.0f ? 0.0f : -r);
  }
}

// ---------------------------------------------------------- observation ---
// TwixTState::ObservationTensor (twixt.cc:101-132): [12, n, n-2] float32 per
// env.  The 12 planes are first formed as column words in shared memory
// (obs_plane_word), then every thread writes 4 consecutive floats.
constexpr int kObsThreads = 256;
constexpr int kObsPlaneWords = 12 * 24;

template <bool kVec4>
__global__ void __launch_bounds__(kObsThreads) observation_kernel(const uint32_t* __restrict__ records, int64_t count,
                                                                  int n, int rw, float* __restrict__ out) {
  __shared__ uint32_t planes[kObsPlaneWords];  // [12][n]
  const int64_t env = blockIdx.x;
  if (env >= count) return;
  RecordRef<1> b{const_cast<uint32_t*>(records + env * rw), n};
  for (int t = threadIdx.x; t < 12 * n; t += kObsThreads) {
    const int p = t / n;
    planes[t] = obs_plane_word(b, p, t - p * n);
  }
  __syncthreads();
  const int w = n - 2;
  const int plane_size = n * w;
  const int total = 12 * plane_size;
  float* dst = out + env * static_cast<int64_t>(total);
  // exact for the ranges used (j < 2^13, divisors < 2^10): see DESIGN.md
  const uint32_t m_plane = ((1u << 24) + plane_size - 1) / plane_size;
  const uint32_t m_row = ((1u << 24) + w - 1) / w;
  constexpr int kPer = kVec4 ? 4 : 1;
  for (int j = threadIdx.x * kPer; j < total; j += kObsThreads * kPer) {
    int p = static_cast<int>((static_cast<uint32_t>(j) * m_plane) >> 24);
    const int rem = j - p * plane_size;
    int r = static_cast<int>((static_cast<uint32_t>(rem) * m_row) >> 24);
    int c = rem - r * w;
    float v[kPer];
#pragma unroll
    for (int e = 0; e < kPer; ++e) {
      int x, y;
      obs_cell(n, p, r, c, x, y);
      v[e] = ((planes[p * n + x] >> y) & 1u) ? 1.0f : 0.0f;
      if (++c == w) {
        c = 0;
        if (++r == n) { r = 0; ++p; }
      }
      if (p >= 12) { p = 11; }  // tail of the last vector stays in bounds (never stored past total)
    }
    if (kVec4) {
      *reinterpret_cast<float4*>(dst + j) = make_float4(v[0], v[1 % kPer], v[2 % kPer], v[3 % kPer]);
    } else {
      dst[j] = v[0];
    }
  }
}

inline int grid_for(int64_t items, int threads, int max_blocks = 148 * 64) {
  int64_t blocks = (items + threads - 1) / threads;
  if (blocks < 1) blocks = 1;
  if (blocks > max_blocks) blocks = max_blocks;
  return static_cast<int>(blocks);
}

}  // namespace

cudaError_t launch_reset(uint32_t* records, int64_t count, int n, cudaStream_t s) {
  if (count <= 0) return cudaSuccess;
  const int quads = record_words(n) / 4;
  reset_kernel<<<grid_for(count * quads, 256), 256, 0, s>>>(records, count, n, quads);
  return cudaGetLastError();
}

cudaError_t launch_clone(uint32_t* dst, const uint32_t* src, const int64_t* src_ids, int64_t count, int n,
                         cudaStream_t s) {
  if (count <= 0) return cudaSuccess;
  const int quads = record_words(n) / 4;
  clone_kernel<<<grid_for(count * quads, 256), 256, 0, s>>>(dst, src, src_ids, count, quads);
  return cudaGetLastError();
}

cudaError_t launch_legal_actions(const uint32_t* records, int64_t count, int n, void* out_actions, int elem_bytes,
                                 int64_t stride, int32_t* out_counts, cudaStream_t s) {
  if (count <= 0) return cudaSuccess;
  const int threads = 256;
  const int64_t blocks = (count * 32 + threads - 1) / threads;
  const int rw = record_words(n);
  if (elem_bytes == 2)
    legal_actions_kernel<uint16_t><<<static_cast<unsigned>(blocks), threads, 0, s>>>(
        records, count, n, rw, static_cast<uint16_t*>(out_actions), stride, out_counts);
  else if (elem_bytes == 4)
    legal_actions_kernel<int32_t><<<static_cast<unsigned>(blocks), threads, 0, s>>>(
        records, count, n, rw, static_cast<int32_t*>(out_actions), stride, out_counts);
  else
    legal_actions_kernel<int64_t><<<static_cast<unsigned>(blocks), threads, 0, s>>>(
        records, count, n, rw, static_cast<int64_t*>(out_actions), stride, out_counts);
  return cudaGetLastError();
}

cudaError_t launch_legal_mask(const uint32_t* records, int64_t count, int n, uint8_t* out, cudaStream_t s) {
  if (count <= 0) return cudaSuccess;
  const int threads = 256;
  const int64_t blocks = (count * 32 + threads - 1) / threads;
  legal_mask_kernel<<<static_cast<unsigned>(blocks), threads, 0, s>>>(records, count, n, record_words(n), out);
  return cudaGetLastError();
}

cudaError_t launch_apply(uint32_t* records, int64_t count, int n, const int32_t* actions, int32_t* out_status,
                         DeviceStats* stats, cudaStream_t s) {
  if (count <= 0) return cudaSuccess;
  const int threads = 128;
  const int64_t blocks = (count + threads - 1) / threads;
  apply_kernel<<<static_cast<unsigned>(blocks), threads, 0, s>>>(records, count, n, record_words(n), actions,
                                                                 out_status, stats);
  return cudaGetLastError();
}

cudaError_t launch_query(const uint32_t* records, int64_t count, int n, int8_t* out_player, uint8_t* out_terminal,
                         float* out_returns, cudaStream_t s) {
  if (count <= 0) return cudaSuccess;
  const int threads = 256;
  const int64_t blocks = (count + threads - 1) / threads;
  query_kernel<<<static_cast<unsigned>(blocks), threads, 0, s>>>(records, count, record_words(n), out_player,
                                                                 out_terminal, out_returns);
  return cudaGetLastError();
}

cudaError_t launch_observation(const uint32_t* records, int64_t count, int n, float* out, cudaStream_t s) {
  if (count <= 0) return cudaSuccess;
  const int rw = record_words(n);
  const bool vec = (reinterpret_cast<uintptr_t>(out) & 15u) == 0;
  // grid.x is limited to 2^31-1 blocks, far above any batch that fits in HBM
  if (vec)
    observation_kernel<true><<<static_cast<unsigned>(count), kObsThreads, 0, s>>>(records, count, n, rw, out);
  else
    observation_kernel<false><<<static_cast<unsigned>(count), kObsThreads, 0, s>>>(records, count, n, rw, out);
  return cudaGetLastError();
}

}  // namespace twixt
