// twixt_batch.cu -- host class TwixTBatch and the extern "C" ABI of
// include/twixt_b200.h.
//
// TwixTBatch owns the env records in HBM, a CUDA stream, a scratch arena used
// to stage host-side inputs/outputs, and the device counters.  It mirrors the
// open_spiel Game/State surface of the reference (twixt.h:31-146) one batched
// method per State method; the C functions at the bottom are thin wrappers
// that translate exceptions-free status codes and keep the last error text.
// There is no CPU implementation behind any of these calls.
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>

#include "../../include/twixt_b200.h"
#include "twixt_engine.cuh"
#include "twixt_kernels.cuh"

namespace twixt {

namespace {

thread_local std::string g_last_error;

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}

#define TW_CUDA(expr)                                                                      \
  do {                                                                                     \
    cudaError_t e_ = (expr);                                                               \
    if (e_ != cudaSuccess) return fail(TWIXT_ECUDA, "CUDA error %s at %s:%d (%s)",         \
                                       cudaGetErrorString(e_), __FILE__, __LINE__, #expr); \
  } while (0)

#define TW_TRY(expr)        \
  do {                      \
    int rc_ = (expr);       \
    if (rc_ != TWIXT_OK) return rc_; \
  } while (0)

bool is_device_pointer(const void* p) {
  cudaPointerAttributes attr;
  cudaError_t e = cudaPointerGetAttributes(&attr, p);
  if (e != cudaSuccess) {
    cudaGetLastError();  // unregistered host memory on old drivers: clear and treat as host
    return false;
  }
  return attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged;
}

class DeviceGuard {
 public:
  explicit DeviceGuard(int device) : prev_(-1) {
    cudaGetDevice(&prev_);
    if (prev_ != device) cudaSetDevice(device);
    dev_ = device;
  }
  ~DeviceGuard() {
    if (prev_ >= 0 && prev_ != dev_) cudaSetDevice(prev_);
  }

 private:
  int prev_, dev_;
};

int fill_game_info(int n, twixt_game_info* out) {
  if (n < TWIXT_MIN_BOARD_SIZE || n > TWIXT_MAX_BOARD_SIZE)  // twixt.cc:139-144, same text
    return fail(TWIXT_EINVAL, "board_size out of range [%d..%d]: %d", TWIXT_MIN_BOARD_SIZE, TWIXT_MAX_BOARD_SIZE, n);
  if (out == nullptr) return fail(TWIXT_EINVAL, "null output pointer");
  out->board_size = n;
  out->num_distinct_actions = n * n;
  out->num_players = 2;
  out->max_game_length = n * n - 4 + 1;
  out->obs_shape[0] = TWIXT_NUM_OBS_PLANES;
  out->obs_shape[1] = n;
  out->obs_shape[2] = n - 2;
  out->obs_size = TWIXT_NUM_OBS_PLANES * n * (n - 2);
  out->max_legal_actions = n * (n - 2);
  out->record_words = record_words(n);
  out->min_utility = -1.0;
  out->max_utility = 1.0;
  out->utility_sum = 0.0;
  return TWIXT_OK;
}

}  // namespace

// A host output/input that may live on either side of the bus.
struct Staged {
  void* user = nullptr;     // what the caller passed
  void* dev = nullptr;      // what the kernel uses
  void* pin = nullptr;      // host alias of `dev` when the small-transfer arena is used (see PinAlloc)
  size_t bytes = 0;
  bool host = false;
};

class TwixTBatch {
 public:
  TwixTBatch() = default;
  ~TwixTBatch() {
    DeviceGuard g(device_);
    if (stream_ != nullptr) cudaStreamSynchronize(stream_);
    if (records_ != nullptr) cudaFree(records_);
    if (d_stats_ != nullptr) cudaFree(d_stats_);
    if (pinned_ != nullptr) cudaFreeHost(pinned_);
    for (int k = 0; k < kSlots; ++k)
      if (slot_[k] != nullptr) cudaFree(slot_[k]);
    if (own_stream_ && stream_ != nullptr) cudaStreamDestroy(stream_);
  }

  int Init(int n, int64_t num_envs, int device, uint64_t seed) {
    twixt_game_info info;
    TW_TRY(fill_game_info(n, &info));
    if (num_envs <= 0) return fail(TWIXT_EINVAL, "num_envs must be positive: %lld", static_cast<long long>(num_envs));
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
      cudaGetLastError();
      return fail(TWIXT_ECUDA, "no CUDA device available (%s): libtwixt_b200 has no CPU fallback",
                  e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    }
    if (device < 0 || device >= ndev) return fail(TWIXT_EINVAL, "device %d out of range [0..%d)", device, ndev);
    cudaDeviceProp prop;
    TW_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
      return fail(TWIXT_ECUDA, "device %d is sm_%d%d; libtwixt_b200 is built for sm_100a only", device, prop.major,
                  prop.minor);
    n_ = n;
    num_envs_ = num_envs;
    device_ = device;
    seed_ = seed;
    rw_ = info.record_words;
    DeviceGuard g(device_);
    TW_CUDA(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
    own_stream_ = true;
    const size_t bytes = static_cast<size_t>(num_envs_) * rw_ * sizeof(uint32_t);
    e = cudaMalloc(&records_, bytes);
    if (e != cudaSuccess) {
      cudaGetLastError();
      return fail(TWIXT_ENOMEM, "cudaMalloc of %zu bytes for %lld envs failed: %s", bytes,
                  static_cast<long long>(num_envs_), cudaGetErrorString(e));
    }
    TW_CUDA(cudaMalloc(&d_stats_, sizeof(DeviceStats)));
    if (cudaHostAlloc(&pinned_, kPinBytes, cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess ||
        cudaHostGetDevicePointer(&pinned_dev_, pinned_, 0) != cudaSuccess) {
      cudaGetLastError();  // no mapped memory on this system: every host transfer takes the staged path
      if (pinned_ != nullptr) cudaFreeHost(pinned_);
      pinned_ = pinned_dev_ = nullptr;
    }
    TW_CUDA(cudaMemsetAsync(d_stats_, 0, sizeof(DeviceStats), stream_));
    TW_CUDA(playout_setup());
    return Reset(0, num_envs_);
  }

  int n() const { return n_; }
  int64_t num_envs() const { return num_envs_; }
  int rw() const { return rw_; }
  int device() const { return device_; }
  const uint32_t* records() const { return records_; }
  cudaStream_t stream() const { return stream_; }

  int SetStream(cudaStream_t s) {
    DeviceGuard g(device_);
    TW_CUDA(cudaStreamSynchronize(stream_));
    if (own_stream_) cudaStreamDestroy(stream_);
    stream_ = s;
    own_stream_ = false;
    return TWIXT_OK;
  }
  int Synchronize() {
    DeviceGuard g(device_);
    TW_CUDA(cudaStreamSynchronize(stream_));
    return TWIXT_OK;
  }

  // One State step for an unbatched caller (see twixt_step in the header): apply (or reset, or nothing),
  // then player / terminal / returns / ascending legal actions of the resulting state, by ONE kernel launch
  // whose outputs land in the pinned arena.
  int Step(int64_t env, int32_t action, twixt_step_result* out, int64_t* out_legal) {
    TW_TRY(CheckRange(env, 1));
    if (out == nullptr) return fail(TWIXT_EINVAL, "null output pointer");
    if (action < TWIXT_STEP_RESET) return fail(TWIXT_EINVAL, "action %d: use an action >= 0, TWIXT_STEP_QUERY or TWIXT_STEP_RESET", action);
    DeviceGuard g(device_);
    ScratchReset();
    Staged res, legal;
    TW_TRY(StageOut(out, sizeof(twixt_step_result), &res, false));
    TW_TRY(StageOut(out_legal, static_cast<size_t>(n_) * (n_ - 2) * sizeof(int64_t), &legal, false));
    TW_CUDA(launch_step(rec(env), n_, action, static_cast<twixt_step_result*>(res.dev), static_cast<int64_t*>(legal.dev),
                        stream_));
    launches_ += 1;
    TW_TRY(Finish(&res));
    TW_TRY(Finish(&legal));
    if (!res.host && !legal.host) return TWIXT_OK;
    TW_TRY(Sync());
    if (res.host && out->status == 1) return fail(TWIXT_EILLEGAL, "Not a legal action: %d", action);  // twixt.h:96
    return TWIXT_OK;
  }
  void SetSeed(uint64_t seed) { seed_ = seed; }
  void SetStreamBase(uint64_t base) { stream_base_ = base; }

  int CheckRange(int64_t first, int64_t count) const {
    if (first < 0 || count < 0 || first + count > num_envs_)
      return fail(TWIXT_EINVAL, "env range [%lld, %lld) outside [0, %lld)", static_cast<long long>(first),
                  static_cast<long long>(first + count), static_cast<long long>(num_envs_));
    return TWIXT_OK;
  }

  int Reset(int64_t first, int64_t count) {
    TW_TRY(CheckRange(first, count));
    DeviceGuard g(device_);
    TW_CUDA(launch_reset(rec(first), count, n_, stream_));
    launches_ += count > 0;
    return TWIXT_OK;
  }

  int Clone(int64_t src_first, int64_t dst_first, int64_t count) {
    TW_TRY(CheckRange(src_first, count));
    TW_TRY(CheckRange(dst_first, count));
    if (count > 0 && src_first < dst_first + count && dst_first < src_first + count)
      return fail(TWIXT_EINVAL, "clone ranges overlap");
    DeviceGuard g(device_);
    TW_CUDA(launch_clone(rec(dst_first), rec(src_first), nullptr, count, n_, num_envs_, dst_first, d_stats_, stream_));
    launches_ += count > 0;
    return TWIXT_OK;
  }

  int CloneGather(const int64_t* src_ids, int64_t dst_first, int64_t count) {
    TW_TRY(CheckRange(dst_first, count));
    if (count == 0) return TWIXT_OK;
    if (src_ids == nullptr) return fail(TWIXT_EINVAL, "null src_ids");
    DeviceGuard g(device_);
    ScratchReset();
    Staged ids;
    TW_TRY(StageIn(src_ids, static_cast<size_t>(count) * sizeof(int64_t), &ids));
    if (ids.host) {
      for (int64_t i = 0; i < count; ++i) {
        const int64_t s = src_ids[i];
        if (s < 0 || s >= num_envs_) return fail(TWIXT_EINVAL, "src_ids[%lld] = %lld out of range", (long long)i, (long long)s);
        if (s >= dst_first && s < dst_first + count) return fail(TWIXT_EINVAL, "src_ids[%lld] lies in the destination range", (long long)i);
      }
    }
    // device-resident ids cannot be checked here: the kernel checks every id (it skips a bad one and reports
    // the lowest offending position), and the flag is read back before the call returns
    TW_CUDA(cudaMemsetAsync(&d_stats_->bad_clone_index, 0xFF, sizeof(unsigned int), stream_));
    TW_CUDA(launch_clone(rec(dst_first), records_, static_cast<const int64_t*>(ids.dev), count, n_, num_envs_,
                         dst_first, d_stats_, stream_));
    launches_ += 1;
    unsigned int bad = 0xFFFFFFFFu;
    TW_CUDA(cudaMemcpyAsync(&bad, &d_stats_->bad_clone_index, sizeof(bad), cudaMemcpyDeviceToHost, stream_));
    TW_TRY(Sync());
    if (bad != 0xFFFFFFFFu)
      return fail(TWIXT_EINVAL, "src_ids[%u] is out of range or lies in the destination range (that env was not copied)", bad);
    return TWIXT_OK;
  }

  int CloneFrom(int64_t dst_first, const TwixTBatch& src, int64_t src_first, int64_t count) {
    TW_TRY(CheckRange(dst_first, count));
    TW_TRY(src.CheckRange(src_first, count));
    if (src.n_ != n_) return fail(TWIXT_EINVAL, "board sizes differ: %d vs %d", src.n_, n_);
    if (count == 0) return TWIXT_OK;
    if (&src == this && src_first < dst_first + count && dst_first < src_first + count)
      return fail(TWIXT_EINVAL, "clone ranges overlap");
    DeviceGuard g(device_);
    const size_t bytes = static_cast<size_t>(count) * rw_ * sizeof(uint32_t);
    // order after the source batch's pending work
    if (src.stream_ != stream_) {
      DeviceGuard gs(src.device_);
      TW_CUDA(cudaStreamSynchronize(src.stream_));
    }
    TW_CUDA(cudaMemcpyAsync(rec(dst_first), src.records_ + src_first * src.rw_, bytes, cudaMemcpyDefault, stream_));
    return TWIXT_OK;
  }

  int LegalActions(int64_t first, int64_t count, void* out_actions, int elem_bytes, int64_t stride,
                   int32_t* out_counts) {
    TW_TRY(CheckRange(first, count));
    if (elem_bytes != 2 && elem_bytes != 4 && elem_bytes != 8)
      return fail(TWIXT_EINVAL, "elem_bytes must be 2, 4 or 8: %d", elem_bytes);
    if (out_actions != nullptr && stride < n_ * (n_ - 2))
      return fail(TWIXT_EINVAL, "stride %lld < max_legal_actions %d", static_cast<long long>(stride), n_ * (n_ - 2));
    if (count == 0) return TWIXT_OK;
    DeviceGuard g(device_);
    ScratchReset();
    Staged acts, cnts;
    TW_TRY(StageOut(out_actions, static_cast<size_t>(count) * stride * elem_bytes, &acts, /*preserve=*/true));
    TW_TRY(StageOut(out_counts, static_cast<size_t>(count) * sizeof(int32_t), &cnts, false));
    TW_CUDA(launch_legal_actions(rec(first), count, n_, acts.dev, elem_bytes, stride,
                                 static_cast<int32_t*>(cnts.dev), stream_));
    launches_ += 1;
    TW_TRY(Finish(&acts));
    TW_TRY(Finish(&cnts));
    return SyncIfHost(acts, cnts);
  }

  int LegalMask(int64_t first, int64_t count, uint8_t* out) {
    TW_TRY(CheckRange(first, count));
    if (count == 0) return TWIXT_OK;
    if (out == nullptr) return fail(TWIXT_EINVAL, "null output pointer");
    DeviceGuard g(device_);
    ScratchReset();
    Staged m;
    TW_TRY(StageOut(out, static_cast<size_t>(count) * n_ * n_, &m, false));
    TW_CUDA(launch_legal_mask(rec(first), count, n_, static_cast<uint8_t*>(m.dev), stream_));
    launches_ += 1;
    TW_TRY(Finish(&m));
    return SyncIfHost(m, m);
  }

  int Apply(int64_t first, int64_t count, const int32_t* actions, int32_t* out_status) {
    TW_TRY(CheckRange(first, count));
    if (count == 0) return TWIXT_OK;
    if (actions == nullptr) return fail(TWIXT_EINVAL, "null actions pointer");
    DeviceGuard g(device_);
    ScratchReset();
    Staged in, st;
    TW_TRY(StageIn(actions, static_cast<size_t>(count) * sizeof(int32_t), &in));
    const size_t status_bytes = static_cast<size_t>(count) * sizeof(int32_t);
    if (out_status == nullptr) PinAlloc(status_bytes, &st);  // small batches: statuses into the arena, read below
    else TW_TRY(StageOut(out_status, status_bytes, &st, false));
    // Where the statuses end up in host-visible memory the first illegal action is found by reading them;
    // otherwise the kernel's device-side flag is reset, and read back after the launch.
    const bool scan_host = st.pin != nullptr;
    if (!scan_host) TW_CUDA(cudaMemsetAsync(&d_stats_->illegal_index, 0xFF, sizeof(unsigned int), stream_));
    TW_CUDA(launch_apply(rec(first), count, n_, static_cast<const int32_t*>(in.dev), static_cast<int32_t*>(st.dev),
                         d_stats_, stream_));
    launches_ += 1;
    TW_TRY(Finish(&st));
    if (out_status != nullptr && !st.host) return TWIXT_OK;  // fully asynchronous: caller inspects the statuses
    int64_t bad = -1;
    if (scan_host) {
      TW_TRY(Sync());
      const int32_t* seen = static_cast<const int32_t*>(st.pin);
      for (int64_t i = 0; i < count && bad < 0; ++i)
        if (seen[i] == 1) bad = i;
    } else {
      unsigned int flag = 0xFFFFFFFFu;
      TW_CUDA(cudaMemcpyAsync(&flag, &d_stats_->illegal_index, sizeof(flag), cudaMemcpyDeviceToHost, stream_));
      TW_TRY(Sync());
      if (flag != 0xFFFFFFFFu) bad = flag;
    }
    if (bad >= 0) {
      int32_t a = 0;
      if (in.host) a = actions[bad];
      else TW_CUDA(cudaMemcpy(&a, actions + bad, sizeof(a), cudaMemcpyDeviceToHost));
      return fail(TWIXT_EILLEGAL, "Not a legal action: %d", a);  // twixt.h:96, same text
    }
    return TWIXT_OK;
  }

  int Query(int64_t first, int64_t count, int8_t* player, uint8_t* terminal, float* returns) {
    TW_TRY(CheckRange(first, count));
    if (count == 0) return TWIXT_OK;
    DeviceGuard g(device_);
    ScratchReset();
    Staged p, t, r;
    TW_TRY(StageOut(player, static_cast<size_t>(count), &p, false));
    TW_TRY(StageOut(terminal, static_cast<size_t>(count), &t, false));
    TW_TRY(StageOut(returns, static_cast<size_t>(count) * 2 * sizeof(float), &r, false));
    TW_CUDA(launch_query(rec(first), count, n_, static_cast<int8_t*>(p.dev), static_cast<uint8_t*>(t.dev),
                         static_cast<float*>(r.dev), stream_));
    launches_ += 1;
    TW_TRY(Finish(&p));
    TW_TRY(Finish(&t));
    TW_TRY(Finish(&r));
    if (p.host || t.host || r.host) TW_TRY(Sync());
    return TWIXT_OK;
  }

  // out_mask != nullptr: the fused producer (tensor + legal mask in one pass over the records)
  int Observation(int64_t first, int64_t count, float* out, uint8_t* out_mask) {
    TW_TRY(CheckRange(first, count));
    if (count == 0) return TWIXT_OK;
    if (out == nullptr) return fail(TWIXT_EINVAL, "null output pointer");
    DeviceGuard g(device_);
    ScratchReset();
    Staged o, m;
    TW_TRY(StageOut(out, static_cast<size_t>(count) * 12 * n_ * (n_ - 2) * sizeof(float), &o, false));
    TW_TRY(StageOut(out_mask, static_cast<size_t>(count) * n_ * n_, &m, false));
    TW_CUDA(launch_observation(rec(first), count, n_, static_cast<float*>(o.dev), static_cast<uint8_t*>(m.dev),
                               stream_));
    launches_ += 1;
    TW_TRY(Finish(&o));
    TW_TRY(Finish(&m));
    return SyncIfHost(o, m);
  }

  int Replay(int64_t first, int64_t count, const int32_t* actions, int64_t stride, const int32_t* lengths,
             int32_t* out_applied) {
    TW_TRY(CheckRange(first, count));
    if (stride < 0) return fail(TWIXT_EINVAL, "stride must be >= 0");
    if (count == 0) return TWIXT_OK;
    if (actions == nullptr && stride > 0) return fail(TWIXT_EINVAL, "null actions pointer");
    DeviceGuard g(device_);
    ScratchReset();
    Staged in, len, ap;
    TW_TRY(StageIn(actions, static_cast<size_t>(count) * stride * sizeof(int32_t), &in));
    TW_TRY(StageIn(lengths, static_cast<size_t>(count) * sizeof(int32_t), &len));
    TW_TRY(StageOut(out_applied, static_cast<size_t>(count) * sizeof(int32_t), &ap, false));
    TW_CUDA(cudaMemsetAsync(&d_stats_->replay_illegal, 0xFF, sizeof(unsigned long long), stream_));
    TW_CUDA(launch_replay(rec(first), count, n_, static_cast<const int32_t*>(in.dev), stride,
                          static_cast<const int32_t*>(len.dev), static_cast<int32_t*>(ap.dev), d_stats_, stream_));
    launches_ += 1;
    TW_TRY(Finish(&ap));
    if (out_applied != nullptr && !ap.host && !in.host && !len.host) return TWIXT_OK;  // asynchronous: caller inspects out_applied
    unsigned long long bad = ~0ull;
    TW_CUDA(cudaMemcpyAsync(&bad, &d_stats_->replay_illegal, sizeof(bad), cudaMemcpyDeviceToHost, stream_));
    TW_TRY(Sync());
    if (bad != ~0ull)  // twixt.h:96, same text
      return fail(TWIXT_EILLEGAL, "Not a legal action: %d", static_cast<int>(static_cast<int32_t>(bad & 0xFFFFFFFFull)));
    return TWIXT_OK;
  }

  int Playout(int64_t first, int64_t count, int32_t max_plies, const uint64_t* stream_ids, float* out_returns,
              int32_t* out_lengths, uint16_t* out_actions, int32_t trace_plies) {
    TW_TRY(CheckRange(first, count));
    if (max_plies < 0) return fail(TWIXT_EINVAL, "max_plies must be >= 0");
    if (out_actions != nullptr && trace_plies <= 0) return fail(TWIXT_EINVAL, "trace_plies must be positive with out_actions");
    if (count == 0) return TWIXT_OK;
    DeviceGuard g(device_);
    ScratchReset();
    Staged ids, ret, len, act;
    TW_TRY(StageIn(stream_ids, static_cast<size_t>(count) * sizeof(uint64_t), &ids));
    TW_TRY(StageOut(out_returns, static_cast<size_t>(count) * 2 * sizeof(float), &ret, false));
    TW_TRY(StageOut(out_lengths, static_cast<size_t>(count) * sizeof(int32_t), &len, false));
    const size_t trace_bytes = out_actions ? static_cast<size_t>(trace_plies) * count * sizeof(uint16_t) : 0;
    TW_TRY(StageOut(out_actions, trace_bytes, &act, false));
    if (act.dev != nullptr) TW_CUDA(cudaMemsetAsync(act.dev, 0xFF, trace_bytes, stream_));
    PlayoutArgs a;
    a.records = rec(first);
    a.count = count;
    a.n = n_;
    a.max_plies = max_plies;
    a.seed = seed_;
    a.stream_base = stream_base_ + static_cast<uint64_t>(first);
    a.stream_ids = static_cast<const uint64_t*>(ids.dev);
    a.out_returns = static_cast<float*>(ret.dev);
    a.out_lengths = static_cast<int32_t*>(len.dev);
    a.out_actions = static_cast<uint16_t*>(act.dev);
    a.trace_plies = out_actions ? trace_plies : 0;
    a.stats = d_stats_;
    a.tickets = &d_stats_->tickets;
    TW_CUDA(cudaMemsetAsync(&d_stats_->tickets, 0, sizeof(unsigned long long), stream_));
    TW_CUDA(launch_playout(a, stream_));
    launches_ += 1;
    TW_TRY(Finish(&ret));
    TW_TRY(Finish(&len));
    TW_TRY(Finish(&act));
    if (ids.host || ret.host || len.host || act.host) TW_TRY(Sync());
    return TWIXT_OK;
  }

  int Export(int64_t first, int64_t count, uint32_t* out) {
    TW_TRY(CheckRange(first, count));
    if (count == 0) return TWIXT_OK;
    if (out == nullptr) return fail(TWIXT_EINVAL, "null output pointer");
    DeviceGuard g(device_);
    ScratchReset();
    const bool dev = is_device_pointer(out);
    TW_CUDA(cudaMemcpyAsync(out, rec(first), static_cast<size_t>(count) * rw_ * sizeof(uint32_t), cudaMemcpyDefault,
                            stream_));
    if (!dev) TW_TRY(Sync());
    return TWIXT_OK;
  }

  // Records from the caller are validated BEFORE they replace any env (validate_kernel,
  // twixt_kernels_api.cu): a host array is staged in device scratch and checked there, a device array is
  // checked where it lies; only then are the records copied in.  On failure the batch is untouched.
  int Import(int64_t first, int64_t count, const uint32_t* in) {
    TW_TRY(CheckRange(first, count));
    if (count == 0) return TWIXT_OK;
    if (in == nullptr) return fail(TWIXT_EINVAL, "null input pointer");
    DeviceGuard g(device_);
    const size_t bytes = static_cast<size_t>(count) * rw_ * sizeof(uint32_t);
    ScratchReset();
    if (!validate_) {  // twixt_set_validation(b, 0): trusted records (e.g. our own exports)
      const bool dev = is_device_pointer(in);
      TW_CUDA(cudaMemcpyAsync(rec(first), in, bytes, cudaMemcpyDefault, stream_));
      if (!dev) TW_TRY(Sync());
      return TWIXT_OK;
    }
    Staged src;
    TW_TRY(StageIn(in, bytes, &src));
    TW_CUDA(cudaMemsetAsync(&d_stats_->invalid_code, 0xFF, sizeof(unsigned long long), stream_));
    TW_CUDA(launch_validate(static_cast<const uint32_t*>(src.dev), count, n_, d_stats_, stream_));
    launches_ += 1;
    unsigned long long code = ~0ull;
    TW_CUDA(cudaMemcpyAsync(&code, &d_stats_->invalid_code, sizeof(code), cudaMemcpyDeviceToHost, stream_));
    TW_TRY(Sync());
    if (code != ~0ull)
      return fail(TWIXT_EINVAL, "invalid state record at index %lld: %s", static_cast<long long>(code >> 8),
                  invalid_reason_text(static_cast<unsigned>(code & 0xFFull)));
    TW_CUDA(cudaMemcpyAsync(rec(first), src.dev, bytes, cudaMemcpyDeviceToDevice, stream_));
    if (src.host) TW_TRY(Sync());
    return TWIXT_OK;
  }
  void SetValidation(bool on) { validate_ = on; }

  int GetStats(twixt_stats* out) {
    if (out == nullptr) return fail(TWIXT_EINVAL, "null output pointer");
    DeviceGuard g(device_);
    ScratchReset();
    DeviceStats h;
    TW_CUDA(cudaMemcpyAsync(&h, d_stats_, sizeof(h), cudaMemcpyDeviceToHost, stream_));
    TW_TRY(Sync());
    out->plies = static_cast<int64_t>(h.plies);
    out->games = static_cast<int64_t>(h.games);
    out->red_wins = static_cast<int64_t>(h.red_wins);
    out->blue_wins = static_cast<int64_t>(h.blue_wins);
    out->draws = static_cast<int64_t>(h.draws);
    out->swaps = static_cast<int64_t>(h.swaps);
    out->max_length = static_cast<int64_t>(h.max_length);
    out->kernel_launches = launches_;
    out->debug_violations = static_cast<int64_t>(h.bounds_violations);
    return TWIXT_OK;
  }

  int ResetStats() {
    DeviceGuard g(device_);
    TW_CUDA(cudaMemsetAsync(d_stats_, 0, sizeof(DeviceStats), stream_));
    return TWIXT_OK;
  }

 private:
  uint32_t* rec(int64_t env) const { return records_ + env * rw_; }

  // ---- device staging for host-side buffers: a few grow-only slots ------
  void ScratchReset() {
    next_slot_ = 0;
    pin_off_ = 0;
    num_pending_ = 0;
  }

  // Small host-side transfers (the count = 1 calls of a drop-in adapter: a status word, a 4 KB legal list, a
  // 25 KB tensor) go through a pinned, device-mapped arena instead of a device slot + cudaMemcpyAsync: the
  // kernel reads / writes the host memory directly and the call costs one launch and one stream
  // synchronisation.  Returns false when the request does not fit (the caller then uses a device slot).
  bool PinAlloc(size_t bytes, Staged* s) {
    const size_t need = (bytes + 255) & ~static_cast<size_t>(255);
    if (pinned_ == nullptr || bytes > kPinMaxTransfer || pin_off_ + need > kPinBytes || num_pending_ >= kMaxPending)
      return false;
    s->pin = static_cast<char*>(pinned_) + pin_off_;
    s->dev = static_cast<char*>(pinned_dev_) + pin_off_;
    pin_off_ += need;
    return true;
  }

  // stream synchronisation + delivery of what kernels wrote into the pinned arena
  int Sync() {
    TW_CUDA(cudaStreamSynchronize(stream_));
    for (int i = 0; i < num_pending_; ++i) memcpy(pending_[i].user, pending_[i].pin, pending_[i].bytes);
    num_pending_ = 0;
    return TWIXT_OK;
  }

  int SlotAlloc(size_t bytes, void** out) {
    if (next_slot_ >= kSlots) return fail(TWIXT_ENOMEM, "internal: out of staging slots");
    const int k = next_slot_++;
    if (bytes > slot_cap_[k]) {
      TW_CUDA(cudaStreamSynchronize(stream_));
      if (slot_[k] != nullptr) cudaFree(slot_[k]);
      slot_[k] = nullptr;
      slot_cap_[k] = 0;
      const size_t want = bytes + (bytes >> 2) + 4096;
      cudaError_t e = cudaMalloc(&slot_[k], want);
      if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(TWIXT_ENOMEM, "cudaMalloc of %zu staging bytes failed: %s", want, cudaGetErrorString(e));
      }
      slot_cap_[k] = want;
    }
    *out = slot_[k];
    return TWIXT_OK;
  }

  int StageIn(const void* user, size_t bytes, Staged* s) {
    s->user = const_cast<void*>(user);
    s->bytes = bytes;
    if (user == nullptr || bytes == 0) { s->dev = nullptr; s->host = false; return TWIXT_OK; }
    if (is_device_pointer(user)) { s->dev = s->user; s->host = false; return TWIXT_OK; }
    s->host = true;
    if (PinAlloc(bytes, s)) {
      memcpy(s->pin, user, bytes);  // visible to the kernel: written before its launch
      return TWIXT_OK;
    }
    TW_TRY(SlotAlloc(bytes, &s->dev));
    TW_CUDA(cudaMemcpyAsync(s->dev, user, bytes, cudaMemcpyHostToDevice, stream_));
    return TWIXT_OK;
  }

  int StageOut(void* user, size_t bytes, Staged* s, bool preserve) {
    s->user = user;
    s->bytes = bytes;
    if (user == nullptr || bytes == 0) { s->dev = nullptr; s->host = false; return TWIXT_OK; }
    if (is_device_pointer(user)) { s->dev = user; s->host = false; return TWIXT_OK; }
    s->host = true;
    if (PinAlloc(bytes, s)) {
      if (preserve) memcpy(s->pin, user, bytes);
      return TWIXT_OK;
    }
    TW_TRY(SlotAlloc(bytes, &s->dev));
    if (preserve) TW_CUDA(cudaMemcpyAsync(s->dev, user, bytes, cudaMemcpyHostToDevice, stream_));
    return TWIXT_OK;
  }

  int Finish(Staged* s) {
    if (s->host && s->pin != nullptr) {
      pending_[num_pending_++] = {s->user, s->pin, s->bytes};  // copied out by Sync()
      return TWIXT_OK;
    }
    if (s->host && s->dev != nullptr)
      TW_CUDA(cudaMemcpyAsync(s->user, s->dev, s->bytes, cudaMemcpyDeviceToHost, stream_));
    return TWIXT_OK;
  }

  int SyncIfHost(const Staged& a, const Staged& b) {
    if (a.host || b.host) TW_TRY(Sync());
    return TWIXT_OK;
  }

  int n_ = 0;
  int rw_ = 0;
  int device_ = 0;
  int64_t num_envs_ = 0;
  uint64_t seed_ = 0;
  uint64_t stream_base_ = 0;
  bool validate_ = true;
  uint32_t* records_ = nullptr;
  DeviceStats* d_stats_ = nullptr;
  cudaStream_t stream_ = nullptr;
  bool own_stream_ = false;
  // pinned, device-mapped arena for small host transfers
  static constexpr size_t kPinBytes = 256 * 1024, kPinMaxTransfer = 96 * 1024;
  static constexpr int kMaxPending = 8;
  struct Pending {
    void* user;
    void* pin;
    size_t bytes;
  };
  void* pinned_ = nullptr;
  void* pinned_dev_ = nullptr;
  size_t pin_off_ = 0;
  Pending pending_[kMaxPending];
  int num_pending_ = 0;
  static constexpr int kSlots = 4;
  void* slot_[kSlots] = {nullptr, nullptr, nullptr, nullptr};
  size_t slot_cap_[kSlots] = {0, 0, 0, 0};
  int next_slot_ = 0;
  int64_t launches_ = 0;
};

}  // namespace twixt

struct twixt_batch {
  twixt::TwixTBatch impl;
};

using twixt::fail;

extern "C" {

const char* twixt_last_error(void) { return twixt::g_last_error.c_str(); }
const char* twixt_version(void) { return "twixt_b200 0.1 (sm_100a)"; }

int twixt_game_info_for(int board_size, twixt_game_info* out) { return twixt::fill_game_info(board_size, out); }

int twixt_create(int board_size, int64_t num_envs, int device, uint64_t seed, twixt_batch** out) {
  if (out == nullptr) return fail(TWIXT_EINVAL, "null output pointer");
  *out = nullptr;
  twixt_batch* b = new (std::nothrow) twixt_batch();
  if (b == nullptr) return fail(TWIXT_ENOMEM, "out of host memory");
  int rc = b->impl.Init(board_size, num_envs, device, seed);
  if (rc != TWIXT_OK) {
    std::string keep = twixt::g_last_error;
    delete b;
    twixt::g_last_error = keep;
    return rc;
  }
  *out = b;
  return TWIXT_OK;
}

void twixt_destroy(twixt_batch* b) { delete b; }

#define TW_NEED(b) \
  if ((b) == nullptr) return fail(TWIXT_EINVAL, "null batch handle")

int twixt_get_info(const twixt_batch* b, twixt_game_info* out) {
  TW_NEED(b);
  return twixt::fill_game_info(b->impl.n(), out);
}
int64_t twixt_num_envs(const twixt_batch* b) { return b ? b->impl.num_envs() : 0; }
int twixt_set_stream(twixt_batch* b, uintptr_t s) {
  TW_NEED(b);
  return b->impl.SetStream(reinterpret_cast<cudaStream_t>(s));
}
uintptr_t twixt_get_stream(const twixt_batch* b) { return b ? reinterpret_cast<uintptr_t>(b->impl.stream()) : 0; }
int twixt_synchronize(twixt_batch* b) {
  TW_NEED(b);
  return b->impl.Synchronize();
}
int twixt_set_seed(twixt_batch* b, uint64_t seed) {
  TW_NEED(b);
  b->impl.SetSeed(seed);
  return TWIXT_OK;
}
int twixt_set_stream_base(twixt_batch* b, uint64_t base) {
  TW_NEED(b);
  b->impl.SetStreamBase(base);
  return TWIXT_OK;
}
int twixt_reset(twixt_batch* b, int64_t first, int64_t count) {
  TW_NEED(b);
  return b->impl.Reset(first, count);
}
int twixt_clone(twixt_batch* b, int64_t src_first, int64_t dst_first, int64_t count) {
  TW_NEED(b);
  return b->impl.Clone(src_first, dst_first, count);
}
int twixt_clone_gather(twixt_batch* b, const int64_t* src_ids, int64_t dst_first, int64_t count) {
  TW_NEED(b);
  return b->impl.CloneGather(src_ids, dst_first, count);
}
int twixt_clone_from(twixt_batch* dst, int64_t dst_first, const twixt_batch* src, int64_t src_first, int64_t count) {
  TW_NEED(dst);
  TW_NEED(src);
  return dst->impl.CloneFrom(dst_first, src->impl, src_first, count);
}
int twixt_legal_actions(twixt_batch* b, int64_t first, int64_t count, void* out_actions, int32_t elem_bytes,
                        int64_t stride, int32_t* out_counts) {
  TW_NEED(b);
  return b->impl.LegalActions(first, count, out_actions, elem_bytes, stride, out_counts);
}
int twixt_legal_mask(twixt_batch* b, int64_t first, int64_t count, uint8_t* out_mask) {
  TW_NEED(b);
  return b->impl.LegalMask(first, count, out_mask);
}
int twixt_apply(twixt_batch* b, int64_t first, int64_t count, const int32_t* actions, int32_t* out_status) {
  TW_NEED(b);
  return b->impl.Apply(first, count, actions, out_status);
}
int twixt_current_player(twixt_batch* b, int64_t first, int64_t count, int8_t* out) {
  TW_NEED(b);
  return b->impl.Query(first, count, out, nullptr, nullptr);
}
int twixt_is_terminal(twixt_batch* b, int64_t first, int64_t count, uint8_t* out) {
  TW_NEED(b);
  return b->impl.Query(first, count, nullptr, out, nullptr);
}
int twixt_returns(twixt_batch* b, int64_t first, int64_t count, float* out) {
  TW_NEED(b);
  return b->impl.Query(first, count, nullptr, nullptr, out);
}
int twixt_observation(twixt_batch* b, int64_t first, int64_t count, float* out) {
  TW_NEED(b);
  return b->impl.Observation(first, count, out, nullptr);
}
int twixt_observation_and_mask(twixt_batch* b, int64_t first, int64_t count, float* out_obs, uint8_t* out_mask) {
  TW_NEED(b);
  if (out_mask == nullptr) return fail(TWIXT_EINVAL, "null output pointer");
  return b->impl.Observation(first, count, out_obs, out_mask);
}
int twixt_replay(twixt_batch* b, int64_t first, int64_t count, const int32_t* actions, int64_t stride,
                 const int32_t* lengths, int32_t* out_applied) {
  TW_NEED(b);
  return b->impl.Replay(first, count, actions, stride, lengths, out_applied);
}
int twixt_step(twixt_batch* b, int64_t env, int32_t action, twixt_step_result* out, int64_t* out_legal) {
  TW_NEED(b);
  return b->impl.Step(env, action, out, out_legal);
}
int twixt_set_validation(twixt_batch* b, int enabled) {
  TW_NEED(b);
  b->impl.SetValidation(enabled != 0);
  return TWIXT_OK;
}
int twixt_playout(twixt_batch* b, int64_t first, int64_t count, int32_t max_plies, const uint64_t* stream_ids,
                  float* out_returns, int32_t* out_lengths, uint16_t* out_actions, int32_t trace_plies) {
  TW_NEED(b);
  return b->impl.Playout(first, count, max_plies, stream_ids, out_returns, out_lengths, out_actions, trace_plies);
}
int twixt_export_state(twixt_batch* b, int64_t first, int64_t count, uint32_t* out_records) {
  TW_NEED(b);
  return b->impl.Export(first, count, out_records);
}
int twixt_import_state(twixt_batch* b, int64_t first, int64_t count, const uint32_t* records) {
  TW_NEED(b);
  return b->impl.Import(first, count, records);
}
// ---- multi-GPU helpers for a host that is not Python (twixt_for_open_spiel_b200/sharding.py restated) ----
int twixt_shard_range(int64_t global_envs, int32_t world, int32_t rank, int64_t* out_first, int64_t* out_count) {
  if (world < 1 || rank < 0 || rank >= world || global_envs < 0)
    return fail(TWIXT_EINVAL, "bad shard arguments: %lld envs, world %d, rank %d", static_cast<long long>(global_envs),
                world, rank);
  if (out_first == nullptr || out_count == nullptr) return fail(TWIXT_EINVAL, "null output pointer");
  const int64_t base = global_envs / world, extra = global_envs % world;
  *out_first = rank * base + (rank < extra ? rank : extra);
  *out_count = base + (rank < extra ? 1 : 0);
  return TWIXT_OK;
}
int twixt_stats_accumulate(twixt_stats* acc, const twixt_stats* part) {
  if (acc == nullptr || part == nullptr) return fail(TWIXT_EINVAL, "null pointer");
  acc->plies += part->plies;
  acc->games += part->games;
  acc->red_wins += part->red_wins;
  acc->blue_wins += part->blue_wins;
  acc->draws += part->draws;
  acc->swaps += part->swaps;
  acc->kernel_launches += part->kernel_launches;
  if (part->max_length > acc->max_length) acc->max_length = part->max_length;
  return TWIXT_OK;
}

int twixt_get_stats(twixt_batch* b, twixt_stats* out) {
  TW_NEED(b);
  return b->impl.GetStats(out);
}
int twixt_stats_reset(twixt_batch* b) {
  TW_NEED(b);
  return b->impl.ResetStats();
}

}  // extern "C"
