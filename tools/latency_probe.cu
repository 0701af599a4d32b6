// tools/latency_probe.cu -- DEVELOPMENT TOOL: where do the microseconds of a single-state call go?
// Times, on the current device: an empty kernel launch + stream synchronise (the floor of any count = 1 call),
// cudaPointerGetAttributes on a pageable host pointer, and the library's twixt_is_terminal / twixt_step /
// twixt_observation / twixt_clone with host buffers.
//   nvcc -O2 -gencode arch=compute_100a,code=sm_100a -I include tools/latency_probe.cu -o /tmp/latency_probe \
//        -L twixt_for_open_spiel_b200 -ltwixt_b200 -Xlinker -rpath=$PWD/twixt_for_open_spiel_b200
#include <chrono>
#include <cstdio>
#include <vector>

#include <cuda_runtime.h>

#include "twixt_b200.h"

__global__ void empty_kernel(int* p) {
  if (p != nullptr && threadIdx.x == 1000) *p = 1;
}

template <class F>
double per_call_us(F f, int reps = 2000) {
  for (int i = 0; i < 50; ++i) f();
  auto t0 = std::chrono::steady_clock::now();
  for (int i = 0; i < reps; ++i) f();
  return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count() / reps;
}

int main() {
  twixt_batch* b = nullptr;
  if (twixt_create(24, 8, 0, 1, &b) != TWIXT_OK) {
    std::printf("create failed: %s\n", twixt_last_error());
    return 1;
  }
  twixt_playout(b, 0, 8, 150, nullptr, nullptr, nullptr, nullptr, 0);
  twixt_synchronize(b);
  cudaStream_t s;
  cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
  std::vector<int64_t> legal(528);
  std::vector<float> obs(12 * 24 * 22);
  twixt_step_result res;
  unsigned char term = 0;
  cudaPointerAttributes attr;
  std::printf("empty launch + sync            %7.2f us\n", per_call_us([&] { empty_kernel<<<1, 32, 0, s>>>(nullptr); cudaStreamSynchronize(s); }));
  std::printf("empty launch only              %7.2f us\n", per_call_us([&] { empty_kernel<<<1, 32, 0, s>>>(nullptr); }));
  cudaStreamSynchronize(s);
  std::printf("cudaPointerGetAttributes(host) %7.2f us\n", per_call_us([&] { cudaPointerGetAttributes(&attr, legal.data()); cudaGetLastError(); }));
  std::printf("twixt_is_terminal(count=1)     %7.2f us\n", per_call_us([&] { twixt_is_terminal(b, 0, 1, &term); }));
  std::printf("twixt_step(query)              %7.2f us\n", per_call_us([&] { twixt_step(b, 0, TWIXT_STEP_QUERY, &res, legal.data()); }));
  std::printf("twixt_observation(count=1)     %7.2f us\n", per_call_us([&] { twixt_observation(b, 0, 1, obs.data()); }));
  std::printf("twixt_clone(1 env, async)      %7.2f us\n", per_call_us([&] { twixt_clone(b, 0, 1, 1); }));
  twixt_synchronize(b);
  twixt_destroy(b);
  return 0;
}
