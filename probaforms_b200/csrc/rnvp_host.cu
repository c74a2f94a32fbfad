// Host side of ingestion / egress (no device code in this file): what the reference does with
// torch.tensor(X, dtype=torch.float32) (realnvp.py:226-228) and .cpu().detach().numpy() (realnvp.py:281) once the
// kernels are fast enough for those copies to be the wall-clock bound.
//
// rnvp_host_gather_rows: dst[r][:] = (float) src[idx ? idx[r] : row0 + r][:] -- the rows of one optimisation step (a
// slice of the epoch permutation), converted from the caller's float64 / float32 numpy array straight into a pinned
// staging buffer, split over a few host threads.  RealNVP.fit streams these buffers to the GPU one step ahead of
// the kernels, so every rank uploads only the rows of its own shard.
// rnvp_host_copy: multi-threaded memcpy (first-touch of a fresh numpy result array is page-fault bound on one thread).
#include <stdint.h>
#include <string.h>
#include <algorithm>
#include <thread>
#include <vector>
#include "../../include/rnvp.h"

namespace {

template <typename T>
void gather_range(const T* src, int64_t width, const int64_t* idx, int64_t row0, int64_t r0, int64_t r1, float* dst) {
  constexpr int AHEAD = 16;                      // random rows: prefetch every cache line of the row 16 rows ahead
  const int64_t row_bytes = width * (int64_t)sizeof(T);
  for (int64_t r = r0; r < r1; ++r) {
    if (idx && r + AHEAD < r1) {
      const char* nx = (const char*)(src + idx[r + AHEAD] * width);
      for (int64_t b = 0; b < row_bytes; b += 64) __builtin_prefetch(nx + b);
    }
    const T* s = src + (idx ? idx[r] : row0 + r) * width;
    float* d = dst + r * width;
    for (int64_t j = 0; j < width; ++j) d[j] = (float)s[j];
  }
}

template <typename F>
void run_threads(int64_t n, int64_t min_per_thread, int threads, F f) {
  int t = (int)std::max<int64_t>(1, std::min<int64_t>(threads, n / std::max<int64_t>(min_per_thread, 1)));
  if (t <= 1) { f(0, n); return; }
  std::vector<std::thread> pool;
  const int64_t per = (n + t - 1) / t;
  for (int i = 1; i < t; ++i) pool.emplace_back(f, std::min(n, i * per), std::min(n, (i + 1) * per));
  f(0, std::min(n, per));
  for (auto& th : pool) th.join();
}

}  // namespace

extern "C" {

int rnvp_host_gather_rows(const void* src, int src_is_f64, int64_t width, const int64_t* idx, int64_t row0, int64_t n,
                          float* dst, int threads) {
  if (!src || !dst || width < 1 || n < 0 || row0 < 0) return RNVP_EINVAL;
  auto work = [=](int64_t r0, int64_t r1) {
    if (src_is_f64) gather_range((const double*)src, width, idx, row0, r0, r1, dst);
    else gather_range((const float*)src, width, idx, row0, r0, r1, dst);
  };
  run_threads(n, std::max<int64_t>(1, 65536 / width), threads, work);
  return 0;
}

int rnvp_host_copy(void* dst, const void* src, int64_t bytes, int threads) {
  if (!src || !dst || bytes < 0) return RNVP_EINVAL;
  auto work = [=](int64_t b0, int64_t b1) { memcpy((char*)dst + b0, (const char*)src + b0, (size_t)(b1 - b0)); };
  run_threads(bytes, 1 << 20, threads, work);
  return 0;
}

}  // extern "C"
