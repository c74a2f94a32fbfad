// Host-side epoch row order, produced incrementally (no device code in this file).
//
// The reference draws every epoch's batches from DataLoader(dataset, batch_size, shuffle=True) (realnvp.py:237), i.e.
// RandomSampler -> torch.randperm(n, generator=Generator().manual_seed(seed)).  On the CPU that is a sequential forward
// Fisher-Yates shuffle driven by a 32-bit Mersenne Twister (ATen/native/TensorFactories.cpp randperm_cpu, the n < 2^32/20
// branch: for i in [0, n-1): swap(r[i], r[i + random() % (n - i)])), so entry i is FINAL as soon as step i has run.
// rnvp_perm_advance exposes exactly that: the fit loop starts on the first batches while a helper thread is still
// shuffling the tail, instead of waiting tens of ns per row for the whole permutation up front.
#include <stdint.h>
#if defined(__linux__)
#include <sys/mman.h>
#endif
#include <new>
#include <random>
#include <thread>
#include <vector>
#include "../../include/rnvp.h"

constexpr int PERM_AHEAD = 32;   // swap targets are drawn this many steps ahead and prefetched: the shuffle is bound by
                                 // the latency of one random access per row (~30 ns), the draws do not depend on the data

struct rnvp_perm {
  std::mt19937 eng;        // at::mt19937(seed) == std::mt19937 seeded with the low 32 bits (same init, same tempering)
  int64_t n, next;         // entries [0, next) are final
  int64_t drawn;           // swap targets of steps [next, drawn) are already in `target`
  int64_t target[PERM_AHEAD];
  int64_t* out;
};

extern "C" {

int rnvp_perm_create(uint64_t seed, int64_t n, int64_t* out, rnvp_perm** p) {
  if (!p || (!out && n > 0) || n < 0) return RNVP_EINVAL;
  if (n >= (int64_t)(UINT32_MAX / 20)) return RNVP_ESHAPE;     // torch switches to a 64-bit rejection scheme there
  rnvp_perm* q = new (std::nothrow) rnvp_perm;
  if (!q) return RNVP_EINVAL;
  q->eng.seed((uint32_t)(seed & 0xffffffffu));
  q->n = n; q->next = 0; q->drawn = -1; q->out = out;         // drawn < 0: `out` not initialised yet (done by the first advance,
                                                               // i.e. on the helper thread, not on the caller's critical path)
  *p = q;
  return 0;
}

int64_t rnvp_perm_advance(rnvp_perm* p, int64_t upto) {
  if (!p) return RNVP_EINVAL;
  if (upto > p->n) upto = p->n;
  int64_t i = p->next;
  const int64_t n = p->n;
  int64_t* r = p->out;
  if (p->drawn < 0) {
#if defined(__linux__)
    // every swap is one random access into an 8n-byte array: with 4 KB pages each one is also a TLB miss.  Ask for
    // transparent huge pages on the whole-2MB part of the buffer before it is first touched (a hint: ignored where THP is
    // off or the buffer is pinned / already faulted in).
    if (n >= (1 << 18)) {
      const uintptr_t a0 = ((uintptr_t)r + (2u << 20) - 1) & ~(uintptr_t)((2u << 20) - 1);
      const uintptr_t a1 = ((uintptr_t)(r + n)) & ~(uintptr_t)((2u << 20) - 1);
      if (a1 > a0) madvise((void*)a0, a1 - a0, MADV_HUGEPAGE);
    }
#endif
    // identity fill (first touch of a fresh buffer: page faults): split over a few threads for large n so the first
    // batch does not wait tens of milliseconds for it
    const int nt = n >= (1 << 20) ? 8 : 1;
    if (nt == 1) {
      for (int64_t k = 0; k < n; ++k) r[k] = k;
    } else {
      std::vector<std::thread> th;
      for (int t = 0; t < nt; ++t)
        th.emplace_back([=] {
          const int64_t lo = n * t / nt, hi = n * (t + 1) / nt;
          for (int64_t k = lo; k < hi; ++k) r[k] = k;
        });
      for (auto& x : th) x.join();
    }
    p->drawn = 0;
  }
  for (; i < upto && i < n - 1; ++i) {
    while (p->drawn < i + PERM_AHEAD && p->drawn < n - 1) {      // draw ahead (same order as the reference: one per step)
      const int64_t k = p->drawn++;
      const int64_t t = k + (int64_t)((uint32_t)p->eng() % (uint32_t)(n - k));   // n < 2^32/20: 32-bit modulo
      p->target[k % PERM_AHEAD] = t;
      __builtin_prefetch(r + t, 1);
    }
    const int64_t t = p->target[i % PERM_AHEAD];
    const int64_t sav = r[i];
    r[i] = r[t];
    r[t] = sav;
  }
  if (upto >= n) i = n;                                         // the last entry needs no draw
  p->next = i > p->next ? i : p->next;
  return p->next;
}

void rnvp_perm_destroy(rnvp_perm* p) { delete p; }

}  // extern "C"
