// Shared pieces of the host emulations of CUDA-core kernels (TEST INFRASTRUCTURE): CUDA intrinsics with one rounding
// as the plain operation (build with -ffp-contract=off), and the thread context - one of 256 OS threads per CTA,
// __syncthreads = pthread barrier, cp.async = a 16-byte copy at issue time.
#pragma once
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#include <thread>
#include <vector>

static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __frcp_rn(float a) { return 1.0f / a; }

namespace emul {
struct HostCtx {
  int tid, block, nblocks;
  pthread_barrier_t* bar;
#ifdef NEUS_EMUL_DROP_BARRIER  // negative control of the sanitizer tests: every thread skips its N-th __syncthreads
  int n_sync = 0;
  void sync() {
    if (++n_sync != NEUS_EMUL_DROP_BARRIER) pthread_barrier_wait(bar);
  }
#else
  void sync() { pthread_barrier_wait(bar); }
#endif
  void cp16(void* dst, const void* src) { memcpy(dst, src, 16); }
  void cp_commit() {}
  void cp_wait_1() {}
  void cp_wait_0() {}
};

// One grid: CTAs run one after the other, each on `threads` OS threads over a NaN-filled block of "shared memory"
// (uninitialised shared memory must never be consumed); `shift` floats of misalignment for the negative control.
template <class Body>
inline void run_grid(int nblocks, int threads, size_t smem_floats, int shift, Body body) {
  for (int blk = 0; blk < nblocks; ++blk) {
    float* smem = (float*)aligned_alloc(64, (smem_floats * sizeof(float) + 63) / 64 * 64 + 64);
    for (size_t i = 0; i < smem_floats + 16; ++i) smem[i] = NAN;
    pthread_barrier_t bar;
    pthread_barrier_init(&bar, nullptr, threads);
    std::vector<std::thread> th;
    for (int t = 0; t < threads; ++t)
      th.emplace_back([&, t] {
        HostCtx cx{t, blk, nblocks, &bar};
        body(cx, smem + shift);
      });
    for (auto& x : th) x.join();
    pthread_barrier_destroy(&bar);
    free(smem);
  }
}
}  // namespace emul
