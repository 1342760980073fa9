// Hardware probes for the CTA-pair (cta_group::2) field kernel: they pin the conventions the kernel
// relies on and measure the two rates its design rests on.
//
//   neddf_tc_pair_selftest : C[256, n] = A[256, k] * B[n, k]^T with one tcgen05.mma.cta_group::2 per
//       K-step and product (A = "weights", fp16 hi/lo, written to each CTA's tensor memory with
//       tcgen05.st: CTA r owns rows 128r..128r+127; B = "activations", fp16 hi/lo, MN-major in shared
//       memory: CTA r owns rows (n/2)r..(n/2)(r+1)-1).  Checks the operand split between the two CTAs,
//       the multicast commit and the accumulator layout; also times `reps` repetitions.
//   neddf_dsmem_bench      : 16-byte st.shared::cluster stores into the peer CTA's shared memory
//       (the activation exchange of the pair kernel), bytes per cycle.
#include <cstdlib>

#include "tc_ptx.cuh"

namespace neddf {
namespace tc {

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
    tc_pair_selftest_kernel(const float* __restrict__ A, const float* __restrict__ B, int n, int k,
                            float* __restrict__ C, long long* __restrict__ cyc, int reps, int kc, int row0, int rmode, int c8, int boff) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* h_hi = smem + boff;
  unsigned char* h_lo = smem + boff + 32768;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const int nh = n / 2;  // B rows held by this CTA
  for (int idx = tid; idx < nh * k; idx += 128) {
    int kk = idx % k, r = idx / k;
    float b = B[(size_t)(nh * rank + r) * k + kk];
    __half hi = __float2half_rn(b), lo = __float2half_rn(b - __half2float(hi));
    *reinterpret_cast<__half*>(h_hi + act_off(row0 + r, kk, kc)) = hi;
    *reinterpret_cast<__half*>(h_lo + act_off(row0 + r, kk, kc)) = lo;
  }
  if (tid == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc2(&tmem_base, 512);
  fence_async_smem();
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = tmem_base;
  // A: lane = row m of this CTA's half; per K-step 8 columns hi at 256 + ks*16, 8 columns lo after
  const int m = 32 * warp + lane;
  for (int ks = 0; ks < k / 16; ++ks) {
    uint32_t hi[8], lo[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float a0 = A[(size_t)(128 * rank + m) * k + ks * 16 + 2 * j], a1 = A[(size_t)(128 * rank + m) * k + ks * 16 + 2 * j + 1];
      float amax = 0.f;
      split2(a0, a1, hi[j], lo[j], amax);
    }
    const uint32_t ta = tmem + ((uint32_t)(32 * warp) << 16) + 256 + ks * 16;
    tmem_st8(ta, hi);
    tmem_st8(ta + 8, lo);
  }
  tmem_st_wait();
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  if (warp == 0 && rank == 0) {
    const uint32_t s_hhi = smem_u32(h_hi), s_hlo = smem_u32(h_lo);
    const uint32_t idesc = make_idesc(256, n, 0, 1);
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      for (int ks = 0; ks < k / 16; ++ks) {
        const uint32_t roff = (uint32_t)(row0 >> 3) * (kc * 16);
        const uint64_t db_hi = make_desc(s_hhi + roff + ks * 256, 128, kc * 16), db_lo = make_desc(s_hlo + roff + ks * 256, 128, kc * 16);
        const uint32_t ta = tmem + 256 + ks * 16;
        mma2_f16_ts_elect(tmem, ta, db_hi, idesc, (r | ks) > 0);
        mma2_f16_ts_elect(tmem, ta + 8, db_hi, idesc, 1);
        mma2_f16_ts_elect(tmem, ta, db_lo, idesc, 1);
      }
    }
    mma2_commit_elect(smem_u32(&bar), 3);
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    if (lane == 0 && cyc) cyc[0] = t1 - t0;
  } else {
    mbar_wait(&bar, 0);
  }
  __syncthreads();
  tc_fence_after();
  if (rmode == 0) {
    for (int cb = 0; cb < n / 16; ++cb) {
      float v[16];
      tmem_ld16(tmem + ((uint32_t)(32 * warp) << 16) + cb * 16, v);
#pragma unroll
      for (int i = 0; i < 16; ++i) C[(size_t)(128 * rank + m) * n + cb * 16 + i] = v[i] / (float)reps;
    }
  } else if (rmode == 1) {
    // x8 loads, compile-time column offsets inside 32-column blocks (ptxas folds them into the LDTM immediate)
    for (int cb = 0; cb < n / 32; ++cb) {
      const uint32_t base = tmem + ((uint32_t)(32 * warp) << 16) + cb * 32;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float v[8];
        tmem_ld8(base + 8 * q, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) C[(size_t)(128 * rank + m) * n + cb * 32 + 8 * q + i] = v[i] / (float)reps;
      }
    }
  } else {
    // x8 loads, run-time column offsets (c8 == 8, a kernel argument)
    for (int cb = 0; cb < n / 32; ++cb) {
      const uint32_t base = tmem + ((uint32_t)(32 * warp) << 16) + cb * 32;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float v[8];
        tmem_ld8(base + c8 * q, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) C[(size_t)(128 * rank + m) * n + cb * 32 + 8 * q + i] = v[i] / (float)reps;
      }
    }
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 0) tmem_dealloc2(tmem, 512);
}

// mode 0: local st.shared.v4; 1: both CTAs store into the peer; 2: only rank 0 stores into rank 1;
// 3: both store into the peer with a 512-byte-per-warp contiguous pattern (as the pair kernel's
//    epilogue does) but the warp's rows 2 KB apart (different row groups)
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(512, 1)
    dsmem_bench_kernel(int mode, int reps, int bytes, long long* __restrict__ out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const uint32_t rank = cluster_ctarank();
  const int tid = threadIdx.x;
  const uint32_t local = smem_u32(smem);
  const uint32_t peer = mapa_u32(local, rank ^ 1u);
  for (int i = tid; i < bytes / 16; i += 512) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  cluster_sync_all();
  // modes 4 / 5: 8-byte stores at 16-byte stride / only 8 of the 16 warps store (the pair kernel's epilogue patterns)
  const bool active = (mode == 0) || (mode == 1) || (mode == 3) || (mode == 2 && rank == 0) || (mode == 4) || (mode == 5 && tid < 256);
  const uint32_t base = (mode == 0) ? local : peer;
  long long t0 = clock64();
  if (active) {
    for (int r = 0; r < reps; ++r) {
      for (int off = tid * 16; off < bytes; off += 512 * 16) {
        uint32_t o = off;
        if (mode == 3) {  // permute 512-byte blocks
          uint32_t blk = off >> 9, in = off & 511;
          o = (((blk * 5) % (bytes >> 9)) << 9) | in;
        }
        if (mode == 4) st_cluster_v2(base + o, make_uint2(r, tid));
        else st_cluster_v4(base + o, make_uint4(r, tid, off, 1));
      }
    }
  }
  asm volatile("fence.acq_rel.cluster;" ::: "memory");
  cluster_sync_all();
  long long t1 = clock64();
  if (tid == 0) out[blockIdx.x] = t1 - t0;
  // keep the stores observable
  if (tid == 0 && reinterpret_cast<volatile uint32_t*>(smem)[3] == 0xdeadbeefu) out[blockIdx.x] = -1;
}

// tcgen05.cp layout probe: shared memory holds the 16-bit pattern value[i] = i (i = half index); one
// 128x256b copy with the given descriptor strides lands in tensor-memory columns 0..7; out[lane][col] =
// the 32-bit column value (two half indices), so the host can read off which bytes went where.
__global__ void __launch_bounds__(128, 1) tc_cp_probe_kernel(int lbo, int sbo, uint32_t* __restrict__ out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 16384; i += 128) reinterpret_cast<uint16_t*>(smem)[i] = (uint16_t)i;
  if (tid == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(&tmem_base, 32);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base;
  if (tid == 0) {
    const uint64_t desc = make_desc(smem_u32(smem), (uint32_t)lbo, (uint32_t)sbo);
    asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(tmem), "l"(desc) : "memory");
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  float v[8];
  tmem_ld8(tmem + ((uint32_t)(32 * warp) << 16), v);
#pragma unroll
  for (int i = 0; i < 8; ++i) out[tid * 8 + i] = __float_as_uint(v[i]);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 32);
}

}  // namespace tc
}  // namespace neddf

using namespace neddf;

extern "C" int32_t neddf_tc_cp_probe(int32_t lbo, int32_t sbo, uint32_t* d_out, void* stream) {
  if (!d_out || lbo < 0 || sbo < 0 || (lbo % 16) || (sbo % 16)) return fail(NEDDF_E_INVALID, "neddf_tc_cp_probe: bad arguments");
  NEDDF_CUDA_CHECK(cudaFuncSetAttribute(tc::tc_cp_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768));
  tc::tc_cp_probe_kernel<<<1, 128, 32768, (cudaStream_t)stream>>>(lbo, sbo, d_out);
  NEDDF_LAUNCH_CHECK();
  return NEDDF_OK;
}


extern "C" int32_t neddf_tc_pair_selftest(const float* d_a, const float* d_b, int32_t n, int32_t k, float* d_c,
                                          int64_t* d_cycles, int32_t reps, void* stream) {
  // probe variants: NEDDF_PAIR_KC = K capacity of the B buffer (row-group stride kc*16 bytes),
  // NEDDF_PAIR_ROW0 = first row of the B tile inside the buffer (multiple of 8)
  int kc = tc::kHK, row0 = 0;
  if (const char* e = std::getenv("NEDDF_PAIR_KC")) kc = std::atoi(e);
  if (const char* e = std::getenv("NEDDF_PAIR_ROW0")) row0 = std::atoi(e);
  int rmode = 0;  // NEDDF_PAIR_READ: 0 = x16 loads, 1 = x8 loads with immediate column offsets, 2 = x8 with run-time offsets
  if (const char* e = std::getenv("NEDDF_PAIR_READ")) rmode = std::atoi(e);
  int boff = 0;  // NEDDF_PAIR_BASE: byte offset of the B buffers inside the dynamic shared memory (hi, lo 32 KB apart)
  if (const char* e = std::getenv("NEDDF_PAIR_BASE")) boff = std::atoi(e);
  if (boff < 0 || boff > 131072 || (boff % 1024) != 0 || (size_t)(row0 + n / 2 + 7) / 8 * kc * 16 > 32768)
    return fail(NEDDF_E_INVALID, "neddf_tc_pair_selftest: bad NEDDF_PAIR_BASE / tile does not fit 32 KB");
  if (kc < k || kc > 256 || (row0 % 8) != 0 || row0 + n / 2 > 128) return fail(NEDDF_E_INVALID, "neddf_tc_pair_selftest: bad probe variant");
  if (k < 16 || k > 256 || (k % 16) != 0 || reps < 1 || n < 32 || n > 256 || (n % 32) != 0)
    return fail(NEDDF_E_INVALID, "neddf_tc_pair_selftest: need k in [16,256] (multiple of 16), n in [32,256] (multiple of 32)");
  if (!d_a || !d_b || !d_c) return fail(NEDDF_E_INVALID, "neddf_tc_pair_selftest: NULL pointer");
  size_t smem = 131072 + 65536;
  NEDDF_CUDA_CHECK(cudaFuncSetAttribute(tc::tc_pair_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  tc::tc_pair_selftest_kernel<<<2, 128, smem, (cudaStream_t)stream>>>(d_a, d_b, n, k, d_c, reinterpret_cast<long long*>(d_cycles), reps, kc, row0, rmode, 8, boff);
  NEDDF_LAUNCH_CHECK();
  return NEDDF_OK;
}

extern "C" int32_t neddf_dsmem_bench(int32_t mode, int32_t reps, int32_t bytes, int32_t n_clusters, int64_t* d_cycles,
                                     void* stream) {
  if (mode < 0 || mode > 5 || reps < 1 || bytes < 8192 || bytes > 196608 || (bytes % 8192) != 0 || n_clusters < 1 || !d_cycles)
    return fail(NEDDF_E_INVALID, "neddf_dsmem_bench: bad arguments");
  NEDDF_CUDA_CHECK(cudaFuncSetAttribute(tc::dsmem_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  tc::dsmem_bench_kernel<<<2 * n_clusters, 512, bytes, (cudaStream_t)stream>>>(mode, reps, bytes, reinterpret_cast<long long*>(d_cycles));
  NEDDF_LAUNCH_CHECK();
  return NEDDF_OK;
}
