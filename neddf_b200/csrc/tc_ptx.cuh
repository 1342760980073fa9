// tcgen05 / TMEM / mbarrier PTX wrappers, operand layouts and the split-fp16 helpers shared by the
// tensor-core field kernels (field_tc.cu: one CTA per SM; field_tc2.cu: CTA pairs, cta_group::2).
#pragma once

#include "field_math.cuh"

#include <cuda_fp16.h>

namespace neddf {
namespace tc {

constexpr int kTileS = 32;              // samples per CTA tile
constexpr int kRows = 4 * kTileS;       // 128 = MMA N (hidden) / M (heads)
constexpr int kHK = 256;                // K capacity of H
constexpr int kAuxK = 96;               // K capacity of AUX
constexpr int kChunkBytes = 8192;       // one weight chunk: 128 rows x (8 words hi | 8 words lo)
constexpr uint32_t kHBytes = kRows * kHK * 2;      // 65536 per hi / lo
constexpr uint32_t kAuxBytes = kRows * kAuxK * 2;  // 24576 per hi / lo

// ---------------------------------------------------------------------------------------------
// layouts
// ---------------------------------------------------------------------------------------------
// B/A operand, MN-major, no swizzle: element (row, k) of a buffer with K capacity KC
__host__ __device__ __forceinline__ uint32_t act_off(int row, int k, int KC) {
  return (uint32_t)((row >> 3) * (KC * 16) + k * 16 + (row & 7) * 2);
}
// weight chunk (K-major, no swizzle): [m/8][k/8][m%8][k%8] fp16, m in [0,128), k in [0,16)
__host__ __device__ __forceinline__ uint32_t wchunk_off(int m, int k) {
  return (uint32_t)((m >> 3) * 256 + (k >> 3) * 128 + (m & 7) * 16 + (k & 7) * 2);
}
// head-weight chunk (K-major): [n/8][k/8][n%8][k%8] fp16, n in [0,16), k in [0,256)
__host__ __device__ __forceinline__ uint32_t hchunk_off(int n, int k) {
  return (uint32_t)((n >> 3) * 4096 + (k >> 3) * 128 + (n & 7) * 16 + (k & 7) * 2);
}

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "TC_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra TC_DONE;\n"
      "bra TC_WAIT;\n"
      "TC_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}

// shared-memory matrix descriptor, SWIZZLE_NONE, version 1 (cute::UMMA::SmemDescriptor)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// instruction descriptor, kind::f16: fp16 x fp16 -> fp32 (cute::UMMA::InstrDescriptor)
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}
constexpr uint32_t kIdescHidden = make_idesc(128, kRows, 0, 1);      // A weights K-major, B activations MN-major
constexpr uint32_t kIdescHiddenValue = make_idesc(128, kTileS, 0, 1);  // value rows only (eval colour trunk)
constexpr uint32_t kIdescHead = make_idesc(kRows, 16, 1, 0);     // A activations MN-major, B head weights K-major

__device__ __forceinline__ void mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// A operand from tensor memory ("TS"): A[m][k] lives in lane m, 32-bit column k/2 (two fp16 per
// column, even k in the low half) - cute::UMMA::tmem_frg_1sm<half_t>.
__device__ __forceinline__ void mma_f16_ts_elect(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                                 uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p, q;\n"
      "elect.sync _|q, 0xffffffff;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}
// One weight chunk: D += A_hi*B_hi + A_lo*B_hi + A_hi*B_lo (A_hi at TMEM columns a..a+7, A_lo at a+8..),
// then tcgen05.commit to `bar` (shared-memory address) - one elected lane, one asm block.
__device__ __forceinline__ void chunk_mma_elect(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_hi, uint64_t b_lo,
                                                uint32_t idesc, uint32_t accumulate, uint32_t bar) {
  asm volatile(
      "{\n"
      ".reg .pred p, q, t;\n"
      ".reg .b32 alo;\n"
      "elect.sync _|q, 0xffffffff;\n"
      "setp.ne.b32 p, %5, 0;\n"
      "setp.ne.b32 t, %8, 0;\n"
      "add.u32 alo, %1, 8;\n"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %4, {%7, %7, %7, %7}, p;\n"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [alo], %2, %4, {%7, %7, %7, %7}, t;\n"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %3, %4, {%7, %7, %7, %7}, t;\n"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%6];\n"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_hi), "l"(b_lo), "r"(idesc), "r"(accumulate), "r"(bar), "r"(0u), "r"(1u)
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t r[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t a[8], const uint32_t b[8]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(b[0]), "r"(b[1]), "r"(b[2]),
      "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// Warp-uniform variants: the whole warp executes them with identical operands and one elected
// lane issues.  Keeping the issuing code convergent lets ptxas hold descriptors in uniform
// registers; under `if (lane == 0)` it wraps every UTCHMMA in an ELECT/R2UR.BROADCAST loop.
__device__ __forceinline__ void mma_f16_elect(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p, q;\n"
      "elect.sync _|q, 0xffffffff;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_commit_elect(uint64_t* bar) {
  asm volatile(
      "{\n"
      ".reg .pred q;\n"
      "elect.sync _|q, 0xffffffff;\n"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n"
      "}\n" ::"r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float v[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
      "tcgen05.wait::ld.sync.aligned;\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float v[8]) {
  uint32_t r[8];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
      "tcgen05.wait::ld.sync.aligned;\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float v[4]) {
  uint32_t r[4];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];\n"
      "tcgen05.wait::ld.sync.aligned;\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// tanhExp with first derivative on the SFU (MUFU.EX2 x2 + MUFU.RCP), ~20 instructions instead of
// ~55 for expf + tanhf.  e^x carries the rounding residual of x*log2(e) (relative error ~2e-7);
// tanh(u), u = e^x >= 0, is 1 - 2/(e^{2u}+1) for u >= 1/8 (absolute error ~1.2e-7) and its odd
// series below (truncation < 2e-9), so y and f' stay within a few fp32 ulp of the reference's
// libm evaluation in absolute terms.  Same masks as nn_module/with_grad/tanh_exp.py:38-45.
__device__ __forceinline__ void tanhexp_fast(float x, float& y, float& d1) {
  const float kL2E = 1.4426950408889634f, kL2E_lo = 1.9259629911266175e-8f, kLn2 = 0.6931471805599453f;
  float t = x * kL2E;
  float r = fmaf(x, kL2E, -t) + x * kL2E_lo;  // what rounding t dropped
  float ex = ex2_approx(t);
  ex = fmaf(ex, r * kLn2, ex);
  float E = ex2_approx(ex * (2.0f * kL2E));
  float tx_big = fmaf(-2.0f, rcp_approx(E + 1.0f), 1.0f);
  float u2 = ex * ex;
  float poly = fmaf(u2, fmaf(u2, fmaf(u2, -17.0f / 315.0f, 2.0f / 15.0f), -1.0f / 3.0f), 1.0f);
  float tx = (ex < 0.125f) ? ex * poly : tx_big;
  float yy = x * tx;
  float dd = tx - x * ex * (tx * tx - 1.0f);
  const bool big = x > 20.0f;
  y = big ? x : yy;
  d1 = big ? 1.0f : dd;
}

template <int ACT>
__device__ __forceinline__ void tc_hidden_act(float x, float& y, float& d1) {
  if (ACT == NEDDF_ACT_TANHEXP) tanhexp_fast(x, y, d1);
  else hidden_act<ACT>(x, y, d1);
}

// x = hi + lo with hi, lo fp16 (round to nearest); returns packed pairs and flags fp16 overflow
// `amax` tracks max |value| (fp16 range check, one FMNMX per value; NaN shows up in the outputs)
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo, float& amax) {
  __half2 h = __floats2half2_rn(a, b);
  float2 hf = __half22float2(h);
  __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
  hi = *reinterpret_cast<uint32_t*>(&h);
  lo = *reinterpret_cast<uint32_t*>(&l);
  amax = fmaxf(amax, fmaxf(fabsf(a), fabsf(b)));
}

// same split; the range check on packed halves (one HMNMX2 per pair instead of two FMNMX + FABS: an operand
// beyond fp16 range rounds to inf and is caught at the end of the kernel)
__device__ __forceinline__ void split2h(float a, float b, uint32_t& hi, uint32_t& lo, __half2& amax) {
  __half2 h = __floats2half2_rn(a, b);
  float2 hf = __half22float2(h);
  __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
  hi = *reinterpret_cast<uint32_t*>(&h);
  lo = *reinterpret_cast<uint32_t*>(&l);
  amax = __hmax2(amax, __habs2(h));
}

// write the rows (value, Jx, Jy, Jz) of sample s at K index k into an operand buffer pair;
// type-major row order: row = 32*j + s.  rows = 4, or 1 when only the value row is consumed.
__device__ __forceinline__ void store_sample(unsigned char* hi_buf, unsigned char* lo_buf, int KC, int s, int k,
                                             float v0, float v1, float v2, float v3, float& bad, int rows = 4) {
  uint32_t h0, l0, h1, l1;
  split2(v0, v1, h0, l0, bad);
  split2(v2, v3, h1, l1, bad);
  const uint32_t off = act_off(s, k, KC);
  const uint32_t tstride = (uint32_t)(4 * KC * 16);  // 32 rows = 4 row groups
  *reinterpret_cast<uint16_t*>(hi_buf + off) = (uint16_t)(h0 & 0xffffu);
  *reinterpret_cast<uint16_t*>(lo_buf + off) = (uint16_t)(l0 & 0xffffu);
  if (rows > 1) {
    *reinterpret_cast<uint16_t*>(hi_buf + off + tstride) = (uint16_t)(h0 >> 16);
    *reinterpret_cast<uint16_t*>(lo_buf + off + tstride) = (uint16_t)(l0 >> 16);
    *reinterpret_cast<uint16_t*>(hi_buf + off + 2 * tstride) = (uint16_t)(h1 & 0xffffu);
    *reinterpret_cast<uint16_t*>(lo_buf + off + 2 * tstride) = (uint16_t)(l1 & 0xffffu);
    *reinterpret_cast<uint16_t*>(hi_buf + off + 3 * tstride) = (uint16_t)(h1 >> 16);
    *reinterpret_cast<uint16_t*>(lo_buf + off + 3 * tstride) = (uint16_t)(l1 >> 16);
  }
}

// Sum v[i] over the 32 lanes of the warp for N values per lane; afterwards lane L holds the totals of
// indices [L * N / 32, (L + 1) * N / 32) in v[0 .. N/32).  N / 32 * 31 shuffles instead of 5 N.
template <int N>
__device__ __forceinline__ void warp_transpose_reduce(float (&v)[N], int lane) {
#pragma unroll
  for (int d = 16, n = N / 2; d >= 1; d >>= 1, n >>= 1) {
    const bool upper = (lane & d) != 0;
#pragma unroll
    for (int t = 0; t < n; ++t) {
      const float send = upper ? v[t] : v[t + n];
      const float keep = upper ? v[t + n] : v[t];
      v[t] = keep + __shfl_xor_sync(0xffffffffu, send, d);
    }
  }
}

// N = 16: lane L ends with the total of index L >> 1 (both lanes of a pair hold it)
__device__ __forceinline__ void warp_transpose_reduce16(float (&v)[16], int lane) {
#pragma unroll
  for (int d = 16, n = 8; d >= 2; d >>= 1, n >>= 1) {
    const bool upper = (lane & d) != 0;
#pragma unroll
    for (int t = 0; t < n; ++t) {
      const float send = upper ? v[t] : v[t + n];
      const float keep = upper ? v[t + n] : v[t];
      v[t] = keep + __shfl_xor_sync(0xffffffffu, send, d);
    }
  }
  v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
}

// ---------------------------------------------------------------------------------------------
// CTA pairs: clusters of two CTAs on one TPC, tcgen05 cta_group::2
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_id_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `saddr` (a shared::cta address of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_v2(uint32_t caddr, uint2 v) {
  asm volatile("st.shared::cluster.v2.b32 [%0], {%1,%2};" ::"r"(caddr), "r"(v.x), "r"(v.y) : "memory");
}
__device__ __forceinline__ void st_cluster_v4(uint32_t caddr, uint4 v) {
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(caddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void st_cluster_f32(uint32_t caddr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(caddr), "f"(v) : "memory");
}
// arrive (release at cluster scope) on an mbarrier given by its shared::cluster address
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t caddr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(caddr) : "memory");
}
// arrive without any memory fence (the data handed over is tensor memory, ordered by tcgen05 fences, or
// was already published by an explicit fence)
__device__ __forceinline__ void mbar_arrive_cluster_relaxed(uint32_t caddr) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(caddr) : "memory");
}
// default semantics (release at CTA scope): what CUTLASS's ClusterBarrier::arrive(cta_id) uses for DSMEM hand-offs
__device__ __forceinline__ void mbar_arrive_cluster_default(uint32_t caddr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(caddr) : "memory");
}
// four 8-column loads in flight, one wait (lane = TMEM lane, 4 x 8 consecutive columns at taddr + stride * j)
__device__ __forceinline__ void tmem_ld8x4(uint32_t t0, uint32_t t1, uint32_t t2, uint32_t t3, float a[8], float b[8], float c[8], float d[8]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%32];\n"
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%8,%9,%10,%11,%12,%13,%14,%15}, [%33];\n"
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%16,%17,%18,%19,%20,%21,%22,%23}, [%34];\n"
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%24,%25,%26,%27,%28,%29,%30,%31}, [%35];\n"
      "tcgen05.wait::ld.sync.aligned;\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(t0), "r"(t1), "r"(t2), "r"(t3)
      : "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    a[i] = __uint_as_float(r[i]);
    b[i] = __uint_as_float(r[8 + i]);
    c[i] = __uint_as_float(r[16 + i]);
    d[i] = __uint_as_float(r[24 + i]);
  }
}
__device__ __forceinline__ void fence_acq_rel_cluster() { asm volatile("fence.acq_rel.cluster;" ::: "memory"); }
// wait with acquire at cluster scope (the barrier is local; the arrivals may come from the peer CTA)
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "TC_WAITC:\n"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n"
      "@p bra TC_DONEC;\n"
      "bra TC_WAITC;\n"
      "TC_DONEC:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// TMA 1-D bulk copy global -> shared memory, completion counted in bytes on an mbarrier
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void fence_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc2(uint32_t* dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
// M = 256 (128 rows of A / D per CTA), A from each CTA's tensor memory, B split between the CTAs
__device__ __forceinline__ void mma2_f16_ts_elect(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                                  uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p, q;\n"
      "elect.sync _|q, 0xffffffff;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "@q tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}
// commit of all prior cta_group::2 MMAs to the mbarrier at the same shared-memory offset in the CTAs of `mask`
__device__ __forceinline__ void mma2_commit_elect(uint32_t bar, uint32_t mask) {
  asm volatile(
      "{\n"
      ".reg .pred q;\n"
      ".reg .b16 lo, hi;\n"
      "mov.b32 {lo, hi}, %1;\n"
      "elect.sync _|q, 0xffffffff;\n"
      "@q tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], lo;\n"
      "}\n" ::"r"(bar),
      "r"(mask)
      : "memory");
}

}  // namespace tc
}  // namespace neddf
