// K2 (tensor-core engine): the NeDDF field network as one persistent tcgen05 megakernel.
//
// Reference: NeDDF.forward (neddf/network/neddf.py:162-309), sample geometry of
// neddf/ray/ray.py:88-194 fused into the prologue.
//
// Precision.  Parity with the fp32 reference (1e-4) rules out single-pass TF32/BF16/FP16 operands
// (measured 1e-3..1e-2).  Every GEMM operand is therefore split x = hi + lo into two
// fp16 values and three products are accumulated in fp32 in TMEM:
//     hi_w*hi_x + lo_w*hi_x + hi_w*lo_x        (kind::f16, fp32 accumulate)
// which is as accurate as 3xTF32 (emulated on the oracle: density 3.5e-6, colour 1.7e-6) at twice
// TF32's tensor rate.  fp16's range (65504) is checked in the epilogue; a tile that exceeds it sets
// status bit 2 and the host raises (the fp32 engine covers such networks).
//
// Orientation.  Every sample carries 4 rows (value + d/dx, d/dy, d/dz).  The MMAs are issued
// "swapped": A = weights (M = 128 output channels, K-major), B = activations (N = 128 rows =
// 32 samples x 4, MN-major), so the accumulator has one output channel per TMEM lane and the four
// rows of a sample in four adjacent columns.  The epilogue thread that owns a channel therefore
// holds x and its three Jacobian entries in registers: y = f(x), G = f'(x) J need no cross-thread
// traffic, and it writes 8 consecutive rows (16 bytes) of the next layer's B operand per store.
// The narrow heads (256->1,1,3) run in the standard orientation (A = activations, N = 16) so that
// their result has one sample row per lane.
//
// Per SM: one CTA of 10 warps, persistent over 32-sample tiles.
//   warps 0-7  epilogue: TMEM -> registers (tcgen05.ld), bias + activation + Jacobian, fp16
//              hi/lo split, st.shared into the next B operand; also prologue (geometry, PE)
//   warp  8    TMA producer: cp.async.bulk of 8 KB weight chunks through a 5-stage ring
//   warp  9    MMA issuer: one thread, tcgen05.mma + tcgen05.commit onto mbarriers
// Shared memory (B operands, canonical no-swizzle MN-major: [row/8][k][row%8] fp16):
//   H   hi/lo  128 rows x 256 k   2 x 64 KB   hidden activations, rewritten layer after layer
//   AUX hi/lo  128 rows x  96 k   2 x 24 KB   E_s (trunk input + skip) then [E0|D|n] (colour input)
//   W ring     5 x 8 KB                      weight chunks [hi 4 KB | lo 4 KB], K-major
// TMEM: cols [0,128) channels 0-127, [128,256) channels 128-255, [256,272) head results.
#include "field_math.cuh"

#include <cuda_fp16.h>

#include <algorithm>
#include <cstring>

namespace neddf {
namespace tc {

constexpr int kTileS = 32;              // samples per tile
constexpr int kRows = 4 * kTileS;       // 128 = MMA N (hidden) / M (heads)
constexpr int kHK = 256;                // K capacity of H
constexpr int kAuxK = 96;               // K capacity of AUX
constexpr int kStageBytes = 8192;
constexpr int kNumStages = 5;
constexpr int kEpiWarps = 8;
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kThreads = kEpiThreads + 64;
constexpr int kMaxSteps = kMaxHidden + 2;
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kHeadCol = 256;

constexpr uint32_t kHBytes = kRows * kHK * 2;      // 65536 per hi / lo
constexpr uint32_t kAuxBytes = kRows * kAuxK * 2;  // 24576 per hi / lo
constexpr uint32_t kOffHHi = 0;
constexpr uint32_t kOffHLo = kOffHHi + kHBytes;
constexpr uint32_t kOffAuxHi = kOffHLo + kHBytes;
constexpr uint32_t kOffAuxLo = kOffAuxHi + kAuxBytes;
constexpr uint32_t kOffStages = kOffAuxLo + kAuxBytes;
constexpr uint32_t kOffScratch = kOffStages + kNumStages * kStageBytes;

struct Scratch {
  float geo[kTileS][12];   // pos[3], dir[3], var[3], pad
  HeadOut head[kTileS];
  uint64_t full[kNumStages];
  uint64_t empty[kNumStages];
  uint64_t act_ready;
  uint64_t acc_ready;
  uint32_t tmem_base;
  uint32_t pad;
};
constexpr uint32_t kSmemBytes = kOffScratch + sizeof(Scratch);
static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");

enum StepKind { kStepHidden = 0, kStepHeadDA = 1, kStepHeadCol = 2 };

struct Step {
  int kind;
  int aux_ksteps;  // K-steps (16) taken from AUX first ...
  int h_ksteps;    // ... then from H
  int bias_off;    // hidden layers: offset into the plain-order bias array
};

struct TcParams {
  FieldParams f;
  int n_steps;
  int chunks_per_tile;
  Step step[kMaxSteps];
  const unsigned char* w_tc;  // packed chunks, kStageBytes each, in consumption order
  const float* bias;          // [n_hidden][256] plain channel order
  int* status;
};

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
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
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
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
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
constexpr uint32_t kIdescHidden = make_idesc(128, kRows, 0, 1);  // A weights K-major, B activations MN-major
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

// x = hi + lo with hi, lo fp16 (round to nearest); returns packed pairs and flags fp16 overflow
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo, uint32_t& bad) {
  __half2 h = __floats2half2_rn(a, b);
  float2 hf = __half22float2(h);
  __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
  hi = *reinterpret_cast<uint32_t*>(&h);
  lo = *reinterpret_cast<uint32_t*>(&l);
  bad |= ((hi & 0x7c00u) == 0x7c00u) | ((hi & 0x7c000000u) == 0x7c000000u);
}

// write the 4 rows (value, Jx, Jy, Jz) of sample s, K index k into an operand buffer pair
__device__ __forceinline__ void store_sample(unsigned char* hi_buf, unsigned char* lo_buf, int KC, int s, int k,
                                             float v0, float v1, float v2, float v3, uint32_t& bad) {
  uint32_t h0, l0, h1, l1;
  split2(v0, v1, h0, l0, bad);
  split2(v2, v3, h1, l1, bad);
  uint32_t off = act_off(4 * s, k, KC);
  *reinterpret_cast<uint2*>(hi_buf + off) = make_uint2(h0, h1);
  *reinterpret_cast<uint2*>(lo_buf + off) = make_uint2(l0, l1);
}

// ---------------------------------------------------------------------------------------------
// prologue pieces (epilogue warps)
// ---------------------------------------------------------------------------------------------
// geometry of the tile's samples -> scratch (one thread per sample)
__device__ __forceinline__ void tile_geometry(const FieldParams& p, Scratch* sc, int64_t n0, int s) {
  float pos[3] = {0.f, 0.f, 0.f}, dir[3] = {0.f, 0.f, 1.f}, var[3] = {0.f, 0.f, 0.f};
  const int64_t n = n0 + s;
  if (n < p.n) {
    if (p.dists) {
      int64_t b = n / p.n_edges;
      int j = (int)(n % p.n_edges);
      const float* row = p.dists + b * p.n_edges;
      float o[3];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        o[i] = p.ray_orig[3 * b + i];
        dir[i] = p.ray_dir[3 * b + i];
      }
      sample_geometry(p.sampling_type, p.ray_radius, o, dir, row[j], far_edge(row, j, p.n_edges), pos, var);
    } else {
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        pos[i] = p.pos[3 * n + i];
        dir[i] = p.dir[3 * n + i];
        var[i] = p.var[3 * n + i];
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    sc->geo[s][i] = pos[i];
    sc->geo[s][3 + i] = dir[i];
    sc->geo[s][6 + i] = var[i];
  }
}

// position embedding of sample s into AUX at K offset k0; scaled = distance-trunk scaling
// (neddf.py:200-204) else plain (neddf.py:205-209).  `sub`/`nsub` split the 3*E entries.
__device__ __forceinline__ void write_pos_embedding(const FieldParams& p, const Scratch* sc, unsigned char* aux_hi,
                                                    unsigned char* aux_lo, int s, int sub, int nsub, bool scaled,
                                                    uint32_t& bad) {
  const int half = 3 * p.embed_pos;
  for (int idx = sub; idx < half; idx += nsub) {
    int e = idx / 3, d = idx - 3 * e;
    PeEntry q = pe_entry(e, sc->geo[s][d], sc->geo[s][6 + d], p.lowpass[e]);
    float sc_ = scaled ? q.scale_s : q.scale_0;
    float g = q.freq * sc_;
    float js = g * q.c, jc = -g * q.s;
    float vs[4] = {sc_ * q.s, 0.f, 0.f, 0.f}, vc[4] = {sc_ * q.c, 0.f, 0.f, 0.f};
    vs[1 + d] = js;
    vc[1 + d] = jc;
    store_sample(aux_hi, aux_lo, kAuxK, s, idx, vs[0], vs[1], vs[2], vs[3], bad);
    store_sample(aux_hi, aux_lo, kAuxK, s, half + idx, vc[0], vc[1], vc[2], vc[3], bad);
  }
}

// ---------------------------------------------------------------------------------------------
// the megakernel
// ---------------------------------------------------------------------------------------------
template <int ACT>
__global__ void __launch_bounds__(kThreads, 1) field_tc_kernel(const __grid_constant__ TcParams P) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const FieldParams& p = P.f;
  unsigned char* h_hi = smem + kOffHHi;
  unsigned char* h_lo = smem + kOffHLo;
  unsigned char* aux_hi = smem + kOffAuxHi;
  unsigned char* aux_lo = smem + kOffAuxLo;
  unsigned char* stages = smem + kOffStages;
  Scratch* sc = reinterpret_cast<Scratch*>(smem + kOffScratch);

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;

  const int64_t n_tiles = (p.n + kTileS - 1) / kTileS;
  int64_t my_tiles = 0;
  if ((int64_t)blockIdx.x < n_tiles) my_tiles = (n_tiles - 1 - blockIdx.x) / gridDim.x + 1;
  const int64_t total_chunks = my_tiles * P.chunks_per_tile;

  if (tid == 0) {
    for (int i = 0; i < kNumStages; ++i) {
      mbar_init(&sc->full[i], 1);
      mbar_init(&sc->empty[i], 1);
    }
    mbar_init(&sc->act_ready, kEpiThreads);
    mbar_init(&sc->acc_ready, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 9) tmem_alloc(&sc->tmem_base, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sc->tmem_base;

  if (warp == 8) {
    // ===================== TMA producer =====================================================
    if (lane == 0) {
      for (int64_t g = 0; g < total_chunks; ++g) {
        const int stage = (int)(g % kNumStages);
        if (g >= kNumStages) mbar_wait(&sc->empty[stage], (uint32_t)(((g / kNumStages) - 1) & 1));
        mbar_expect_tx(&sc->full[stage], kStageBytes);
        bulk_g2s(stages + stage * kStageBytes, P.w_tc + (size_t)(g % P.chunks_per_tile) * kStageBytes, kStageBytes,
                 &sc->full[stage]);
      }
    }
  } else if (warp == 9) {
    // ===================== MMA issuer ========================================================
    if (lane == 0) {
      const uint32_t s_hhi = smem_u32(h_hi), s_hlo = smem_u32(h_lo);
      const uint32_t s_ahi = smem_u32(aux_hi), s_alo = smem_u32(aux_lo);
      const uint32_t s_stage = smem_u32(stages);
      int64_t g = 0;
      uint32_t act_phase = 0;
      for (int64_t t = 0; t < my_tiles; ++t) {
        for (int si = 0; si < P.n_steps; ++si) {
          const Step& st = P.step[si];
          mbar_wait(&sc->act_ready, act_phase);
          act_phase ^= 1;
          tc_fence_after();
          if (st.kind == kStepHidden) {
            const int ksteps = st.aux_ksteps + st.h_ksteps;
            for (int ks = 0; ks < ksteps; ++ks) {
              uint32_t b_hi, b_lo, sbo;
              if (ks < st.aux_ksteps) {
                b_hi = s_ahi + ks * 256;
                b_lo = s_alo + ks * 256;
                sbo = kAuxK * 16;
              } else {
                b_hi = s_hhi + (ks - st.aux_ksteps) * 256;
                b_lo = s_hlo + (ks - st.aux_ksteps) * 256;
                sbo = kHK * 16;
              }
              const uint64_t db_hi = make_desc(b_hi, 128, sbo), db_lo = make_desc(b_lo, 128, sbo);
#pragma unroll
              for (int half = 0; half < 2; ++half, ++g) {
                const int stage = (int)(g % kNumStages);
                mbar_wait(&sc->full[stage], (uint32_t)((g / kNumStages) & 1));
                tc_fence_after();
                const uint32_t a = s_stage + stage * kStageBytes;
                const uint64_t da_hi = make_desc(a, 128, 256), da_lo = make_desc(a + 4096, 128, 256);
                const uint32_t d = tmem + half * kRows;
                mma_f16(d, da_hi, db_hi, kIdescHidden, ks > 0);
                mma_f16(d, da_lo, db_hi, kIdescHidden, 1);
                mma_f16(d, da_hi, db_lo, kIdescHidden, 1);
                mma_commit(&sc->empty[stage]);
              }
            }
          } else {
            // heads: D[row, n] = sum_k H[row,k] * Wh[n,k]; chunk g = hi, g+1 = lo
            const int st0 = (int)(g % kNumStages), st1 = (int)((g + 1) % kNumStages);
            mbar_wait(&sc->full[st0], (uint32_t)((g / kNumStages) & 1));
            mbar_wait(&sc->full[st1], (uint32_t)(((g + 1) / kNumStages) & 1));
            tc_fence_after();
            const uint32_t w_hi = s_stage + st0 * kStageBytes, w_lo = s_stage + st1 * kStageBytes;
            const uint32_t d = tmem + kHeadCol;
            for (int ks = 0; ks < kHK / 16; ++ks) {
              const uint64_t da_hi = make_desc(s_hhi + ks * 256, 128, kHK * 16);
              const uint64_t da_lo = make_desc(s_hlo + ks * 256, 128, kHK * 16);
              const uint64_t db_hi = make_desc(w_hi + ks * 256, 128, 4096);
              const uint64_t db_lo = make_desc(w_lo + ks * 256, 128, 4096);
              mma_f16(d, da_hi, db_hi, kIdescHead, ks > 0);
              mma_f16(d, da_lo, db_hi, kIdescHead, 1);
              mma_f16(d, da_hi, db_lo, kIdescHead, 1);
            }
            mma_commit(&sc->empty[st0]);
            mma_commit(&sc->empty[st1]);
            g += 2;
          }
          mma_commit(&sc->acc_ready);
        }
      }
    }
  } else {
    // ===================== epilogue warps =====================================================
    const int quarter = warp & 3;   // TMEM lane quarter this warp may access
    const int half = warp >> 2;     // accumulator half (channels 128*half ..)
    const int ch = 128 * half + 32 * quarter + lane;
    const uint32_t lane_addr = (uint32_t)(32 * quarter) << 16;
    uint32_t bad = 0;
    uint32_t acc_phase = 0;

    // prologue of the first tile
    auto prologue = [&](int64_t tile) {
      const int64_t n0 = tile * kTileS;
      if (tid < kTileS) tile_geometry(p, sc, n0, tid);
      asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
      const int s = tid >> 3, sub = tid & 7;
      write_pos_embedding(p, sc, aux_hi, aux_lo, s, sub, 8, true, bad);
      // zero the K padding of E_s (its weights are zero, the operand must still be finite)
      for (int k = p.n_e0 + sub; k < 64; k += 8) store_sample(aux_hi, aux_lo, kAuxK, s, k, 0.f, 0.f, 0.f, 0.f, bad);
    };
    if (my_tiles > 0) {
      prologue(blockIdx.x);
      fence_async_smem();
      mbar_arrive(&sc->act_ready);
    }

    for (int64_t t = 0; t < my_tiles; ++t) {
      const int64_t tile = blockIdx.x + t * gridDim.x;
      const int64_t n0 = tile * kTileS;
      for (int si = 0; si < P.n_steps; ++si) {
        const Step& st = P.step[si];
        mbar_wait(&sc->acc_ready, acc_phase);
        acc_phase ^= 1;
        tc_fence_after();
        if (st.kind == kStepHidden) {
          const float bias = __ldg(P.bias + st.bias_off + ch);
          const uint32_t tbase = tmem + lane_addr + half * kRows;
#pragma unroll 1
          for (int cb = 0; cb < kRows / 16; ++cb) {  // 16 columns = 4 samples per load
            float v[16];
            tmem_ld16(tbase + cb * 16, v);
#pragma unroll
            for (int q = 0; q < 2; ++q) {  // 2 samples = one 16-byte row group
              float y0, d0, y1, d1;
              hidden_act<ACT>(v[8 * q + 0] + bias, y0, d0);
              hidden_act<ACT>(v[8 * q + 4] + bias, y1, d1);
              uint32_t h[4], l[4];
              split2(y0, d0 * v[8 * q + 1], h[0], l[0], bad);
              split2(d0 * v[8 * q + 2], d0 * v[8 * q + 3], h[1], l[1], bad);
              split2(y1, d1 * v[8 * q + 5], h[2], l[2], bad);
              split2(d1 * v[8 * q + 6], d1 * v[8 * q + 7], h[3], l[3], bad);
              const uint32_t off = (uint32_t)((2 * cb + q) * (kHK * 16) + ch * 16);
              *reinterpret_cast<uint4*>(h_hi + off) = make_uint4(h[0], h[1], h[2], h[3]);
              *reinterpret_cast<uint4*>(h_lo + off) = make_uint4(l[0], l[1], l[2], l[3]);
            }
          }
        } else if (st.kind == kStepHeadDA) {
          if (warp < 4) {
            // one row per lane: columns 0,1 = ddf_out, aux_out of that row (neddf.py:220-230)
            float v[4];
            tmem_ld4(tmem + lane_addr + kHeadCol, v);
            float ddf[4], aux[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              ddf[j] = __shfl_down_sync(0xffffffffu, v[0], j);
              aux[j] = __shfl_down_sync(0xffffffffu, v[1], j);
            }
            if ((lane & 3) == 0) {
              const int s = 8 * quarter + (lane >> 2);
              ddf[0] += __ldg(p.b_head + 0);
              aux[0] += __ldg(p.b_head + 1);
              HeadOut h;
              head_density(ddf, aux, p.d_near, p.aux_grad_scale, p.density_act, h);
              sc->head[s] = h;
              const int kn = p.n_e0 + p.n_d;  // normal: detached, zero Jacobian (neddf.py:243-253)
#pragma unroll
              for (int i = 0; i < 3; ++i) store_sample(aux_hi, aux_lo, kAuxK, s, kn + i, h.normal[i], 0.f, 0.f, 0.f, bad);
            }
          } else {
            // colour-trunk inputs E0 | D (| pad) into AUX while warps 0-3 finish the heads
            const int t2 = tid - 128;
            const int s = t2 >> 2, sub = t2 & 3;
            write_pos_embedding(p, sc, aux_hi, aux_lo, s, sub, 4, false, bad);
            const int dhalf = 3 * p.embed_dir;
            for (int idx = sub; idx < dhalf; idx += 4) {
              int e = idx / 3, d = idx - 3 * e;
              float sn, cs;
              sincosf((float)(1u << e) * sc->geo[s][3 + d], &sn, &cs);
              store_sample(aux_hi, aux_lo, kAuxK, s, p.n_e0 + idx, sn, 0.f, 0.f, 0.f, bad);
              store_sample(aux_hi, aux_lo, kAuxK, s, p.n_e0 + dhalf + idx, cs, 0.f, 0.f, 0.f, bad);
            }
            for (int k = p.n_e0 + p.n_d + 3 + sub; k < kAuxK; k += 4)
              store_sample(aux_hi, aux_lo, kAuxK, s, k, 0.f, 0.f, 0.f, 0.f, bad);
          }
        } else {
          // colour head (neddf.py:257) + penalties (:259-300) + outputs, then next tile's prologue
          if (warp < 4) {
            float v[4];
            tmem_ld4(tmem + lane_addr + kHeadCol, v);
            float cv[4][3];
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
              for (int c = 0; c < 3; ++c) cv[j][c] = __shfl_down_sync(0xffffffffu, v[c], j);
            const int s = 8 * quarter + (lane >> 2);
            const int64_t n = n0 + s;
            if ((lane & 3) == 0 && n < p.n) {
              const HeadOut& h = sc->head[s];
              float col[3], colJ[3][3];
#pragma unroll
              for (int c = 0; c < 3; ++c) {
                col[c] = cv[0][c] + __ldg(p.b_head + 2 + c);
#pragma unroll
                for (int i = 0; i < 3; ++i) colJ[i][c] = cv[1 + i][c];
              }
              if (p.distance) p.distance[n] = h.distance;
              if (p.density) p.density[n] = h.density;
              if (p.aux_grad) p.aux_grad[n] = h.aux;
              if (p.color) {
                p.color[3 * n + 0] = col[0];
                p.color[3 * n + 1] = col[1];
                p.color[3 * n + 2] = col[2];
              }
              if (p.penalty) p.penalty[n] = field_penalty(h, col, colJ, p.distance_range_max, p.penalty_weight);
            }
          }
          asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");  // scratch reads done
          if (t + 1 < my_tiles) prologue(tile + gridDim.x);
        }
        tc_fence_before();
        fence_async_smem();
        if (t + 1 < my_tiles || si + 1 < P.n_steps) mbar_arrive(&sc->act_ready);
      }
    }
    if (bad && P.status) atomicOr(P.status, 4);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc(tmem, kTmemCols);
}

// ---------------------------------------------------------------------------------------------
// weight packing: fp32 [in,out] -> fp16 hi/lo chunks in consumption order
// ---------------------------------------------------------------------------------------------
struct TcPackArgs {
  const float* w[kMaxHidden + 3];
  const float* b[kMaxHidden + 3];
  int k_in[kMaxHidden];
  int aux_real[kMaxHidden];  // leading input channels that live in AUX (0 if none)
  int aux_pad[kMaxHidden];   // their padded K extent in AUX
  int ksteps[kMaxHidden];
  int chunk0[kMaxHidden + 3];  // first chunk index of the layer / head
  int n_hidden;
};

__global__ void tc_pack_hidden_kernel(TcPackArgs a, unsigned char* __restrict__ dst, float* __restrict__ bias) {
  const int l = blockIdx.y;
  const int total = a.ksteps[l] * 2 * 128 * 16;  // (kstep, half, m, k)
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    int k = idx & 15, m = (idx >> 4) & 127, half = (idx >> 11) & 1, ks = idx >> 12;
    int kk = ks * 16 + k;  // K index in operand space (AUX part first, then H)
    int row;               // reference weight row, -1 = zero padding
    if (kk < a.aux_pad[l]) row = (kk < a.aux_real[l]) ? kk : -1;
    else row = a.aux_real[l] + (kk - a.aux_pad[l]);
    if (row >= a.k_in[l]) row = -1;
    float w = (row >= 0) ? a.w[l][(size_t)row * kWidth + 128 * half + m] : 0.f;
    __half hi = __float2half_rn(w);
    __half lo = __float2half_rn(w - __half2float(hi));
    unsigned char* chunk = dst + (size_t)(a.chunk0[l] + ks * 2 + half) * kStageBytes;
    *reinterpret_cast<__half*>(chunk + wchunk_off(m, k)) = hi;
    *reinterpret_cast<__half*>(chunk + 4096 + wchunk_off(m, k)) = lo;
  }
  if (blockIdx.x == 0)
    for (int c = threadIdx.x; c < kWidth; c += blockDim.x) bias[l * kWidth + c] = a.b[l][c];
}

// heads: chunk0[n_hidden] = (ddf,aux) hi chunk, +1 lo; chunk0[n_hidden+1] = colour hi, +1 lo
__global__ void tc_pack_heads_kernel(TcPackArgs a, unsigned char* __restrict__ dst) {
  const int nh = a.n_hidden;
  for (int idx = threadIdx.x; idx < 16 * kWidth; idx += blockDim.x) {
    int k = idx % kWidth, n = idx / kWidth;
    float wda = (n == 0) ? a.w[nh + 0][k] : (n == 1) ? a.w[nh + 1][k] : 0.f;
    float wc = (n < 3) ? a.w[nh + 2][3 * k + n] : 0.f;
    __half h1 = __float2half_rn(wda), h2 = __float2half_rn(wc);
    __half l1 = __float2half_rn(wda - __half2float(h1)), l2 = __float2half_rn(wc - __half2float(h2));
    *reinterpret_cast<__half*>(dst + (size_t)a.chunk0[nh] * kStageBytes + hchunk_off(n, k)) = h1;
    *reinterpret_cast<__half*>(dst + (size_t)(a.chunk0[nh] + 1) * kStageBytes + hchunk_off(n, k)) = l1;
    *reinterpret_cast<__half*>(dst + (size_t)a.chunk0[nh + 1] * kStageBytes + hchunk_off(n, k)) = h2;
    *reinterpret_cast<__half*>(dst + (size_t)(a.chunk0[nh + 1] + 1) * kStageBytes + hchunk_off(n, k)) = l2;
  }
}

struct TcStorage {
  unsigned char* d_w = nullptr;
  float* d_bias = nullptr;
  int* d_status = nullptr;
  int n_steps = 0;
  int chunks_per_tile = 0;
  Step step[kMaxSteps];
  TcPackArgs pack;
};

// ---------------------------------------------------------------------------------------------
// self-test of the MMA building block (pins descriptor / layout conventions on hardware)
// ---------------------------------------------------------------------------------------------
// n == 128: hidden configuration.  C[m, n] = sum_k A[m,k] B[n,k], A = "weights" (K-major chunks),
//           B = "activations" (MN-major in H).
// n == 16 : head configuration.    C[m, n] = sum_k A[m,k] B[n,k], A = "activations" (MN-major in H),
//           B = "head weights" (K-major chunk).
__global__ void __launch_bounds__(128, 1) tc_selftest_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                             int n, int k, float* __restrict__ C) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* h_hi = smem;                 // 64 KB
  unsigned char* h_lo = smem + kHBytes;       // 64 KB
  unsigned char* w_hi = smem + 2 * kHBytes;   // up to 8 chunks of 4 KB (hidden, k <= 128) or 8 KB (head)
  unsigned char* w_lo = w_hi + 32768;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool hidden = (n == kRows);
  uint32_t bad = 0;
  // operands -> fp16 hi/lo in the kernel layouts
  for (int idx = tid; idx < 128 * k; idx += 128) {
    int kk = idx % k, r = idx / k;
    float a = A[(size_t)r * k + kk];
    __half hi = __float2half_rn(a), lo = __float2half_rn(a - __half2float(hi));
    if (hidden) {  // A = weights: chunk per k-step of 16, 4 KB each
      uint32_t off = (uint32_t)(kk >> 4) * 4096 + wchunk_off(r, kk & 15);
      *reinterpret_cast<__half*>(w_hi + off) = hi;
      *reinterpret_cast<__half*>(w_lo + off) = lo;
    } else {       // A = activations in H
      *reinterpret_cast<__half*>(h_hi + act_off(r, kk, kHK)) = hi;
      *reinterpret_cast<__half*>(h_lo + act_off(r, kk, kHK)) = lo;
    }
  }
  for (int idx = tid; idx < n * k; idx += 128) {
    int kk = idx % k, r = idx / k;
    float b = B[(size_t)r * k + kk];
    __half hi = __float2half_rn(b), lo = __float2half_rn(b - __half2float(hi));
    if (hidden) {
      *reinterpret_cast<__half*>(h_hi + act_off(r, kk, kHK)) = hi;
      *reinterpret_cast<__half*>(h_lo + act_off(r, kk, kHK)) = lo;
    } else {
      *reinterpret_cast<__half*>(w_hi + hchunk_off(r, kk)) = hi;
      *reinterpret_cast<__half*>(w_lo + hchunk_off(r, kk)) = lo;
    }
  }
  (void)bad;
  if (tid == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(&tmem_base, 128);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base;
  if (tid == 0) {
    const uint32_t s_hhi = smem_u32(h_hi), s_hlo = smem_u32(h_lo), s_whi = smem_u32(w_hi), s_wlo = smem_u32(w_lo);
    for (int ks = 0; ks < k / 16; ++ks) {
      if (hidden) {
        uint64_t da_hi = make_desc(s_whi + ks * 4096, 128, 256), da_lo = make_desc(s_wlo + ks * 4096, 128, 256);
        uint64_t db_hi = make_desc(s_hhi + ks * 256, 128, kHK * 16), db_lo = make_desc(s_hlo + ks * 256, 128, kHK * 16);
        mma_f16(tmem, da_hi, db_hi, kIdescHidden, ks > 0);
        mma_f16(tmem, da_lo, db_hi, kIdescHidden, 1);
        mma_f16(tmem, da_hi, db_lo, kIdescHidden, 1);
      } else {
        uint64_t da_hi = make_desc(s_hhi + ks * 256, 128, kHK * 16), da_lo = make_desc(s_hlo + ks * 256, 128, kHK * 16);
        uint64_t db_hi = make_desc(s_whi + ks * 256, 128, 4096), db_lo = make_desc(s_wlo + ks * 256, 128, 4096);
        mma_f16(tmem, da_hi, db_hi, kIdescHead, ks > 0);
        mma_f16(tmem, da_lo, db_hi, kIdescHead, 1);
        mma_f16(tmem, da_hi, db_lo, kIdescHead, 1);
      }
    }
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  // lane = row m of D; columns = n
  const int m = 32 * warp + lane;
  for (int cb = 0; cb < n / 16; ++cb) {
    float v[16];
    tmem_ld16(tmem + ((uint32_t)(32 * warp) << 16) + cb * 16, v);
#pragma unroll
    for (int i = 0; i < 16; ++i) C[(size_t)m * n + cb * 16 + i] = v[i];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 128);
}

}  // namespace tc

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static bool is_skip_layer(const neddf_field* f, int l) {
  // hidden layer l (>0) of the distance trunk takes [E_s | h] when layer l-1 is a skip layer
  for (int i = 0; i < f->cfg.n_skips; ++i)
    if (f->cfg.skips[i] == l - 1) return true;
  return false;
}

bool tc_supported(const neddf_field* f) {
  // built for the reference network shape: 60-channel position embedding, 24-channel direction
  // embedding (AUX holds 64 / 96 K), 256-wide layers
  return f->cfg.embed_pos_rank == 10 && f->cfg.embed_dir_rank == 4 && f->cfg.ddf_layer_width == kWidth &&
         f->cfg.col_layer_width == kWidth;
}

static int32_t tc_ensure(neddf_field* f) {
  if (f->tc) return NEDDF_OK;
  tc::TcStorage* S = new tc::TcStorage();
  const int n_hidden = f->n_ddf + f->n_col;
  int chunk = 0, si = 0;
  std::memset(&S->pack, 0, sizeof(S->pack));
  S->pack.n_hidden = n_hidden;
  for (int l = 0; l < n_hidden; ++l) {
    int aux_real = 0, aux_pad = 0, h_k = 0;
    if (l == 0) { aux_real = f->proto.n_e0; aux_pad = 64; }
    else if (l < f->n_ddf) { if (is_skip_layer(f, l)) { aux_real = f->proto.n_e0; aux_pad = 64; } h_k = kWidth; }
    else if (l == f->n_ddf) { aux_real = f->proto.off_h; aux_pad = tc::kAuxK; h_k = kWidth; }
    else h_k = kWidth;
    tc::Step& st = S->step[si++];
    st.kind = tc::kStepHidden;
    st.aux_ksteps = aux_pad / 16;
    st.h_ksteps = h_k / 16;
    st.bias_off = l * kWidth;
    S->pack.k_in[l] = f->shape_in[l];
    S->pack.aux_real[l] = aux_real;
    S->pack.aux_pad[l] = aux_pad;
    S->pack.ksteps[l] = st.aux_ksteps + st.h_ksteps;
    S->pack.chunk0[l] = chunk;
    chunk += 2 * S->pack.ksteps[l];
    if (l == f->n_ddf - 1) {  // distance / aux heads after the trunk
      tc::Step& hs = S->step[si++];
      hs.kind = tc::kStepHeadDA; hs.aux_ksteps = 0; hs.h_ksteps = kWidth / 16; hs.bias_off = 0;
      S->pack.chunk0[n_hidden] = chunk;
      chunk += 2;
    }
  }
  tc::Step& cs = S->step[si++];
  cs.kind = tc::kStepHeadCol; cs.aux_ksteps = 0; cs.h_ksteps = kWidth / 16; cs.bias_off = 0;
  S->pack.chunk0[n_hidden + 1] = chunk;
  chunk += 2;
  S->n_steps = si;
  S->chunks_per_tile = chunk;
  if (cudaMalloc(&S->d_w, (size_t)chunk * tc::kStageBytes) != cudaSuccess ||
      cudaMalloc(&S->d_bias, (size_t)n_hidden * kWidth * sizeof(float)) != cudaSuccess ||
      cudaMalloc(&S->d_status, sizeof(int)) != cudaSuccess) {
    delete S;
    return fail(NEDDF_E_CUDA, "tensor-core engine: cudaMalloc failed");
  }
  cudaMemset(S->d_status, 0, sizeof(int));
  f->tc = S;
  return NEDDF_OK;
}

void tc_destroy(neddf_field* f) {
  if (!f->tc) return;
  tc::TcStorage* S = static_cast<tc::TcStorage*>(f->tc);
  cudaFree(S->d_w);
  cudaFree(S->d_bias);
  cudaFree(S->d_status);
  delete S;
  f->tc = nullptr;
}

int32_t tc_pack_weights(neddf_field* f, const float* const* d_w, const float* const* d_b, cudaStream_t s) {
  int32_t rc = tc_ensure(f);
  if (rc != NEDDF_OK) return rc;
  tc::TcStorage* S = static_cast<tc::TcStorage*>(f->tc);
  tc::TcPackArgs a = S->pack;
  const int n_hidden = f->n_ddf + f->n_col;
  for (int l = 0; l < n_hidden + 3; ++l) {
    a.w[l] = d_w[l];
    a.b[l] = d_b[l];
  }
  tc::tc_pack_hidden_kernel<<<dim3(64, n_hidden), 256, 0, s>>>(a, S->d_w, S->d_bias);
  NEDDF_LAUNCH_CHECK();
  tc::tc_pack_heads_kernel<<<1, 256, 0, s>>>(a, S->d_w);
  NEDDF_LAUNCH_CHECK();
  return NEDDF_OK;
}

int32_t launch_field_tc(const neddf_field* f, FieldParams& p, int flags, cudaStream_t s) {
  (void)flags;
  tc::TcStorage* S = static_cast<tc::TcStorage*>(f->tc);
  if (!S) return fail(NEDDF_E_INVALID, "tensor-core engine: weights were never packed");
  tc::TcParams P;
  P.f = p;
  P.n_steps = S->n_steps;
  P.chunks_per_tile = S->chunks_per_tile;
  for (int i = 0; i < S->n_steps; ++i) P.step[i] = S->step[i];
  P.w_tc = S->d_w;
  P.bias = S->d_bias;
  P.status = S->d_status;
  int64_t n_tiles = (p.n + tc::kTileS - 1) / tc::kTileS;
  int grid = (int)std::min<int64_t>(n_tiles, sm_count());
  auto launch = [&](auto kern) -> int32_t {
    NEDDF_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::kSmemBytes));
    kern<<<grid, tc::kThreads, tc::kSmemBytes, s>>>(P);
    NEDDF_LAUNCH_CHECK();
    return NEDDF_OK;
  };
  switch (p.hidden_act) {
    case NEDDF_ACT_TANHEXP: return launch(tc::field_tc_kernel<NEDDF_ACT_TANHEXP>);
    case NEDDF_ACT_RELU: return launch(tc::field_tc_kernel<NEDDF_ACT_RELU>);
    case NEDDF_ACT_LEAKYRELU: return launch(tc::field_tc_kernel<NEDDF_ACT_LEAKYRELU>);
  }
  return fail(NEDDF_E_INVALID, "tensor-core engine: unknown activation");
}

int32_t tc_read_status(const neddf_field* f, int* out, cudaStream_t s) {
  tc::TcStorage* S = static_cast<tc::TcStorage*>(f->tc);
  *out = 0;
  if (!S) return NEDDF_OK;
  NEDDF_CUDA_CHECK(cudaMemcpyAsync(out, S->d_status, sizeof(int), cudaMemcpyDeviceToHost, s));
  NEDDF_CUDA_CHECK(cudaStreamSynchronize(s));
  NEDDF_CUDA_CHECK(cudaMemsetAsync(S->d_status, 0, sizeof(int), s));
  return NEDDF_OK;
}

}  // namespace neddf

extern "C" int32_t neddf_tc_selftest(const float* d_a, const float* d_b, int32_t m, int32_t n, int32_t k, float* d_c,
                                     void* stream) {
  using namespace neddf;
  if (m != 128 || (n != 128 && n != 16) || k < 16 || k > (n == 128 ? 128 : 256) || (k % 16) != 0)
    return fail(NEDDF_E_INVALID, "neddf_tc_selftest: need m=128, n in {128,16}, k multiple of 16 (<=128 for n=128, <=256 for n=16)");
  if (!d_a || !d_b || !d_c) return fail(NEDDF_E_INVALID, "neddf_tc_selftest: NULL pointer");
  size_t smem = 2 * tc::kHBytes + 65536;
  NEDDF_CUDA_CHECK(cudaFuncSetAttribute(tc::tc_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  tc::tc_selftest_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(d_a, d_b, n, k, d_c);
  NEDDF_LAUNCH_CHECK();
  return NEDDF_OK;
}
