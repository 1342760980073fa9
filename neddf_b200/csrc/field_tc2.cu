// K2 (tensor-core engine, CTA pairs): the NeDDF field network as one persistent tcgen05 megakernel
// in which the two SMs of a TPC share every weight chunk (tcgen05.mma.cta_group::2).
//
// Reference: NeDDF.forward (neddf/network/neddf.py:162-309), sample geometry of
// neddf/ray/ray.py:88-194 fused into the prologue.  Arithmetic, precision scheme (fp16 hi/lo operand
// split, three products, fp32 accumulation) and the orientation of the MMAs are those of field_tc.cu;
// what changes is who shares what.
//
// Why pairs.  field_tc.cu streams the whole packed weight set (2.56 MB) from L2 for every 32-sample
// tile; at 6.8e7 evaluations/s that is 5.9 TB/s of L2->SM traffic against a measured chip-wide ceiling
// of ~7.2 TB/s: the single-CTA kernel cannot pass ~9e7 whatever its instruction stream looks like
// (profiles/r02_summary.md).  Here a cluster of two CTAs processes a 64-sample tile with M = 256
// MMAs: CTA r keeps output channels 128r..128r+127 (its half of every weight chunk, in ITS tensor
// memory) and the B-operand rows of ITS 32 samples; the tensor cores of the pair read both halves of
// B.  Per evaluation the weight stream is halved.
//
// The price is an exchange: CTA r's epilogue produces channels 128r.. for all 64 samples, and the
// rows of the peer's 32 samples must land in the peer's shared memory (16-byte st.shared::cluster,
// measured 19-21 B/clk per CTA in both directions at once, tools/pair_probe.py): 64 KB per layer and
// CTA, hidden under the 6.1k-cycle MMA phase of a layer.
//
// Rows of a CTA's B operands (128 = 32 samples x {value, d/dx, d/dy, d/dz}):
//     row = 64 * hs + 16 * j + s'        s = 16 hs + s' local sample, j row type
// so that the "sample half" hs is a contiguous block of 8 row groups: one MMA covers N = 128 rows
// = half hs of BOTH CTAs (64 rows each), accumulating into TMEM columns [128 hs, 128 hs + 128) where
// column 64 c + 16 j + s' belongs to sample (CTA c, 16 hs + s').  The two halves ping-pong: while the
// epilogue warps drain half 0 the tensor cores run half 1 of the same layer on the SAME weight chunks,
// which therefore stay in the 16-stage tensor-memory ring for two passes (a chunk is freed by the
// commit of its second pass).  When only images are wanted (no fields_penalty) the colour trunk runs
// on the 16 value rows of each half (N = 32).
//
// The three narrow heads (256 -> 1, 1, 3) do not go to the tensor cores at all: the epilogue thread
// that owns a channel multiplies its 32 fresh activations by the head weights of that channel and
// the warp reduces over its 32 channels with a transposing butterfly (62 shuffles for 64 sums); the
// eight partial sums per sample (4 lane quarters x 2 CTAs) meet in the shared memory of the CTA that
// owns the sample, where one warp applies softplus / sigmoid / density / penalties.
//
// Weights.  A chunk = two K-steps of this CTA's 128 channels (16 KB: per K-step 4 KB hi | 4 KB lo in
// the K-major core-matrix layout).  One thread per CTA streams the chunks from L2 into a two-stage
// shared-memory ring with 1-D TMA bulk copies (cp.async.bulk + mbarrier complete_tx); one thread of the
// leader moves each staged chunk of BOTH CTAs into the tensor-memory ring (8 stages x 32 columns) with
// four tcgen05.cp.cta_group::2.128x256b and commits; the MMAs take A from tensor memory at the 64-cycle
// rate.  (Round-2 first version: eight loader warps per CTA went L2 -> registers -> tcgen05.st as in
// field_tc.cu; their ~1400-cycle period per chunk and warp was the bottleneck - tools/loader_timeline.py.)
//
// Per CTA: 19 warps.  warps 0-15 epilogue + prologue (lane quarter w % 4; sub-block w / 4 = (target
// CTA, 8 samples)); warp 16 MMA issuer (leader CTA only); warp 17 TMA producer; warp 18 tcgen05.cp
// issuer (leader) / forwarder of "my chunk landed" to the leader (peer).
#include "tc_ptx.cuh"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace neddf {
namespace tc2 {

using namespace tc;

constexpr int kPairS = 2 * kTileS;  // samples per pair tile
constexpr int kChunkK = 2;                            // K-steps per weight chunk
constexpr int kWChunkBytes = kChunkK * kChunkBytes;   // 16 KB per CTA and chunk
constexpr int kARing = 8;                             // chunks resident in tensor memory (32 columns each)
#ifndef NEDDF_TC2_TMA_SPLIT
#define NEDDF_TC2_TMA_SPLIT 1
#endif
constexpr int kTmaSplit = NEDDF_TC2_TMA_SPLIT;          // bulk copies per chunk
#ifndef NEDDF_TC2_SSTAGES
#define NEDDF_TC2_SSTAGES 3
#endif
constexpr int kSStages = NEDDF_TC2_SSTAGES;                           // shared-memory staging ring of the TMA copies
constexpr int kEpiWarps = 16;
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kMmaWarp = kEpiWarps;
constexpr int kTmaWarp = kEpiWarps + 1;
constexpr int kCpWarp = kEpiWarps + 2;
constexpr int kFinWarp = kEpiWarps + 3;  // finishes the heads (density, normals, colour, penalties, outputs)
#ifndef NEDDF_TC2_CP_ISSUERS
#define NEDDF_TC2_CP_ISSUERS 1
#endif
constexpr int kCpIssuers = NEDDF_TC2_CP_ISSUERS;  // threads (one per warp) that move staged chunks to tensor memory, chunk c by issuer c % kCpIssuers
constexpr int kCpWarp2 = kEpiWarps + 4;           // the second one
constexpr int kThreads = kEpiThreads + (3 + kCpIssuers) * 32;
constexpr int kWsFloats = 8 * kTileS * 12;  // per-CTA global workspace: head partial sums
constexpr uint32_t kACol = 256;  // first TMEM column of the weight ring
constexpr uint32_t kTmemCols = 512;
constexpr int kMaxSteps = kMaxHidden;
#ifndef NEDDF_TC2_GROUP
#define NEDDF_TC2_GROUP 8  // chunks per two-pass group (<= kARing)
#endif

constexpr uint32_t kOffHHi = 0;
constexpr uint32_t kOffHLo = kOffHHi + kHBytes;
constexpr uint32_t kOffAuxHi = kOffHLo + kHBytes;
constexpr uint32_t kOffAuxLo = kOffAuxHi + kAuxBytes;
constexpr uint32_t kOffStage = kOffAuxLo + kAuxBytes;
constexpr uint32_t kOffScratch = kOffStage + kSStages * kWChunkBytes;

struct Scratch {
  float geo[kTileS][12];  // pos[3], dir[3], var[3], pad
  uint64_t s_full[kSStages];     // TMA -> (this CTA's) cp issuer / forwarder: chunk landed in shared memory
  uint64_t s_empty[kSStages];    // tcgen05.commit (multicast) -> TMA producers: staged chunk copied to tensor memory
  uint64_t peer_full[kSStages];  // (leader) the peer's chunk landed in the peer's shared memory
  uint64_t a_full[kARing];       // (leader) tcgen05.commit -> MMA: chunk of both CTAs is in tensor memory
  uint64_t a_empty[kARing];      // (leader) tcgen05.commit -> cp issuer: chunk consumed by both passes
  uint64_t act_ready[2];     // (leader) epilogue warps of both CTAs -> MMA: accumulator hs drained, B rows of half hs rewritten
  uint64_t acc_ready[2];     // MMA -> epilogue warps (multicast commit): accumulator hs complete
  uint64_t head_ready[2];    // partial head sums of this CTA's samples of half hs are in the workspace (one arrival per CTA)
  uint64_t norm_ready[2];    // (leader) both CTAs wrote the surface normals of half hs into AUX
  uint64_t fin_done[2];      // the finishing warps of both CTAs consumed the sums of head type k (0 = distance / aux, 1 = colour)
  long long t_issue[kSStages];  // profiling aid: when the producer issued the copy into each staging slot
  uint32_t tmem_base;
  uint32_t pad;
};
constexpr uint32_t kSmemBytes = kOffScratch + sizeof(Scratch);
static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");
static_assert(kCpIssuers == 1 || kSStages % kCpIssuers == 0, "every copy issuer must see every phase of the staging barriers it waits on");

struct Step {
  int aux_ksteps;  // K-steps (16) taken from AUX
  int h_ksteps;    // K-steps taken from H
  int aux_first;   // 1: AUX K-steps precede the H K-steps in the chunk stream, 0: they follow
  int bias_off;    // offset into the plain-order bias array
  int post;        // work after this step's epilogue, under the next step's MMAs: 1 = colour inputs
                   // E0|D into AUX, 2 = next tile's prologue
  int head;        // 1 = distance / aux heads follow this layer, 2 = colour head
  int colour;      // layer of the colour trunk (value rows only in images-only mode)
};

struct Tc2Params {
  FieldParams f;
  int n_steps;
  int chunks_per_tile;  // per CTA: one chunk per two K-steps
  Step step[kMaxSteps];
  const unsigned char* w;  // packed chunks: [(chunk, rank)] x kWChunkBytes in consumption order
  const float* bias;       // [n_hidden][256]
  const float* w_head;     // [256][8]: ddf, aux, r, g, b, 0, 0, 0
  float* ws;               // [CTA][contributor = 4 * rank + lane quarter][local sample][3 * row type + output] head partial sums
  int* status;
  int eval;             // 1 = images only
  long long* timeline;  // optional: CTA 0 writes 6 values per step
  int timeline_cap;
  int debug;            // NEDDF_TC2_DEBUG (timing experiments; 4..64 make the results garbage): 1 = one asm statement per MMA,
                        // 2 = delay the first MMA of every step, 4 = MMA warp does not wait for weight chunks, 8 = loaders skip
                        // the L2 reads, 16 = epilogue skips its math and stores, 32 = no remote stores (peer rows stay stale),
                        // 64 = no MMAs
  int group;            // chunks per two-pass group (1..kARing)
  int col8;             // == 8: TMEM column offsets are formed at run time (see tmem column note in the epilogue)
  float* dump;          // debugging aid: cluster 0 dumps AUX (hi) and the accumulators of (tile 0, step dump_step)
  int dump_step;
};

// row of local sample s, type j
__host__ __device__ __forceinline__ int row_of(int s, int j) { return 64 * (s >> 4) + 16 * j + (s & 15); }

// write the rows (value, Jx, Jy, Jz) of local sample s at K index k into an operand buffer pair
__device__ __forceinline__ void store_sample2(unsigned char* hi_buf, unsigned char* lo_buf, int KC, int s, int k,
                                              float v0, float v1, float v2, float v3, __half2& bad, int rows = 4) {
  uint32_t h0, l0, h1, l1;
  split2h(v0, v1, h0, l0, bad);
  split2h(v2, v3, h1, l1, bad);
  const uint32_t off = act_off(row_of(s, 0), k, KC);
  const uint32_t tstride = (uint32_t)(2 * KC * 16);  // 16 rows = 2 row groups
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

// One weight chunk (two K-steps), one pass: for each K-step D += A_hi*B_hi + A_lo*B_hi + A_hi*B_lo over
// both CTAs (M = 256).  A of K-step u at tensor-memory columns a + 16u (hi) / a + 16u + 8 (lo); the B
// descriptors advance by 16 (256 bytes) per K-step.  On the second pass the commit that frees the ring stage
// (the leader's barrier `bar`).
template <bool COMMIT>
__device__ __forceinline__ void chunk_mma2_elect(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_hi, uint64_t b_lo,
                                                 uint32_t idesc, uint32_t accumulate, uint32_t bar) {
  if (COMMIT) {
    asm volatile(
        "{\n"
        ".reg .pred p, q, t;\n"
        ".reg .b32 a1, a2, a3;\n"
        ".reg .b64 bh1, bl1;\n"
        ".reg .b16 mlo, mhi;\n"
        "elect.sync _|q, 0xffffffff;\n"
        "setp.ne.b32 p, %5, 0;\n"
        "setp.ne.b32 t, %8, 0;\n"
        "add.u32 a1, %1, 8;\n"
        "add.u32 a2, %1, 16;\n"
        "add.u32 a3, %1, 24;\n"
        "add.u64 bh1, %2, 16;\n"
        "add.u64 bl1, %3, 16;\n"
        "mov.b32 {mlo, mhi}, %9;\n"
        "@q tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %4, {%7, %7, %7, %7, %7, %7, %7, %7}, p;\n"
        "@q tcgen05.mma.cta_group::2.kind::f16 [%0], [a1], %2, %4, {%7, %7, %7, %7, %7, %7, %7, %7}, t;\n"
        "@q tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %3, %4, {%7, %7, %7, %7, %7, %7, %7, %7}, t;\n"
        "@q tcgen05.mma.cta_group::2.kind::f16 [%0], [a2], bh1, %4, {%7, %7, %7, %7, %7, %7, %7, %7}, t;\n"
        "@q tcgen05.mma.cta_group::2.kind::f16 [%0], [a3], bh1, %4, {%7, %7, %7, %7, %7, %7, %7, %7}, t;\n"
        "@q tcgen05.mma.cta_group::2.kind::f16 [%0], [a2], bl1, %4, {%7, %7, %7, %7, %7, %7, %7, %7}, t;\n"
        "@q tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%6], mlo;\n"
        "}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_hi), "l"(b_lo), "r"(idesc), "r"(accumulate), "r"(bar), "r"(0u), "r"(1u), "r"(1u)
        : "memory");
  } else {
    asm volatile(
        "{\n"
        ".reg .pred p, q, t;\n"
        ".reg .b32 a1, a2, a3;\n"
        ".reg .b64 bh1, bl1;\n"
        "elect.sync _|q, 0xffffffff;\n"
        "setp.ne.b32 p, %5, 0;\n"
        "setp.ne.b32 t, %7, 0;\n"
        "add.u32 a1, %1, 8;\n"
        "add.u32 a2, %1, 16;\n"
        "add.u32 a3, %1, 24;\n"
        "add.u64 bh1, %2, 16;\n"
        "add.u64 bl1, %3, 16;\n"
        "@q tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %4, {%6, %6, %6, %6, %6, %6, %6, %6}, p;\n"
        "@q tcgen05.mma.cta_group::2.kind::f16 [%0], [a1], %2, %4, {%6, %6, %6, %6, %6, %6, %6, %6}, t;\n"
        "@q tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %3, %4, {%6, %6, %6, %6, %6, %6, %6, %6}, t;\n"
        "@q tcgen05.mma.cta_group::2.kind::f16 [%0], [a2], bh1, %4, {%6, %6, %6, %6, %6, %6, %6, %6}, t;\n"
        "@q tcgen05.mma.cta_group::2.kind::f16 [%0], [a3], bh1, %4, {%6, %6, %6, %6, %6, %6, %6, %6}, t;\n"
        "@q tcgen05.mma.cta_group::2.kind::f16 [%0], [a2], bl1, %4, {%6, %6, %6, %6, %6, %6, %6, %6}, t;\n"
        "}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_hi), "l"(b_lo), "r"(idesc), "r"(accumulate), "r"(0u), "r"(1u)
        : "memory");
  }
}

// shared memory -> tensor memory in both CTAs: 128 lanes x 8 columns from a 4 KB K-major core-matrix block
__device__ __forceinline__ void tmem_cp2_128x256b(uint32_t taddr, uint64_t sdesc) {
  asm volatile("tcgen05.cp.cta_group::2.128x256b [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory");
}
// commit without elect (a single thread issues): all prior async tcgen05 operations of this thread
__device__ __forceinline__ void mma2_commit(uint32_t bar, uint32_t mask) {
  asm volatile(
      "{\n"
      ".reg .b16 lo, hi;\n"
      "mov.b32 {lo, hi}, %1;\n"
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], lo;\n"
      "}\n" ::"r"(bar),
      "r"(mask)
      : "memory");
}
__device__ __forceinline__ void tile_geometry2(const FieldParams& p, Scratch* sc, int64_t n0, int s, int64_t n_total) {
  float pos[3] = {0.f, 0.f, 0.f}, dir[3] = {0.f, 0.f, 1.f}, var[3] = {0.f, 0.f, 0.f};
  const int64_t n = n0 + s;
  if (n < n_total) {
    if (p.dists) {
      int64_t b, out;
      int j;
      field_map(p, n, b, j, out);
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

// position embedding of local sample s into AUX; scaled = distance-trunk scaling (neddf.py:200-204)
// else plain (neddf.py:205-209)
__device__ __forceinline__ void write_pos_embedding2(const FieldParams& p, const Scratch* sc, unsigned char* aux_hi,
                                                     unsigned char* aux_lo, int s, int sub, int nsub, bool scaled,
                                                     __half2& bad, int rows = 4) {
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
    store_sample2(aux_hi, aux_lo, kAuxK, s, idx, vs[0], vs[1], vs[2], vs[3], bad, rows);
    store_sample2(aux_hi, aux_lo, kAuxK, s, half + idx, vc[0], vc[1], vc[2], vc[3], bad, rows);
  }
}

// ---------------------------------------------------------------------------------------------
// the megakernel
// ---------------------------------------------------------------------------------------------
template <int ACT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1) field_tc2_kernel(const __grid_constant__ Tc2Params P) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const FieldParams& p = P.f;
  unsigned char* h_hi = smem + kOffHHi;
  unsigned char* h_lo = smem + kOffHLo;
  unsigned char* aux_hi = smem + kOffAuxHi;
  unsigned char* aux_lo = smem + kOffAuxLo;
  Scratch* sc = reinterpret_cast<Scratch*>(smem + kOffScratch);

  const int tid = threadIdx.x;
  // warp index, tile count and tensor-memory base through a shuffle: provably warp-uniform, so the single-thread
  // roles (MMA issuer, copy issuers, producer) compile to uniform-datapath code (see field_tc.cu)
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const int64_t cid = blockIdx.x >> 1;
  const int64_t n_clusters = gridDim.x >> 1;

  const int64_t n_total = field_total(p);
  const int64_t n_tiles = (n_total + kPairS - 1) / kPairS;
  int64_t my_tiles = 0;
  if (cid < n_tiles) my_tiles = (n_tiles - 1 - cid) / n_clusters + 1;
  my_tiles = __shfl_sync(0xffffffffu, my_tiles, 0);
  const int64_t total_chunks = my_tiles * P.chunks_per_tile;

  if (tid == 0) {
    for (int i = 0; i < kSStages; ++i) {
      mbar_init(&sc->s_full[i], 1);     // expect_tx arrival of the producer + the TMA bytes
      mbar_init(&sc->s_empty[i], 1);    // tcgen05.commit (multicast)
      mbar_init(&sc->peer_full[i], 1);  // the peer's forwarder
    }
    for (int i = 0; i < kARing; ++i) {
      mbar_init(&sc->a_full[i], 1);   // tcgen05.commit of the copies
      mbar_init(&sc->a_empty[i], 1);  // tcgen05.commit of the second MMA pass
    }
    for (int h = 0; h < 2; ++h) {
      mbar_init(&sc->act_ready[h], 2);              // one arrival per CTA (after a barrier of its epilogue threads)
      mbar_init(&sc->acc_ready[h], 1);
      mbar_init(&sc->head_ready[h], 2);             // one arrival per CTA: its partial sums for this CTA's samples are written
    }
    for (int h = 0; h < 2; ++h) {
      mbar_init(&sc->norm_ready[h], 2 * 16);          // the 16 lanes (samples of half h) of the finishing warp of both CTAs
      mbar_init(&sc->fin_done[h], 2);                 // lane 0 of the finishing warp of both CTAs
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kMmaWarp) tmem_alloc2(&sc->tmem_base, kTmemCols);
  fence_async_smem();
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, sc->tmem_base, 0);

  if (warp == kTmaWarp) {
    // ===================== TMA producer: this CTA's chunks, L2 -> shared-memory staging ring =========
    if (lane == 0) {
      const unsigned char* src = P.w + (size_t)rank * kWChunkBytes;
      int idx = 0;  // chunk within the tile
      for (int64_t c = 0; c < total_chunks; ++c) {
        const int s = (int)(c % kSStages);
        if (c >= kSStages) mbar_wait(&sc->s_empty[s], (uint32_t)((c / kSStages - 1) & 1));
        if (P.debug & 128) sc->t_issue[s] = clock64();
        mbar_expect_tx(&sc->s_full[s], kWChunkBytes);
        // several smaller copies on one barrier: their L2 round trips overlap (one 16 KB request is served serially)
#pragma unroll
        for (int u = 0; u < kTmaSplit; ++u)
          tma_bulk_g2s(smem_u32(smem + kOffStage + s * kWChunkBytes + u * (kWChunkBytes / kTmaSplit)),
                       src + (size_t)idx * (2 * kWChunkBytes) + u * (kWChunkBytes / kTmaSplit), kWChunkBytes / kTmaSplit,
                       &sc->s_full[s]);
        if (++idx == P.chunks_per_tile) idx = 0;
      }
    }
  } else if (warp == kCpWarp || warp == kCpWarp2) {
    if (lane == 0) {
      const int64_t c_first = (warp == kCpWarp) ? 0 : 1;
      if (rank != 0) {
        // ===================== peer: tell the leader when a chunk has landed here ==================
        const uint32_t pf0 = mapa_u32(smem_u32(&sc->peer_full[0]), 0);
        for (int64_t c = c_first; c < total_chunks; c += kCpIssuers) {
          const int s = (int)(c % kSStages);
          mbar_wait(&sc->s_full[s], (uint32_t)((c / kSStages) & 1));
          mbar_arrive_cluster_relaxed(pf0 + 8 * s);  // the data stays in this CTA; its tcgen05.cp is ordered by the TMA completion observed here
        }
      } else {
        // ===================== leader: staged chunks of both CTAs -> tensor-memory ring ============
        const uint32_t st0 = smem_u32(smem + kOffStage);
        for (int64_t c = c_first; c < total_chunks; c += kCpIssuers) {
          const int s = (int)(c % kSStages), t = (int)(c % kARing);
          const uint32_t ph = (uint32_t)((c / kSStages) & 1);
          const bool cstamp = P.timeline && (P.debug & 128) && blockIdx.x == 0 && (c + 1) * 6 <= P.timeline_cap;
          const long long ct0 = cstamp ? clock64() : 0;
          mbar_wait(&sc->s_full[s], ph);
          const long long ct1 = cstamp ? clock64() : 0;
          const long long ti = cstamp ? sc->t_issue[s] : 0;
          mbar_wait(&sc->peer_full[s], ph);
          const long long ct2 = cstamp ? clock64() : 0;
          if (c >= kARing) mbar_wait(&sc->a_empty[t], (uint32_t)((c / kARing - 1) & 1));
          tc_fence_after();
          if (cstamp) {
            P.timeline[6 * c + 0] = ct0;          // start waiting for this chunk
            P.timeline[6 * c + 1] = ti;           // producer issued its TMA
            P.timeline[6 * c + 2] = ct1;          // landed here
            P.timeline[6 * c + 3] = ct2;          // peer's landed
            P.timeline[6 * c + 4] = clock64();    // tensor-memory stage free
          }
          const uint32_t ta = tmem + kACol + t * (16 * kChunkK);
          // per K-step: 4 KB hi -> columns +0..7, 4 KB lo -> columns +8..15 (K-major core matrices: LBO 128, SBO 256)
#pragma unroll
          for (int u = 0; u < 2 * kChunkK; ++u)
            tmem_cp2_128x256b(ta + 8 * u, make_desc(st0 + s * kWChunkBytes + u * 4096, 128, 256));
          const long long cm0 = cstamp ? clock64() : 0;
          mma2_commit(smem_u32(&sc->a_full[t]), 1);   // MMA warp: the chunk is in tensor memory (both CTAs)
          const long long cm1 = cstamp ? clock64() : 0;
          mma2_commit(smem_u32(&sc->s_empty[s]), 3);  // producers of both CTAs: the staging slot is free
          if (cstamp) P.timeline[6 * c + 5] = ((clock64() - cm1) << 40) | ((cm1 - cm0) << 20) | (cm0 - P.timeline[6 * c + 4]);
        }
      }
    }
  } else if (warp == kMmaWarp) {
    if (rank == 0) {
      // ===================== MMA issuer (leader CTA; whole warp, one elected lane issues) =========
      const uint32_t s_hhi = smem_u32(h_hi), s_ahi = smem_u32(aux_hi), s_hlo = smem_u32(h_lo), s_alo = smem_u32(aux_lo);
      int stage = 0;          // ring position of the next chunk
      uint32_t full_par = 0;  // parity to wait for on a_full[stage]
      uint32_t act_phase = 0;
      uint32_t norm_phase = 0;
      int mst = 0;  // profiling stamps of the MMA warp (NEDDF_TC2_DEBUG & 512)
      for (int64_t t = 0; t < my_tiles; ++t) {
        for (int si = 0; si < P.n_steps; ++si) {
          const Step& st = P.step[si];
          const int tl = (int)(t * P.n_steps + si);
          const bool stamp = P.timeline && !(P.debug & (128 | 512)) && blockIdx.x == 0 && lane == 0 && (tl + 1) * 6 <= P.timeline_cap;
          const bool value_only = P.eval && st.colour;
          const uint32_t idesc = value_only ? make_idesc(256, 2 * 16, 0, 1) : make_idesc(256, 2 * 64, 0, 1);
          // in chunks of kChunkK K-steps (AUX and H extents are multiples of it)
          const int nA = st.aux_ksteps / kChunkK, nH = st.h_ksteps / kChunkK, nT = nA + nH;
          // chunk i of the stream reads AUX or H: [0, nA) AUX then H when aux_first, else H then AUX
          const int first_n = st.aux_first ? nA : nH;
          // two-pass groups of at most NEDDF_TC2_GROUP chunks (the ring holds a whole group)
          const int n_groups = (nT + P.group - 1) / P.group;
          const int g_len = (nT + n_groups - 1) / n_groups;
          const bool needs_norm = st.colour && st.aux_ksteps > 0;  // colour layer 0 reads the normals from AUX
          if (stamp) P.timeline[6 * tl + 0] = clock64();
          for (int gb = 0; gb < nT; gb += g_len) {
            const int ge = (gb + g_len < nT) ? gb + g_len : nT;
            const int stage0 = stage;
            const uint32_t par0 = full_par;
            for (int hs = 0; hs < 2; ++hs) {
              if (gb == 0) {  // accumulator hs drained by the previous step's epilogue, B rows of half hs rewritten
                mbar_wait(&sc->act_ready[hs], act_phase);
                tc_fence_after();
              }
              const uint32_t d = tmem + 128 * hs;
              // descriptors of half hs: 8 row groups from row group 8 hs
              const uint64_t dba_hi = make_desc(s_ahi + hs * 8 * (kAuxK * 16), 128, kAuxK * 16);
              const uint64_t dba_lo = make_desc(s_alo + hs * 8 * (kAuxK * 16), 128, kAuxK * 16);
              const uint64_t dbh_hi = make_desc(s_hhi + hs * 8 * (kHK * 16), 128, kHK * 16);
              const uint64_t dbh_lo = make_desc(s_hlo + hs * 8 * (kHK * 16), 128, kHK * 16);
              stage = stage0;
              full_par = par0;
              uint32_t acc = gb > 0;
              bool norm_waited = !needs_norm;
              if ((P.debug & 2) && gb == 0) {
                const long long t0 = clock64();
                while (clock64() - t0 < 200000) {}
                tc_fence_after();
              }
              auto run = [&](uint64_t db_hi, uint64_t db_lo, int n) {
                for (int i = 0; i < n; ++i) {
                  const bool mstamp = P.timeline && (P.debug & 512) && blockIdx.x == 0 && lane == 0 && t == 1 && (mst + 1) * 4 <= P.timeline_cap;
                  const long long mt0 = mstamp ? clock64() : 0;
                  if (P.debug & 1) {
                    if (hs == 0) {
                      mbar_wait(&sc->a_full[stage], full_par);
                      tc_fence_after();
                    }
                    const uint32_t ta = tmem + kACol + stage * (16 * kChunkK);
#pragma unroll
                    for (int u = 0; u < kChunkK; ++u) {
                      mma2_f16_ts_elect(d, ta + 16 * u, db_hi + 16 * u, idesc, u ? 1u : acc);
                      mma2_f16_ts_elect(d, ta + 16 * u + 8, db_hi + 16 * u, idesc, 1);
                      mma2_f16_ts_elect(d, ta + 16 * u, db_lo + 16 * u, idesc, 1);
                    }
                    if (hs == 1) mma2_commit_elect(smem_u32(&sc->a_empty[stage]), 1);
                  } else if (hs == 0) {
                    if (!(P.debug & 4)) mbar_wait(&sc->a_full[stage], full_par);
                    tc_fence_after();
                    if (!(P.debug & 64)) chunk_mma2_elect<false>(d, tmem + kACol + stage * (16 * kChunkK), db_hi, db_lo, idesc, acc, 0);
                  } else if (P.debug & 64) {
                    mma2_commit_elect(smem_u32(&sc->a_empty[stage]), 1);
                  } else {
                    chunk_mma2_elect<true>(d, tmem + kACol + stage * (16 * kChunkK), db_hi, db_lo, idesc, acc, smem_u32(&sc->a_empty[stage]));
                  }
                  if (mstamp) {
                    P.timeline[4 * mst + 0] = mt0;
                    P.timeline[4 * mst + 1] = clock64();
                    P.timeline[4 * mst + 2] = si * 16 + hs * 8 + stage;
                    ++mst;
                  }
                  acc = 1;
                  db_hi += 16 * kChunkK;  // 16 K = 256 bytes = 16 descriptor units per K-step
                  db_lo += 16 * kChunkK;
                  if (++stage == kARing) {
                    stage = 0;
                    full_par ^= 1;
                  }
                }
              };
              // part 1 = K-steps [0, first_n) of the stream, part 2 = the rest
              const int b1 = gb < first_n ? gb : first_n, e1 = ge < first_n ? ge : first_n;
              const int b2 = gb > first_n ? gb : first_n, e2 = ge > first_n ? ge : first_n;
              const bool part1_aux = st.aux_first != 0;
              if (b1 < e1) {
                if (part1_aux) {
                  run(dba_hi + b1 * (16 * kChunkK), dba_lo + b1 * (16 * kChunkK), e1 - b1);
                } else {
                  run(dbh_hi + b1 * (16 * kChunkK), dbh_lo + b1 * (16 * kChunkK), e1 - b1);
                }
              }
              if (b2 < e2) {
                if (!part1_aux) {
                  if (!norm_waited) {  // the surface normals (AUX K rows n_e0 + n_d ..) of both CTAs
                    mbar_wait(&sc->norm_ready[hs], norm_phase);
                    tc_fence_after();
                    norm_waited = true;
                  }
                  run(dba_hi + (b2 - first_n) * (16 * kChunkK), dba_lo + (b2 - first_n) * (16 * kChunkK), e2 - b2);
                } else {
                  run(dbh_hi + (b2 - first_n) * (16 * kChunkK), dbh_lo + (b2 - first_n) * (16 * kChunkK), e2 - b2);
                }
              }
              if (ge == nT) mma2_commit_elect(smem_u32(&sc->acc_ready[hs]), 3);
            }
          }
          act_phase ^= 1;
          if (needs_norm) norm_phase ^= 1;
          if (stamp) P.timeline[6 * tl + 1] = clock64();
        }
      }
    }
  } else if (warp == kFinWarp) {
    // ===================== finishing warp: heads -> density / normals / colour / penalties / outputs ========
    // lane = local sample of this CTA; the two sample halves are finished as their sums arrive, so the
    // normals of half 0 reach AUX while the last trunk layer still works on half 1.
    const int s = lane, hs_own = lane >> 4;
    const float* hsum = P.ws + (size_t)blockIdx.x * kWsFloats;  // [contributor][sample][12]
    const uint32_t norm0 = mapa_u32(smem_u32(&sc->norm_ready[0]), 0);
    const uint32_t fin_c0 = mapa_u32(smem_u32(&sc->fin_done[0]), 0), fin_c1 = mapa_u32(smem_u32(&sc->fin_done[0]), 1);
    __half2 bad = __floats2half2_rn(0.f, 0.f);
    uint32_t head_phase = 0;
    HeadOut ho;  // of sample `lane`, kept from the distance heads to the colour head
    for (int64_t t = 0; t < my_tiles; ++t) {
      const int64_t tile = cid + t * n_clusters;
      const int64_t n = tile * kPairS + kTileS * rank + s;
      // ---- distance / aux heads (neddf.py:212-241), normals into AUX for the colour trunk (:243-253) ----
      for (int hs = 0; hs < 2; ++hs) {
        mbar_wait(&sc->head_ready[hs], head_phase);
        fence_acq_rel_cluster();  // the sums were written by both CTAs (global memory, read from L2)
        if (hs == hs_own) {
          float ddf[4] = {0.f, 0.f, 0.f, 0.f}, aux[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const float4* q = reinterpret_cast<const float4*>(hsum + ((size_t)c * kTileS + s) * 12);
            const float4 q0 = __ldcg(q), q1 = __ldcg(q + 1), q2 = __ldcg(q + 2);
            ddf[0] += q0.x; aux[0] += q0.y;
            ddf[1] += q0.w; aux[1] += q1.x;
            ddf[2] += q1.z; aux[2] += q1.w;
            ddf[3] += q2.y; aux[3] += q2.z;
          }
          ddf[0] += __ldg(p.b_head + 0);
          aux[0] += __ldg(p.b_head + 1);
          head_density(ddf, aux, p.d_near, p.aux_grad_scale, p.density_act, ho);
          const int kn = p.n_e0 + p.n_d;  // normal: detached, zero Jacobian
#pragma unroll
          for (int i = 0; i < 3; ++i)
            store_sample2(aux_hi, aux_lo, kAuxK, s, kn + i, ho.normal[i], 0.f, 0.f, 0.f, bad, P.eval ? 1 : 4);
          fence_async_all();
          mbar_arrive_cluster(norm0 + 8 * hs);  // every lane after its own writes
        }
      }
      head_phase ^= 1;
      __syncwarp();
      if (lane == 0) {
        mbar_arrive_cluster(fin_c0);
        mbar_arrive_cluster(fin_c1);
      }
      // ---- colour head (neddf.py:257) + penalties (:259-300) + outputs ----
      for (int hs = 0; hs < 2; ++hs) {
        mbar_wait(&sc->head_ready[hs], head_phase);
        fence_acq_rel_cluster();
        if (hs == hs_own) {
          float col[3] = {0.f, 0.f, 0.f}, colJ[3][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const float4* q = reinterpret_cast<const float4*>(hsum + ((size_t)c * kTileS + s) * 12);
            const float4 q0 = __ldcg(q);
            col[0] += q0.x; col[1] += q0.y; col[2] += q0.z;
            if (!P.eval) {
              const float4 q1 = __ldcg(q + 1), q2 = __ldcg(q + 2);
              colJ[0][0] += q0.w; colJ[0][1] += q1.x; colJ[0][2] += q1.y;
              colJ[1][0] += q1.z; colJ[1][1] += q1.w; colJ[1][2] += q2.x;
              colJ[2][0] += q2.y; colJ[2][1] += q2.z; colJ[2][2] += q2.w;
            }
          }
          if (n < n_total) {
#pragma unroll
            for (int o = 0; o < 3; ++o) col[o] += __ldg(p.b_head + 2 + o);
            int64_t ray_, on;  // where this sample's outputs go (segment view: [ray, edge] of the full arrays)
            int j_;
            field_map(p, n, ray_, j_, on);
            if (p.distance) p.distance[on] = ho.distance;
            if (p.density) p.density[on] = ho.density;
            if (p.aux_grad) p.aux_grad[on] = ho.aux;
            if (p.color) {
              p.color[3 * on + 0] = col[0];
              p.color[3 * on + 1] = col[1];
              p.color[3 * on + 2] = col[2];
            }
            // (in images-only mode the colour Jacobian rows are not computed and no penalty is asked)
            if (p.penalty) p.penalty[on] = field_penalty(ho, col, colJ, p.distance_range_max, p.penalty_weight);
          }
        }
      }
      head_phase ^= 1;
      __syncwarp();
      if (lane == 0) {
        mbar_arrive_cluster(fin_c0 + 8);
        mbar_arrive_cluster(fin_c1 + 8);
      }
    }
    {
      const float2 m = __half22float2(bad);
      if (!(fmaxf(m.x, m.y) < 65504.0f) && P.status) atomicOr(P.status, 4);
    }
  } else {
    // ===================== epilogue warps =====================================================
    const int quarter = warp & 3;  // TMEM lane quarter this warp may access
    const int blk = warp >> 2;     // sub-block: samples of CTA `tc`, 8 samples `sg` of each half
    const uint32_t tcta = (uint32_t)(blk >> 1);
    const int sg = blk & 1;
    const int chl = 32 * quarter + lane;     // channel within this CTA's half
    const int ch = 128 * (int)rank + chl;    // output channel = K index of the next layer
    const uint32_t lane_addr = (uint32_t)(32 * quarter) << 16;
    const bool remote = tcta != rank;
    const bool skip_remote = remote && (P.debug & 32);
    // destination operand buffers: own shared memory or the peer's
    const uint32_t dst_hhi = mapa_u32(smem_u32(h_hi), tcta), dst_hlo = mapa_u32(smem_u32(h_lo), tcta);
    // head partial sums of this warp go to the owner of its samples (global workspace, L2)
    float* dst_hsum = P.ws + ((size_t)(2 * cid + tcta) * 8 + 4 * rank + quarter) * (kTileS * 12);
    const uint32_t act0 = mapa_u32(smem_u32(&sc->act_ready[0]), 0);
    __half2 bad = __floats2half2_rn(0.f, 0.f);  // max |operand hi part| seen (fp16 range check)
    uint32_t acc_phase = 0;
    uint32_t fin_phase[2] = {0, 0};
    // head weights of this thread's channel: ddf, aux, r, g, b
    float wh[5];
    {
      const float4 a = __ldg(reinterpret_cast<const float4*>(P.w_head + 8 * ch));
      wh[0] = a.x; wh[1] = a.y; wh[2] = a.z; wh[3] = a.w;
      wh[4] = __ldg(P.w_head + 8 * ch + 4);
    }

    auto prologue = [&](int64_t tile) {
      const int64_t n0 = tile * kPairS + kTileS * rank;
      if (tid < kTileS) tile_geometry2(p, sc, n0, tid, n_total);
      asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
      const int s = tid >> 4, sub = tid & 15;
      write_pos_embedding2(p, sc, aux_hi, aux_lo, s, sub, 16, true, bad);
      for (int k = p.n_e0 + sub; k < 64; k += 16) store_sample2(aux_hi, aux_lo, kAuxK, s, k, 0.f, 0.f, 0.f, 0.f, bad);
    };
    // colour-trunk inputs E0 | D (| zero pad) into AUX (neddf.py:205-210, 243); 16 threads per sample
    auto colour_prep = [&]() {
      const int s = tid >> 4, sub = tid & 15;
      const int rows = P.eval ? 1 : 4;
      write_pos_embedding2(p, sc, aux_hi, aux_lo, s, sub, 16, false, bad, rows);
      const int dhalf = 3 * p.embed_dir;
      for (int idx = sub; idx < dhalf; idx += 16) {
        int e = idx / 3, d = idx - 3 * e;
        float sn, cs;
        sincosf((float)(1u << e) * sc->geo[s][3 + d], &sn, &cs);
        store_sample2(aux_hi, aux_lo, kAuxK, s, p.n_e0 + idx, sn, 0.f, 0.f, 0.f, bad, rows);
        store_sample2(aux_hi, aux_lo, kAuxK, s, p.n_e0 + dhalf + idx, cs, 0.f, 0.f, 0.f, bad, rows);
      }
      for (int k = p.n_e0 + p.n_d + 3 + sub; k < kAuxK; k += 16)
        store_sample2(aux_hi, aux_lo, kAuxK, s, k, 0.f, 0.f, 0.f, 0.f, bad, rows);
    };
    // Publish the epilogue threads' shared-memory writes (own and peer CTA) to the tensor cores and signal
    // mbarriers (given by shared::cluster address; 0 = none).  Every thread fences its own writes towards the
    // async proxy, a hardware named barrier orders the 512 threads, one thread arrives with release at
    // cluster scope.  (A warp-level variant - fence, __syncwarp, lane 0 arrives - let the MMAs read operand
    // rows before they were visible: tools/tc2_debug.py.)
    auto cta_arrive = [&](int bar_id, uint32_t cbar0, uint32_t cbar1, uint32_t cbar2) {
      fence_async_all();
      asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(kEpiThreads) : "memory");
      if (tid == 0) {
        if (P.debug & 256) {  // experiment: CTA-scope release only (no MEMBAR.ALL.GPU)
          if (cbar0) mbar_arrive_cluster_default(cbar0);
          if (cbar1) mbar_arrive_cluster_default(cbar1);
          if (cbar2) mbar_arrive_cluster_default(cbar2);
        } else {
          fence_acq_rel_cluster();  // one release for the (up to) three signals
          if (cbar0) mbar_arrive_cluster_relaxed(cbar0);
          if (cbar1) mbar_arrive_cluster_relaxed(cbar1);
          if (cbar2) mbar_arrive_cluster_relaxed(cbar2);
        }
      }
    };
    const uint32_t head_ready_c0 = mapa_u32(smem_u32(&sc->head_ready[0]), 0), head_ready_c1 = mapa_u32(smem_u32(&sc->head_ready[0]), 1);

    if (my_tiles > 0) {
      prologue(cid);
      cta_arrive(2, act0, act0 + 8, 0);
    }

    for (int64_t t = 0; t < my_tiles; ++t) {
      const int64_t tile = cid + t * n_clusters;
      for (int si = 0; si < P.n_steps; ++si) {
        const Step& st = P.step[si];
        const float bias = __ldg(P.bias + st.bias_off + ch);
        const bool value_only = P.eval && st.colour;
        const bool last = (t + 1 == my_tiles) && (si + 1 == P.n_steps);
        const int tl = (int)(t * P.n_steps + si);
        const bool stamp = P.timeline && !(P.debug & (128 | 512)) && blockIdx.x == 0 && tid == 0 && (tl + 1) * 6 <= P.timeline_cap;
        for (int hs = 0; hs < 2; ++hs) {
          if (lane == 0) {
            mbar_wait(&sc->acc_ready[hs], acc_phase);
            // the previous tile's sums of this head type were consumed by the finishing warps of both CTAs
            if (st.head != 0 && hs == 0 && t > 0) mbar_wait(&sc->fin_done[st.head - 1], fin_phase[st.head - 1]);
          }
          __syncwarp();
          tc_fence_after();
          if (stamp) P.timeline[6 * tl + 2 + 2 * hs] = clock64();
          if (P.dump && cid == 0 && t == 0 && si == P.dump_step) {
            // [rank][ AUX hi words 6144 | acc: hs x 128 lanes x 128 columns ]
            float* base = P.dump + (size_t)rank * (6144 + 2 * 128 * 128);
            if (hs == 0)
              for (int w = tid; w < 6144; w += kEpiThreads) base[w] = __uint_as_float(reinterpret_cast<const uint32_t*>(aux_hi)[w]);
            if (blk == 0) {  // the probe's access pattern: one warp per lane quarter, x16 loads, run-time loop
              for (int cb = 0; cb < 8; ++cb) {
                float v[16];
                tmem_ld16(tmem + lane_addr + 128 * hs + cb * 16, v);
#pragma unroll
                for (int i = 0; i < 16; ++i) base[6144 + ((size_t)hs * 128 + 32 * quarter + lane) * 128 + cb * 16 + i] = v[i];
              }
            }
          }
          if (!(P.debug & 16)) {
          // this thread: channel ch, samples (CTA tcta, 16 hs + 8 sg + i), i = 0..7
          const uint32_t tbase = tmem + lane_addr + 128 * hs + (value_only ? 16 * tcta + 8 * sg : 64 * tcta + 8 * sg);
          const int s0 = 16 * hs + 8 * sg;                                         // local sample of i = 0
          const int64_t ng0 = tile * kPairS + kTileS * (int64_t)tcta + s0;         // its global index
          float x[8], d1[8], gj[3][8];
          if (value_only) tmem_ld8(tbase, x);
          else tmem_ld8x4(tbase, tbase + 2 * P.col8, tbase + 4 * P.col8, tbase + 6 * P.col8, x, gj[0], gj[1], gj[2]);  // one wait for the four row types
          float* save = nullptr;  // training: pre-activations [layer][sample][row type][channel]
          if (p.save_pre) save = p.save_pre + (((size_t)st.bias_off / kWidth * p.n + ng0) * 4) * kWidth + ch;
          if (save) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
              if (ng0 + i < n_total) save[(size_t)i * 4 * kWidth] = x[i] + bias;
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) tc_hidden_act<ACT>(x[i] + bias, x[i], d1[i]);
          uint32_t h[4], l[4];
          uint32_t off = (uint32_t)((8 * hs + sg) * (kHK * 16) + ch * 16);  // row group of (hs, j = 0, sg)
          const bool write_h = st.head != 2;  // the last colour layer feeds only the colour head
          if (write_h) {
#pragma unroll
            for (int i = 0; i < 4; ++i) split2h(x[2 * i], x[2 * i + 1], h[i], l[i], bad);
            if (skip_remote) {
            } else if (remote) {
              st_cluster_v4(dst_hhi + off, make_uint4(h[0], h[1], h[2], h[3]));
              st_cluster_v4(dst_hlo + off, make_uint4(l[0], l[1], l[2], l[3]));
            } else {
              *reinterpret_cast<uint4*>(h_hi + off) = make_uint4(h[0], h[1], h[2], h[3]);
              *reinterpret_cast<uint4*>(h_lo + off) = make_uint4(l[0], l[1], l[2], l[3]);
            }
          }
          // heads: this channel's products with the head weights, reduced over the warp's 32 channels
          // one row type at a time and parked with the owner of the sample
          auto head_rows = [&](const float (&v)[8], int j) {
            if (st.head == 1) {
              float hv[16];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                hv[2 * i + 0] = wh[0] * v[i];
                hv[2 * i + 1] = wh[1] * v[i];
              }
              warp_transpose_reduce16(hv, lane);  // lane L: sample L >> 2, output (L >> 1) & 1
              if ((lane & 1) == 0)
                __stcg(dst_hsum + (s0 + (lane >> 2)) * 12 + 3 * j + ((lane >> 1) & 1), hv[0]);
            } else if (st.head == 2) {
              float hv[32];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                hv[4 * i + 0] = wh[2] * v[i];
                hv[4 * i + 1] = wh[3] * v[i];
                hv[4 * i + 2] = wh[4] * v[i];
                hv[4 * i + 3] = 0.f;
              }
              warp_transpose_reduce<32>(hv, lane);  // lane L: sample L >> 2, output L & 3
              if ((lane & 3) != 3) __stcg(dst_hsum + (s0 + (lane >> 2)) * 12 + 3 * j + (lane & 3), hv[0]);
            }
          };
          head_rows(x, 0);
          if (!value_only) {
#pragma unroll
            for (int j = 1; j < 4; ++j) {  // Jacobian rows: G = f'(x) J (tanh_exp.py:47-48)
              float (&g)[8] = gj[j - 1];
              if (save) {
#pragma unroll
                for (int i = 0; i < 8; ++i)
                  if (ng0 + i < n_total) save[((size_t)i * 4 + j) * kWidth] = g[i];
              }
#pragma unroll
              for (int i = 0; i < 8; ++i) g[i] *= d1[i];
              if (write_h) {
#pragma unroll
                for (int i = 0; i < 4; ++i) split2h(g[2 * i], g[2 * i + 1], h[i], l[i], bad);
                off += 2 * (kHK * 16);
                if (skip_remote) {
                } else if (remote) {
                  st_cluster_v4(dst_hhi + off, make_uint4(h[0], h[1], h[2], h[3]));
                  st_cluster_v4(dst_hlo + off, make_uint4(l[0], l[1], l[2], l[3]));
                } else {
                  *reinterpret_cast<uint4*>(h_hi + off) = make_uint4(h[0], h[1], h[2], h[3]);
                  *reinterpret_cast<uint4*>(h_lo + off) = make_uint4(l[0], l[1], l[2], l[3]);
                }
              }
              head_rows(g, j);
            }
          }
          }  // debug & 16
          tc_fence_before();
          if (stamp) P.timeline[6 * tl + 3 + 2 * hs] = clock64();
          // accumulator hs drained, operand rows of half hs rewritten in both CTAs, head sums parked
          if (st.head != 0) cta_arrive(2 + hs, last ? 0u : act0 + 8 * hs, head_ready_c0 + 8 * hs, head_ready_c1 + 8 * hs);
          else if (!last) cta_arrive(2 + hs, act0 + 8 * hs, 0, 0);
        }
        acc_phase ^= 1;
        // work that only feeds later steps runs here, under the next step's MMA phase; its
        // shared-memory writes are published by the fence + arrive of the following steps
        if (st.post == 1) colour_prep();
        if (st.post == 2 && t + 1 < my_tiles) prologue(tile + n_clusters);
        if (st.head != 0 && t > 0) fin_phase[st.head - 1] ^= 1;
      }
    }
    {
      const float2 m = __half22float2(bad);
      if (!(fmaxf(m.x, m.y) < 65504.0f) && P.status) atomicOr(P.status, 4);
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == kMmaWarp) tmem_dealloc2(tmem, kTmemCols);
}

// ---------------------------------------------------------------------------------------------
// weight packing: fp32 [in,out] -> fp16 hi/lo chunks [(K-step, CTA rank)] in consumption order
// ---------------------------------------------------------------------------------------------
struct PackArgs {
  const float* w[kMaxHidden + 3];
  const float* b[kMaxHidden + 3];
  int k_in[kMaxHidden];
  int aux_real[kMaxHidden];   // leading reference input channels that live in AUX (0 if none)
  int aux_pad[kMaxHidden];    // their padded K extent in AUX
  int aux_first[kMaxHidden];  // AUX K-steps first (1) or after the H K-steps (0)
  int ksteps[kMaxHidden];
  int chunk0[kMaxHidden];     // first per-CTA chunk index (chunks of kChunkK K-steps) of the layer
  int n_hidden;
};

__global__ void pack_hidden_kernel(PackArgs a, unsigned char* __restrict__ dst, float* __restrict__ bias) {
  const int l = blockIdx.y;
  const int total = a.ksteps[l] * 2 * 128 * 16;  // (kstep, rank, m, k)
  const int aux_ks = a.aux_pad[l] / 16, h_ks = a.ksteps[l] - aux_ks;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    int k = idx & 15, m = (idx >> 4) & 127, rank = (idx >> 11) & 1, ks = idx >> 12;
    // operand-space K index of stream K-step ks: AUX part occupies [0, aux_pad), H follows
    int ks_op = a.aux_first[l] ? ks : (ks < h_ks ? aux_ks + ks : ks - h_ks);
    int kk = ks_op * 16 + k;
    int row;  // reference weight row, -1 = zero padding
    if (kk < a.aux_pad[l]) row = (kk < a.aux_real[l]) ? kk : -1;
    else row = a.aux_real[l] + (kk - a.aux_pad[l]);
    if (row >= a.k_in[l]) row = -1;
    float w = (row >= 0) ? a.w[l][(size_t)row * kWidth + 128 * rank + m] : 0.f;
    __half hi = __float2half_rn(w);
    __half lo = __float2half_rn(w - __half2float(hi));
    // chunk = kChunkK K-steps of this rank; per K-step 4 KB hi then 4 KB lo, each a K-major core-matrix block
    // [m/8][k/8][m%8][k%8] (what tcgen05.cp.128x256b moves to lane m, columns k/2)
    unsigned char* chunk = dst + ((size_t)(a.chunk0[l] + ks / kChunkK) * 2 + rank) * kWChunkBytes + (size_t)(ks % kChunkK) * kChunkBytes;
    *reinterpret_cast<__half*>(chunk + wchunk_off(m, k)) = hi;
    *reinterpret_cast<__half*>(chunk + 4096 + wchunk_off(m, k)) = lo;
  }
  if (blockIdx.x == 0)
    for (int c = threadIdx.x; c < kWidth; c += blockDim.x) bias[l * kWidth + c] = a.b[l][c];
}

__global__ void pack_heads_kernel(PackArgs a, float* __restrict__ w_head) {
  const int nh = a.n_hidden;
  for (int k = threadIdx.x; k < kWidth; k += blockDim.x) {
    w_head[8 * k + 0] = a.w[nh + 0][k];
    w_head[8 * k + 1] = a.w[nh + 1][k];
    w_head[8 * k + 2] = a.w[nh + 2][3 * k + 0];
    w_head[8 * k + 3] = a.w[nh + 2][3 * k + 1];
    w_head[8 * k + 4] = a.w[nh + 2][3 * k + 2];
    w_head[8 * k + 5] = 0.f;
    w_head[8 * k + 6] = 0.f;
    w_head[8 * k + 7] = 0.f;
  }
}

struct Storage {
  unsigned char* d_w = nullptr;
  float* d_bias = nullptr;
  float* d_w_head = nullptr;
  int* d_status = nullptr;
  float* d_ws = nullptr;  // head partial sums, one block per CTA of the largest grid
  int n_steps = 0;
  int chunks_per_tile = 0;
  Step step[kMaxSteps];
  PackArgs pack;
  long long* timeline = nullptr;
  int timeline_cap = 0;
  float* dump = nullptr;
  int dump_step = 0;
};

}  // namespace tc2

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static bool tc2_is_skip_layer(const neddf_field* f, int l) {
  for (int i = 0; i < f->cfg.n_skips; ++i)
    if (f->cfg.skips[i] == l - 1) return true;
  return false;
}

bool tc2_supported(const neddf_field* f) {
  // AUX holds 64 K (position embedding of the distance trunk) / 96 K (colour inputs E0 | D | n)
  const int n_e0 = 6 * f->cfg.embed_pos_rank, off_h = n_e0 + 6 * f->cfg.embed_dir_rank + 3;
  return n_e0 <= 64 && off_h <= tc::kAuxK && f->cfg.ddf_layer_width == kWidth && f->cfg.col_layer_width == kWidth;
}

static int32_t tc2_ensure(neddf_field* f) {
  if (f->tc2) return NEDDF_OK;
  tc2::Storage* S = new tc2::Storage();
  const int n_hidden = f->n_ddf + f->n_col;
  int chunk = 0;
  std::memset(&S->pack, 0, sizeof(S->pack));
  S->pack.n_hidden = n_hidden;
  int last_aux = 0;
  for (int l = 0; l < n_hidden; ++l) {
    int aux_real = 0, aux_pad = 0, h_k = 0, aux_first = 1;
    if (l == 0) { aux_real = f->proto.n_e0; aux_pad = 64; }
    else if (l < f->n_ddf) { if (tc2_is_skip_layer(f, l)) { aux_real = f->proto.n_e0; aux_pad = 64; } h_k = kWidth; }
    else if (l == f->n_ddf) { aux_real = f->proto.off_h; aux_pad = tc::kAuxK; h_k = kWidth; aux_first = 0; }
    else h_k = kWidth;
    tc2::Step& st = S->step[l];
    st.aux_ksteps = aux_pad / 16;
    st.h_ksteps = h_k / 16;
    st.aux_first = aux_first;
    st.bias_off = l * kWidth;
    st.post = 0;
    st.head = (l == f->n_ddf - 1) ? 1 : (l == n_hidden - 1) ? 2 : 0;
    st.colour = l >= f->n_ddf;
    if (l < f->n_ddf && aux_pad > 0) last_aux = l;
    S->pack.k_in[l] = f->shape_in[l];
    S->pack.aux_real[l] = aux_real;
    S->pack.aux_pad[l] = aux_pad;
    S->pack.aux_first[l] = aux_first;
    S->pack.ksteps[l] = st.aux_ksteps + st.h_ksteps;
    S->pack.chunk0[l] = chunk;
    chunk += S->pack.ksteps[l] / tc2::kChunkK;  // AUX (64 / 96 K) and H (256 K) extents are multiples of 32 K
  }
  S->n_steps = n_hidden;
  S->chunks_per_tile = chunk;
  // AUX holds E_s until the last trunk layer that reads it; after that layer's epilogue the colour
  // inputs E0|D are written while later trunk layers run.  After the first colour layer's epilogue
  // AUX is dead again: the next tile's prologue goes there.
  S->step[last_aux].post = 1;
  S->step[f->n_ddf].post = 2;
  if (cudaMalloc(&S->d_w, (size_t)chunk * 2 * tc2::kWChunkBytes) != cudaSuccess ||
      cudaMalloc(&S->d_bias, (size_t)n_hidden * kWidth * sizeof(float)) != cudaSuccess ||
      cudaMalloc(&S->d_w_head, (size_t)kWidth * 8 * sizeof(float)) != cudaSuccess ||
      cudaMalloc(&S->d_status, sizeof(int)) != cudaSuccess ||
      cudaMalloc(&S->d_ws, (size_t)sm_count() * tc2::kWsFloats * sizeof(float)) != cudaSuccess) {
    cudaFree(S->d_w); cudaFree(S->d_bias); cudaFree(S->d_w_head); cudaFree(S->d_status); cudaFree(S->d_ws);
    delete S;
    return fail(NEDDF_E_CUDA, "tensor-core pair engine: cudaMalloc failed");
  }
  cudaMemset(S->d_status, 0, sizeof(int));
  cudaMemset(S->d_ws, 0, (size_t)sm_count() * tc2::kWsFloats * sizeof(float));
  f->tc2 = S;
  return NEDDF_OK;
}

void tc2_destroy(neddf_field* f) {
  if (!f->tc2) return;
  tc2::Storage* S = static_cast<tc2::Storage*>(f->tc2);
  cudaFree(S->d_w);
  cudaFree(S->d_bias);
  cudaFree(S->d_w_head);
  cudaFree(S->d_status);
  cudaFree(S->d_ws);
  delete S;
  f->tc2 = nullptr;
}

int32_t tc2_pack_weights(neddf_field* f, const float* const* d_w, const float* const* d_b, cudaStream_t s) {
  int32_t rc = tc2_ensure(f);
  if (rc != NEDDF_OK) return rc;
  tc2::Storage* S = static_cast<tc2::Storage*>(f->tc2);
  tc2::PackArgs a = S->pack;
  const int n_hidden = f->n_ddf + f->n_col;
  for (int l = 0; l < n_hidden + 3; ++l) {
    a.w[l] = d_w[l];
    a.b[l] = d_b[l];
  }
  tc2::pack_hidden_kernel<<<dim3(64, n_hidden), 256, 0, s>>>(a, S->d_w, S->d_bias);
  NEDDF_LAUNCH_CHECK();
  tc2::pack_heads_kernel<<<1, 256, 0, s>>>(a, S->d_w_head);
  NEDDF_LAUNCH_CHECK();
  return NEDDF_OK;
}

int32_t launch_field_tc2(const neddf_field* f, FieldParams& p, int flags, cudaStream_t s) {
  tc2::Storage* S = static_cast<tc2::Storage*>(f->tc2);
  if (!S) return fail(NEDDF_E_INVALID, "tensor-core pair engine: weights were never packed");
  tc2::Tc2Params P;
  P.f = p;
  P.n_steps = S->n_steps;
  P.chunks_per_tile = S->chunks_per_tile;
  for (int i = 0; i < S->n_steps; ++i) P.step[i] = S->step[i];
  P.w = S->d_w;
  P.bias = S->d_bias;
  P.w_head = S->d_w_head;
  P.status = S->d_status;
  P.ws = S->d_ws;
  P.eval = (flags == NEDDF_OUT_EVAL && p.penalty == nullptr && p.save_pre == nullptr) ? 1 : 0;
  P.timeline = S->timeline;
  P.timeline_cap = S->timeline_cap;
  P.col8 = 8;
  P.group = NEDDF_TC2_GROUP;
  if (const char* e = std::getenv("NEDDF_TC2_GROUP")) P.group = std::max(1, std::min(tc2::kARing, std::atoi(e)));
  P.debug = 0;
  if (const char* e = std::getenv("NEDDF_TC2_DEBUG")) P.debug = std::atoi(e);
  P.dump = S->dump;
  P.dump_step = S->dump_step;
  int64_t n_tiles = (p.n + tc2::kPairS - 1) / tc2::kPairS;
  int grid = 2 * (int)std::min<int64_t>(n_tiles, sm_count() / 2);
  auto launch = [&](auto kern) -> int32_t {
    NEDDF_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc2::kSmemBytes));
    kern<<<grid, tc2::kThreads, tc2::kSmemBytes, s>>>(P);
    NEDDF_LAUNCH_CHECK();
    return NEDDF_OK;
  };
  switch (p.hidden_act) {
    case NEDDF_ACT_TANHEXP: return launch(tc2::field_tc2_kernel<NEDDF_ACT_TANHEXP>);
    case NEDDF_ACT_RELU: return launch(tc2::field_tc2_kernel<NEDDF_ACT_RELU>);
    case NEDDF_ACT_LEAKYRELU: return launch(tc2::field_tc2_kernel<NEDDF_ACT_LEAKYRELU>);
  }
  return fail(NEDDF_E_INVALID, "tensor-core pair engine: unknown activation");
}

int32_t tc2_set_timeline(neddf_field* f, long long* d_buf, int cap) {
  int32_t rc = tc2_ensure(f);
  if (rc != NEDDF_OK) return rc;
  tc2::Storage* S = static_cast<tc2::Storage*>(f->tc2);
  S->timeline = d_buf;
  S->timeline_cap = d_buf ? cap : 0;
  return NEDDF_OK;
}

int32_t tc2_set_dump(neddf_field* f, float* d_buf, int step) {
  int32_t rc = tc2_ensure(f);
  if (rc != NEDDF_OK) return rc;
  tc2::Storage* S = static_cast<tc2::Storage*>(f->tc2);
  S->dump = d_buf;
  S->dump_step = step;
  return NEDDF_OK;
}

int32_t tc2_read_status(const neddf_field* f, int* out, cudaStream_t s) {
  tc2::Storage* S = static_cast<tc2::Storage*>(f->tc2);
  *out = 0;
  if (!S) return NEDDF_OK;
  NEDDF_CUDA_CHECK(cudaMemcpyAsync(out, S->d_status, sizeof(int), cudaMemcpyDeviceToHost, s));
  NEDDF_CUDA_CHECK(cudaStreamSynchronize(s));
  NEDDF_CUDA_CHECK(cudaMemsetAsync(S->d_status, 0, sizeof(int), s));
  return NEDDF_OK;
}

}  // namespace neddf
