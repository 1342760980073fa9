// NeRF field variant, training backward: CUDA thread context, weight packing and C ABI around
// the tile program of nerf_train_kernel.cuh (which is also compiled by g++ into a host emulation for the CPU tests).
// One launch recomputes the forward per 64-sample tile and walks back through the network; it leaves the operands of
// the weight-gradient GEMMs (layer inputs X, pre-activation gradients G) in global memory for neddf_wgrad.
#include "nerf_train_kernel.cuh"

#include <algorithm>
#include <cstring>

namespace neddf {
namespace nerft {

struct CudaCtx {
  int tid, block, nblocks;
  __device__ __forceinline__ void sync() { __syncthreads(); }
  __device__ __forceinline__ void cp16(void* smem, const void* gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
  }
  __device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
  __device__ __forceinline__ void cp_wait_1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }
  __device__ __forceinline__ void cp_wait_0() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
};

__global__ void __launch_bounds__(kThreads, 1) nerf_train_kernel(const __grid_constant__ Params P) {
  extern __shared__ __align__(16) float smem[];
  CudaCtx cx{(int)threadIdx.x, (int)blockIdx.x, (int)gridDim.x};
  tile_program(cx, P, smem);
}

// forward pack [k_pad][256] + bias [256] + transposed pack [kt_pad][256] of one torch nn.Linear ([out][in], [out])
__global__ void nerf_train_pack_kernel(const float* __restrict__ w, const float* __restrict__ b, int n_in, int n_out, int k_pad,
                                       int kt_pad, float* __restrict__ dst_w, float* __restrict__ dst_b, float* __restrict__ dst_wt) {
  const int stride = gridDim.x * blockDim.x, t0 = blockIdx.x * blockDim.x + threadIdx.x;
  for (int idx = t0; idx < k_pad * kW; idx += stride) dst_w[idx] = pack_fwd(w, n_in, n_out, idx / kW, idx % kW);
  for (int idx = t0; idx < kt_pad * kW; idx += stride) dst_wt[idx] = pack_bwd(w, n_in, n_out, idx / kW, idx % kW);
  if (blockIdx.x == 0)
    for (int c = threadIdx.x; c < kW; c += blockDim.x) dst_b[c] = c < n_out ? b[c] : 0.f;
}
// density [256] + bias, colour output [3][128] + 3 biases, as stored by torch
__global__ void nerf_train_pack_heads_kernel(const float* __restrict__ wd, const float* __restrict__ bd, const float* __restrict__ wc,
                                             const float* __restrict__ bc, float* __restrict__ dst_d, float* __restrict__ dst_c) {
  for (int i = threadIdx.x; i < kW; i += blockDim.x) dst_d[i] = wd[i];
  if (threadIdx.x == 0) dst_d[kW] = bd[0];
  for (int i = threadIdx.x; i < 3 * (kW / 2); i += blockDim.x) dst_c[i] = wc[i];
  if (threadIdx.x < 3) dst_c[3 * (kW / 2) + threadIdx.x] = bc[threadIdx.x];
}

}  // namespace nerft
}  // namespace neddf

using namespace neddf;

struct neddf_nerf_train {
  neddf_nerf_config_t cfg;
  int n_layers = 0;
  int shape_in[nerft::kMaxLayers + 3];
  int shape_out[nerft::kMaxLayers + 3];
  nerft::Params proto;
  float* d_w = nullptr;
  size_t w_floats = 0;
  bool packed = false;
};

extern "C" int32_t neddf_nerf_train_create(const neddf_nerf_config_t* cfg, neddf_nerf_train_t** out) {
  if (!cfg || !out) return fail(NEDDF_E_INVALID, "neddf_nerf_train_create: null argument");
  if (const char* why = nerft::unsupported(cfg)) return fail(NEDDF_E_UNSUPPORTED, std::string("neddf_nerf_train_create: ") + why);
  neddf_nerf_train* h = new neddf_nerf_train();
  h->cfg = *cfg;
  h->n_layers = nerft::layer_shapes(cfg, h->shape_in, h->shape_out);
  std::memset(&h->proto, 0, sizeof(h->proto));
  h->w_floats = nerft::build_program(cfg, h->proto);
  if (cudaMalloc(&h->d_w, h->w_floats * sizeof(float)) != cudaSuccess) {
    delete h;
    return fail(NEDDF_E_CUDA, "neddf_nerf_train_create: cudaMalloc failed");
  }
  *out = h;
  return NEDDF_OK;
}

extern "C" void neddf_nerf_train_destroy(neddf_nerf_train_t* h) {
  if (!h) return;
  cudaFree(h->d_w);
  delete h;
}

extern "C" int32_t neddf_nerf_train_set_weights(neddf_nerf_train_t* h, const float* const* d_w, const float* const* d_b, int32_t n_layers,
                                                void* stream) {
  if (!h || !d_w || !d_b) return fail(NEDDF_E_INVALID, "neddf_nerf_train_set_weights: null argument");
  const int L = h->cfg.layer_count;
  if (n_layers != L + 3) return fail(NEDDF_E_INVALID, "neddf_nerf_train_set_weights: expected layer_count + 3 layers");
  cudaStream_t s = (cudaStream_t)stream;
  // hidden layers 0..L-1 are tensors 0..L-1; the colour branch's first layer (kernel layer L) is tensor L + 1
  for (int l = 0; l <= L; ++l) {
    const int t = (l < L) ? l : L + 1;
    const nerft::Layer& ly = h->proto.layer[l];
    nerft::nerf_train_pack_kernel<<<64, 256, 0, s>>>(d_w[t], d_b[t], h->shape_in[t], h->shape_out[t], ly.k_pad, ly.kt_pad,
                                                     h->d_w + ly.w_off, h->d_w + ly.b_off, h->d_w + ly.wt_off);
    NEDDF_LAUNCH_CHECK();
  }
  nerft::nerf_train_pack_heads_kernel<<<1, 256, 0, s>>>(d_w[L], d_b[L], d_w[L + 2], d_b[L + 2], h->d_w + h->proto.w_density_off,
                                                        h->d_w + h->proto.w_col2_off);
  NEDDF_LAUNCH_CHECK();
  h->packed = true;
  return NEDDF_OK;
}

static int32_t nerf_train_launch(const neddf_nerf_train_t* h, nerft::Params& P, const float* lowpass, const float* g_density,
                                 const float* g_color, float* X, float* G, float* E, float* D, float* C1, float* GC1, float* GZD,
                                 void* stream) {
  if (!h->packed) return fail(NEDDF_E_INVALID, "neddf_nerf_train_backward: weights were never set");
  if (!lowpass || !g_density || !g_color || !X || !G || !E || !D || !C1 || !GC1 || !GZD)
    return fail(NEDDF_E_INVALID, "neddf_nerf_train_backward: null argument");
  if (P.n <= 0) return NEDDF_OK;
  for (int e = 0; e < h->cfg.embed_pos_rank; ++e) P.lowpass[e] = lowpass[e];
  P.w = h->d_w;
  P.g_density = g_density; P.g_color = g_color;
  P.X = X; P.G = G; P.Eo = E; P.Do = D; P.C1 = C1; P.GC1 = GC1; P.GZD = GZD;
  const int64_t n_tiles = (P.n + nerft::kT - 1) / nerft::kT;
  const int grid = (int)std::min<int64_t>(n_tiles, sm_count());
  NEDDF_CUDA_CHECK(cudaFuncSetAttribute(nerft::nerf_train_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)nerft::kSmemBytes));
  nerft::nerf_train_kernel<<<grid, nerft::kThreads, nerft::kSmemBytes, (cudaStream_t)stream>>>(P);
  NEDDF_LAUNCH_CHECK();
  return NEDDF_OK;
}

extern "C" int32_t neddf_nerf_train_backward(const neddf_nerf_train_t* h, const float* lowpass, const float* d_pos, const float* d_dir,
                                             const float* d_var, int64_t n, const float* d_g_density, const float* d_g_color, float* d_x,
                                             float* d_g, float* d_e, float* d_d, float* d_c1, float* d_gc1, float* d_gzd, void* stream) {
  if (!h || !d_pos || !d_dir || !d_var) return fail(NEDDF_E_INVALID, "neddf_nerf_train_backward: null argument");
  nerft::Params P = h->proto;
  P.n = n;
  P.pos = d_pos; P.dir = d_dir; P.var = d_var;
  return nerf_train_launch(h, P, lowpass, d_g_density, d_g_color, d_x, d_g, d_e, d_d, d_c1, d_gc1, d_gzd, stream);
}

extern "C" int32_t neddf_nerf_train_backward_rays(const neddf_nerf_train_t* h, const float* lowpass, const float* d_ray_dir,
                                                  const float* d_ray_orig, const float* d_dists, int64_t n_rays, int32_t n_edges,
                                                  int32_t sampling_type, float ray_radius, const float* d_g_density,
                                                  const float* d_g_color, float* d_x, float* d_g, float* d_e, float* d_d, float* d_c1,
                                                  float* d_gc1, float* d_gzd, void* stream) {
  if (!h || !d_ray_dir || !d_ray_orig || !d_dists) return fail(NEDDF_E_INVALID, "neddf_nerf_train_backward_rays: null argument");
  if (n_edges < 1 || (sampling_type != NEDDF_SAMPLING_POINT && sampling_type != NEDDF_SAMPLING_CONE))
    return fail(NEDDF_E_INVALID, "neddf_nerf_train_backward_rays: bad n_edges / sampling_type");
  nerft::Params P = h->proto;
  P.n = n_rays * n_edges;
  P.ray_dir = d_ray_dir; P.ray_orig = d_ray_orig; P.dists = d_dists;
  P.n_edges = n_edges; P.sampling_type = sampling_type; P.ray_radius = ray_radius;
  return nerf_train_launch(h, P, lowpass, d_g_density, d_g_color, d_x, d_g, d_e, d_d, d_c1, d_gc1, d_gzd, stream);
}
