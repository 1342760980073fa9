// K3: alpha compositing along rays as a warp scan (one warp per ray).
//
// Reference: BaseNeuralRender.integrate_volume_render (neddf/render/base_neural_render.py:144-172)
// and the penalty integration of render_rays (neddf/render/nerf_render.py:153-159).
//
// HBM-bound: per sample it reads dist(4)+density(4)+color(12)[+penalty(4)] and writes weight(4):
// 24-28 algorithmic bytes.  Lane l handles samples l, l+32, ... so that every global access of
// a warp is a contiguous run; the exclusive transmittance product is a shuffle scan carried
// across the 32-sample blocks.  The product is accumulated in fp64 because torch's CPU cumprod
// accumulates float in double (at::acc_type<float,false>) and rounds each output once.
#include "common.cuh"

namespace neddf {

constexpr int kWarpsPerBlock = 8;

__global__ void __launch_bounds__(kWarpsPerBlock * 32)
composite_kernel(const float* __restrict__ dists, const float* __restrict__ density,
                 const float* __restrict__ color, const float* __restrict__ penalty, int64_t n_rays,
                 int n_edges, float max_dist, float* __restrict__ weight, float* __restrict__ depth,
                 float* __restrict__ color_out, float* __restrict__ transmittance,
                 float* __restrict__ penalty_out, int* __restrict__ status) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  const int64_t n_warps = (int64_t)gridDim.x * kWarpsPerBlock;
  const int n_int = n_edges - 1;  // intervals; the last edge only closes the last interval
  bool saw_nan = false;

  for (int64_t ray = warp; ray < n_rays; ray += n_warps) {
    const float* drow = dists + ray * n_edges;
    const float* srow = density + ray * n_edges;
    const float* crow = color + ray * n_edges * 3;
    const float* prow = penalty ? penalty + ray * n_edges : nullptr;
    float* wrow = weight ? weight + ray * n_int : nullptr;

    double carry = 1.0;  // T at the start of the current 32-sample block
    float acc_d = 0.f, acc_r = 0.f, acc_g = 0.f, acc_b = 0.f, acc_p = 0.f;
    for (int base = 0; base < n_int; base += 32) {
      const int j = base + lane;
      const bool live = j < n_int;
      float dj = 0.f, delta = 0.f, o = 0.f;
      if (live) {
        dj = drow[j];
        delta = drow[j + 1] - dj;
        // exp evaluated in fp64 and rounded once: CUDA's fp32 expf is up to 2 ulp with a small
        // one-sided bias just below 1, which accumulates over ~200 factors of the
        // transmittance product (measured 2.5e-6 vs torch's 1-ulp exp); the count is tiny.
        o = 1.0f - (float)exp((double)(-srow[j] * delta));
      }
      // factor_j = 1 - o + 1e-7 evaluated in fp32 like the reference's tensor expression
      double f = live ? (double)(1.0f - o + 1e-7f) : 1.0;
      // inclusive product scan over the block
      double incl = f;
#pragma unroll
      for (int s = 1; s < 32; s <<= 1) {
        double up = __shfl_up_sync(0xffffffffu, incl, s);
        if (lane >= s) incl *= up;
      }
      double excl = __shfl_up_sync(0xffffffffu, incl, 1);
      if (lane == 0) excl = 1.0;
      // torch rounds every cumprod output to fp32 before it is used
      float t_j = (float)(carry * excl);
      if (live) {
        float w = o * t_j;
        saw_nan |= (w != w);
        if (wrow) wrow[j] = w;
        acc_d += w * dj;
        acc_r += w * crow[3 * j + 0];
        acc_g += w * crow[3 * j + 1];
        acc_b += w * crow[3 * j + 2];
        if (prow) acc_p += delta * prow[j];
      }
      carry = carry * __shfl_sync(0xffffffffu, incl, 31);
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
      acc_d += __shfl_xor_sync(0xffffffffu, acc_d, s);
      acc_r += __shfl_xor_sync(0xffffffffu, acc_r, s);
      acc_g += __shfl_xor_sync(0xffffffffu, acc_g, s);
      acc_b += __shfl_xor_sync(0xffffffffu, acc_b, s);
      acc_p += __shfl_xor_sync(0xffffffffu, acc_p, s);
    }
    if (lane == 0) {
      float t_last = (float)carry;
      if (depth) depth[ray] = acc_d + t_last * max_dist;  // black background, :165
      if (color_out) {
        color_out[3 * ray + 0] = acc_r;
        color_out[3 * ray + 1] = acc_g;
        color_out[3 * ray + 2] = acc_b;
      }
      if (transmittance) transmittance[ray] = t_last;
      if (penalty_out) penalty_out[ray] = acc_p;
    }
  }
  if (status && __any_sync(0xffffffffu, saw_nan) && lane == 0) atomicOr(status, 1);
}

}  // namespace neddf

using namespace neddf;

extern "C" int32_t neddf_composite(const float* d_dists, const float* d_density, const float* d_color,
                                   const float* d_penalty, int64_t n_rays, int32_t n_edges, float max_dist,
                                   float* d_weight, float* d_depth, float* d_color_out, float* d_transmittance,
                                   float* d_penalty_out, int32_t* d_status, void* stream) {
  if (n_rays < 0 || n_edges < 1) return fail(NEDDF_E_INVALID, "neddf_composite: bad sizes");
  if (n_rays == 0) return NEDDF_OK;
  if (!d_dists || !d_density || !d_color) return fail(NEDDF_E_INVALID, "neddf_composite: null input pointer");
  if (d_penalty_out && !d_penalty) return fail(NEDDF_E_INVALID, "neddf_composite: penalty_out without penalty");
  int64_t blocks = (n_rays + kWarpsPerBlock - 1) / kWarpsPerBlock;
  int64_t cap = (int64_t)sm_count() * 8;  // 8 resident 256-thread CTAs per SM
  if (blocks > cap) blocks = cap;
  composite_kernel<<<(unsigned)blocks, kWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(
      d_dists, d_density, d_color, d_penalty_out ? d_penalty : nullptr, n_rays, n_edges, max_dist, d_weight,
      d_depth, d_color_out, d_transmittance, d_penalty_out, d_status);
  NEDDF_LAUNCH_CHECK();
  return NEDDF_OK;
}
