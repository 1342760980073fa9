// K4: hierarchical (inverse-CDF) resampling, one warp per ray.
//
// Reference: BaseNeuralRender.sample_pdf (neddf/render/base_neural_render.py:27-115), both modes:
// cat_coarse=True (what render_rays uses: new samples merged with the coarse edges) and cat_coarse=False
// (neighbour-max smoothing of the weights, :61-68, and only the new samples are returned).
//
// Per ray: sanitise + bias the coarse weights, L1-normalise, cumulative sum (fp64 accumulate,
// fp32 outputs -- torch's CPU cumsum accumulates float in double), binary search
// (searchsorted right=True) of every uniform in the cdf held in shared memory, linear
// interpolation, then a bitonic sort of [new | coarse] distances in shared memory.
// HBM traffic per ray: reads 4*(E + E-1 + F) B, writes 4*(E+F) B  (E edges, F new samples).
#include "common.cuh"

namespace neddf {

constexpr int kPdfWarps = 4;

__device__ __forceinline__ int upper_bound(const float* __restrict__ cdf, int n, float u) {
  // number of entries <= u  == torch.searchsorted(cdf, u, right=True)
  int lo = 0, hi = n;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (cdf[mid] <= u) lo = mid + 1;
    else hi = mid;
  }
  return lo;
}

__device__ __forceinline__ float invert_one(const float* __restrict__ cdf, const float* __restrict__ dists,
                                            int n_edges, float u, int& id) {
  id = upper_bound(cdf, n_edges, u);
  int below = max(0, id - 1);
  int above = min(n_edges - 1, id);
  float c0 = cdf[below], c1 = cdf[above];
  float d0 = dists[below], d1 = dists[above];
  float denom = NS(c1, c0);
  if (denom < 1e-5f) denom = 1.0f;
  float t = __fdiv_rn(NS(u, c0), denom);
  return NA(d0, NM(t, NS(d1, d0)));
}

// smem per warp: cdf[n_edges] | dists[n_edges] | merged[p2]
__global__ void __launch_bounds__(kPdfWarps * 32)
sample_pdf_kernel(const float* __restrict__ dists, float* __restrict__ weights, const float* __restrict__ u,
                  int64_t n_rays, int n_edges, int n_new, int p2, int cat_coarse, float* __restrict__ dists_fine,
                  int64_t* __restrict__ ids_out, float* __restrict__ cdf_out, int* __restrict__ status) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  float* s_cdf = smem + (size_t)wib * (2 * n_edges + p2);
  float* s_dist = s_cdf + n_edges;
  float* s_merge = s_dist + n_edges;
  const int n_w = n_edges - 1;
  const int n_out = cat_coarse ? n_edges + n_new : n_new;
  bool saw_nan = false;

  const int64_t n_warps = (int64_t)gridDim.x * kPdfWarps;
  for (int64_t ray = (int64_t)blockIdx.x * kPdfWarps + wib; ray < n_rays; ray += n_warps) {
    float* wrow = weights + ray * n_w;
    const float* drow = dists + ray * n_edges;
    // ---- sanitise (:52-55, in place), +1e-2 (:58), L1 norm (:70)
    double part = 0.0;
    for (int j = lane; j < n_w; j += 32) {
      float w = wrow[j];
      bool touched = false;
      if (w < 0.0f) { w = w * 0.0f; touched = true; }
      if (w != w) { w = 0.0f; touched = true; }
      if (touched) wrow[j] = w;
      w = w + 1e-2f;
      s_merge[j] = w;  // staging for the scan
      if (cat_coarse) part += (double)fabsf(w);
    }
    if (!cat_coarse) {  // :61-68: interior weights <- 0.5 (max(w[j+1], w[j]) + max(w[j-1], w[j])), from the biased weights
      __syncwarp();
      for (int j = lane; j < n_w; j += 32) {
        float w = s_merge[j];
        if (j >= 1 && j + 1 < n_w) w = NM(0.5f, NA(fmaxf(s_merge[j + 1], w), fmaxf(s_merge[j - 1], w)));
        s_cdf[j] = w;
      }
      __syncwarp();
      for (int j = lane; j < n_w; j += 32) {
        s_merge[j] = s_cdf[j];
        part += (double)fabsf(s_cdf[j]);
      }
    }
    for (int j = lane; j < n_edges; j += 32) s_dist[j] = drow[j];
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) part += __shfl_xor_sync(0xffffffffu, part, s);
    float denom = fmaxf((float)part, 1e-12f);  // F.normalize eps
    __syncwarp();
    // ---- cdf = [0, cumsum(pdf)] (:72-73)
    double carry = 0.0;
    if (lane == 0) s_cdf[0] = 0.0f;
    for (int base = 0; base < n_w; base += 32) {
      int j = base + lane;
      double v = (j < n_w) ? (double)(s_merge[j] / denom) : 0.0;
#pragma unroll
      for (int s = 1; s < 32; s <<= 1) {
        double up = __shfl_up_sync(0xffffffffu, v, s);
        if (lane >= s) v += up;
      }
      if (j < n_w) s_cdf[j + 1] = (float)(carry + v);
      carry += __shfl_sync(0xffffffffu, v, 31);
    }
    __syncwarp();
    if (cdf_out)
      for (int j = lane; j < n_edges; j += 32) cdf_out[ray * n_edges + j] = s_cdf[j];
    // ---- inverse CDF (:77-98)
    const float* urow = u + ray * n_new;
    for (int i = lane; i < n_new; i += 32) {
      int id;
      float smp = invert_one(s_cdf, s_dist, n_edges, urow[i], id);
      if (ids_out) ids_out[ray * n_new + i] = id;
      s_merge[i] = smp;
      saw_nan |= (smp != smp);
    }
    // ---- cat([samples, dists]) (:102) or the samples alone (:104), pad to a power of two with +inf
    if (cat_coarse) {
      for (int j = lane; j < n_edges; j += 32) {
        float d = s_dist[j];
        s_merge[n_new + j] = d;
        saw_nan |= (d != d);
      }
    }
    for (int j = n_out + lane; j < p2; j += 32) s_merge[j] = __int_as_float(0x7f800000);
    __syncwarp();
    // ---- bitonic sort ascending
    for (int k = 2; k <= p2; k <<= 1) {
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int t = lane; t < (p2 >> 1); t += 32) {
          int lo = ((t & ~(j - 1)) << 1) | (t & (j - 1));
          int hi = lo | j;
          bool up = ((lo & k) == 0);
          float a = s_merge[lo], b = s_merge[hi];
          if ((a > b) == up) {
            s_merge[lo] = b;
            s_merge[hi] = a;
          }
        }
        __syncwarp();
      }
    }
    for (int j = lane; j < n_out; j += 32) dists_fine[ray * n_out + j] = s_merge[j];
    __syncwarp();
  }
  if (status && __any_sync(0xffffffffu, saw_nan) && lane == 0) atomicOr(status + 1, 2);  // this launch's flag
}

// Batch-wide fallback of the reference (:105-114): if any merged sample of THIS launch was NaN, every
// ray of the batch gets linspace(dists[0,0], dists[0,-1], n_out).  Runs on device so the host never
// synchronises.  status[1] is the per-launch flag (zeroed before sample_pdf_kernel); it is folded into
// the persistent word status[0] that the host reads and clears (the reference prints each time).
__global__ void pdf_nan_fallback_kernel(int* __restrict__ status, const float* __restrict__ dists,
                                        int n_edges, int64_t total, int n_out, float* __restrict__ dists_fine) {
  if ((status[1] & 2) == 0) return;
  if (blockIdx.x == 0 && threadIdx.x == 0) atomicOr(status, 2);
  float a = dists[0], b = dists[n_edges - 1];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
    dists_fine[i] = linspace_at(a, b, n_out, (int)(i % n_out));
}

__global__ void invert_cdf_kernel(const float* __restrict__ dists, const float* __restrict__ cdf,
                                  const float* __restrict__ u, int64_t total, int n_edges, int n_new,
                                  float* __restrict__ samples, int64_t* __restrict__ ids) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int64_t ray = i / n_new;
  int id;
  float s = invert_one(cdf + ray * n_edges, dists + ray * n_edges, n_edges, u[i], id);
  if (samples) samples[i] = s;
  if (ids) ids[i] = id;
}

}  // namespace neddf

using namespace neddf;

extern "C" int32_t neddf_sample_pdf(const float* d_dists, float* d_weights, const float* d_u, int64_t n_rays,
                                    int32_t n_edges, int32_t n_new, int32_t cat_coarse, float* d_dists_fine,
                                    int64_t* d_ids, float* d_cdf, int32_t* d_status, void* stream) {
  if (n_rays < 0 || n_edges < 2 || n_new < 0) return fail(NEDDF_E_INVALID, "neddf_sample_pdf: bad sizes");
  if (n_rays == 0) return NEDDF_OK;
  if (!d_dists || !d_weights || (!d_u && n_new > 0) || !d_dists_fine)
    return fail(NEDDF_E_INVALID, "neddf_sample_pdf: null device pointer");
  int n_out = cat_coarse ? n_edges + n_new : n_new;
  if (n_out < 1) return fail(NEDDF_E_INVALID, "neddf_sample_pdf: nothing to produce");
  int p2 = 2;
  while (p2 < n_out) p2 <<= 1;
  if (p2 < n_edges) p2 = n_edges;  // the merge buffer also stages the weights
  { int q = 2; while (q < p2) q <<= 1; p2 = q; }
  size_t smem = (size_t)kPdfWarps * (2 * n_edges + p2) * sizeof(float);
  if (smem > 200 * 1024) return fail(NEDDF_E_UNSUPPORTED, "neddf_sample_pdf: too many samples per ray for shared memory");
  cudaStream_t s = (cudaStream_t)stream;
  if (smem > 48 * 1024)
    NEDDF_CUDA_CHECK(cudaFuncSetAttribute(sample_pdf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int64_t blocks = (n_rays + kPdfWarps - 1) / kPdfWarps;
  int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  if (d_status) NEDDF_CUDA_CHECK(cudaMemsetAsync(d_status + 1, 0, sizeof(int32_t), s));
  sample_pdf_kernel<<<(unsigned)blocks, kPdfWarps * 32, smem, s>>>(d_dists, d_weights, d_u, n_rays, n_edges, n_new,
                                                                   p2, cat_coarse ? 1 : 0, d_dists_fine, d_ids, d_cdf, d_status);
  NEDDF_LAUNCH_CHECK();
  if (d_status) {
    pdf_nan_fallback_kernel<<<sm_count(), 256, 0, s>>>(d_status, d_dists, n_edges, n_rays * (int64_t)n_out, n_out,
                                                       d_dists_fine);
    NEDDF_LAUNCH_CHECK();
  }
  return NEDDF_OK;
}

extern "C" int32_t neddf_invert_cdf(const float* d_dists, const float* d_cdf, const float* d_u, int64_t n_rays,
                                    int32_t n_edges, int32_t n_new, float* d_samples, int64_t* d_ids, void* stream) {
  if (n_rays < 0 || n_edges < 1 || n_new < 0) return fail(NEDDF_E_INVALID, "neddf_invert_cdf: bad sizes");
  int64_t total = n_rays * (int64_t)n_new;
  if (total == 0) return NEDDF_OK;
  if (!d_dists || !d_cdf || !d_u) return fail(NEDDF_E_INVALID, "neddf_invert_cdf: null device pointer");
  int threads = 256;
  int64_t blocks = (total + threads - 1) / threads;
  invert_cdf_kernel<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(d_dists, d_cdf, d_u, total, n_edges,
                                                                            n_new, d_samples, d_ids);
  NEDDF_LAUNCH_CHECK();
  return NEDDF_OK;
}
