// Per-sample head math shared by the fp32 and tensor-core field kernels:
// softplus/sigmoid heads, distance -> density, surface normal, field-constraint penalties.
#pragma once

#include "field.cuh"

namespace neddf {

struct HeadOut {
  float ddf_out, aux_out;  // pre-activation head outputs (needed by the range penalties)
  float distance, density, aux, dist_inv, grad_norm, dDdt;
  float grad_d[3];  // nabla distance
  float aux_gg[3];  // nabla aux
  float normal[3];
};

// neddf/network/neddf.py:220-241.  ddf[4] / aux[4] = (value, d/dx, d/dy, d/dz) of the two
// 256->1 linear heads, bias already added to the value.
__device__ __forceinline__ void head_density(const float ddf[4], const float aux[4], float d_near,
                                             float aux_grad_scale, int density_act_id, HeadOut& h) {
  h.ddf_out = ddf[0];
  h.aux_out = aux[0];
  // SoftplusGradFunction, nn_module/with_grad/softplus.py:38-48
  float sp, sp_d1;
  if (ddf[0] > 20.0f) {
    sp = ddf[0];
    sp_d1 = 1.0f;
  } else {
    sp = logf(1.0f + expf(ddf[0]));
    sp_d1 = 1.0f / (1.0f + expf(-ddf[0]));
  }
  h.distance = sp + d_near;
  // SigmoidGradFunction, nn_module/with_grad/sigmoid.py:38-43
  float t = (1.0f + tanhf(aux[0] * 0.5f)) * 0.5f;
  float sg_d1 = t * (1.0f - t);
  h.aux = aux_grad_scale * t;
  float n2 = 0.f;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    h.grad_d[i] = sp_d1 * ddf[1 + i];
    h.aux_gg[i] = aux_grad_scale * (sg_d1 * aux[1 + i]);
    n2 += h.grad_d[i] * h.grad_d[i];
  }
  h.grad_norm = sqrtf(n2);
  h.dDdt = sqrtf(n2 + h.aux * h.aux);  // norm of [grad_d, aux], neddf.py:232-236
  h.dist_inv = 1.0f / h.distance;
  h.density = density_act(density_act_id, h.dist_inv * (1.0f - h.dDdt));
  float inv = 1.0f / (h.grad_norm + 1e-7f);
#pragma unroll
  for (int i = 0; i < 3; ++i) h.normal[i] = inv * h.grad_d[i];
}

// neddf/network/neddf.py:259-300.  col[c] = colour value, colJ[i][c] = d colour_c / d pos_i.
// pw[] in the reference's insertion order (see neddf_field_config_t.penalty_weight).
__device__ __forceinline__ float field_penalty(const HeadOut& h, const float col[3], const float colJ[3][3],
                                               float distance_range_max, const float* pw) {
  float d2 = h.aux_gg[0] * h.normal[0] + h.aux_gg[1] * h.normal[1] + h.aux_gg[2] * h.normal[2];
  float rest = 3.0f * h.aux * h.dist_inv;
  float ag_scale = h.aux * h.grad_norm * h.distance;
  float diff = d2 - rest;
  float p_aux = ag_scale * (diff * diff);
  float r = fmaxf(-1.0f + h.dDdt, 0.0f);
  float p_dDdt = r * r;
  float rd = fmaxf(-4.6f - h.ddf_out, 0.0f) + fmaxf(-distance_range_max + h.ddf_out, 0.0f);
  float p_rd = rd * rd;
  float ra = fmaxf(-4.6f - h.aux_out, 0.0f) + fmaxf(-4.6f + h.aux_out, 0.0f);
  float p_ra = ra * ra;
  float p_rc = 0.f, p_cc = 0.f;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float rc = fmaxf(-0.0f - col[c], 0.0f) + fmaxf(-1.0f + col[c], 0.0f);
    p_rc += rc * rc;
    float dot = colJ[0][c] * h.grad_d[0] + colJ[1][c] * h.grad_d[1] + colJ[2][c] * h.grad_d[2];
    p_cc += dot * dot;
  }
  return ((((p_aux * pw[0] + p_dDdt * pw[1]) + p_rd * pw[2]) + p_ra * pw[3]) + p_rc * pw[4]) + p_cc * pw[5];
}

// Position embedding entry (e, d): sin/cos of 2^e * x_d with the cone weight
// exp(-0.5 * 4^e * var_d) (ray/sampling.py:58-71) and the two scalings of neddf.py:193-209.
struct PeEntry {
  float s, c;        // sin p, cos p
  float scale_s;     // scaled embedding (distance trunk): (2/2^e * lowpass_e) * w
  float scale_0;     // plain embedding (colour trunk):      lowpass_e * w
  float freq;        // 2^e
};

__device__ __forceinline__ PeEntry pe_entry(int e, float x, float var, float lowpass_e) {
  PeEntry r;
  r.freq = (float)(1u << e);
  float p = r.freq * x;
  sincosf(p, &r.s, &r.c);
  float w = expf(-0.5f * (r.freq * r.freq) * var);
  float s_grad = 1.0f / (0.5f * r.freq);
  r.scale_s = (s_grad * lowpass_e) * w;
  r.scale_0 = lowpass_e * w;
  return r;
}

}  // namespace neddf
