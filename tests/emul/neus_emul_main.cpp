// Sanitizer driver for the emulated NeuS kernel (TEST INFRASTRUCTURE): the tile program of csrc/neus_kernel.cuh on
// 256 OS threads per CTA, built with -fsanitize=address,undefined (out-of-bounds / misaligned vector accesses to the
// emulated shared memory, the packed weights and the I/O arrays: what compute-sanitizer memcheck looks for) or with
// -fsanitize=thread (unsynchronised conflicting accesses between the threads of a CTA: racecheck; the pthread barrier
// that stands in for __syncthreads is the only synchronisation ThreadSanitizer sees).  Exit code 0 = clean.
#include "neus_emul.cpp"

#include <stdio.h>

#include <random>

static int run(const neddf_neus_config_t& cfg, int n, bool rays, unsigned seed) {
  int sin[neus::kMaxSdf + neus::kMaxCol + 2], sout[neus::kMaxSdf + neus::kMaxCol + 2];
  const int nl = neus::layer_shapes(&cfg, sin, sout);
  std::mt19937 g(seed);
  std::normal_distribution<float> nd(0.f, 1.f);
  std::vector<std::vector<float>> W(nl), B(nl);
  std::vector<const float*> wp(nl), bp(nl);
  for (int t = 0; t < nl; ++t) {
    W[t].resize((size_t)sin[t] * sout[t]);
    B[t].resize(sout[t]);
    const float s = sqrtf(2.f / (sin[t] + sout[t]));
    for (auto& v : W[t]) v = s * nd(g);
    for (auto& v : B[t]) v = 0.05f * nd(g);
    wp[t] = W[t].data();
    bp[t] = B[t].data();
  }
  const float variance = 0.3f;
  const int n_edges = rays ? 7 : 0;
  const long long total = rays ? (long long)n * n_edges : n;
  std::vector<float> pos(3 * total), dir(3 * total), rd(3 * n), ro(3 * n), dists((size_t)n * (rays ? n_edges : 1));
  for (auto& v : pos) v = 0.8f * nd(g);
  for (long long i = 0; i < total; ++i) {
    float a = nd(g), b = nd(g), c = nd(g), r = sqrtf(a * a + b * b + c * c) + 1e-6f;
    dir[3 * i] = a / r; dir[3 * i + 1] = b / r; dir[3 * i + 2] = c / r;
  }
  for (int i = 0; i < n; ++i) {
    float a = nd(g), b = nd(g), c = nd(g), r = sqrtf(a * a + b * b + c * c) + 1e-6f;
    rd[3 * i] = a / r; rd[3 * i + 1] = b / r; rd[3 * i + 2] = c / r;
    ro[3 * i] = 0.1f * nd(g); ro[3 * i + 1] = 0.1f * nd(g); ro[3 * i + 2] = 0.1f * nd(g);
    for (int j = 0; j < n_edges; ++j) dists[(size_t)i * n_edges + j] = 2.f + 0.5f * j + 0.1f * fabsf(nd(g));
  }
  // exact-size outputs: a write past sample n - 1 is a heap overflow
  std::vector<float> sdf(total), den(total), col(3 * total), nrm(3 * total);
  const int rc = neus_emul_forward(&cfg, wp.data(), bp.data(), nl, &variance, rays ? nullptr : pos.data(), rays ? nullptr : dir.data(),
                                   rays ? rd.data() : nullptr, rays ? ro.data() : nullptr, rays ? dists.data() : nullptr, n, n_edges,
                                   NEDDF_SAMPLING_CONE, 2.6e-4f, sdf.data(), den.data(), col.data(), nrm.data(), 2);
  double sum = 0;
  for (long long i = 0; i < total; ++i) sum += sdf[i] + den[i] + col[3 * i] + nrm[3 * i + 2];
  printf("rc %d checksum %.6f (%lld samples)\n", rc, sum, total);
  return (rc == 0 && sum == sum) ? 0 : 1;
}

int main() {
  neddf_neus_config_t a = {6, 4, 8, 256, 8, 256, NEDDF_ACT_RELU, 1, {4}};      // config/network/neus.yaml
  neddf_neus_config_t b = {10, 4, 3, 256, 2, 256, NEDDF_ACT_TANHEXP, 2, {0, 1}};  // largest embeddings, skips after layers 0 and 1
  int bad = 0;
  bad |= run(a, 70, false, 1);  // one full tile + a ragged one (last SDF sub-tile partly empty), two CTAs
  bad |= run(b, 19, true, 2);   // fused ray geometry, 133 samples: three tiles over two CTAs (CTA 0 runs two of them)
  return bad;
}
