// lip_eval.cu — the consumer of predict_lla_scalable in the reference's evaluation loop (scale_experiments/evaluate.py:98-146,
// SURVEY §8 row f4): Monte-Carlo softmax predictive from S logit samples.
//   log_avg_prob[b] = logsumexp_s( log_softmax(logits[s, b, :])[y_b] ) - log S          (evaluate.py:126-137)
//   mean_probs[b,c] = (1/S) sum_s softmax(logits[s, b, :])[c]                            (evaluate.py:141-142)
// One thread per example: consecutive threads read consecutive rows of the [S, B, C] block (coalesced), the C-wide row lives in
// registers, the logsumexp over samples is the streaming (running max, rescaled sum) form.  HBM-bound: 4·S·B·C bytes read once.
#include <math.h>

#include "lip_common.cuh"

namespace lip {
namespace {

constexpr int EVAL_CMAX = 64;

__global__ void mc_softmax_kernel(const float* __restrict__ logits, const int* __restrict__ labels, float* __restrict__ log_avg,
                                  float* __restrict__ mean_probs, int64_t S, int64_t B, int C) {
  const int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (b >= B) return;
  float acc[EVAL_CMAX];
#pragma unroll
  for (int c = 0; c < EVAL_CMAX; ++c) acc[c] = 0.f;
  const int y = labels ? labels[b] : -1;
  float run_max = -INFINITY, run_sum = 0.f;
  for (int64_t s = 0; s < S; ++s) {
    const float* row = logits + (s * B + b) * C;
    float row_v[EVAL_CMAX];
    float mx = -INFINITY;
#pragma unroll
    for (int c = 0; c < EVAL_CMAX; ++c)
      if (c < C) { row_v[c] = row[c]; mx = fmaxf(mx, row_v[c]); }
    float z = 0.f;
#pragma unroll
    for (int c = 0; c < EVAL_CMAX; ++c)
      if (c < C) { row_v[c] = expf(row_v[c] - mx); z += row_v[c]; }
    const float inv = 1.f / z;
#pragma unroll
    for (int c = 0; c < EVAL_CMAX; ++c)
      if (c < C) acc[c] += row_v[c] * inv;
    if (y >= 0 && y < C) {
      const float lp = (row[y] - mx) - logf(z);                 // log_softmax at the true class
      if (lp > run_max) { run_sum = run_sum * expf(run_max - lp) + 1.f; run_max = lp; }
      else run_sum += expf(lp - run_max);
    }
  }
  const float invS = 1.f / (float)S;
  if (mean_probs)
    for (int c = 0; c < C; ++c) mean_probs[b * C + c] = acc[c] * invS;
  if (log_avg) log_avg[b] = (y >= 0 && y < C) ? run_max + logf(run_sum) - logf((float)S) : NAN;
}

}  // namespace
}  // namespace lip

using namespace lip;

extern "C" int lip_mc_softmax_predictive(const float* logits, const int32_t* labels, float* log_avg_prob, float* mean_probs, int64_t S,
                                         int64_t B, int32_t C, lip_stream_t stream) {
  LIP_REQUIRE(logits && S > 0 && B > 0 && C > 0, "lip_mc_softmax_predictive: null logits or empty shape");
  LIP_REQUIRE(C <= EVAL_CMAX, "lip_mc_softmax_predictive: at most %d classes (got %d)", EVAL_CMAX, C);
  LIP_REQUIRE(log_avg_prob == nullptr || labels != nullptr, "lip_mc_softmax_predictive: log_avg_prob needs labels");
  mc_softmax_kernel<<<(unsigned)ceil_div(B, 128), 128, 0, (cudaStream_t)stream>>>(logits, labels, log_avg_prob, mean_probs, S, B, C);
  LIP_LAUNCH_CHECK();
  return LIP_OK;
}
