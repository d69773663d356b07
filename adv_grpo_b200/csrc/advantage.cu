// Group-relative advantage (per-prompt mean, global or per-prompt population std) and
// the GRPO clipped loss.  Both are tiny, latency-bound problems ([N<=~10^4, T<=8] and
// [B<=64]); they exist to keep the rollout->update loop free of the reference's
// device->host->numpy->device round trips.  float64 arithmetic like the reference.
//
// Reference: adv_grpo/stat_tracking.py:18-47, scripts/train_sd3_fast_pickscore.py:195-229,
//            :962-970, :1111-1162.
#include "common.cuh"

namespace advgrpo {
namespace {

__device__ __forceinline__ uint64_t mix64(uint64_t z) {  // splitmix64 finaliser
  z ^= z >> 30; z *= 0xbf58476d1ce4e5b9ull;
  z ^= z >> 27; z *= 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}

// One warp per row: two independent 64-bit digests of the key row (128-bit identity).
__global__ void hash_rows_kernel(const int64_t* keys, int64_t key_len, int64_t N, uint64_t* h1,
                                 uint64_t* h2) {
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= N) return;
  const int lane = threadIdx.x & 31;
  uint64_t a = 0, b = 0;
  for (int64_t i = lane; i < key_len; i += 32) {
    uint64_t v = (uint64_t)keys[row * key_len + i];
    a += mix64(v + 0x9e3779b97f4a7c15ull * (uint64_t)(i + 1));
    b ^= mix64((v ^ 0xc2b2ae3d27d4eb4full) + 0x165667b19e3779f9ull * (uint64_t)(i + 1));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b ^= __shfl_xor_sync(0xffffffffu, b, o);
  }
  if (lane == 0) { h1[row] = a; h2[row] = b; }
}

__device__ double block_sum_d(double v, double* scratch) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();
  if (l == 0) scratch[w] = v;
  __syncthreads();
  double t = (l < nw) ? scratch[l] : 0.0;
  t = warp_sum(t);
  return t;
}

// Single CTA. O(N^2) group scan out of L1/L2 -- N is the number of images of one epoch.
__global__ void __launch_bounds__(1024) group_advantage_kernel(const float* __restrict__ r,
                                                               const uint64_t* __restrict__ h1,
                                                               const uint64_t* __restrict__ h2,
                                                               int64_t N, int64_t T, int global_std,
                                                               int mode, double* __restrict__ adv,
                                                               double* __restrict__ stats) {
  __shared__ double scratch[32];
  __shared__ double col_std[64];
  // column statistics (np.std(rewards, axis=0): population std, two-pass)
  for (int64_t t = 0; t < T; ++t) {
    double s = 0.0;
    for (int64_t i = threadIdx.x; i < N; i += blockDim.x) s += (double)r[i * T + t];
    const double mean = block_sum_d(s, scratch) / (double)N;
    double q = 0.0;
    for (int64_t i = threadIdx.x; i < N; i += blockDim.x) {
      double d = (double)r[i * T + t] - mean;
      q += d * d;
    }
    const double var = block_sum_d(q, scratch) / (double)N;
    if (threadIdx.x == 0) col_std[t] = sqrt(var);
    __syncthreads();
  }
  double n_groups = 0.0, zero_std = 0.0, std_sum = 0.0;
  for (int64_t i = threadIdx.x; i < N; i += blockDim.x) {
    const uint64_t a = h1[i], b = h2[i];
    bool leader = true;
    if (mode != 0) {
      // the other `type`s of PerPromptStatTracker.update (stat_tracking.py:48-70); group members in array order
      for (int64_t j = 0; j < i; ++j)
        if (h1[j] == a && h2[j] == b) leader = false;
      if (mode == 1) {                                   // 'rwr': the rewards themselves
        for (int64_t t = 0; t < T; ++t) adv[i * T + t] = (double)r[i * T + t];
      } else if (mode == 2) {                            // 'sft': 1 where the reward equals the maximum of the group's
        float mx = -INFINITY;                            //        WHOLE [n, T] block (torch.max over all elements)
        for (int64_t j = 0; j < N; ++j)
          if (h1[j] == a && h2[j] == b)
            for (int64_t t = 0; t < T; ++t) mx = fmaxf(mx, r[j * T + t]);
        for (int64_t t = 0; t < T; ++t) adv[i * T + t] = r[i * T + t] == mx ? 1.0 : 0.0;
      } else {                                           // 'dpo' (T == 1): +1 at the first arg-max, -1 at the first arg-min
        float mx = -INFINITY, mn = INFINITY;
        int64_t jmax = -1, jmin = -1, first = -1, second = -1;
        for (int64_t j = 0; j < N; ++j)
          if (h1[j] == a && h2[j] == b) {
            const float v = r[j];
            if (v > mx) { mx = v; jmax = j; }
            if (v < mn) { mn = v; jmin = j; }
            if (first < 0) first = j; else if (second < 0) second = j;
          }
        if (jmax == jmin) { jmin = first; jmax = second; }   // all equal: min_idx = 0, max_idx = 1 (stat_tracking.py:60-62)
        adv[i] = i == jmax ? 1.0 : (i == jmin ? -1.0 : 0.0);
      }
    }
    for (int64_t t = 0; mode == 0 && t < T; ++t) {
      double s = 0.0;
      int64_t cnt = 0;
      for (int64_t j = 0; j < N; ++j) {
        if (h1[j] == a && h2[j] == b) {
          s += (double)r[j * T + t];
          ++cnt;
          if (t == 0 && j < i) leader = false;
        }
      }
      const double mean = s / (double)cnt;
      double sd;
      if (global_std) {
        sd = col_std[t];
      } else {
        double q = 0.0;
        for (int64_t j = 0; j < N; ++j)
          if (h1[j] == a && h2[j] == b) {
            double d = (double)r[j * T + t] - mean;
            q += d * d;
          }
        sd = sqrt(q / (double)cnt);
      }
      adv[i * T + t] = ((double)r[i * T + t] - mean) / (sd + 1e-4);   // stat_tracking.py:41-47
    }
    if (leader && stats) {   // calculate_zero_std_ratio on column 0 ('ori_avg')
      double s = 0.0; int64_t cnt = 0; bool all_equal = true;
      const float r0 = r[i * T];
      for (int64_t j = 0; j < N; ++j)
        if (h1[j] == a && h2[j] == b) { s += (double)r[j * T]; ++cnt; all_equal &= (r[j * T] == r0); }
      const double mean = s / (double)cnt;
      double q = 0.0;
      for (int64_t j = 0; j < N; ++j)
        if (h1[j] == a && h2[j] == b) { double d = (double)r[j * T] - mean; q += d * d; }
      n_groups += 1.0;
      std_sum += all_equal ? 0.0 : sqrt(q / (double)cnt);
      zero_std += all_equal ? 1.0 : 0.0;
    }
  }
  if (stats) {
    n_groups = block_sum_d(n_groups, scratch);
    zero_std = block_sum_d(zero_std, scratch);
    std_sum = block_sum_d(std_sum, scratch);
    if (threadIdx.x == 0) {
      stats[0] = n_groups;
      stats[1] = (double)N / n_groups;
      stats[2] = zero_std / n_groups;
      stats[3] = std_sum / n_groups;
    }
  }
}

// type == 'grpo' (the only one the two scripts use), N <= kFastMaxN: the O(N^2) part -- who is in my group? -- runs ONCE per
// row out of shared memory (first row with the same 128-bit digest = the group's leader), the T column statistics are
// reduced by one warp each instead of 2 T block-wide reductions, and the group mean / std are computed once per
// (group, column) by the leader's owner thread (members visited in ascending row order: the same summation order, hence
// the same doubles, as the general kernel above) instead of once per (row, column).  Single CTA: the data is N T floats.
constexpr int kFastMaxN = 8192;
__global__ void __launch_bounds__(1024) group_advantage_grpo_kernel(const float* __restrict__ r,
                                                                    const uint64_t* __restrict__ h1,
                                                                    const uint64_t* __restrict__ h2, int N, int T,
                                                                    int global_std, double* __restrict__ adv,
                                                                    double* __restrict__ stats, double* __restrict__ gm,
                                                                    double* __restrict__ gs) {
  extern __shared__ __align__(16) uint8_t adv_smem[];
  uint64_t* sh1 = reinterpret_cast<uint64_t*>(adv_smem);
  uint64_t* sh2 = sh1 + N;
  int* lead = reinterpret_cast<int*>(sh2 + N);
  __shared__ double scratch[32];
  __shared__ double col_std[64];
  const int tid = threadIdx.x, nth = blockDim.x, warp = tid >> 5, lane = tid & 31, nw = nth >> 5;
  for (int i = tid; i < N; i += nth) { sh1[i] = h1[i]; sh2[i] = h2[i]; }
  __syncthreads();
  for (int i = tid; i < N; i += nth) {
    const uint64_t a = sh1[i], b = sh2[i];
    int j = 0;
    while (!(sh1[j] == a && sh2[j] == b)) ++j;          // terminates at j == i at the latest
    lead[i] = j;
  }
  // np.std(rewards, axis=0): population std, two-pass, one warp per column
  for (int t = warp; t < T; t += nw) {
    double s = 0.0;
    for (int i = lane; i < N; i += 32) s += (double)r[(int64_t)i * T + t];
    const double mean = warp_sum(s) / (double)N;
    double q = 0.0;
    for (int i = lane; i < N; i += 32) {
      const double d = (double)r[(int64_t)i * T + t] - mean;
      q += d * d;
    }
    q = warp_sum(q);
    if (lane == 0) col_std[t] = sqrt(q / (double)N);
  }
  __syncthreads();
  double n_groups = 0.0, zero_std = 0.0, std_sum = 0.0;
  for (int p = tid; p < N * T; p += nth) {
    const int i = p / T, t = p - i * T;
    if (lead[i] != i) continue;
    double s = 0.0;
    int cnt = 0;
    bool all_equal = true;
    const float r0 = r[(int64_t)i * T + t];
    for (int j = i; j < N; ++j)
      if (lead[j] == i) {
        const float v = r[(int64_t)j * T + t];
        s += (double)v;
        ++cnt;
        all_equal &= (v == r0);
      }
    const double mean = s / (double)cnt;
    double sd = col_std[t];
    if (!global_std || (t == 0 && stats)) {
      double q = 0.0;
      for (int j = i; j < N; ++j)
        if (lead[j] == i) {
          const double d = (double)r[(int64_t)j * T + t] - mean;
          q += d * d;
        }
      const double gsd = sqrt(q / (double)cnt);
      if (!global_std) sd = gsd;
      if (t == 0) {                                      // calculate_zero_std_ratio on column 0 ('ori_avg')
        n_groups += 1.0;
        std_sum += all_equal ? 0.0 : gsd;
        zero_std += all_equal ? 1.0 : 0.0;
      }
    }
    gm[p] = mean;
    gs[p] = sd;
  }
  __syncthreads();
  for (int p = tid; p < N * T; p += nth) {
    const int i = p / T, t = p - i * T;
    const int l = lead[i] * T + t;
    adv[p] = ((double)r[p] - gm[l]) / (gs[l] + 1e-4);     // stat_tracking.py:41-47
  }
  if (stats) {
    n_groups = block_sum_d(n_groups, scratch);
    zero_std = block_sum_d(zero_std, scratch);
    std_sum = block_sum_d(std_sum, scratch);
    if (tid == 0) {
      stats[0] = n_groups;
      stats[1] = (double)N / n_groups;
      stats[2] = zero_std / n_groups;
      stats[3] = std_sum / n_groups;
    }
  }
}

// One warp; B is a per-rank micro-batch (8..64).
__global__ void grpo_clip_loss_kernel(const float* __restrict__ lp, const float* __restrict__ lp_old,
                                      const double* __restrict__ adv, int64_t adv_stride, int64_t B,
                                      double clip, double adv_clip, double grad_scale,
                                      double* __restrict__ out, float* __restrict__ grad_lp) {
  double loss = 0, kl = 0, cf = 0, cf_gt = 0, cf_lt = 0;
  for (int64_t i = threadIdx.x; i < B; i += 32) {
    double A = adv[i * adv_stride];
    A = A < -adv_clip ? -adv_clip : (A > adv_clip ? adv_clip : A);        // :1111-1115
    const float diff = lp[i] - lp_old[i];
    const float ratio_f = expf(diff);                                     // :1116 (fp32 tensor op)
    const double ratio = (double)ratio_f;
    const double lo = 1.0 - clip, hi = 1.0 + clip;
    // torch.clamp(ratio_f32, lo, hi): python scalars are cast to the tensor dtype (fp32)
    const float lo_f = (float)lo, hi_f = (float)hi;
    const float rc_f = ratio_f < lo_f ? lo_f : (ratio_f > hi_f ? hi_f : ratio_f);
    const double unclipped = -A * ratio;                                  // :1117
    const double clipped = -A * (double)rc_f;                             // :1118-1122
    loss += unclipped > clipped ? unclipped : clipped;                    // :1123
    kl += (double)(diff * diff);
    const float dev = fabsf(ratio_f - 1.0f);
    const float clip_f = (float)clip;
    cf += dev > clip_f ? 1.0 : 0.0;
    cf_gt += (ratio_f - 1.0f > clip_f) ? 1.0 : 0.0;
    cf_lt += (1.0f - ratio_f > clip_f) ? 1.0 : 0.0;
    if (grad_lp) {
      const bool inside = (ratio_f >= lo_f) && (ratio_f <= hi_f);
      double g;
      if (inside || unclipped > clipped) g = -A * ratio;                  // d(-A rho)/d lp
      else if (unclipped == clipped) g = 0.5 * (-A * ratio);              // torch.maximum tie split (A == 0)
      else g = 0.0;
      grad_lp[i] = (float)(grad_scale * g / (double)B);
    }
  }
  loss = warp_sum(loss); kl = warp_sum(kl); cf = warp_sum(cf);
  cf_gt = warp_sum(cf_gt); cf_lt = warp_sum(cf_lt);
  if (threadIdx.x == 0) {
    const double inv = 1.0 / (double)B;
    out[0] = loss * inv;          // loss (beta = 0: loss == policy_loss)
    out[1] = 0.5 * kl * inv;      // approx_kl, :1132-1135
    out[2] = cf * inv;            // clipfrac
    out[3] = cf_gt * inv;         // clipfrac_gt_one
    out[4] = cf_lt * inv;         // clipfrac_lt_one
    out[5] = loss * inv;          // policy_loss
  }
}

}  // namespace
}  // namespace advgrpo

using namespace advgrpo;

extern "C" {

size_t advgrpo_group_advantage_workspace_bytes(int64_t N, int64_t T) {
  // two 64-bit digests per row + group mean / std per (row, column)
  return (size_t)N * 2 * sizeof(uint64_t) + (size_t)N * (size_t)(T > 0 ? T : 1) * 2 * sizeof(double) + 16;
}

int advgrpo_group_advantage_mode(const float* rewards, const int64_t* group_keys, int64_t key_len,
                                 int64_t N, int64_t T, int global_std, int mode, double* advantages,
                                 double* stats, void* workspace, size_t workspace_bytes,
                                 advgrpo_stream_t stream) {
  ADVGRPO_CHECK_ARG(rewards && group_keys && advantages, "group_advantage: null pointer");
  ADVGRPO_CHECK_ARG(mode >= ADVGRPO_ADV_GRPO && mode <= ADVGRPO_ADV_DPO, "group_advantage: unknown mode %d", mode);
  ADVGRPO_CHECK_ARG(mode != ADVGRPO_ADV_DPO || T == 1, "group_advantage: mode 'dpo' needs 1-D rewards (T = 1)");
  ADVGRPO_CHECK_ARG(N >= 0 && T >= 1 && T <= 64 && key_len >= 1,
                    "group_advantage: need N >= 0, 1 <= T <= 64, key_len >= 1 (got N=%lld T=%lld key_len=%lld)",
                    (long long)N, (long long)T, (long long)key_len);
  if (N == 0) return ADVGRPO_OK;
  if (!workspace || workspace_bytes < advgrpo_group_advantage_workspace_bytes(N, T))
    return set_error(ADVGRPO_ERR_WORKSPACE, "group_advantage: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  uint64_t* h1 = (uint64_t*)workspace;
  uint64_t* h2 = h1 + N;
  const int warps = 8;
  hash_rows_kernel<<<(unsigned)((N + warps - 1) / warps), warps * 32, 0, st>>>(group_keys, key_len, N, h1, h2);
  ADVGRPO_CUDA_LAUNCH_CHECK();
  if (mode == ADVGRPO_ADV_GRPO && N <= kFastMaxN) {
    double* gm = (double*)(h2 + N);
    double* gs = gm + N * T;
    const size_t smem = (size_t)N * 20;
    static bool attr_set = false;
    if (!attr_set) {
      ADVGRPO_CUDA_CALL(cudaFuncSetAttribute(group_advantage_grpo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFastMaxN * 20));
      attr_set = true;
    }
    int64_t work = N * T;
    int threads = work >= 1024 ? 1024 : (int)(((work + 31) / 32) * 32);
    group_advantage_grpo_kernel<<<1, threads, smem, st>>>(rewards, h1, h2, (int)N, (int)T, global_std, advantages, stats, gm, gs);
  } else {
    int threads = N >= 1024 ? 1024 : (int)(((N + 31) / 32) * 32);
    group_advantage_kernel<<<1, threads, 0, st>>>(rewards, h1, h2, N, T, global_std, mode, advantages, stats);
  }
  ADVGRPO_CUDA_LAUNCH_CHECK();
  return ADVGRPO_OK;
}

int advgrpo_group_advantage(const float* rewards, const int64_t* group_keys, int64_t key_len,
                            int64_t N, int64_t T, int global_std, double* advantages,
                            double* stats, void* workspace, size_t workspace_bytes,
                            advgrpo_stream_t stream) {
  return advgrpo_group_advantage_mode(rewards, group_keys, key_len, N, T, global_std, ADVGRPO_ADV_GRPO, advantages, stats,
                                      workspace, workspace_bytes, stream);
}

int advgrpo_grpo_clip_loss(const float* log_prob, const float* old_log_prob,
                           const double* advantages, int64_t adv_stride, int64_t B,
                           double clip_range, double adv_clip_max, double grad_scale, double* out,
                           float* grad_log_prob, advgrpo_stream_t stream) {
  ADVGRPO_CHECK_ARG(log_prob && old_log_prob && advantages && out, "grpo_clip_loss: null pointer");
  ADVGRPO_CHECK_ARG(B >= 1 && adv_stride >= 1, "grpo_clip_loss: B and adv_stride must be >= 1");
  grpo_clip_loss_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(log_prob, old_log_prob, advantages,
                                                            adv_stride, B, clip_range, adv_clip_max,
                                                            grad_scale, out, grad_log_prob);
  ADVGRPO_CUDA_LAUNCH_CHECK();
  return ADVGRPO_OK;
}

}  // extern "C"
