// Fused Adam over the 28 parameter tensors + GradScaler-style non-finite check.
// Replaces `scaler.step(optimizer)` / `torch.optim.Adam.step()` of the reference train loop
// (scripts/gnn_train.py:111,118,204-207): the foreach implementation costs ~12 launches per step, this is one.
#include <math.h>

#include "pdg_common.cuh"

namespace pdg {

struct AdamPtrs {
  float* p[PDG_NUM_PARAMS];
  const float* g[PDG_NUM_PARAMS];
};
struct ParamTable { int off[PDG_NUM_PARAMS + 1]; };
static ParamTable make_table() {
  ParamTable t;
  int o = 0;
  for (int i = 0; i < PDG_NUM_PARAMS; ++i) { t.off[i] = o; o += param_size(i); }
  t.off[PDG_NUM_PARAMS] = o;
  return t;
}
// flat element index -> (tensor, element): the 29 offsets are kernel parameters, the search is 5 compares
__device__ __forceinline__ int find_tensor(const ParamTable& tab, int i) {
  int lo = 0, hi = PDG_NUM_PARAMS;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (i >= tab.off[mid]) lo = mid; else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(256)
k_grads_check_finite(AdamPtrs ptrs, ParamTable tab, int* __restrict__ found_inf) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  bool bad = false;
  if (i < PDG_PARAM_ELEMS) {
    const int t = find_tensor(tab, i);
    const float g = ptrs.g[t][i - tab.off[t]];
    bad = !isfinite(g);
  }
  if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(found_inf, 1);
}

// torch.optim.Adam (amsgrad=False, maximize=False), same operation order as torch's _multi_tensor_adam:
//   g' = g * inv_scale + weight_decay * p;  m = m + (1 - b1) (g' - m);  v = b2 v + (1 - b2) g'^2
//   p -= step_size * m / (sqrt(v) / sqrt(1 - b2^t) + eps),   step_size = lr / (1 - b1^t)
// step_dev != nullptr: the step count lives on the DEVICE (number of updates applied so far) and the bias corrections
// are formed here from it, in double like the host path -- a skipped step (found_inf) then leaves the count alone,
// exactly like GradScaler.step, which does not call optimizer.step() on overflow (gnn_train.py:205-207).
__global__ void __launch_bounds__(256)
k_adam_step(AdamPtrs ptrs, ParamTable tab, float* __restrict__ exp_avg, float* __restrict__ exp_avg_sq, float one_m_b1,
            float b2, float one_m_b2, float step_size, float bc2_sqrt, float eps, float weight_decay, float inv_scale,
            const int* __restrict__ found_inf, const int* __restrict__ step_dev, double lr, double beta1, double beta2) {
  if (found_inf != nullptr && *found_inf != 0) return;  // GradScaler: skip the whole step
  if (step_dev != nullptr) {
    __shared__ float s_cfg[2];
    if (threadIdx.x == 0) {
      const double t = (double)(*step_dev + 1);
      s_cfg[0] = (float)(lr / (1.0 - pow(beta1, t)));
      s_cfg[1] = (float)sqrt(1.0 - pow(beta2, t));
    }
    __syncthreads();
    step_size = s_cfg[0];
    bc2_sqrt = s_cfg[1];
  }
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= PDG_PARAM_ELEMS) return;
  const int t = find_tensor(tab, i);
  const int e = i - tab.off[t];
  float p = ptrs.p[t][e];
  float g = ptrs.g[t][e] * inv_scale;
  if (weight_decay != 0.f) g = fmaf(weight_decay, p, g);
  float m = exp_avg[i], v = exp_avg_sq[i];
  m = m + one_m_b1 * (g - m);
  v = v * b2 + one_m_b2 * g * g;
  const float denom = sqrtf(v) / bc2_sqrt + eps;
  p = p - step_size * (m / denom);
  exp_avg[i] = m;
  exp_avg_sq[i] = v;
  ptrs.p[t][e] = p;
}

// runs after k_adam_step (stream order): count the update unless it was skipped
__global__ void k_adam_advance(const int* __restrict__ found_inf, int* __restrict__ step_dev) {
  if (found_inf == nullptr || *found_inf == 0) *step_dev += 1;
}

}  // namespace pdg

using namespace pdg;

static int fill_ptrs(AdamPtrs& P, float* const* params, const float* const* grads, const char* who) {
  for (int i = 0; i < PDG_NUM_PARAMS; ++i) {
    P.p[i] = params ? params[i] : nullptr;
    P.g[i] = grads[i];
    if (grads[i] == nullptr || (params && params[i] == nullptr)) { set_error("%s: null pointer for tensor %d", who, i); return -1; }
  }
  return 0;
}

extern "C" int pdg_grads_check_finite(const float* const* grads, int* found_inf, void* stream) {
  AdamPtrs P;
  if (found_inf == nullptr) { set_error("pdg_grads_check_finite: found_inf is null"); return -1; }
  if (fill_ptrs(P, nullptr, grads, "pdg_grads_check_finite")) return -1;
  k_grads_check_finite<<<(PDG_PARAM_ELEMS + 255) / 256, 256, 0, (cudaStream_t)stream>>>(P, make_table(), found_inf);
  PDG_LAUNCH_CHECK();
  return 0;
}

static int adam_launch(float* const* params, const float* const* grads, float* exp_avg, float* exp_avg_sq,
                       const pdg_adam_t* cfg, const int* found_inf, int* step_dev, void* stream) {
  if (cfg == nullptr || exp_avg == nullptr || exp_avg_sq == nullptr) { set_error("pdg_adam_step: null argument"); return -1; }
  if (step_dev == nullptr && cfg->step < 1) { set_error("pdg_adam_step: step must be >= 1 (got %d)", cfg->step); return -1; }
  if (!(cfg->beta1 >= 0. && cfg->beta1 < 1. && cfg->beta2 >= 0. && cfg->beta2 < 1.) || !(cfg->eps >= 0.) || !(cfg->lr >= 0.)) {
    set_error("pdg_adam_step: invalid hyper-parameters");
    return -1;
  }
  AdamPtrs P;
  if (fill_ptrs(P, params, grads, "pdg_adam_step")) return -1;
  // bias corrections in double on the host, like torch's Python scalars
  const int host_step = cfg->step < 1 ? 1 : cfg->step;  // unused by the kernel when step_dev is given
  const double bc1 = 1.0 - pow(cfg->beta1, (double)host_step);
  const double bc2 = 1.0 - pow(cfg->beta2, (double)host_step);
  const float step_size = (float)(cfg->lr / bc1);
  const float bc2_sqrt = (float)sqrt(bc2);
  k_adam_step<<<(PDG_PARAM_ELEMS + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
      P, make_table(), exp_avg, exp_avg_sq, (float)(1.0 - cfg->beta1), (float)cfg->beta2, (float)(1.0 - cfg->beta2),
      step_size, bc2_sqrt, (float)cfg->eps, (float)cfg->weight_decay, (float)cfg->inv_scale, found_inf, step_dev, cfg->lr,
      cfg->beta1, cfg->beta2);
  PDG_LAUNCH_CHECK();
  if (step_dev != nullptr) {
    k_adam_advance<<<1, 1, 0, (cudaStream_t)stream>>>(found_inf, step_dev);
    PDG_LAUNCH_CHECK();
  }
  return 0;
}

extern "C" int pdg_adam_step(float* const* params, const float* const* grads, float* exp_avg, float* exp_avg_sq,
                             const pdg_adam_t* cfg, const int* found_inf, void* stream) {
  return adam_launch(params, grads, exp_avg, exp_avg_sq, cfg, found_inf, nullptr, stream);
}

extern "C" int pdg_adam_step_counted(float* const* params, const float* const* grads, float* exp_avg, float* exp_avg_sq,
                                     const pdg_adam_t* cfg, const int* found_inf, int* step_count, void* stream) {
  if (step_count == nullptr) { set_error("pdg_adam_step_counted: step_count is null"); return -1; }
  return adam_launch(params, grads, exp_avg, exp_avg_sq, cfg, found_inf, step_count, stream);
}
