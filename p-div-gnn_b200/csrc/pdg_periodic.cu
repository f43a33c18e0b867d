// Device-side periodicity check of a batch of 2-D meshes.
//
// Replaces microgen.mesh.is_periodic / microgen.remesh.is_periodic, which the reference asserts on the host for every
// mesh it generates or benchmarks (generate_dataset.py:191, generate_dataset_hyperelast.py:160,237,
// benchmark_gnn_fem.py:195, always on the [N,2] in-plane coordinates with tol = 1e-8):
//   * nodes within tol of the bounding-box minimum / maximum of an axis are the two opposite sides of that axis;
//   * each side is sorted along the other axis;
//   * periodic <=> opposite sides hold the same number of nodes and no sorted coordinate of the max side exceeds its
//     partner on the min side by more than tol (the published test is one-sided, no absolute value).
// Whole batch at once: 4 candidate entries per node (left, right, bottom, top), keyed by the order-preserving bit image
// of the coordinate along the side, segment id = 4 * graph + side (non-members get a sentinel segment); one stable
// 64-bit radix sort by coordinate followed by one stable sort by segment id leaves every side contiguous and sorted;
// one CTA per mesh then compares the partners.  No host sync, no size limit on a side.
#include <cub/cub.cuh>

#include "pdg_common.cuh"

namespace pdg {

typedef unsigned long long u64;

struct PeriodicLayout {
  size_t off_box, off_val, off_val2, off_seg, off_seg2, off_sort, sort_bytes, total;
  PeriodicLayout(int64_t n, int64_t b) {
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o += (size_t)round_up((int64_t)bytes, 256); return r; };
    off_box = take((size_t)b * 4 * sizeof(double));
    off_val = take((size_t)4 * n * sizeof(u64));
    off_val2 = take((size_t)4 * n * sizeof(u64));
    off_seg = take((size_t)4 * n * sizeof(uint32_t));
    off_seg2 = take((size_t)4 * n * sizeof(uint32_t));
    size_t s1 = 0, s2 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, s1, (const u64*)nullptr, (u64*)nullptr, (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)(4 * n));
    cub::DeviceRadixSort::SortPairs(nullptr, s2, (const uint32_t*)nullptr, (uint32_t*)nullptr, (const u64*)nullptr, (u64*)nullptr, (int)(4 * n));
    sort_bytes = s1 > s2 ? s1 : s2;
    off_sort = take(sort_bytes);
    total = o;
  }
};

// order-preserving map double -> u64 (and back): negative values flip all bits, the others set the sign bit
__device__ __forceinline__ u64 ordered_bits(double x) {
  const u64 u = (u64)__double_as_longlong(x);
  return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double ordered_value(u64 k) {
  const u64 u = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
  return __longlong_as_double((long long)u);
}

constexpr int PER_NT = 256;
// bounding box of every mesh: box[g] = {xmin, xmax, ymin, ymax}
__global__ void __launch_bounds__(PER_NT)
k_periodic_bbox(const double* __restrict__ pos, const int64_t* __restrict__ node_ptr, double* __restrict__ box) {
  __shared__ double s_box[4][PER_NT / 32];
  const int g = blockIdx.x, tid = threadIdx.x;
  const int64_t n0 = node_ptr[g], n1 = node_ptr[g + 1];
  double xmin = 1e300, xmax = -1e300, ymin = 1e300, ymax = -1e300;
  for (int64_t n = n0 + tid; n < n1; n += PER_NT) {
    const double x = pos[2 * n], y = pos[2 * n + 1];
    xmin = fmin(xmin, x); xmax = fmax(xmax, x); ymin = fmin(ymin, y); ymax = fmax(ymax, y);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    xmin = fmin(xmin, __shfl_xor_sync(0xffffffffu, xmin, o)); xmax = fmax(xmax, __shfl_xor_sync(0xffffffffu, xmax, o));
    ymin = fmin(ymin, __shfl_xor_sync(0xffffffffu, ymin, o)); ymax = fmax(ymax, __shfl_xor_sync(0xffffffffu, ymax, o));
  }
  if ((tid & 31) == 0) { s_box[0][tid >> 5] = xmin; s_box[1][tid >> 5] = xmax; s_box[2][tid >> 5] = ymin; s_box[3][tid >> 5] = ymax; }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < PER_NT / 32; ++w) {
      s_box[0][0] = fmin(s_box[0][0], s_box[0][w]); s_box[1][0] = fmax(s_box[1][0], s_box[1][w]);
      s_box[2][0] = fmin(s_box[2][0], s_box[2][w]); s_box[3][0] = fmax(s_box[3][0], s_box[3][w]);
    }
    for (int q = 0; q < 4; ++q) box[4 * g + q] = s_box[q][0];
  }
}

// entry 4 n + s (s: 0 left, 1 right, 2 bottom, 3 top): coordinate ALONG the side + segment id
__global__ void k_periodic_entries(const double* __restrict__ pos, const int64_t* __restrict__ node_ptr, int B, int64_t N,
                                   const double* __restrict__ box, double tol, u64* __restrict__ val, uint32_t* __restrict__ seg) {
  const int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (n >= N) return;
  int lo = 0, hi = B;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (node_ptr[mid] <= n) lo = mid; else hi = mid;
  }
  const double c[2] = {pos[2 * n], pos[2 * n + 1]};
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    const int axis = s >> 1;
    const bool member = fabs(c[axis] - box[4 * lo + s]) < tol;  // np.abs(crd[:, axis] - bound) < tol
    val[4 * n + s] = ordered_bits(c[1 - axis]);
    seg[4 * n + s] = member ? (uint32_t)(4 * lo + s) : (uint32_t)(4 * B);
  }
}

__device__ __forceinline__ int64_t lower_bound_u32(const uint32_t* __restrict__ a, int64_t n, uint32_t v) {
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (a[mid] < v) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// seg / val: sorted by (segment, coordinate).  One CTA per mesh.
__global__ void __launch_bounds__(PER_NT)
k_periodic_check(const uint32_t* __restrict__ seg, const u64* __restrict__ val, int64_t M, double tol, int32_t* __restrict__ periodic) {
  __shared__ int64_t s_start[5];
  const int g = blockIdx.x, tid = threadIdx.x;
  if (tid < 5) s_start[tid] = lower_bound_u32(seg, M, (uint32_t)(4 * g + tid));
  __syncthreads();
  int bad = 0;
#pragma unroll
  for (int axis = 0; axis < 2; ++axis) {
    const int64_t l0 = s_start[2 * axis], h0 = s_start[2 * axis + 1], h1 = s_start[2 * axis + 2];
    const int64_t cl = h0 - l0, chh = h1 - h0;
    if (cl != chh) { bad = 1; continue; }
    for (int64_t k = tid; k < cl; k += PER_NT)
      if (ordered_value(val[h0 + k]) - ordered_value(val[l0 + k]) > tol) bad = 1;
  }
  bad = __syncthreads_or(bad);
  if (tid == 0) periodic[g] = bad ? 0 : 1;
}

}  // namespace pdg

using namespace pdg;

extern "C" size_t pdg_periodic_tmp_bytes(int64_t n_nodes, int64_t n_graphs) {
  if (n_nodes <= 0 || n_graphs <= 0) return 0;
  return PeriodicLayout(n_nodes, n_graphs).total;
}

extern "C" int pdg_is_periodic(const double* pos, const int64_t* node_ptr, int64_t n_graphs, int64_t n_nodes, double tol, void* tmp,
                               size_t tmp_bytes, int32_t* periodic, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  if (n_graphs <= 0 || n_nodes <= 0) { set_error("pdg_is_periodic: empty batch"); return -1; }
  if (4 * n_nodes >= (1ll << 31) || 4 * n_graphs + 1 >= (1ll << 31)) { set_error("pdg_is_periodic: batch too large for int32 entry ids"); return -1; }
  if (!(tol >= 0.0)) { set_error("pdg_is_periodic: tolerance must be >= 0"); return -1; }
  PeriodicLayout L(n_nodes, n_graphs);
  if (tmp_bytes < L.total) { set_error("pdg_is_periodic: workspace %zu < %zu", tmp_bytes, L.total); return -1; }
  char* base = (char*)tmp;
  double* box = (double*)(base + L.off_box);
  u64* val = (u64*)(base + L.off_val);
  u64* val2 = (u64*)(base + L.off_val2);
  uint32_t* seg = (uint32_t*)(base + L.off_seg);
  uint32_t* seg2 = (uint32_t*)(base + L.off_seg2);
  const int M = (int)(4 * n_nodes);
  k_periodic_bbox<<<(unsigned)n_graphs, PER_NT, 0, st>>>(pos, node_ptr, box);
  PDG_LAUNCH_CHECK();
  k_periodic_entries<<<(unsigned)((n_nodes + 255) / 256), 256, 0, st>>>(pos, node_ptr, (int)n_graphs, n_nodes, box, tol, val, seg);
  PDG_LAUNCH_CHECK();
  size_t sb = L.sort_bytes;
  PDG_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(base + L.off_sort, sb, val, val2, seg, seg2, M, 0, 64, st));
  int end_bit = 1;
  while (end_bit < 32 && ((uint32_t)(4 * n_graphs) >> end_bit)) ++end_bit;
  sb = L.sort_bytes;
  PDG_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(base + L.off_sort, sb, seg2, seg, val2, val, M, 0, end_bit, st));
  count_launches(8);
  k_periodic_check<<<(unsigned)n_graphs, PER_NT, 0, st>>>(seg, val, (int64_t)M, tol, periodic);
  PDG_LAUNCH_CHECK();
  return 0;
}
