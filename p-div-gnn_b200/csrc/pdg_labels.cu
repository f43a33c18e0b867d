// Device-side node labelling of plate-with-hole meshes.
//
// Replaces datasets.compute_node_labels (datasets.py:133-179), which runs three VTK filters per mesh on the
// host (extract_feature_edges(boundary_edges) -> connectivity() -> cell_data_to_point_data()):
//   * a mesh edge used by exactly ONE cell (triangle or quad) is a boundary edge;
//   * the boundary edges form closed loops = connected regions;
//   * the region that touches the mesh bounding box is the EXTERNAL boundary (the reference takes VTK's
//     RegionId 0 and swaps when its first point is not on the bounds, datasets.py:120-130,166-172: same set),
//     every other region an INTERNAL boundary (the hole).
// labels: NodeType  -1 internal boundary, 0 internal, 1 external boundary (datasets.py:33-36), int64 like
// the reference's np.int_.  n_regions[g] lets the caller repeat the reference's `assert n_regions == 2`.
// One stable radix sort of the 3F undirected edge keys for the whole batch, then ONE CTA per mesh.
#include <cub/cub.cuh>

#include "pdg_common.cuh"

namespace pdg {

typedef unsigned long long u64;

struct LabelLayout {
  size_t off_keys, off_keys2, off_comp, off_ext, off_sort, sort_bytes, total;
  LabelLayout(int64_t n, int64_t f, int npf) {
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o += (size_t)round_up((int64_t)bytes, 256); return r; };
    off_keys = take((size_t)npf * f * 8);
    off_keys2 = take((size_t)npf * f * 8);
    off_comp = take((size_t)n * 4);
    off_ext = take((size_t)n * 4);
    sort_bytes = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, sort_bytes, (const u64*)nullptr, (u64*)nullptr, (int)(npf * f));
    off_sort = take(sort_bytes);
    total = o;
  }
};

// one undirected key per face side (3 per triangle, 4 per quad): min(a,b) * Ntot + max(a,b), batch-global node ids
template <int NPF>
__global__ void k_label_edge_keys(const int64_t* __restrict__ faces, int64_t F, const int64_t* __restrict__ node_ptr,
                                  const int64_t* __restrict__ face_ptr, int B, int64_t Ntot, u64* __restrict__ keys) {
  const int64_t f = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (f >= F) return;
  int lo = 0, hi = B;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (face_ptr[mid] <= f) lo = mid; else hi = mid;
  }
  const u64 off = (u64)node_ptr[lo];
  u64 v[NPF];
#pragma unroll
  for (int i = 0; i < NPF; ++i) v[i] = (u64)faces[i * F + f] + off;
  auto key = [&](u64 p, u64 q) { return p < q ? p * (u64)Ntot + q : q * (u64)Ntot + p; };
#pragma unroll
  for (int i = 0; i < NPF; ++i) keys[NPF * f + i] = key(v[i], v[(i + 1) % NPF]);
}

__device__ __forceinline__ int64_t lower_bound_u64(const u64* __restrict__ a, int64_t n, u64 v) {
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (a[mid] < v) lo = mid + 1; else hi = mid;
  }
  return lo;
}

constexpr int LAB_NT = 256;
__global__ void __launch_bounds__(LAB_NT)
k_label_mesh(const double* __restrict__ pos, const u64* __restrict__ keys, int64_t K, u64* __restrict__ bedges,
             const int64_t* __restrict__ node_ptr, int64_t Ntot, int* __restrict__ comp, int* __restrict__ ext,
             int64_t* __restrict__ labels, int* __restrict__ n_regions) {
  __shared__ int s_cnt, s_regions;
  __shared__ double s_box[4][LAB_NT / 32];
  __shared__ double s_b[4];
  const int g = blockIdx.x, tid = threadIdx.x;
  const int64_t n0 = node_ptr[g], n1 = node_ptr[g + 1];
  const int64_t k0 = lower_bound_u64(keys, K, (u64)n0 * (u64)Ntot), k1 = lower_bound_u64(keys, K, (u64)n1 * (u64)Ntot);
  if (tid == 0) { s_cnt = 0; s_regions = 0; }
  // bounding box of the mesh (exact doubles, compared with == below like the reference's `coord in bounds`)
  double xmin = 1e300, xmax = -1e300, ymin = 1e300, ymax = -1e300;
  for (int64_t n = n0 + tid; n < n1; n += LAB_NT) {
    comp[n] = -1;
    ext[n] = 0;
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
    for (int w = 1; w < LAB_NT / 32; ++w) {
      s_box[0][0] = fmin(s_box[0][0], s_box[0][w]); s_box[1][0] = fmax(s_box[1][0], s_box[1][w]);
      s_box[2][0] = fmin(s_box[2][0], s_box[2][w]); s_box[3][0] = fmax(s_box[3][0], s_box[3][w]);
    }
    for (int q = 0; q < 4; ++q) s_b[q] = s_box[q][0];
  }
  __syncthreads();
  // boundary edges = keys that occur exactly once; compacted to the front of this mesh's range of `bedges`
  for (int64_t k = k0 + tid; k < k1; k += LAB_NT) {
    const u64 key = keys[k];
    const bool once = (k == k0 || keys[k - 1] != key) && (k + 1 == k1 || keys[k + 1] != key);
    if (once) {
      const int a = (int)(key / (u64)Ntot), b = (int)(key % (u64)Ntot);
      comp[a] = a;  // benign race: every writer stores the node's own id
      comp[b] = b;
      bedges[k0 + atomicAdd(&s_cnt, 1)] = key;
    }
  }
  __syncthreads();
  const int nb = s_cnt;
  // connected components of the boundary graph: min-label propagation along the edges until nothing changes
  // (labels only decrease and only this CTA touches this mesh: plain __syncthreads rounds)
  for (;;) {
    int changed = 0;
    for (int e = tid; e < nb; e += LAB_NT) {
      const u64 key = bedges[k0 + e];
      const int a = (int)(key / (u64)Ntot), b = (int)(key % (u64)Ntot);
      const int la = comp[a], lb = comp[b];
      if (la < lb) { atomicMin(&comp[b], la); changed = 1; }
      else if (lb < la) { atomicMin(&comp[a], lb); changed = 1; }
    }
    // pointer jumping: follow the label of my label (halves the remaining rounds on long loops)
    __syncthreads();
    for (int64_t n = n0 + tid; n < n1; n += LAB_NT) {
      const int l = comp[n];
      if (l >= 0) {
        const int ll = comp[l];
        if (ll < l) { comp[n] = ll; changed = 1; }
      }
    }
    if (!__syncthreads_or(changed)) break;
  }
  // a region is external when one of its nodes lies on the bounding box
  for (int64_t n = n0 + tid; n < n1; n += LAB_NT) {
    const int l = comp[n];
    if (l >= 0) {
      const double x = pos[2 * n], y = pos[2 * n + 1];
      if (x == s_b[0] || x == s_b[1] || y == s_b[2] || y == s_b[3]) ext[l] = 1;
      if (l == (int)n) atomicAdd(&s_regions, 1);
    }
  }
  __syncthreads();
  for (int64_t n = n0 + tid; n < n1; n += LAB_NT) {
    const int l = comp[n];
    labels[n] = l < 0 ? 0 : (ext[l] ? 1 : -1);
  }
  if (tid == 0) n_regions[g] = s_regions;
}

}  // namespace pdg

using namespace pdg;

extern "C" size_t pdg_labels_tmp_bytes(int64_t n_nodes, int64_t n_faces, int nodes_per_face, int64_t n_graphs) {
  (void)n_graphs;
  if (n_nodes <= 0 || n_faces <= 0 || (nodes_per_face != 3 && nodes_per_face != 4)) return 0;
  return LabelLayout(n_nodes, n_faces, nodes_per_face).total;
}

extern "C" int pdg_node_labels(const double* pos, const int64_t* faces, const int64_t* node_ptr, const int64_t* face_ptr,
                               int64_t n_graphs, int64_t n_nodes, int64_t n_faces, int nodes_per_face, void* tmp,
                               size_t tmp_bytes, int64_t* labels, int32_t* n_regions, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  if (n_graphs <= 0 || n_nodes <= 0 || n_faces <= 0) { set_error("pdg_node_labels: empty batch"); return -1; }
  if (nodes_per_face != 3 && nodes_per_face != 4) { set_error("pdg_node_labels: nodes_per_face = %d (3 or 4)", nodes_per_face); return -1; }
  if (n_nodes >= (1ll << 31) || nodes_per_face * n_faces >= (1ll << 31)) { set_error("pdg_node_labels: batch too large for int32 ids"); return -1; }
  LabelLayout L(n_nodes, n_faces, nodes_per_face);
  if (tmp_bytes < L.total) { set_error("pdg_node_labels: workspace %zu < %zu", tmp_bytes, L.total); return -1; }
  char* base = (char*)tmp;
  u64* keys = (u64*)(base + L.off_keys);
  u64* keys2 = (u64*)(base + L.off_keys2);
  const int64_t K = (int64_t)nodes_per_face * n_faces;
  if (nodes_per_face == 3)
    k_label_edge_keys<3><<<(unsigned)((n_faces + 255) / 256), 256, 0, st>>>(faces, n_faces, node_ptr, face_ptr, (int)n_graphs, n_nodes, keys);
  else
    k_label_edge_keys<4><<<(unsigned)((n_faces + 255) / 256), 256, 0, st>>>(faces, n_faces, node_ptr, face_ptr, (int)n_graphs, n_nodes, keys);
  PDG_LAUNCH_CHECK();
  int end_bit = 1;
  while (end_bit < 64 && ((u64)n_nodes * (u64)n_nodes) >> end_bit) ++end_bit;
  size_t sb = L.sort_bytes;
  PDG_CUDA_CHECK(cub::DeviceRadixSort::SortKeys(base + L.off_sort, sb, keys, keys2, (int)K, 0, end_bit, st));
  count_launches(4);
  // keys2 = sorted keys; keys is reused for the compacted boundary-edge lists
  k_label_mesh<<<(unsigned)n_graphs, LAB_NT, 0, st>>>(pos, keys2, K, keys, node_ptr, n_nodes, (int*)(base + L.off_comp),
                                                      (int*)(base + L.off_ext), labels, n_regions);
  PDG_LAUNCH_CHECK();
  return 0;
}
