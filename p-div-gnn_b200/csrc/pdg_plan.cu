// Graph plan: receiver-sorted edge order + receiver CSR + sender CSR (device side).
//
// The processor aggregates messages at the TARGET node (edge_index[1], PyG
// flow="source_to_target"; reference models.py:215-217 via MessagePassing.propagate).
// Latent edge features live in receiver-sorted order inside the library (internal edge
// order is free: only node fields leave the module and graph-LayerNorm statistics are
// permutation invariant), which makes every segment contiguous and the aggregation
// atomic-free.  A stable LSD radix sort keeps ties in input order, so the plan -- and
// with it every floating-point summation order downstream -- is deterministic.
#include <cub/cub.cuh>

#include "pdg_common.cuh"

namespace pdg {

struct PlanLayout {
  int64_t N, E, E_pad;
  size_t off_perm, off_recv, off_send, off_rowptr, off_sptr, off_slist, off_status, total;
  __host__ PlanLayout(int64_t n, int64_t e) : N(n), E(e) {
    E_pad = round_up(e > 0 ? e : 1, TM);
    size_t o = 0;
    auto take = [&](size_t elems) { size_t r = o; o += round_up((int64_t)(elems * sizeof(int32_t)), 256); return r; };
    off_perm = take(E_pad);
    off_recv = take(E_pad);
    off_send = take(E_pad);
    off_rowptr = take(N + 1);
    off_sptr = take(N + 1);
    off_slist = take(E_pad);
    off_status = take(1);  // != 0: edge_index held ids outside [0, N) (they were clamped; see pdg_plan_status)
    total = o;
  }
};

// Node ids outside [0, N) would make every gather of the forward read out of bounds: they are clamped into range
// (the kernels stay memory-safe whatever the caller passes) and reported through the plan's status word.
__device__ __forceinline__ int32_t checked_id(int64_t v, int64_t N, int32_t* status) {
  if (v < 0 || v >= N) {
    atomicOr(status, 1);
    return v < 0 ? 0 : (int32_t)(N - 1);
  }
  return (int32_t)v;
}
__global__ void k_split_edge_index(const int64_t* __restrict__ ei, int64_t E, int64_t E_pad, int64_t N, int32_t* __restrict__ key_col,
                                   int32_t* __restrict__ iota, int32_t* __restrict__ status) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < E_pad) {
    key_col[i] = i < E ? checked_id(ei[E + i], N, status) : 0x7fffffff;  // padding sorts last
    iota[i] = (int32_t)i;
  }
}
__global__ void k_finish_recv(const int64_t* __restrict__ ei, int64_t E, int64_t E_pad, int64_t N, int32_t* __restrict__ recv,
                              int32_t* __restrict__ perm, int32_t* __restrict__ send, int32_t* __restrict__ status) {
  const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (p < E_pad) {
    if (p < E) {
      send[p] = checked_id(ei[perm[p]], N, status);
    } else {  // padding rows point at node 0 / edge 0 and are masked by position
      recv[p] = 0;
      perm[p] = 0;
      send[p] = 0;
    }
  }
}
// rowptr[n] = first position p with key[p] >= n  (key sorted ascending, length E)
__global__ void k_lower_bound(const int32_t* __restrict__ key, int64_t E, int64_t N, int32_t* __restrict__ rowptr) {
  const int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (n <= N) {
    int64_t lo = 0, hi = E;
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      if (key[mid] < (int32_t)n) lo = mid + 1; else hi = mid;
    }
    rowptr[n] = (int32_t)lo;
  }
}
__global__ void k_iota(int32_t* __restrict__ v, int64_t n) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) v[i] = (int32_t)i;
}

static size_t sort_tmp_bytes(int64_t n) {
  size_t b = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, b, (const int32_t*)nullptr, (int32_t*)nullptr, (const int32_t*)nullptr,
                                  (int32_t*)nullptr, (int)n);
  return b;
}

}  // namespace pdg

using namespace pdg;

extern "C" size_t pdg_plan_bytes(int64_t n_nodes, int64_t n_edges) { return PlanLayout(n_nodes, n_edges).total; }

extern "C" size_t pdg_plan_tmp_bytes(int64_t n_nodes, int64_t n_edges) {
  const int64_t E_pad = round_up(n_edges > 0 ? n_edges : 1, TM);
  return (size_t)round_up((int64_t)sort_tmp_bytes(E_pad), 256) + 3 * (size_t)round_up(E_pad * 4, 256);
}

extern "C" int pdg_plan_views(void* plan, int64_t n_nodes, int64_t n_edges, int32_t** perm, int32_t** recv,
                              int32_t** send, int32_t** rowptr, int32_t** send_ptr, int32_t** send_list) {
  PlanLayout L(n_nodes, n_edges);
  char* b = (char*)plan;
  if (perm) *perm = (int32_t*)(b + L.off_perm);
  if (recv) *recv = (int32_t*)(b + L.off_recv);
  if (send) *send = (int32_t*)(b + L.off_send);
  if (rowptr) *rowptr = (int32_t*)(b + L.off_rowptr);
  if (send_ptr) *send_ptr = (int32_t*)(b + L.off_sptr);
  if (send_list) *send_list = (int32_t*)(b + L.off_slist);
  return 0;
}

extern "C" int pdg_plan_build(const int64_t* edge_index, int64_t n_nodes, int64_t n_edges, void* plan, void* tmp,
                              size_t tmp_bytes, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  if (n_nodes <= 0 || n_edges <= 0 || n_nodes >= 0x7fffffff || n_edges >= 0x7fffffff - TM) {
    set_error("pdg_plan_build: N=%lld E=%lld out of range", (long long)n_nodes, (long long)n_edges);
    return -1;
  }
  if (tmp_bytes < pdg_plan_tmp_bytes(n_nodes, n_edges)) {
    set_error("pdg_plan_build: tmp too small");
    return -1;
  }
  PlanLayout L(n_nodes, n_edges);
  int32_t *perm, *recv, *send, *rowptr, *sptr, *slist;
  pdg_plan_views(plan, n_nodes, n_edges, &perm, &recv, &send, &rowptr, &sptr, &slist);
  const int64_t E = n_edges, E_pad = L.E_pad, N = n_nodes;
  size_t sb = sort_tmp_bytes(E_pad);
  char* t = (char*)tmp;
  void* sort_tmp = t;
  t += round_up((int64_t)sb, 256);
  int32_t* key_in = (int32_t*)t;
  t += round_up(E_pad * 4, 256);
  int32_t* val_in = (int32_t*)t;
  t += round_up(E_pad * 4, 256);
  int32_t* key_out = (int32_t*)t;
  const int TB = 256;
  const int gb = (int)((E_pad + TB - 1) / TB);
  int bits = 1;
  while ((1ll << bits) <= N && bits < 31) ++bits;  // keys < N, padding key handled below
  // receiver sort: key = col (target), value = input edge id
  int32_t* status = (int32_t*)((char*)plan + L.off_status);
  PDG_CUDA_CHECK(cudaMemsetAsync(status, 0, sizeof(int32_t), st));
  k_split_edge_index<<<gb, TB, 0, st>>>(edge_index, E, E_pad, N, key_in, val_in, status);
  PDG_LAUNCH_CHECK();
  PDG_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(sort_tmp, sb, key_in, recv, val_in, perm, (int)E_pad, 0, 32, st));
  k_finish_recv<<<gb, TB, 0, st>>>(edge_index, E, E_pad, N, recv, perm, send, status);
  PDG_LAUNCH_CHECK();
  k_lower_bound<<<(int)((N + 1 + TB - 1) / TB), TB, 0, st>>>(recv, E, N, rowptr);
  PDG_LAUNCH_CHECK();
  // sender CSR over sorted positions: key = send[p], value = p   (only the E real edges)
  k_iota<<<gb, TB, 0, st>>>(val_in, E);
  PDG_LAUNCH_CHECK();
  PDG_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(sort_tmp, sb, send, key_out, val_in, slist, (int)E, 0, bits, st));
  k_lower_bound<<<(int)((N + 1 + TB - 1) / TB), TB, 0, st>>>(key_out, E, N, sptr);
  PDG_LAUNCH_CHECK();
  return 0;
}

// Optional validation (the ONE entry point of the plan family that synchronises): copies the plan's status word to the
// host and waits for it.  *status_host != 0: edge_index held node ids outside [0, n_nodes) -- the reference would
// have raised an IndexError in x[col] (models.py:233-238); here they were clamped, so results are defined but
// meaningless.  Returns -3 in that case (0 otherwise) with the text in pdg_last_error().
extern "C" int pdg_plan_status(const void* plan, int64_t n_nodes, int64_t n_edges, int32_t* status_host, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  PlanLayout L(n_nodes, n_edges);
  int32_t v = 0;
  PDG_CUDA_CHECK(cudaMemcpyAsync(&v, (const char*)plan + L.off_status, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  PDG_CUDA_CHECK(cudaStreamSynchronize(st));
  if (status_host) *status_host = v;
  if (v != 0) {
    set_error("pdg_plan_status: edge_index contains node ids outside [0, %lld)", (long long)n_nodes);
    return -3;
  }
  return 0;
}
