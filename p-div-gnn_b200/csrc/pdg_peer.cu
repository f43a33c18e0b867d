// One-shot all-reduce (mean) of the flat gradient over NVLink peer memory.
//
// The data-parallel path exchanges ONE 669 KB buffer per step.  At that size a ring / tree collective is all latency:
// every rank instead PUSHES its gradient into a slot of each peer's exchange area (CUDA IPC mapping, NVLink P2P through
// NVSwitch; posted stores, no read round trips), raises a flag in every peer, waits for its own flags and sums the slots
// it now holds locally -- in rank order, so every rank computes bit-identical means (the parameters stay identical
// without a broadcast).
//   exchange area of a rank:  float x[2][world][n_pad]   slot [parity of the sequence number][source rank]
//                             int   ctrl[64]             [16 p + r]: sequence number rank r has delivered for parity p
//                                                        [32 + p]: CTA arrival counter (local)
// One flag barrier per step is enough: a peer can push step s+2 (same parity as s) only after its all-reduce of s+1
// has completed, which needs THIS rank's flag s+1, raised after this rank finished reading step s.
// A rank that never arrives would leave the others spinning: the wait gives up after PDG_PEER_TIMEOUT_CYCLES and
// reports it through *status (device int); the results are then undefined but nothing hangs.
#include "pdg_common.cuh"

namespace pdg {

constexpr int PEER_MAX = 16;
constexpr int PEER_CTRL_INTS = 64;
constexpr long long PDG_PEER_TIMEOUT_CYCLES = 6000000000ll;  // ~3 s at 1.9 GHz

struct PeerPtrs {
  float* p[PEER_MAX];
};

__device__ __forceinline__ void st_release_sys(int* p, int v) { asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ int ld_acquire_sys(const int* p) {
  int v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_slot(const float* p) {  // slots are written by peers: read at L2, never from a stale L1 line
  float4 v;
  asm("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

template <int WORLD>
__device__ __forceinline__ void reduce_slots(float* __restrict__ flat, const float* __restrict__ slots, int n, int n_pad, int world,
                                              int gtid, int gsz) {
  const int n4 = n >> 2;
  const float inv = 1.f / (float)world;
  for (int i = gtid; i < n4; i += gsz) {
    float4 v[WORLD];
#pragma unroll
    for (int r = 0; r < WORLD; ++r)
      if (r < world) v[r] = ld_slot(slots + (size_t)r * n_pad + 4 * (size_t)i);  // all loads in flight, then the sum in rank order
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int r = 0; r < WORLD; ++r)
      if (r < world) { s.x += v[r].x; s.y += v[r].y; s.z += v[r].z; s.w += v[r].w; }
    reinterpret_cast<float4*>(flat)[i] = make_float4(s.x * inv, s.y * inv, s.z * inv, s.w * inv);
  }
  if (gtid == 0)
    for (int i = 4 * n4; i < n; ++i) {
      float s = 0.f;
      for (int r = 0; r < world; ++r) s += __ldcg(slots + (size_t)r * n_pad + i);
      flat[i] = s * inv;
    }
}

__global__ void __launch_bounds__(256)
k_allreduce_mean(float* __restrict__ flat, int n, int n_pad, PeerPtrs P, int rank, int world, int seq, int* __restrict__ status) {
  pdl_sync();
  const int par = seq & 1;
  const size_t slot_off = ((size_t)par * world + rank) * n_pad;  // my slot inside every rank's area
  const size_t ctrl_off = 2 * (size_t)world * n_pad;
  int* ctrl = reinterpret_cast<int*>(P.p[rank] + ctrl_off);
  const int tid = threadIdx.x, gtid = blockIdx.x * blockDim.x + tid, gsz = gridDim.x * blockDim.x;
  const int n4 = n >> 2;  // whole float4s; the 0..3 trailing floats go one by one (thread 0 of the grid)
  // 1. push my gradient into my slot of every rank's area (own area included)
  for (int i = gtid; i < n4; i += gsz) {
    const float4 v = reinterpret_cast<const float4*>(flat)[i];
    for (int r = 0; r < world; ++r) reinterpret_cast<float4*>(P.p[r] + slot_off)[i] = v;
  }
  if (gtid == 0)
    for (int i = 4 * n4; i < n; ++i)
      for (int r = 0; r < world; ++r) P.p[r][slot_off + i] = flat[i];
  __threadfence_system();
  __syncthreads();
  __shared__ int s_last, s_bad;
  if (tid == 0) {
    s_bad = 0;
    const int done = atomicAdd(&ctrl[32 + par], 1);
    s_last = done == (int)gridDim.x - 1;
    if (s_last) ctrl[32 + par] = 0;
    __threadfence_system();
  }
  __syncthreads();
  if (s_last && tid < world)  // last CTA of this rank: everything is out -- tell every rank (itself included)
    st_release_sys(reinterpret_cast<int*>(P.p[tid] + ctrl_off) + 16 * par + rank, seq);
  // 2. wait until every rank has delivered into MY area (local polling)
  if (tid < world) {
    const int* f = ctrl + 16 * par + tid;
    const long long t0 = clock64();
    while (ld_acquire_sys(f) < seq) {
      if (clock64() - t0 > PDG_PEER_TIMEOUT_CYCLES) { s_bad = 1; break; }
      __nanosleep(32);
    }
  }
  __syncthreads();
  if (s_bad) {
    if (tid == 0 && status != nullptr) atomicExch(status, 1);
    return;
  }
  // 3. mean of the slots I hold, in rank order (identical arithmetic on every rank)
  const float* slots = P.p[rank] + (size_t)par * world * n_pad;
  if (world <= 2) reduce_slots<2>(flat, slots, n, n_pad, world, gtid, gsz);
  else if (world <= 4) reduce_slots<4>(flat, slots, n, n_pad, world, gtid, gsz);
  else if (world <= 8) reduce_slots<8>(flat, slots, n, n_pad, world, gtid, gsz);
  else reduce_slots<PEER_MAX>(flat, slots, n, n_pad, world, gtid, gsz);
}

}  // namespace pdg

using namespace pdg;

static size_t peer_pad(int64_t n) { return (size_t)round_up(n, 64); }

extern "C" size_t pdg_peer_bytes(int64_t n_floats, int world) {
  return 2 * (size_t)world * peer_pad(n_floats) * sizeof(float) + PEER_CTRL_INTS * sizeof(int);
}

// The exchange buffer is the one allocation this library makes itself: CUDA IPC needs the base of a cudaMalloc'ed
// range, which a sub-allocating caller (torch's caching allocator) cannot hand over.
extern "C" int pdg_peer_alloc(int64_t n_floats, int world, void** dev_ptr, void* handle64) {
  if (n_floats <= 0 || world < 1 || world > PEER_MAX || dev_ptr == nullptr || handle64 == nullptr) { set_error("pdg_peer_alloc: bad arguments"); return -1; }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  void* p = nullptr;
  const size_t bytes = pdg_peer_bytes(n_floats, world);
  PDG_CUDA_CHECK(cudaMalloc(&p, bytes));
  PDG_CUDA_CHECK(cudaMemset(p, 0, bytes));
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) { cudaFree(p); set_error("cudaIpcGetMemHandle: %s", cudaGetErrorString(e)); return -2; }
  memcpy(handle64, &h, 64);
  *dev_ptr = p;
  return 0;
}
extern "C" int pdg_peer_open(const void* handle64, void** dev_ptr) {
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  cudaError_t e = cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) { set_error("cudaIpcOpenMemHandle: %s", cudaGetErrorString(e)); return -2; }
  return 0;
}
extern "C" int pdg_peer_close(void* dev_ptr) {
  cudaError_t e = cudaIpcCloseMemHandle(dev_ptr);
  if (e != cudaSuccess) { set_error("cudaIpcCloseMemHandle: %s", cudaGetErrorString(e)); return -2; }
  return 0;
}
extern "C" int pdg_peer_free(void* dev_ptr) {
  cudaError_t e = cudaFree(dev_ptr);
  if (e != cudaSuccess) { set_error("cudaFree: %s", cudaGetErrorString(e)); return -2; }
  return 0;
}

extern "C" int pdg_allreduce_mean(float* flat, int64_t n_floats, void* const* peer_bufs_host, int rank, int world, int64_t seq,
                                  int* status, void* stream_) {
  if (world < 1 || world > PEER_MAX || rank < 0 || rank >= world) { set_error("pdg_allreduce_mean: world %d / rank %d (max %d ranks)", world, rank, PEER_MAX); return -1; }
  if (n_floats <= 0 || n_floats > 0x7fffffff) { set_error("pdg_allreduce_mean: bad buffer size"); return -1; }
  if (seq < 1 || seq > 0x7fffffff) { set_error("pdg_allreduce_mean: sequence numbers start at 1"); return -1; }
  if (((uintptr_t)flat & 15) != 0) { set_error("pdg_allreduce_mean: buffer must be 16-byte aligned"); return -1; }
  PeerPtrs P;
  for (int r = 0; r < PEER_MAX; ++r) P.p[r] = r < world ? (float*)peer_bufs_host[r] : nullptr;
  const int n_pad = (int)peer_pad(n_floats);
  const int grid = 96;
  cudaError_t e = launch_pdl(k_allreduce_mean, dim3(grid), dim3(256), 0, (cudaStream_t)stream_, flat, (int)n_floats, n_pad, P, rank, world, (int)seq, status);
  if (e != cudaSuccess) { set_error("k_allreduce_mean launch: %s", cudaGetErrorString(e)); return -2; }
  count_launches(1);
  return 0;
}
