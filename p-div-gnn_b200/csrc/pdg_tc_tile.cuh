// Row-per-thread / column-per-thread helpers shared by the tcgen05 tile kernels.
#pragma once
#include "pdg_common.cuh"
#include "pdg_tc.cuh"

namespace pdg {

// thread = (row, 64-column half); j = 16-byte chunk (8 columns) inside the half
__device__ __forceinline__ void row_store8(uint8_t* tile, int row, int half, int j, const float* v8) {
  *reinterpret_cast<uint4*>(tile + tc::sw128_chunk(row, half * 8 + j)) = tc::pack8_bf16(v8);
}
__device__ __forceinline__ void row_load8(const uint8_t* tile, int row, int half, int j, float* v8) {
  const uint4 u = *reinterpret_cast<const uint4*>(tile + tc::sw128_chunk(row, half * 8 + j));
  const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
  const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
  const float2 c = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.z));
  const float2 d = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.w));
  v8[0] = a.x; v8[1] = a.y; v8[2] = b.x; v8[3] = b.y; v8[4] = c.x; v8[5] = c.y; v8[6] = d.x; v8[7] = d.y;
}
__device__ __forceinline__ float tile_elem(const uint8_t* tile, int r, int c) {
  return __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(tile + tc::sw128_off(r, c)));
}
// column-thread partial sum over rows [hf*64, hf*64+64) of channel (tid & 127)
__device__ __forceinline__ float tile_colsum_bf16(const uint8_t* tile) {
  const int ch = threadIdx.x & (H - 1), hf = threadIdx.x >> 7;
  float s = 0.f;
#pragma unroll 8
  for (int r = hf * 64; r < hf * 64 + 64; ++r) s += tile_elem(tile, r, ch);
  return s;
}
__device__ __forceinline__ void tile_segsum_bf16(const uint8_t* tile, const int* recv_s, const int32_t* __restrict__ rowptr,
                                                 int row0, int nvalid, int sp, float* __restrict__ dst) {
  const int ch = threadIdx.x & (H - 1), half = threadIdx.x >> 7;
  const int r0 = half ? sp : 0, r1 = half ? nvalid : sp;
  float seg = 0.f;
  for (int r = r0; r < r1; ++r) {
    seg += tile_elem(tile, r, ch);
    if (r == r1 - 1 || recv_s[r + 1] != recv_s[r]) {
      const int c = recv_s[r];
      const int lo = rowptr[c], hi = rowptr[c + 1];
      float* d = dst + (size_t)c * H + ch;
      if (lo >= row0 + r0 && hi <= row0 + r1) *d = seg; else atomicAdd(d, seg);
      seg = 0.f;
    }
  }
}
// fp32 staging tile [128][128] with the float4-chunk index XOR-swizzled by the row, so that both the
// row-per-thread writes and the column-per-thread reads are bank-conflict free
__device__ __forceinline__ float* s32_ptr(float* S, int r, int c) { return S + r * H + ((((c >> 2) ^ (r & 31)) << 2) | (c & 3)); }


// thread context of the 256-thread tile kernels: TMEM lane = tile row, two 64-column halves
struct TcThread {
  int tid, warp, lane, row, half;
  uint32_t lane_base;
  __device__ __forceinline__ TcThread() {
    tid = threadIdx.x; warp = tid >> 5; lane = tid & 31;
    row = 32 * (warp & 3) + lane; half = warp >> 2;
    lane_base = (uint32_t)(32 * (warp & 3)) << 16;
  }
};
// 1024-byte aligned base of the dynamic shared memory window
__device__ __forceinline__ uint8_t* tc_smem_base(uint8_t* raw) { return raw + ((1024u - (tc::smem_u32(raw) & 1023u)) & 1023u); }
// write 64 fp32 values of this thread's row/half to a global fp32 row-major [.,128] array
__device__ __forceinline__ void row_store_global32(float* __restrict__ dst_row_half, const float (&v)[32], int hh) {
#pragma unroll
  for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(dst_row_half + hh * 32 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
}
// column-thread combine of two row-half partials (threads tid and tid+128 share a channel)
__device__ __forceinline__ void colpart_flush(float v, float* comb, float* dst, bool add) {
  __syncthreads();
  comb[(threadIdx.x >> 7) * H + (threadIdx.x & (H - 1))] = v;
  __syncthreads();
  if (threadIdx.x < H) {
    const float s = comb[threadIdx.x] + comb[H + threadIdx.x];
    dst[threadIdx.x] = add ? dst[threadIdx.x] + s : s;
  }
}
// column sums of a [128][64-per-thread] fp32 row fragment through the swizzled fp32 staging tile
__device__ __forceinline__ float s32_colsum(float* S32) {
  const int chn = threadIdx.x & (H - 1), hf = threadIdx.x >> 7;
  float s = 0.f;
#pragma unroll 8
  for (int r = hf * 64; r < hf * 64 + 64; ++r) s += *s32_ptr(S32, r, chn);
  return s;
}

}  // namespace pdg
