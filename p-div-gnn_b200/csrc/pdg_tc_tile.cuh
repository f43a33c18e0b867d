// Row-per-thread / column-per-thread helpers shared by the tcgen05 tile kernels.
#pragma once
#include "pdg_common.cuh"
#include "pdg_tc.cuh"

namespace pdg {

// 8 fp16 (one 16-byte chunk) -> floats: forward-valued tiles / rows
__device__ __forceinline__ void unpack8_f16(const uint4& u, float* v8) {
  const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
  const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
  const float2 c = __half22float2(*reinterpret_cast<const __half2*>(&u.z));
  const float2 d = __half22float2(*reinterpret_cast<const __half2*>(&u.w));
  v8[0] = a.x; v8[1] = a.y; v8[2] = b.x; v8[3] = b.y; v8[4] = c.x; v8[5] = c.y; v8[6] = d.x; v8[7] = d.y;
}
__device__ __forceinline__ float4 unpack4_f16(const uint2& u) {
  const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
  const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
  return make_float4(a.x, a.y, b.x, b.y);
}
// thread = (row, 64-column half); j = 16-byte chunk (8 columns) inside the half.  Every operand tile is fp16.
__device__ __forceinline__ void row_store8(uint8_t* tile, int row, int half, int j, const float* v8) {
  *reinterpret_cast<uint4*>(tile + tc::sw128_chunk(row, half * 8 + j)) = tc::pack8_f16(v8);
}
__device__ __forceinline__ void row_load8(const uint8_t* tile, int row, int half, int j, float* v8) {
  unpack8_f16(*reinterpret_cast<const uint4*>(tile + tc::sw128_chunk(row, half * 8 + j)), v8);
}
__device__ __forceinline__ float tile_elem(const uint8_t* tile, int r, int c) {
  return __half2float(*reinterpret_cast<const __half*>(tile + tc::sw128_off(r, c)));
}
// column-thread partial sum over rows [hf*64, hf*64+64) of channel (tid & 127)
__device__ __forceinline__ float tile_colsum_f16(const uint8_t* tile) {
  const int ch = threadIdx.x & (H - 1), hf = threadIdx.x >> 7;
  float s = 0.f;
#pragma unroll 8
  for (int r = hf * 64; r < hf * 64 + 64; ++r) s += tile_elem(tile, r, ch);
  return s;
}
// fp32 staging tile [128][128] with the float4-chunk index XOR-swizzled by the row, so that both the
// row-per-thread writes and the column-per-thread reads are bank-conflict free
__device__ __forceinline__ float* s32_ptr(float* S, int r, int c) { return S + r * H + ((((c >> 2) ^ (r & 31)) << 2) | (c & 3)); }

// column sums of a bf16 tile: thread = (channel pair, 32-row quarter); adds into acc[2]
__device__ __forceinline__ void tile_colsum2_f16(const uint8_t* tile, float (&acc)[2]) {
  const int cp = threadIdx.x & 63, q = threadIdx.x >> 6;
  float s0 = 0.f, s1 = 0.f;
#pragma unroll 8
  for (int r = q * 32; r < q * 32 + 32; ++r) {
    const float2 v = __half22float2(*reinterpret_cast<const __half2*>(tile + tc::sw128_off(r, 2 * cp)));
    s0 += v.x;
    s1 += v.y;
  }
  acc[0] += s0;
  acc[1] += s1;
}
// flush (channel pair, quarter) partials: comb = [4][H] floats scratch
__device__ __forceinline__ void colpart2_flush(const float (&v)[2], float* comb, float* dst, bool add) {
  __syncthreads();
  const int cp = threadIdx.x & 63, q = threadIdx.x >> 6;
  comb[q * H + 2 * cp] = v[0];
  comb[q * H + 2 * cp + 1] = v[1];
  __syncthreads();
  if (threadIdx.x < H) {
    const float s = (comb[threadIdx.x] + comb[H + threadIdx.x]) + (comb[2 * H + threadIdx.x] + comb[3 * H + threadIdx.x]);
    dst[threadIdx.x] = add ? dst[threadIdx.x] + s : s;
  }
}



// ---- item-parallel receiver-segment sums of a bf16 tile --------------------------------------------------
// Work item = (receiver segment s, 16-byte chunk of 8 channels): the 16 lanes of a half-warp own one segment and
// read whole 256-byte rows with one 128-bit load each (4x fewer shared-memory instructions than a column walk,
// no per-row branch).  seg_row[0..nseg]: first row of every segment (seg_row[nseg] = nvalid); seg_cut[s]:
// 1 = segment wholly inside the tile (plain store), 2 = cut by a tile boundary (exactly two partial sums meet:
// atomicAdd onto a zeroed row, order-free).  Rows are added in row order => deterministic.  cs[8] accumulates
// the thread's column sums (chunk = tid & 15 is fixed per thread) for the bias gradient.  256 threads.
__device__ __forceinline__ void tile_segsum_items(const uint8_t* tile, const int* recv_s, const unsigned char* seg_row,
                                                  const unsigned char* seg_cut, int nseg, float* __restrict__ dst, float (&cs)[8]) {
  const int chunk = threadIdx.x & 15;
  for (int s = threadIdx.x >> 4; s < nseg; s += 16) {
    const int r0 = seg_row[s], r1 = seg_row[s + 1];
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int r = r0; r < r1; r += 2) {  // two rows in flight
      const uint4 u0 = *reinterpret_cast<const uint4*>(tile + tc::sw128_chunk(r, chunk));
      uint4 u1 = make_uint4(0u, 0u, 0u, 0u);
      if (r + 1 < r1) u1 = *reinterpret_cast<const uint4*>(tile + tc::sw128_chunk(r + 1, chunk));
      float v0[8], v1[8];
      unpack8_f16(u0, v0);
      unpack8_f16(u1, v1);
#pragma unroll
      for (int k = 0; k < 8; ++k) { acc[k] += v0[k]; acc[k] += v1[k]; }
    }
    float* d = dst + (size_t)recv_s[r0] * H + chunk * 8;
    if (seg_cut[s] == 1) {
      *reinterpret_cast<float4*>(d) = make_float4(acc[0], acc[1], acc[2], acc[3]);
      *reinterpret_cast<float4*>(d + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) atomicAdd(d + k, acc[k]);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) cs[k] += acc[k];
  }
}
// column sums of a bf16 tile, chunk-mapped: thread = (chunk = tid & 15, row group = tid >> 4), 8 rows each
__device__ __forceinline__ void tile_colsum_chunks(const uint8_t* tile, float (&cs)[8]) {
  const int chunk = threadIdx.x & 15, rg = threadIdx.x >> 4;
  uint4 u[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) u[i] = *reinterpret_cast<const uint4*>(tile + tc::sw128_chunk(rg + 16 * i, chunk));
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float v[8];
    unpack8_f16(u[i], v);
#pragma unroll
    for (int k = 0; k < 8; ++k) cs[k] += v[k];
  }
}

// TMEM accumulator (row-per-thread) -> swizzled fp32 staging tile
__device__ __forceinline__ void tmem_to_s32(uint32_t tacc, float* S32, int row, int half, uint32_t lane_base) {
#pragma unroll
  for (int hh = 0; hh < 2; ++hh) {
    float v[32];
    tc::tmem_ld32(tacc + lane_base + (uint32_t)(half * 64 + hh * 32), v);
    tc::tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; j += 4)
      *reinterpret_cast<float4*>(s32_ptr(S32, row, half * 64 + hh * 32 + j)) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
  }
}
// ---- weight-gradient flush: TMEM accumulator [128][128] += into the CTA's gradient slice ----------------
// TMEM -> fp32 staging in smem (thread = row x 64-column half) -> TWO 32 KB TMA reduce-adds (cp.reduce.async.bulk
// .add.f32): the copy engine performs the read-modify-write at L2, so the CTA issues no loads and does not wait for them.
// One 512-byte bulk operation per row (round 1) cost ~4.6 k cycles per accumulator -- per-operation overhead of the copy
// engine, 7 accumulators per message-passing step; whole 32 KB blocks need the SLICE to hold the accumulator in the
// staging order, so in the tensor-core path a [128][128] block of a weight gradient lives in its CTA slice as
//   element (r, c)  ->  block_base + grad_block_offset(r, c)            (float4 chunks XOR-swizzled by the row: the
//   row-per-thread staging stores are bank-conflict free), column blocks of a wider matrix ([128][256], [128][384])
//   one after the other (block b = columns [128 b, 128 b + 128)) -- same floats, same region of the slice, permuted.
// k_grad_reduce undoes the permutation while it sums the slices (pdg_backward.cu).  Each slice has one writer and
// launches are stream-ordered, so the sums stay bit-reproducible.
//   Sa / Sb: staging of rows 0-63 / 64-127 (32 KB each, 16-byte aligned, may be anywhere in shared memory).
//   Call order: acc_stage -> fence_async_smem + barrier (caller) -> acc_reduce_issue (ONE thread) -> the same thread
//   calls acc_reduce_drain before the staging is overwritten and tc::bulk_wait_all() before the kernel exits.
__host__ __device__ __forceinline__ int grad_block_offset(int r, int c) { return r * H + ((((c >> 2) ^ (r & 31)) << 2) | (c & 3)); }
__device__ __forceinline__ void acc_stage(uint32_t tacc, float* Sa, float* Sb, int row, int half, uint32_t lane_base) {
  float* base = row < 64 ? Sa + row * H : Sb + (row - 64) * H;
#pragma unroll
  for (int hh = 0; hh < 2; ++hh) {
    float v[32];
    tc::tmem_ld32(tacc + lane_base + (uint32_t)(half * 64 + hh * 32), v);
    tc::tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const int c = half * 64 + hh * 32 + j;
      *reinterpret_cast<float4*>(base + ((((c >> 2) ^ (row & 31)) << 2))) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    }
  }
}
// one thread; block = first float of the [128][128] block inside this CTA's slice
__device__ __forceinline__ void acc_reduce_issue(const float* Sa, const float* Sb, float* __restrict__ block) {
  tc::bulk_reduce_add_f32(block, Sa, 64 * H * 4);
  tc::bulk_reduce_add_f32(block + 64 * H, Sb, 64 * H * 4);
  tc::bulk_commit();
}
__device__ __forceinline__ void acc_reduce_drain() { tc::bulk_wait_read(); }  // issuing thread: staging may be overwritten afterwards
// __syncthreads flavour for the 256-thread kernels with ONE accumulator: S = 64 KB of contiguous staging
__device__ __forceinline__ void tmem_acc_flush(uint32_t tacc, float* S, float* __restrict__ block, int row, int half, uint32_t lane_base) {
  acc_stage(tacc, S, S + 64 * H, row, half, lane_base);
  tc::fence_async_smem();
  __syncthreads();
  if (threadIdx.x == 0) acc_reduce_issue(S, S + 64 * H, block);
}
// flush chunk-mapped column partials (16 row groups x columns {ch*4..+3, 64+ch*4..+3}); scr = [16][H] floats
__device__ __forceinline__ void chunkpart_flush(const float (&v)[8], float* scr, float* dst) {
  const int ch = threadIdx.x & 15, grp = threadIdx.x >> 4;
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    scr[grp * H + ch * 4 + j] = v[j];
    scr[grp * H + 64 + ch * 4 + j] = v[4 + j];
  }
  __syncthreads();
  if (threadIdx.x < H) {
    float s = 0.f;
#pragma unroll
    for (int g2 = 0; g2 < 16; ++g2) s += scr[g2 * H + threadIdx.x];
    dst[threadIdx.x] = s;
  }
}

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
  for (int j = 0; j < 32; j += 8) {  // 256-bit stores: half the L1 tag-stage cycles of a row-per-thread pattern
    const uint32_t* u = reinterpret_cast<const uint32_t*>(&v[j]);
    tc::stg256(dst_row_half + hh * 32 + j, make_uint4(u[0], u[1], u[2], u[3]), make_uint4(u[4], u[5], u[6], u[7]));
  }
}
// same values as fp16 (raw edge-MLP outputs are stored as fp16 rows in the tensor-core path)
__device__ __forceinline__ void row_store_global32_f16(__half* __restrict__ dst_row_half, const float (&v)[32], int hh) {
#pragma unroll
  for (int j = 0; j < 32; j += 16) tc::stg256(dst_row_half + hh * 32 + j, tc::pack8_f16(&v[j]), tc::pack8_f16(&v[j + 8]));
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

}  // namespace pdg
