// Self-test of the tcgen05 tile engine (pdg_tc.cuh): one CTA, one 128x128x128 GEMM.
//   mode 0: D = A . B^T          (both operands K-major)
//   mode 1: D = A^T . B          (both operands MN-major; weight-gradient shape)
//   mode 2: D = A . B            (A K-major, B MN-major; data-gradient shape)
// A, B: [128][128] fp32 row-major in global memory, rounded to the 16-bit operand formats of the real kernels on
// the way into the swizzled smem tiles (fp16 for both operands in every mode, see pdg_tc.cuh); B's tile additionally
// travels through a pre-swizzled global image + 1-D bulk copy (the route the weights take).  D: [128][128] fp32.
#include "pdg_common.cuh"
#include "pdg_tc.cuh"

namespace pdg {

__global__ void k_tc_make_image(const float* __restrict__ src, uint8_t* __restrict__ img) {
  // one thread per 16-byte chunk: 128 rows x 16 chunks
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 128 * 16) return;
  const int r = idx >> 4, ch = idx & 15;
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = src[r * 128 + ch * 8 + j];
  *reinterpret_cast<uint4*>(img + tc::sw128_chunk(r, ch)) = tc::pack8_f16(v);
}

__global__ void __launch_bounds__(256, 1)
k_tc_selftest(const float* __restrict__ A, const uint8_t* __restrict__ Bimg, float* __restrict__ D, int mode) {
  extern __shared__ __align__(1024) uint8_t sm[];
  uint8_t* tA = sm;
  uint8_t* tB = sm + tc::TILE_BF16_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + 2 * tc::TILE_BF16_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    tc::mbar_init(&bars[0], 1);
    tc::mbar_init(&bars[1], 1);
    tc::mbar_init_fence();
  }
  if (warp == 0) tc::tmem_alloc(tmem_slot, 128);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  if (tid == 0) {
    tc::mbar_expect_tx(&bars[0], tc::TILE_BF16_BYTES);
    tc::bulk_g2s(tB, Bimg, tc::TILE_BF16_BYTES, &bars[0]);
  }
  for (int idx = tid; idx < 128 * 16; idx += 256) {
    const int r = idx >> 4, ch = idx & 15;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = A[r * 128 + ch * 8 + j];
    *reinterpret_cast<uint4*>(tA + tc::sw128_chunk(r, ch)) = tc::pack8_f16(v);
  }
  tc::fence_async_smem();
  __syncthreads();
  if (tid == 0) {
    tc::mbar_wait(&bars[0], 0);
    tc::fence_after_sync();
    if (mode == 0) tc::issue_gemm_kmajor(tmem, tc::smem_u32(tA), tc::smem_u32(tB), 128, false);
    else if (mode == 1) tc::issue_gemm_mnmajor(tmem, tc::smem_u32(tA), tc::smem_u32(tB), false);
    else tc::issue_gemm_k_mn(tmem, tc::smem_u32(tA), tc::smem_u32(tB), false);
    tc::mma_commit(&bars[1]);
  }
  tc::mbar_wait(&bars[1], 0);
  tc::fence_after_sync();
  const int row = 32 * (warp & 3) + lane, c0 = 64 * (warp >> 2);
  float v[32];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    tc::tmem_ld32(tmem + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(c0 + 32 * h), v);
    tc::tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j) D[row * 128 + c0 + 32 * h + j] = v[j];
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 128);
}

}  // namespace pdg

using namespace pdg;
// A, B, D: device [128][128] fp32; img: device scratch of 32 KB
extern "C" int pdg_tc_selftest(int mode, const float* A, const float* B, float* D, void* img, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  const size_t smem = 2 * tc::TILE_BF16_BYTES + 1024;
  PDG_CUDA_CHECK(cudaFuncSetAttribute((const void*)k_tc_selftest, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_tc_make_image<<<8, 256, 0, st>>>(B, (uint8_t*)img);
  PDG_LAUNCH_CHECK();
  k_tc_selftest<<<1, 256, smem, st>>>(A, (const uint8_t*)img, D, mode);
  PDG_LAUNCH_CHECK();
  return 0;
}
