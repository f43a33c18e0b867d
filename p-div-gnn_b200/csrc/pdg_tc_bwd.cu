// bf16 tensor-core (tcgen05 / TMEM) version of the backward edge kernel (PDG_PREC_BF16).
//
// Same math as k_edge_step_bwd (pdg_backward.cu).  Per 128-edge tile, eight 128^3 GEMMs:
//   G   = E  . We^T          (K/K)    work0      recompute of the shared layer-1 term
//   y1  = HM . W2^T          (K/K)    work1      recompute of the message pre-activation
//   dW2 += DY^T . HM         (MN/MN)  acc[0]     persistent TMEM accumulator (whole kernel)
//   dhm = DY . W2            (K/MN)   work0
//   dW2 += DY^T . HN ; dhn = DY . W2  work1      (edge-update path, skipped on the last step)
//   de  = DG . We            (K/MN)   work0
//   dWe += DG^T . E          (MN/MN)  acc[1]     persistent TMEM accumulator
// Tiles E, HM, HN, DY are bf16 SWIZZLE_128B images that serve as K-major operands of the
// forward/dgrad GEMMs and, unchanged, as MN-major operands of the weight-gradient GEMMs.
// The two weight-gradient accumulators live in TMEM for the whole kernel and are added to
// the CTA's gradient slice once at the end (the FFMA path pays a 64 KB RMW per tile).
// dhm / dhn leave as bf16 rows (half the sender-gather traffic of the fp32 path).
#include "pdg_ws.cuh"
#include "pdg_tc_tile.cuh"

namespace pdg {

constexpr int TC_SMEM_EDGE_BWD = 6 * tc::TILE_BF16_BYTES  // We, W2, E, HM, HN, DY
                                 + 2 * TM * 4             // recv / send
                                 + 3 * H * 4              // b1, b2, ln weight
                                 + 2 * H * 4              // column-sum combine scratch
                                 + 512 + 2048;

__global__ void __launch_bounds__(NT, 1)
k_edge_step_bwd_tc(EdgeBwdArgs a, const uint8_t* __restrict__ imgWe, const uint8_t* __restrict__ imgW2) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sm = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sWe = sm;
  uint8_t* sW2 = sWe + tc::TILE_BF16_BYTES;
  uint8_t* tE = sW2 + tc::TILE_BF16_BYTES;
  uint8_t* tHM = tE + tc::TILE_BF16_BYTES;
  uint8_t* tHN = tHM + tc::TILE_BF16_BYTES;
  uint8_t* tDY = tHN + tc::TILE_BF16_BYTES;
  float* S32 = reinterpret_cast<float*>(tHM);  // fp32 staging, aliases HM + HN (64 KB) once both are dead
  int* recv_s = reinterpret_cast<int*>(tDY + tc::TILE_BF16_BYTES);
  int* send_s = recv_s + TM;
  float* b1s = reinterpret_cast<float*>(send_s + TM);
  float* b2s = b1s + H;
  float* lws = b2s + H;
  float* comb = lws + H;  // [2][H]
  float* smf = comb + 2 * H;
  int* smi = reinterpret_cast<int*>(smf + 4);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smi + 4);  // [0] weights, [1..5] MMA groups
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row = 32 * (warp & 3) + lane, half = warp >> 2;
  const uint32_t lane_base = (uint32_t)(32 * (warp & 3)) << 16;
  float* cg = a.cta_grads + (size_t)blockIdx.x * GRADP;

  if (tid == 0) {
    for (int i = 0; i < 6; ++i) tc::mbar_init(&bars[i], 1);
    tc::mbar_init_fence();
  }
  if (warp == 0) tc::tmem_alloc(tmem_slot, 512);
  if (tid < H) { b1s[tid] = a.b1[tid]; b2s[tid] = a.b2[tid]; lws[tid] = a.lnw[tid]; }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  const uint32_t ACC_W2 = tmem, ACC_WE = tmem + 128, WORK0 = tmem + 256, WORK1 = tmem + 384;
  if (tid == 0) {
    tc::mbar_expect_tx(&bars[0], 2 * tc::TILE_BF16_BYTES);
    tc::bulk_g2s(sWe, imgWe, tc::TILE_BF16_BYTES, &bars[0]);
    tc::bulk_g2s(sW2, imgW2, tc::TILE_BF16_BYTES, &bars[0]);
  }
  const float mu_prev = ln_stat_block(a.parts_prev, a.count_prev, smf).mu;
  const float c1m = a.scal1[0], c2m = a.scal1[1], mu1 = a.scal1[2], rstd1 = a.scal1[3];
  float c1n = 0.f, c2n = 0.f, mu2 = 0.f, rstd2 = 0.f;
  if (!a.last) { c1n = a.scal2[0]; c2n = a.scal2[1]; mu2 = a.scal2[2]; rstd2 = a.scal2[3]; }
  const int ch = tid & 15;
  float db2 = 0.f, db1 = 0.f, cge = 0.f, cgye = 0.f;  // column-thread partials (channel tid&127, row half tid>>7)
  uint32_t ph = 0;
  bool first = true;
  const uint32_t sE = tc::smem_u32(tE), sHM = tc::smem_u32(tHM), sHN = tc::smem_u32(tHN), sDY = tc::smem_u32(tDY);
  const uint32_t aWe = tc::smem_u32(sWe), aW2 = tc::smem_u32(sW2);

  for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
    const int row0 = tile * TM;
    const int nvalid = min(TM, a.E - row0);
    const bool ok = row < nvalid;
    const size_t grow = ((size_t)row0 + row) * H + half * 64;  // this thread's 64 floats of a global fp32 row
    if (tid < TM) {
      recv_s[tid] = a.recv[row0 + tid];
      send_s[tid] = a.send[row0 + tid];
    }
    {  // pull the next tile's streamed rows into L2 while this one computes
      const int nt = tile + gridDim.x;
      if (nt < a.n_tiles) {
        const size_t pg = ((size_t)nt * TM + row) * H + half * 64;
        tc::prefetch_l2(a.e_t + pg);
        tc::prefetch_l2(a.e_t + pg + 32);
        tc::prefetch_l2(a.yprev + pg);
        tc::prefetch_l2(a.yprev + pg + 32);
        if (!a.last) {
          tc::prefetch_l2(a.y2_t + pg);
          tc::prefetch_l2(a.y2_t + pg + 32);
          tc::prefetch_l2(a.ge + pg);
          tc::prefetch_l2(a.ge + pg + 32);
        }
      }
    }
    // ---- E <- bf16(e_t) ----
#pragma unroll 2
    for (int it = 0; it < 8; ++it) {
      const int r = (tid >> 4) + it * 16;
      const size_t g = ((size_t)row0 + r) * H + ch * 8;
      float v[8];
      *reinterpret_cast<float4*>(v) = *reinterpret_cast<const float4*>(a.e_t + g);
      *reinterpret_cast<float4*>(v + 4) = *reinterpret_cast<const float4*>(a.e_t + g + 4);
      *reinterpret_cast<uint4*>(tE + tc::sw128_chunk(r, ch)) = tc::pack8_bf16(v);
    }
    tc::fence_async_smem();
    __syncthreads();
    if (tid == 0) {
      if (first) tc::mbar_wait(&bars[0], 0);
      tc::fence_after_sync();
      tc::issue_gemm_kmajor(WORK0, sE, aWe, H, false);
      tc::mma_commit(&bars[1]);
    }
    if (tid == 32) {
      int sp = nvalid;
      if (nvalid > 64)
        for (int r = 64; r < nvalid; ++r)
          if (recv_s[r] != recv_s[r - 1]) { sp = r; break; }
      smi[0] = sp;
    }
    const int rc = recv_s[row], sd = send_s[row];
    tc::mbar_wait(&bars[1], ph);
    tc::fence_after_sync();
    // ---- hidden activations (same arithmetic as the forward) -> HM, HN ----
    {
      const float* par = a.Pa + (size_t)rc * H + half * 64;
      const float* pbs = a.Pb + (size_t)sd * H + half * 64;
      const float* pas = a.Pa + (size_t)sd * H + half * 64;
      const float* pbr = a.Pb + (size_t)rc * H + half * 64;
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        float gacc[32];
        tc::tmem_ld32(WORK0 + lane_base + (uint32_t)(half * 64 + hh * 32), gacc);
        tc::tmem_ld_wait();
#pragma unroll
        for (int c8 = 0; c8 < 4; ++c8) {
          const int co = hh * 32 + c8 * 8;
          float pr[8], ps[8], qs[8], qr[8];
          *reinterpret_cast<float4*>(pr) = __ldg(reinterpret_cast<const float4*>(par + co));
          *reinterpret_cast<float4*>(pr + 4) = __ldg(reinterpret_cast<const float4*>(par + co + 4));
          *reinterpret_cast<float4*>(ps) = __ldg(reinterpret_cast<const float4*>(pbs + co));
          *reinterpret_cast<float4*>(ps + 4) = __ldg(reinterpret_cast<const float4*>(pbs + co + 4));
          *reinterpret_cast<float4*>(qs) = __ldg(reinterpret_cast<const float4*>(pas + co));
          *reinterpret_cast<float4*>(qs + 4) = __ldg(reinterpret_cast<const float4*>(pas + co + 4));
          *reinterpret_cast<float4*>(qr) = __ldg(reinterpret_cast<const float4*>(pbr + co));
          *reinterpret_cast<float4*>(qr + 4) = __ldg(reinterpret_cast<const float4*>(pbr + co + 4));
          float hm[8], hn[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float g = gacc[c8 * 8 + j] + b1s[half * 64 + co + j];
            hm[j] = fmaxf(g + pr[j] + ps[j], 0.f);
            hn[j] = fmaxf(g + qs[j] + qr[j], 0.f);
          }
          row_store8(tHM, row, half, hh * 4 + c8, hm);
          row_store8(tHN, row, half, hh * 4 + c8, hn);
        }
      }
    }
    tc::fence_before_sync();
    tc::fence_async_smem();
    __syncthreads();
    if (tid == 0) {
      tc::fence_after_sync();
      tc::issue_gemm_kmajor(WORK1, sHM, aW2, H, false);  // y1 pre-activation
      tc::mma_commit(&bars[2]);
    }
    tc::mbar_wait(&bars[2], ph);
    tc::fence_after_sync();
    // ---- dy1 -> DY ----
    {
      const float* gp = a.gagg + (size_t)rc * H + half * 64;
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        float v[32];
        tc::tmem_ld32(WORK1 + lane_base + (uint32_t)(half * 64 + hh * 32), v);
        tc::tmem_ld_wait();
#pragma unroll
        for (int c8 = 0; c8 < 4; ++c8) {
          const int co = hh * 32 + c8 * 8;
          float gg[8], d[8];
          *reinterpret_cast<float4*>(gg) = __ldg(reinterpret_cast<const float4*>(gp + co));
          *reinterpret_cast<float4*>(gg + 4) = __ldg(reinterpret_cast<const float4*>(gp + co + 4));
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int c = half * 64 + co + j;
            const float y = fmaxf(v[c8 * 8 + j] + b2s[c], 0.f);
            d[j] = (ok && y > 0.f) ? rstd1 * gg[j] * lws[c] - c1m - c2m * (y - mu1) : 0.f;
          }
          row_store8(tDY, row, half, hh * 4 + c8, d);
        }
      }
    }
    tc::fence_before_sync();
    tc::fence_async_smem();
    __syncthreads();
    if (tid == 0) {
      tc::fence_after_sync();
      tc::issue_gemm_mnmajor(ACC_W2, sDY, sHM, !first);  // dW2 += dy1^T hm
      tc::issue_gemm_k_mn(WORK0, sDY, aW2, false);       // dhm_pre = dy1 W2
      tc::mma_commit(&bars[3]);
    }
    db2 += tile_colsum_bf16(tDY);  // overlaps the MMAs (reads only)
    tc::mbar_wait(&bars[3], ph);
    tc::fence_after_sync();
    // ---- dhm = dhm_pre * [hm > 0] -> HM tile (bf16) + DHM rows (bf16) ----
    {
      __nv_bfloat16* dh = reinterpret_cast<__nv_bfloat16*>(a.DHM) + grow;
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        float v[32];
        tc::tmem_ld32(WORK0 + lane_base + (uint32_t)(half * 64 + hh * 32), v);
        tc::tmem_ld_wait();
#pragma unroll
        for (int c8 = 0; c8 < 4; ++c8) {
          float h[8], d[8];
          row_load8(tHM, row, half, hh * 4 + c8, h);
#pragma unroll
          for (int j = 0; j < 8; ++j) d[j] = h[j] > 0.f ? v[c8 * 8 + j] : 0.f;
          const uint4 pk = tc::pack8_bf16(d);
          *reinterpret_cast<uint4*>(tHM + tc::sw128_chunk(row, half * 8 + hh * 4 + c8)) = pk;
          *reinterpret_cast<uint4*>(dh + hh * 32 + c8 * 8) = pk;
        }
      }
    }
    tc::fence_before_sync();
    __syncthreads();
    tile_segsum_bf16(tHM, recv_s, a.rowptr, row0, nvalid, smi[0], a.RA);
    // ---- edge-update path ----
    if (!a.last) {
      {
        const float* yp = a.y2_t + grow;
        const float* gp = a.ge + grow;
#pragma unroll
        for (int c8 = 0; c8 < 8; ++c8) {
          float y[8], g[8], d[8];
          *reinterpret_cast<float4*>(y) = *reinterpret_cast<const float4*>(yp + c8 * 8);
          *reinterpret_cast<float4*>(y + 4) = *reinterpret_cast<const float4*>(yp + c8 * 8 + 4);
          *reinterpret_cast<float4*>(g) = *reinterpret_cast<const float4*>(gp + c8 * 8);
          *reinterpret_cast<float4*>(g + 4) = *reinterpret_cast<const float4*>(gp + c8 * 8 + 4);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int c = half * 64 + c8 * 8 + j;
            d[j] = (ok && y[j] > 0.f) ? rstd2 * g[j] * lws[c] - c1n - c2n * (y[j] - mu2) : 0.f;
          }
          row_store8(tDY, row, half, c8, d);
        }
      }
      tc::fence_async_smem();
      __syncthreads();
      if (tid == 0) {
        tc::fence_after_sync();
        tc::issue_gemm_mnmajor(ACC_W2, sDY, sHN, true);  // dW2 += dy2^T hn
        tc::issue_gemm_k_mn(WORK1, sDY, aW2, false);     // dhn_pre = dy2 W2
        tc::mma_commit(&bars[4]);
      }
      db2 += tile_colsum_bf16(tDY);
      tc::mbar_wait(&bars[4], ph);
      tc::fence_after_sync();
      {
        __nv_bfloat16* dh = reinterpret_cast<__nv_bfloat16*>(a.DHN) + grow;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          float v[32];
          tc::tmem_ld32(WORK1 + lane_base + (uint32_t)(half * 64 + hh * 32), v);
          tc::tmem_ld_wait();
#pragma unroll
          for (int c8 = 0; c8 < 4; ++c8) {
            float h[8], d[8];
            row_load8(tHN, row, half, hh * 4 + c8, h);
#pragma unroll
            for (int j = 0; j < 8; ++j) d[j] = h[j] > 0.f ? v[c8 * 8 + j] : 0.f;
            const uint4 pk = tc::pack8_bf16(d);
            *reinterpret_cast<uint4*>(tHN + tc::sw128_chunk(row, half * 8 + hh * 4 + c8)) = pk;
            *reinterpret_cast<uint4*>(dh + hh * 32 + c8 * 8) = pk;
          }
        }
      }
      tc::fence_before_sync();
      __syncthreads();
      tile_segsum_bf16(tHN, recv_s, a.rowptr, row0, nvalid, smi[0], a.RB);
    }
    // ---- dG = dhm + dhn -> DY ----
#pragma unroll
    for (int c8 = 0; c8 < 8; ++c8) {
      float m[8];
      row_load8(tHM, row, half, c8, m);
      if (!a.last) {
        float n8[8];
        row_load8(tHN, row, half, c8, n8);
#pragma unroll
        for (int j = 0; j < 8; ++j) m[j] += n8[j];
      }
      row_store8(tDY, row, half, c8, m);
    }
    tc::fence_async_smem();
    __syncthreads();
    if (tid == 0) {
      tc::fence_after_sync();
      tc::issue_gemm_k_mn(WORK0, sDY, aWe, false);       // de = dG We
      tc::issue_gemm_mnmajor(ACC_WE, sDY, sE, !first);   // dWe += dG^T e_t
      tc::mma_commit(&bars[5]);
    }
    db1 += tile_colsum_bf16(tDY);
    tc::mbar_wait(&bars[5], ph);
    tc::fence_after_sync();
    // ---- ge_t = ge_{t+1} + de ; column sums for the LayerNorm that produced e_t's increment ----
    {
      float* gp = a.ge + grow;
      const float* yp = a.yprev + grow;
      float de[64];
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        float v[32];
        tc::tmem_ld32(WORK0 + lane_base + (uint32_t)(half * 64 + hh * 32), v);
        tc::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          float4 o = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          if (!a.last) {
            const float4 g = *reinterpret_cast<const float4*>(gp + hh * 32 + j);
            o.x += g.x; o.y += g.y; o.z += g.z; o.w += g.w;
          }
          *reinterpret_cast<float4*>(gp + hh * 32 + j) = o;
          de[hh * 32 + j] = o.x; de[hh * 32 + j + 1] = o.y; de[hh * 32 + j + 2] = o.z; de[hh * 32 + j + 3] = o.w;
        }
      }
      // pass 1: colsum(de)
#pragma unroll
      for (int j = 0; j < 64; j += 4)
        *reinterpret_cast<float4*>(s32_ptr(S32, row, half * 64 + j)) = make_float4(de[j], de[j + 1], de[j + 2], de[j + 3]);
      __syncthreads();
      {
        const int chn = tid & (H - 1), hf = tid >> 7;
        float s = 0.f;
#pragma unroll 8
        for (int r = hf * 64; r < hf * 64 + 64; ++r) s += *s32_ptr(S32, r, chn);
        cge += s;
      }
      __syncthreads();
      // pass 2: colsum(de * (y_prev - mu))
#pragma unroll
      for (int j = 0; j < 64; j += 4) {
        const float4 y = *reinterpret_cast<const float4*>(yp + j);
        *reinterpret_cast<float4*>(s32_ptr(S32, row, half * 64 + j)) =
            make_float4(de[j] * (y.x - mu_prev), de[j + 1] * (y.y - mu_prev), de[j + 2] * (y.z - mu_prev),
                        de[j + 3] * (y.w - mu_prev));
      }
      __syncthreads();
      {
        const int chn = tid & (H - 1), hf = tid >> 7;
        float s = 0.f;
#pragma unroll 8
        for (int r = hf * 64; r < hf * 64 + 64; ++r) s += *s32_ptr(S32, r, chn);
        cgye += s;
      }
    }
    ph ^= 1u;
    first = false;
    tc::fence_before_sync();
    __syncthreads();
  }
  // ---- flush: TMEM weight-gradient accumulators -> this CTA's gradient slice ----
  {
    float* w2 = cg + param_offset(PE_W2) + (size_t)row * H + half * 64;
    float* we = cg + param_offset(PE_W0) + (size_t)row * 3 * H + 2 * H + half * 64;
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      float v[32];
      tc::tmem_ld32(ACC_W2 + lane_base + (uint32_t)(half * 64 + hh * 32), v);
      tc::tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) w2[hh * 32 + j] += v[j];
      tc::tmem_ld32(ACC_WE + lane_base + (uint32_t)(half * 64 + hh * 32), v);
      tc::tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) we[hh * 32 + j] += v[j];
    }
  }
  // column-thread partials: combine the two row halves, then add / store
  auto flush = [&](float v, float* dst, bool add) {
    __syncthreads();
    comb[(tid >> 7) * H + (tid & (H - 1))] = v;
    __syncthreads();
    if (tid < H) {
      const float s = comb[tid] + comb[H + tid];
      dst[tid] = add ? dst[tid] + s : s;
    }
  };
  flush(db2, cg + param_offset(PE_B2), true);
  flush(db1, cg + param_offset(PE_B0), true);
  flush(cge, a.cs2 + (size_t)blockIdx.x * 2 * H, false);
  flush(cgye, a.cs2 + (size_t)blockIdx.x * 2 * H + H, false);
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

int launch_edge_step_bwd_tc(const EdgeBwdArgs& a, const uint8_t* img, int grid, cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute((const void*)k_edge_step_bwd_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_EDGE_BWD);
  if (e != cudaSuccess) { set_error("k_edge_step_bwd_tc smem attribute: %s", cudaGetErrorString(e)); return -2; }
  k_edge_step_bwd_tc<<<grid, NT, TC_SMEM_EDGE_BWD, st>>>(a, img + (size_t)IMG_PE_WE * tc::TILE_BF16_BYTES,
                                                         img + (size_t)IMG_PE_W2 * tc::TILE_BF16_BYTES);
  return 0;
}

}  // namespace pdg
