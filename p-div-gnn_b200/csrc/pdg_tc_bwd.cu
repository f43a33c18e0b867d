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

#ifdef PDG_PHASE_TIMERS
__device__ unsigned long long g_phase[32];
#define PH(i) do { if (blockIdx.x == 0 && threadIdx.x == 0) { unsigned long long _t = clock64(); g_phase[i] += _t - _tl; _tl = _t; } } while (0)
#else
#define PH(i) do {} while (0)
#endif

constexpr int TC_SMEM_EDGE_BWD = 6 * tc::TILE_BF16_BYTES  // We, W2, E, HM, HN, DY
                                 + 2 * TM * 4             // recv / send
                                 + 3 * H * 4              // b1, b2, ln weight
                                 + 4 * H * 4              // column-sum combine scratch
                                 + 1024 + 2048;

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
  float* comb = lws + H;  // [4][H]
  float* smf = comb + 4 * H;
  int* qs = reinterpret_cast<int*>(smf + 4);          // [5] quarter boundaries (+3 pad)
  unsigned* masks = reinterpret_cast<unsigned*>(qs + 8);  // [4]
  unsigned char* code_s = reinterpret_cast<unsigned char*>(masks + 4);  // [128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(code_s + TM);  // [0] weights, [1..5] MMA groups
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
  float db2[2] = {0.f, 0.f}, db1[2] = {0.f, 0.f};  // (channel pair, row quarter) partials
  float cge8[8] = {0}, cgye8[8] = {0};               // chunk-mapped column partials: cols ch*4..+3 and 64+ch*4..+3
  uint32_t ph = 0;
  bool first = true;
  const uint32_t sE = tc::smem_u32(tE), sHM = tc::smem_u32(tHM), sHN = tc::smem_u32(tHN), sDY = tc::smem_u32(tDY);
  const uint32_t aWe = tc::smem_u32(sWe), aW2 = tc::smem_u32(sW2);

#ifdef PDG_PHASE_TIMERS
  unsigned long long _tl = clock64();
#endif
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
    // ---- E <- bf16(e_t): all 16 loads of the thread in flight before the first use ----
    {
      float4 ld[16];
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const size_t g = ((size_t)row0 + (tid >> 4) + it * 16) * H + ch * 8;
        ld[2 * it] = *reinterpret_cast<const float4*>(a.e_t + g);
        ld[2 * it + 1] = *reinterpret_cast<const float4*>(a.e_t + g + 4);
      }
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const float v[8] = {ld[2 * it].x, ld[2 * it].y, ld[2 * it].z, ld[2 * it].w,
                            ld[2 * it + 1].x, ld[2 * it + 1].y, ld[2 * it + 1].z, ld[2 * it + 1].w};
        *reinterpret_cast<uint4*>(tE + tc::sw128_chunk((tid >> 4) + it * 16, ch)) = tc::pack8_bf16(v);
      }
    }
    tc::fence_async_smem();
    __syncthreads();
    if (tid == 0) {
      if (first) tc::mbar_wait(&bars[0], 0);
      tc::fence_after_sync();
      tc::issue_gemm_kmajor(WORK0, sE, aWe, H, false);
      tc::mma_commit(&bars[1]);
    }
    tile_segment_codes(recv_s, a.rowptr, row0, nvalid, code_s, qs, masks);
    const int rc = recv_s[row], sd = send_s[row];
    tc::mbar_wait(&bars[1], ph);
    tc::fence_after_sync();
    // ---- hidden activations (same arithmetic as the forward) -> HM, HN ----
    {
      const __nv_bfloat16* par = reinterpret_cast<const __nv_bfloat16*>(a.Pa) + (size_t)rc * H + half * 64;
      const __nv_bfloat16* pbs = reinterpret_cast<const __nv_bfloat16*>(a.Pb) + (size_t)sd * H + half * 64;
      const __nv_bfloat16* pas = reinterpret_cast<const __nv_bfloat16*>(a.Pa) + (size_t)sd * H + half * 64;
      const __nv_bfloat16* pbr = reinterpret_cast<const __nv_bfloat16*>(a.Pb) + (size_t)rc * H + half * 64;
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        float gacc[32];
        tc::tmem_ld32(WORK0 + lane_base + (uint32_t)(half * 64 + hh * 32), gacc);
        tc::tmem_ld_wait();
#pragma unroll
        for (int c8 = 0; c8 < 4; ++c8) {
          const int co = hh * 32 + c8 * 8;
          float pr[8], ps[8], qs[8], qr[8];
          ldg8_bf16(par + co, pr);
          ldg8_bf16(pbs + co, ps);
          ldg8_bf16(pas + co, qs);
          ldg8_bf16(pbr + co, qr);
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
      const __nv_bfloat16* gp = reinterpret_cast<const __nv_bfloat16*>(a.gagg) + (size_t)rc * H + half * 64;
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        float v[32];
        tc::tmem_ld32(WORK1 + lane_base + (uint32_t)(half * 64 + hh * 32), v);
        tc::tmem_ld_wait();
#pragma unroll
        for (int c8 = 0; c8 < 4; ++c8) {
          const int co = hh * 32 + c8 * 8;
          float gg[8], d[8];
          ldg8_bf16(gp + co, gg);
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
    tile_colsum2_bf16(tDY, db2);  // overlaps the MMAs (reads only)
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
    tile_segsum2_bf16(tHM, recv_s, code_s, qs, a.RA);
    // ---- edge-update path ----
    if (!a.last) {
      // dy2 -> DY, elementwise in the coalesced loader mapping (16 lanes per row)
#pragma unroll 4
      for (int it = 0; it < 8; ++it) {
        const int r = (tid >> 4) + it * 16;
        const size_t g = ((size_t)row0 + r) * H + ch * 8;
        float d[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (r < nvalid) {
          float y[8], gg[8];
          *reinterpret_cast<float4*>(y) = *reinterpret_cast<const float4*>(a.y2_t + g);
          *reinterpret_cast<float4*>(y + 4) = *reinterpret_cast<const float4*>(a.y2_t + g + 4);
          *reinterpret_cast<float4*>(gg) = *reinterpret_cast<const float4*>(a.ge + g);
          *reinterpret_cast<float4*>(gg + 4) = *reinterpret_cast<const float4*>(a.ge + g + 4);
#pragma unroll
          for (int j = 0; j < 8; ++j) d[j] = y[j] > 0.f ? rstd2 * gg[j] * lws[ch * 8 + j] - c1n - c2n * (y[j] - mu2) : 0.f;
        }
        *reinterpret_cast<uint4*>(tDY + tc::sw128_chunk(r, ch)) = tc::pack8_bf16(d);
      }
      tc::fence_async_smem();
      __syncthreads();
      if (tid == 0) {
        tc::fence_after_sync();
        tc::issue_gemm_mnmajor(ACC_W2, sDY, sHN, true);  // dW2 += dy2^T hn
        tc::issue_gemm_k_mn(WORK1, sDY, aW2, false);     // dhn_pre = dy2 W2
        tc::mma_commit(&bars[4]);
      }
      tile_colsum2_bf16(tDY, db2);
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
      tile_segsum2_bf16(tHN, recv_s, code_s, qs, a.RB);
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
    tile_colsum2_bf16(tDY, db1);
    tc::mbar_wait(&bars[5], ph);
    tc::fence_after_sync();
    // ---- ge_t = ge_{t+1} + de ; column sums for the LayerNorm that produced e_t's increment ----
    // de leaves TMEM in the row-per-thread layout, is transposed through the fp32 staging tile and then
    // handled in the coalesced loader mapping (ge read-modify-write, y_prev read, column partials).
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      float v[32];
      tc::tmem_ld32(WORK0 + lane_base + (uint32_t)(half * 64 + hh * 32), v);
      tc::tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; j += 4)
        *reinterpret_cast<float4*>(s32_ptr(S32, row, half * 64 + hh * 32 + j)) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    }
    tc::fence_before_sync();
    __syncthreads();
#pragma unroll 4
    for (int it = 0; it < 8; ++it) {
      const int r = (tid >> 4) + it * 16;
      const size_t g = ((size_t)row0 + r) * H + ch * 4;  // columns ch*4..+3 and 64+ch*4..+3
      float4 d0 = *reinterpret_cast<const float4*>(s32_ptr(S32, r, ch * 4));
      float4 d1 = *reinterpret_cast<const float4*>(s32_ptr(S32, r, 64 + ch * 4));
      if (!a.last) {
        const float4 g0 = *reinterpret_cast<const float4*>(a.ge + g);
        const float4 g1 = *reinterpret_cast<const float4*>(a.ge + g + 64);
        d0.x += g0.x; d0.y += g0.y; d0.z += g0.z; d0.w += g0.w;
        d1.x += g1.x; d1.y += g1.y; d1.z += g1.z; d1.w += g1.w;
      }
      *reinterpret_cast<float4*>(a.ge + g) = d0;
      *reinterpret_cast<float4*>(a.ge + g + 64) = d1;
      const float4 y0 = *reinterpret_cast<const float4*>(a.yprev + g);
      const float4 y1 = *reinterpret_cast<const float4*>(a.yprev + g + 64);
      cge8[0] += d0.x; cge8[1] += d0.y; cge8[2] += d0.z; cge8[3] += d0.w;
      cge8[4] += d1.x; cge8[5] += d1.y; cge8[6] += d1.z; cge8[7] += d1.w;
      cgye8[0] = fmaf(d0.x, y0.x - mu_prev, cgye8[0]); cgye8[1] = fmaf(d0.y, y0.y - mu_prev, cgye8[1]);
      cgye8[2] = fmaf(d0.z, y0.z - mu_prev, cgye8[2]); cgye8[3] = fmaf(d0.w, y0.w - mu_prev, cgye8[3]);
      cgye8[4] = fmaf(d1.x, y1.x - mu_prev, cgye8[4]); cgye8[5] = fmaf(d1.y, y1.y - mu_prev, cgye8[5]);
      cgye8[6] = fmaf(d1.z, y1.z - mu_prev, cgye8[6]); cgye8[7] = fmaf(d1.w, y1.w - mu_prev, cgye8[7]);
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
  colpart2_flush(db2, comb, cg + param_offset(PE_B2), true);
  colpart2_flush(db1, comb, cg + param_offset(PE_B0), true);
  {  // chunk-mapped partials: 16 row groups (tid >> 4) x columns {ch*4..+3, 64+ch*4..+3}
    float* scr = reinterpret_cast<float*>(tE);  // [16][H], the tiles are dead now
    auto flush8 = [&](const float (&v)[8], float* dst) {
      __syncthreads();
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        scr[(tid >> 4) * H + ch * 4 + j] = v[j];
        scr[(tid >> 4) * H + 64 + ch * 4 + j] = v[4 + j];
      }
      __syncthreads();
      if (tid < H) {
        float s = 0.f;
#pragma unroll
        for (int g2 = 0; g2 < 16; ++g2) s += scr[g2 * H + tid];
        dst[tid] = s;
      }
    };
    flush8(cge8, a.cs2 + (size_t)blockIdx.x * 2 * H);
    flush8(cgye8, a.cs2 + (size_t)blockIdx.x * 2 * H + H);
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

#ifdef PDG_PHASE_TIMERS
extern "C" int pdg_phase_read(unsigned long long* out32) {
  return cudaMemcpyFromSymbol(out32, g_phase, sizeof(unsigned long long) * 32) == cudaSuccess ? 0 : -1;
}
#endif

int launch_edge_step_bwd_tc(const EdgeBwdArgs& a, const uint8_t* img, int grid, cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute((const void*)k_edge_step_bwd_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_EDGE_BWD);
  if (e != cudaSuccess) { set_error("k_edge_step_bwd_tc smem attribute: %s", cudaGetErrorString(e)); return -2; }
  k_edge_step_bwd_tc<<<grid, NT, TC_SMEM_EDGE_BWD, st>>>(a, img + (size_t)IMG_PE_WE * tc::TILE_BF16_BYTES,
                                                         img + (size_t)IMG_PE_W2 * tc::TILE_BF16_BYTES);
  return 0;
}

}  // namespace pdg
