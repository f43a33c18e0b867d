// bf16 tensor-core (tcgen05 / TMEM) version of the forward edge kernel (PDG_PREC_BF16).
//
// Same math and data flow as k_edge_step (pdg_forward.cu; reference models.py:215-238), but
// the three 128x128x128 GEMMs of a tile run on the 5th-gen tensor cores:
//   operands  bf16 in SWIZZLE_128B shared-memory tiles (activations written by the fused
//             prologue/epilogues, weights staged once per CTA by 1-D TMA bulk copies of
//             pre-swizzled images), fp32 accumulation in TMEM (3 x 128 columns),
//   epilogues read TMEM with tcgen05.ld (thread = one edge row x 64 channels), add bias /
//             gathered node projections / ReLU in fp32, and either re-quantise to bf16 for
//             the next GEMM or leave through fp32 (LayerNorm statistics, segment sums, HBM).
// Latent storage, LayerNorm and all reductions stay fp32; tolerance of this mode: 2e-2.
#include "pdg_ws.cuh"
#include "pdg_tc_tile.cuh"

namespace pdg {

constexpr int TC_SMEM_EDGE = 2 * tc::TILE_BF16_BYTES      // weight images We, W2
                             + 2 * tc::TILE_BF16_BYTES    // A0 (e_t / hn), A1 (hm)
                             + TM * LDS * 4               // fp32 staging of y1 for the segment sum
                             + 2 * TM * 4                 // recv / send
                             + 2 * H * 4                  // b1, b2
                             + 1024 + 2048;               // scalars, segment codes, barriers, alignment slack

__global__ void __launch_bounds__(NT, 1)
k_edge_step_tc(EdgeStepArgs a, const uint8_t* __restrict__ imgWe, const uint8_t* __restrict__ imgW2) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sm = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sWe = sm;
  uint8_t* sW2 = sWe + tc::TILE_BF16_BYTES;
  uint8_t* A0 = sW2 + tc::TILE_BF16_BYTES;
  uint8_t* A1 = A0 + tc::TILE_BF16_BYTES;
  float* S = reinterpret_cast<float*>(A1 + tc::TILE_BF16_BYTES);
  int* recv_s = reinterpret_cast<int*>(S + TM * LDS);
  int* send_s = recv_s + TM;
  float* b1s = reinterpret_cast<float*>(send_s + TM);
  float* b2s = b1s + H;
  double* red = reinterpret_cast<double*>(b2s + H);
  float* smf = reinterpret_cast<float*>(red + 16);
  int* qs = reinterpret_cast<int*>(smf + 4);  // [5] (+3 pad)
  unsigned* masks = reinterpret_cast<unsigned*>(qs + 8);
  unsigned char* code_s = reinterpret_cast<unsigned char*>(masks + 4);  // [128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(code_s + TM);  // [0] weights, [1..3] accumulators 0..2
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row = 32 * (warp & 3) + lane, half = warp >> 2;
  const uint32_t lane_base = (uint32_t)(32 * (warp & 3)) << 16;

  if (tid == 0) {
    tc::mbar_init(&bars[0], 1);
    tc::mbar_init(&bars[1], 1);
    tc::mbar_init(&bars[2], 1);
    tc::mbar_init(&bars[3], 1);
    tc::mbar_init_fence();
  }
  if (warp == 0) tc::tmem_alloc(tmem_slot, 512);
  if (tid < H) { b1s[tid] = a.b1[tid]; b2s[tid] = a.b2[tid]; }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  if (tid == 0) {
    tc::mbar_expect_tx(&bars[0], 2 * tc::TILE_BF16_BYTES);
    tc::bulk_g2s(sWe, imgWe, tc::TILE_BF16_BYTES, &bars[0]);
    tc::bulk_g2s(sW2, imgW2, tc::TILE_BF16_BYTES, &bars[0]);
  }
  const LnStat st = ln_stat_block(a.prev_parts, a.prev_count, smf);
  // loader mapping: 16 lanes cover one row (16 chunks of 8 floats), 16 rows per pass, 8 passes
  const int ch = tid & 15;
  float lw[8], lb[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { lw[j] = a.prev_w[ch * 8 + j]; lb[j] = a.prev_b[ch * 8 + j]; }
  double t1s = 0, t1ss = 0, t2s = 0, t2ss = 0;
  uint32_t ph = 0;
  bool weights_ready = false;
  for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
    const int row0 = tile * TM;
    const int nvalid = min(TM, a.E - row0);
    if (tid < TM) {
      recv_s[tid] = a.recv[row0 + tid];
      send_s[tid] = a.send[row0 + tid];
    }
    {  // pull the next tile's rows into L2 while this one computes (thread = row x 256-byte half)
      const int nt = tile + gridDim.x;
      if (nt < a.n_tiles) {
        const size_t pg = ((size_t)nt * TM + row) * H + half * 64;
        tc::prefetch_l2(a.yprev + pg);
        tc::prefetch_l2(a.yprev + pg + 32);
        if (a.base != nullptr) { tc::prefetch_l2(a.base + pg); tc::prefetch_l2(a.base + pg + 32); }
      }
    }
    // ---- e_t tile: lazy LayerNorm + residual, fp32 to HBM, bf16 to the A0 operand tile ----
#pragma unroll 4
    for (int it = 0; it < 8; ++it) {
      const int r = (tid >> 4) + it * 16;
      const size_t g = ((size_t)row0 + r) * H + ch * 8;
      const float4 y0 = *reinterpret_cast<const float4*>(a.yprev + g);
      const float4 y1 = *reinterpret_cast<const float4*>(a.yprev + g + 4);
      float v[8] = {y0.x, y0.y, y0.z, y0.w, y1.x, y1.y, y1.z, y1.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = (v[j] - st.mu) * st.rstd * lw[j] + lb[j];
      if (a.base != nullptr) {
        const float4 x0 = *reinterpret_cast<const float4*>(a.base + g);
        const float4 x1 = *reinterpret_cast<const float4*>(a.base + g + 4);
        v[0] += x0.x; v[1] += x0.y; v[2] += x0.z; v[3] += x0.w;
        v[4] += x1.x; v[5] += x1.y; v[6] += x1.z; v[7] += x1.w;
      }
      if (a.e_out != nullptr) {
        *reinterpret_cast<float4*>(a.e_out + g) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(a.e_out + g + 4) = make_float4(v[4], v[5], v[6], v[7]);
      }
      *reinterpret_cast<uint4*>(A0 + tc::sw128_chunk(r, ch)) = tc::pack8_bf16(v);
    }
    tc::fence_async_smem();
    __syncthreads();
    if (tid == 0) {
      if (!weights_ready) tc::mbar_wait(&bars[0], 0);
      tc::fence_after_sync();
      tc::issue_gemm_kmajor(tmem, tc::smem_u32(A0), tc::smem_u32(sWe), H, false);  // G = e_t We^T
      tc::mma_commit(&bars[1]);
    }
    weights_ready = true;
    tile_segment_codes(recv_s, a.rowptr, row0, nvalid, code_s, qs, masks);
    tc::mbar_wait(&bars[1], ph);
    tc::fence_after_sync();
    // ---- hidden activations of both edge-MLP evaluations -> A1 (message), A0 (edge update) ----
    {
      const int rc = recv_s[row], sd = send_s[row];
      const __nv_bfloat16* par = reinterpret_cast<const __nv_bfloat16*>(a.Pa) + (size_t)rc * H + half * 64;
      const __nv_bfloat16* pbs = reinterpret_cast<const __nv_bfloat16*>(a.Pb) + (size_t)sd * H + half * 64;
      const __nv_bfloat16* pas = reinterpret_cast<const __nv_bfloat16*>(a.Pa) + (size_t)sd * H + half * 64;
      const __nv_bfloat16* pbr = reinterpret_cast<const __nv_bfloat16*>(a.Pb) + (size_t)rc * H + half * 64;
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        float gacc[32];
        tc::tmem_ld32(tmem + lane_base + (uint32_t)(half * 64 + hh * 32), gacc);
        tc::tmem_ld_wait();
#pragma unroll
        for (int c8 = 0; c8 < 4; ++c8) {
          const int co = hh * 32 + c8 * 8;  // column offset inside this thread's 64
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
          const int chunk = half * 8 + hh * 4 + c8;
          *reinterpret_cast<uint4*>(A1 + tc::sw128_chunk(row, chunk)) = tc::pack8_bf16(hm);
          *reinterpret_cast<uint4*>(A0 + tc::sw128_chunk(row, chunk)) = tc::pack8_bf16(hn);
        }
      }
    }
    tc::fence_before_sync();
    tc::fence_async_smem();
    __syncthreads();
    if (tid == 0) {
      tc::fence_after_sync();
      tc::issue_gemm_kmajor(tmem + 128, tc::smem_u32(A1), tc::smem_u32(sW2), H, false);  // message layer 2
      tc::mma_commit(&bars[2]);
      if (a.y2_out != nullptr) {
        tc::issue_gemm_kmajor(tmem + 256, tc::smem_u32(A0), tc::smem_u32(sW2), H, false);  // edge-update layer 2
        tc::mma_commit(&bars[3]);
      }
    }
    // ---- message: y1 = relu(acc1 + b2) -> fp32 staging -> receiver-segment sums + LN1 partials ----
    tc::mbar_wait(&bars[2], ph);
    tc::fence_after_sync();
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      float v[32];
      tc::tmem_ld32(tmem + 128 + lane_base + (uint32_t)(half * 64 + hh * 32), v);
      tc::tmem_ld_wait();
      float* dst = S + row * LDS + half * 64 + hh * 32;
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const int c = half * 64 + hh * 32 + j;
        *reinterpret_cast<float4*>(dst + j) =
            make_float4(fmaxf(v[j] + b2s[c], 0.f), fmaxf(v[j + 1] + b2s[c + 1], 0.f), fmaxf(v[j + 2] + b2s[c + 2], 0.f),
                        fmaxf(v[j + 3] + b2s[c + 3], 0.f));
      }
    }
    __syncthreads();
    {
      // receiver-segment sums: thread = (channel pair, row quarter); rows walked in order => fixed summation order
      const int cp = tid & 63, q = tid >> 6;
      const int r0 = qs[q], r1 = qs[q + 1];
      float g0 = 0.f, g1 = 0.f, s = 0.f, ss = 0.f;
      for (int r = r0; r < r1; ++r) {
        const float2 v = *reinterpret_cast<const float2*>(S + r * LDS + 2 * cp);
        g0 += v.x;
        g1 += v.y;
        s += v.x + v.y;
        ss = fmaf(v.x, v.x, fmaf(v.y, v.y, ss));
        const int code = code_s[r];
        if (code) {
          float* dst = a.aggraw + (size_t)recv_s[r] * H + 2 * cp;
          if (code == 1) *reinterpret_cast<float2*>(dst) = make_float2(g0, g1);  // whole segment seen here
          else { atomicAdd(dst, g0); atomicAdd(dst + 1, g1); }  // cut by a tile boundary: two addends, order-free
          g0 = 0.f; g1 = 0.f;
        }
      }
      double ds = s, dss = ss;
      block_sum2(ds, dss, red);
      if (tid == 0) { t1s += ds; t1ss += dss; }
    }
    // ---- edge update: y2 = relu(acc2 + b2) raw to HBM + LN2 partials ----
    if (a.y2_out != nullptr) {
      tc::mbar_wait(&bars[3], ph);
      tc::fence_after_sync();
      float s = 0.f, ss = 0.f;
      const bool ok = row < nvalid;
      float* dst = S + row * LDS + half * 64;  // block_sum2 above already fenced the segment-sum reads of S
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        float v[32];
        tc::tmem_ld32(tmem + 256 + lane_base + (uint32_t)(half * 64 + hh * 32), v);
        tc::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const int c = half * 64 + hh * 32 + j;
          float4 o = make_float4(fmaxf(v[j] + b2s[c], 0.f), fmaxf(v[j + 1] + b2s[c + 1], 0.f),
                                 fmaxf(v[j + 2] + b2s[c + 2], 0.f), fmaxf(v[j + 3] + b2s[c + 3], 0.f));
          if (ok) {
            s += (o.x + o.y) + (o.z + o.w);
            ss = fmaf(o.x, o.x, fmaf(o.y, o.y, fmaf(o.z, o.z, fmaf(o.w, o.w, ss))));
          }
          *reinterpret_cast<float4*>(dst + hh * 32 + j) = o;
        }
      }
      __syncthreads();
      // coalesced copy-out: 16 lanes per row
#pragma unroll 4
      for (int it = 0; it < 8; ++it) {
        const int r = (tid >> 4) + it * 16;
        const float* sp = S + r * LDS + ch * 8;
        float* gp = a.y2_out + ((size_t)row0 + r) * H + ch * 8;
        *reinterpret_cast<float4*>(gp) = *reinterpret_cast<const float4*>(sp);
        *reinterpret_cast<float4*>(gp + 4) = *reinterpret_cast<const float4*>(sp + 4);
      }
      double ds = s, dss = ss;
      block_sum2(ds, dss, red);
      if (tid == 0) { t2s += ds; t2ss += dss; }
    }
    ph ^= 1u;
    tc::fence_before_sync();
    __syncthreads();
  }
  if (tid == 0) {
    a.parts1[2 * blockIdx.x] = t1s;
    a.parts1[2 * blockIdx.x + 1] = t1ss;
    if (a.parts2 != nullptr) {
      a.parts2[2 * blockIdx.x] = t2s;
      a.parts2[2 * blockIdx.x + 1] = t2ss;
    }
  }
  if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

int launch_edge_step_tc(const EdgeStepArgs& a, const uint8_t* img, int grid, cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute((const void*)k_edge_step_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_EDGE);
  if (e != cudaSuccess) { set_error("k_edge_step_tc smem attribute: %s", cudaGetErrorString(e)); return -2; }
  k_edge_step_tc<<<grid, NT, TC_SMEM_EDGE, st>>>(a, img + (size_t)IMG_PE_WE * tc::TILE_BF16_BYTES,
                                                 img + (size_t)IMG_PE_W2 * tc::TILE_BF16_BYTES);
  return 0;
}

}  // namespace pdg
