// tcgen05 flash attention (non-causal, no mask) for the CogVideoX DiT: head_dim 64, q/k/v read straight from the
// fused [rows, 3*heads*64] QKV buffer through one 2-D tensor map (no head-major re-layout).
//
// CTA = one 128-row query tile of one head; 192 threads:
//   warp 0  TMA producer (Q once, then a ring of K/V tiles of 128 keys)
//   warp 1  TMEM owner + single-thread MMA issuer:  S_j = Q K_j^T  (M128 N128 K64, both operands K-major smem)
//                                                   PV_j = P_j V_j (M128 N64  K128, A = P from TMEM (or smem),
//                                                                   B = V tile consumed MN-major as it sits)
//   warps 2..5  softmax: one query row per thread (tcgen05.ld 32x32b), online max/sum in log2 domain, P packed to
//               bf16 and written back to TMEM; the PV partial product of the previous tile is folded into the
//               register accumulator O while the tensor core works on the current one.
// TMEM columns: S[2] = 0..255, PV[2] = 256..383, P[2] = 384..511 (bf16 pairs).
#include "common.cuh"
#include "ptx.cuh"

namespace dove {

struct AttnParams {
  int rows, heads, nkv;
  float scale_log2;
  bf16* out;
};

constexpr int ATT_STAGES = 3;
constexpr uint32_t ATT_TILE_BYTES = 128 * 128;   // 128 rows x 64 bf16
constexpr size_t ATT_SMEM = 1024 + ATT_TILE_BYTES * (1 + 2 * ATT_STAGES + 2 * 2) + 256;

template <bool kPTmem>
__global__ void __launch_bounds__(192, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tm, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + ATT_TILE_BYTES;
  uint8_t* sV = sK + ATT_STAGES * ATT_TILE_BYTES;
  uint8_t* sP = sV + ATT_STAGES * ATT_TILE_BYTES;   // [2 buffers][2 k-halves][128 x 128 B] (smem-P variant only)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 4 * ATT_TILE_BYTES);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;
  uint64_t* kv_empty = kv_full + ATT_STAGES;
  uint64_t* s_full = kv_empty + ATT_STAGES;
  uint64_t* p_full = s_full + 2;
  uint64_t* o_full = p_full + 2;
  uint64_t* o_empty = o_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(o_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int head = blockIdx.y;
  const int q0 = blockIdx.x * 128;
  const int qcol = head * 64, kcol = (p.heads + head) * 64, vcol = (2 * p.heads + head) * 64;
  const int nkv = p.nkv;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm);
    mbar_init(q_full, 1);
    for (int i = 0; i < ATT_STAGES; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 4);
      mbar_init(&o_full[i], 1);
      mbar_init(&o_empty[i], 4);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t TS = tmem_base, TO = tmem_base + 256, TP = tmem_base + 384;

  if (warp == 0 && lane == 0) {
    // ===================== TMA producer =====================
    mbar_expect_tx(q_full, ATT_TILE_BYTES);
    tma_load_2d(sQ, &tm, q_full, qcol, q0);
    int stage = 0;
    uint32_t phase = 0;
    for (int j = 0; j < nkv; ++j) {
      mbar_wait(&kv_empty[stage], phase ^ 1);
      mbar_expect_tx(&kv_full[stage], 2 * ATT_TILE_BYTES);
      tma_load_2d(sK + stage * ATT_TILE_BYTES, &tm, &kv_full[stage], kcol, j * 128);
      tma_load_2d(sV + stage * ATT_TILE_BYTES, &tm, &kv_full[stage], vcol, j * 128);
      if (++stage == ATT_STAGES) {
        stage = 0;
        phase ^= 1;
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);   // Q (K-major) x K (K-major)
    constexpr uint32_t idesc_o = umma_idesc_bf16(128, 64, 0, 1);    // P (K-major / TMEM) x V (MN-major)
    const uint64_t qdesc = umma_desc_sw128(smem_u32(sQ));
    auto issue_s = [&](int j, int stage) {
      const uint64_t kdesc = umma_desc_sw128(smem_u32(sK + stage * ATT_TILE_BYTES));
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_ss(TS + (j & 1) * 128, qdesc + 2 * k, kdesc + 2 * k, idesc_s, k != 0);
      umma_commit(&s_full[j & 1]);
    };
    mbar_wait(q_full, 0);
    int stage_s = 0;
    uint32_t phase_s = 0;   // ring position of the next tile whose S is to be issued
    int stage_o = 0;        // ring position of tile j (PV)
    mbar_wait(&kv_full[0], 0);
    tc_fence_after();
    issue_s(0, 0);
    if (++stage_s == ATT_STAGES) { stage_s = 0; phase_s ^= 1; }
    for (int j = 0; j < nkv; ++j) {
      const int b = j & 1;
      const uint32_t ph2 = (j >> 1) & 1;
      if (j + 1 < nkv) {
        mbar_wait(&kv_full[stage_s], phase_s);
        tc_fence_after();
        issue_s(j + 1, stage_s);
        if (++stage_s == ATT_STAGES) { stage_s = 0; phase_s ^= 1; }
      }
      mbar_wait(&p_full[b], ph2);
      mbar_wait(&o_empty[b], ph2 ^ 1);
      tc_fence_after();
      const uint64_t vdesc = umma_desc_sw128(smem_u32(sV + stage_o * ATT_TILE_BYTES));
#pragma unroll
      for (int k = 0; k < 8; ++k) {   // K = 128 keys, 16 per instruction: V advances 16 rows = 2048 B
        if (kPTmem) {
          umma_ts(TO + b * 64, TP + b * 64 + k * 8, vdesc + 128 * k, idesc_o, k != 0);
        } else {
          const uint64_t pdesc =
              umma_desc_sw128(smem_u32(sP + (b * 2 + (k >> 2)) * ATT_TILE_BYTES)) + 2 * (k & 3);
          umma_ss(TO + b * 64, pdesc, vdesc + 128 * k, idesc_o, k != 0);
        }
      }
      umma_commit(&kv_empty[stage_o]);
      umma_commit(&o_full[b]);
      if (++stage_o == ATT_STAGES) stage_o = 0;
    }
  } else if (warp >= 2) {
    // ===================== softmax / accumulate =====================
    const int q = warp & 3;
    const int r = q * 32 + lane;            // query row in tile == TMEM lane
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    float O[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) O[i] = 0.f;
    float m_run = -INFINITY, l_run = 0.f, alpha_prev = 1.f;

    auto fold_pv = [&](int jb, uint32_t ph) {   // O = O*alpha_prev + PV[jb]
      mbar_wait(&o_full[jb], ph);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        tmem_ld32(TO + lane_off + jb * 64 + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) O[c * 32 + i] = O[c * 32 + i] * alpha_prev + __uint_as_float(v[i]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_empty[jb]);
    };

    for (int j = 0; j < nkv; ++j) {
      const int b = j & 1;
      const uint32_t ph2 = (j >> 1) & 1;
      mbar_wait(&s_full[b], ph2);
      tc_fence_after();
      const uint32_t ts = TS + lane_off + b * 128;
      const int kbase = j * 128;
      const bool tail = kbase + 128 > p.rows;
      // pass 1: row max
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld32(ts + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          float s = __uint_as_float(v[i]);
          if (tail && kbase + c * 32 + i >= p.rows) s = -INFINITY;
          mx = fmaxf(mx, s);
        }
      }
      const float m_new = fmaxf(m_run, mx * p.scale_log2);
      const float alpha = exp2f(m_run - m_new);
      float lsum = 0.f;
      // pass 2: p = exp2(s*scale - m), pack to bf16, store
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld32(ts + c * 32, v);
        tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float s0 = __uint_as_float(v[2 * i]) * p.scale_log2 - m_new;
          float s1 = __uint_as_float(v[2 * i + 1]) * p.scale_log2 - m_new;
          float p0 = exp2f(s0), p1 = exp2f(s1);
          if (tail) {
            if (kbase + c * 32 + 2 * i >= p.rows) p0 = 0.f;
            if (kbase + c * 32 + 2 * i + 1 >= p.rows) p1 = 0.f;
          }
          lsum += p0 + p1;
          pk[i] = pack_bf16x2(p0, p1);
        }
        if (kPTmem) {
          tmem_st16(TP + lane_off + b * 64 + c * 16, pk);
        } else {
          // K-major SW128 tile: row r, 16-byte chunk index (c*4+jj) within the 64-wide half (c >> 1)
          uint8_t* base = sP + (b * 2 + (c >> 1)) * ATT_TILE_BYTES + (r >> 3) * 1024 + (r & 7) * 128;
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            const int chunk = ((c & 1) * 4 + jj) ^ (r & 7);
            *reinterpret_cast<uint4*>(base + chunk * 16) =
                make_uint4(pk[jj * 4], pk[jj * 4 + 1], pk[jj * 4 + 2], pk[jj * 4 + 3]);
          }
        }
      }
      if (kPTmem) tmem_st_wait(); else fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[b]);
      l_run = l_run * alpha + lsum;
      if (j > 0) fold_pv(b ^ 1, ((j - 1) >> 1) & 1);
      alpha_prev = alpha;
      m_run = m_new;
    }
    fold_pv((nkv - 1) & 1, ((nkv - 1) >> 1) & 1);
    const int row = q0 + r;
    if (row < p.rows) {
      const float inv = 1.0f / l_run;
      uint4* op = reinterpret_cast<uint4*>(p.out + static_cast<long long>(row) * (p.heads * 64) + head * 64);
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        uint4 o;
        o.x = pack_bf16x2(O[jj * 8 + 0] * inv, O[jj * 8 + 1] * inv);
        o.y = pack_bf16x2(O[jj * 8 + 2] * inv, O[jj * 8 + 3] * inv);
        o.z = pack_bf16x2(O[jj * 8 + 4] * inv, O[jj * 8 + 5] * inv);
        o.w = pack_bf16x2(O[jj * 8 + 6] * inv, O[jj * 8 + 7] * inv);
        op[jj] = o;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

static int attention_launch(const void* qkv, void* out, int rows, int heads, float scale, int variant,
                            cudaStream_t st) {
  if (int e = ensure_init()) return e;
  DOVE_CHECK_ARG(rows > 0 && heads > 0, "attention: empty problem");
  DOVE_CHECK_ARG(reinterpret_cast<uintptr_t>(qkv) % 16 == 0 && reinterpret_cast<uintptr_t>(out) % 16 == 0,
                 "attention: pointers must be 16-byte aligned");
  CUtensorMap tm;
  uint64_t dims[2] = {static_cast<uint64_t>(3) * heads * 64, static_cast<uint64_t>(rows)};
  uint64_t strides[1] = {static_cast<uint64_t>(3) * heads * 64 * 2};
  uint32_t box[2] = {64, 128};
  if (int e = make_tmap_bf16(&tm, qkv, 2, dims, strides, box, nullptr)) return e;
  AttnParams p;
  p.rows = rows;
  p.heads = heads;
  p.nkv = (rows + 127) / 128;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.out = static_cast<bf16*>(out);
  dim3 grid((rows + 127) / 128, heads);
  cudaError_t e;
  if (variant == 0) {
    e = cudaFuncSetAttribute(attn_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             static_cast<int>(ATT_SMEM));
    if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(attn)");
    attn_fwd_kernel<true><<<grid, 192, ATT_SMEM, st>>>(tm, p);
  } else {
    e = cudaFuncSetAttribute(attn_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             static_cast<int>(ATT_SMEM));
    if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(attn)");
    attn_fwd_kernel<false><<<grid, 192, ATT_SMEM, st>>>(tm, p);
  }
  DOVE_LAUNCH_CHECK("attn_fwd_kernel");
  return DOVE_OK;
}

}  // namespace dove

extern "C" int dove_attention_bf16(const void* qkv, void* out, int rows, int heads, float scale, void* stream) {
  return dove::attention_launch(qkv, out, rows, heads, scale, 0, static_cast<cudaStream_t>(stream));
}

// Test hook: variant 0 = P through TMEM (TS MMA), 1 = P through swizzled shared memory (SS MMA).
extern "C" int dove_attention_bf16_variant(const void* qkv, void* out, int rows, int heads, float scale,
                                           int variant, void* stream) {
  return dove::attention_launch(qkv, out, rows, heads, scale, variant, static_cast<cudaStream_t>(stream));
}
