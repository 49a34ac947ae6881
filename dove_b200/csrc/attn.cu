// tcgen05 flash attention (non-causal, no mask) for the CogVideoX DiT: head_dim 64, q/k/v read straight from the
// fused [rows, 3*heads*64] QKV buffer through one 2-D tensor map (no head-major re-layout).
//
// CTA = one 128-row query tile of one head, 192 threads, TWO CTAs PER SM (80 KB smem, 256 TMEM columns each) so
// that one CTA's softmax overlaps the other CTA's MMAs:
//   warp 0      TMA producer (Q once, then a 2-stage ring of K/V tiles of 128 keys)
//   warp 1      TMEM owner + single-thread MMA issuer:
//                   S_j  = Q K_j^T      M128 N128 K64, both operands K-major smem (SW128)
//                   O   += P_j V_j      M128 N64 K128, A = P from TMEM (bf16), B = V tile consumed MN-major
//   warps 2..5  softmax, one query row per thread (tcgen05.ld 32x32b): single pass over the 128 S values held in
//               registers, online max/sum in the log2 domain, P packed to bf16 and written to TMEM.
// O accumulates in TMEM across all KV tiles.  Rescaling is LAZY: the running max used for the exponent is only
// advanced when the true row max grew by more than 8 (log2 units), so O needs a TMEM read-modify-write only on
// the rare tiles where that happens (P stays <= 2^8, well inside bf16/fp32 range); the final 1/l normalisation
// uses the same stale max, so the result is exact softmax.
// TMEM columns: S = 0..127, O = 128..191, P = 192..255 (bf16 pairs).
#include <atomic>
#include <mutex>

#include "common.cuh"
#include "ptx.cuh"

namespace dove {

struct AttnParams {
  int rows, heads, nkv;
  float scale_log2;
  bf16* out;
};

constexpr int ATT_STAGES = 2;
constexpr uint32_t ATT_TILE_BYTES = 128 * 128;   // 128 rows x 64 bf16
constexpr size_t ATT_SMEM = 1024 + ATT_TILE_BYTES * (1 + 2 * ATT_STAGES) + 128;
constexpr float ATT_LAZY_THRESHOLD = 8.0f;

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// packed fp32x2 arithmetic (sm_100): one issue slot for two lanes of the softmax's scale-subtract and row-sum
__device__ __forceinline__ uint64_t pack_f32x2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack_f32x2(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t fma_f32x2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t add_f32x2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

__global__ void __launch_bounds__(192, 2)
attn_fwd_v2_kernel(const __grid_constant__ CUtensorMap tm, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + ATT_TILE_BYTES;
  uint8_t* sV = sK + ATT_STAGES * ATT_TILE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + ATT_STAGES * ATT_TILE_BYTES);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;
  uint64_t* kv_empty = kv_full + ATT_STAGES;
  uint64_t* s_full = kv_empty + ATT_STAGES;
  uint64_t* p_full = s_full + 1;
  uint64_t* pv_done = p_full + 1;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(pv_done + 1);

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
    mbar_init(s_full, 1);
    mbar_init(p_full, 4);
    mbar_init(pv_done, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<256>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t TS = tmem_base, TO = tmem_base + 128, TP = tmem_base + 192;

  if (warp == 0 && elect_one()) {
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
  } else if (warp == 1 && elect_one()) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);   // Q (K-major) x K (K-major)
    constexpr uint32_t idesc_o = umma_idesc_bf16(128, 64, 0, 1);    // P (TMEM) x V (MN-major)
    const uint64_t qdesc = umma_desc_sw128(smem_u32(sQ));
    auto issue_s = [&](int stage) {
      const uint64_t kdesc = umma_desc_sw128(smem_u32(sK + stage * ATT_TILE_BYTES));
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_ss(TS, qdesc + 2 * k, kdesc + 2 * k, idesc_s, k != 0);
      umma_commit(s_full);
    };
    mbar_wait(q_full, 0);
    mbar_wait(&kv_full[0], 0);
    tc_fence_after();
    issue_s(0);
    for (int j = 0; j < nkv; ++j) {
      const int stage = j & 1;
      mbar_wait(p_full, j & 1);            // softmax_j has read all of S_j and written P_j
      tc_fence_after();
      if (j + 1 < nkv) {                   // S_{j+1} first: softmax_{j+1} can start while PV_j runs
        mbar_wait(&kv_full[stage ^ 1], ((j + 1) >> 1) & 1);
        tc_fence_after();
        issue_s(stage ^ 1);
      }
      const uint64_t vdesc = umma_desc_sw128(smem_u32(sV + stage * ATT_TILE_BYTES));
#pragma unroll
      for (int k = 0; k < 8; ++k)   // K = 128 keys, 16 per instruction: P +8 TMEM columns, V +16 rows = 2048 B
        umma_ts(TO, TP + k * 8, vdesc + 128 * k, idesc_o, (j | k) != 0);
      umma_commit(&kv_empty[stage]);
      umma_commit(pv_done);
    }
  } else if (warp >= 2) {
    // ===================== softmax =====================
    const int q = warp & 3;
    const int r = q * 32 + lane;            // query row in tile == TMEM lane
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    float m_used = -INFINITY, l_run = 0.f;
    const float sl2 = p.scale_log2;

    for (int j = 0; j < nkv; ++j) {
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      uint32_t v[128];
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld32(TS + lane_off + c * 32, v + c * 32);
      tmem_ld_wait();
      const int kbase = j * 128;
      if (kbase + 128 > p.rows) {            // tail tile: keys beyond `rows` are masked out
#pragma unroll
        for (int i = 0; i < 128; ++i)
          if (kbase + i >= p.rows) v[i] = 0xff800000u;   // -inf
      }
      float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};   // 4 independent chains (ILP)
#pragma unroll
      for (int i = 0; i < 128; i += 8) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
          mx4[u] = fmaxf(mx4[u], fmaxf(__uint_as_float(v[i + 2 * u]), __uint_as_float(v[i + 2 * u + 1])));
      }
      const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
      const float m_new = fmaxf(m_used, mx * sl2);
      const bool need = (m_new - m_used) > ATT_LAZY_THRESHOLD;     // first tile: inf > 8
      const bool warp_need = __any_sync(0xffffffffu, need);
      float alpha = 1.0f;
      if (need) {
        alpha = ex2_approx(m_used - m_new);
        m_used = m_new;
      }
      // exponentials first (MUFU-bound phase, overlaps PV_{j-1} still running on the tensor pipe) ...
      const float neg_m = -m_used;
      const uint64_t sl2_2 = pack_f32x2(sl2, sl2), negm_2 = pack_f32x2(neg_m, neg_m);
      uint64_t lsum2a = 0ull, lsum2b = 0ull;      // two packed accumulators (4 independent add chains)
      uint32_t pk[64];
#pragma unroll
      for (int i = 0; i < 64; i += 2) {
        const uint64_t xa = fma_f32x2(pack_f32x2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1])), sl2_2, negm_2);
        const uint64_t xb = fma_f32x2(pack_f32x2(__uint_as_float(v[2 * i + 2]), __uint_as_float(v[2 * i + 3])), sl2_2, negm_2);
        float a0, a1, b0, b1;
        unpack_f32x2(xa, a0, a1);
        unpack_f32x2(xb, b0, b1);
        // (emulating part of the exp2s with an FMA-pipe polynomial was measured 39 % SLOWER here: the softmax warps are
        //  as issue-limited as they are MUFU-limited, profiles/r01_microbench.txt)
        const float p0 = ex2_approx(a0), p1 = ex2_approx(a1), p2 = ex2_approx(b0), p3 = ex2_approx(b1);
        lsum2a = add_f32x2(lsum2a, pack_f32x2(p0, p1));
        lsum2b = add_f32x2(lsum2b, pack_f32x2(p2, p3));
        pk[i] = pack_bf16x2(p0, p1);
        pk[i + 1] = pack_bf16x2(p2, p3);
      }
      float ls0, ls1, ls2, ls3;
      unpack_f32x2(lsum2a, ls0, ls1);
      unpack_f32x2(lsum2b, ls2, ls3);
      const float lsum = (ls0 + ls1) + (ls2 + ls3);
      // ... then wait for PV_{j-1}: the P buffer is free and O is stable from here on
      if (j > 0) {
        mbar_wait(pv_done, (j - 1) & 1);
        tc_fence_after();
        if (warp_need) {                     // rare after the first few tiles: O *= alpha (TMEM read-modify-write)
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint32_t o[32];
            tmem_ld32(TO + lane_off + c * 32, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st32(TO + lane_off + c * 32, o);
          }
        }
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_st16(TP + lane_off + c * 16, pk + c * 16);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
      l_run = l_run * alpha + lsum;
    }
    mbar_wait(pv_done, (nkv - 1) & 1);
    tc_fence_after();
    const int row = q0 + r;
    const float inv = 1.0f / l_run;
    uint4* op = reinterpret_cast<uint4*>(p.out + static_cast<long long>(row) * (p.heads * 64) + head * 64);
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t o[32];
      tmem_ld32(TO + lane_off + c * 32, o);
      tmem_ld_wait();
      if (row < p.rows) {
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          uint4 w;
          w.x = pack_bf16x2(__uint_as_float(o[jj * 8 + 0]) * inv, __uint_as_float(o[jj * 8 + 1]) * inv);
          w.y = pack_bf16x2(__uint_as_float(o[jj * 8 + 2]) * inv, __uint_as_float(o[jj * 8 + 3]) * inv);
          w.z = pack_bf16x2(__uint_as_float(o[jj * 8 + 4]) * inv, __uint_as_float(o[jj * 8 + 5]) * inv);
          w.w = pack_bf16x2(__uint_as_float(o[jj * 8 + 6]) * inv, __uint_as_float(o[jj * 8 + 7]) * inv);
          op[c * 4 + jj] = w;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<256>(tmem_base);
  }
}

// =====================================================================================================================
// v3: ONE CTA PER SM, TWO 128-row query tiles per CTA (256 query rows) that share every staged K/V tile, so the
// L2 -> SM traffic per query row is half of v2's, and a leaner softmax:
//   * no row-max pass in the steady state: the exponent reference m_used is kept from earlier tiles and the
//     probabilities are computed directly; the row sum of the tile (and the largest argument of the FMA-pipe
//     exponential) detects the rare case where a score exceeded m_used by more than 64 log2 units, and only then the
//     exact path (row max, rescale of O in TMEM) is run.
//     The first KV tile and a ragged last tile always take the exact path.
//   * the MUFU unit (16 ex2/clk/SM) bounds a head_dim-64 attention at 50 % tensor pipe, so EMU8/8 of the exponentials
//     are evaluated on the FMA pipe instead (Cody-Waite split with the 1.5*2^23 magic add + a degree-3 minimax
//     polynomial on [-0.5, 0.5], relative error 7.5e-5 << the bf16 rounding of P; exponent merged with one
//     integer shift-add), all in packed f32x2 FMAs.
//   warp 0 TMA producer (Q0, Q1 once; 3-stage ring of K/V tiles), warp 1 TMEM owner + MMA issuer,
//   warps 2..5 softmax of query tile 0, warps 6..9 softmax of query tile 1 (one row per thread).
// TMEM (512 columns): tile t at base + 256 t:  S = +0..127 (fp32), O = +128..191 (fp32), P = +192..255 (bf16 pairs).
constexpr int A3_STAGES = 3;
constexpr size_t A3_SMEM = 1024 + ATT_TILE_BYTES * (2 + 2 * A3_STAGES) + 256;
// Steady-state validity bound: the stale reference may lag the true row max by up to 2^6 = 64 log2 units before the tile is
// redone exactly.  Nothing needs the probabilities to be <= 1: P is bf16 (fp32 exponent range), S/P/O/l are fp32, and
// p <= 2^64 keeps every partial sum (128 keys x 152 tiles x |v|) far below fp32 overflow.  A TIGHT bound would be wrong
// here: with a sum threshold of 2^8 a row whose first key tile (the 226 text tokens) scores a few units below the video
// keys would take the exact path on every later tile (measured: 2x slower on the real DiT activations).
constexpr float A3_SUM_LIMIT = 1.8446744e19f * 128.0f;   // 128 keys x 2^64
constexpr float A3_ARG_LIMIT = 12582912.0f + 64.0f;      // magic-shifted exponent argument of the FMA-pipe path

__host__ __device__ constexpr bool a3_emu_pair(int i, int emu8) { return ((i * emu8) % 8) < emu8 && emu8 > 0; }

// 2^x on the FMA pipe, two lanes at a time.  x below -126 is clamped (result ~ 0, exact enough: such a key has weight
// < 2^-126); x above 127 would wrap the exponent field SILENTLY, so the largest magic-shifted argument seen is
// tracked in tmax (one 3-input max per pair) and the caller reruns the tile on the exact path when it exceeds 127.
__device__ __forceinline__ void ex2_emu_x2(uint64_t x2, float& p0, float& p1, float& tmax) {
  float x0, x1;
  unpack_f32x2(x2, x0, x1);
  x0 = fmaxf(x0, -126.0f);
  x1 = fmaxf(x1, -126.0f);
  const uint64_t xc = pack_f32x2(x0, x1);
  const uint64_t magic = pack_f32x2(12582912.0f, 12582912.0f);          // 1.5 * 2^23: low mantissa bits = round(x)
  const uint64_t nmagic = pack_f32x2(-12582912.0f, -12582912.0f);
  const uint64_t t2 = add_f32x2(xc, magic);
  const uint64_t n2 = add_f32x2(t2, nmagic);                            // round(x) as a float
  const uint64_t r2 = fma_f32x2(n2, pack_f32x2(-1.0f, -1.0f), xc);      // r = x - round(x) in [-0.5, 0.5]
  uint64_t q2 = fma_f32x2(pack_f32x2(0.05517132207751274f, 0.05517132207751274f), r2,
                          pack_f32x2(0.24261054396629333f, 0.24261054396629333f));
  q2 = fma_f32x2(q2, r2, pack_f32x2(0.6932609677314758f, 0.6932609677314758f));
  q2 = fma_f32x2(q2, r2, pack_f32x2(0.9999281167984009f, 0.9999281167984009f));
  float q0, q1, t0, t1;
  unpack_f32x2(q2, q0, q1);
  unpack_f32x2(t2, t0, t1);
  tmax = fmaxf(tmax, fmaxf(t0, t1));
  p0 = __uint_as_float(__float_as_uint(q0) + (__float_as_uint(t0) << 23));
  p1 = __uint_as_float(__float_as_uint(q1) + (__float_as_uint(t1) << 23));
}

template <int EMU8>
__global__ void __launch_bounds__(320, 1)
attn_fwd_v3_kernel(const __grid_constant__ CUtensorMap tm, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                                   // 2 query tiles
  uint8_t* sK = sQ + 2 * ATT_TILE_BYTES;
  uint8_t* sV = sK + A3_STAGES * ATT_TILE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + A3_STAGES * ATT_TILE_BYTES);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;
  uint64_t* kv_empty = kv_full + A3_STAGES;
  uint64_t* s_full = kv_empty + A3_STAGES;              // [2]
  uint64_t* p_full = s_full + 2;                        // [2]
  uint64_t* pv_done = p_full + 2;                       // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(pv_done + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int head = blockIdx.y;
  const int q0 = blockIdx.x * 256;
  const int qcol = head * 64, kcol = (p.heads + head) * 64, vcol = (2 * p.heads + head) * 64;
  const int nkv = p.nkv;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm);
    mbar_init(q_full, 1);
    for (int i = 0; i < A3_STAGES; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&s_full[t], 1);
      mbar_init(&p_full[t], 4);
      mbar_init(&pv_done[t], 1);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0 && elect_one()) {
    // ===================== TMA producer =====================
    mbar_expect_tx(q_full, 2 * ATT_TILE_BYTES);
    tma_load_2d(sQ, &tm, q_full, qcol, q0);
    tma_load_2d(sQ + ATT_TILE_BYTES, &tm, q_full, qcol, q0 + 128);
    int stage = 0;
    uint32_t phase = 0;
    for (int j = 0; j < nkv; ++j) {
      mbar_wait(&kv_empty[stage], phase ^ 1);
      mbar_expect_tx(&kv_full[stage], 2 * ATT_TILE_BYTES);
      tma_load_2d(sK + stage * ATT_TILE_BYTES, &tm, &kv_full[stage], kcol, j * 128);
      tma_load_2d(sV + stage * ATT_TILE_BYTES, &tm, &kv_full[stage], vcol, j * 128);
      if (++stage == A3_STAGES) {
        stage = 0;
        phase ^= 1;
      }
    }
  } else if (warp == 1 && elect_one()) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);   // Q (K-major) x K (K-major)
    constexpr uint32_t idesc_o = umma_idesc_bf16(128, 64, 0, 1);    // P (TMEM) x V (MN-major)
    auto issue_s = [&](int t, int stage) {
      const uint64_t qdesc = umma_desc_sw128(smem_u32(sQ + t * ATT_TILE_BYTES));
      const uint64_t kdesc = umma_desc_sw128(smem_u32(sK + stage * ATT_TILE_BYTES));
      const uint32_t TS = tmem_base + t * 256;
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_ss(TS, qdesc + 2 * k, kdesc + 2 * k, idesc_s, k != 0);
      umma_commit(&s_full[t]);
    };
    mbar_wait(q_full, 0);
    mbar_wait(&kv_full[0], 0);
    tc_fence_after();
    issue_s(0, 0);
    issue_s(1, 0);
    int stage = 0, nstage = 1;
    uint32_t nphase = 0;                      // parity of kv_full[nstage] for tile j + 1
    for (int j = 0; j < nkv; ++j) {
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        mbar_wait(&p_full[t], j & 1);        // softmax_t(j) has read all of S_t(j) and written P_t(j)
        tc_fence_after();
        if (j + 1 < nkv) {                   // S_t(j+1) first: softmax_t(j+1) can start while PV_t(j) runs
          if (t == 0) {
            mbar_wait(&kv_full[nstage], nphase);
            tc_fence_after();
          }
          issue_s(t, nstage);
        }
        const uint64_t vdesc = umma_desc_sw128(smem_u32(sV + stage * ATT_TILE_BYTES));
        const uint32_t TO = tmem_base + t * 256 + 128, TP = tmem_base + t * 256 + 192;
#pragma unroll
        for (int k = 0; k < 8; ++k)   // K = 128 keys, 16 per instruction: P +8 TMEM columns, V +16 rows = 2048 B
          umma_ts(TO, TP + k * 8, vdesc + 128 * k, idesc_o, (j | k) != 0);
        umma_commit(&pv_done[t]);
      }
      umma_commit(&kv_empty[stage]);          // both tiles are done with K_j and V_j once everything above retires
      stage = nstage;
      if (++nstage == A3_STAGES) {
        nstage = 0;
        nphase ^= 1;
      }
    }
  } else if (warp >= 2) {
    // ===================== softmax =====================
    const int t = (warp - 2) >> 2;          // query tile of this warpgroup
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int r = q * 32 + lane;            // query row in tile == TMEM lane
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t TS = tmem_base + t * 256 + lane_off, TO = TS + 128, TP = TS + 192;
    float m_used = -INFINITY, l_run = 0.f;
    const float sl2 = p.scale_log2;
    const uint64_t sl2_2 = pack_f32x2(sl2, sl2);
    const bool ragged = (nkv * 128 != p.rows);

    for (int j = 0; j < nkv; ++j) {
      mbar_wait(&s_full[t], j & 1);
      tc_fence_after();
      bool exact = (j == 0) || (ragged && j == nkv - 1);
      bool waited_pv = (j == 0);            // P_t / O_t are only touched after PV_t(j-1) has retired
      float lsum = 0.f;
      if (!exact) {
        // ---------- steady state: probabilities straight from the stale reference max ----------
        const float neg_m = -m_used;
        const uint64_t negm_2 = pack_f32x2(neg_m, neg_m);
        uint64_t lsA = 0ull, lsB = 0ull;
        float tmax = 0.f;
        uint32_t vbuf[2][32];                  // TMEM loads run one 32-column chunk ahead of the arithmetic
        tmem_ld32(TS, vbuf[0]);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const uint32_t* v = vbuf[c & 1];
          tmem_ld_wait();
          if (c < 3) tmem_ld32(TS + (c + 1) * 32, vbuf[(c + 1) & 1]);
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const uint64_t x2 = fma_f32x2(pack_f32x2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1])), sl2_2,
                                          negm_2);
            float p0, p1;
            if (a3_emu_pair(i, EMU8)) {
              ex2_emu_x2(x2, p0, p1, tmax);
            } else {
              float x0, x1;
              unpack_f32x2(x2, x0, x1);
              p0 = ex2_approx(x0);
              p1 = ex2_approx(x1);
            }
            if (i & 1) lsB = add_f32x2(lsB, pack_f32x2(p0, p1));
            else lsA = add_f32x2(lsA, pack_f32x2(p0, p1));
            pk[i] = pack_bf16x2(p0, p1);
          }
          if (!waited_pv) {
            mbar_wait(&pv_done[t], (j - 1) & 1);
            tc_fence_after();
            waited_pv = true;
          }
          tmem_st16(TP + c * 16, pk);
        }
        float a0, a1, b0, b1;
        unpack_f32x2(lsA, a0, a1);
        unpack_f32x2(lsB, b0, b1);
        lsum = (a0 + a1) + (b0 + b1);
        // a score more than 8 log2 units above m_used (or a NaN) shows up in the row sum: redo this tile exactly
        exact = __any_sync(0xffffffffu, !(lsum <= A3_SUM_LIMIT) || tmax > A3_ARG_LIMIT);
      }
      float alpha = 1.0f;
      if (exact) {
        // ---------- exact path: row max, lazy rescale of O, masked tail (first / last / overflowing tiles) ----------
        const int kbase = j * 128;
        const bool tail = kbase + 128 > p.rows;
        float mx = -INFINITY;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t v[32];
          tmem_ld32(TS + c * 32, v);
          tmem_ld_wait();
          if (tail) {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (kbase + c * 32 + i >= p.rows) v[i] = 0xff800000u;   // -inf
          }
          float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
#pragma unroll
            for (int u = 0; u < 4; ++u)
              m4[u] = fmaxf(m4[u], fmaxf(__uint_as_float(v[i + 2 * u]), __uint_as_float(v[i + 2 * u + 1])));
          }
          mx = fmaxf(mx, fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])));
        }
        const float m_new = fmaxf(m_used, mx * sl2);
        const bool need = (m_new - m_used) > ATT_LAZY_THRESHOLD;     // first tile: inf > 8
        const bool warp_need = __any_sync(0xffffffffu, need);
        if (need) {
          alpha = ex2_approx(m_used - m_new);
          m_used = m_new;
        }
        if (!waited_pv) {
          mbar_wait(&pv_done[t], (j - 1) & 1);
          tc_fence_after();
          waited_pv = true;
        }
        if (j > 0 && warp_need) {            // O *= alpha (TMEM read-modify-write; rare after the first tiles)
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint32_t o[32];
            tmem_ld32(TO + c * 32, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st32(TO + c * 32, o);
          }
        }
        const float neg_m = -m_used;
        const uint64_t negm_2 = pack_f32x2(neg_m, neg_m);
        uint64_t lsA = 0ull, lsB = 0ull;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t v[32];
          tmem_ld32(TS + c * 32, v);
          tmem_ld_wait();
          if (tail) {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (kbase + c * 32 + i >= p.rows) v[i] = 0xff800000u;
          }
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const uint64_t x2 = fma_f32x2(pack_f32x2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1])), sl2_2,
                                          negm_2);
            float x0, x1;
            unpack_f32x2(x2, x0, x1);
            const float p0 = ex2_approx(x0), p1 = ex2_approx(x1);
            if (i & 1) lsB = add_f32x2(lsB, pack_f32x2(p0, p1));
            else lsA = add_f32x2(lsA, pack_f32x2(p0, p1));
            pk[i] = pack_bf16x2(p0, p1);
          }
          tmem_st16(TP + c * 16, pk);
        }
        float a0, a1, b0, b1;
        unpack_f32x2(lsA, a0, a1);
        unpack_f32x2(lsB, b0, b1);
        lsum = (a0 + a1) + (b0 + b1);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[t]);
      l_run = l_run * alpha + lsum;
    }
    mbar_wait(&pv_done[t], (nkv - 1) & 1);
    tc_fence_after();
    const int row = q0 + t * 128 + r;
    const float inv = 1.0f / l_run;
    uint4* op = reinterpret_cast<uint4*>(p.out + static_cast<long long>(row) * (p.heads * 64) + head * 64);
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t o[32];
      tmem_ld32(TO + c * 32, o);
      tmem_ld_wait();
      if (row < p.rows) {
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          uint4 w;
          w.x = pack_bf16x2(__uint_as_float(o[jj * 8 + 0]) * inv, __uint_as_float(o[jj * 8 + 1]) * inv);
          w.y = pack_bf16x2(__uint_as_float(o[jj * 8 + 2]) * inv, __uint_as_float(o[jj * 8 + 3]) * inv);
          w.z = pack_bf16x2(__uint_as_float(o[jj * 8 + 4]) * inv, __uint_as_float(o[jj * 8 + 5]) * inv);
          w.w = pack_bf16x2(__uint_as_float(o[jj * 8 + 6]) * inv, __uint_as_float(o[jj * 8 + 7]) * inv);
          op[c * 4 + jj] = w;
        }
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

// =====================================================================================================================
// v4: v3's tiling (one CTA per SM, two 128-row query tiles sharing every staged K/V tile) with the S(j+1) bubble removed.
// ncu on v3 (profiles/r02_ncu_attn_v3.txt): the softmax warps spend 24 % of their samples waiting for S(j+1), which v3
// can only issue after softmax(j) has stored all of P(j).  v4 breaks that dependency with REGISTERS instead of TMEM
// (TMEM is full: 2 x (S 128 + O 64 + P 64) = 512 columns):
//   * 384 threads = 3 warpgroups; `setmaxnreg` moves registers from warpgroup 0 (TMA producer, two MMA issuers, 1 idle
//     warp: 168 -> 56) to the two softmax warpgroups (168 -> 224), so a softmax thread holds its whole S row (128 fp32) in
//     registers without spilling (v3's dead end #2 was exactly this at 168 registers; a 40 / 232 split makes the MMA
//     issuers spill and react late: 5.53 vs 5.06 ms);
//   * the row arrives in eight 16-column tcgen05.ld pieces, four in flight (loading all of it up front measured 6.4 ms,
//     both warpgroups then sit in the TMEM read together); the S buffer is released at the midpoint of the tile (s_free)
//     and the MMA warp issues S_t(j+1) = Q_t K(j+1)^T while softmax_t(j) is still computing; P of the first half of a
//     tile is held in registers until PV_t(j-1) has retired, and a scheduling token keeps ptxas from issuing all
//     MUFU.EX2 of a tile in one burst (5.06 vs 5.40 ms);
//   * TWO MMA issuer threads (warp 1: query tile 0, warp 2: query tile 1), each on a fixed S_t(n+1), PV_t(n) sequence;
//   * K and V have separate 4-stage rings (K(j+1) is consumed a whole softmax period before V(j) is released);
//   * the exponent reference m_used is kept INTEGER (any reference is valid, the normalisation uses the same one), which
//     makes the range reduction of the FMA-pipe exponential 3 packed operations: t = s*c + (magic - m) (exact integer
//     part), d = (magic - m) - t = -m - round(x), r = s*c + d in [-0.5, 0.5]; then the degree-3 polynomial and one
//     shift-add.  The exact path (first tile, ragged last tile, overflow of the stale reference) reuses the registers.
#ifdef DOVE_ATTN_TRACE   // profiles/trace_attn.cu: per-actor event timeline (clock64) of one CTA, tiles 40..55
__device__ unsigned long long g_attn_trace[4][1024];
__device__ int g_attn_trace_n[4];
#define ATR_DECL int atr_k = 0
#define ATR(actor, ev, tile)   /* plain store, counter in a register: no round trip on the traced thread */       \
  do {                                                                                                          \
    if (blockIdx.x == 2 && blockIdx.y == 1 && (tile) >= 40 && (tile) < 56 && atr_k < 1024)                      \
      g_attn_trace[actor][atr_k++] = (static_cast<unsigned long long>(clock64()) & 0xffffffffffull) |           \
                                     (static_cast<unsigned long long>(ev) << 40) |                              \
                                     (static_cast<unsigned long long>(tile) << 48);                             \
  } while (0)
#define ATR_END(actor) do { if (blockIdx.x == 2 && blockIdx.y == 1) g_attn_trace_n[actor] = atr_k; } while (0)
#else
#define ATR_DECL do {} while (0)
#define ATR(actor, ev, tile) do {} while (0)
#define ATR_END(actor) do {} while (0)
#endif
constexpr int A4_KS = 4, A4_VS = 4;
constexpr size_t A4_SMEM = 1024 + ATT_TILE_BYTES * (2 + A4_KS + A4_VS) + 256;

template <int EMU8, bool TOKEN = true, int REG_MISC = 56, int REG_SOFTMAX = 224>
__global__ void __launch_bounds__(384, 1)
attn_fwd_v4_kernel(const __grid_constant__ CUtensorMap tm, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                                   // 2 query tiles
  uint8_t* sK = sQ + 2 * ATT_TILE_BYTES;
  uint8_t* sV = sK + A4_KS * ATT_TILE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + A4_VS * ATT_TILE_BYTES);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;
  uint64_t* k_empty = k_full + A4_KS;
  uint64_t* v_full = k_empty + A4_KS;
  uint64_t* v_empty = v_full + A4_VS;
  uint64_t* s_full = v_empty + A4_VS;                   // [2]  MMA -> softmax: S_t(j) complete
  uint64_t* s_free = s_full + 2;                        // [2]  softmax -> MMA: S_t(j) is in registers
  uint64_t* p_full = s_free + 2;                        // [2]  softmax -> MMA: P_t(j) stored
  uint64_t* pv_done = p_full + 2;                       // [2]  MMA -> softmax: O_t += P_t(j) V(j) retired
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(pv_done + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int head = blockIdx.y;
  const int q0 = blockIdx.x * 256;
  const int qcol = head * 64, kcol = (p.heads + head) * 64, vcol = (2 * p.heads + head) * 64;
  const int nkv = p.nkv;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm);
    mbar_init(q_full, 1);
    for (int i = 0; i < A4_KS; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 2);
    }
    for (int i = 0; i < A4_VS; ++i) {
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 2);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&s_full[t], 1);
      mbar_init(&s_free[t], 4);
      mbar_init(&p_full[t], 4);
      mbar_init(&pv_done[t], 1);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REG_MISC));
    if (warp == 0 && elect_one()) {
      // ===================== TMA producer =====================
      mbar_expect_tx(q_full, 2 * ATT_TILE_BYTES);
      tma_load_2d(sQ, &tm, q_full, qcol, q0);
      tma_load_2d(sQ + ATT_TILE_BYTES, &tm, q_full, qcol, q0 + 128);
      int ks = 0, vs = 0;
      uint32_t kph = 0, vph = 0;
      for (int j = 0; j < nkv; ++j) {
        mbar_wait(&k_empty[ks], kph ^ 1);
        mbar_expect_tx(&k_full[ks], ATT_TILE_BYTES);
        tma_load_2d(sK + ks * ATT_TILE_BYTES, &tm, &k_full[ks], kcol, j * 128);
        if (++ks == A4_KS) ks = 0, kph ^= 1;
        mbar_wait(&v_empty[vs], vph ^ 1);
        mbar_expect_tx(&v_full[vs], ATT_TILE_BYTES);
        tma_load_2d(sV + vs * ATT_TILE_BYTES, &tm, &v_full[vs], vcol, j * 128);
        if (++vs == A4_VS) vs = 0, vph ^= 1;
      }
    } else if ((warp == 1 || warp == 2) && elect_one()) {
      // ===================== MMA issuers: warp 1 drives query tile 0, warp 2 query tile 1 =====================
      // Per query tile the hand-offs alternate strictly — s_free_t(n) at the midpoint of softmax_t(n), p_full_t(n) at
      // its end — so each issuer runs a fixed S_t(n+1), PV_t(n) sequence on blocking waits.  ONE thread serving both
      // tiles (polling four barriers, or a fixed S0 PV1 S1 PV0 order) needs ~450 clk per action and was the critical
      // path (event trace, profiles/trace_attn.cu: 500-clk reaction time to every hand-off, 3100 clk per key tile).
      const int t = warp - 1;
      ATR_DECL;
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);   // Q (K-major) x K (K-major)
      constexpr uint32_t idesc_o = umma_idesc_bf16(128, 64, 0, 1);    // P (TMEM) x V (MN-major)
      const uint64_t qdesc = umma_desc_sw128(smem_u32(sQ + t * ATT_TILE_BYTES));
      const uint64_t kdesc0 = umma_desc_sw128(smem_u32(sK));
      const uint64_t vdesc0 = umma_desc_sw128(smem_u32(sV));
      const uint32_t TS = tmem_base + t * 256, TO = TS + 128, TP = TS + 192;
      auto issue_s = [&](int kstage) {
        const uint64_t kdesc = kdesc0 + static_cast<uint64_t>(kstage) * (ATT_TILE_BYTES >> 4);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_ss(TS, qdesc + 2 * k, kdesc + 2 * k, idesc_s, k != 0);
        umma_commit(&s_full[t]);
        umma_commit(&k_empty[kstage]);          // count 2: released when both issuers' reads of K have retired
      };
      auto issue_pv = [&](int vstage, bool first) {
        const uint64_t vdesc = vdesc0 + static_cast<uint64_t>(vstage) * (ATT_TILE_BYTES >> 4);
#pragma unroll
        for (int k = 0; k < 8; ++k)   // K = 128 keys, 16 per instruction: P +8 TMEM columns, V +16 rows = 2048 B
          umma_ts(TO, TP + k * 8, vdesc + 128 * k, idesc_o, !(first && k == 0));
        umma_commit(&pv_done[t]);
        umma_commit(&v_empty[vstage]);
      };
      static_assert(A4_KS == 4 && A4_VS == 4, "stage = n & 3, parity = (n >> 2) & 1");
      mbar_wait(q_full, 0);
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      issue_s(0);
      for (int n = 0; n < nkv; ++n) {
        if (n + 1 < nkv) {                      // S_t(n+1): softmax_t(n) has its S row in registers
          mbar_wait(&k_full[(n + 1) & 3], ((n + 1) >> 2) & 1);
          mbar_wait(&s_free[t], n & 1);
          tc_fence_after();
          ATR(t, 1 + t, n + 1);
          issue_s((n + 1) & 3);
        }
        mbar_wait(&v_full[n & 3], (n >> 2) & 1);
        mbar_wait(&p_full[t], n & 1);           // PV_t(n): softmax_t(n) has stored P_t(n)
        tc_fence_after();
        ATR(t, 3 + t, n);
        issue_pv(n & 3, n == 0);
      }
      ATR_END(t);
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REG_SOFTMAX));
    // ===================== softmax =====================
    const int t = (warp - 4) >> 2;          // query tile of this warpgroup
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int r = q * 32 + lane;            // query row in tile == TMEM lane
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t TS = tmem_base + t * 256 + lane_off, TO = TS + 128, TP = TS + 192;
    float m_used = -INFINITY, l_run = 0.f;      // m_used is an integer once set
    ATR_DECL;
    const float sl2 = p.scale_log2;
    const float inv_sl2 = 1.0f / sl2;
    const uint64_t sl2_2 = pack_f32x2(sl2, sl2);
    const bool ragged = (nkv * 128 != p.rows);

    for (int j = 0; j < nkv; ++j) {
      if (q == 0 && lane == 0) ATR(2 + t, 9, j);
      mbar_wait(&s_full[t], j & 1);
      tc_fence_after();
      if (q == 0 && lane == 0) ATR(2 + t, 10, j);
      // The S row arrives in eight 16-column pieces, four in flight: every tcgen05.wait::ld below only waits for a load
      // issued a whole compute block earlier, and it is a scheduling fence — ptxas cannot hoist the next block's
      // exponentials over it, so MUFU bursts stay 16 instructions long and the previous block's packing / row-sum
      // arithmetic fills the gaps (with the half-row loads of the first v4 ptxas issued 54 MUFU back to back and a
      // warpgroup alone on the SM sub-partition ran the MUFU pipe at half rate).
      uint32_t v[128];
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld16(TS + c * 16, v + c * 16);
      bool exact = (j == 0) || (ragged && j == nkv - 1);
      auto release_s = [&]() {                  // the whole row is in registers: S_t(j+1) may overwrite the buffer
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_free[t]);
        if (q == 0 && lane == 0) ATR(2 + t, 13, j);
      };
      if (exact) {
#pragma unroll
        for (int c = 4; c < 8; ++c) tmem_ld16(TS + c * 16, v + c * 16);
        tmem_ld_wait();
        release_s();
      }
      bool waited_pv = (j == 0);            // P_t / O_t are only touched after PV_t(j-1) has retired
      float lsum = 0.f;
      if (!exact) {
        // ---------- steady state: probabilities straight from the stale integer reference ----------
        const float neg_m = -m_used;
        const float cm = 12582912.0f - m_used;                     // exact: m_used is an integer, |m_used| < 2^21
        const uint64_t cm_2 = pack_f32x2(cm, cm);
        const float smin = (m_used - 125.0f) * inv_sl2;            // scores below this contribute < 2^-125: clamp
        const uint64_t neg1_2 = pack_f32x2(-1.0f, -1.0f);
        uint64_t lsA = 0ull, lsB = 0ull;
        float tmax = 0.f;
        uint32_t pk[32];                       // pieces 0..3 are held back until PV_t(j-1) has retired
        uint32_t tok[2] = {0u, 0u};
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          tmem_ld_wait();                      // pieces <= c + 3 have landed
          if (c + 4 < 8) tmem_ld16(TS + (c + 4) * 16, v + (c + 4) * 16);
          if (c == 4) release_s();
          // scheduling token: the exponent arguments of piece c formally depend on a packed result of piece c - 2
          // (x * 0 + (-m) is not foldable in IEEE arithmetic), so ptxas cannot pull all MUFU.EX2 of the tile into one
          // burst ahead of the packing / row-sum work — two interleaved chains keep the MUFU queue fed instead
          const float negm_c = (TOKEN && c >= 2) ? fmaf(__uint_as_float(tok[c & 1]), 0.0f, neg_m) : neg_m;
          const uint64_t negm_c2 = pack_f32x2(negm_c, negm_c);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int g = c * 8 + i;
            float p0, p1;
            if (a3_emu_pair(i, EMU8)) {
              const float s0 = fmaxf(__uint_as_float(v[2 * g]), smin), s1 = fmaxf(__uint_as_float(v[2 * g + 1]), smin);
              const uint64_t s2 = pack_f32x2(s0, s1);
              const uint64_t t2 = fma_f32x2(s2, sl2_2, cm_2);        // magic + round(x), x = s*c - m
              const uint64_t d2 = fma_f32x2(t2, neg1_2, cm_2);       // -m - round(x)   (exact)
              const uint64_t r2 = fma_f32x2(s2, sl2_2, d2);          // x - round(x) in [-0.5, 0.5]
              uint64_t q2 = fma_f32x2(pack_f32x2(0.05517132207751274f, 0.05517132207751274f), r2,
                                      pack_f32x2(0.24261054396629333f, 0.24261054396629333f));
              q2 = fma_f32x2(q2, r2, pack_f32x2(0.6932609677314758f, 0.6932609677314758f));
              q2 = fma_f32x2(q2, r2, pack_f32x2(0.9999281167984009f, 0.9999281167984009f));
              float qa, qb, ta, tb;
              unpack_f32x2(q2, qa, qb);
              unpack_f32x2(t2, ta, tb);
              tmax = fmaxf(tmax, fmaxf(ta, tb));
              p0 = __uint_as_float(__float_as_uint(qa) + (__float_as_uint(ta) << 23));
              p1 = __uint_as_float(__float_as_uint(qb) + (__float_as_uint(tb) << 23));
            } else {
              const uint64_t x2 = fma_f32x2(pack_f32x2(__uint_as_float(v[2 * g]), __uint_as_float(v[2 * g + 1])),
                                            sl2_2, negm_c2);
              float x0, x1;
              unpack_f32x2(x2, x0, x1);
              p0 = ex2_approx(x0);
              p1 = ex2_approx(x1);
            }
            if (i & 1) lsB = add_f32x2(lsB, pack_f32x2(p0, p1));
            else lsA = add_f32x2(lsA, pack_f32x2(p0, p1));
            pk[(c & 3) * 8 + i] = pack_bf16x2(p0, p1);
          }
          tok[c & 1] = pk[(c & 3) * 8 + 7];
          if (c == 3) {                        // PV_t(j-1) was issued at the end of tile j-1: half a tile to retire
            if (!waited_pv) {
              if (q == 0 && lane == 0) ATR(2 + t, 11, j);
              mbar_wait(&pv_done[t], (j - 1) & 1);
              tc_fence_after();
              waited_pv = true;
              if (q == 0 && lane == 0) ATR(2 + t, 12, j);
            }
            tmem_st16(TP, pk);
            tmem_st16(TP + 16, pk + 16);
          } else if (c == 5) {
            tmem_st16(TP + 32, pk);
          } else if (c == 7) {
            tmem_st16(TP + 48, pk + 16);
          }
        }
        float a0, a1, b0, b1;
        unpack_f32x2(lsA, a0, a1);
        unpack_f32x2(lsB, b0, b1);
        lsum = (a0 + a1) + (b0 + b1);
        // a score more than 64 log2 units above m_used (or a NaN) shows up in the row sum / the largest shifted
        // argument: redo this tile exactly
        exact = __any_sync(0xffffffffu, !(lsum <= A3_SUM_LIMIT) || tmax > A3_ARG_LIMIT);
        if (exact) tmem_st_wait();
      }
      float alpha = 1.0f;
      if (exact) {
        // ---------- exact path: row max, lazy rescale of O, masked tail (first / last / overflowing tiles) ----------
        const int kbase = j * 128;
        if (kbase + 128 > p.rows) {
#pragma unroll
          for (int i = 0; i < 128; ++i)
            if (kbase + i >= p.rows) v[i] = 0xff800000u;   // -inf
        }
        float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int i = 0; i < 128; i += 8) {
#pragma unroll
          for (int u = 0; u < 4; ++u)
            m4[u] = fmaxf(m4[u], fmaxf(__uint_as_float(v[i + 2 * u]), __uint_as_float(v[i + 2 * u + 1])));
        }
        const float mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
        // integer reference, clamped so that magic - m_used stays exact (|scores| beyond 2^20 log2 units are not
        // representable differences anyway)
        const float m_new = fmaxf(m_used, fminf(fmaxf(ceilf(mx * sl2), -1048576.0f), 1048576.0f));
        const bool need = (m_new - m_used) > ATT_LAZY_THRESHOLD;     // first tile: inf > 8
        const bool warp_need = __any_sync(0xffffffffu, need);
        if (need) {
          alpha = ex2_approx(m_used - m_new);
          m_used = m_new;
        }
        if (!waited_pv) {
          mbar_wait(&pv_done[t], (j - 1) & 1);
          tc_fence_after();
          waited_pv = true;
        }
        if (j > 0 && warp_need) {            // O *= alpha (TMEM read-modify-write; rare after the first tiles)
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint32_t o[32];
            tmem_ld32(TO + c * 32, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st32(TO + c * 32, o);
          }
        }
        const float neg_m = -m_used;
        const uint64_t negm_2 = pack_f32x2(neg_m, neg_m);
        uint64_t lsA = 0ull, lsB = 0ull;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int g = c * 16 + i;
            const uint64_t x2 = fma_f32x2(pack_f32x2(__uint_as_float(v[2 * g]), __uint_as_float(v[2 * g + 1])), sl2_2,
                                          negm_2);
            float x0, x1;
            unpack_f32x2(x2, x0, x1);
            const float p0 = ex2_approx(x0), p1 = ex2_approx(x1);
            if (i & 1) lsB = add_f32x2(lsB, pack_f32x2(p0, p1));
            else lsA = add_f32x2(lsA, pack_f32x2(p0, p1));
            pk[i] = pack_bf16x2(p0, p1);
          }
          tmem_st16(TP + c * 16, pk);
        }
        float a0, a1, b0, b1;
        unpack_f32x2(lsA, a0, a1);
        unpack_f32x2(lsB, b0, b1);
        lsum = (a0 + a1) + (b0 + b1);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[t]);
      if (q == 0 && lane == 0) ATR(2 + t, 14, j);
      l_run = l_run * alpha + lsum;
    }
    if (q == 0 && lane == 0) ATR_END(2 + t);
    mbar_wait(&pv_done[t], (nkv - 1) & 1);
    tc_fence_after();
    const int row = q0 + t * 128 + r;
    const float inv = 1.0f / l_run;
    uint4* op = reinterpret_cast<uint4*>(p.out + static_cast<long long>(row) * (p.heads * 64) + head * 64);
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t o[32];
      tmem_ld32(TO + c * 32, o);
      tmem_ld_wait();
      if (row < p.rows) {
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          uint4 w;
          w.x = pack_bf16x2(__uint_as_float(o[jj * 8 + 0]) * inv, __uint_as_float(o[jj * 8 + 1]) * inv);
          w.y = pack_bf16x2(__uint_as_float(o[jj * 8 + 2]) * inv, __uint_as_float(o[jj * 8 + 3]) * inv);
          w.z = pack_bf16x2(__uint_as_float(o[jj * 8 + 4]) * inv, __uint_as_float(o[jj * 8 + 5]) * inv);
          w.w = pack_bf16x2(__uint_as_float(o[jj * 8 + 6]) * inv, __uint_as_float(o[jj * 8 + 7]) * inv);
          op[c * 4 + jj] = w;
        }
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

// Measured dead ends on B200 (round 2, N = 19 426, 48 heads; shipping v4 with 1/8 emulation = 5.05-5.4 ms depending on the
// box, 5.19 ms / tensor pipe 42 % / XU 74 % under ncu):
//   * 64-key tiles with double-buffered S and P in TMEM (so that S(j+1) is always ready): 6.9-7.0 ms — every per-tile
//     fixed cost (mbarrier hand-offs, tcgen05.wait::st, fences) doubles and the N = 64 QK^T MMA is shared-memory bound;
//   * holding the whole S row in registers at 168 registers per thread (v3 without setmaxnreg): 6.3-6.5 ms, spills;
//   * more than 1/8 of the exponentials on the FMA pipe: slower in v3 and v4 (the softmax warps are in-order and
//     latency-bound: every emulated pair is 11 dependent-ish instructions against 2 MUFU issue slots);
//   * ONE MMA issuer thread for both query tiles (fixed order or polling four barriers): 5.4-10 ms, erratic — the issuer
//     needs ~450 clk per action and becomes the critical path; `lane == 0` instead of elect.sync made it worse still;
//   * loading the whole S row before the first exponential: 6.4 ms (both warpgroups sit in the TMEM read together);
//   * SIXTEEN softmax warps (each query row split between two threads, half-row maxima exchanged through shared memory
//     and a 64-thread named barrier every tile, 96/104 registers): correct (142 attention tests) but 5.84-5.94 ms, XU
//     69 % — 36 % more instructions (row-max pass, exchange) and the MUFU.EX2 stalls move from `wait` to `mio_throttle`;
//     four warps per sub-partition do not keep the special-function unit busier than two;
//   * nanosleep back-off in the TMA / MMA threads' mbarrier spins (to take their SYNCS out of the MIO queue): no change;
//   * v4 without the scheduling token: 5.40 ms (with: 5.06); v4 with a 40 / 232 register split: 5.53 ms (56 / 224: 5.06 —
//     at 40 registers the MMA issuer threads spill and react late);
//   * packed ex2.approx.bf16x2 / f16x2: two MUFU ops per instruction in SASS, same 8 clk each (profiles/microbench_sm.cu).
// The special-function unit sustains ~74 % of its 16 results/clk/SM in this instruction mix in every variant that keeps
// the hand-offs off the critical path; at head_dim 64 that caps the tensor pipe near 42-45 %.
// variant: -1 = auto (v4 with 1/8 emulation from 3 000 rows — on par with v2 at 3 216 / 4 978, ahead above —, v2 below,
// where its 2x finer CTA grain wins);
// 0 = v2 (2 CTAs/SM, one query tile each); 1 + e (e = 0..5) = v3 with e/8 of the exponentials on the FMA pipe;
// 7 + e (e = 0..4) = v4 (S row in registers via setmaxnreg, two MMA issuer warps) with e/8 emulated.
// In-step A/B at cfg-2 (42 layers, N = 19 426, power-capped clocks ~1.45 GHz): v3 289.6 ms, v4 266.9 ms per clip.
static std::atomic<int> g_attn_variant{-1};
int set_attn_variant(int v) {
  if (v < -1 || v > 11) return set_error(DOVE_E_BAD_ARG, "attn_variant must be -1..11");
  g_attn_variant.store(v);
  return DOVE_OK;
}

template <typename K>
static int set_smem_once(K kernel, size_t bytes, std::once_flag& flag) {
  cudaError_t err = cudaSuccess;
  std::call_once(flag, [&] {
    err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes));
  });
  return err == cudaSuccess ? DOVE_OK : check_cuda(err, "cudaFuncSetAttribute(attention)");
}

template <int EMU8>
static int launch_v3(const CUtensorMap& tm, const AttnParams& p, cudaStream_t st) {
  static std::once_flag flag;
  if (int e = set_smem_once(attn_fwd_v3_kernel<EMU8>, A3_SMEM, flag)) return e;
  dim3 grid((p.rows + 255) / 256, p.heads);
  attn_fwd_v3_kernel<EMU8><<<grid, 320, A3_SMEM, st>>>(tm, p);
  DOVE_LAUNCH_CHECK("attn_fwd_v3_kernel");
  return DOVE_OK;
}

template <int EMU8, bool TOKEN = true, int REG_MISC = 56, int REG_SOFTMAX = 224>
static int launch_v4(const CUtensorMap& tm, const AttnParams& p, cudaStream_t st) {
  static std::once_flag flag;
  if (int e = set_smem_once(attn_fwd_v4_kernel<EMU8, TOKEN, REG_MISC, REG_SOFTMAX>, A4_SMEM, flag)) return e;
  dim3 grid((p.rows + 255) / 256, p.heads);
  attn_fwd_v4_kernel<EMU8, TOKEN, REG_MISC, REG_SOFTMAX><<<grid, 384, A4_SMEM, st>>>(tm, p);
  DOVE_LAUNCH_CHECK("attn_fwd_v4_kernel");
  return DOVE_OK;
}

static int attention_launch(const void* qkv, void* out, int rows, int heads, float scale, cudaStream_t st) {
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
  int variant = g_attn_variant.load();
  if (variant < 0) variant = rows >= 3000 ? 8 : 0;
  switch (variant) {
    case 1: return launch_v3<0>(tm, p, st);
    case 2: return launch_v3<1>(tm, p, st);
    case 3: return launch_v3<2>(tm, p, st);
    case 4: return launch_v3<3>(tm, p, st);
    case 5: return launch_v3<4>(tm, p, st);
    case 6: return launch_v3<5>(tm, p, st);
    case 7: return launch_v4<0>(tm, p, st);
    case 8: return launch_v4<1>(tm, p, st);
    case 9: return launch_v4<2>(tm, p, st);
    case 10: return launch_v4<3>(tm, p, st);
    case 11: return launch_v4<4>(tm, p, st);
    default: break;
  }
  static std::once_flag flag;
  if (int e = set_smem_once(attn_fwd_v2_kernel, ATT_SMEM, flag)) return e;
  dim3 grid((rows + 127) / 128, heads);
  attn_fwd_v2_kernel<<<grid, 192, ATT_SMEM, st>>>(tm, p);
  DOVE_LAUNCH_CHECK("attn_fwd_v2_kernel");
  return DOVE_OK;
}

}  // namespace dove

extern "C" int dove_attention_bf16(const void* qkv, void* out, int rows, int heads, float scale, void* stream) {
  return dove::attention_launch(qkv, out, rows, heads, scale, static_cast<cudaStream_t>(stream));
}
