// CTA-pair (cta_group::2) dense GEMM: C[M,N] = epi(A[M,K] W[N,K]^T + bias) with 256 x 256 pair tiles.
// One tcgen05.mma.cta_group::2 instruction spans two SMs (M = 256: CTA r owns rows r*128..r*128+127 of the pair
// tile) and N = 256; each SM stages its own 128 A rows and only HALF of the weight tile (128 of the 256 N rows), so
// shared-memory traffic per FLOP is 2/3 of the 1-CTA 128 x 256 tile and the tensor pipe is no longer operand-bound.
// Same roles as gemm.cu: warp 0 TMA producer (both CTAs), warp 1 TMEM owner (both) + MMA issuer (leader),
// warps 2..9 epilogue (both CTAs).  "full" barriers live in the leader; "empty" barriers are released in both CTAs by
// multicast tcgen05.commit; accumulators are double-buffered in TMEM (2 x 256 columns).
#include <mutex>

#include "gemm_common.cuh"

namespace dove {

struct Gemm2Cfg {
  static constexpr int STAGES = 6;
  static constexpr uint32_t A_BYTES = 128 * 128;
  static constexpr uint32_t B_BYTES = 128 * 128;     // half of the 256-row weight tile
  static constexpr size_t SMEM = 1024 + STAGES * (A_BYTES + B_BYTES) + 256;
};

// (320 threads get 168 registers each: the SM allocates registers for 12 warps — warps come in groups of four — so
// 200 registers x 320 threads fails to launch with "too many resources"; measured on B200.)
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(320, 1)
gemm2cta_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const GemmParams p) {
  using Cfg = Gemm2Cfg;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int BN = 256, CH = 32;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * Cfg::A_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(sB + STAGES * Cfg::B_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int total_tiles = p.num_m_tiles * p.num_n_tiles;     // m-tiles are 256-row pair tiles

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 16); }
    fence_mbar_init();
  }
  __syncwarp();
  if (warp == 1) tmem_alloc_2cta<512>(tmem_ptr);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0 && elect_one()) {
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = pair; tile < total_tiles; tile += npairs) {
      const int mt = tile / p.num_n_tiles, nt = tile % p.num_n_tiles;
      const int row0 = mt * 256 + static_cast<int>(rank) * 128;
      const int n0 = nt * BN + static_cast<int>(rank) * 128;
      for (int kb = 0; kb < p.num_kb; ++kb) {
        mbar_wait(&empty[stage], phase ^ 1);
        const uint32_t fb = mapa_shared(smem_u32(&full[stage]), 0);
        if (rank == 0) mbar_expect_tx(&full[stage], 2 * (Cfg::A_BYTES + Cfg::B_BYTES));
        tma_load_2d_2cta(sA + stage * Cfg::A_BYTES, &tmA, fb, kb * 64, row0);
        tma_load_2d_2cta(sB + stage * Cfg::B_BYTES, &tmB, fb, kb * 64, n0);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1 && rank == 0 && elect_one()) {
    constexpr uint32_t idesc = umma_idesc_bf16(256, BN, 0, 0);
    int stage = 0, acc = 0;
    uint32_t phase = 0, acc_phase = 0;
    for (int tile = pair; tile < total_tiles; tile += npairs) {
      mbar_wait(&tempty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int kb = 0; kb < p.num_kb; ++kb) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint64_t adesc = umma_desc_sw128(smem_u32(sA + stage * Cfg::A_BYTES));
        const uint64_t bdesc = umma_desc_sw128(smem_u32(sB + stage * Cfg::B_BYTES));
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_ss_2cta(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
        umma_commit_2cta(&empty[stage]);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      umma_commit_2cta(&tfull[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  } else if (warp >= 2) {
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int r_in_tile = q * 32 + lane;
    const uint32_t tempty_leader0 = mapa_shared(smem_u32(&tempty[0]), 0);
    const uint32_t tempty_leader1 = mapa_shared(smem_u32(&tempty[1]), 0);
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = pair; tile < total_tiles; tile += npairs) {
      const int mt = tile / p.num_n_tiles, nt = tile % p.num_n_tiles;
      const long long row = static_cast<long long>(mt) * 256 + static_cast<int>(rank) * 128 + r_in_tile;
      const bool valid = row < p.M;
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN;
      const bf16* gate = (p.epi == DOVE_EPI_GATED_RES) ? (row < p.split_row ? p.gate0 : p.gate1) : nullptr;
      if (p.epi == DOVE_EPI_QKV_NORM_ROPE) {
        // the 256-column tile holds 4 heads (64 columns each, all q, all k or all v since heads*64 % 256 == 0); the
        // two epilogue warps of a lane quarter take heads {0, 2} and {1, 3}, so one thread owns whole heads of its row
#pragma unroll 1
        for (int hh = half; hh < 4; hh += 2) {
          uint32_t v[64];
          tmem_ld32(t_row + hh * 64, v);
          tmem_ld32(t_row + hh * 64 + 32, v + 32);
          tmem_ld_wait();
          if (valid) qkv_head_epilogue(p, v, row, nt * BN + hh * 64);
        }
      } else {
#pragma unroll 1
        for (int c0 = half * CH; c0 < BN; c0 += 2 * CH) {
          uint32_t v[CH];
          tmem_ld32(t_row + c0, v);
          tmem_ld_wait();
          if (valid) epilogue_chunk<CH>(p, v, row, nt * BN + c0, gate);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(acc ? tempty_leader1 : tempty_leader0);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }
  __syncwarp();
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2cta<512>(tmem_base);
  }
}

int gemm2cta_launch(const void* A, long long lda, const void* W, long long ldw, int M, int N, int K, GemmParams p,
                    cudaStream_t st) {
  CUtensorMap tmA, tmB;
  {
    uint64_t dims[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(M)};
    uint64_t strides[1] = {static_cast<uint64_t>(lda) * 2};
    uint32_t box[2] = {64, 128};
    if (int e = make_tmap_bf16(&tmA, A, 2, dims, strides, box, nullptr)) return e;
  }
  {
    uint64_t dims[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(N)};
    uint64_t strides[1] = {static_cast<uint64_t>(ldw) * 2};
    uint32_t box[2] = {64, 128};
    if (int e = make_tmap_bf16(&tmB, W, 2, dims, strides, box, nullptr)) return e;
  }
  p.num_m_tiles = (M + 255) / 256;
  p.num_n_tiles = N / 256;
  p.num_kb = K / 64;
  static std::once_flag attr_once;   // thread-safe one-time opt-in to > 48 KB dynamic shared memory
  cudaError_t attr_err = cudaSuccess;
  std::call_once(attr_once, [&] { attr_err = cudaFuncSetAttribute(gemm2cta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(Gemm2Cfg::SMEM)); });
  if (attr_err != cudaSuccess) return check_cuda(attr_err, "cudaFuncSetAttribute(gemm2cta_kernel)");
  const int total = p.num_m_tiles * p.num_n_tiles;
  const int max_pairs = num_sms() / 2;
  const int pairs = total < max_pairs ? total : max_pairs;
  gemm2cta_kernel<<<pairs * 2, 320, Gemm2Cfg::SMEM, st>>>(tmA, tmB, p);
  DOVE_LAUNCH_CHECK("gemm2cta_kernel");
  return DOVE_OK;
}

}  // namespace dove

// QKV projection with the per-head q/k LayerNorm and the 3-D RoPE applied in the GEMM epilogue (one head = 64
// accumulator columns = one thread's registers), so the separate in-place pass over the [rows, 3*heads*64] buffer
// (dove_qk_norm_rope_bf16) disappears from the DiT block.
extern "C" int dove_gemm_qkv_norm_rope_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, void* C, int64_t ldc,
                                            int M, int heads, int K, const void* bias, const void* q_w, const void* q_b,
                                            const void* k_w, const void* k_b, float eps, const float* cos_t,
                                            const float* sin_t, int text_len, void* stream) {
  using namespace dove;
  if (int e = ensure_init()) return e;
  const int N = 3 * heads * 64;
  DOVE_CHECK_ARG(M > 0 && heads > 0 && K > 0 && K % 64 == 0, "gemm_qkv: bad shape M=%d heads=%d K=%d", M, heads, K);
  DOVE_CHECK_ARG((heads * 64) % 256 == 0, "gemm_qkv: heads*64 = %d must be a multiple of 256", heads * 64);
  DOVE_CHECK_ARG(lda % 8 == 0 && ldw % 8 == 0 && ldc % 8 == 0, "gemm_qkv: leading dims must be multiples of 8");
  DOVE_CHECK_ARG((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(W) | reinterpret_cast<uintptr_t>(C)) % 16 == 0,
                 "gemm_qkv: pointers must be 16-byte aligned");
  DOVE_CHECK_ARG(bias && q_w && q_b && k_w && k_b, "gemm_qkv: bias and norm_q / norm_k parameters required");
  DOVE_CHECK_ARG(text_len >= M || (cos_t && sin_t), "gemm_qkv: cos/sin tables required");
  GemmParams p{};
  p.M = M;
  p.epi = DOVE_EPI_QKV_NORM_ROPE;
  p.C = static_cast<bf16*>(C);
  p.ldc = ldc;
  p.bias = static_cast<const bf16*>(bias);
  p.n_valid = N;
  p.rows_total = M;
  p.qk_w[0] = static_cast<const bf16*>(q_w);
  p.qk_b[0] = static_cast<const bf16*>(q_b);
  p.qk_w[1] = static_cast<const bf16*>(k_w);
  p.qk_b[1] = static_cast<const bf16*>(k_b);
  p.rope_cos = cos_t;
  p.rope_sin = sin_t;
  p.rope_ld = M - text_len;
  p.text_len = text_len;
  p.heads = heads;
  p.qk_eps = eps;
  return gemm2cta_launch(A, lda, W, ldw, M, N, K, p, static_cast<cudaStream_t>(stream));
}
