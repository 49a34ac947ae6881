// Persistent warp-specialised tcgen05 GEMM for sm_100a, shared by
//   * dense GEMM  C[M,N] = epi(A[M,K] W[N,K]^T + bias)           (DiT linears, VAE 1x1x1 convs)
//   * implicit-GEMM convolution on channels-last activations      (CausalConv3d / Conv2d of the VAE)
// The two differ only in the TMA producer: dense loads a [128 x 64] box of A; conv loads, for each filter
// tap, the shifted [th x tw x 64ch] window of the activation tensor (4-D tensor map, OOB zero fill = spatial
// zero padding) — no im2col buffer ever exists.
//
// CTA = 320 threads: warp 0 = TMA producer, warp 1 = TMEM owner + single-thread MMA issuer,
// warps 2..9 = epilogue (TMEM -> registers -> fused epilogue -> global; two warps per TMEM lane quarter, each
// taking every other 32-column chunk, so the epilogue has 2 warps per SMSP to hide its own latencies).  Accumulators are double-buffered in
// TMEM (2 x BN fp32 columns) so the epilogue of tile i overlaps the main loop of tile i+1.
#include <mutex>

#include "gemm_common.cuh"

namespace dove {

int conv2cta_dispatch(const void* x, const void* x_prev, int Tin, const void* w, int Tout, int Hin, int Win, int Cin,
                      int Cout_pad, int kt, int Ho, int Wo, GemmParams p, cudaStream_t st);
int conv_trans_halo_dispatch(const void* x, const void* x_prev, int Tin, const void* w, int Tout, int Hin, int Win, int Cin,
                             int Cout_pad, int kt, int Ho, int Wo, GemmParams p, cudaStream_t st);
int conv_narrow_dispatch(const void* x, const void* x_prev, int Tin, const void* w, int Tout, int H, int W, int Cin,
                         GemmParams p, cudaStream_t st);
int get_option_conv2cta();
int gemm2cta_launch(const void* A, long long lda, const void* W, long long ldw, int M, int N, int K, GemmParams p,
                    cudaStream_t st);

template <int BN>
struct GemmCfg {
  static constexpr int STAGES = (BN == 256) ? 4 : 6;
  static constexpr uint32_t A_BYTES = 128 * 128;
  static constexpr uint32_t B_BYTES = BN * 128;
  static constexpr uint32_t TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128
                                        : (2 * BN <= 256) ? 256 : 512;
  static constexpr int CH = (BN >= 32) ? 32 : 16;
  static constexpr size_t SMEM = 1024 + STAGES * (A_BYTES + B_BYTES) + 256;
};

// kTrans (conv only, BN = 256): operand roles are swapped — the M = 128 operand is the WEIGHT tile (128 output
// channels), the N = 256 operand is the activation box (256 voxels) — so that 128-channel convs still issue N = 256
// MMAs (an M128 x N128 MMA re-reads 128 B/clk of operands and saturates shared memory at ~42 % tensor pipe).
// The accumulator is then [channel lane][voxel column] and the epilogue writes it back channels-last.
template <int BN, bool kConv, bool kTrans = false>
__global__ void __launch_bounds__(320, 1)
umma_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmP, const GemmParams p) {
  using Cfg = GemmCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int CH = Cfg::CH;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * Cfg::A_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(sB + STAGES * Cfg::B_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty + 2);

  __shared__ float gn_red[8 * 64];                  // per-epilogue-warp GroupNorm partials [warp][group][2] (convs)
  if (kConv)
    for (int i = threadIdx.x; i < 8 * 64; i += blockDim.x) gn_red[i] = 0.f;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_tiles = p.num_m_tiles * p.num_n_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 8);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<Cfg::TMEM_COLS>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0 && elect_one()) {
    // ===================== TMA producer =====================
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int mt = tile / p.num_n_tiles, nt = tile % p.num_n_tiles;
      int t = 0, h0 = 0, w0 = 0;
      if (kConv) {
        int hy, wx;
        conv_tile_coords(p, mt, t, hy, wx);
        h0 = hy * p.th;
        w0 = wx * p.tw;
      }
      for (int kb = 0; kb < p.num_kb; ++kb) {
        mbar_wait(&empty[stage], phase ^ 1);
        mbar_expect_tx(&full[stage], Cfg::A_BYTES + Cfg::B_BYTES);
        if (kConv) {
          const int tap = kb / p.cin_blocks;
          const int c0 = (kb - tap * p.cin_blocks) * 64;
          const int dw = tap % p.kw;
          const int dh = (tap / p.kw) % p.kh;
          const int dt = tap / (p.kw * p.kh);
          int f = t + dt;
          const CUtensorMap* src = conv_frame_src(p, &tmA, &tmP, f);
          tma_load_4d(kTrans ? sB + stage * Cfg::B_BYTES : sA + stage * Cfg::A_BYTES, src, &full[stage], c0,
                      w0 * p.stride + dw - p.pad, h0 * p.stride + dh - p.pad, f);
        } else {
          tma_load_2d(sA + stage * Cfg::A_BYTES, &tmA, &full[stage], kb * 64, mt * 128);
        }
        if (kTrans) tma_load_2d(sA + stage * Cfg::A_BYTES, &tmB, &full[stage], kb * 64, nt * 128);
        else tma_load_2d(sB + stage * Cfg::B_BYTES, &tmB, &full[stage], kb * 64, nt * BN);
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1 && elect_one()) {
    // ===================== MMA issuer (one thread) =====================
    constexpr uint32_t idesc = umma_idesc_bf16(128, BN, 0, 0);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      mbar_wait(&tempty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int kb = 0; kb < p.num_kb; ++kb) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint64_t adesc = umma_desc_sw128(smem_u32(sA + stage * Cfg::A_BYTES));
        const uint64_t bdesc = umma_desc_sw128(smem_u32(sB + stage * Cfg::B_BYTES));
#pragma unroll
        for (int k = 0; k < 4; ++k)   // 4 x (K = 16) per 64-wide k-block; +32 B per step inside the swizzle atom
          umma_ss(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
        umma_commit(&empty[stage]);
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
      umma_commit(&tfull[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  } else if (warp >= 2) {
    // ===================== epilogue warps =====================
    const int q = warp & 3;               // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;     // two warps per quarter: even / odd column chunks
    const int r_in_tile = q * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    float gn_s = 0.f, gn_ss = 0.f;          // kTrans: this thread's channel-pair sums over every voxel it stores
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int mt = tile / p.num_n_tiles, nt = tile % p.num_n_tiles;
      if (kTrans) {
        // lane = output channel, accumulator columns = the tile's 256 voxels (th rows x tw cols, tw a power of 2)
        int t, hy, wx;
        conv_tile_coords(p, mt, t, hy, wx);
        const int ch = nt * 128 + r_in_tile;
        const float bias_v = p.bias ? __bfloat162float(p.bias[ch]) : 0.f;
        const int tw_mask = p.tw - 1, tw_shift = 31 - __clz(p.tw);
        mbar_wait(&tfull[acc], acc_phase);
        tc_fence_after();
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN;
#pragma unroll 1
        for (int c0 = half * 32; c0 < BN; c0 += 64) {
          uint32_t v[32];
          tmem_ld32(t_row + c0, v);
          tmem_ld_wait();
          // voxels c0+i (even lanes store it) and c0+i+1 (odd lanes): lane pairs swap one value so that every lane
          // stores a packed (channel, channel+1) pair -> 64 contiguous bytes per voxel per half-warp.
          // All residual loads are issued BEFORE any store (C and aux may alias, so the compiler would otherwise
          // serialise load -> store -> load and expose one DRAM latency per voxel pair).
          long long offs[16];
          uint32_t res[16];
          bool oks[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int vi = c0 + 2 * i + (lane & 1);
            const int h = hy * p.th + (vi >> tw_shift), w = wx * p.tw + (vi & tw_mask);
            oks[i] = (h < p.Ho) && (w < p.Wo);
            const long long row = (static_cast<long long>(t) * p.Ho + h) * p.Wo + w;
            offs[i] = row * p.ldc + (ch & ~1);
            res[i] = 0;
            if (p.epi == DOVE_EPI_ADD && oks[i])
              res[i] = *reinterpret_cast<const uint32_t*>(p.aux + row * p.ld_aux + (ch & ~1));
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float a = bf16_round(__uint_as_float(v[2 * i]) + bias_v);
            const float b = bf16_round(__uint_as_float(v[2 * i + 1]) + bias_v);
            const float send = (lane & 1) ? a : b;
            const float recv = __shfl_xor_sync(0xffffffffu, send, 1);
            float lo = (lane & 1) ? recv : a;     // channel (ch & ~1)
            float hi = (lane & 1) ? b : recv;     // channel (ch | 1)
            const float2 x2 = unpack_bf16x2(res[i]);
            lo += x2.x;
            hi += x2.y;
            if (oks[i]) {
              const uint32_t packed = pack_bf16x2(lo, hi);
              *reinterpret_cast<uint32_t*>(p.C + offs[i]) = packed;
              const float2 fr = unpack_bf16x2(packed);
              gn_s += fr.x + fr.y;
              gn_ss += fr.x * fr.x + fr.y * fr.y;
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
        continue;
      }
      long long row;
      bool valid;
      if (kConv) {
        int t, hy, wx;
        conv_tile_coords(p, mt, t, hy, wx);
        const int h = hy * p.th + r_in_tile / p.tw;
        const int w = wx * p.tw + r_in_tile % p.tw;
        valid = (h < p.Ho) && (w < p.Wo);
        row = (static_cast<long long>(t) * p.Ho + h) * p.Wo + w;
      } else {
        row = static_cast<long long>(mt) * 128 + r_in_tile;
        valid = row < p.M;
      }
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN;
      const bf16* gate = (p.epi == DOVE_EPI_GATED_RES) ? (row < p.split_row ? p.gate0 : p.gate1) : nullptr;
#pragma unroll 1
      for (int c0 = half * CH; c0 < BN; c0 += 2 * CH) {
        uint32_t v[CH];
        if (CH == 32) tmem_ld32(t_row + c0, v); else tmem_ld16(t_row + c0, v);
        tmem_ld_wait();
        const int n0 = nt * BN + c0;
        if (kConv && CH == 32 && p.gn_partial) {      // fused GroupNorm statistics of the output (as in conv2cta)
          float qs[8], qss[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) qs[i] = qss[i] = 0.f;
          if (valid) epilogue_chunk<CH, true>(p, v, row, n0, gate, qs, qss);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
              qs[i] += __shfl_xor_sync(0xffffffffu, qs[i], o);
              qss[i] += __shfl_xor_sync(0xffffffffu, qss[i], o);
            }
          }
          if (lane == 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int g = (n0 + 4 * i) / p.gn_cpg;
              gn_red[(warp - 2) * 64 + g * 2] += qs[i];
              gn_red[(warp - 2) * 64 + g * 2 + 1] += qss[i];
            }
          }
        } else if (valid) {
          epilogue_chunk<CH>(p, v, row, n0, gate);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (kTrans && p.gn_partial) {          // 4 lanes = one group of 4 channels (C = 128, 32 groups)
      gn_s += __shfl_xor_sync(0xffffffffu, gn_s, 1);
      gn_ss += __shfl_xor_sync(0xffffffffu, gn_ss, 1);
      gn_s += __shfl_xor_sync(0xffffffffu, gn_s, 2);
      gn_ss += __shfl_xor_sync(0xffffffffu, gn_ss, 2);
      if ((lane & 3) == 0) {
        const int g = q * 8 + (lane >> 2);
        gn_red[(warp - 2) * 64 + g * 2] = gn_s;
        gn_red[(warp - 2) * 64 + g * 2 + 1] = gn_ss;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (kConv && p.gn_partial && threadIdx.x < 64) {    // fixed-order combine of the 8 epilogue warps
    float a8 = 0.f;
#pragma unroll
    for (int w8 = 0; w8 < 8; ++w8) a8 += gn_red[w8 * 64 + threadIdx.x];
    p.gn_partial[static_cast<long long>(blockIdx.x) * 64 + threadIdx.x] = a8;
  }
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

template <int BN, bool kConv>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmP, const GemmParams& p,
                       cudaStream_t st) {
  using Cfg = GemmCfg<BN>;
  static std::once_flag attr_once;   // thread-safe one-time opt-in to > 48 KB dynamic shared memory
  cudaError_t attr_err = cudaSuccess;
  std::call_once(attr_once, [&] { attr_err = cudaFuncSetAttribute(umma_gemm_kernel<BN, kConv>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(Cfg::SMEM)); });
  if (attr_err != cudaSuccess) return check_cuda(attr_err, "cudaFuncSetAttribute(umma_gemm_kernel)");
  const int total = p.num_m_tiles * p.num_n_tiles;
  const int grid = total < num_sms() ? total : num_sms();
  umma_gemm_kernel<BN, kConv><<<grid, 320, Cfg::SMEM, st>>>(tmA, tmB, tmP, p);
  DOVE_LAUNCH_CHECK("umma_gemm_kernel");
  return DOVE_OK;
}

static int launch_conv_trans(const CUtensorMap& tmX, const CUtensorMap& tmW, const CUtensorMap& tmP, const GemmParams& p,
                             cudaStream_t st) {
  using Cfg = GemmCfg<256>;
  static std::once_flag attr_once;   // thread-safe one-time opt-in to > 48 KB dynamic shared memory
  cudaError_t attr_err = cudaSuccess;
  std::call_once(attr_once, [&] { attr_err = cudaFuncSetAttribute(umma_gemm_kernel<256, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(Cfg::SMEM)); });
  if (attr_err != cudaSuccess) return check_cuda(attr_err, "cudaFuncSetAttribute(umma_gemm_kernel trans)");
  const int total = p.num_m_tiles * p.num_n_tiles;
  const int grid = total < num_sms() ? total : num_sms();
  umma_gemm_kernel<256, true, true><<<grid, 320, Cfg::SMEM, st>>>(tmX, tmW, tmP, p);
  DOVE_LAUNCH_CHECK("umma_gemm_kernel<trans>");
  return DOVE_OK;
}

template <bool kConv>
static int dispatch_bn(int bn, const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmP,
                       const GemmParams& p, cudaStream_t st) {
  switch (bn) {
    case 256: return launch_gemm<256, kConv>(tmA, tmB, tmP, p, st);
    case 128: return launch_gemm<128, kConv>(tmA, tmB, tmP, p, st);
    case 64: return launch_gemm<64, kConv>(tmA, tmB, tmP, p, st);
    case 32: return launch_gemm<32, kConv>(tmA, tmB, tmP, p, st);
    case 16: return launch_gemm<16, kConv>(tmA, tmB, tmP, p, st);
  }
  return set_error(DOVE_E_BAD_ARG, "unsupported BN %d", bn);
}

static int pick_bn(int N) {
  if (N % 256 == 0) return 256;
  if (N % 128 == 0) return 128;
  if (N % 64 == 0) return 64;
  if (N % 32 == 0) return 32;
  return 16;
}

}  // namespace dove

using namespace dove;

extern "C" int dove_gemm_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, void* C, int64_t ldc, int M,
                              int N, int K, const void* bias, int epilogue, const void* aux, int64_t ld_aux,
                              const void* gate0, const void* gate1, int split_row, void* stream) {
  if (int e = ensure_init()) return e;
  DOVE_CHECK_ARG(M > 0 && N > 0 && K > 0, "gemm: empty problem M=%d N=%d K=%d", M, N, K);
  DOVE_CHECK_ARG(K % 64 == 0, "gemm: K=%d must be a multiple of 64", K);
  DOVE_CHECK_ARG(N % 16 == 0, "gemm: N=%d must be a multiple of 16", N);
  DOVE_CHECK_ARG(lda % 8 == 0 && ldw % 8 == 0 && ldc % 8 == 0, "gemm: leading dims must be multiples of 8");
  DOVE_CHECK_ARG((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(W) |
                  reinterpret_cast<uintptr_t>(C)) % 16 == 0, "gemm: pointers must be 16-byte aligned");
  DOVE_CHECK_ARG(epilogue >= 0 && epilogue <= 3, "gemm: bad epilogue %d", epilogue);
  if (epilogue == DOVE_EPI_GATED_RES)
    DOVE_CHECK_ARG(aux && gate0 && gate1 && ld_aux % 8 == 0, "gemm: gated residual needs aux, gate0, gate1");
  if (epilogue == DOVE_EPI_ADD) DOVE_CHECK_ARG(aux && ld_aux % 8 == 0, "gemm: add epilogue needs aux");
  const int bn = pick_bn(N);
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
    uint32_t box[2] = {64, static_cast<uint32_t>(bn)};
    if (int e = make_tmap_bf16(&tmB, W, 2, dims, strides, box, nullptr)) return e;
  }
  GemmParams p{};
  p.num_m_tiles = (M + 127) / 128;
  p.num_n_tiles = N / bn;
  p.num_kb = K / 64;
  p.M = M;
  p.epi = epilogue;
  p.C = static_cast<bf16*>(C);
  p.ldc = ldc;
  p.bias = static_cast<const bf16*>(bias);
  p.aux = static_cast<const bf16*>(aux);
  p.ld_aux = ld_aux;
  p.gate0 = static_cast<const bf16*>(gate0);
  p.gate1 = static_cast<const bf16*>(gate1);
  p.split_row = split_row;
  p.n_valid = N;
  p.out_mode = 0;
  p.rows_total = M;
  if (get_option_conv2cta() != 0 && N % 256 == 0 && M >= 2048)   // large GEMMs: CTA-pair kernel, 256 x 256 tiles
    return gemm2cta_launch(A, lda, W, ldw, M, N, K, p, static_cast<cudaStream_t>(stream));
  return dispatch_bn<false>(bn, tmA, tmB, tmA, p, static_cast<cudaStream_t>(stream));
}

// cached != 0: causal 3x3x3 conv on the UN-padded frame batch x [Tout,H,W,Cin]; the two preceding frames come from
// x_prev [2,H,W,Cin] (the conv cache) or, when x_prev is NULL, frame 0 is replicated (first frame batch).
static int conv_impl(const void* x, const void* x_prev, int cached, const void* w, const void* bias, void* y, int Tout,
                     int Hin, int Win, int Cin, int Cout_pad, int cout_valid, int64_t ldy, int kt, int kh, int kw,
                     int stride, int pad, int Ho, int Wo, int epilogue, const void* aux, int64_t ld_aux, int out_mode,
                     float* gn_partial, int* gn_done, void* stream) {
  if (int e = ensure_init()) return e;
  if (gn_done) *gn_done = 0;
  DOVE_CHECK_ARG(!cached || (kt == 3 && stride == 1), "conv: cached mode is for causal kt = 3, stride 1 convs");
  const int t_shift = cached ? kt - 1 : 0;
  const int has_prev = (cached && x_prev) ? 1 : 0;
  const int Tin_all = cached ? Tout : Tout + kt - 1;      // frames present in x
  auto make_prev_map = [&](CUtensorMap* tm, const uint32_t* box, const uint32_t* es) -> int {
    uint64_t dims[4] = {static_cast<uint64_t>(Cin), static_cast<uint64_t>(Win), static_cast<uint64_t>(Hin), 2};
    uint64_t strides[3] = {static_cast<uint64_t>(Cin) * 2, static_cast<uint64_t>(Win) * Cin * 2,
                           static_cast<uint64_t>(Hin) * Win * Cin * 2};
    return make_tmap_bf16(tm, x_prev, 4, dims, strides, box, es);
  };
  DOVE_CHECK_ARG(Tout > 0 && Hin > 0 && Win > 0 && Ho > 0 && Wo > 0, "conv: empty problem");
  DOVE_CHECK_ARG(Cin % 64 == 0, "conv: Cin=%d must be a multiple of 64 (pad channels)", Cin);
  DOVE_CHECK_ARG(Cout_pad % 16 == 0 && cout_valid <= Cout_pad, "conv: Cout_pad=%d must be a multiple of 16", Cout_pad);
  DOVE_CHECK_ARG(stride == 1 || stride == 2, "conv: stride must be 1 or 2");
  DOVE_CHECK_ARG(epilogue == DOVE_EPI_BIAS || epilogue == DOVE_EPI_ADD, "conv: epilogue must be BIAS or ADD");
  DOVE_CHECK_ARG(out_mode >= 0 && out_mode <= 3, "conv: bad out_mode");
  if (out_mode == 0) DOVE_CHECK_ARG(ldy % 8 == 0, "conv: ldy must be a multiple of 8");
  if (out_mode >= 1) DOVE_CHECK_ARG(ldy >= static_cast<long long>(Tout) * Ho * Wo, "conv: planar plane stride too small");
  if (epilogue == DOVE_EPI_ADD) DOVE_CHECK_ARG(aux && ld_aux % 8 == 0, "conv: add epilogue needs aux");
  {   // CTA-pair kernel with in-smem reuse of the W taps for the big stride-1 3x3(x3) convs on wide images
    const int opt = get_option_conv2cta();
    if (opt != 0 && stride == 1 && kh == 3 && kw == 3 && pad == 1 && out_mode == 0 && Wo >= 256 && Hin == Ho &&
        Win == Wo && Cout_pad % 256 == 0 && cout_valid == Cout_pad) {
      GemmParams q{};
      q.Ho = Ho;
      q.Wo = Wo;
      q.kh = kh;
      q.kw = kw;
      q.stride = 1;
      q.pad = 1;
      q.epi = epilogue;
      q.C = static_cast<bf16*>(y);
      q.ldc = ldy;
      q.bias = static_cast<const bf16*>(bias);
      q.aux = static_cast<const bf16*>(aux);
      q.ld_aux = ld_aux;
      q.n_valid = cout_valid;
      q.out_mode = 0;
      q.rows_total = static_cast<long long>(Tout) * Ho * Wo;
      q.t_shift = t_shift;
      q.has_prev = has_prev;
      if (gn_partial && gn_done && cout_valid % 32 == 0) {
        if (int e = check_cuda(cudaMemsetAsync(gn_partial, 0, sizeof(float) * GN_PARTIAL_ROWS * 64,
                                               static_cast<cudaStream_t>(stream)), "gn partial memset")) return e;
        q.gn_partial = gn_partial;
        q.gn_cpg = cout_valid / 32;
        *gn_done = 1;
      }
      return conv2cta_dispatch(x, has_prev ? x_prev : nullptr, Tin_all, w, Tout, Hin, Win, Cin, Cout_pad, kt, Ho, Wo, q,
                               static_cast<cudaStream_t>(stream));
    }
  }
  // narrow outputs (decoder conv_out, 128 -> 3): resident weights + (dh, dw) tap reuse out of one halo box (conv3.cu)
  if (get_option_conv2cta() != 0 && Cout_pad == 16 && kt == 3 && kh == 3 && kw == 3 && stride == 1 && pad == 1 &&
      Cin <= 128 && Wo >= 128 && Hin == Ho && Win == Wo && (epilogue == DOVE_EPI_BIAS || out_mode == 0)) {
    GemmParams q{};
    q.Ho = Ho;
    q.Wo = Wo;
    q.kh = kh;
    q.kw = kw;
    q.stride = 1;
    q.pad = 1;
    q.t_shift = t_shift;
    q.has_prev = has_prev;
    q.epi = epilogue;
    q.C = static_cast<bf16*>(y);
    q.ldc = ldy;
    q.bias = static_cast<const bf16*>(bias);
    q.aux = static_cast<const bf16*>(aux);
    q.ld_aux = ld_aux;
    q.n_valid = cout_valid;
    q.out_mode = out_mode;
    q.rows_total = out_mode >= 1 ? ldy : static_cast<long long>(Tout) * Ho * Wo;
    return conv_narrow_dispatch(x, has_prev ? x_prev : nullptr, Tin_all, w, Tout, Hin, Win, Cin, q,
                                static_cast<cudaStream_t>(stream));
  }
  // 128-channel-out convs on large images: swapped operand roles (weights = M 128, voxels = N 256)
  const bool trans = get_option_conv2cta() != 0 && stride == 1 && out_mode == 0 && Cout_pad == 128 &&
                     cout_valid == 128 && static_cast<long long>(Ho) * Wo >= 4096;
  if (trans && get_option_conv2cta() == 1 && kh == 3 && kw == 3 && pad == 1 && Hin == Ho && Win == Wo && Wo >= 256 &&
      (kt == 1 || kt == 3)) {
    // one image row of 256 voxels per tile + W-tap reuse out of a 258-voxel halo box (conv4.cu)
    GemmParams q{};
    q.Ho = Ho;
    q.Wo = Wo;
    q.kh = 3;
    q.kw = 3;
    q.stride = 1;
    q.pad = 1;
    q.t_shift = t_shift;
    q.has_prev = has_prev;
    if (gn_partial && gn_done) {
      if (int e = check_cuda(cudaMemsetAsync(gn_partial, 0, sizeof(float) * GN_PARTIAL_ROWS * 64,
                                             static_cast<cudaStream_t>(stream)), "gn partial memset")) return e;
      q.gn_partial = gn_partial;
      q.gn_cpg = 4;
      *gn_done = 1;
    }
    q.epi = epilogue;
    q.C = static_cast<bf16*>(y);
    q.ldc = ldy;
    q.bias = static_cast<const bf16*>(bias);
    q.aux = static_cast<const bf16*>(aux);
    q.ld_aux = ld_aux;
    q.n_valid = cout_valid;
    q.out_mode = 0;
    q.rows_total = static_cast<long long>(Tout) * Ho * Wo;
    return conv_trans_halo_dispatch(x, has_prev ? x_prev : nullptr, Tin_all, w, Tout, Hin, Win, Cin, Cout_pad, kt, Ho, Wo,
                                    q, static_cast<cudaStream_t>(stream));
  }
  if (trans) {
    int best_tw = 256;
    long long best_cost = -1;
    for (int tw = 256; tw >= 8; tw >>= 1) {
      const int th = 256 / tw;
      const long long cost = static_cast<long long>((Wo + tw - 1) / tw) * ((Ho + th - 1) / th);
      if (best_cost < 0 || cost < best_cost) {
        best_cost = cost;
        best_tw = tw;
      }
    }
    const int tw = best_tw, th = 256 / tw;
    const int Tin = Tin_all;
    CUtensorMap tmX, tmW, tmP;
    {
      uint64_t dims[4] = {static_cast<uint64_t>(Cin), static_cast<uint64_t>(Win), static_cast<uint64_t>(Hin),
                          static_cast<uint64_t>(Tin)};
      uint64_t strides[3] = {static_cast<uint64_t>(Cin) * 2, static_cast<uint64_t>(Win) * Cin * 2,
                             static_cast<uint64_t>(Hin) * Win * Cin * 2};
      uint32_t box[4] = {64, static_cast<uint32_t>(tw), static_cast<uint32_t>(th), 1};
      if (int e = make_tmap_bf16(&tmX, x, 4, dims, strides, box, nullptr)) return e;
      tmP = tmX;
      if (has_prev)
        if (int e = make_prev_map(&tmP, box, nullptr)) return e;
    }
    const int Ktot = kt * kh * kw * Cin;
    {
      uint64_t dims[2] = {static_cast<uint64_t>(Ktot), static_cast<uint64_t>(Cout_pad)};
      uint64_t strides[1] = {static_cast<uint64_t>(Ktot) * 2};
      uint32_t box[2] = {64, 128};
      if (int e = make_tmap_bf16(&tmW, w, 2, dims, strides, box, nullptr)) return e;
    }
    GemmParams q{};
    q.tw = tw;
    q.th = th;
    q.tiles_w = (Wo + tw - 1) / tw;
    q.tiles_h = (Ho + th - 1) / th;
    q.num_m_tiles = Tout * q.tiles_w * q.tiles_h;
    q.num_n_tiles = Cout_pad / 128;
    q.num_kb = Ktot / 64;
    q.To = Tout;
    {
      const double row_bytes = static_cast<double>(Win) * Cin * 2.0 * Tin * th;
      int bh = static_cast<int>(32.0e6 / row_bytes);
      if (bh < 1) bh = 1;
      if (bh > q.tiles_h) bh = q.tiles_h;
      q.band_h = bh;
    }
    q.Ho = Ho;
    q.Wo = Wo;
    q.kh = kh;
    q.kw = kw;
    q.cin_blocks = Cin / 64;
    q.stride = 1;
    q.pad = pad;
    q.t_shift = t_shift;
    q.has_prev = has_prev;
    if (gn_partial && gn_done) {
      if (int e = check_cuda(cudaMemsetAsync(gn_partial, 0, sizeof(float) * GN_PARTIAL_ROWS * 64,
                                             static_cast<cudaStream_t>(stream)), "gn partial memset")) return e;
      q.gn_partial = gn_partial;
      q.gn_cpg = 4;
      *gn_done = 1;
    }
    q.epi = epilogue;
    q.C = static_cast<bf16*>(y);
    q.ldc = ldy;
    q.bias = static_cast<const bf16*>(bias);
    q.aux = static_cast<const bf16*>(aux);
    q.ld_aux = ld_aux;
    q.n_valid = cout_valid;
    q.out_mode = 0;
    q.rows_total = static_cast<long long>(Tout) * Ho * Wo;
    return launch_conv_trans(tmX, tmW, tmP, q, static_cast<cudaStream_t>(stream));
  }
  const int bn = pick_bn(Cout_pad);
  // output tile geometry: tw x th = 128 voxels of one frame, tw a power of two minimising padded waste
  int best_tw = 128;
  long long best_cost = -1;
  for (int tw = 128; tw >= 8; tw >>= 1) {
    const int th = 128 / tw;
    if (stride == 2 && (tw * 2 > 256 || th * 2 > 256)) continue;
    const long long cost = static_cast<long long>((Wo + tw - 1) / tw) * ((Ho + th - 1) / th);
    if (best_cost < 0 || cost < best_cost) {
      best_cost = cost;
      best_tw = tw;
    }
  }
  const int tw = best_tw, th = 128 / tw;
  const int Tin = Tin_all;
  CUtensorMap tmA, tmB, tmP;
  {
    uint64_t dims[4] = {static_cast<uint64_t>(Cin), static_cast<uint64_t>(Win), static_cast<uint64_t>(Hin),
                        static_cast<uint64_t>(Tin)};
    uint64_t strides[3] = {static_cast<uint64_t>(Cin) * 2, static_cast<uint64_t>(Win) * Cin * 2,
                           static_cast<uint64_t>(Hin) * Win * Cin * 2};
    uint32_t box[4] = {64, static_cast<uint32_t>(tw * stride), static_cast<uint32_t>(th * stride), 1};
    uint32_t es[4] = {1, static_cast<uint32_t>(stride), static_cast<uint32_t>(stride), 1};
    if (int e = make_tmap_bf16(&tmA, x, 4, dims, strides, box, es)) return e;
    tmP = tmA;
    if (has_prev)
      if (int e = make_prev_map(&tmP, box, es)) return e;
  }
  const int Ktot = kt * kh * kw * Cin;
  {
    uint64_t dims[2] = {static_cast<uint64_t>(Ktot), static_cast<uint64_t>(Cout_pad)};
    uint64_t strides[1] = {static_cast<uint64_t>(Ktot) * 2};
    uint32_t box[2] = {64, static_cast<uint32_t>(bn)};
    if (int e = make_tmap_bf16(&tmB, w, 2, dims, strides, box, nullptr)) return e;
  }
  GemmParams p{};
  p.tw = tw;
  p.th = th;
  p.tiles_w = (Wo + tw - 1) / tw;
  p.tiles_h = (Ho + th - 1) / th;
  p.num_m_tiles = Tout * p.tiles_w * p.tiles_h;
  p.To = Tout;
  {   // band height: keep (band rows + halo) x all input frames of the band within ~32 MB of L2
    const double row_bytes = static_cast<double>(Win) * Cin * 2.0 * Tin * th * stride;
    int bh = static_cast<int>(32.0e6 / row_bytes);
    if (bh < 1) bh = 1;
    if (bh > p.tiles_h) bh = p.tiles_h;
    p.band_h = bh;
  }
  p.num_n_tiles = Cout_pad / bn;
  p.num_kb = Ktot / 64;
  p.Ho = Ho;
  p.Wo = Wo;
  p.kh = kh;
  p.kw = kw;
  p.cin_blocks = Cin / 64;
  p.stride = stride;
  p.pad = pad;
  p.t_shift = t_shift;
  p.has_prev = has_prev;
  p.epi = epilogue;
  p.C = static_cast<bf16*>(y);
  p.ldc = ldy;
  p.bias = static_cast<const bf16*>(bias);
  p.aux = static_cast<const bf16*>(aux);
  p.ld_aux = ld_aux;
  p.n_valid = cout_valid;
  p.out_mode = out_mode;
  p.rows_total = out_mode >= 1 ? ldy : static_cast<long long>(Tout) * Ho * Wo;
  // fused GroupNorm statistics on the generic kernel too (32-column epilogue chunks, >= 4 channels per group)
  if (gn_partial && gn_done && out_mode == 0 && bn >= 32 && cout_valid == Cout_pad && cout_valid % 128 == 0) {
    if (int e = check_cuda(cudaMemsetAsync(gn_partial, 0, sizeof(float) * GN_PARTIAL_ROWS * 64,
                                           static_cast<cudaStream_t>(stream)), "gn partial memset")) return e;
    p.gn_partial = gn_partial;
    p.gn_cpg = cout_valid / 32;
    *gn_done = 1;
  }
  return dispatch_bn<true>(bn, tmA, tmB, tmP, p, static_cast<cudaStream_t>(stream));
}

extern "C" int dove_conv_cl_bf16(const void* x, const void* w, const void* bias, void* y, int Tout, int Hin,
                                 int Win, int Cin, int Cout_pad, int cout_valid, int64_t ldy, int kt, int kh,
                                 int kw, int stride, int pad, int Ho, int Wo, int epilogue, const void* aux,
                                 int64_t ld_aux, int out_mode, float* gn_partial, int* gn_done, void* stream) {
  return conv_impl(x, nullptr, 0, w, bias, y, Tout, Hin, Win, Cin, Cout_pad, cout_valid, ldy, kt, kh, kw, stride, pad, Ho,
                   Wo, epilogue, aux, ld_aux, out_mode, gn_partial, gn_done, stream);
}

extern "C" int dove_conv3d_causal_bf16(const void* x, const void* x_prev, const void* w, const void* bias, void* y,
                                       int T, int H, int W, int Cin, int Cout_pad, int cout_valid, int64_t ldy,
                                       int epilogue, const void* aux, int64_t ld_aux, int out_mode, float* gn_partial,
                                       int* gn_done, void* stream) {
  return conv_impl(x, x_prev, 1, w, bias, y, T, H, W, Cin, Cout_pad, cout_valid, ldy, 3, 3, 3, 1, 1, H, W, epilogue, aux,
                   ld_aux, out_mode, gn_partial, gn_done, stream);
}
