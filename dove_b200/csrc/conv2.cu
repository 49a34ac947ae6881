// CTA-pair (cta_group::2) implicit-GEMM convolution with in-smem reuse of the W-direction filter taps.
//
// The 1-CTA kernel in gemm.cu is bounded by shared-memory traffic for the VAE's big 3x3(x3) convs: every filter
// tap re-stages a [128 voxel x 64 ch] activation box (16 KB TMA write) and the M128 x N128 MMA re-reads
// A (16 KB) + B (16 KB) per 64-wide k-block = 128 B/clk, so the tensor pipe sat at 42 % (profiles/r01_ncu_conv_igemm_bn128.txt).
// This kernel cuts both:
//   * CTA PAIR: one tcgen05.mma.cta_group::2 instruction computes M = 256 (two output rows h, h+1 of 128 voxels,
//     one per SM) x N = BN; each SM stages only HALF of the weight tile (BN/2 rows) and the pair shares it.
//   * kw REUSE: the tile is one image row of 128 consecutive voxels, so the three W-taps (dw = 0,1,2) of a
//     (dt, dh) tap group are the same smem box shifted by one voxel = one 128-byte row.  The producer stages ONE
//     [130 voxel x 64 ch] halo box per (dt, dh, channel block) and the MMA issuer forms the three taps with shifted
//     shared-memory descriptors (start address + dw*128 B; the swizzle follows absolute address bits).  A-staging traffic drops 3x.
// Per SM and k-block: TMA writes 5.5 KB (A) + BN/2*128 B (B), MMA reads 16 KB (A) + BN/2*128 B (B).
//
// CTA = 320 threads: warp 0 TMA producer (both CTAs), warp 1 TMEM owner (both) + MMA issuer (leader only),
// warps 2..9 epilogue (both CTAs, each on its own 128 accumulator lanes).  Separate A and B mbarrier rings; all
// "full" barriers live in the leader CTA (both CTAs' TMA bytes are credited there), "empty" barriers are released
// in both CTAs by multicast tcgen05.commit.
#include <mutex>

#include "gemm_common.cuh"

namespace dove {

template <int BN>
struct Conv2Cfg {
  static constexpr int SA = 4;                      // A (halo box) stages
  static constexpr int SB = 8;                      // B (weight half-tile) stages
  static constexpr uint32_t A_BOX_BYTES = 130 * 128;
  static constexpr uint32_t A_SLOT = 17 * 1024;     // 1024-aligned slot holding one 130-row box
  static constexpr uint32_t B_BYTES = (BN / 2) * 128;
  static constexpr uint32_t TMEM_COLS = (2 * BN <= 256) ? 256 : 512;
  static constexpr size_t SMEM = 1024 + SA * A_SLOT + SB * B_BYTES + 512;
};

template <int BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(320, 1)
conv2cta_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ CUtensorMap tmP, const GemmParams p, const int kt) {
  using Cfg = Conv2Cfg<BN>;
  constexpr int SA = Cfg::SA, SB = Cfg::SB;
  constexpr int CH = 32;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + SA * Cfg::A_SLOT;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(sB + SB * Cfg::B_BYTES);
  uint64_t* a_empty = a_full + SA;
  uint64_t* b_full = a_empty + SA;
  uint64_t* b_empty = b_full + SB;
  uint64_t* tfull = b_empty + SB;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty + 2);

  __shared__ float gn_red[8 * 64];                  // per-epilogue-warp GroupNorm partials [warp][group][2]
  for (int i = threadIdx.x; i < 8 * 64; i += blockDim.x) gn_red[i] = 0.f;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();          // 0 = leader
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int total_tiles = p.num_m_tiles * p.num_n_tiles;   // m-tiles here are PAIR tiles (two output rows)
  const int groups = kt * 3 * p.cin_blocks;                // (dt, dh, channel block) tap groups

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < SA; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < SB; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 16); }
    fence_mbar_init();
  }
  __syncwarp();
  if (warp == 1) tmem_alloc_2cta<Cfg::TMEM_COLS>(tmem_ptr);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0 && elect_one()) {
    // ===================== TMA producer (both CTAs) =====================
    int sa = 0, sb = 0;
    uint32_t pa = 0, pb = 0;
    for (int tile = pair; tile < total_tiles; tile += npairs) {
      const int mt = tile / p.num_n_tiles, nt = tile % p.num_n_tiles;
      int t, hy, wx;
      conv_tile_coords(p, mt, t, hy, wx);
      const int h = hy * 2 + static_cast<int>(rank);       // this CTA's output row
      const int w0 = wx * 128;
      for (int g = 0; g < groups; ++g) {
        const int cb = g % p.cin_blocks;
        const int dh = (g / p.cin_blocks) % 3;
        const int dt = g / (p.cin_blocks * 3);
        mbar_wait(&a_empty[sa], pa ^ 1);
        if (rank == 0) mbar_expect_tx(&a_full[sa], 2 * Cfg::A_BOX_BYTES);
        int f = t + dt;
        const CUtensorMap* src = conv_frame_src(p, &tmA, &tmP, f);
        tma_load_4d_2cta(sA + sa * Cfg::A_SLOT, src, mapa_shared(smem_u32(&a_full[sa]), 0), cb * 64, w0 - 1,
                         h + dh - 1, f);
        if (++sa == SA) { sa = 0; pa ^= 1; }
#pragma unroll 1
        for (int dw = 0; dw < 3; ++dw) {
          const int kofs = (((dt * 3 + dh) * 3 + dw) * p.cin_blocks + cb) * 64;
          mbar_wait(&b_empty[sb], pb ^ 1);
          if (rank == 0) mbar_expect_tx(&b_full[sb], 2 * Cfg::B_BYTES);
          tma_load_2d_2cta(sB + sb * Cfg::B_BYTES, &tmB, mapa_shared(smem_u32(&b_full[sb]), 0), kofs,
                           nt * BN + static_cast<int>(rank) * (BN / 2));
          if (++sb == SB) { sb = 0; pb ^= 1; }
        }
      }
    }
  } else if (warp == 1 && rank == 0 && elect_one()) {
    // ===================== MMA issuer (leader CTA, one thread) =====================
    constexpr uint32_t idesc = umma_idesc_bf16(256, BN, 0, 0);
    int sa = 0, sb = 0, acc = 0;
    uint32_t pa = 0, pb = 0, acc_phase = 0;
    for (int tile = pair; tile < total_tiles; tile += npairs) {
      mbar_wait(&tempty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int g = 0; g < groups; ++g) {
        mbar_wait(&a_full[sa], pa);
        const uint32_t a_base = smem_u32(sA + sa * Cfg::A_SLOT);
#pragma unroll 1
        for (int dw = 0; dw < 3; ++dw) {
          mbar_wait(&b_full[sb], pb);
          tc_fence_after();
          const uint32_t a_addr = a_base + dw * 128;        // shift by one voxel row inside the halo box
          // the 128B swizzle is a function of the absolute smem address bits, so a row-shifted start address needs
          // no descriptor base offset (verified on B200: setting the base-offset field gives wrong results)
          const uint64_t adesc = umma_desc_sw128(a_addr);
          const uint64_t bdesc = umma_desc_sw128(smem_u32(sB + sb * Cfg::B_BYTES));
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_ss_2cta(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (g | dw | k) != 0);
          umma_commit_2cta(&b_empty[sb]);
          if (++sb == SB) { sb = 0; pb ^= 1; }
        }
        umma_commit_2cta(&a_empty[sa]);
        if (++sa == SA) { sa = 0; pa ^= 1; }
      }
      umma_commit_2cta(&tfull[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  } else if (warp >= 2) {
    // ===================== epilogue (both CTAs; each owns the 128 accumulator lanes of its output row) ==========
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int r_in_tile = q * 32 + lane;
    const uint32_t tempty_leader0 = mapa_shared(smem_u32(&tempty[0]), 0);
    const uint32_t tempty_leader1 = mapa_shared(smem_u32(&tempty[1]), 0);
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = pair; tile < total_tiles; tile += npairs) {
      const int mt = tile / p.num_n_tiles, nt = tile % p.num_n_tiles;
      int t, hy, wx;
      conv_tile_coords(p, mt, t, hy, wx);
      const int h = hy * 2 + static_cast<int>(rank);
      const int w = wx * 128 + r_in_tile;
      const bool valid = (h < p.Ho) && (w < p.Wo);
      const long long row = (static_cast<long long>(t) * p.Ho + h) * p.Wo + w;
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN;
#pragma unroll 1
      for (int c0 = half * CH; c0 < BN; c0 += 2 * CH) {
        uint32_t v[CH];
        tmem_ld32(t_row + c0, v);
        tmem_ld_wait();
        if (p.gn_partial) {
          float qs[8], qss[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) qs[i] = qss[i] = 0.f;
          if (valid) epilogue_chunk<CH, true>(p, v, row, nt * BN + c0, nullptr, qs, qss);
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
              const int g = (nt * BN + c0 + 4 * i) / p.gn_cpg;
              gn_red[(warp - 2) * 64 + g * 2] += qs[i];
              gn_red[(warp - 2) * 64 + g * 2 + 1] += qss[i];
            }
          }
        } else if (valid) {
          epilogue_chunk<CH>(p, v, row, nt * BN + c0, nullptr);
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
  if (p.gn_partial && threadIdx.x < 64) {            // fixed-order combine of the 8 epilogue warps -> deterministic
    float acc = 0.f;
#pragma unroll
    for (int w8 = 0; w8 < 8; ++w8) acc += gn_red[w8 * 64 + threadIdx.x];
    p.gn_partial[static_cast<long long>(blockIdx.x) * 64 + threadIdx.x] = acc;
  }
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2cta<Cfg::TMEM_COLS>(tmem_base);
  }
}

template <int BN>
static int launch_conv2(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmP, const GemmParams& p,
                        int kt, cudaStream_t st) {
  using Cfg = Conv2Cfg<BN>;
  static std::once_flag attr_once;   // thread-safe one-time opt-in to > 48 KB dynamic shared memory
  cudaError_t attr_err = cudaSuccess;
  std::call_once(attr_once, [&] { attr_err = cudaFuncSetAttribute(conv2cta_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(Cfg::SMEM)); });
  if (attr_err != cudaSuccess) return check_cuda(attr_err, "cudaFuncSetAttribute(conv2cta_kernel)");
  const int total = p.num_m_tiles * p.num_n_tiles;
  const int max_pairs = num_sms() / 2;
  const int pairs = total < max_pairs ? total : max_pairs;
  conv2cta_kernel<BN><<<pairs * 2, 320, Cfg::SMEM, st>>>(tmA, tmB, tmP, p, kt);
  DOVE_LAUNCH_CHECK("conv2cta_kernel");
  return DOVE_OK;
}

// Called from dove_conv_cl_bf16 for stride-1 3x3(x3) convs on wide images.  Returns DOVE_OK or an error.
int conv2cta_dispatch(const void* x, const void* x_prev, int Tin, const void* w, int Tout, int Hin, int Win, int Cin,
                      int Cout_pad, int kt, int Ho, int Wo, GemmParams p, cudaStream_t st) {
  const int bn = (Cout_pad % 256 == 0) ? 256 : 128;
  CUtensorMap tmA, tmB, tmP;
  {
    uint64_t dims[4] = {static_cast<uint64_t>(Cin), static_cast<uint64_t>(Win), static_cast<uint64_t>(Hin),
                        static_cast<uint64_t>(Tin)};
    uint64_t strides[3] = {static_cast<uint64_t>(Cin) * 2, static_cast<uint64_t>(Win) * Cin * 2,
                           static_cast<uint64_t>(Hin) * Win * Cin * 2};
    uint32_t box[4] = {64, 130, 1, 1};
    if (int e = make_tmap_bf16(&tmA, x, 4, dims, strides, box, nullptr)) return e;
    tmP = tmA;
    if (x_prev) {
      dims[3] = 2;
      if (int e = make_tmap_bf16(&tmP, x_prev, 4, dims, strides, box, nullptr)) return e;
    }
  }
  const int Ktot = kt * 9 * Cin;
  {
    uint64_t dims[2] = {static_cast<uint64_t>(Ktot), static_cast<uint64_t>(Cout_pad)};
    uint64_t strides[1] = {static_cast<uint64_t>(Ktot) * 2};
    uint32_t box[2] = {64, static_cast<uint32_t>(bn / 2)};
    if (int e = make_tmap_bf16(&tmB, w, 2, dims, strides, box, nullptr)) return e;
  }
  p.tw = 128;
  p.th = 2;                                   // a pair tile covers two output rows
  p.tiles_w = (Wo + 127) / 128;
  p.tiles_h = (Ho + 1) / 2;
  p.num_m_tiles = Tout * p.tiles_w * p.tiles_h;
  p.num_n_tiles = Cout_pad / bn;
  p.To = Tout;
  {
    const double row_bytes = static_cast<double>(Win) * Cin * 2.0 * Tin * 2;
    int bh = static_cast<int>(32.0e6 / row_bytes);
    if (bh < 1) bh = 1;
    if (bh > p.tiles_h) bh = p.tiles_h;
    p.band_h = bh;
  }
  p.cin_blocks = Cin / 64;
  return bn == 256 ? launch_conv2<256>(tmA, tmB, tmP, p, kt, st) : launch_conv2<128>(tmA, tmB, tmP, p, kt, st);
}

}  // namespace dove
