// Swapped-operand implicit-GEMM convolution (weights = M 128, voxels = N 256) WITH in-smem reuse of the W-direction taps,
// for the 128-output-channel 3x3(x3) stride-1 convs on wide images (the dominant kernel class of the cfg-2 step).
//
// The generic swapped-operand kernel (gemm.cu, kTrans) stages a fresh [256 voxel x 64 ch] activation box (32 KB) and a
// [128 cout x 64 ch] weight tile (16 KB) for EVERY filter tap and channel block through one 4-stage ring: 83 GB of
// L2 -> SM traffic per 128->128 full-resolution launch (15 TB/s) and only 4 k-blocks (2 048 MMA cycles) of look-ahead —
// ncu: tensor pipe 84 %.  Here the tile is ONE image row of 256 consecutive voxels, so the three taps dw = 0, 1, 2 of a
// (dt, dh) group are the same shared-memory box shifted by one voxel = one 128-byte row (the trick of conv2.cu, applied to
// the N operand):
//   * B ring: one [258 voxel x 64 ch] halo box per (dt, dh, channel block) — two TMA boxes (128 + 130 voxels; a box
//     dimension may not exceed 256) landing back to back — consumed by three MMAs groups through row-shifted descriptors;
//     activation staging drops 2.9x;
//   * A ring: the weight tiles on their own 5-stage ring, so the two operand streams never wait for each other;
//   * look-ahead: 4 halo boxes = 12 taps x 4 MMAs = 6 144 MMA cycles.
// CTA = 320 threads: warp 0 TMA producer, warp 1 TMEM owner + MMA issuer (M128 x N256 x K16, accumulator [channel lane]
// [voxel column], double-buffered: 512 TMEM columns), warps 2..9 epilogue (identical to the kTrans epilogue of gemm.cu:
// lane pairs swap one value so every store is a packed (c, c+1) pair; fused residual add and GroupNorm statistics).
#include <mutex>

#include "gemm_common.cuh"

namespace dove {

struct Conv4Cfg {
  static constexpr int SA = 5;                         // weight-tile stages
  static constexpr int SB = 4;                         // halo-box stages
  static constexpr uint32_t A_BYTES = 128 * 128;       // 128 cout x 64 ch
  static constexpr uint32_t B_BOX1 = 128 * 128;        // voxels w0-1 .. w0+126
  static constexpr uint32_t B_BOX2 = 130 * 128;        // voxels w0+127 .. w0+tw (tw <= 256)
  static constexpr uint32_t B_SLOT = 33 * 1024;        // 1024-aligned slot holding the 258 rows
  static constexpr size_t SMEM = 1024 + SA * A_BYTES + SB * B_SLOT + 512;
};

__global__ void __launch_bounds__(320, 1)
conv_trans_halo_kernel(const __grid_constant__ CUtensorMap tmX1, const __grid_constant__ CUtensorMap tmX2,
                       const __grid_constant__ CUtensorMap tmP1, const __grid_constant__ CUtensorMap tmP2,
                       const __grid_constant__ CUtensorMap tmW, const GemmParams p, const int kt) {
  using Cfg = Conv4Cfg;
  constexpr int SA = Cfg::SA, SB = Cfg::SB;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sB = smem;
  uint8_t* sA = smem + SB * Cfg::B_SLOT;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(sA + SA * Cfg::A_BYTES);
  uint64_t* a_empty = a_full + SA;
  uint64_t* b_full = a_empty + SA;
  uint64_t* b_empty = b_full + SB;
  uint64_t* tfull = b_empty + SB;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty + 2);

  __shared__ float gn_red[8 * 64];                  // per-epilogue-warp GroupNorm partials [warp][group][2]
  for (int i = threadIdx.x; i < 8 * 64; i += blockDim.x) gn_red[i] = 0.f;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_tiles = p.num_m_tiles * p.num_n_tiles;
  const int groups = kt * 3 * p.cin_blocks;         // (dt, dh, channel block) tap groups

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX1);
    tma_prefetch_desc(&tmX2);
    tma_prefetch_desc(&tmW);
    for (int i = 0; i < SA; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < SB; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 8); }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0 && elect_one()) {
    // ===================== TMA producer =====================
    int sa = 0, sb = 0;
    uint32_t pa = 0, pb = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int mt = tile / p.num_n_tiles, nt = tile % p.num_n_tiles;
      int t, h, wx;
      conv_tile_coords(p, mt, t, h, wx);             // th = 1: the tile row IS the image row
      const int w0 = wx * p.tw;
      const uint32_t box2_bytes = static_cast<uint32_t>(p.tw - 126) * 128;   // voxels w0+127 .. w0+tw
      for (int g = 0; g < groups; ++g) {
        const int cb = g % p.cin_blocks;
        const int dh = (g / p.cin_blocks) % 3;
        const int dt = g / (p.cin_blocks * 3);
        int f = t + dt;
        const bool from_prev = conv_frame_src(p, &tmX1, &tmP1, f) == &tmP1;
        mbar_wait(&b_empty[sb], pb ^ 1);
        mbar_expect_tx(&b_full[sb], Cfg::B_BOX1 + box2_bytes);
        uint8_t* slot = sB + sb * Cfg::B_SLOT;
        tma_load_4d(slot, from_prev ? &tmP1 : &tmX1, &b_full[sb], cb * 64, w0 - 1, h + dh - 1, f);
        tma_load_4d(slot + Cfg::B_BOX1, from_prev ? &tmP2 : &tmX2, &b_full[sb], cb * 64, w0 + 127, h + dh - 1, f);
        if (++sb == SB) { sb = 0; pb ^= 1; }
#pragma unroll 1
        for (int dw = 0; dw < 3; ++dw) {
          const int kofs = (((dt * 3 + dh) * 3 + dw) * p.cin_blocks + cb) * 64;
          mbar_wait(&a_empty[sa], pa ^ 1);
          mbar_expect_tx(&a_full[sa], Cfg::A_BYTES);
          tma_load_2d(sA + sa * Cfg::A_BYTES, &tmW, &a_full[sa], kofs, nt * 128);
          if (++sa == SA) { sa = 0; pa ^= 1; }
        }
      }
    }
  } else if (warp == 1 && elect_one()) {
    // ===================== MMA issuer (one thread) =====================
    int sa = 0, sb = 0, acc = 0;
    uint32_t pa = 0, pb = 0, acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      // weights (K-major) x voxels (K-major), N = the tile's valid voxels rounded up to 16: rows are cut into EQUAL tiles
      // of tw <= 256 voxels (dispatch), so any width >= 256 runs at >= 94 % MMA utilisation
      int t_, h_, wx;
      conv_tile_coords(p, tile / p.num_n_tiles, t_, h_, wx);
      const int nvox = min(p.tw, p.Wo - wx * p.tw);
      const uint32_t idesc = umma_idesc_bf16(128, (nvox + 15) & ~15, 0, 0);
      mbar_wait(&tempty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * 256;
      for (int g = 0; g < groups; ++g) {
        mbar_wait(&b_full[sb], pb);
        const uint32_t b_base = smem_u32(sB + sb * Cfg::B_SLOT);
#pragma unroll 1
        for (int dw = 0; dw < 3; ++dw) {
          mbar_wait(&a_full[sa], pa);
          tc_fence_after();
          // rows dw .. dw+255 of the halo box: the 128B swizzle is a function of the absolute smem address bits, so a
          // row-shifted start address needs no descriptor base offset (conv2.cu, verified on B200)
          const uint64_t adesc = umma_desc_sw128(smem_u32(sA + sa * Cfg::A_BYTES));
          const uint64_t bdesc = umma_desc_sw128(b_base + dw * 128);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_ss(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (g | dw | k) != 0);
          umma_commit(&a_empty[sa]);
          if (++sa == SA) { sa = 0; pa ^= 1; }
        }
        umma_commit(&b_empty[sb]);
        if (++sb == SB) { sb = 0; pb ^= 1; }
      }
      umma_commit(&tfull[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  } else if (warp >= 2) {
    // ===================== epilogue: lane = output channel, accumulator columns = the row's 256 voxels =====================
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int r_in_tile = q * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    float gn_s = 0.f, gn_ss = 0.f;          // this thread's channel-pair sums over every voxel it stores
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int mt = tile / p.num_n_tiles, nt = tile % p.num_n_tiles;
      int t, h, wx;
      conv_tile_coords(p, mt, t, h, wx);
      const int ch = nt * 128 + r_in_tile;
      const float bias_v = p.bias ? __bfloat162float(p.bias[ch]) : 0.f;
      const long long row0 = (static_cast<long long>(t) * p.Ho + h) * p.Wo + wx * p.tw;
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * 256;
      const int ncols = min(p.tw, p.Wo - wx * p.tw);        // columns beyond the tile / row end were not computed
#pragma unroll 1
      for (int c0 = half * 32; c0 < ncols; c0 += 64) {
        uint32_t v[32];
        tmem_ld32(t_row + c0, v);
        tmem_ld_wait();
        // voxels c0+2i (even lanes store it) and c0+2i+1 (odd lanes): lane pairs swap one value so that every lane stores
        // a packed (channel, channel+1) pair.  All residual loads are issued BEFORE any store (C and aux may alias).
        long long offs[16];
        uint32_t res[16];
        bool oks[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int vi = c0 + 2 * i + (lane & 1);
          oks[i] = vi < ncols;
          const long long row = row0 + vi;
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
    }
    if (p.gn_partial) {                      // 4 lanes = one group of 4 channels (C = 128, 32 groups)
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
  if (p.gn_partial && threadIdx.x < 64) {    // fixed-order combine of the 8 epilogue warps -> deterministic
    float a8 = 0.f;
#pragma unroll
    for (int w8 = 0; w8 < 8; ++w8) a8 += gn_red[w8 * 64 + threadIdx.x];
    p.gn_partial[static_cast<long long>(blockIdx.x) * 64 + threadIdx.x] = a8;
  }
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// Called from conv_impl (gemm.cu) for stride-1 3x3(x3) convs with exactly 128 output channels on images at least 256
// voxels wide.  q carries the epilogue / GroupNorm / causal-cache fields.
int conv_trans_halo_dispatch(const void* x, const void* x_prev, int Tin, const void* w, int Tout, int Hin, int Win, int Cin,
                             int Cout_pad, int kt, int Ho, int Wo, GemmParams p, cudaStream_t st) {
  using Cfg = Conv4Cfg;
  CUtensorMap tmX1, tmX2, tmP1, tmP2, tmW;
  {
    uint64_t dims[4] = {static_cast<uint64_t>(Cin), static_cast<uint64_t>(Win), static_cast<uint64_t>(Hin),
                        static_cast<uint64_t>(Tin)};
    uint64_t strides[3] = {static_cast<uint64_t>(Cin) * 2, static_cast<uint64_t>(Win) * Cin * 2,
                           static_cast<uint64_t>(Hin) * Win * Cin * 2};
    const int tiles_w_ = (Wo + 255) / 256;
    const int tw_ = (((Wo + tiles_w_ - 1) / tiles_w_) + 15) & ~15;              // 144 .. 256
    uint32_t box1[4] = {64, 128, 1, 1}, box2[4] = {64, static_cast<uint32_t>(tw_ - 126), 1, 1};
    if (int e = make_tmap_bf16(&tmX1, x, 4, dims, strides, box1, nullptr)) return e;
    if (int e = make_tmap_bf16(&tmX2, x, 4, dims, strides, box2, nullptr)) return e;
    tmP1 = tmX1;
    tmP2 = tmX2;
    if (x_prev) {
      dims[3] = 2;
      if (int e = make_tmap_bf16(&tmP1, x_prev, 4, dims, strides, box1, nullptr)) return e;
      if (int e = make_tmap_bf16(&tmP2, x_prev, 4, dims, strides, box2, nullptr)) return e;
    }
  }
  const int Ktot = kt * 9 * Cin;
  {
    uint64_t dims[2] = {static_cast<uint64_t>(Ktot), static_cast<uint64_t>(Cout_pad)};
    uint64_t strides[1] = {static_cast<uint64_t>(Ktot) * 2};
    uint32_t box[2] = {64, 128};
    if (int e = make_tmap_bf16(&tmW, w, 2, dims, strides, box, nullptr)) return e;
  }
  // row tiles of equal width: tiles_w = ceil(Wo / 256), tw = ceil(Wo / tiles_w) rounded up to the MMA N granularity (16).
  // (A 256 + remainder split leaves narrow tiles that re-read the whole 0.9 MB weight set for few voxels: measured 1 163
  // vs 1 301 TFLOP/s of the generic kernel at Wo = 368.)
  p.tiles_w = (Wo + 255) / 256;
  p.tw = (((Wo + p.tiles_w - 1) / p.tiles_w) + 15) & ~15;
  p.th = 1;
  p.tiles_h = Ho;
  p.num_m_tiles = Tout * p.tiles_w * p.tiles_h;
  p.num_n_tiles = Cout_pad / 128;
  p.To = Tout;
  {
    const double row_bytes = static_cast<double>(Win) * Cin * 2.0 * Tin;
    int bh = static_cast<int>(32.0e6 / row_bytes);
    if (bh < 1) bh = 1;
    if (bh > p.tiles_h) bh = p.tiles_h;
    p.band_h = bh;
  }
  p.cin_blocks = Cin / 64;
  static std::once_flag attr_once;   // thread-safe one-time opt-in to > 48 KB dynamic shared memory
  cudaError_t attr_err = cudaSuccess;
  std::call_once(attr_once, [&] {
    attr_err = cudaFuncSetAttribute(conv_trans_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    static_cast<int>(Cfg::SMEM));
  });
  if (attr_err != cudaSuccess) return check_cuda(attr_err, "cudaFuncSetAttribute(conv_trans_halo_kernel)");
  const int total = p.num_m_tiles * p.num_n_tiles;
  const int grid = total < num_sms() ? total : num_sms();
  conv_trans_halo_kernel<<<grid, 320, Cfg::SMEM, st>>>(tmX1, tmX2, tmP1, tmP2, tmW, p, kt);
  DOVE_LAUNCH_CHECK("conv_trans_halo_kernel");
  return DOVE_OK;
}

}  // namespace dove
