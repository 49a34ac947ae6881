// Implicit-GEMM 3x3x3 convolution for a NARROW output (Cout_pad = 16: the decoder's conv_out, 128 -> 3 channels).
//
// On the generic kernel this layer is bound by L2 -> SM traffic, not by math: every one of the 27 filter taps re-stages a
// [128 voxel x 64 ch] activation box (16 KB) to feed an M128 x N16 MMA of 8 clocks, 972 KB of TMA per 128 output voxels,
// 5.8 TB/s of L2 reads for 16 TFLOP/s (bench class "conv 128 3 3 8 768 1280": 10.3 ms per frame batch, 3.2 % of the clip).
// This kernel removes the re-staging:
//   * the WEIGHTS (27 taps x Cin x 16 rows, 108 KB for Cin = 128) are loaded into shared memory ONCE per CTA and stay;
//   * the activations arrive as one HALO box per (frame tap dt, channel block): 3 rows x 130 voxels x 64 ch (48.75 KB),
//     and the 9 (dh, dw) taps are formed by the MMA issuer with row-shifted shared-memory descriptors (start address +
//     (dh*130 + dw) * 128 B; the 128B swizzle follows absolute address bits, as in conv2.cu).  A-staging drops from 27
//     boxes of 16 KB to 6 boxes of 48.75 KB per 128 voxels (432 KB -> 293 KB incl. no weight traffic: 3.3x less).
// CTA = 192 threads: warp 0 TMA producer, warp 1 TMEM owner + MMA issuer, warps 2..5 epilogue (one voxel per thread,
// 16 accumulator columns: bias, planar store with the optional fused `*0.5+0.5, clamp` / uint8 quantisation).
#include <mutex>

#include "gemm_common.cuh"

namespace dove {

struct Conv3Cfg {
  static constexpr int SA = 2;                          // halo-box stages
  static constexpr uint32_t A_BOX_BYTES = 3 * 130 * 128;
  static constexpr uint32_t A_SLOT = 49 * 1024;         // 1024-aligned slot holding one 3 x 130-row box
  static constexpr uint32_t B_KB_BYTES = 16 * 128;      // one 64-wide k-block of the 16-row weight tile
  static constexpr int MAX_KB = 54;                     // 27 taps x 2 channel blocks (Cin <= 128)
  static constexpr size_t SMEM = 1024 + SA * A_SLOT + MAX_KB * B_KB_BYTES + 256;
};

__global__ void __launch_bounds__(192, 1)
conv_narrow_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                   const __grid_constant__ CUtensorMap tmP, const GemmParams p) {
  using Cfg = Conv3Cfg;
  constexpr int SA = Cfg::SA;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + SA * Cfg::A_SLOT;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(sB + Cfg::MAX_KB * Cfg::B_KB_BYTES);
  uint64_t* a_empty = a_full + SA;
  uint64_t* b_full = a_empty + SA;
  uint64_t* tfull = b_full + 1;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_tiles = p.num_m_tiles;                // one tile = 128 consecutive voxels of one image row
  const int num_kb = 27 * p.cin_blocks;
  const int groups = 3 * p.cin_blocks;                  // (dt, channel block) halo boxes per tile

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < SA; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    mbar_init(b_full, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4); }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<32>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0 && elect_one()) {
    // ===================== TMA producer =====================
    mbar_expect_tx(b_full, static_cast<uint32_t>(num_kb) * Cfg::B_KB_BYTES);       // resident weights, loaded once
    for (int kb = 0; kb < num_kb; ++kb) tma_load_2d(sB + kb * Cfg::B_KB_BYTES, &tmB, b_full, kb * 64, 0);
    int sa = 0;
    uint32_t pa = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      int t, h, wx;
      conv_tile_coords(p, tile, t, h, wx);
      const int w0 = wx * 128;
      for (int g = 0; g < groups; ++g) {
        const int cb = g % p.cin_blocks, dt = g / p.cin_blocks;
        mbar_wait(&a_empty[sa], pa ^ 1);
        mbar_expect_tx(&a_full[sa], Cfg::A_BOX_BYTES);
        int f = t + dt;
        const CUtensorMap* src = conv_frame_src(p, &tmA, &tmP, f);
        tma_load_4d(sA + sa * Cfg::A_SLOT, src, &a_full[sa], cb * 64, w0 - 1, h - 1, f);
        if (++sa == SA) { sa = 0; pa ^= 1; }
      }
    }
  } else if (warp == 1 && elect_one()) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = umma_idesc_bf16(128, 16, 0, 0);
    mbar_wait(b_full, 0);
    int sa = 0, acc = 0;
    uint32_t pa = 0, acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      mbar_wait(&tempty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * 16;
      for (int g = 0; g < groups; ++g) {
        const int cb = g % p.cin_blocks, dt = g / p.cin_blocks;
        mbar_wait(&a_full[sa], pa);
        tc_fence_after();
        const uint32_t a_base = smem_u32(sA + sa * Cfg::A_SLOT);
#pragma unroll 1
        for (int tap = 0; tap < 9; ++tap) {              // (dh, dw): the halo box shifted by whole 128-byte voxel rows
          const int dh = tap / 3, dw = tap - dh * 3;
          const uint64_t adesc = umma_desc_sw128(a_base + (dh * 130 + dw) * 128);
          const int kb = ((dt * 3 + dh) * 3 + dw) * p.cin_blocks + cb;
          const uint64_t bdesc = umma_desc_sw128(smem_u32(sB + kb * Cfg::B_KB_BYTES));
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_ss(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (g | tap | k) != 0);
        }
        umma_commit(&a_empty[sa]);
        if (++sa == SA) { sa = 0; pa ^= 1; }
      }
      umma_commit(&tfull[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  } else if (warp >= 2) {
    // ===================== epilogue (4 warps, one output voxel per thread) =====================
    const int q = warp & 3;
    const int r_in_tile = q * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      int t, h, wx;
      conv_tile_coords(p, tile, t, h, wx);
      const int w = wx * 128 + r_in_tile;
      const bool valid = w < p.Wo;
      const long long row = (static_cast<long long>(t) * p.Ho + h) * p.Wo + w;
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      uint32_t v[16];
      tmem_ld16(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * 16, v);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);          // accumulator is in registers: release it before the stores
      if (valid) epilogue_chunk<16>(p, v, row, 0, nullptr);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<32>(tmem_base);
  }
}

// Called from conv_impl for stride-1 3x3x3 convs with Cout_pad == 16 and Cin <= 128 on rows >= 128 voxels wide.
int conv_narrow_dispatch(const void* x, const void* x_prev, int Tin, const void* w, int Tout, int H, int W, int Cin,
                         GemmParams p, cudaStream_t st) {
  using Cfg = Conv3Cfg;
  CUtensorMap tmA, tmB, tmP;
  {
    uint64_t dims[4] = {static_cast<uint64_t>(Cin), static_cast<uint64_t>(W), static_cast<uint64_t>(H),
                        static_cast<uint64_t>(Tin)};
    uint64_t strides[3] = {static_cast<uint64_t>(Cin) * 2, static_cast<uint64_t>(W) * Cin * 2,
                           static_cast<uint64_t>(H) * W * Cin * 2};
    uint32_t box[4] = {64, 130, 3, 1};
    if (int e = make_tmap_bf16(&tmA, x, 4, dims, strides, box, nullptr)) return e;
    tmP = tmA;
    if (x_prev) {
      dims[3] = 2;
      if (int e = make_tmap_bf16(&tmP, x_prev, 4, dims, strides, box, nullptr)) return e;
    }
  }
  const int Ktot = 27 * Cin;
  {
    uint64_t dims[2] = {static_cast<uint64_t>(Ktot), 16};
    uint64_t strides[1] = {static_cast<uint64_t>(Ktot) * 2};
    uint32_t box[2] = {64, 16};
    if (int e = make_tmap_bf16(&tmB, w, 2, dims, strides, box, nullptr)) return e;
  }
  p.tw = 128;
  p.th = 1;
  p.tiles_w = (W + 127) / 128;
  p.tiles_h = H;
  p.num_m_tiles = Tout * p.tiles_w * p.tiles_h;
  p.num_n_tiles = 1;
  p.To = Tout;
  {   // band height as in the other conv kernels: a band of rows x all input frames of the band within ~32 MB of L2
    const double row_bytes = static_cast<double>(W) * Cin * 2.0 * Tin;
    int bh = static_cast<int>(32.0e6 / row_bytes);
    if (bh < 1) bh = 1;
    if (bh > p.tiles_h) bh = p.tiles_h;
    p.band_h = bh;
  }
  p.cin_blocks = Cin / 64;
  static std::once_flag attr_once;
  cudaError_t attr_err = cudaSuccess;
  std::call_once(attr_once, [&] {
    attr_err = cudaFuncSetAttribute(conv_narrow_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    static_cast<int>(Cfg::SMEM));
  });
  if (attr_err != cudaSuccess) return check_cuda(attr_err, "cudaFuncSetAttribute(conv_narrow_kernel)");
  const int grid = p.num_m_tiles < num_sms() ? p.num_m_tiles : num_sms();
  conv_narrow_kernel<<<grid, 192, Cfg::SMEM, st>>>(tmA, tmB, tmP, p);
  DOVE_LAUNCH_CHECK("conv_narrow_kernel");
  return DOVE_OK;
}

}  // namespace dove
