// HBM-bound kernels of the DOVE hot path (norms, RoPE, layout glue, resampling).  All are coalesced,
// 16-byte vectorised, fp32 math with the reference's bf16 rounding points reproduced explicitly.
#include "common.cuh"
#include "ptx.cuh"

namespace dove {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float silu_f(float x) { return __fdividef(x, 1.0f + __expf(-x)); }
// SiLU with ONE MUFU op per element (x * sigmoid(x) = h + h * tanh(h), h = x/2) instead of ex2 + rcp: the GroupNorm-apply
// pass was MUFU-bound at ~0.5 of HBM speed with two.  tanh.approx.f32 is accurate to ~2^-11 of tanh, i.e. an absolute
// error <= |x| * 2.5e-4 in the result, far below the bf16 rounding of the stored output (2^-9 relative).
__device__ __forceinline__ float silu_fast(float x) {
  const float h = 0.5f * x;
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return fmaf(h, t, h);
}
__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  return make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]),
                    pack_bf16x2(f[6], f[7]));
}
static inline int grid_for(long long work_items, int threads, int max_blocks_per_sm = 16) {
  long long b = (work_items + threads - 1) / threads;
  long long cap = static_cast<long long>(num_sms()) * max_blocks_per_sm;
  return static_cast<int>(b < 1 ? 1 : (b > cap ? cap : b));
}

// ------------------------------------------------------------------------------------------------ LayerNorm
// one warp per row; D % 256 == 0, D <= 4096 (each lane holds D/32 values as packed bf16)
__global__ void __launch_bounds__(256) layernorm_mod_kernel(const bf16* __restrict__ x, bf16* __restrict__ out,
                                                            int rows, int D, const bf16* __restrict__ w,
                                                            const bf16* __restrict__ b, float eps,
                                                            const bf16* __restrict__ scale0,
                                                            const bf16* __restrict__ shift0,
                                                            const bf16* __restrict__ scale1,
                                                            const bf16* __restrict__ shift1, int split_row) {
  const int lane = threadIdx.x & 31;
  const int wid = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int nwarps = gridDim.x * (blockDim.x >> 5);
  const int nv = D >> 8;   // uint4 vectors per lane
  for (int row = wid; row < rows; row += nwarps) {
    const uint4* xr = reinterpret_cast<const uint4*>(x + static_cast<long long>(row) * D);
    uint4 v[16];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i)
      if (i < nv) {
        v[i] = xr[i * 32 + lane];
        float f[8];
        unpack8(v[i], f);
#pragma unroll
        for (int k = 0; k < 8; ++k) s += f[k];
      }
    const float mean = warp_sum(s) / D;
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i)
      if (i < nv) {
        float f[8];
        unpack8(v[i], f);
#pragma unroll
        for (int k = 0; k < 8; ++k) ss += (f[k] - mean) * (f[k] - mean);
      }
    const float rstd = rsqrtf(warp_sum(ss) / D + eps);
    const bf16* sc = scale0 ? (row < split_row ? scale0 : scale1) : nullptr;
    const bf16* sh = scale0 ? (row < split_row ? shift0 : shift1) : nullptr;
    uint4* orow = reinterpret_cast<uint4*>(out + static_cast<long long>(row) * D);
    const __nv_bfloat162 one2 = __floats2bfloat162_rn(1.0f, 1.0f);
#pragma unroll
    for (int i = 0; i < 16; ++i)
      if (i < nv) {
        const int col = (i * 32 + lane) * 8;
        float f[8], wv[8], bv[8];
        unpack8(v[i], f);
        unpack8(*reinterpret_cast<const uint4*>(w + col), wv);
        unpack8(*reinterpret_cast<const uint4*>(b + col), bv);
        uint32_t n[4];
#pragma unroll
        for (int k = 0; k < 4; ++k)      // LayerNorm in fp32, ONE rounding to bf16 (F.layer_norm's output)
          n[k] = pack_bf16x2((f[2 * k] - mean) * rstd * wv[2 * k] + bv[2 * k],
                             (f[2 * k + 1] - mean) * rstd * wv[2 * k + 1] + bv[2 * k + 1]);
        if (sc) {
          // norm * (1 + scale) + shift as three bf16 tensor ops (each rounds to nearest even): packed bf16x2
          // arithmetic gives exactly those roundings (the product / sum of two bf16 is exact in fp32) at a quarter of
          // the instructions of the unpack -> fp32 -> round sequence; _rn forms so that mul + add never contract
          const uint4 s4 = *reinterpret_cast<const uint4*>(sc + col);
          const uint4 h4 = *reinterpret_cast<const uint4*>(sh + col);
          const uint32_t sw[4] = {s4.x, s4.y, s4.z, s4.w}, hw[4] = {h4.x, h4.y, h4.z, h4.w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const __nv_bfloat162 t1 = __hadd2_rn(one2, *reinterpret_cast<const __nv_bfloat162*>(&sw[k]));
            const __nv_bfloat162 t2 = __hmul2_rn(*reinterpret_cast<const __nv_bfloat162*>(&n[k]), t1);
            const __nv_bfloat162 t3 = __hadd2_rn(t2, *reinterpret_cast<const __nv_bfloat162*>(&hw[k]));
            n[k] = *reinterpret_cast<const uint32_t*>(&t3);
          }
        }
        orow[i * 32 + lane] = make_uint4(n[0], n[1], n[2], n[3]);
      }
  }
}

// ------------------------------------------------------------------------------------------------ QK-LN + RoPE
// one warp per (row, head); lane i owns elements (2i, 2i+1) of q and of k
__global__ void __launch_bounds__(256) qk_norm_rope_kernel(bf16* __restrict__ qkv, int rows, int heads,
                                                           const bf16* __restrict__ qw, const bf16* __restrict__ qb,
                                                           const bf16* __restrict__ kw, const bf16* __restrict__ kb,
                                                           float eps, const float* __restrict__ cosT,
                                                           const float* __restrict__ sinT, int text_len) {
  const int lane = threadIdx.x & 31;
  const long long wid = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = static_cast<long long>(gridDim.x) * (blockDim.x >> 5);
  const long long total = static_cast<long long>(rows) * heads;
  const float2 qwv = unpack_bf16x2(reinterpret_cast<const uint32_t*>(qw)[lane]);
  const float2 qbv = unpack_bf16x2(reinterpret_cast<const uint32_t*>(qb)[lane]);
  const float2 kwv = unpack_bf16x2(reinterpret_cast<const uint32_t*>(kw)[lane]);
  const float2 kbv = unpack_bf16x2(reinterpret_cast<const uint32_t*>(kb)[lane]);
  for (long long it = wid; it < total; it += nwarps) {
    const int row = static_cast<int>(it / heads), head = static_cast<int>(it % heads);
    uint32_t* base = reinterpret_cast<uint32_t*>(qkv + static_cast<long long>(row) * 3 * heads * 64);
    float2 c2 = make_float2(1.f, 1.f), s2 = make_float2(0.f, 0.f);
    const bool rope = row >= text_len;
    if (rope) {
      c2 = reinterpret_cast<const float2*>(cosT + static_cast<long long>(row - text_len) * 64)[lane];
      s2 = reinterpret_cast<const float2*>(sinT + static_cast<long long>(row - text_len) * 64)[lane];
    }
#pragma unroll
    for (int which = 0; which < 2; ++which) {
      uint32_t* ptr = base + (which * heads + head) * 32 + lane;
      float2 xv = unpack_bf16x2(*ptr);
      const float mean = warp_sum(xv.x + xv.y) * (1.0f / 64.0f);
      const float dx = xv.x - mean, dy = xv.y - mean;
      const float rstd = rsqrtf(warp_sum(dx * dx + dy * dy) * (1.0f / 64.0f) + eps);
      const float2 wv = which ? kwv : qwv, bv = which ? kbv : qbv;
      float y0 = bf16_round(dx * rstd * wv.x + bv.x);
      float y1 = bf16_round(dy * rstd * wv.y + bv.y);
      if (rope) {   // out = x*cos + rot(x)*sin, rot(x)[2i] = -x[2i+1], rot(x)[2i+1] = x[2i]  (fp32, no fma)
        const float o0 = __fadd_rn(__fmul_rn(y0, c2.x), __fmul_rn(-y1, s2.x));
        const float o1 = __fadd_rn(__fmul_rn(y1, c2.y), __fmul_rn(y0, s2.y));
        y0 = o0;
        y1 = o1;
      }
      *ptr = pack_bf16x2(y0, y1);
    }
  }
}

// ------------------------------------------------------------------------------------------------ GEMV
__global__ void __launch_bounds__(256) gemv_kernel(const bf16* __restrict__ x, const bf16* __restrict__ W,
                                                   const bf16* __restrict__ b, bf16* __restrict__ y, int N, int K,
                                                   int silu_in) {
  const int lane = threadIdx.x & 31;
  const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (n >= N) return;
  const bf16* wr = W + static_cast<long long>(n) * K;
  float acc = 0.f;
  for (int k = lane * 2; k < K; k += 64) {
    float2 xv = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(x + k));
    if (silu_in) {
      xv.x = bf16_round(silu_f(xv.x));
      xv.y = bf16_round(silu_f(xv.y));
    }
    const float2 wv = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(wr + k));
    acc += xv.x * wv.x + xv.y * wv.y;
  }
  acc = warp_sum(acc);
  if (lane == 0) y[n] = __float2bfloat16_rn(acc + (b ? __bfloat162float(b[n]) : 0.f));
}

// ------------------------------------------------------------------------------------------------ patchify
__global__ void patchify_kernel(const bf16* __restrict__ lat, bf16* __restrict__ tok, int F, int C, int h, int w) {
  const int hp = h / 2, wp = w / 2, feat = C * 8;
  const long long total = static_cast<long long>(F / 2) * hp * wp * feat / 2;   // element pairs (pw = 0,1)
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int fp = static_cast<int>(i % (feat / 2));
    const long long token = i / (feat / 2);
    const int xq = static_cast<int>(token % wp), yq = static_cast<int>((token / wp) % hp);
    const int tq = static_cast<int>(token / (static_cast<long long>(wp) * hp));
    const int ph = fp & 1, pt = (fp >> 1) & 1, c = fp >> 2;
    const long long src = ((static_cast<long long>(2 * tq + pt) * C + c) * h + (2 * yq + ph)) * w + 2 * xq;
    reinterpret_cast<uint32_t*>(tok)[i] = *reinterpret_cast<const uint32_t*>(lat + src);
  }
}

__global__ void unpatchify_velocity_kernel(const bf16* __restrict__ tok, const bf16* __restrict__ lat,
                                           bf16* __restrict__ x0, bf16* __restrict__ pred, int F, int C, int h,
                                           int w, float a, float b) {
  const int hp = h / 2, wp = w / 2, feat = C * 8;
  const long long total = static_cast<long long>(F) * C * h * w / 2;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long e = i * 2;
    const int x = static_cast<int>(e % w), y = static_cast<int>((e / w) % h);
    const int c = static_cast<int>((e / (static_cast<long long>(w) * h)) % C);
    const int f = static_cast<int>(e / (static_cast<long long>(w) * h * C));
    const long long token = (static_cast<long long>(f / 2) * hp + y / 2) * wp + x / 2;
    const int fidx = ((c * 2 + (f & 1)) * 2 + (y & 1)) * 2;
    const uint32_t pv = *reinterpret_cast<const uint32_t*>(tok + token * feat + fidx);
    if (pred) reinterpret_cast<uint32_t*>(pred)[i] = pv;
    if (x0) {
      const float2 p2 = unpack_bf16x2(pv);
      const float2 l2 = unpack_bf16x2(reinterpret_cast<const uint32_t*>(lat)[i]);
      reinterpret_cast<uint32_t*>(x0)[i] = pack_bf16x2(bf16_round(a * l2.x) - bf16_round(b * p2.x),
                                                       bf16_round(a * l2.y) - bf16_round(b * p2.y));
    }
  }
}

__global__ void velocity_kernel(const bf16* __restrict__ sample, const bf16* __restrict__ noise,
                                bf16* __restrict__ out, long long n, float a, float b) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    out[i] = __float2bfloat16_rn(bf16_round(a * __bfloat162float(noise[i])) -
                                 bf16_round(b * __bfloat162float(sample[i])));
}

// ------------------------------------------------------------------------------------------------ GroupNorm
constexpr int GN_MAX_BLOCKS = 1184;   // == GN_PARTIAL_ROWS in gemm_common.cuh

// x [nvox, C] channels-last.  thread -> fixed 8-channel vector column, strided over voxels.
__global__ void __launch_bounds__(256) gn_partial_kernel(const bf16* __restrict__ x, long long nvox, int C,
                                                         int groups, float* __restrict__ partial) {
  // per-thread partials are combined in a FIXED order (no atomics) so results are run-to-run deterministic
  __shared__ float sh[256 * 4];
  const int vcols = C >> 3;
  const int vcol = threadIdx.x % vcols, vlane = threadIdx.x / vcols, vper = blockDim.x / vcols;
  const int cpg = C / groups;
  float s0 = 0.f, q0 = 0.f, s1 = 0.f, q1 = 0.f;
  for (long long v = static_cast<long long>(blockIdx.x) * vper + vlane; v < nvox;
       v += static_cast<long long>(gridDim.x) * vper) {
    float f[8];
    unpack8(*reinterpret_cast<const uint4*>(x + v * C + vcol * 8), f);
#pragma unroll
    for (int k = 0; k < 4; ++k) { s0 += f[k]; q0 += f[k] * f[k]; }
#pragma unroll
    for (int k = 4; k < 8; ++k) { s1 += f[k]; q1 += f[k] * f[k]; }
  }
  sh[threadIdx.x * 4 + 0] = s0;
  sh[threadIdx.x * 4 + 1] = q0;
  sh[threadIdx.x * 4 + 2] = s1;
  sh[threadIdx.x * 4 + 3] = q1;
  __syncthreads();
  if (threadIdx.x < groups * 2) {
    const int g = threadIdx.x >> 1, which = threadIdx.x & 1;   // which: 0 = sum, 1 = sum of squares
    float acc = 0.f;
    for (int t = 0; t < 256; ++t) {
      const int tv = t % vcols;
      if (cpg == 4) {
        if (tv == (g >> 1)) acc += sh[t * 4 + (g & 1) * 2 + which];
      } else if (tv * 8 / cpg == g) {
        acc += sh[t * 4 + which] + sh[t * 4 + 2 + which];
      }
    }
    partial[static_cast<long long>(blockIdx.x) * groups * 2 + threadIdx.x] = acc;
  }
}

// one 256-thread block per group: threads stride over the per-CTA partials, fp64 accumulation, fixed-order tree combine
// (deterministic).  (Round 1 used one WARP per group in a single block: 37 dependent fp64 adds per lane, ~58 us per call x
// 264 calls per clip.)
__global__ void __launch_bounds__(256) gn_finalize_kernel(const float* __restrict__ partial, int nblocks, int groups,
                                                          double count, float eps, float* __restrict__ stats) {
  __shared__ double sh_s[256], sh_q[256];
  const int g = blockIdx.x;
  double s = 0.0, q = 0.0;
  for (int b = threadIdx.x; b < nblocks; b += 256) {
    const float2 v = *reinterpret_cast<const float2*>(partial + (static_cast<long long>(b) * groups + g) * 2);
    s += v.x;
    q += v.y;
  }
  sh_s[threadIdx.x] = s;
  sh_q[threadIdx.x] = q;
  __syncthreads();
#pragma unroll
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      sh_s[threadIdx.x] += sh_s[threadIdx.x + o];
      sh_q[threadIdx.x] += sh_q[threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double mean = sh_s[0] / count;
    double var = sh_q[0] / count - mean * mean;
    if (var < 0.0) var = 0.0;
    stats[g * 2] = static_cast<float>(mean);
    stats[g * 2 + 1] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
  }
}

// One image row (t, y) per block iteration; 256 % (C/8) == 0, so every thread keeps ONE 8-channel column for the
// whole kernel and the affine coefficients a = rstd*gamma, b = beta - mean*a are computed once per thread.  All
// divisions are per ROW (frame / row / nearest-neighbour source row of the SpatialNorm3D latent); the per-voxel
// index math is shifts and adds (the x upsampling ratio is a power of two in the CogVideoX decoder; a generic
// divide is kept for other ratios).
template <int U>
__global__ void __launch_bounds__(256) gn_apply_kernel(const bf16* __restrict__ x, bf16* __restrict__ out, int T,
                                                       int H, int W, int C, int groups,
                                                       const float* __restrict__ stats,
                                                       const bf16* __restrict__ gamma, const bf16* __restrict__ beta,
                                                       int apply_silu, const bf16* __restrict__ zy,
                                                       const bf16* __restrict__ zb, int Tz, int hz, int wz,
                                                       int x_shift) {
  const int vcols = C >> 3;
  const int cpg = C / groups;
  const int vcol = threadIdx.x % vcols;
  const int c0 = vcol * 8;
  float a[8], b[8];
  {
    float g[8], bt[8];
    unpack8(*reinterpret_cast<const uint4*>(gamma + c0), g);
    unpack8(*reinterpret_cast<const uint4*>(beta + c0), bt);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int grp = (c0 + k) / cpg;
      const float mean = stats[grp * 2], rstd = stats[grp * 2 + 1];
      a[k] = rstd * g[k];
      b[k] = bt[k] - mean * a[k];
    }
  }
  const int vper = 256 / vcols;                 // voxels of a row covered per block pass
  const int xw0 = threadIdx.x / vcols;
  const int rows = T * H;
  for (int row = blockIdx.x; row < rows; row += gridDim.x) {
    const int tf = row / H, yh = row - tf * H;
    const bf16* xr = x + static_cast<long long>(row) * W * C + c0;
    bf16* orow = out + static_cast<long long>(row) * W * C + c0;
    const bf16* zyr = nullptr;
    const bf16* zbr = nullptr;
    if (zy) {
      int tz;
      if (T > 1 && (T & 1)) tz = (tf == 0) ? 0 : 1 + ((tf - 1) * (Tz - 1)) / (T - 1);
      else tz = (tf * Tz) / T;
      const int yz = (yh * hz) / H;
      const long long zo = (static_cast<long long>(tz) * hz + yz) * wz * C + c0;
      zyr = zy + zo;
      zbr = zb + zo;
    }
    for (int xw = xw0; xw < W; xw += vper * U) {
      uint4 raw[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int xx = xw + u * vper;
        if (xx < W) raw[u] = *reinterpret_cast<const uint4*>(xr + static_cast<long long>(xx) * C);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int xx = xw + u * vper;
        if (xx >= W) break;
        float f[8];
        unpack8(raw[u], f);
#pragma unroll
        for (int k = 0; k < 8; ++k) f[k] = bf16_round(f[k] * a[k] + b[k]);
        if (zy) {
          const int xz = x_shift >= 0 ? (xx >> x_shift) : (xx * wz) / W;
          float yv[8], bv[8];
          unpack8(*reinterpret_cast<const uint4*>(zyr + static_cast<long long>(xz) * C), yv);
          unpack8(*reinterpret_cast<const uint4*>(zbr + static_cast<long long>(xz) * C), bv);
#pragma unroll
          for (int k = 0; k < 8; ++k) f[k] = bf16_round(bf16_round(f[k] * yv[k]) + bv[k]);
        }
        if (apply_silu) {
#pragma unroll
          for (int k = 0; k < 8; ++k) f[k] = silu_fast(f[k]);
        }
        *reinterpret_cast<uint4*>(orow + static_cast<long long>(xx) * C) = pack8(f);
      }
    }
  }
}

// SpatialNorm3D apply for the CogVideoX decoder, where the latent is upsampled by G = W / wz in x (G = 1, 2, 4, 8): a
// thread owns ONE 8-channel column and G CONSECUTIVE voxels per step, which all share one conv_y / conv_b vector of the
// latent row — those two vectors are loaded once per G voxels and stay PACKED: `bf16(bf16(n * y) + b)` is exactly what
// HMUL2.BF16 / HADD2.BF16 compute (one rounding of the exact product / sum), so the per-element unpack + fp32 multiply +
// re-round chain of the generic kernel (which made this variant issue-bound at 0.48 of HBM speed) disappears.
template <int G>
__global__ void __launch_bounds__(256) gn_apply_spatial_kernel(const bf16* __restrict__ x, bf16* __restrict__ out, int T,
                                                               int H, int W, int C, int groups,
                                                               const float* __restrict__ stats,
                                                               const bf16* __restrict__ gamma,
                                                               const bf16* __restrict__ beta, int apply_silu,
                                                               const bf16* __restrict__ zy, const bf16* __restrict__ zb,
                                                               int Tz, int hz, int wz) {
  const int vcols = C >> 3;
  const int cpg = C / groups;
  const int vcol = threadIdx.x % vcols;
  const int c0 = vcol * 8;
  float a[8], b[8];
  {
    float g[8], bt[8];
    unpack8(*reinterpret_cast<const uint4*>(gamma + c0), g);
    unpack8(*reinterpret_cast<const uint4*>(beta + c0), bt);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int grp = (c0 + k) / cpg;
      const float mean = stats[grp * 2], rstd = stats[grp * 2 + 1];
      a[k] = rstd * g[k];
      b[k] = bt[k] - mean * a[k];
    }
  }
  const int gper = 256 / vcols;                 // voxel groups of a row covered per block pass
  const int g0 = threadIdx.x / vcols;
  const int rows = T * H;
  for (int row = blockIdx.x; row < rows; row += gridDim.x) {
    const int tf = row / H, yh = row - tf * H;
    int tz;
    if (T > 1 && (T & 1)) tz = (tf == 0) ? 0 : 1 + ((tf - 1) * (Tz - 1)) / (T - 1);
    else tz = (tf * Tz) / T;
    const int yz = (yh * hz) / H;
    const long long zo = (static_cast<long long>(tz) * hz + yz) * wz * C + c0;
    const bf16* xr = x + static_cast<long long>(row) * W * C + c0;
    bf16* orow = out + static_cast<long long>(row) * W * C + c0;
    for (int xz = g0; xz < wz; xz += gper) {
      uint4 raw[G];
#pragma unroll
      for (int u = 0; u < G; ++u) raw[u] = *reinterpret_cast<const uint4*>(xr + static_cast<long long>(xz * G + u) * C);
      const uint4 y4 = *reinterpret_cast<const uint4*>(zy + zo + static_cast<long long>(xz) * C);
      const uint4 b4 = *reinterpret_cast<const uint4*>(zb + zo + static_cast<long long>(xz) * C);
      const uint32_t yw[4] = {y4.x, y4.y, y4.z, y4.w}, bw[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int u = 0; u < G; ++u) {
        float f[8];
        unpack8(raw[u], f);
        uint32_t w4[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          __nv_bfloat162 n2 = __floats2bfloat162_rn(f[2 * k] * a[2 * k] + b[2 * k], f[2 * k + 1] * a[2 * k + 1] + b[2 * k + 1]);
          // _rn variants: never contracted into one fma (the reference rounds the product to bf16 before the add)
          n2 = __hadd2_rn(__hmul2_rn(n2, *reinterpret_cast<const __nv_bfloat162*>(&yw[k])),
                          *reinterpret_cast<const __nv_bfloat162*>(&bw[k]));
          if (apply_silu) {
            const float2 v2 = __bfloat1622float2(n2);
            n2 = __floats2bfloat162_rn(silu_fast(v2.x), silu_fast(v2.y));
          }
          w4[k] = *reinterpret_cast<uint32_t*>(&n2);
        }
        *reinterpret_cast<uint4*>(orow + static_cast<long long>(xz * G + u) * C) = make_uint4(w4[0], w4[1], w4[2], w4[3]);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ resampling
__global__ void time_pool_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, int T, long long fvec) {
  const int odd = T & 1;
  const int Tout = odd ? 1 + (T - 1) / 2 : T / 2;
  const long long total = static_cast<long long>(Tout) * fvec;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int to = static_cast<int>(i / fvec);
    const long long e = i % fvec;
    const uint4* xp = reinterpret_cast<const uint4*>(x);
    if (odd && to == 0) {
      reinterpret_cast<uint4*>(y)[i] = xp[e];
    } else {
      const int t0 = odd ? 1 + 2 * (to - 1) : 2 * to;
      float a[8], b[8];
      unpack8(xp[static_cast<long long>(t0) * fvec + e], a);
      unpack8(xp[static_cast<long long>(t0 + 1) * fvec + e], b);
#pragma unroll
      for (int k = 0; k < 8; ++k) a[k] = (a[k] + b[k]) * 0.5f;
      reinterpret_cast<uint4*>(y)[i] = pack8(a);
    }
  }
}

__global__ void upsample_nearest_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, int T, int H, int W,
                                        int C, int Tout, int time_mode) {
  const int vcols = C >> 3;
  const long long total = static_cast<long long>(Tout) * (2 * H) * (2 * W) * vcols;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int vc = static_cast<int>(i % vcols);
    long long r = i / vcols;
    const int xo = static_cast<int>(r % (2 * W));
    r /= (2 * W);
    const int yo = static_cast<int>(r % (2 * H));
    const int to = static_cast<int>(r / (2 * H));
    int ti;
    if (time_mode == 0) ti = to;                       // no temporal upsampling
    else if (time_mode == 1) ti = to / 2;              // T even: x2
    else ti = (to == 0) ? 0 : 1 + (to - 1) / 2;        // T odd > 1: first frame kept single
    const long long src = ((static_cast<long long>(ti) * H + yo / 2) * W + xo / 2) * vcols + vc;
    reinterpret_cast<uint4*>(y)[i] = reinterpret_cast<const uint4*>(x)[src];
  }
}

// ------------------------------------------------------------------------------------------------ pre-processing
// lr [F, 3, h, w] fp32 (0..255) -> out [3, F, S*h, S*w] fp32 = bilinear(align_corners=False) upscale, then /255*2-1
// (ref: inference_script.py:672-679, F.interpolate(scale_factor=upscale, mode="bilinear") on the 0-255 floats)
__global__ void upscale_normalize_kernel(const float* __restrict__ lr, float* __restrict__ out, int F, int h, int w,
                                         int S) {
  const int H = h * S, W = w * S;
  const long long total = static_cast<long long>(3) * F * H * W;
  const float rs = 1.0f / static_cast<float>(S);
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(i % W);
    const int y = static_cast<int>((i / W) % H);
    const int f = static_cast<int>((i / (static_cast<long long>(W) * H)) % F);
    const int c = static_cast<int>(i / (static_cast<long long>(W) * H * F));
    float sy = rs * (y + 0.5f) - 0.5f, sx = rs * (x + 0.5f) - 0.5f;
    sy = sy < 0.f ? 0.f : sy;
    sx = sx < 0.f ? 0.f : sx;
    const int y0 = static_cast<int>(sy), x0 = static_cast<int>(sx);
    const int y1 = y0 + (y0 < h - 1), x1 = x0 + (x0 < w - 1);
    const float ly = sy - y0, lx = sx - x0, hy = 1.f - ly, hx = 1.f - lx;
    const float* src = lr + (static_cast<long long>(f) * 3 + c) * h * w;
    const float v = hy * (hx * src[y0 * w + x0] + lx * src[y0 * w + x1]) +
                    ly * (hx * src[y1 * w + x0] + lx * src[y1 * w + x1]);
    out[i] = v / 255.0f * 2.0f - 1.0f;
  }
}

// ------------------------------------------------------------------------------------------------ tile blend
// b[o, p, q, c] = bf16(bf16(a[o, P_a - e + p, q, c]*(1 - p/e)) + bf16(b[o, p, q, c]*(p/e)))  for p < e  (in place)
// generic strides: (outer, blended axis, other axis, inner) — serves blend_v and blend_h on channels-last latents and
// on planar pixel tiles.
__global__ void blend_kernel(const bf16* __restrict__ a, bf16* __restrict__ b, int outer, int e, int other,
                             int inner, long long a_so, long long a_sp, long long a_sq, int a_len,
                             long long b_so, long long b_sp, long long b_sq) {
  const long long total = static_cast<long long>(outer) * e * other * inner;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % inner);
    long long r = i / inner;
    const int q = static_cast<int>(r % other);
    r /= other;
    const int pp = static_cast<int>(r % e);
    const int o = static_cast<int>(r / e);
    const float wb = static_cast<float>(static_cast<double>(pp) / e);
    const float wa = static_cast<float>(1.0 - static_cast<double>(pp) / e);
    const float av = __bfloat162float(a[o * a_so + (a_len - e + pp) * a_sp + q * a_sq + c]);
    bf16* bp = b + o * b_so + pp * b_sp + q * b_sq + c;
    const float bv = __bfloat162float(*bp);
    *bp = __float2bfloat16_rn(bf16_round(av * wa) + bf16_round(bv * wb));
  }
}

// ------------------------------------------------------------------------------------------------ layout glue
template <typename TIn>
__global__ void pixels_to_cl_kernel(const TIn* __restrict__ x, bf16* __restrict__ y, long long nvox, int Cpad) {
  for (long long v = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; v < nvox;
       v += static_cast<long long>(gridDim.x) * blockDim.x) {
    float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    f[0] = static_cast<float>(x[v]);
    f[1] = static_cast<float>(x[nvox + v]);
    f[2] = static_cast<float>(x[2 * nvox + v]);
    uint4* yp = reinterpret_cast<uint4*>(y + v * Cpad);
    yp[0] = pack8(f);
    const uint4 z = make_uint4(0, 0, 0, 0);
    for (int k = 1; k < Cpad / 8; ++k) yp[k] = z;
  }
}

__global__ void ncthw_to_cl_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, int C, long long nvox,
                                   int Cpad, float scale) {
  const long long total = nvox * Cpad;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % Cpad);
    const long long v = i / Cpad;
    float val = 0.f;
    if (c < C) {
      val = __bfloat162float(x[static_cast<long long>(c) * nvox + v]);
      if (scale != 1.0f) val = val * scale;
    }
    y[i] = __float2bfloat16_rn(val);
  }
}

__global__ void cl_to_ncthw_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, int C, long long nvox,
                                   int ldx) {
  const long long total = nvox * C;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long v = i % nvox;
    const int c = static_cast<int>(i / nvox);
    y[i] = x[v * ldx + c];
  }
}

__global__ void gaussian_sample_kernel(const bf16* __restrict__ mom, const bf16* __restrict__ noise,
                                       bf16* __restrict__ z, long long nvox, float scaling) {
  const long long total = nvox * 16;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long v = i % nvox;
    const int c = static_cast<int>(i / nvox);
    const float mean = __bfloat162float(mom[v * 32 + c]);
    float lv = __bfloat162float(mom[v * 32 + 16 + c]);
    lv = fminf(fmaxf(lv, -30.0f), 20.0f);
    const float stdv = bf16_round(expf(bf16_round(0.5f * lv)));
    const float s = bf16_round(mean + bf16_round(stdv * __bfloat162float(noise[i])));
    z[i] = __float2bfloat16_rn(s * scaling);
  }
}

__global__ void post_scale_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, long long n) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float v = bf16_round(bf16_round(__bfloat162float(x[i]) * 0.5f) + 0.5f);
    y[i] = __float2bfloat16_rn(fminf(fmaxf(v, 0.0f), 1.0f));
  }
}

}  // namespace dove

using namespace dove;
#define ST(s) static_cast<cudaStream_t>(s)

extern "C" int dove_layernorm_mod_bf16(const void* x, void* out, int rows, int D, const void* ln_w,
                                       const void* ln_b, float eps, const void* scale0, const void* shift0,
                                       const void* scale1, const void* shift1, int split_row, void* stream) {
  if (int e = ensure_init()) return e;
  DOVE_CHECK_ARG(rows > 0 && D % 256 == 0 && D <= 4096, "layernorm: D=%d must be a multiple of 256 and <= 4096", D);
  DOVE_CHECK_ARG(ln_w && ln_b, "layernorm: affine parameters required");
  DOVE_CHECK_ARG((scale0 == nullptr) == (shift0 == nullptr), "layernorm: scale/shift must come together");
  if (scale0 && !scale1) { scale1 = scale0; shift1 = shift0; }
  const int blocks = grid_for(static_cast<long long>(rows) * 32, 256, 8);
  layernorm_mod_kernel<<<blocks, 256, 0, ST(stream)>>>(
      static_cast<const bf16*>(x), static_cast<bf16*>(out), rows, D, static_cast<const bf16*>(ln_w),
      static_cast<const bf16*>(ln_b), eps, static_cast<const bf16*>(scale0), static_cast<const bf16*>(shift0),
      static_cast<const bf16*>(scale1), static_cast<const bf16*>(shift1), split_row);
  DOVE_LAUNCH_CHECK("layernorm_mod_kernel");
  return DOVE_OK;
}

extern "C" int dove_qk_norm_rope_bf16(void* qkv, int rows, int heads, const void* q_w, const void* q_b,
                                      const void* k_w, const void* k_b, float eps, const float* cos,
                                      const float* sin, int text_len, void* stream) {
  if (int e = ensure_init()) return e;
  DOVE_CHECK_ARG(rows > 0 && heads > 0, "qk_norm_rope: empty problem");
  DOVE_CHECK_ARG(text_len >= rows || (cos && sin), "qk_norm_rope: cos/sin tables required");
  const int blocks = grid_for(static_cast<long long>(rows) * heads * 32, 256, 8);
  qk_norm_rope_kernel<<<blocks, 256, 0, ST(stream)>>>(
      static_cast<bf16*>(qkv), rows, heads, static_cast<const bf16*>(q_w), static_cast<const bf16*>(q_b),
      static_cast<const bf16*>(k_w), static_cast<const bf16*>(k_b), eps, cos, sin, text_len);
  DOVE_LAUNCH_CHECK("qk_norm_rope_kernel");
  return DOVE_OK;
}

extern "C" int dove_gemv_bf16(const void* x, const void* W, const void* b, void* y, int N, int K, int silu_in,
                              void* stream) {
  if (int e = ensure_init()) return e;
  DOVE_CHECK_ARG(N > 0 && K > 0 && K % 2 == 0, "gemv: bad shape N=%d K=%d", N, K);
  gemv_kernel<<<(N + 7) / 8, 256, 0, ST(stream)>>>(static_cast<const bf16*>(x), static_cast<const bf16*>(W),
                                                   static_cast<const bf16*>(b), static_cast<bf16*>(y), N, K, silu_in);
  DOVE_LAUNCH_CHECK("gemv_kernel");
  return DOVE_OK;
}

extern "C" int dove_patchify_bf16(const void* latent, void* tokens, int F, int C, int h, int w, void* stream) {
  if (int e = ensure_init()) return e;
  DOVE_CHECK_ARG(F > 0 && F % 2 == 0 && h % 2 == 0 && w % 2 == 0, "patchify: F,h,w must be even (F=%d h=%d w=%d)", F, h, w);
  const long long total = static_cast<long long>(F) * C * h * w / 2;
  patchify_kernel<<<grid_for(total, 256), 256, 0, ST(stream)>>>(static_cast<const bf16*>(latent),
                                                               static_cast<bf16*>(tokens), F, C, h, w);
  DOVE_LAUNCH_CHECK("patchify_kernel");
  return DOVE_OK;
}

extern "C" int dove_unpatchify_velocity_bf16(const void* tokens, const void* latent, void* x0, void* pred_out,
                                             int F, int C, int h, int w, float a, float b, void* stream) {
  if (int e = ensure_init()) return e;
  DOVE_CHECK_ARG(F > 0 && F % 2 == 0 && h % 2 == 0 && w % 2 == 0, "unpatchify: F,h,w must be even");
  DOVE_CHECK_ARG(x0 || pred_out, "unpatchify: no output requested");
  DOVE_CHECK_ARG(!x0 || latent, "unpatchify: latent required for x0");
  const long long total = static_cast<long long>(F) * C * h * w / 2;
  unpatchify_velocity_kernel<<<grid_for(total, 256), 256, 0, ST(stream)>>>(
      static_cast<const bf16*>(tokens), static_cast<const bf16*>(latent), static_cast<bf16*>(x0),
      static_cast<bf16*>(pred_out), F, C, h, w, a, b);
  DOVE_LAUNCH_CHECK("unpatchify_velocity_kernel");
  return DOVE_OK;
}

extern "C" int dove_velocity_bf16(const void* sample, const void* noise, void* out, int64_t n, float a, float b,
                                  void* stream) {
  if (int e = ensure_init()) return e;
  DOVE_CHECK_ARG(n > 0, "velocity: empty");
  velocity_kernel<<<grid_for(n, 256), 256, 0, ST(stream)>>>(static_cast<const bf16*>(sample),
                                                           static_cast<const bf16*>(noise), static_cast<bf16*>(out),
                                                           n, a, b);
  DOVE_LAUNCH_CHECK("velocity_kernel");
  return DOVE_OK;
}

extern "C" size_t dove_gn_partial_floats(int64_t nvox, int groups) {
  (void)nvox;
  return static_cast<size_t>(GN_MAX_BLOCKS) * groups * 2;
}

extern "C" int dove_gn_stats_bf16(const void* x, int64_t nvox, int C, int groups, float eps, float* partial,
                                  float* stats, void* stream) {
  if (int e = ensure_init()) return e;
  DOVE_CHECK_ARG(groups == 32 && C % 128 == 0 && C <= 2048, "gn_stats: groups must be 32 and C a multiple of 128 (C=%d)", C);
  DOVE_CHECK_ARG(nvox > 0, "gn_stats: empty");
  const int vper = 256 / (C >> 3);
  DOVE_CHECK_ARG(vper >= 1 && 256 % (C >> 3) == 0, "gn_stats: unsupported C=%d", C);
  long long blocks = (nvox + vper * 8 - 1) / (vper * 8);
  if (blocks > GN_MAX_BLOCKS) blocks = GN_MAX_BLOCKS;
  if (blocks < 1) blocks = 1;
  gn_partial_kernel<<<static_cast<int>(blocks), 256, 0, ST(stream)>>>(static_cast<const bf16*>(x), nvox, C, groups,
                                                                    partial);
  DOVE_LAUNCH_CHECK("gn_partial_kernel");
  gn_finalize_kernel<<<groups, 256, 0, ST(stream)>>>(partial, static_cast<int>(blocks), groups,
                                              static_cast<double>(nvox) * (C / groups), eps, stats);
  DOVE_LAUNCH_CHECK("gn_finalize_kernel");
  return DOVE_OK;
}

extern "C" int dove_gn_finalize(const float* partial, int64_t nvox, int C, int groups, float eps, float* stats,
                               void* stream) {
  if (int e = ensure_init()) return e;
  DOVE_CHECK_ARG(groups == 32 && C % 32 == 0 && nvox > 0, "gn_finalize: bad shape");
  gn_finalize_kernel<<<groups, 256, 0, ST(stream)>>>(partial, GN_MAX_BLOCKS, groups,
                                                   static_cast<double>(nvox) * (C / groups), eps, stats);
  DOVE_LAUNCH_CHECK("gn_finalize_kernel");
  return DOVE_OK;
}

extern "C" int dove_gn_apply_bf16(const void* x, void* out, int T, int H, int W, int C, int groups,
                                  const float* stats, const void* gamma, const void* beta, int apply_silu,
                                  const void* zq_y, const void* zq_b, int Tz, int hz, int wz, void* stream) {
  if (int e = ensure_init()) return e;
  DOVE_CHECK_ARG(T > 0 && H > 0 && W > 0 && C % 8 == 0 && C % groups == 0, "gn_apply: bad shape");
  DOVE_CHECK_ARG((zq_y == nullptr) == (zq_b == nullptr), "gn_apply: zq_y/zq_b must come together");
  DOVE_CHECK_ARG(256 % (C >> 3) == 0, "gn_apply: C/8 must divide 256 (C=%d)", C);
  DOVE_CHECK_ARG(static_cast<long long>(T) * H * W < (1ll << 31), "gn_apply: too many voxels");
  const long long nvec = static_cast<long long>(T) * H * W * (C >> 3);
  int x_shift = -1;                               // log2(W / wz) when the x ratio is a power of two
  if (zq_y && wz > 0 && W % wz == 0) {
    const int r = W / wz;
    if ((r & (r - 1)) == 0) {
      x_shift = 0;
      while ((1 << x_shift) < r) ++x_shift;
    }
  }
  const int rows = T * H;
  const int blocks = rows < num_sms() * 8 ? rows : num_sms() * 8;
  if (zq_y && x_shift >= 0 && x_shift <= 3 && (wz << x_shift) == W) {     // CogVideoX decoder ratios 1, 2, 4, 8
#define DOVE_GN_SPATIAL(G)                                                                                          \
  gn_apply_spatial_kernel<G><<<blocks, 256, 0, ST(stream)>>>(                                                       \
      static_cast<const bf16*>(x), static_cast<bf16*>(out), T, H, W, C, groups, stats, static_cast<const bf16*>(gamma), \
      static_cast<const bf16*>(beta), apply_silu, static_cast<const bf16*>(zq_y), static_cast<const bf16*>(zq_b), Tz, hz, wz)
    switch (x_shift) {
      case 0: DOVE_GN_SPATIAL(1); break;
      case 1: DOVE_GN_SPATIAL(2); break;
      case 2: DOVE_GN_SPATIAL(4); break;
      default: DOVE_GN_SPATIAL(8); break;
    }
#undef DOVE_GN_SPATIAL
    DOVE_LAUNCH_CHECK("gn_apply_spatial_kernel");
    return DOVE_OK;
  }
  gn_apply_kernel<4><<<blocks, 256, 0, ST(stream)>>>(
      static_cast<const bf16*>(x), static_cast<bf16*>(out), T, H, W, C, groups, stats, static_cast<const bf16*>(gamma),
      static_cast<const bf16*>(beta), apply_silu, static_cast<const bf16*>(zq_y), static_cast<const bf16*>(zq_b), Tz, hz,
      wz, x_shift);
  DOVE_LAUNCH_CHECK("gn_apply_kernel");
  return DOVE_OK;
}

extern "C" int dove_time_pool_bf16(const void* x, void* y, int T, int64_t frame_elems, void* stream) {
  if (int e = ensure_init()) return e;
  DOVE_CHECK_ARG(T >= 1 && frame_elems % 8 == 0, "time_pool: frame_elems must be a multiple of 8");
  const int Tout = (T & 1) ? 1 + (T - 1) / 2 : T / 2;
  const long long fvec = frame_elems / 8;
  time_pool_kernel<<<grid_for(Tout * fvec, 256), 256, 0, ST(stream)>>>(static_cast<const bf16*>(x),
                                                                       static_cast<bf16*>(y), T, fvec);
  DOVE_LAUNCH_CHECK("time_pool_kernel");
  return DOVE_OK;
}

extern "C" int dove_upsample_nearest_bf16(const void* x, void* y, int T, int H, int W, int C, int time_x2,
                                          void* stream) {
  if (int e = ensure_init()) return e;
  DOVE_CHECK_ARG(T >= 1 && C % 8 == 0, "upsample: bad shape");
  int mode = 0, Tout = T;
  if (time_x2 && T > 1) {
    if (T & 1) { mode = 2; Tout = 1 + 2 * (T - 1); }
    else { mode = 1; Tout = 2 * T; }
  }
  const long long total = static_cast<long long>(Tout) * 2 * H * 2 * W * (C >> 3);
  upsample_nearest_kernel<<<grid_for(total, 256), 256, 0, ST(stream)>>>(static_cast<const bf16*>(x),
                                                                        static_cast<bf16*>(y), T, H, W, C, Tout, mode);
  DOVE_LAUNCH_CHECK("upsample_nearest_kernel");
  return DOVE_OK;
}

extern "C" int dove_pixels_to_cl_bf16(const void* x, int x_is_fp32, void* y, int T, int H, int W, int Cpad,
                                      void* stream) {
  if (int e = ensure_init()) return e;
  DOVE_CHECK_ARG(Cpad % 8 == 0 && Cpad >= 8, "pixels_to_cl: Cpad must be a multiple of 8");
  const long long nvox = static_cast<long long>(T) * H * W;
  if (x_is_fp32)
    pixels_to_cl_kernel<float><<<grid_for(nvox, 256), 256, 0, ST(stream)>>>(static_cast<const float*>(x),
                                                                            static_cast<bf16*>(y), nvox, Cpad);
  else
    pixels_to_cl_kernel<bf16><<<grid_for(nvox, 256), 256, 0, ST(stream)>>>(static_cast<const bf16*>(x),
                                                                           static_cast<bf16*>(y), nvox, Cpad);
  DOVE_LAUNCH_CHECK("pixels_to_cl_kernel");
  return DOVE_OK;
}

extern "C" int dove_ncthw_to_cl_bf16(const void* x, void* y, int C, int T, int H, int W, int Cpad, float scale,
                                     void* stream) {
  if (int e = ensure_init()) return e;
  const long long nvox = static_cast<long long>(T) * H * W;
  ncthw_to_cl_kernel<<<grid_for(nvox * Cpad, 256), 256, 0, ST(stream)>>>(static_cast<const bf16*>(x),
                                                                         static_cast<bf16*>(y), C, nvox, Cpad, scale);
  DOVE_LAUNCH_CHECK("ncthw_to_cl_kernel");
  return DOVE_OK;
}

extern "C" int dove_cl_to_ncthw_bf16(const void* x, void* y, int C, int T, int H, int W, int ldx, void* stream) {
  if (int e = ensure_init()) return e;
  const long long nvox = static_cast<long long>(T) * H * W;
  cl_to_ncthw_kernel<<<grid_for(nvox * C, 256), 256, 0, ST(stream)>>>(static_cast<const bf16*>(x),
                                                                      static_cast<bf16*>(y), C, nvox, ldx);
  DOVE_LAUNCH_CHECK("cl_to_ncthw_kernel");
  return DOVE_OK;
}

extern "C" int dove_gaussian_sample_bf16(const void* moments, const void* noise, void* z, int64_t nvox,
                                         float scaling, void* stream) {
  if (int e = ensure_init()) return e;
  gaussian_sample_kernel<<<grid_for(nvox * 16, 256), 256, 0, ST(stream)>>>(
      static_cast<const bf16*>(moments), static_cast<const bf16*>(noise), static_cast<bf16*>(z), nvox, scaling);
  DOVE_LAUNCH_CHECK("gaussian_sample_kernel");
  return DOVE_OK;
}

extern "C" int dove_upscale_normalize_f32(const float* lr, float* out, int F, int h, int w, int scale, void* stream) {
  if (int e = ensure_init()) return e;
  DOVE_CHECK_ARG(F > 0 && h > 0 && w > 0 && scale >= 1, "upscale: bad shape");
  const long long total = static_cast<long long>(3) * F * h * scale * w * scale;
  upscale_normalize_kernel<<<grid_for(total, 256), 256, 0, ST(stream)>>>(lr, out, F, h, w, scale);
  DOVE_LAUNCH_CHECK("upscale_normalize_kernel");
  return DOVE_OK;
}

extern "C" int dove_blend_bf16(const void* a, void* b, int outer, int extent, int other, int inner, int64_t a_so,
                               int64_t a_sp, int64_t a_sq, int a_len, int64_t b_so, int64_t b_sp, int64_t b_sq,
                               void* stream) {
  if (int e = ensure_init()) return e;
  DOVE_CHECK_ARG(outer > 0 && extent > 0 && other > 0 && inner > 0 && a_len >= extent, "blend: bad shape");
  const long long total = static_cast<long long>(outer) * extent * other * inner;
  blend_kernel<<<grid_for(total, 256), 256, 0, ST(stream)>>>(static_cast<const bf16*>(a), static_cast<bf16*>(b), outer,
                                                             extent, other, inner, a_so, a_sp, a_sq, a_len, b_so, b_sp,
                                                             b_sq);
  DOVE_LAUNCH_CHECK("blend_kernel");
  return DOVE_OK;
}

extern "C" int dove_post_scale_bf16(const void* x, void* y, int64_t n, void* stream) {
  if (int e = ensure_init()) return e;
  post_scale_kernel<<<grid_for(n, 256), 256, 0, ST(stream)>>>(static_cast<const bf16*>(x), static_cast<bf16*>(y), n);
  DOVE_LAUNCH_CHECK("post_scale_kernel");
  return DOVE_OK;
}
