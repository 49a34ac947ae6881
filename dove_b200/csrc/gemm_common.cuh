// Shared pieces of the tcgen05 GEMM / implicit-GEMM conv kernels: parameters, tile-order mapping and the fused
// epilogue applied to one chunk of CH accumulator columns held in registers.
#pragma once
#include "common.cuh"
#include "ptx.cuh"

namespace dove {

struct GemmParams {
  int num_m_tiles, num_n_tiles, num_kb;
  int M;   // dense: valid rows
  // conv geometry
  int tw, th, tiles_w, tiles_h;
  int band_h, To;   // L2-friendly tile order: bands of band_h tile-rows, all To frames of a band before the next
  int Ho, Wo;
  int kh, kw, cin_blocks;
  int stride, pad;
  int t_shift, has_prev;   // causal convs on an un-padded input: input frame = t + dt - t_shift; frames < 0 come
                           // from the `prev` tensor map (the conv cache) or, without one, replicate frame 0
  // epilogue
  int epi;
  bf16* C;
  long long ldc;
  const bf16* bias;
  const bf16* aux;
  long long ld_aux;
  const bf16* gate0;
  const bf16* gate1;
  int split_row;
  int n_valid;    // columns stored
  int out_mode;   // 0 row-major [rows, ldc]; planar [n][rows_total]: 1 bf16, 2 bf16 post-scaled to [0,1], 3 uint8
  long long rows_total;
  // DOVE_EPI_QKV_NORM_ROPE (dense CTA-pair GEMM only): per-head LayerNorm(64) on the q / k column blocks + 3-D RoPE
  const bf16* qk_w[2];   // norm_q.weight, norm_k.weight  [64]
  const bf16* qk_b[2];   // norm_q.bias,   norm_k.bias    [64]
  const float* rope_cos; // TRANSPOSED, pair-deduplicated tables [32][rope_ld] fp32 (null when text_len >= rows)
  const float* rope_sin;
  long long rope_ld;     // = rows - text_len
  int text_len, heads;
  float qk_eps;
  // fused GroupNorm statistics of the OUTPUT (consumed by the next GroupNorm): per-CTA partial sums
  float* gn_partial;   // [GN_PARTIAL_ROWS][32 groups][2] (sum, sum of squares) or nullptr
  int gn_cpg;          // channels per group of the output tensor
};

constexpr int GN_PARTIAL_ROWS = 1184;

__device__ __forceinline__ float gelu_tanh_f(float x) {
  // 0.5*x*(1+tanh(sqrt(2/pi)*(x+0.044715x^3))), tanh(y) = 1 - 2/(exp(2y)+1)
  const float kBeta = 0.7978845608028654f, kKappa = 0.044715f;
  float inner = kBeta * (x + kKappa * x * x * x);
  float t = 1.0f - __fdividef(2.0f, __expf(2.0f * inner) + 1.0f);
  return 0.5f * x * (1.0f + t);
}

// conv m-tile index -> (frame t, tile row hy, tile col wx).  Tiles are ordered band-major: a band is `band_h` tile
// rows; within a band all frames are visited before moving on, so the 3 output frames that share an input
// frame (and the 3 tile rows that share an input row) are processed while that input is still in L2.
__device__ __forceinline__ void conv_tile_coords(const GemmParams& p, int mt, int& t, int& hy, int& wx) {
  const int full_band = p.band_h * p.tiles_w * p.To;
  const int b = mt / full_band;
  const int r = mt - b * full_band;
  const int gb = min(p.band_h, p.tiles_h - b * p.band_h);
  const int pb = gb * p.tiles_w;
  t = r / pb;
  const int rr = r - t * pb;
  hy = b * p.band_h + rr / p.tiles_w;
  wx = rr % p.tiles_w;
}

// Source of input frame (t + dt) of a causal conv: the current frame batch, the cached 2 frames, or frame 0.
__device__ __forceinline__ const CUtensorMap* conv_frame_src(const GemmParams& p, const CUtensorMap* cur,
                                                             const CUtensorMap* prev, int& f) {
  f -= p.t_shift;
  if (f >= 0) return cur;
  if (p.has_prev) {
    f += 2;
    return prev;
  }
  f = 0;
  return cur;
}

// Fused epilogue for CH consecutive accumulator columns (n0 .. n0+CH) of output row `row`.
// kStats: also accumulates, per quad of 4 consecutive channels, the sum / sum of squares of the bf16-ROUNDED
// outputs into qs[CH/4], qss[CH/4] (fused GroupNorm statistics).
template <int CH, bool kStats = false>
__device__ __forceinline__ void epilogue_chunk(const GemmParams& p, const uint32_t* v, long long row, int n0,
                                               const bf16* gate, float* qs = nullptr, float* qss = nullptr) {
  float r[CH];
  if (p.bias) {
    const uint4* bp = reinterpret_cast<const uint4*>(p.bias + n0);
#pragma unroll
    for (int j = 0; j < CH / 8; ++j) {
      const uint4 b4 = bp[j];
      const uint32_t bu[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 b2 = unpack_bf16x2(bu[k]);
        r[j * 8 + 2 * k] = bf16_round(__uint_as_float(v[j * 8 + 2 * k]) + b2.x);
        r[j * 8 + 2 * k + 1] = bf16_round(__uint_as_float(v[j * 8 + 2 * k + 1]) + b2.y);
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < CH; ++i) r[i] = bf16_round(__uint_as_float(v[i]));
  }
  if (p.epi == DOVE_EPI_GELU_TANH) {
#pragma unroll
    for (int i = 0; i < CH; ++i) r[i] = gelu_tanh_f(r[i]);
  } else if (p.epi == DOVE_EPI_GATED_RES || p.epi == DOVE_EPI_ADD) {
    const uint4* ap = reinterpret_cast<const uint4*>(p.aux + row * p.ld_aux + n0);
    const uint4* gp = reinterpret_cast<const uint4*>((gate ? gate : p.aux) + n0);
#pragma unroll
    for (int j = 0; j < CH / 8; ++j) {
      const uint4 a4 = ap[j];
      const uint32_t au[4] = {a4.x, a4.y, a4.z, a4.w};
      uint4 g4 = make_uint4(0, 0, 0, 0);
      if (p.epi == DOVE_EPI_GATED_RES) g4 = gp[j];
      const uint32_t gu[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 a2 = unpack_bf16x2(au[k]);
        const int i = j * 8 + k * 2;
        if (p.epi == DOVE_EPI_GATED_RES) {
          const float2 g2 = unpack_bf16x2(gu[k]);
          r[i] = a2.x + bf16_round(g2.x * r[i]);
          r[i + 1] = a2.y + bf16_round(g2.y * r[i + 1]);
        } else {
          r[i] += a2.x;
          r[i + 1] += a2.y;
        }
      }
    }
  }
  if (kStats) {
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      const float rr = bf16_round(r[i]);
      qs[i >> 2] += rr;
      qss[i >> 2] += rr * rr;
    }
  }
  if (p.out_mode == 0 && n0 + CH <= p.n_valid) {
    uint4* cp = reinterpret_cast<uint4*>(p.C + row * p.ldc + n0);
#pragma unroll
    for (int j = 0; j < CH / 8; ++j) {
      uint4 o;
      o.x = pack_bf16x2(r[j * 8 + 0], r[j * 8 + 1]);
      o.y = pack_bf16x2(r[j * 8 + 2], r[j * 8 + 3]);
      o.z = pack_bf16x2(r[j * 8 + 4], r[j * 8 + 5]);
      o.w = pack_bf16x2(r[j * 8 + 6], r[j * 8 + 7]);
      cp[j] = o;
    }
  } else {
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      const int n = n0 + i;
      if (n < p.n_valid) {
        if (p.out_mode == 0) {
          p.C[row * p.ldc + n] = __float2bfloat16_rn(r[i]);
        } else {
          const long long o = static_cast<long long>(n) * p.rows_total + row;
          if (p.out_mode == 1) {
            p.C[o] = __float2bfloat16_rn(r[i]);
          } else {   // fused post-processing of the last decoder conv: (x*0.5+0.5).clamp(0,1) with bf16 rounding points
            const float v = fminf(fmaxf(bf16_round(bf16_round(bf16_round(r[i]) * 0.5f) + 0.5f), 0.0f), 1.0f);
            if (p.out_mode == 2) p.C[o] = __float2bfloat16_rn(v);
            else reinterpret_cast<uint8_t*>(p.C)[o] = static_cast<uint8_t>(v * 255.0f);   // trunc, as .to(uint8)
          }
        }
      }
    }
  }
}

// Epilogue of the fused QKV projection for ONE head (64 consecutive output columns col0.. of row `row`):
//   r = bf16(acc + bias)                                           (nn.Linear to_q / to_k / to_v)
//   q, k heads:  y = bf16(LayerNorm_64(r; w, b, eps))              (Attention.norm_q / norm_k, fp32 statistics)
//                rows >= text_len:  y = bf16(y*cos + rot(y)*sin)   (apply_rotary_emb, fp32, interleaved pairs)
//   v heads:     y = r
// One thread owns the whole head of its row, so the statistics need no shuffles.
__device__ __forceinline__ void qkv_head_epilogue(const GemmParams& p, const uint32_t* v, long long row, int col0) {
  float r[64];
  {
    const uint4* bp = reinterpret_cast<const uint4*>(p.bias + col0);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint4 b4 = bp[j];
      const uint32_t bu[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 b2 = unpack_bf16x2(bu[k]);
        r[j * 8 + 2 * k] = bf16_round(__uint_as_float(v[j * 8 + 2 * k]) + b2.x);
        r[j * 8 + 2 * k + 1] = bf16_round(__uint_as_float(v[j * 8 + 2 * k + 1]) + b2.y);
      }
    }
  }
  const int which = col0 / (p.heads * 64);          // 0 = q, 1 = k, 2 = v
  if (which < 2) {
    float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 64; ++i) s4[i & 3] += r[i];
    const float mean = ((s4[0] + s4[1]) + (s4[2] + s4[3])) * (1.0f / 64.0f);
    float q4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 64; ++i) {
      r[i] -= mean;
      q4[i & 3] += r[i] * r[i];
    }
    const float rstd = rsqrtf(((q4[0] + q4[1]) + (q4[2] + q4[3])) * (1.0f / 64.0f) + p.qk_eps);
    const uint4* wp = reinterpret_cast<const uint4*>(p.qk_w[which]);
    const uint4* bp = reinterpret_cast<const uint4*>(p.qk_b[which]);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint4 w4 = wp[j], b4 = bp[j];
      const uint32_t wu[4] = {w4.x, w4.y, w4.z, w4.w}, bu[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 w2 = unpack_bf16x2(wu[k]), b2 = unpack_bf16x2(bu[k]);
        const int i = j * 8 + 2 * k;
        r[i] = bf16_round(r[i] * rstd * w2.x + b2.x);
        r[i + 1] = bf16_round(r[i + 1] * rstd * w2.y + b2.y);
      }
    }
  }
  // RoPE (rows >= text_len of q / k heads) fused with the store, 8 pairs at a time.  The tables arrive TRANSPOSED and
  // pair-deduplicated, [32][nv] fp32 (cos[i][token] = cos_table[token][2i] = cos_table[token][2i+1]), so that the 32
  // lanes of a warp (32 consecutive rows) read 128 contiguous bytes per column instead of 32 rows 256 B apart.
  const bool rope = which < 2 && row >= p.text_len;
  const float* cp = p.rope_cos + (row - p.text_len);
  const float* sp = p.rope_sin + (row - p.text_len);
  uint4* op = reinterpret_cast<uint4*>(p.C + row * p.ldc + col0);
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    if (rope) {      // out = x*cos + rot(x)*sin, rot(x)[2i] = -x[2i+1], rot(x)[2i+1] = x[2i]  (fp32, no fma)
      float c[8], sn[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        c[k] = cp[static_cast<long long>(g * 8 + k) * p.rope_ld];
        sn[k] = sp[static_cast<long long>(g * 8 + k) * p.rope_ld];
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int i = g * 16 + 2 * k;
        const float a0 = r[i], a1 = r[i + 1];
        r[i] = __fadd_rn(__fmul_rn(a0, c[k]), __fmul_rn(-a1, sn[k]));
        r[i + 1] = __fadd_rn(__fmul_rn(a1, c[k]), __fmul_rn(a0, sn[k]));
      }
    }
#pragma unroll
    for (int h2 = 0; h2 < 2; ++h2) {
      const int i = g * 16 + h2 * 8;
      uint4 o;
      o.x = pack_bf16x2(r[i + 0], r[i + 1]);
      o.y = pack_bf16x2(r[i + 2], r[i + 3]);
      o.z = pack_bf16x2(r[i + 4], r[i + 5]);
      o.w = pack_bf16x2(r[i + 6], r[i + 7]);
      op[g * 2 + h2] = o;
    }
  }
}

}  // namespace dove
