// Swin-T backbone kernels (modules/swin_transformer.py) on the haloed NHWC token grid: patch embedding
// (+LayerNorm), LayerNorm, shifted-window attention, patch merging (+LayerNorm).  The linear layers
// (qkv, proj, fc1+GELU, fc2, reduction) run as 1x1 convolutions on the tcgen05 / CUDA-core conv
// kernels, with the residual adds in their epilogues.
#include "layers.cuh"
#include <math.h>
#include <stdlib.h>
#include <algorithm>

namespace yb {

template <typename T> struct Tok {
  static __device__ __forceinline__ float ld(const T* p) { return (float)*p; }
};
template <> struct Tok<__half> { static __device__ __forceinline__ float ld(const __half* p) { return __half2float(*p); } };
template <> struct Tok<__nv_bfloat16> { static __device__ __forceinline__ float ld(const __nv_bfloat16* p) { return __bfloat162float(*p); } };
template <typename T> __device__ __forceinline__ void tok_st(T* p, float v);
template <> __device__ __forceinline__ void tok_st<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void tok_st<__half>(__half* p, float v) { *p = __float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f)); }
template <> __device__ __forceinline__ void tok_st<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------------------------------------
// PatchEmbed: conv 4x4 s4 (3->96, bias) on the zero-padded image + LayerNorm(96)
// (modules/swin_transformer.py:419-433).  One warp per token, lane l owns channels l, l+32, l+64.
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) k_patch_embed(const float* __restrict__ img, const float* __restrict__ w /*[48][96]*/,
                                                     const float* __restrict__ bias, const float* __restrict__ g,
                                                     const float* __restrict__ be, T* __restrict__ out, int B, int S, int Hg) {
  __shared__ float s_w[48 * 96];
  for (int i = threadIdx.x; i < 48 * 96; i += 256) s_w[i] = w[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int Hp = Hg + 2;
  const int b = blockIdx.y;                                        // 32-bit index math only
  for (int ti = blockIdx.x * 8 + (threadIdx.x >> 5); ti < Hp * Hp; ti += gridDim.x * 8) {
    const int yp = ti / Hp, xp = ti - yp * Hp;
    const size_t tok = (size_t)b * Hp * Hp + ti;
    T* o = out + tok * 96;
    if (yp == 0 || yp == Hg + 1 || xp == 0 || xp == Hg + 1) {
      for (int c = lane; c < 96; c += 32) tok_st<T>(o + c, 0.f);
      continue;
    }
    const int y0 = (yp - 1) * 4, x0 = (xp - 1) * 4;
    float a0 = bias[lane], a1 = bias[lane + 32], a2 = bias[lane + 64];
    // the token's 48 inputs: lane l fetches input l (and l+32 for l < 16), then shuffle-broadcast
    float in0, in1 = 0.f;
    {
      const int k = lane, ci = k >> 4, r = (k >> 2) & 3, q = k & 3;
      in0 = (y0 + r < S && x0 + q < S) ? __ldg(img + (((size_t)b * 3 + ci) * S + y0 + r) * S + x0 + q) : 0.f;
      if (lane < 16) {
        const int k2 = lane + 32, c2 = k2 >> 4, r2 = (k2 >> 2) & 3, q2 = k2 & 3;
        in1 = (y0 + r2 < S && x0 + q2 < S) ? __ldg(img + (((size_t)b * 3 + c2) * S + y0 + r2) * S + x0 + q2) : 0.f;
      }
    }
#pragma unroll
    for (int k = 0; k < 48; ++k) {
      const float v = __shfl_sync(0xffffffffu, k < 32 ? in0 : in1, k & 31);
      a0 = fmaf(v, s_w[k * 96 + lane], a0);
      a1 = fmaf(v, s_w[k * 96 + lane + 32], a1);
      a2 = fmaf(v, s_w[k * 96 + lane + 64], a2);
    }
    const float mean = warp_sum(a0 + a1 + a2) * (1.f / 96.f);
    const float d0 = a0 - mean, d1 = a1 - mean, d2 = a2 - mean;
    const float rstd = rsqrtf(warp_sum(d0 * d0 + d1 * d1 + d2 * d2) * (1.f / 96.f) + 1e-5f);
    tok_st<T>(o + lane, d0 * rstd * g[lane] + be[lane]);
    tok_st<T>(o + lane + 32, d1 * rstd * g[lane + 32] + be[lane + 32]);
    tok_st<T>(o + lane + 64, d2 * rstd * g[lane + 64] + be[lane + 64]);
  }
}

int launch_patch_embed(const float* img, const float* w, const float* bias, const float* g, const float* be, void* out, int dt,
                       int B, int S, int Hg, cudaStream_t s) {
  dim3 grid(std::min(((Hg + 2) * (Hg + 2) + 7) / 8, 148 * 4), B);
  YB_DISPATCH_DT(dt, (k_patch_embed<T><<<grid, 256, 0, s>>>(img, w, bias, g, be, (T*)out, B, S, Hg)));
  YB_CHECK_LAUNCH();
  return YB_OK;
}

// ------------------------------------------------------------------------------------------------
// LayerNorm over C channels of every token of a haloed grid (halo rows written as zeros).
// One warp per token, two-pass statistics in fp32 (what torch's layer_norm computes).
// ------------------------------------------------------------------------------------------------
// 16-byte vector of tokens' channels
template <typename T> struct TokVec;
template <> struct TokVec<float> {
  static constexpr int N = 4;
  static __device__ __forceinline__ void ld(const float* p, float* f) { const float4 v = *reinterpret_cast<const float4*>(p); f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w; }
  static __device__ __forceinline__ void st(float* p, const float* f) { *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]); }
};
template <> struct TokVec<__half> {
  static constexpr int N = 8;
  static __device__ __forceinline__ void ld(const __half* p, float* f) {
    const uint4 v = *reinterpret_cast<const uint4*>(p); const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float2 t = __half22float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
  }
  static __device__ __forceinline__ void st(__half* p, const float* f) {
    uint4 v; __half2* h = reinterpret_cast<__half2*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(fminf(fmaxf(f[2 * i], -65504.f), 65504.f), fminf(fmaxf(f[2 * i + 1], -65504.f), 65504.f));
    *reinterpret_cast<uint4*>(p) = v;
  }
};
template <> struct TokVec<__nv_bfloat16> {
  static constexpr int N = 8;
  static __device__ __forceinline__ void ld(const __nv_bfloat16* p, float* f) {
    const uint4 v = *reinterpret_cast<const uint4*>(p); const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float2 t = __bfloat1622float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
  }
  static __device__ __forceinline__ void st(__nv_bfloat16* p, const float* f) {
    uint4 v; __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = v;
  }
};

// G lanes per token (16 or 32), NV 16-byte vectors per lane: one global read, statistics in registers
template <typename T, int NV, int G>
__global__ void __launch_bounds__(256) k_layernorm(const T* __restrict__ in, T* __restrict__ out, const float* __restrict__ g,
                                                   const float* __restrict__ be, int B, int C, int H) {
  constexpr int N = TokVec<T>::N, TPW = 32 / G;
  const int lane = threadIdx.x & 31, sub = lane % G;
  const int Hp = H + 2, nvec = C / N;
  const int per_img = Hp * Hp, b = blockIdx.y;
  for (int t0 = (blockIdx.x * 8 + (threadIdx.x >> 5)) * TPW; t0 < per_img; t0 += gridDim.x * 8 * TPW) {
    const int ti = t0 + lane / G;
    const bool live = ti < per_img;
    const int yp = ti / Hp, xp = ti - yp * Hp;
    const size_t tok = (size_t)b * per_img + ti;
    const T* x = in + tok * C;
    T* o = out + tok * C;
    const bool halo = yp == 0 || yp == H + 1 || xp == 0 || xp == H + 1;
    float buf[NV][N];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int v = sub + G * i;
      if (live && v < nvec && !halo) {
        TokVec<T>::ld(x + v * N, buf[i]);
#pragma unroll
        for (int e = 0; e < N; ++e) s += buf[i][e];
      } else {
#pragma unroll
        for (int e = 0; e < N; ++e) buf[i][e] = 0.f;
      }
    }
#pragma unroll
    for (int o2 = G / 2; o2 > 0; o2 >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o2);
    const float mean = s / (float)C;
    float var = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i)
      if (sub + G * i < nvec) {
#pragma unroll
        for (int e = 0; e < N; ++e) { const float d = buf[i][e] - mean; var += d * d; }
      }
#pragma unroll
    for (int o2 = G / 2; o2 > 0; o2 >>= 1) var += __shfl_xor_sync(0xffffffffu, var, o2);
    const float rstd = rsqrtf(var / (float)C + 1e-5f);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int v = sub + G * i;
      if (live && v < nvec) {
        float r[N];
#pragma unroll
        for (int e = 0; e < N; ++e) r[e] = halo ? 0.f : (buf[i][e] - mean) * rstd * __ldg(g + v * N + e) + __ldg(be + v * N + e);
        TokVec<T>::st(o + v * N, r);
      }
    }
  }
}

template <typename T>
static void layernorm_dispatch(const T* in, T* out, const float* g, const float* be, int B, int C, int H, int nvec, dim3 blocks, cudaStream_t s) {
  if (nvec <= 16) k_layernorm<T, 1, 16><<<blocks, 256, 0, s>>>(in, out, g, be, B, C, H);
  else if (nvec <= 32) k_layernorm<T, 1, 32><<<blocks, 256, 0, s>>>(in, out, g, be, B, C, H);
  else if (nvec <= 64) k_layernorm<T, 2, 32><<<blocks, 256, 0, s>>>(in, out, g, be, B, C, H);
  else if (nvec <= 96) k_layernorm<T, 3, 32><<<blocks, 256, 0, s>>>(in, out, g, be, B, C, H);
  else k_layernorm<T, 6, 32><<<blocks, 256, 0, s>>>(in, out, g, be, B, C, H);
}

int launch_layernorm(const void* in, void* out, const float* g, const float* be, int dt, int B, int C, int H, cudaStream_t s) {
  const int nvec = C / (dt == DT_F32 ? 4 : 8);
  YB_REQUIRE(C % (dt == DT_F32 ? 4 : 8) == 0 && nvec <= 192, YB_ERR_UNSUPPORTED, "layernorm: C=%d", C);
  const dim3 blocks(std::min(((H + 2) * (H + 2) + 7) / 8, 148 * 4), B);
  YB_DISPATCH_DT(dt, (layernorm_dispatch<T>((const T*)in, (T*)out, g, be, B, C, H, nvec, blocks, s)));
  YB_CHECK_LAUNCH();
  return YB_OK;
}

// ------------------------------------------------------------------------------------------------
// PatchMerging gather + LayerNorm(4C) (modules/swin_transformer.py:299-325): output token (y,x)
// concatenates input tokens (2y,2x), (2y+1,2x), (2y,2x+1), (2y+1,2x+1) (zero beyond an odd edge).
// ------------------------------------------------------------------------------------------------
template <typename T, int NV>
__global__ void __launch_bounds__(256) k_patch_merge_ln(const T* __restrict__ in, T* __restrict__ out, const float* __restrict__ g,
                                                        const float* __restrict__ be, int B, int C, int Hin, int Hout) {
  constexpr int N = TokVec<T>::N;
  const int lane = threadIdx.x & 31;
  const int Hpi = Hin + 2, Hpo = Hout + 2, C4 = 4 * C, nvec = C4 / N, vpp = C / N;     // vectors per source token
  const int b = blockIdx.y;
  for (int ti = blockIdx.x * 8 + (threadIdx.x >> 5); ti < Hpo * Hpo; ti += gridDim.x * 8) {
    const int yp = ti / Hpo, xp = ti - yp * Hpo;
    const size_t tok = (size_t)b * Hpo * Hpo + ti;
    T* o = out + tok * C4;
    const bool halo = yp == 0 || yp == Hout + 1 || xp == 0 || xp == Hout + 1;
    const int y = yp - 1, x = xp - 1;
    float buf[NV][N];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int v = lane + 32 * i;
      bool have = false;
      if (v < nvec && !halo) {
        const int part = v / vpp, cv = v - part * vpp;
        const int iy = 2 * y + (part & 1), ix = 2 * x + (part >> 1);
        if (iy < Hin && ix < Hin) { TokVec<T>::ld(in + (((size_t)b * Hpi + iy + 1) * Hpi + ix + 1) * C + cv * N, buf[i]); have = true; }
      }
      if (!have) {
#pragma unroll
        for (int e = 0; e < N; ++e) buf[i][e] = 0.f;
      }
#pragma unroll
      for (int e = 0; e < N; ++e) s += buf[i][e];
    }
    const float mean = warp_sum(s) / (float)C4;
    float var = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i)
      if (lane + 32 * i < nvec) {
#pragma unroll
        for (int e = 0; e < N; ++e) { const float d = buf[i][e] - mean; var += d * d; }
      }
    const float rstd = rsqrtf(warp_sum(var) / (float)C4 + 1e-5f);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int v = lane + 32 * i;
      if (v < nvec) {
        float r[N];
#pragma unroll
        for (int e = 0; e < N; ++e) r[e] = halo ? 0.f : (buf[i][e] - mean) * rstd * __ldg(g + v * N + e) + __ldg(be + v * N + e);
        TokVec<T>::st(o + v * N, r);
      }
    }
  }
}

template <typename T>
static void patch_merge_dispatch(const T* in, T* out, const float* g, const float* be, int B, int C, int Hin, int Hout, int nvec, dim3 blocks,
                                 cudaStream_t s) {
  if (nvec <= 64) k_patch_merge_ln<T, 2><<<blocks, 256, 0, s>>>(in, out, g, be, B, C, Hin, Hout);
  else if (nvec <= 96) k_patch_merge_ln<T, 3><<<blocks, 256, 0, s>>>(in, out, g, be, B, C, Hin, Hout);
  else if (nvec <= 192) k_patch_merge_ln<T, 6><<<blocks, 256, 0, s>>>(in, out, g, be, B, C, Hin, Hout);
  else k_patch_merge_ln<T, 12><<<blocks, 256, 0, s>>>(in, out, g, be, B, C, Hin, Hout);
}

int launch_patch_merge_ln(const void* in, void* out, const float* g, const float* be, int dt, int B, int C, int Hin, int Hout,
                          cudaStream_t s) {
  const int N = dt == DT_F32 ? 4 : 8, nvec = 4 * C / N;
  YB_REQUIRE(C % N == 0 && nvec <= 384, YB_ERR_UNSUPPORTED, "patch_merge_ln: C=%d", C);
  const dim3 blocks(std::min(((Hout + 2) * (Hout + 2) + 7) / 8, 148 * 4), B);
  YB_DISPATCH_DT(dt, (patch_merge_dispatch<T>((const T*)in, (T*)out, g, be, B, C, Hin, Hout, nvec, blocks, s)));
  YB_CHECK_LAUNCH();
  return YB_OK;
}

// ------------------------------------------------------------------------------------------------
// (Shifted-)window attention, one block per (window, head, image)
// (modules/swin_transformer.py:172-200, :246-283, mask :369-387).
//   qkv   haloed grid [B,(H+2)^2][3C] (q | k | v, each head-major), output of the qkv linear
//   pad tokens (grid padded to a multiple of 7 AFTER norm1) have qkv == the linear's bias
//   shift > 0: window coordinates live in the rolled frame; region ids give the -100 mask
//   out   haloed grid [B,(H+2)^2][C], written for real tokens only (halo/pad untouched -> the
//         projection's epilogue zeroes the halo)
// ------------------------------------------------------------------------------------------------
constexpr int WS = 7, WT = 49, HD = 32;

template <typename T>
__global__ void __launch_bounds__(64) k_window_attention(const T* __restrict__ qkv, const float* __restrict__ qkv_bias,
                                                         const float* __restrict__ table /*[169][nH]*/, T* __restrict__ out, int H,
                                                         int C, int nH, int shift) {
  __shared__ __align__(16) float s_k[WT][HD], s_v[WT][HD];
  __shared__ int s_row[WT];          // haloed row index of each window token (-1 = pad token)
  __shared__ int s_reg[WT];
  const int Hpad = (H + WS - 1) / WS * WS, nW = Hpad / WS;
  const int win = blockIdx.x, head = blockIdx.y, b = blockIdx.z;
  const int wy = win / nW, wx = win - wy * nW;
  const int tid = threadIdx.x;
  const int Hp = H + 2;
  if (tid < WT) {
    const int iy = tid / WS, ix = tid - iy * WS;
    const int ys = wy * WS + iy, xs = wx * WS + ix;                 // rolled-frame coordinates
    const int yo = (ys + shift) % Hpad, xo = (xs + shift) % Hpad;   // original (padded-grid) coordinates
    s_row[tid] = (yo < H && xo < H) ? ((b * Hp + yo + 1) * Hp + xo + 1) : -1;
    const int ry = ys < Hpad - WS ? 0 : (ys < Hpad - shift ? 1 : 2), rx = xs < Hpad - WS ? 0 : (xs < Hpad - shift ? 1 : 2);
    s_reg[tid] = ry * 3 + rx;
  }
  __syncthreads();
  const int C3 = 3 * C;
  for (int e = tid; e < WT * HD; e += 64) {
    const int t = e / HD, d = e - t * HD;
    const int row = s_row[t];
    const int ck = C + head * HD + d, cv = 2 * C + head * HD + d;
    s_k[t][d] = row >= 0 ? Tok<T>::ld(qkv + (size_t)row * C3 + ck) : qkv_bias[ck];
    s_v[t][d] = row >= 0 ? Tok<T>::ld(qkv + (size_t)row * C3 + cv) : qkv_bias[cv];
  }
  __syncthreads();
  if (tid >= WT) return;
  const int row = s_row[tid];
  if (row < 0) return;                                             // pad tokens are cropped away (:283)
  const float scale = 0.17677669529663687f;                        // 32^-0.5
  float q[HD];
#pragma unroll
  for (int d = 0; d < HD; d += TokVec<T>::N) {
    TokVec<T>::ld(qkv + (size_t)row * C3 + head * HD + d, q + d);
#pragma unroll
    for (int e = 0; e < TokVec<T>::N; ++e) q[d + e] *= scale;
  }
  const int iy = tid / WS, ix = tid - iy * WS;
  const int reg = s_reg[tid];
  float sc[WT];
  float mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < WT; ++j) {
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;                               // 4 independent chains
#pragma unroll
    for (int d = 0; d < HD; d += 4) {
      const float4 kk = *reinterpret_cast<const float4*>(&s_k[j][d]);          // broadcast read, 4 FMAs per LDS
      a0 = fmaf(q[d], kk.x, a0); a1 = fmaf(q[d + 1], kk.y, a1); a2 = fmaf(q[d + 2], kk.z, a2); a3 = fmaf(q[d + 3], kk.w, a3);
    }
    float a = (a0 + a1) + (a2 + a3);
    const int jy = j / WS, jx = j - jy * WS;
    a += __ldg(table + ((iy - jy + WS - 1) * (2 * WS - 1) + (ix - jx + WS - 1)) * nH + head);
    if (shift > 0 && s_reg[j] != reg) a += -100.f;
    sc[j] = a;
    mx = fmaxf(mx, a);
  }
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < WT; ++j) { sc[j] = expf(sc[j] - mx); sum += sc[j]; }
  const float inv = 1.f / sum;
  float o[HD];
#pragma unroll
  for (int d = 0; d < HD; ++d) o[d] = 0.f;
#pragma unroll
  for (int j = 0; j < WT; ++j) {
    const float p = sc[j] * inv;
#pragma unroll
    for (int d = 0; d < HD; d += 4) {
      const float4 vv = *reinterpret_cast<const float4*>(&s_v[j][d]);
      o[d] = fmaf(p, vv.x, o[d]); o[d + 1] = fmaf(p, vv.y, o[d + 1]); o[d + 2] = fmaf(p, vv.z, o[d + 2]); o[d + 3] = fmaf(p, vv.w, o[d + 3]);
    }
  }
  T* op = out + (size_t)row * C + head * HD;
#pragma unroll
  for (int d = 0; d < HD; d += TokVec<T>::N) TokVec<T>::st(op + d, o + d);
}

// ------------------------------------------------------------------------------------------------
// Tensor-core window attention for the 16-bit modes: one WARP per (window, head), 49 tokens padded to
// 64.  S = Q K^T and O = P V run on mma.sync m16n8k16 (fp32 accumulate) with the flash-attention
// register trick: the score accumulators of two adjacent 8-column tiles ARE the A fragment of the
// P.V product, so probabilities never leave registers.  Scale, relative-position bias, shift mask and
// softmax are applied to the fp32 accumulators.
// ------------------------------------------------------------------------------------------------
template <typename T> struct Mma16;
template <> struct Mma16<__half> {
  static __device__ __forceinline__ void mma(float* d, const uint32_t* a, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
  static __device__ __forceinline__ uint32_t pack(float x, float y) { __half2 v = __floats2half2_rn(x, y); return *reinterpret_cast<uint32_t*>(&v); }
  static __device__ __forceinline__ __half from_f(float x) { return __float2half_rn(x); }
};
template <> struct Mma16<__nv_bfloat16> {
  static __device__ __forceinline__ void mma(float* d, const uint32_t* a, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
  static __device__ __forceinline__ uint32_t pack(float x, float y) { __nv_bfloat162 v = __floats2bfloat162_rn(x, y); return *reinterpret_cast<uint32_t*>(&v); }
  static __device__ __forceinline__ __nv_bfloat16 from_f(float x) { return __float2bfloat16_rn(x); }
};

constexpr int TCA_WARPS = 4;
constexpr int QK_LD = 40;      // halves per Q/K row (32 + 8 pad: conflict-free fragment loads)
constexpr int VT_LD = 72;      // halves per V^T row (64 + 8 pad)

template <typename T>
struct __align__(16) AttnSmem {
  T q[64][QK_LD];
  T k[64][QK_LD];
  T vt[HD][VT_LD];
  float tab[176];
  int row[64];
  int reg[64];
};

template <typename T>
__global__ void __launch_bounds__(TCA_WARPS * 32) k_window_attention_tc(const T* __restrict__ qkv, const float* __restrict__ qkv_bias,
                                                                          const float* __restrict__ table, T* __restrict__ out, int B,
                                                                          int H, int C, int nH, int shift) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  AttnSmem<T>& sm = reinterpret_cast<AttnSmem<T>*>(smem_raw)[warp];
  const int Hpad = (H + WS - 1) / WS * WS, nW = Hpad / WS;
  const long long units = (long long)B * nW * nW * nH;
  const long long unit = (long long)blockIdx.x * TCA_WARPS + warp;
  if (unit >= units) return;
  const int head = (int)(unit % nH);
  const int win = (int)((unit / nH) % (nW * nW));
  const int b = (int)(unit / ((long long)nH * nW * nW));
  const int wy = win / nW, wx = win - wy * nW;
  const int Hp = H + 2, C3 = 3 * C;

  for (int t = lane; t < 64; t += 32) {
    int row = -2, reg = 0;                                          // -2: tile padding (t >= 49)
    if (t < WT) {
      const int iy = t / WS, ix = t - iy * WS;
      const int ys = wy * WS + iy, xs = wx * WS + ix;
      const int yo = (ys + shift) % Hpad, xo = (xs + shift) % Hpad;
      row = (yo < H && xo < H) ? ((b * Hp + yo + 1) * Hp + xo + 1) : -1;   // -1: grid padding token (qkv == bias)
      const int ry = ys < Hpad - WS ? 0 : (ys < Hpad - shift ? 1 : 2), rx = xs < Hpad - WS ? 0 : (xs < Hpad - shift ? 1 : 2);
      reg = ry * 3 + rx;
    }
    sm.row[t] = row; sm.reg[t] = reg;
  }
  for (int i = lane; i < 169; i += 32) sm.tab[i] = __ldg(table + i * nH + head);
  __syncwarp();
  // Q, K (row-major) and V^T into shared memory
  for (int idx = lane; idx < 64 * 4; idx += 32) {
    const int t = idx >> 2, v8 = (idx & 3) * 8;
    const int row = sm.row[t];
    T qv[8], kv[8], vv[8];
    if (row >= 0) {
      const T* base = qkv + (size_t)row * C3 + head * HD + v8;
      *reinterpret_cast<uint4*>(qv) = *reinterpret_cast<const uint4*>(base);
      *reinterpret_cast<uint4*>(kv) = *reinterpret_cast<const uint4*>(base + C);
      *reinterpret_cast<uint4*>(vv) = *reinterpret_cast<const uint4*>(base + 2 * C);
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int c = head * HD + v8 + e;
        qv[e] = Mma16<T>::from_f(row == -1 ? qkv_bias[c] : 0.f);
        kv[e] = Mma16<T>::from_f(row == -1 ? qkv_bias[C + c] : 0.f);
        vv[e] = Mma16<T>::from_f(row == -1 ? qkv_bias[2 * C + c] : 0.f);
      }
    }
    *reinterpret_cast<uint4*>(&sm.q[t][v8]) = *reinterpret_cast<uint4*>(qv);
    *reinterpret_cast<uint4*>(&sm.k[t][v8]) = *reinterpret_cast<uint4*>(kv);
#pragma unroll
    for (int e = 0; e < 8; ++e) sm.vt[v8 + e][t] = vv[e];
  }
  __syncwarp();

  const int g = lane >> 2, q4 = lane & 3;
  const float scale = 0.17677669529663687f;
  // the lane's 16 key columns: j = 8*nt + 2*q4 + e
  for (int mt = 0; mt < 4; ++mt) {
    float acc[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) { acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f; }
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      uint32_t a[4];
      a[0] = *reinterpret_cast<const uint32_t*>(&sm.q[16 * mt + g][16 * ks + 2 * q4]);
      a[1] = *reinterpret_cast<const uint32_t*>(&sm.q[16 * mt + g + 8][16 * ks + 2 * q4]);
      a[2] = *reinterpret_cast<const uint32_t*>(&sm.q[16 * mt + g][16 * ks + 8 + 2 * q4]);
      a[3] = *reinterpret_cast<const uint32_t*>(&sm.q[16 * mt + g + 8][16 * ks + 8 + 2 * q4]);
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const uint32_t b0 = *reinterpret_cast<const uint32_t*>(&sm.k[8 * nt + g][16 * ks + 2 * q4]);
        const uint32_t b1 = *reinterpret_cast<const uint32_t*>(&sm.k[8 * nt + g][16 * ks + 8 + 2 * q4]);
        Mma16<T>::mma(acc[nt], a, b0, b1);
      }
    }
    // scores -> probabilities for rows i0 = 16mt+g and i1 = i0+8
    const int i0 = 16 * mt + g, i1 = i0 + 8;
    const int iy0 = i0 / WS, ix0 = i0 - iy0 * WS, iy1 = i1 / WS, ix1 = i1 - iy1 * WS;
    const int reg0 = sm.reg[i0 & 63], reg1 = sm.reg[i1 & 63];
    float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int j = 8 * nt + 2 * q4 + e;
        float s0 = -INFINITY, s1 = -INFINITY;
        if (j < WT) {
          const int jy = j / WS, jx = j - jy * WS;
          const int regj = sm.reg[j];
          if (i0 < WT) s0 = acc[nt][e] * scale + sm.tab[(iy0 - jy + WS - 1) * (2 * WS - 1) + (ix0 - jx + WS - 1)] + ((shift > 0 && regj != reg0) ? -100.f : 0.f);
          if (i1 < WT) s1 = acc[nt][2 + e] * scale + sm.tab[(iy1 - jy + WS - 1) * (2 * WS - 1) + (ix1 - jx + WS - 1)] + ((shift > 0 && regj != reg1) ? -100.f : 0.f);
        }
        acc[nt][e] = s0; acc[nt][2 + e] = s1;
        m0 = fmaxf(m0, s0); m1 = fmaxf(m1, s1);
      }
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
    if (m0 == -INFINITY) m0 = 0.f;                                  // padded query rows
    if (m1 == -INFINITY) m1 = 0.f;
    float l0 = 0.f, l1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const float p0 = __expf(acc[nt][e] - m0), p1 = __expf(acc[nt][2 + e] - m1);
        acc[nt][e] = p0; acc[nt][2 + e] = p1;
        l0 += p0; l1 += p1;
      }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    // O = P V
    float o[4][4];
#pragma unroll
    for (int dt = 0; dt < 4; ++dt) { o[dt][0] = o[dt][1] = o[dt][2] = o[dt][3] = 0.f; }
#pragma unroll
    for (int kt = 0; kt < 4; ++kt) {
      uint32_t a[4];
      a[0] = Mma16<T>::pack(acc[2 * kt][0], acc[2 * kt][1]);
      a[1] = Mma16<T>::pack(acc[2 * kt][2], acc[2 * kt][3]);
      a[2] = Mma16<T>::pack(acc[2 * kt + 1][0], acc[2 * kt + 1][1]);
      a[3] = Mma16<T>::pack(acc[2 * kt + 1][2], acc[2 * kt + 1][3]);
#pragma unroll
      for (int dt = 0; dt < 4; ++dt) {
        const uint32_t b0 = *reinterpret_cast<const uint32_t*>(&sm.vt[8 * dt + g][16 * kt + 2 * q4]);
        const uint32_t b1 = *reinterpret_cast<const uint32_t*>(&sm.vt[8 * dt + g][16 * kt + 8 + 2 * q4]);
        Mma16<T>::mma(o[dt], a, b0, b1);
      }
    }
    const float inv0 = l0 > 0.f ? 1.f / l0 : 0.f, inv1 = l1 > 0.f ? 1.f / l1 : 0.f;
    const int r0 = i0 < WT ? sm.row[i0] : -1, r1 = i1 < WT ? sm.row[i1] : -1;
#pragma unroll
    for (int dt = 0; dt < 4; ++dt) {
      if (r0 >= 0) *reinterpret_cast<uint32_t*>(out + (size_t)r0 * C + head * HD + 8 * dt + 2 * q4) = Mma16<T>::pack(o[dt][0] * inv0, o[dt][1] * inv0);
      if (r1 >= 0) *reinterpret_cast<uint32_t*>(out + (size_t)r1 * C + head * HD + 8 * dt + 2 * q4) = Mma16<T>::pack(o[dt][2] * inv1, o[dt][3] * inv1);
    }
  }
}

int launch_window_attention(const void* qkv, const float* qkv_bias, const float* table, void* out, int dt, int B, int H, int C, int nH,
                            int shift, cudaStream_t s) {
  YB_REQUIRE(C == nH * HD, YB_ERR_UNSUPPORTED, "window_attention: head_dim must be 32 (C=%d heads=%d)", C, nH);
  const int nW = (H + WS - 1) / WS;
  if (dt != DT_F32 && !getenv("YOLACT_B200_NO_TC_ATTN")) {
    const long long units = (long long)B * nW * nW * nH;
    const int blocks = (int)((units + TCA_WARPS - 1) / TCA_WARPS);
    if (dt == DT_F16) {
      const size_t smem = sizeof(AttnSmem<__half>) * TCA_WARPS;
      YB_CHECK_CUDA(cudaFuncSetAttribute(k_window_attention_tc<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k_window_attention_tc<__half><<<blocks, TCA_WARPS * 32, smem, s>>>((const __half*)qkv, qkv_bias, table, (__half*)out, B, H, C, nH, shift);
    } else {
      const size_t smem = sizeof(AttnSmem<__nv_bfloat16>) * TCA_WARPS;
      YB_CHECK_CUDA(cudaFuncSetAttribute(k_window_attention_tc<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k_window_attention_tc<__nv_bfloat16><<<blocks, TCA_WARPS * 32, smem, s>>>((const __nv_bfloat16*)qkv, qkv_bias, table, (__nv_bfloat16*)out, B, H, C, nH, shift);
    }
    YB_CHECK_LAUNCH();
    return YB_OK;
  }
  dim3 grid(nW * nW, nH, B);
  YB_DISPATCH_DT(dt, (k_window_attention<T><<<grid, 64, 0, s>>>((const T*)qkv, qkv_bias, table, (T*)out, H, C, nH, shift)));
  YB_CHECK_LAUNCH();
  return YB_OK;
}

}  // namespace yb
