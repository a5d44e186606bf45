// Swin-T backbone kernels (modules/swin_transformer.py) on the haloed NHWC token grid: patch embedding
// (+LayerNorm), LayerNorm, shifted-window attention, patch merging (+LayerNorm).  The linear layers
// (qkv, proj, fc1+GELU, fc2, reduction) run as 1x1 convolutions on the tcgen05 / CUDA-core conv
// kernels, with the residual adds in their epilogues.
#include "layers.cuh"
#include <math.h>
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
  const long long total = (long long)B * Hp * Hp;
  for (long long tok = (long long)blockIdx.x * 8 + (threadIdx.x >> 5); tok < total; tok += (long long)gridDim.x * 8) {
    const int xp = (int)(tok % Hp), yp = (int)((tok / Hp) % Hp), b = (int)(tok / ((long long)Hp * Hp));
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
  const long long total = (long long)B * (Hg + 2) * (Hg + 2);
  const int blocks = (int)std::min<long long>((total + 7) / 8, 148LL * 16);
  YB_DISPATCH_DT(dt, (k_patch_embed<T><<<blocks, 256, 0, s>>>(img, w, bias, g, be, (T*)out, B, S, Hg)));
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
  const long long total = (long long)B * Hp * Hp;
  const long long stride = (long long)gridDim.x * 8 * TPW;
  for (long long tok0 = ((long long)blockIdx.x * 8 + (threadIdx.x >> 5)) * TPW; tok0 < total; tok0 += stride) {
    const long long tok = tok0 + lane / G;
    const bool live = tok < total;
    const int xp = (int)(tok % Hp), yp = (int)((tok / Hp) % Hp);
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
static void layernorm_dispatch(const T* in, T* out, const float* g, const float* be, int B, int C, int H, int nvec, int blocks, cudaStream_t s) {
  if (nvec <= 16) k_layernorm<T, 1, 16><<<blocks, 256, 0, s>>>(in, out, g, be, B, C, H);
  else if (nvec <= 32) k_layernorm<T, 1, 32><<<blocks, 256, 0, s>>>(in, out, g, be, B, C, H);
  else if (nvec <= 64) k_layernorm<T, 2, 32><<<blocks, 256, 0, s>>>(in, out, g, be, B, C, H);
  else if (nvec <= 96) k_layernorm<T, 3, 32><<<blocks, 256, 0, s>>>(in, out, g, be, B, C, H);
  else k_layernorm<T, 6, 32><<<blocks, 256, 0, s>>>(in, out, g, be, B, C, H);
}

int launch_layernorm(const void* in, void* out, const float* g, const float* be, int dt, int B, int C, int H, cudaStream_t s) {
  const int nvec = C / (dt == DT_F32 ? 4 : 8);
  YB_REQUIRE(C % (dt == DT_F32 ? 4 : 8) == 0 && nvec <= 192, YB_ERR_UNSUPPORTED, "layernorm: C=%d", C);
  const long long total = (long long)B * (H + 2) * (H + 2);
  const int blocks = (int)std::min<long long>((total + 7) / 8, 148LL * 16);
  YB_DISPATCH_DT(dt, (layernorm_dispatch<T>((const T*)in, (T*)out, g, be, B, C, H, nvec, blocks, s)));
  YB_CHECK_LAUNCH();
  return YB_OK;
}

// ------------------------------------------------------------------------------------------------
// PatchMerging gather + LayerNorm(4C) (modules/swin_transformer.py:299-325): output token (y,x)
// concatenates input tokens (2y,2x), (2y+1,2x), (2y,2x+1), (2y+1,2x+1) (zero beyond an odd edge).
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) k_patch_merge_ln(const T* __restrict__ in, T* __restrict__ out, const float* __restrict__ g,
                                                        const float* __restrict__ be, int B, int C, int Hin, int Hout) {
  const int lane = threadIdx.x & 31;
  const int Hpi = Hin + 2, Hpo = Hout + 2, C4 = 4 * C;
  const long long total = (long long)B * Hpo * Hpo;
  for (long long tok = (long long)blockIdx.x * 8 + (threadIdx.x >> 5); tok < total; tok += (long long)gridDim.x * 8) {
    const int xp = (int)(tok % Hpo), yp = (int)((tok / Hpo) % Hpo), b = (int)(tok / ((long long)Hpo * Hpo));
    T* o = out + tok * C4;
    if (yp == 0 || yp == Hout + 1 || xp == 0 || xp == Hout + 1) {
      for (int c = lane; c < C4; c += 32) tok_st<T>(o + c, 0.f);
      continue;
    }
    const int y = yp - 1, x = xp - 1;
    auto src = [&](int c4) -> float {
      const int part = c4 / C, c = c4 - part * C;
      const int iy = 2 * y + (part & 1), ix = 2 * x + (part >> 1);
      if (iy >= Hin || ix >= Hin) return 0.f;
      return Tok<T>::ld(in + (((size_t)b * Hpi + iy + 1) * Hpi + ix + 1) * C + c);
    };
    float s = 0.f;
    for (int c = lane; c < C4; c += 32) s += src(c);
    const float mean = warp_sum(s) / (float)C4;
    float v = 0.f;
    for (int c = lane; c < C4; c += 32) { const float d = src(c) - mean; v += d * d; }
    const float rstd = rsqrtf(warp_sum(v) / (float)C4 + 1e-5f);
    for (int c = lane; c < C4; c += 32) tok_st<T>(o + c, (src(c) - mean) * rstd * g[c] + be[c]);
  }
}

int launch_patch_merge_ln(const void* in, void* out, const float* g, const float* be, int dt, int B, int C, int Hin, int Hout,
                          cudaStream_t s) {
  const long long total = (long long)B * (Hout + 2) * (Hout + 2);
  const int blocks = (int)std::min<long long>((total + 7) / 8, 148LL * 16);
  YB_DISPATCH_DT(dt, (k_patch_merge_ln<T><<<blocks, 256, 0, s>>>((const T*)in, (T*)out, g, be, B, C, Hin, Hout)));
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

int launch_window_attention(const void* qkv, const float* qkv_bias, const float* table, void* out, int dt, int B, int H, int C, int nH,
                            int shift, cudaStream_t s) {
  YB_REQUIRE(C == nH * HD, YB_ERR_UNSUPPORTED, "window_attention: head_dim must be 32 (C=%d heads=%d)", C, nH);
  const int nW = (H + WS - 1) / WS;
  dim3 grid(nW * nW, nH, B);
  YB_DISPATCH_DT(dt, (k_window_attention<T><<<grid, 64, 0, s>>>((const T*)qkv, qkv_bias, table, (T*)out, H, C, nH, shift)));
  YB_CHECK_LAUNCH();
  return YB_OK;
}

}  // namespace yb
