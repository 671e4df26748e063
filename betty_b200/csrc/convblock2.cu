// Fused second-order rules of an INNER convolution block of a bf16-autocast graph (64 -> 64 channels)
//
//     x_in --conv3x3(W,b)--> y --BatchNorm2d(batch stats; gamma,beta)--> z --[ReLU]--> MaxPool2d(2)--> q
//
// (blocks 2-4 of reference examples/implicit_maml/models.py:9-24).  Every activation-side tensor of the block lives
// in ONE layout: bf16, channels-last, spatially padded -- [N][H+2][W+2][64] with a zero border ("padded NHWC").  It is
// what the halo-resident tensor-core kernels (conv_halo.cu) load by TMA and what their epilogues write, so between the
// products there are only streaming kernels whose lanes run along the channels (128-byte rows) -- no layout
// transposes, no fp32 round trips, no pack kernels.  Tangents / adjoints also enter and leave the block in that
// layout when the neighbouring block is fused too (plan.py `tfmt`), else as the plan's fp32 NCHW buffers.
//
// What makes the block cheap (besides the layout):
//   * a_q and at_q are SPARSE on the conv-output grid (one pixel per window and channel), so the BatchNorm adjoint
//     statistics m1, m2, mt1, mt2 and the gamma / beta / bias slices of H.d are sums over POOLED arrays;
//   * the conv-output-sized adjoint tangent  at_y = sparse + d0 + xhat d1 + t_y d2  (norm.cu's rule expanded, per-channel
//     d*) is produced once, directly as the TMA operand of the input-gradient and weight-gradient kernels;
//   * t_q needs t_y only at each window's arg-max pixel.
//
//   TF   t_y = conv(t_in, W) + conv(x_in, t_W) + t_b          conv_halo_kernel -> bf16 padded NHWC
//        mean_t, sdot (sums over t_y, xhat t_y)                cb2_stats_kernel
//        t_q = mask (gamma rstd (t_y* - mean_t - xhat* sdot) + t_gamma xhat* + t_beta), dxhat*      cb2_final_kernel
//   TB   pooled sums S_at, S_atxh, S_adxh                      cb2_reduce_kernel
//        d0, d1, d2, cw, cd; at_gamma, at_beta, at_b           cb2_coef_kernel
//        at_y                                                  cb2_dense_kernel
//        at_in = dgrad(at_y, W) + dgrad(a_y, t_W)              conv_halo_kernel (flipped taps)
//        at_W += wgrad(at_y, x_in) + wgrad(a_y, t_in)          wgrad_halo_kernel
#include <cuda_bf16.h>
#include <stdlib.h>

#include "../../include/betty_b200.h"
#include "bb_common.cuh"
#include "conv_halo.h"
#include "gemm_tma.h"
#include "plan.h"
#include "tma.h"

namespace {

typedef __nv_bfloat16 bf16;
typedef __nv_bfloat162 bf162;

struct G2 {
  int N, H, W, HP, WP, relu;      // conv output grid = input grid (3x3, padding 1); C = O = 64
};

struct Ws2 {
  double* dsum;        // [64] x: 0 sum y | 1 sum y^2 | 2 sum t_y | 3 sum xhat t_y | 4 S_at | 5 S_atxh | 6 S_adxh | 7 Sa | 8 Saxh | 9 sum xhat
  float* mean;         // [64]
  float* rstd;         // [64]
  float* coef;         // [8][64]: d0, d1, d2, cw, cd, mean_t, sdot
  unsigned char* sel;  // [N][HP][WP][64] pooled, channels last: arg-max code | mask << 2
  bf16* xh;            // pooled: xhat at the arg-max pixel
  bf16* dxh;           // pooled: dxhat at the arg-max pixel (per iteration)
  bf16* aqm;           // pooled: mask * a_q
  bf16* yb;            // padded NHWC copy of the base conv output y (per call)
  bf16* ty;            // padded NHWC: conv tangent (per iteration)
  bf16* aty;           // padded NHWC: adjoint tangent at the conv output (per iteration)
  bf16* ay;            // padded NHWC: base adjoint at the conv output (per call)
  bf16* xin;           // padded NHWC: base input (per call)
  bf16* tin;           // padded NHWC: input tangent, packed here when the producer wrote fp32 NCHW
  bf16* wf;            // [64][9][64] forward operand of W        (per call)
  bf16* wd;            // [64][9][64] input-gradient operand of W (per call)
  bf16* twf;           // same for t_W (per iteration)
  bf16* twd;
  size_t bytes;
};

inline size_t up(size_t x) { return (x + 1023) & ~(size_t)1023; }

Ws2 layout(void* base, const G2& g) {
  Ws2 w{};
  size_t at = 0;
  uint8_t* b = reinterpret_cast<uint8_t*>(base);
  auto take = [&](size_t bytes) { size_t o = at; at = up(at + bytes); return b ? b + o : nullptr; };
  const size_t pooled = (size_t)g.N * g.HP * g.WP * 64, fullp = (size_t)g.N * (g.H + 2) * (g.W + 2) * 64;
  w.dsum = reinterpret_cast<double*>(take(8 * 10 * 64));
  w.mean = reinterpret_cast<float*>(take(4 * 64));
  w.rstd = reinterpret_cast<float*>(take(4 * 64));
  w.coef = reinterpret_cast<float*>(take(4 * 8 * 64));
  w.sel = reinterpret_cast<unsigned char*>(take(pooled));
  w.xh = reinterpret_cast<bf16*>(take(2 * pooled));
  w.dxh = reinterpret_cast<bf16*>(take(2 * pooled));
  w.aqm = reinterpret_cast<bf16*>(take(2 * pooled));
  w.yb = reinterpret_cast<bf16*>(take(2 * fullp));
  w.ty = reinterpret_cast<bf16*>(take(2 * fullp));
  w.aty = reinterpret_cast<bf16*>(take(2 * fullp));
  w.ay = reinterpret_cast<bf16*>(take(2 * fullp));
  w.xin = reinterpret_cast<bf16*>(take(2 * fullp));
  w.tin = reinterpret_cast<bf16*>(take(2 * fullp));
  const size_t wb = (size_t)2 * 64 * 9 * 64;
  w.wf = reinterpret_cast<bf16*>(take(wb));
  w.wd = reinterpret_cast<bf16*>(take(wb));
  w.twf = reinterpret_cast<bf16*>(take(wb));
  w.twd = reinterpret_cast<bf16*>(take(wb));
  w.bytes = at;
  return w;
}

struct A2 {
  G2 g;
  Ws2 w;
  const void* y; int dty;          // base conv output, NCHW (statistics / packing, once per call)
  const void* q; int dtq;          // base pooled output, NCHW (ReLU mask)
  const int64_t* idx;              // arg-max indices, NCHW
  const float* gamma;
  float eps;
  const float *t_b, *t_gamma, *t_beta;
  float *at_b, *at_gamma, *at_beta;
  // pooled-side buffers of q: fp32 NCHW (plan standard) or bf16 padded NHWC [N][HP+2][WP+2][64] (fused neighbour)
  float* tq_f32; bf16* tq_nhwc;
  const float* aq_f32; const bf16* aq_nhwc;
  const float* atq_f32; const bf16* atq_nhwc;
  int base;                        // coef / dense kernels: 1 = base adjoint a_y (per call), 0 = adjoint tangent at_y
};

__device__ __forceinline__ float2 ldbf2(const bf16* p) { return __bfloat1622float2(*reinterpret_cast<const bf162*>(p)); }
__device__ __forceinline__ void stbf2(bf16* p, float a, float b) { *reinterpret_cast<bf162*>(p) = __floats2bfloat162_rn(a, b); }

// ---- once per call -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) cb2_ystats_kernel(const A2 A) {
  __shared__ double red[32];
  const int o = blockIdx.x, HW = A.g.H * A.g.W;
  double s0 = 0, s1 = 0;
  for (int n = blockIdx.y; n < A.g.N; n += gridDim.y) {
    const int64_t base = ((int64_t)n * 64 + o) * HW;
    for (int i = threadIdx.x; i < HW; i += blockDim.x) {
      const float v = bb::ldf(A.y, base + i, A.dty);
      s0 += v;
      s1 += (double)v * v;
    }
  }
  s0 = bb::block_sum<double>(s0, red);
  s1 = bb::block_sum<double>(s1, red);
  if (threadIdx.x == 0) {
    atomicAdd(&A.w.dsum[o], s0);
    atomicAdd(&A.w.dsum[64 + o], s1);
  }
}

__global__ void cb2_ystats_finish_kernel(const A2 A) {
  const int o = threadIdx.x;
  if (o >= 64) return;
  const double cnt = (double)A.g.N * A.g.H * A.g.W;
  const double mean = A.w.dsum[o] / cnt;
  const double var = A.w.dsum[64 + o] / cnt - mean * mean;
  const float mean_f = (float)mean, rstd_f = (float)rsqrt((var > 0 ? var : 0) + (double)A.eps);
  A.w.mean[o] = mean_f;
  A.w.rstd[o] = rstd_f;
  // sum of xhat over y with the ROUNDED mean (what every kernel uses): (sum y - cnt mean_f) rstd_f, no second pass over y
  A.w.dsum[9 * 64 + o] = (A.w.dsum[o] - cnt * (double)mean_f) * (double)rstd_f;
}

// pooled arrays (channels last) from the NCHW arg-max indices; per-channel Sa, Saxh; per-channel sum of xhat over y
// One block per image, one pooled row per step: phase 1 walks the NCHW tensors (runs of WP contiguous elements per
// channel) into a [WP][64] shared tile, phase 2 writes the channels-last rows in full lines and accumulates the channel
// sums (thread t always owns channel t & 63).
__global__ void __launch_bounds__(256) cb2_prep_kernel(const A2 A) {
  extern __shared__ float prep_sm[];
  __shared__ double red[2][4][64];
  const G2& g = A.g;
  const int n = blockIdx.x, PW = g.HP * g.WP, HW = g.H * g.W, cnt = g.WP * 64;
  float* xh_t = prep_sm;
  float* aq_t = prep_sm + cnt;
  unsigned char* sel_t = reinterpret_cast<unsigned char*>(prep_sm + 2 * cnt);
  const int o2 = threadIdx.x & 63;
  double sa = 0, saxh = 0;
  for (int hp = 0; hp < g.HP; ++hp) {
    __syncthreads();
    for (int e = threadIdx.x; e < cnt; e += blockDim.x) {
      const int o = e / g.WP, wp = e - o * g.WP;
      const int64_t pi = ((int64_t)n * 64 + o) * PW + hp * g.WP + wp;
      const int64_t id = A.idx[pi];
      const int iy = (int)(id / g.W), ix = (int)(id - (int64_t)iy * g.W);
      const bool m = g.relu ? (bb::ldf(A.q, pi, A.dtq) > 0.f) : true;
      const int t = wp * 64 + o;
      sel_t[t] = (unsigned char)(((iy - 2 * hp) & 1) * 2 + ((ix - 2 * wp) & 1) + (m ? 4 : 0));
      xh_t[t] = (bb::ldf(A.y, ((int64_t)n * 64 + o) * HW + id, A.dty) - A.w.mean[o]) * A.w.rstd[o];
      if (!A.aq_nhwc) aq_t[t] = A.aq_f32[pi];
    }
    __syncthreads();
    const int64_t p0 = ((int64_t)n * g.HP + hp) * g.WP * 64;
    for (int t = threadIdx.x; t < cnt; t += blockDim.x) {
      const unsigned char code = sel_t[t];
      const int wp = t >> 6;
      float aq = A.aq_nhwc ? __bfloat162float(A.aq_nhwc[((((int64_t)n * (g.HP + 2) + hp + 1) * (g.WP + 2)) + wp + 1) * 64 + o2])
                           : aq_t[t];
      aq = (code & 4) ? aq : 0.f;
      const bf16 xb = __float2bfloat16(xh_t[t]), ab = __float2bfloat16(aq);
      A.w.sel[p0 + t] = code;
      A.w.xh[p0 + t] = xb;
      A.w.aqm[p0 + t] = ab;
      // the sums use the values as the K-loop kernels will read them back (bf16)
      const float aqr = __bfloat162float(ab), xhr = __bfloat162float(xb);
      sa += aqr;
      saxh += (double)aqr * xhr;
    }
  }
  red[0][threadIdx.x >> 6][o2] = sa;
  red[1][threadIdx.x >> 6][o2] = saxh;
  __syncthreads();
  if (threadIdx.x < 128) {
    const int which = threadIdx.x >> 6;
    const double v = red[which][0][o2] + red[which][1][o2] + red[which][2][o2] + red[which][3][o2];
    atomicAdd(&A.w.dsum[(7 + which) * 64 + o2], v);
  }
}

// NCHW (any base dtype) -> bf16 padded NHWC [N][H+2][W+2][64].  grid (N, H)
__global__ void __launch_bounds__(256) cb2_pack_padded_kernel(const void* src, int dt, int C, int H, int W, bf16* dst) {
  extern __shared__ float sm[];                 // [W][C + 1]
  const int n = blockIdx.x, y = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int c = warp; c < C; c += 8) {
    const int64_t row = (((int64_t)n * C + c) * H + y) * W;
    for (int x = lane; x < W; x += 32) sm[x * (C + 1) + c] = bb::ldf(src, row + x, dt);
  }
  __syncthreads();
  bf16* d = dst + ((((int64_t)n * (H + 2) + y + 1) * (W + 2)) + 1) * 64;
  for (int i = threadIdx.x; i < W * 32; i += blockDim.x) {
    const int x = i >> 5, c = (i & 31) * 2;
    const float v0 = c < C ? sm[x * (C + 1) + c] : 0.f, v1 = c + 1 < C ? sm[x * (C + 1) + c + 1] : 0.f;
    reinterpret_cast<bf162*>(d)[i] = __floats2bfloat162_rn(v0, v1);
  }
}

// ---- streaming kernels: a lane owns 8 channels (one 16-byte bf16 chunk), four lanes-groups of a warp take four rows ----
struct F8 {
  float v[8];
};
__device__ __forceinline__ F8 ld8(const bf16* p) {
  const uint4 r = *reinterpret_cast<const uint4*>(p);
  F8 o;
  o.v[0] = __uint_as_float(r.x << 16); o.v[1] = __uint_as_float(r.x & 0xffff0000u);
  o.v[2] = __uint_as_float(r.y << 16); o.v[3] = __uint_as_float(r.y & 0xffff0000u);
  o.v[4] = __uint_as_float(r.z << 16); o.v[5] = __uint_as_float(r.z & 0xffff0000u);
  o.v[6] = __uint_as_float(r.w << 16); o.v[7] = __uint_as_float(r.w & 0xffff0000u);
  return o;
}
__device__ __forceinline__ void st8(bf16* p, const F8& f) {
  bf162 a = __floats2bfloat162_rn(f.v[0], f.v[1]), b = __floats2bfloat162_rn(f.v[2], f.v[3]);
  bf162 c = __floats2bfloat162_rn(f.v[4], f.v[5]), d = __floats2bfloat162_rn(f.v[6], f.v[7]);
  uint4 o;
  o.x = *reinterpret_cast<uint32_t*>(&a); o.y = *reinterpret_cast<uint32_t*>(&b);
  o.z = *reinterpret_cast<uint32_t*>(&c); o.w = *reinterpret_cast<uint32_t*>(&d);
  *reinterpret_cast<uint4*>(p) = o;
}
__device__ __forceinline__ F8 ld8f(const float* p) {       // 8 fp32 parameters / coefficients
  const float4 a = bb::ld4(p), b = bb::ld4(p + 4);
  F8 o;
  o.v[0] = a.x; o.v[1] = a.y; o.v[2] = a.z; o.v[3] = a.w; o.v[4] = b.x; o.v[5] = b.y; o.v[6] = b.z; o.v[7] = b.w;
  return o;
}
// 8 arg-max codes of a window (one byte per channel)
__device__ __forceinline__ void ld_codes(const unsigned char* p, unsigned (&c)[8]) {
  const uint2 r = *reinterpret_cast<const uint2*>(p);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    c[e] = (r.x >> (8 * e)) & 0xffu;
    c[4 + e] = (r.y >> (8 * e)) & 0xffu;
  }
}
// pooled adjoint tangent of a window: bf16 padded NHWC, or the plan's fp32 NCHW buffer (last fused block)
__device__ __forceinline__ F8 load_atq8(const A2& A, int n, int hp, int wp, int c0) {
  const G2& g = A.g;
  if (A.atq_nhwc) return ld8(A.atq_nhwc + ((((int64_t)n * (g.HP + 2) + hp + 1) * (g.WP + 2)) + wp + 1) * 64 + c0);
  F8 o;
  const int64_t b = (((int64_t)n * 64 + c0) * g.HP + hp) * g.WP + wp, cs = (int64_t)g.HP * g.WP;
#pragma unroll
  for (int e = 0; e < 8; ++e) o.v[e] = A.atq_f32[b + e * cs];
  return o;
}

// ---- tangent forward -----------------------------------------------------------------------------------------------
// sum t_y and sum xhat t_y per channel: one pass over the padded rows of t_y and y (border rows of t_y are zeros)
__global__ void __launch_bounds__(256) cb2_stats_kernel(const A2 A) {
  __shared__ float red[8][2][64];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane >> 3, c0 = (lane & 7) * 8;
  const int64_t rows = (int64_t)A.g.N * (A.g.H + 2) * (A.g.W + 2);
  const F8 mean = ld8f(A.w.mean + c0), rstd = ld8f(A.w.rstd + c0);
  float s0[8], s1[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) s0[e] = s1[e] = 0.f;
  const int64_t stride = (int64_t)gridDim.x * 32;
  for (int64_t r = (int64_t)blockIdx.x * 32 + warp * 4 + sub; r < rows; r += stride) {
    const F8 t = ld8(A.w.ty + r * 64 + c0), yv = ld8(A.w.yb + r * 64 + c0);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      s0[e] += t.v[e];
      s1[e] = fmaf((yv.v[e] - mean.v[e]) * rstd.v[e], t.v[e], s1[e]);
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    s0[e] += __shfl_xor_sync(0xffffffffu, s0[e], 8); s0[e] += __shfl_xor_sync(0xffffffffu, s0[e], 16);
    s1[e] += __shfl_xor_sync(0xffffffffu, s1[e], 8); s1[e] += __shfl_xor_sync(0xffffffffu, s1[e], 16);
  }
  if (sub == 0) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      red[warp][0][c0 + e] = s0[e];
      red[warp][1][c0 + e] = s1[e];
    }
  }
  __syncthreads();
  if (threadIdx.x < 128) {
    const int k = threadIdx.x >> 6, o = threadIdx.x & 63;
    float acc = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) acc += red[w][k][o];
    atomicAdd(&A.w.dsum[(2 + k) * 64 + o], (double)acc);
  }
}

// pooled finalize: t_y at the arg-max pixel, dxhat*, t_q
__global__ void __launch_bounds__(256) cb2_final_kernel(const A2 A) {
  const G2& g = A.g;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane >> 3, c0 = (lane & 7) * 8;
  const double P = (double)g.N * g.H * g.W;
  F8 rstd = ld8f(A.w.rstd + c0), mean_t, sdot, gam, tgam, tbet;
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int o = c0 + e;
    const double mt = A.w.dsum[2 * 64 + o] / P;
    mean_t.v[e] = (float)mt;
    sdot.v[e] = (float)((A.w.dsum[3 * 64 + o] - mt * A.w.dsum[9 * 64 + o]) / P);
    gam.v[e] = A.gamma ? A.gamma[o] : 1.f;
    tgam.v[e] = A.t_gamma ? A.t_gamma[o] : 0.f;
    tbet.v[e] = A.t_beta ? A.t_beta[o] : 0.f;
    if (blockIdx.x == 0 && warp == 0 && sub == 0) {
      A.w.coef[5 * 64 + o] = mean_t.v[e];
      A.w.coef[6 * 64 + o] = sdot.v[e];
    }
  }
  const int64_t nwin = (int64_t)g.N * g.HP * g.WP;
  const int Wp = g.W + 2;
  for (int64_t w = (int64_t)blockIdx.x * 32 + warp * 4 + sub; w < nwin; w += (int64_t)gridDim.x * 32) {
    const int wp = (int)(w % g.WP);
    const int64_t t = w / g.WP;
    const int hp = (int)(t % g.HP), n = (int)(t / g.HP);
    const int64_t pi = w * 64 + c0;
    unsigned code[8];
    ld_codes(A.w.sel + pi, code);
    const F8 xh = ld8(A.w.xh + pi);
    // the four candidate pixels of the window (padded rows r0, r0+1, r0+Wp, r0+Wp+1)
    const bf16* t0 = A.w.ty + (((int64_t)n * (g.H + 2) + 2 * hp + 1) * Wp + 2 * wp + 1) * 64 + c0;
    const F8 ta = ld8(t0), tb = ld8(t0 + 64), tc = ld8(t0 + (int64_t)Wp * 64), td = ld8(t0 + (int64_t)(Wp + 1) * 64);
    F8 dx, tq;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const unsigned cd = code[e] & 3u;
      const float ty = cd == 0 ? ta.v[e] : cd == 1 ? tb.v[e] : cd == 2 ? tc.v[e] : td.v[e];
      dx.v[e] = (ty - mean_t.v[e] - xh.v[e] * sdot.v[e]) * rstd.v[e];
      tq.v[e] = (code[e] & 4u) ? fmaf(gam.v[e], dx.v[e], fmaf(tgam.v[e], xh.v[e], tbet.v[e])) : 0.f;
    }
    st8(A.w.dxh + pi, dx);
    if (A.tq_nhwc) {
      st8(A.tq_nhwc + ((((int64_t)n * (g.HP + 2) + hp + 1) * (g.WP + 2)) + wp + 1) * 64 + c0, tq);
    } else {
      const int64_t b = (((int64_t)n * 64 + c0) * g.HP + hp) * g.WP + wp, cs = (int64_t)g.HP * g.WP;
#pragma unroll
      for (int e = 0; e < 8; ++e) A.tq_f32[b + e * cs] = tq.v[e];
    }
  }
}

// ---- tangent backward ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) cb2_reduce_kernel(const A2 A) {
  __shared__ float red[8][3][64];
  const G2& g = A.g;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane >> 3, c0 = (lane & 7) * 8;
  float s0[8], s1[8], s2[8];       // S_at, S_atxh, S_adxh
#pragma unroll
  for (int e = 0; e < 8; ++e) s0[e] = s1[e] = s2[e] = 0.f;
  const int64_t nwin = (int64_t)g.N * g.HP * g.WP;
  for (int64_t w = (int64_t)blockIdx.x * 32 + warp * 4 + sub; w < nwin; w += (int64_t)gridDim.x * 32) {
    const int wp = (int)(w % g.WP);
    const int64_t t = w / g.WP;
    const int hp = (int)(t % g.HP), n = (int)(t / g.HP);
    const int64_t pi = w * 64 + c0;
    unsigned code[8];
    ld_codes(A.w.sel + pi, code);
    const F8 xh = ld8(A.w.xh + pi), aq = ld8(A.w.aqm + pi), dx = ld8(A.w.dxh + pi);
    const F8 at = load_atq8(A, n, hp, wp, c0);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float v = (code[e] & 4u) ? at.v[e] : 0.f;
      s0[e] += v;
      s1[e] = fmaf(v, xh.v[e], s1[e]);
      s2[e] = fmaf(aq.v[e], dx.v[e], s2[e]);
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    s0[e] += __shfl_xor_sync(0xffffffffu, s0[e], 8); s0[e] += __shfl_xor_sync(0xffffffffu, s0[e], 16);
    s1[e] += __shfl_xor_sync(0xffffffffu, s1[e], 8); s1[e] += __shfl_xor_sync(0xffffffffu, s1[e], 16);
    s2[e] += __shfl_xor_sync(0xffffffffu, s2[e], 8); s2[e] += __shfl_xor_sync(0xffffffffu, s2[e], 16);
  }
  if (sub == 0) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      red[warp][0][c0 + e] = s0[e];
      red[warp][1][c0 + e] = s1[e];
      red[warp][2][c0 + e] = s2[e];
    }
  }
  __syncthreads();
  if (threadIdx.x < 192) {
    const int k = threadIdx.x >> 6, o = threadIdx.x & 63;
    float acc = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) acc += red[w][k][o];
    atomicAdd(&A.w.dsum[(4 + k) * 64 + o], (double)acc);
  }
}

// per-channel coefficients of the dense rule; gamma / beta / conv-bias slices of H.d.  base = 1: the base adjoint a_y.
__global__ void cb2_coef_kernel(const A2 A) {
  const G2& g = A.g;
  const int o = threadIdx.x;
  if (o >= 64) return;
  const double P = (double)g.N * g.H * g.W;
  const float rstd = A.w.rstd[o];
  const float gam = A.gamma ? A.gamma[o] : 1.f;
  const double Sa = A.w.dsum[7 * 64 + o], Saxh = A.w.dsum[8 * 64 + o], sx = A.w.dsum[9 * 64 + o];
  const float m1 = (float)(gam * Sa / P), m2 = (float)(gam * Saxh / P);
  float* c = A.w.coef;
  if (A.base) {
    // a_y = rstd (gamma a_z - m1 - xhat m2)
    c[0 * 64 + o] = -rstd * m1;
    c[1 * 64 + o] = -rstd * m2;
    c[2 * 64 + o] = 0.f;
    c[3 * 64 + o] = 0.f;          // coefficient of at_q
    c[4 * 64 + o] = rstd * gam;   // coefficient of mask a_q
    return;
  }
  const float tgam = A.t_gamma ? A.t_gamma[o] : 0.f;
  const float mean_t = c[5 * 64 + o], sdot = c[6 * 64 + o];
  const double S_at = A.w.dsum[4 * 64 + o], S_atxh = A.w.dsum[5 * 64 + o], S_adxh = A.w.dsum[6 * 64 + o];
  const float mt1 = (float)((gam * S_at + tgam * Sa) / P);
  const float mt2 = (float)((gam * S_atxh + tgam * Saxh + gam * S_adxh) / P);
  const float d2 = -rstd * rstd * m2;
  const float d1 = -rstd * mt2 + 2.f * rstd * rstd * sdot * m2;
  const float d0 = -rstd * mt1 + rstd * rstd * sdot * m1 + rstd * rstd * m2 * mean_t;
  const float cw = rstd * gam, cd = rstd * tgam - rstd * rstd * sdot * gam;
  c[0 * 64 + o] = d0; c[1 * 64 + o] = d1; c[2 * 64 + o] = d2; c[3 * 64 + o] = cw; c[4 * 64 + o] = cd;
  if (A.at_gamma) A.at_gamma[o] += (float)(S_atxh + S_adxh);
  if (A.at_beta) A.at_beta[o] += (float)S_at;
  // sum_p at_y in closed form: sparse part + d0 P + d1 sum(xhat) + d2 sum(t_y)
  if (A.at_b) A.at_b[o] += (float)((double)cw * S_at + (double)cd * Sa + (double)d0 * P + (double)d1 * sx + (double)d2 * P * mean_t);
}

// dense rule on the conv-output grid: at_y (or the base a_y) as bf16 padded NHWC.  grid (N, ceil(H/2)): the two image
// rows of one pooled row per block; the pooled window of pixel (y, x) is (y >> 1, x >> 1)
// A warp takes DU consecutive 2x2 windows per step: lane group `sub` owns one pixel of the window (16 bytes = 8 channels
// per lane), the window's codes / pooled adjoints are one broadcast read.  All loads of the DU windows are issued before
// the first use (the row-per-block version with one load batch per iteration sat at 1.9 TB/s).  Windows cover
// ceil(H/2) x ceil(W/2), so the odd last row / column (no pooled contribution) is written too.
constexpr int DU = 4;
__device__ __forceinline__ void unpack8(const uint4& r, float (&o)[8]) {
  o[0] = __uint_as_float(r.x << 16); o[1] = __uint_as_float(r.x & 0xffff0000u);
  o[2] = __uint_as_float(r.y << 16); o[3] = __uint_as_float(r.y & 0xffff0000u);
  o[4] = __uint_as_float(r.z << 16); o[5] = __uint_as_float(r.z & 0xffff0000u);
  o[6] = __uint_as_float(r.w << 16); o[7] = __uint_as_float(r.w & 0xffff0000u);
}
template <bool BASE, bool ATQF32>
__global__ void __launch_bounds__(256, 2) cb2_dense_kernel(const A2 A) {
  const G2& g = A.g;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane >> 3, c0 = (lane & 7) * 8;
  const int dy = sub >> 1, dx = sub & 1;
  // r = k0 + k1 xhat + k2 t  with xhat = (y - mean) rstd  ->  e0 + e1 y + k2 t; the five coefficient rows live in shared
  // memory (40 registers less per thread: two CTAs per SM)
  __shared__ __align__(16) float cf[5][64];
  if (threadIdx.x < 64) {
    const int o = threadIdx.x;
    const float* c = A.w.coef;
    const float e1 = c[64 + o] * A.w.rstd[o];
    cf[0][o] = c[o] - e1 * A.w.mean[o];
    cf[1][o] = e1;
    cf[2][o] = c[128 + o];
    cf[3][o] = c[192 + o];
    cf[4][o] = c[256 + o];
  }
  __syncthreads();
  const int HC = (g.H + 1) >> 1, WC = (g.W + 1) >> 1;
  const int64_t total = (int64_t)g.N * HC * WC;
  const bf16* __restrict__ yb = A.w.yb;
  const bf16* __restrict__ ty = A.w.ty;
  const bf16* __restrict__ aqm = A.w.aqm;
  const unsigned char* __restrict__ sel = A.w.sel;
  bf16* __restrict__ out = BASE ? A.w.ay : A.w.aty;
  const int64_t step = (int64_t)gridDim.x * 8 * DU;
  for (int64_t w0 = ((int64_t)blockIdx.x * 8 + warp) * DU; w0 < total; w0 += step) {
    int n = (int)(w0 / (HC * WC));
    int rem = (int)(w0 - (int64_t)n * HC * WC);
    int hp = rem / WC, wp = rem - hp * WC;
    uint4 yq[DU], tq[DU], aq[DU], at[DU];
    F8 atf[ATQF32 ? DU : 1];
    uint2 cd[DU];
    int64_t off[DU];
    bool inb[DU], pooled[DU];
#pragma unroll
    for (int u = 0; u < DU; ++u) {
      const int y = 2 * hp + dy, x = 2 * wp + dx;
      inb[u] = w0 + u < total && y < g.H && x < g.W;
      pooled[u] = inb[u] && hp < g.HP && wp < g.WP;
      off[u] = ((((int64_t)n * (g.H + 2) + y + 1) * (g.W + 2)) + x + 1) * 64 + c0;
      if (inb[u]) {
        yq[u] = *reinterpret_cast<const uint4*>(yb + off[u]);
        if (!BASE) tq[u] = *reinterpret_cast<const uint4*>(ty + off[u]);
      }
      if (pooled[u]) {
        const int64_t pi = ((((int64_t)n * g.HP + hp) * g.WP) + wp) * 64 + c0;
        cd[u] = *reinterpret_cast<const uint2*>(sel + pi);
        aq[u] = *reinterpret_cast<const uint4*>(aqm + pi);
        if (!BASE) {
          if (ATQF32)
            atf[ATQF32 ? u : 0] = load_atq8(A, n, hp, wp, c0);
          else
            at[u] = *reinterpret_cast<const uint4*>(A.atq_nhwc + ((((int64_t)n * (g.HP + 2) + hp + 1) * (g.WP + 2)) + wp + 1) * 64 + c0);
        }
      }
      if (++wp == WC) {
        wp = 0;
        if (++hp == HC) {
          hp = 0;
          ++n;
        }
      }
    }
#pragma unroll
    for (int u = 0; u < DU; ++u) {
      if (!inb[u]) continue;
      float yv[8], t[8], a[8], b[8];
      unpack8(yq[u], yv);
      if (!BASE) unpack8(tq[u], t);
      if (pooled[u]) {
        unpack8(aq[u], a);
        if (!BASE && !ATQF32) unpack8(at[u], b);
      }
      F8 v;
      const F8 e0 = ld8f(&cf[0][c0]), e1 = ld8f(&cf[1][c0]);
#pragma unroll
      for (int e = 0; e < 8; ++e) v.v[e] = fmaf(yv[e], e1.v[e], e0.v[e]);
      if (!BASE) {
        const F8 k2 = ld8f(&cf[2][c0]);
#pragma unroll
        for (int e = 0; e < 8; ++e) v.v[e] = fmaf(t[e], k2.v[e], v.v[e]);
      }
      if (pooled[u]) {
        const F8 k3 = ld8f(&cf[3][c0]), k4 = ld8f(&cf[4][c0]);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const unsigned code = ((e < 4 ? cd[u].x : cd[u].y) >> (8 * (e & 3))) & 0xffu;
          if ((code & 3u) == (unsigned)sub) {
            v.v[e] = fmaf(k4.v[e], a[e], v.v[e]);
            if (!BASE && (code & 4u)) v.v[e] = fmaf(k3.v[e], ATQF32 ? atf[ATQF32 ? u : 0].v[e] : b[e], v.v[e]);
          }
        }
      }
      st8(out + off[u], v);
    }
  }
}

}  // namespace

// node layout (plan.py _n_convblock2):
//   dims = N,C(=64),H,W,O(=64),3,3,HO(=H),WO(=W),1,1,1,1,HP,WP,relu      f[0] = eps
//   slot 0 = x_in (base, t, a, at), slot 1 = W (base = the autocast copy, t, at), slot 2 = gamma (base fp32; t, at)
//   slot 3 = q (base, t, a, at)
//   aux[0] = workspace, aux[1] = int64 arg-max indices, aux[2] = y (conv output base; its dtype tag in `ndim`)
//   stride[0][0..3] (unused by this op) carry four more pointers: t_b, at_b, t_beta, at_beta
//   kind bit 0: reduced-precision graph (always set); bit 1: t / a / at of x_in are bf16 padded NHWC;
//        bit 2: t / a / at of q are bf16 padded NHWC
int bb_launch_convblock2(const bb_node& nd, int pass, cudaStream_t s) {
  A2 A{};
  G2& g = A.g;
  g.N = (int)nd.dims[0]; g.H = (int)nd.dims[2]; g.W = (int)nd.dims[3];
  g.HP = (int)nd.dims[13]; g.WP = (int)nd.dims[14]; g.relu = (int)nd.dims[15];
  if (nd.dims[1] != 64 || nd.dims[4] != 64 || nd.dims[5] != 3 || nd.dims[6] != 3 || nd.dims[7] != g.H || nd.dims[8] != g.W ||
      !bb_conv_halo_ok(64, 64, g.H, g.W))
    return BB_ERR_UNSUPPORTED;
  A.w = layout(nd.aux[0], g);
  A.y = nd.aux[2]; A.dty = nd.ndim;          // (dtype tag of y travels in the otherwise unused ndim field)
  A.q = nd.base[3]; A.dtq = nd.dt[3];
  A.idx = reinterpret_cast<const int64_t*>(nd.aux[1]);
  A.gamma = reinterpret_cast<const float*>(nd.base[2]);
  A.eps = (float)nd.f[0];
  void* const* pb = reinterpret_cast<void* const*>(&nd.stride[0][0]);
  A.t_b = reinterpret_cast<const float*>(pb[0]); A.at_b = reinterpret_cast<float*>(pb[1]);
  A.t_beta = reinterpret_cast<const float*>(pb[2]); A.at_beta = reinterpret_cast<float*>(pb[3]);
  A.t_gamma = reinterpret_cast<const float*>(nd.t[2]); A.at_gamma = reinterpret_cast<float*>(nd.at[2]);
  const bool in_nhwc = nd.kind & 2, q_nhwc = nd.kind & 4;
  if (q_nhwc) {
    A.tq_nhwc = reinterpret_cast<bf16*>(nd.t[3]);
    A.aq_nhwc = reinterpret_cast<const bf16*>(nd.a[3]);
    A.atq_nhwc = reinterpret_cast<const bf16*>(nd.at[3]);
  } else {
    A.tq_f32 = reinterpret_cast<float*>(nd.t[3]);
    A.aq_f32 = reinterpret_cast<const float*>(nd.a[3]);
    A.atq_f32 = reinterpret_cast<const float*>(nd.at[3]);
  }
  const int taps = 9;
  const int chunks = g.N < 8 ? g.N : (g.N < 64 ? 8 : 32);
  const dim3 per_channel(64, chunks);
  const int64_t rows = (int64_t)g.N * (g.H + 2) * (g.W + 2), nwin = (int64_t)g.N * g.HP * g.WP;
  auto stream_grid = [](int64_t units) {          // 32 rows / windows per block step, a few waves of persistent blocks
    int64_t b = (units + 31) / 32;
    if (b > 8 * BB_SM_COUNT) b = 8 * BB_SM_COUNT;
    return (int)(b < 1 ? 1 : b);
  };
  const int64_t dense_windows = (int64_t)g.N * ((g.H + 1) / 2) * ((g.W + 1) / 2);
  const int dense_grid = (int)std::min<int64_t>((dense_windows + 8 * DU - 1) / (8 * DU), (int64_t)16 * BB_SM_COUNT);
  const size_t pack_smem = 4 * (size_t)g.W * 65;
  int rc;
  if (pass == BB_PASS_BASE_BWD) {
    BB_CUDA_TRY(cudaMemsetAsync(A.w.dsum, 0, sizeof(double) * 10 * 64, s));
    cb2_ystats_kernel<<<per_channel, 256, 0, s>>>(A);
    cb2_ystats_finish_kernel<<<1, 64, 0, s>>>(A);
    cb2_prep_kernel<<<g.N, 256, (size_t)g.WP * 64 * 9, s>>>(A);
    cb2_pack_padded_kernel<<<dim3(g.N, g.H), 256, pack_smem, s>>>(A.y, A.dty, 64, g.H, g.W, A.w.yb);
    cb2_pack_padded_kernel<<<dim3(g.N, g.H), 256, pack_smem, s>>>(nd.base[0], nd.dt[0], 64, g.H, g.W, A.w.xin);
    A.base = 1;
    cb2_coef_kernel<<<1, 64, 0, s>>>(A);
    cb2_dense_kernel<true, false><<<dense_grid, 256, 0, s>>>(A);
    bb_launch_tally += 8;
    BB_LAUNCH_CHECK();
    if ((rc = bb_pack_convw(nd.base[1], nd.dt[1], 64, 64, taps, 0, A.w.wf, 64, s))) return rc;
    if ((rc = bb_pack_convw(nd.base[1], nd.dt[1], 64, 64, taps, 1, A.w.wd, 64, s))) return rc;
    if (nd.pad0 & 1) {
      // a_in = dgrad(a_y, W): the previous block's base adjoint
      const void* src[1] = {A.w.ay};
      const void* wm[1] = {A.w.wd};
      if (in_nhwc)
        rc = bb_conv_halo_run(g.N, g.H, g.W, 1, src, wm, 1, nullptr, 0, nullptr, s, nd.a[0]);
      else
        rc = bb_conv_halo_run(g.N, g.H, g.W, 1, src, wm, 1, reinterpret_cast<float*>(nd.a[0]), nd.beta[0], nullptr, s);
      if (rc) return rc;
    }
    return BB_OK;
  }
  if (pass == BB_PASS_TAN_FWD) {
    const void* tin = nd.t[0];
    if (!in_nhwc) {
      cb2_pack_padded_kernel<<<dim3(g.N, g.H), 256, pack_smem, s>>>(nd.t[0], BB_F32, 64, g.H, g.W, A.w.tin);
      bb_launch_tally += 1;
      tin = A.w.tin;
    }
    if ((rc = bb_pack_convw(nd.t[1], BB_F32, 64, 64, taps, 0, A.w.twf, 64, s))) return rc;
    const void* src[2] = {tin, A.w.xin};
    const void* wm[2] = {A.w.wf, A.w.twf};
    if ((rc = bb_conv_halo_run(g.N, g.H, g.W, 2, src, wm, 0, nullptr, 0, A.t_b, s, A.w.ty))) return rc;
    BB_CUDA_TRY(cudaMemsetAsync(A.w.dsum + 2 * 64, 0, sizeof(double) * 2 * 64, s));     // this pass's sums
    cb2_stats_kernel<<<stream_grid(rows), 256, 0, s>>>(A);
    cb2_final_kernel<<<stream_grid(nwin), 256, 0, s>>>(A);
    bb_launch_tally += 3;
    BB_LAUNCH_CHECK();
    return BB_OK;
  }
  // ---- tangent backward ----
  A.base = 0;
  BB_CUDA_TRY(cudaMemsetAsync(A.w.dsum + 4 * 64, 0, sizeof(double) * 3 * 64, s));
  cb2_reduce_kernel<<<stream_grid(nwin), 256, 0, s>>>(A);
  cb2_coef_kernel<<<1, 64, 0, s>>>(A);
  if (A.atq_nhwc)
    cb2_dense_kernel<false, false><<<dense_grid, 256, 0, s>>>(A);
  else
    cb2_dense_kernel<false, true><<<dense_grid, 256, 0, s>>>(A);
  bb_launch_tally += 4;
  BB_LAUNCH_CHECK();
  if ((rc = bb_pack_convw(nd.t[1], BB_F32, 64, 64, taps, 1, A.w.twd, 64, s))) return rc;
  {
    const void* src[2] = {A.w.aty, A.w.ay};
    const void* wm[2] = {A.w.wd, A.w.twd};
    if (in_nhwc)
      rc = bb_conv_halo_run(g.N, g.H, g.W, 2, src, wm, 1, nullptr, 0, nullptr, s, nd.at[0]);
    else
      rc = bb_conv_halo_run(g.N, g.H, g.W, 2, src, wm, 1, reinterpret_cast<float*>(nd.at[0]), nd.beta[0], nullptr, s);
    if (rc) return rc;
  }
  const void* tin = in_nhwc ? nd.t[0] : (const void*)A.w.tin;     // packed by this iteration's tangent-forward pass
  const void* xs[2] = {A.w.xin, tin};
  const void* gs[2] = {A.w.aty, A.w.ay};
  return bb_wgrad_halo_run(g.N, g.H, g.W, 64, 64, 2, xs, gs, reinterpret_cast<float*>(nd.at[1]), s);
}

extern "C" int64_t bb_convblock2_ws_bytes(int N, int C, int H, int W, int O, int HO, int WO, int HP, int WP) {
  G2 g{};
  (void)C; (void)O; (void)HO; (void)WO;
  g.N = N; g.H = H; g.W = W; g.HP = HP; g.WP = WP;
  return (int64_t)layout(nullptr, g).bytes;
}
