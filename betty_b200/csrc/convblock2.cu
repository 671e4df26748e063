// Fused second-order rules of an INNER convolution block of a bf16-autocast graph
//
//     x_in --conv3x3(W,b)--> y --BatchNorm2d(batch stats; gamma,beta)--> z --[ReLU]--> MaxPool2d(2)--> q
//
// (blocks 2-4 of reference examples/implicit_maml/models.py:9-24).  The convolution products run on the TMA-fed
// tcgen05 kernels of conv_tma.cu / gemm_tma.cu; everything between them -- BatchNorm statistics, BatchNorm apply,
// ReLU mask, pooling -- is restructured so that per iteration
//   * the pooled-side quantities (a_q, at_q) are SPARSE on the conv-output grid (one pixel per window and channel), so
//     the BatchNorm adjoint statistics m1, m2, mt1, mt2 and the gamma/beta slices of H.d are sums over POOLED arrays;
//   * the conv-output-sized adjoint tangent at_y = sparse + d0 + xhat d1 + t_y d2 (norm.cu's rule expanded, with
//     per-channel d*) is produced once, directly as the bf16 NHWC operand the input-gradient and weight-gradient
//     kernels load by TMA -- no fp32 adjoint buffer, no pack kernel;
//   * the tangent t_q is written directly as the next block's bf16 NHWC operand.
// Unfused, the same block costs three BatchNorm sweeps + two pooling kernels over fp32 NCHW buffers plus four pack
// kernels per iteration (profiles/r01_plan_profile_maml.md).
//
//   TF   t_y = conv(t_in, W) + conv(x_in, t_W) + t_b                                  (tensor cores)
//        mean_t = mean(t_y), sdot = mean(xhat t_y) - mean_t mean(xhat)                  cb2_stats_kernel (+ raw select)
//        t_q = mask (gamma rstd (t_y* - mean_t - xhat* sdot) + t_gamma xhat* + t_beta)  cb2_final_kernel
//   TB   pooled sums S_at, S_atxh, S_adxh                                             cb2_reduce_kernel
//        d0, d1, d2, cw, cd; at_gamma, at_beta, at_b                                  cb2_coef_kernel
//        at_y (bf16 NHWC)                                                             cb2_dense_kernel
//        at_in = dgrad(at_y, W) + dgrad(a_y, t_W);  at_W += wgrad(at_y, x_in) + wgrad(a_y, t_in)   (tensor cores)
#include <cuda_bf16.h>
#include <stdlib.h>

#include "../../include/betty_b200.h"
#include "bb_common.cuh"
#include "conv_halo.h"
#include "conv_tma.h"
#include "gemm_tma.h"
#include "plan.h"
#include "tma.h"

namespace {

constexpr int NT = 256;
constexpr int MAXC = 64;

struct G2 {
  int N, C, H, W, O, HO, WO, ph, pw, HP, WP, relu;
};

struct Ws2 {
  double* dsum;        // [2*O] sum y, sum y^2 | [2*O] sum t_y, sum xhat t_y | [3*O] S_at, S_atxh, S_adxh | [2*O] Sa, Saxh | [O] sum xhat
  float* mean;         // [O]
  float* rstd;         // [O]
  float* coef;         // [8*O]: d0, d1, d2, cw, cd, mean_t, sdot, (spare)
  unsigned char* sel;  // [N*O*HP*WP]  NCHW pooled: code | mask << 2
  float* xh;           // pooled: xhat at the arg-max pixel
  float* dxh;          // pooled: dxhat at the arg-max pixel (per iteration)
  float* aqm;          // pooled: mask * a_q
  float* tys;          // pooled: t_y at the arg-max pixel (per iteration)
  float* ty;           // [N*O*HO*WO] fp32 NCHW: conv tangent (per iteration)
  __nv_bfloat16* aty;  // [N][HO+2][WO+2][64] bf16: adjoint tangent at the conv output (per iteration)
  __nv_bfloat16* ay;   // [N][HO+2][WO+2][64] bf16: base adjoint at the conv output (per call)
  __nv_bfloat16* xin;  // [N][H+2][W+2][64]   bf16: base input (per call)
  __nv_bfloat16* tin;  // [N][H+2][W+2][64]   bf16: packed input tangent when the producer wrote fp32 NCHW
  __nv_bfloat16* wf;   // [64*taps*64] forward operand of W        (per call)
  __nv_bfloat16* wd;   // [64*taps*64] input-gradient operand of W (per call)
  __nv_bfloat16* twf;  // same for t_W (per iteration)
  __nv_bfloat16* twd;
  size_t bytes;
};

inline size_t up(size_t x) { return (x + 1023) & ~(size_t)1023; }

Ws2 layout(void* base, const G2& g) {
  Ws2 w{};
  size_t at = 0;
  uint8_t* b = reinterpret_cast<uint8_t*>(base);
  auto take = [&](size_t bytes) { size_t o = at; at = up(at + bytes); return b ? b + o : nullptr; };
  // TMA operands live in the PADDED NHWC layout [N][H+2][W+2][64] (zero border, never written): a tap displacement is
  // a constant row offset there, which is what the halo-resident convolution (conv_halo.cu) needs
  const size_t O = g.O, pooled = (size_t)g.N * g.O * g.HP * g.WP, full = (size_t)g.N * g.HO * g.WO,
               fullp = (size_t)g.N * (g.HO + 2) * (g.WO + 2), finp = (size_t)g.N * (g.H + 2) * (g.W + 2);
  w.dsum = reinterpret_cast<double*>(take(8 * 10 * O));
  w.mean = reinterpret_cast<float*>(take(4 * O));
  w.rstd = reinterpret_cast<float*>(take(4 * O));
  w.coef = reinterpret_cast<float*>(take(4 * 8 * O));
  w.sel = reinterpret_cast<unsigned char*>(take(pooled));
  w.xh = reinterpret_cast<float*>(take(4 * pooled));
  w.dxh = reinterpret_cast<float*>(take(4 * pooled));
  w.aqm = reinterpret_cast<float*>(take(4 * pooled));
  w.tys = reinterpret_cast<float*>(take(4 * pooled));
  w.ty = reinterpret_cast<float*>(take(4 * full * O));
  w.aty = reinterpret_cast<__nv_bfloat16*>(take(2 * fullp * 64));
  w.ay = reinterpret_cast<__nv_bfloat16*>(take(2 * fullp * 64));
  w.xin = reinterpret_cast<__nv_bfloat16*>(take(2 * finp * 64));
  w.tin = reinterpret_cast<__nv_bfloat16*>(take(2 * finp * 64));
  const size_t wb = (size_t)2 * 64 * 9 * 64;
  w.wf = reinterpret_cast<__nv_bfloat16*>(take(wb));
  w.wd = reinterpret_cast<__nv_bfloat16*>(take(wb));
  w.twf = reinterpret_cast<__nv_bfloat16*>(take(wb));
  w.twd = reinterpret_cast<__nv_bfloat16*>(take(wb));
  w.bytes = at;
  return w;
}

struct A2 {
  G2 g;
  Ws2 w;
  const void* y; int dty;
  const void* q; int dtq;
  const int64_t* idx;
  const float* gamma;
  float eps;
  const float *t_b, *t_gamma, *t_beta;
  float *at_b, *at_gamma, *at_beta;
  float* tq;                 // fp32 NCHW pooled tangent (standard plan buffer), or
  __nv_bfloat16* tq_nhwc;    // bf16 padded NHWC [N][HP+2][WP+2][64] when the consumer is a fused block
  const float* a_q;
  const float* at_q;
  int base;                  // dense kernel: 1 = base adjoint a_y (per call), 0 = adjoint tangent at_y
};

// ---- once per call -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) cb2_ystats_kernel(const A2 A) {
  __shared__ double red[32];
  const int o = blockIdx.x, HW = A.g.HO * A.g.WO;
  double s0 = 0, s1 = 0;
  for (int n = blockIdx.y; n < A.g.N; n += gridDim.y) {
    const int64_t base = ((int64_t)n * A.g.O + o) * HW;
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
    atomicAdd(&A.w.dsum[A.g.O + o], s1);
  }
}

__global__ void cb2_ystats_finish_kernel(const A2 A) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= A.g.O) return;
  const double cnt = (double)A.g.N * A.g.HO * A.g.WO;
  const double mean = A.w.dsum[o] / cnt;
  const double var = A.w.dsum[A.g.O + o] / cnt - mean * mean;
  A.w.mean[o] = (float)mean;
  A.w.rstd[o] = (float)rsqrt((var > 0 ? var : 0) + (double)A.eps);
}

// pooled NCHW: codes, mask, xhat*, masked base adjoint; per-channel Sa, Saxh; per-channel sum of xhat over y
__global__ void __launch_bounds__(256) cb2_prep_kernel(const A2 A) {
  __shared__ double red[32];
  const G2& g = A.g;
  const int o = blockIdx.x, PW = g.HP * g.WP, HW = g.HO * g.WO;
  const float mean = A.w.mean[o], rstd = A.w.rstd[o];
  double sa = 0, saxh = 0, sxh = 0;
  for (int n = blockIdx.y; n < g.N; n += gridDim.y) {
    const int64_t pbase = ((int64_t)n * g.O + o) * PW, ybase = ((int64_t)n * g.O + o) * HW;
    for (int i = threadIdx.x; i < PW; i += blockDim.x) {
      const int hp = i / g.WP, wp = i - hp * g.WP;
      const int64_t id = A.idx[pbase + i];
      const int iy = (int)(id / g.WO), ix = (int)(id - (int64_t)iy * g.WO);
      const bool m = g.relu ? (bb::ldf(A.q, pbase + i, A.dtq) > 0.f) : true;
      const float xh = (bb::ldf(A.y, ybase + id, A.dty) - mean) * rstd;
      const float aq = m ? A.a_q[pbase + i] : 0.f;
      A.w.sel[pbase + i] = (unsigned char)(((iy - 2 * hp) & 1) * 2 + ((ix - 2 * wp) & 1) + (m ? 4 : 0));
      A.w.xh[pbase + i] = xh;
      A.w.aqm[pbase + i] = aq;
      sa += aq;
      saxh += (double)aq * xh;
    }
    for (int i = threadIdx.x; i < HW; i += blockDim.x) sxh += (bb::ldf(A.y, ybase + i, A.dty) - mean) * rstd;
  }
  sa = bb::block_sum<double>(sa, red);
  saxh = bb::block_sum<double>(saxh, red);
  sxh = bb::block_sum<double>(sxh, red);
  if (threadIdx.x == 0) {
    atomicAdd(&A.w.dsum[7 * g.O + o], sa);
    atomicAdd(&A.w.dsum[8 * g.O + o], saxh);
    atomicAdd(&A.w.dsum[9 * g.O + o], sxh);
  }
}

// ---- tangent forward -----------------------------------------------------------------------------------------------
// per (channel, image) plane of t_y (fp32 NCHW, just written by the convolution): sum t_y, sum xhat t_y, and the raw
// value at each window's arg-max pixel
__global__ void __launch_bounds__(256) cb2_stats_kernel(const A2 A) {
  __shared__ double red[32];
  const G2& g = A.g;
  const int o = blockIdx.x, PW = g.HP * g.WP, HW = g.HO * g.WO;
  const float mean = A.w.mean[o], rstd = A.w.rstd[o];
  float s0 = 0.f, s1 = 0.f;
  const bool vec = (HW & 3) == 0 && A.dty == BB_BF16;
  for (int n = blockIdx.y; n < g.N; n += gridDim.y) {
    const int64_t pbase = ((int64_t)n * g.O + o) * PW, ybase = ((int64_t)n * g.O + o) * HW;
    const float* ty = A.w.ty + ybase;
    if (vec) {
      // plane offsets are multiples of 4 elements: 128-bit t_y loads, 64-bit loads of the bf16 base activation
      const uint2* yb = reinterpret_cast<const uint2*>(reinterpret_cast<const unsigned short*>(A.y) + ybase);
      for (int i = threadIdx.x; i < (HW >> 2); i += blockDim.x) {
        const float4 t = bb::ld4(ty + 4 * i);
        const uint2 r = yb[i];
        const float y0 = __uint_as_float(r.x << 16), y1 = __uint_as_float(r.x & 0xffff0000u);
        const float y2 = __uint_as_float(r.y << 16), y3 = __uint_as_float(r.y & 0xffff0000u);
        s0 += (t.x + t.y) + (t.z + t.w);
        s1 = fmaf((y0 - mean) * rstd, t.x, s1);
        s1 = fmaf((y1 - mean) * rstd, t.y, s1);
        s1 = fmaf((y2 - mean) * rstd, t.z, s1);
        s1 = fmaf((y3 - mean) * rstd, t.w, s1);
      }
    } else {
      for (int i = threadIdx.x; i < HW; i += blockDim.x) {
        const float t = ty[i];
        const float xh = (bb::ldf(A.y, ybase + i, A.dty) - mean) * rstd;
        s0 += t;
        s1 = fmaf(xh, t, s1);
      }
    }
    for (int i = threadIdx.x; i < PW; i += blockDim.x) {
      const int hp = i / g.WP, wp = i - hp * g.WP;
      const unsigned code = A.w.sel[pbase + i];
      A.w.tys[pbase + i] = ty[(2 * hp + ((code >> 1) & 1)) * g.WO + 2 * wp + (code & 1)];   // same plane: L1 / L2 hits
    }
  }
  const double d0 = bb::block_sum<double>((double)s0, red);
  const double d1 = bb::block_sum<double>((double)s1, red);
  if (threadIdx.x == 0) {
    atomicAdd(&A.w.dsum[2 * g.O + o], d0);
    atomicAdd(&A.w.dsum[3 * g.O + o], d1);
  }
}

// pooled finalize: dxhat*, t_q (fp32 NCHW, or bf16 padded NHWC).  grid (N, HP): one pooled row of every channel per
// block; per-channel constants once per block, warp per channel, lanes along the row
__global__ void __launch_bounds__(256) cb2_final_kernel(const A2 A) {
  extern __shared__ float sm[];                 // [O][8] constants | [WP][O + 1] transposed tile for the NHWC store
  const G2& g = A.g;
  const int n = blockIdx.x, hp = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* cst = sm;
  float* tile = sm + 8 * g.O;
  if ((int)threadIdx.x < g.O) {
    const int o = threadIdx.x;
    const double P = (double)g.N * g.HO * g.WO;
    const double mt = A.w.dsum[2 * g.O + o] / P;
    const float mean_t = (float)mt;
    const float sdot = (float)((A.w.dsum[3 * g.O + o] - mt * A.w.dsum[9 * g.O + o]) / P);
    const float rstd = A.w.rstd[o];
    const float gam = A.gamma ? A.gamma[o] : 1.f, tgam = A.t_gamma ? A.t_gamma[o] : 0.f, tbet = A.t_beta ? A.t_beta[o] : 0.f;
    cst[o * 8 + 0] = rstd;                       // dxhat = rstd * (t_y* - mean_t - xhat* sdot)
    cst[o * 8 + 1] = mean_t;
    cst[o * 8 + 2] = sdot;
    cst[o * 8 + 3] = gam;
    cst[o * 8 + 4] = tgam;
    cst[o * 8 + 5] = tbet;
    if (n == 0 && hp == 0) {
      A.w.coef[5 * g.O + o] = mean_t;
      A.w.coef[6 * g.O + o] = sdot;
    }
  }
  __syncthreads();
  for (int o = warp; o < g.O; o += 8) {
    const float rstd = cst[o * 8], mean_t = cst[o * 8 + 1], sdot = cst[o * 8 + 2], gam = cst[o * 8 + 3],
                tgam = cst[o * 8 + 4], tbet = cst[o * 8 + 5];
    const int64_t row = (((int64_t)n * g.O + o) * g.HP + hp) * g.WP;
    for (int wp = lane; wp < g.WP; wp += 32) {
      const float xh = A.w.xh[row + wp];
      const float dxh = (A.w.tys[row + wp] - mean_t - xh * sdot) * rstd;
      A.w.dxh[row + wp] = dxh;
      const float tq = (A.w.sel[row + wp] & 4) ? fmaf(gam, dxh, fmaf(tgam, xh, tbet)) : 0.f;
      if (A.tq_nhwc)
        tile[wp * (g.O + 1) + o] = tq;
      else
        A.tq[row + wp] = tq;
    }
  }
  if (A.tq_nhwc) {
    __syncthreads();
    __nv_bfloat16* dst = A.tq_nhwc + ((((int64_t)n * (g.HP + 2) + hp + 1) * (g.WP + 2)) + 1) * 64;
    for (int i = threadIdx.x; i < g.WP * 32; i += blockDim.x) {        // two channels per thread
      const int wp = i >> 5, o = (i & 31) * 2;
      const float v0 = o < g.O ? tile[wp * (g.O + 1) + o] : 0.f, v1 = o + 1 < g.O ? tile[wp * (g.O + 1) + o + 1] : 0.f;
      reinterpret_cast<__nv_bfloat162*>(dst)[i] = __floats2bfloat162_rn(v0, v1);
    }
  }
}

// ---- tangent backward ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) cb2_reduce_kernel(const A2 A) {
  __shared__ double red[32];
  const G2& g = A.g;
  const int o = blockIdx.x, PW = g.HP * g.WP;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f;
  for (int n = blockIdx.y; n < g.N; n += gridDim.y) {
    const int64_t pbase = ((int64_t)n * g.O + o) * PW;
    for (int i = threadIdx.x; i < PW; i += blockDim.x) {
      const float v = (A.w.sel[pbase + i] & 4) ? A.at_q[pbase + i] : 0.f;
      const float xh = A.w.xh[pbase + i];
      s0 += v;
      s1 = fmaf(v, xh, s1);
      s2 = fmaf(A.w.aqm[pbase + i], A.w.dxh[pbase + i], s2);
    }
  }
  const double d0 = bb::block_sum<double>((double)s0, red);
  const double d1 = bb::block_sum<double>((double)s1, red);
  const double d2 = bb::block_sum<double>((double)s2, red);
  if (threadIdx.x == 0) {
    atomicAdd(&A.w.dsum[4 * g.O + o], d0);
    atomicAdd(&A.w.dsum[5 * g.O + o], d1);
    atomicAdd(&A.w.dsum[6 * g.O + o], d2);
  }
}

// per-channel coefficients of the dense rule; gamma / beta / conv-bias slices of H.d.  BASE: the base adjoint a_y.
__global__ void cb2_coef_kernel(const A2 A) {
  const G2& g = A.g;
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= g.O) return;
  const double P = (double)g.N * g.HO * g.WO;
  const float rstd = A.w.rstd[o];
  const float gam = A.gamma ? A.gamma[o] : 1.f;
  const double Sa = A.w.dsum[7 * g.O + o], Saxh = A.w.dsum[8 * g.O + o], sx = A.w.dsum[9 * g.O + o];
  const float m1 = (float)(gam * Sa / P), m2 = (float)(gam * Saxh / P);
  float* c = A.w.coef;
  if (A.base) {
    // a_y = rstd (gamma a_z - m1 - xhat m2)
    c[0 * g.O + o] = -rstd * m1;
    c[1 * g.O + o] = -rstd * m2;
    c[2 * g.O + o] = 0.f;
    c[3 * g.O + o] = 0.f;          // coefficient of at_q
    c[4 * g.O + o] = rstd * gam;   // coefficient of mask a_q
    return;
  }
  const float tgam = A.t_gamma ? A.t_gamma[o] : 0.f;
  const float mean_t = c[5 * g.O + o], sdot = c[6 * g.O + o];
  const double S_at = A.w.dsum[4 * g.O + o], S_atxh = A.w.dsum[5 * g.O + o], S_adxh = A.w.dsum[6 * g.O + o];
  const float mt1 = (float)((gam * S_at + tgam * Sa) / P);
  const float mt2 = (float)((gam * S_atxh + tgam * Saxh + gam * S_adxh) / P);
  const float d2 = -rstd * rstd * m2;
  const float d1 = -rstd * mt2 + 2.f * rstd * rstd * sdot * m2;
  const float d0 = -rstd * mt1 + rstd * rstd * sdot * m1 + rstd * rstd * m2 * mean_t;
  const float cw = rstd * gam, cd = rstd * tgam - rstd * rstd * sdot * gam;
  c[0 * g.O + o] = d0; c[1 * g.O + o] = d1; c[2 * g.O + o] = d2; c[3 * g.O + o] = cw; c[4 * g.O + o] = cd;
  if (A.at_gamma) A.at_gamma[o] += (float)(S_atxh + S_adxh);
  if (A.at_beta) A.at_beta[o] += (float)S_at;
  if (A.at_b) A.at_b[o] += (float)((double)cw * S_at + (double)cd * Sa + (double)d0 * P + (double)d1 * sx + (double)d2 * P * mean_t);
}

// dense rule on the conv-output grid, written as the bf16 padded-NHWC TMA operand.  grid (N, ceil(HO/2)): the two
// output rows of one pooled row, every channel, per block; warp per channel (its coefficients are warp-uniform),
// lanes along the two rows (NCHW reads coalesced along x), the store is the transposed tile.
__global__ void __launch_bounds__(256) cb2_dense_kernel(const A2 A) {
  extern __shared__ float sm[];                 // [2][WO][O + 1]
  const G2& g = A.g;
  const int n = blockIdx.x, hpb = blockIdx.y;
  const int hy0 = 2 * hpb;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float* c = A.w.coef;
  const bool in_pool = hpb < g.HP;
  const int rows = (hy0 + 1 < g.HO) ? 2 : 1;
  for (int o = warp; o < g.O; o += 8) {
    const float c0 = c[0 * g.O + o], c1 = c[1 * g.O + o], c2 = c[2 * g.O + o], c3 = c[3 * g.O + o], c4 = c[4 * g.O + o];
    const float mean = A.w.mean[o], rstd = A.w.rstd[o];
    const int64_t ybase = (((int64_t)n * g.O + o) * g.HO + hy0) * g.WO;
    const int64_t pbase = (((int64_t)n * g.O + o) * g.HP + hpb) * g.WP;
    for (int e = lane; e < rows * g.WO; e += 32) {
      const int dy = e >= g.WO ? 1 : 0, x = e - dy * g.WO;
      const float xh = (bb::ldf(A.y, ybase + e, A.dty) - mean) * rstd;
      float v = fmaf(xh, c1, c0);
      if (!A.base) v = fmaf(A.w.ty[ybase + e], c2, v);
      const int wp = x >> 1;
      if (in_pool && wp < g.WP) {
        const unsigned code = A.w.sel[pbase + wp];
        if ((int)((code >> 1) & 1) == dy && (int)(code & 1) == (x & 1)) {
          v = fmaf(c4, A.w.aqm[pbase + wp], v);
          if (!A.base && (code & 4)) v = fmaf(c3, A.at_q[pbase + wp], v);
        }
      }
      sm[e * (g.O + 1) + o] = v;
    }
  }
  __syncthreads();
  __nv_bfloat16* base = A.base ? A.w.ay : A.w.aty;
  for (int dy = 0; dy < rows; ++dy) {
    __nv_bfloat16* dst = base + ((((int64_t)n * (g.HO + 2) + hy0 + dy + 1) * (g.WO + 2)) + 1) * 64;
    const float* src = sm + dy * g.WO * (g.O + 1);
    for (int i = threadIdx.x; i < g.WO * 32; i += blockDim.x) {
      const int x = i >> 5, o = (i & 31) * 2;
      const float v0 = o < g.O ? src[x * (g.O + 1) + o] : 0.f, v1 = o + 1 < g.O ? src[x * (g.O + 1) + o + 1] : 0.f;
      reinterpret_cast<__nv_bfloat162*>(dst)[i] = __floats2bfloat162_rn(v0, v1);
    }
  }
}

// NCHW (any base dtype) -> bf16 padded NHWC [N][H+2][W+2][64].  grid (N, H)
__global__ void __launch_bounds__(256) cb2_pack_padded_kernel(const void* src, int dt, int C, int H, int W,
                                                              __nv_bfloat16* dst) {
  extern __shared__ float sm[];                 // [W][C + 1]
  const int n = blockIdx.x, y = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int c = warp; c < C; c += 8) {
    const int64_t row = (((int64_t)n * C + c) * H + y) * W;
    for (int x = lane; x < W; x += 32) sm[x * (C + 1) + c] = bb::ldf(src, row + x, dt);
  }
  __syncthreads();
  __nv_bfloat16* d = dst + ((((int64_t)n * (H + 2) + y + 1) * (W + 2)) + 1) * 64;
  for (int i = threadIdx.x; i < W * 32; i += blockDim.x) {
    const int x = i >> 5, c = (i & 31) * 2;
    const float v0 = c < C ? sm[x * (C + 1) + c] : 0.f, v1 = c + 1 < C ? sm[x * (C + 1) + c + 1] : 0.f;
    reinterpret_cast<__nv_bfloat162*>(d)[i] = __floats2bfloat162_rn(v0, v1);
  }
}

// forward-form product over padded operands: the halo-resident kernel where it applies, else the per-tap TMA kernel
int corr(const G2& g, const BbConvGeo& cg, int npairs, const void* const* src, int SH, int SW, const void* const* wm, int ncols,
         int flip, float* out, int beta, const float* bias, cudaStream_t s) {
  if (g.C == 64 && g.O == 64 && bb_conv_halo_ok(64, 64, SH, SW))
    return bb_conv_halo_run(g.N, SH, SW, npairs, src, wm, flip, out, beta, bias, s);
  return bb_conv_tma_corr(cg, npairs, src, SH, SW, wm, ncols, SH, SW, flip, out, beta, bias, s, true);
}

}  // namespace

// node layout (plan.py _n_convblock2):
//   dims = N,C,H,W,O,KH,KW,HO,WO,sh,sw,ph,pw,HP,WP,relu      f[0] = eps
//   slot 0 = x_in (base, t, a, at), slot 1 = W (base = bf16/fp32 weights, t, at), slot 2 = gamma (base fp32; t, at)
//   slot 3 = q (base, t, a, at)
//   aux[0] = workspace, aux[1] = int64 arg-max indices, aux[2] = y (conv output base; its dtype tag in `ndim`)
//   stride[0][0..3] (unused by this op) carry four more pointers: t_b, at_b, t_beta, at_beta
//   kind bit 0: reduced-precision graph (always set), bit 1: t of x_in is bf16 NHWC, bit 2: t of q is bf16 NHWC
int bb_launch_convblock2(const bb_node& nd, int pass, cudaStream_t s) {
  A2 A{};
  G2& g = A.g;
  g.N = (int)nd.dims[0]; g.C = (int)nd.dims[1]; g.H = (int)nd.dims[2]; g.W = (int)nd.dims[3]; g.O = (int)nd.dims[4];
  g.HO = (int)nd.dims[7]; g.WO = (int)nd.dims[8]; g.ph = (int)nd.dims[11]; g.pw = (int)nd.dims[12];
  g.HP = (int)nd.dims[13]; g.WP = (int)nd.dims[14]; g.relu = (int)nd.dims[15];
  if (nd.dims[5] != 3 || nd.dims[6] != 3 || g.O > MAXC || g.C > MAXC || g.WO > 64) return BB_ERR_UNSUPPORTED;
  A.w = layout(nd.aux[0], g);
  A.y = nd.aux[2]; A.dty = nd.ndim;          // (dtype tag of y travels in the otherwise unused ndim field)
  A.q = nd.base[3]; A.dtq = nd.dt[3];
  A.idx = reinterpret_cast<const int64_t*>(nd.aux[1]);
  A.gamma = reinterpret_cast<const float*>(nd.base[2]);
  A.eps = (float)nd.f[0];
  void* const* pb = reinterpret_cast<void* const*>(&nd.stride[0][0]);   // host-side pointer block (see plan.py)
  A.t_b = reinterpret_cast<const float*>(pb[0]); A.at_b = reinterpret_cast<float*>(pb[1]);
  A.t_beta = reinterpret_cast<const float*>(pb[2]); A.at_beta = reinterpret_cast<float*>(pb[3]);
  A.t_gamma = reinterpret_cast<const float*>(nd.t[2]); A.at_gamma = reinterpret_cast<float*>(nd.at[2]);
  const bool tin_nhwc = nd.kind & 2, tq_nhwc = nd.kind & 4;
  A.tq = tq_nhwc ? nullptr : reinterpret_cast<float*>(nd.t[3]);
  A.tq_nhwc = tq_nhwc ? reinterpret_cast<__nv_bfloat16*>(nd.t[3]) : nullptr;
  A.a_q = reinterpret_cast<const float*>(nd.a[3]);
  A.at_q = reinterpret_cast<const float*>(nd.at[3]);
  const int taps = 9;
  BbConvGeo cg{g.N, g.C, g.H, g.W, g.O, 3, 3, g.HO, g.WO, g.ph, g.pw};
  const int chunks = g.N < 8 ? g.N : (g.N < 64 ? 8 : 32);
  const dim3 per_channel(g.O, chunks);
  const size_t tile_q = 4 * ((size_t)g.WP * (g.O + 1) + 8 * g.O), tile_y = 4 * (size_t)2 * g.WO * (g.O + 1);
  const dim3 dense_grid(g.N, (g.HO + 1) / 2);
  int rc;
  if (pass == BB_PASS_BASE_BWD) {
    BB_CUDA_TRY(cudaMemsetAsync(A.w.dsum, 0, sizeof(double) * 10 * g.O, s));
    cb2_ystats_kernel<<<per_channel, 256, 0, s>>>(A);
    cb2_ystats_finish_kernel<<<1, 64, 0, s>>>(A);
    cb2_prep_kernel<<<per_channel, 256, 0, s>>>(A);
    A.base = 1;
    cb2_coef_kernel<<<1, 64, 0, s>>>(A);
    cb2_dense_kernel<<<dense_grid, 256, tile_y, s>>>(A);
    bb_launch_tally += 6;
    BB_LAUNCH_CHECK();
    // per-call operand packs: x_in (NHWC bf16), W in its forward and input-gradient layouts
    cb2_pack_padded_kernel<<<dim3(g.N, g.H), 256, 4 * (size_t)g.W * (g.C + 1), s>>>(nd.base[0], nd.dt[0], g.C, g.H, g.W, A.w.xin);
    bb_launch_tally += 1;
    if ((rc = bb_pack_convw(nd.base[1], nd.dt[1], g.O, g.C, taps, 0, A.w.wf, 64, s))) return rc;
    if ((rc = bb_pack_convw(nd.base[1], nd.dt[1], g.O, g.C, taps, 1, A.w.wd, 64, s))) return rc;
    if (nd.pad0 & 1) {
      // a_in (beta) = dgrad(a_y, W): the previous block's base adjoint
      const void* src[1] = {A.w.ay};
      const void* wm[1] = {A.w.wd};
      if ((rc = corr(g, cg, 1, src, g.HO, g.WO, wm, g.C, 1, reinterpret_cast<float*>(nd.a[0]), nd.beta[0], nullptr, s)))
        return rc;
    }
    return BB_OK;
  }
  if (pass == BB_PASS_TAN_FWD) {
    const void* tin = nd.t[0];
    if (!tin_nhwc) {
      cb2_pack_padded_kernel<<<dim3(g.N, g.H), 256, 4 * (size_t)g.W * (g.C + 1), s>>>(nd.t[0], BB_F32, g.C, g.H, g.W, A.w.tin);
      bb_launch_tally += 1;
      tin = A.w.tin;
    }
    if ((rc = bb_pack_convw(nd.t[1], BB_F32, g.O, g.C, taps, 0, A.w.twf, 64, s))) return rc;
    const void* src[2] = {tin, A.w.xin};
    const void* wm[2] = {A.w.wf, A.w.twf};
    if ((rc = corr(g, cg, 2, src, g.H, g.W, wm, g.O, 0, A.w.ty, 0, A.t_b, s))) return rc;
    BB_CUDA_TRY(cudaMemsetAsync(A.w.dsum + 2 * g.O, 0, sizeof(double) * 2 * g.O, s));     // this pass's sums
    cb2_stats_kernel<<<per_channel, 256, 0, s>>>(A);
    cb2_final_kernel<<<dim3(g.N, g.HP), 256, tile_q, s>>>(A);
    bb_launch_tally += 3;
    BB_LAUNCH_CHECK();
    return BB_OK;
  }
  // ---- tangent backward ----
  A.base = 0;
  BB_CUDA_TRY(cudaMemsetAsync(A.w.dsum + 4 * g.O, 0, sizeof(double) * 3 * g.O, s));
  cb2_reduce_kernel<<<per_channel, 256, 0, s>>>(A);
  cb2_coef_kernel<<<1, 64, 0, s>>>(A);
  cb2_dense_kernel<<<dense_grid, 256, tile_y, s>>>(A);
  bb_launch_tally += 4;
  BB_LAUNCH_CHECK();
  if ((rc = bb_pack_convw(nd.t[1], BB_F32, g.O, g.C, taps, 1, A.w.twd, 64, s))) return rc;
  {
    const void* src[2] = {A.w.aty, A.w.ay};
    const void* wm[2] = {A.w.wd, A.w.twd};
    if ((rc = corr(g, cg, 2, src, g.HO, g.WO, wm, g.C, 1, reinterpret_cast<float*>(nd.at[0]), nd.beta[0], nullptr, s)))
      return rc;
  }
  const void* tin = tin_nhwc ? nd.t[0] : (const void*)A.w.tin;     // packed by this iteration's tangent-forward pass
  const void* xs[2] = {A.w.xin, tin};
  const void* gs[2] = {A.w.aty, A.w.ay};
  static const bool no_wh = getenv("BB200_NO_WGRAD_HALO") != nullptr;
  if (!no_wh && bb_conv_halo_ok(64, 64, g.H, g.W))
    return bb_wgrad_halo_run(g.N, g.H, g.W, g.C, g.O, 2, xs, gs, reinterpret_cast<float*>(nd.at[1]), s);
  return bb_conv_tma_wgrad(cg, 2, xs, gs, reinterpret_cast<float*>(nd.at[1]), s, true);
}

extern "C" int64_t bb_convblock2_ws_bytes(int N, int C, int H, int W, int O, int HO, int WO, int HP, int WP) {
  G2 g{};
  g.N = N; g.C = C; g.H = H; g.W = W; g.O = O; g.HO = HO; g.WO = WO; g.HP = HP; g.WP = WP;
  return (int64_t)layout(nullptr, g).bytes;
}
