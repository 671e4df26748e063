// K7 / K8: second-order rules of BatchNorm2d with batch statistics (NCHW) and LayerNorm (last dim).
//   y = gamma * xhat + beta,  xhat = (x - mu) * rstd
//   TF   tc = t_x - mean(t_x); sdot = mean(xhat*tc); dxhat = (tc - xhat*sdot)*rstd
//        t_y = gamma*dxhat + t_gamma*xhat + t_beta
//   BB   gh = a_y*gamma; a_x = rstd*(gh - mean(gh) - xhat*mean(gh*xhat))
//   TB   ght = at_y*gamma + a_y*t_gamma; u = gh - m1 - xhat*m2
//        ut = ght - mean(ght) - dxhat*m2 - xhat*mean(ght*xhat + gh*dxhat)
//        at_x = -rstd^2*sdot*u + rstd*ut ; at_gamma = sum(at_y*xhat + a_y*dxhat) ; at_beta = sum(at_y)
// (verified against jvp-of-vjp in float64 through oracle/plan_interp.py
// _tf_norm/_bb_norm/_tb_norm.)  BatchNorm statistics couple every pixel of a channel, so each pass is
// a per-channel reduction kernel followed by an apply kernel; LayerNorm rows fit in one warp.
#include "../../include/betty_b200.h"
#include "bb_common.cuh"
#include "plan.h"

namespace {

// ---------------------------------------------------------------------------------------------------
// BatchNorm.  scratch: double[16][C]; slots 0,1 stats | 2,3 BB | 4,5 TF | 6..9 TB
// ---------------------------------------------------------------------------------------------------
struct BnArgs {
  const void* x;
  int dtx;
  const float* gamma;   // may be null (affine=False) -> 1
  const float* tx;
  const float* tgamma;  // may be null
  const float* tbeta;   // may be null
  const float* g;       // a_y
  const float* gt;      // at_y
  float* ty;
  float* dst;           // a_x or at_x
  float* at_gamma;
  float* at_beta;
  double* S;
  int N, C, HW;
  float eps;
  int beta;
};

enum { BN_STATS = 0, BN_BB = 1, BN_TF = 2, BN_TB = 3 };

struct ChanConst {
  float mean, rstd, m1, m2, mean_t, sdot, gamma, tgamma, tbeta;
};

__device__ __forceinline__ ChanConst chan_const(const BnArgs& A, int c) {
  ChanConst k;
  const double cnt = (double)A.N * A.HW;
  const double mean = A.S[0 * A.C + c] / cnt;
  const double var = A.S[1 * A.C + c] / cnt - mean * mean;
  k.mean = (float)mean;
  k.rstd = (float)rsqrt((var > 0 ? var : 0) + (double)A.eps);
  k.m1 = (float)(A.S[2 * A.C + c] / cnt);
  k.m2 = (float)(A.S[3 * A.C + c] / cnt);
  k.mean_t = (float)(A.S[4 * A.C + c] / cnt);
  k.sdot = (float)(A.S[5 * A.C + c] / cnt);
  k.gamma = A.gamma ? A.gamma[c] : 1.f;
  k.tgamma = A.tgamma ? A.tgamma[c] : 0.f;
  k.tbeta = A.tbeta ? A.tbeta[c] : 0.f;
  return k;
}

template <int MODE>
__global__ void __launch_bounds__(256) bn_reduce_kernel(const __grid_constant__ BnArgs A) {
  __shared__ double red[32];
  const int c = blockIdx.x;
  ChanConst k{};
  if (MODE != BN_STATS) k = chan_const(A, c);
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  for (int n = blockIdx.y; n < A.N; n += gridDim.y) {
    const int64_t off = ((int64_t)n * A.C + c) * A.HW;
    for (int q = threadIdx.x; q < A.HW; q += blockDim.x) {
      const float x = bb::ldf(A.x, off + q, A.dtx);
      if (MODE == BN_STATS) {
        s0 += x;
        s1 += x * x;
      } else {
        const float xh = (x - k.mean) * k.rstd;
        if (MODE == BN_BB) {
          const float gh = A.g[off + q] * k.gamma;
          s0 += gh;
          s1 += gh * xh;
        } else if (MODE == BN_TF) {
          const float t = A.tx[off + q];
          s0 += t;
          s1 += xh * t;
        } else {
          const float dxh = (A.tx[off + q] - k.mean_t - xh * k.sdot) * k.rstd;
          const float g = A.g[off + q], gt = A.gt[off + q];
          const float gh = g * k.gamma, ght = gt * k.gamma + g * k.tgamma;
          s0 += ght;
          s1 += ght * xh + gh * dxh;
          s2 += gt * xh + g * dxh;
          s3 += gt;
        }
      }
    }
  }
  constexpr int slot = MODE == BN_STATS ? 0 : MODE == BN_BB ? 2 : MODE == BN_TF ? 4 : 6;
  double b0 = bb::block_sum<double>((double)s0, red);
  double b1 = bb::block_sum<double>((double)s1, red);
  if (threadIdx.x == 0) {
    atomicAdd(&A.S[(slot + 0) * A.C + c], b0);
    atomicAdd(&A.S[(slot + 1) * A.C + c], b1);
  }
  if (MODE == BN_TB) {
    double b2 = bb::block_sum<double>((double)s2, red);
    double b3 = bb::block_sum<double>((double)s3, red);
    if (threadIdx.x == 0) {
      atomicAdd(&A.S[8 * A.C + c], b2);
      atomicAdd(&A.S[9 * A.C + c], b3);
    }
  }
}

template <int MODE>
__global__ void __launch_bounds__(256) bn_apply_kernel(const __grid_constant__ BnArgs A) {
  const int c = blockIdx.x;
  const ChanConst k = chan_const(A, c);
  const double cnt = (double)A.N * A.HW;
  float mt1 = 0.f, mt2 = 0.f;
  if (MODE == BN_TB) {
    mt1 = (float)(A.S[6 * A.C + c] / cnt);
    mt2 = (float)(A.S[7 * A.C + c] / cnt);
    if (blockIdx.y == 0 && threadIdx.x == 0) {
      if (A.at_gamma) A.at_gamma[c] += (float)A.S[8 * A.C + c];  // parameter slices always accumulate
      if (A.at_beta) A.at_beta[c] += (float)A.S[9 * A.C + c];
    }
  }
  for (int n = blockIdx.y; n < A.N; n += gridDim.y) {
    const int64_t off = ((int64_t)n * A.C + c) * A.HW;
    for (int q = threadIdx.x; q < A.HW; q += blockDim.x) {
      const float xh = (bb::ldf(A.x, off + q, A.dtx) - k.mean) * k.rstd;
      if (MODE == BN_TF) {
        const float dxh = (A.tx[off + q] - k.mean_t - xh * k.sdot) * k.rstd;
        A.ty[off + q] = k.gamma * dxh + k.tgamma * xh + k.tbeta;
      } else if (MODE == BN_BB) {
        const float gh = A.g[off + q] * k.gamma;
        const float v = k.rstd * (gh - k.m1 - xh * k.m2);
        A.dst[off + q] = A.beta ? A.dst[off + q] + v : v;
      } else {
        const float dxh = (A.tx[off + q] - k.mean_t - xh * k.sdot) * k.rstd;
        const float g = A.g[off + q], gt = A.gt[off + q];
        const float gh = g * k.gamma, ght = gt * k.gamma + g * k.tgamma;
        const float u = gh - k.m1 - xh * k.m2;
        const float ut = ght - mt1 - dxh * k.m2 - xh * mt2;
        const float v = -k.rstd * k.rstd * k.sdot * u + k.rstd * ut;
        A.dst[off + q] = A.beta ? A.dst[off + q] + v : v;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// LayerNorm: one warp per row, statistics recomputed per pass; parameter gradients go through
// block-level shared-memory column accumulators and one global atomic per column per block.
// ---------------------------------------------------------------------------------------------------
struct LnArgs {
  const void* x;
  int dtx;
  const float* gamma;
  const float* tx;
  const float* tgamma;
  const float* tbeta;
  const float* g;
  const float* gt;
  float* ty;
  float* dst;
  float* at_gamma;
  float* at_beta;
  int64_t rows;
  int D;
  float eps;
  int beta;
};

// one row per warp and small blocks: the rows are short (hidden size), so parallelism across rows -- not work per
// warp -- is what keeps the SMs busy (profile: 25 blocks of 8 warps x 4 rows took 0.18 ms per LayerNorm)
constexpr int kLnWarps = 4;
constexpr int kLnRowsPerWarp = 1;

template <int MODE>  // BN_BB / BN_TF / BN_TB reuse the enum
__global__ void __launch_bounds__(kLnWarps * 32) ln_kernel(const __grid_constant__ LnArgs A) {
  extern __shared__ float colacc[];  // [2][D] for MODE == TB
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int D = A.D;
  if (MODE == BN_TB) {
    for (int c = threadIdx.x; c < 2 * D; c += blockDim.x) colacc[c] = 0.f;
    __syncthreads();
  }
  const float invD = 1.f / (float)D;
  for (int rr = 0; rr < kLnRowsPerWarp; ++rr) {
    const int64_t row = ((int64_t)blockIdx.x * kLnWarps + warp) * kLnRowsPerWarp + rr;
    if (row >= A.rows) break;
    const int64_t off = row * D;
    float s = 0.f;
    for (int c = lane; c < D; c += 32) s += bb::ldf(A.x, off + c, A.dtx);
    const float mean = bb::warp_sum(s) * invD;
    float v = 0.f;
    for (int c = lane; c < D; c += 32) {
      const float d = bb::ldf(A.x, off + c, A.dtx) - mean;
      v += d * d;
    }
    const float rstd = rsqrtf(bb::warp_sum(v) * invD + A.eps);
    auto XH = [&](int c) { return (bb::ldf(A.x, off + c, A.dtx) - mean) * rstd; };
    auto GAM = [&](int c) { return A.gamma ? A.gamma[c] : 1.f; };
    float mean_t = 0.f, sdot = 0.f;
    if (MODE != BN_BB) {
      float a = 0.f;
      for (int c = lane; c < D; c += 32) a += A.tx[off + c];
      mean_t = bb::warp_sum(a) * invD;
      float b = 0.f;
      for (int c = lane; c < D; c += 32) b += XH(c) * (A.tx[off + c] - mean_t);
      sdot = bb::warp_sum(b) * invD;
    }
    if (MODE == BN_TF) {
      for (int c = lane; c < D; c += 32) {
        const float xh = XH(c);
        const float dxh = (A.tx[off + c] - mean_t - xh * sdot) * rstd;
        A.ty[off + c] = GAM(c) * dxh + (A.tgamma ? A.tgamma[c] * xh : 0.f) + (A.tbeta ? A.tbeta[c] : 0.f);
      }
      continue;
    }
    float a1 = 0.f, a2 = 0.f;
    for (int c = lane; c < D; c += 32) {
      const float gh = A.g[off + c] * GAM(c);
      a1 += gh;
      a2 += gh * XH(c);
    }
    const float m1 = bb::warp_sum(a1) * invD, m2 = bb::warp_sum(a2) * invD;
    if (MODE == BN_BB) {
      for (int c = lane; c < D; c += 32) {
        const float xh = XH(c);
        const float val = rstd * (A.g[off + c] * GAM(c) - m1 - xh * m2);
        A.dst[off + c] = A.beta ? A.dst[off + c] + val : val;
      }
      continue;
    }
    float b1 = 0.f, b2 = 0.f;
    for (int c = lane; c < D; c += 32) {
      const float xh = XH(c);
      const float dxh = (A.tx[off + c] - mean_t - xh * sdot) * rstd;
      const float g = A.g[off + c], gt = A.gt[off + c];
      const float gh = g * GAM(c), ght = gt * GAM(c) + (A.tgamma ? g * A.tgamma[c] : 0.f);
      b1 += ght;
      b2 += ght * xh + gh * dxh;
      atomicAdd(&colacc[c], gt * xh + g * dxh);
      atomicAdd(&colacc[D + c], gt);
    }
    const float mt1 = bb::warp_sum(b1) * invD, mt2 = bb::warp_sum(b2) * invD;
    for (int c = lane; c < D; c += 32) {
      const float xh = XH(c);
      const float dxh = (A.tx[off + c] - mean_t - xh * sdot) * rstd;
      const float g = A.g[off + c], gt = A.gt[off + c];
      const float gh = g * GAM(c), ght = gt * GAM(c) + (A.tgamma ? g * A.tgamma[c] : 0.f);
      const float u = gh - m1 - xh * m2;
      const float ut = ght - mt1 - dxh * m2 - xh * mt2;
      const float val = -rstd * rstd * sdot * u + rstd * ut;
      A.dst[off + c] = A.beta ? A.dst[off + c] + val : val;
    }
  }
  if (MODE == BN_TB) {
    __syncthreads();
    for (int c = threadIdx.x; c < D; c += blockDim.x) {
      if (A.at_gamma) atomicAdd(A.at_gamma + c, colacc[c]);
      if (A.at_beta) atomicAdd(A.at_beta + c, colacc[D + c]);
    }
  }
}

}  // namespace

int bb_launch_batchnorm(const bb_node& nd, int pass, cudaStream_t s) {
  BnArgs A{};
  A.x = nd.base[0]; A.dtx = nd.dt[0];
  A.gamma = reinterpret_cast<const float*>(nd.base[1]);
  A.tx = reinterpret_cast<const float*>(nd.t[0]);
  A.tgamma = (nd.active & 2) ? reinterpret_cast<const float*>(nd.t[1]) : nullptr;
  A.tbeta = (nd.active & 4) ? reinterpret_cast<const float*>(nd.t[2]) : nullptr;
  A.g = reinterpret_cast<const float*>(nd.a[3]);
  A.gt = reinterpret_cast<const float*>(nd.at[3]);
  A.ty = reinterpret_cast<float*>(nd.t[3]);
  A.at_gamma = (nd.active & 2) ? reinterpret_cast<float*>(nd.at[1]) : nullptr;
  A.at_beta = (nd.active & 4) ? reinterpret_cast<float*>(nd.at[2]) : nullptr;
  A.S = reinterpret_cast<double*>(nd.aux[0]);
  A.N = (int)nd.dims[0]; A.C = (int)nd.dims[1]; A.HW = (int)nd.dims[2];
  A.eps = (float)nd.f[0];
  A.beta = nd.beta[0];
  int split = (4 * BB_SM_COUNT + A.C - 1) / A.C;
  if (split > A.N) split = A.N;
  if (split < 1) split = 1;
  const dim3 grid(A.C, split);
  const size_t cbytes = sizeof(double) * A.C;
  if (pass == BB_PASS_BASE_BWD) {
    A.dst = reinterpret_cast<float*>(nd.a[0]);
    BB_CUDA_TRY(cudaMemsetAsync(A.S, 0, 4 * cbytes, s));
    bn_reduce_kernel<BN_STATS><<<grid, 256, 0, s>>>(A);
    bn_reduce_kernel<BN_BB><<<grid, 256, 0, s>>>(A);
    bn_apply_kernel<BN_BB><<<grid, 256, 0, s>>>(A);
    bb_launch_tally += 4;
  } else if (pass == BB_PASS_TAN_FWD) {
    BB_CUDA_TRY(cudaMemsetAsync(A.S + 4 * A.C, 0, 2 * cbytes, s));
    bn_reduce_kernel<BN_TF><<<grid, 256, 0, s>>>(A);
    bn_apply_kernel<BN_TF><<<grid, 256, 0, s>>>(A);
    bb_launch_tally += 3;
  } else {
    A.dst = reinterpret_cast<float*>(nd.at[0]);
    BB_CUDA_TRY(cudaMemsetAsync(A.S + 6 * A.C, 0, 4 * cbytes, s));
    bn_reduce_kernel<BN_TB><<<grid, 256, 0, s>>>(A);
    bn_apply_kernel<BN_TB><<<grid, 256, 0, s>>>(A);
    bb_launch_tally += 3;
  }
  BB_LAUNCH_CHECK();
  return BB_OK;
}

int bb_launch_layernorm(const bb_node& nd, int pass, cudaStream_t s) {
  LnArgs A{};
  A.x = nd.base[0]; A.dtx = nd.dt[0];
  A.gamma = reinterpret_cast<const float*>(nd.base[1]);
  A.tx = reinterpret_cast<const float*>(nd.t[0]);
  A.tgamma = (nd.active & 2) ? reinterpret_cast<const float*>(nd.t[1]) : nullptr;
  A.tbeta = (nd.active & 4) ? reinterpret_cast<const float*>(nd.t[2]) : nullptr;
  A.g = reinterpret_cast<const float*>(nd.a[3]);
  A.gt = reinterpret_cast<const float*>(nd.at[3]);
  A.ty = reinterpret_cast<float*>(nd.t[3]);
  A.at_gamma = (nd.active & 2) ? reinterpret_cast<float*>(nd.at[1]) : nullptr;
  A.at_beta = (nd.active & 4) ? reinterpret_cast<float*>(nd.at[2]) : nullptr;
  A.rows = nd.dims[0]; A.D = (int)nd.dims[1];
  A.eps = (float)nd.f[0];
  A.beta = nd.beta[0];
  if (A.rows <= 0) return BB_OK;
  const int rows_per_block = kLnWarps * kLnRowsPerWarp;
  const unsigned grid = (unsigned)((A.rows + rows_per_block - 1) / rows_per_block);
  if (pass == BB_PASS_BASE_BWD) {
    A.dst = reinterpret_cast<float*>(nd.a[0]);
    ln_kernel<BN_BB><<<grid, kLnWarps * 32, 0, s>>>(A);
  } else if (pass == BB_PASS_TAN_FWD) {
    ln_kernel<BN_TF><<<grid, kLnWarps * 32, 0, s>>>(A);
  } else {
    A.dst = reinterpret_cast<float*>(nd.at[0]);
    const size_t smem = sizeof(float) * 2 * A.D;
    if (smem > 48 * 1024) return BB_ERR_UNSUPPORTED;
    ln_kernel<BN_TB><<<grid, kLnWarps * 32, smem, s>>>(A);
  }
  bb_launch_tally += 1;
  BB_LAUNCH_CHECK();
  return BB_OK;
}
