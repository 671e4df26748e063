// Fused second-order rules of a DATA-INPUT convolution block
//
//     x (data) --conv3x3(W,b)--> y --BatchNorm2d(batch stats; gamma,beta)--> z --[ReLU]--> MaxPool2d(2)--> q
//
// (reference examples/implicit_maml/models.py:9-24; the first block of the 4-conv backbone is 2/3 of the K-loop's
// algorithmic bytes).  Because x is data, the tangent t_y = X t_W + t_b (X = im2col(x), P x CKK) is
// LINEAR in the direction, so every full-resolution reduction of the BatchNorm rules collapses to a product with a
// per-call constant:
//
//     s[k]    = sum_p X[p,k]                 sx[o]   = sum_p xhat[p,o]
//     G[o,k]  = sum_p xhat[p,o] X[p,k]       S2[k,l] = sum_p X[p,k] X[p,l]
//
//   TF   mean_t = (t_W.s)/P + t_b            sdot = (t_W.G + t_b sx - mean_t sx)/P
//        t_q[w,o] = mask * (gamma*rstd*(t_y[p*] + ... ) ...)   only at the arg-max pixel p*(w,o) of each window
//   TB   at_y = sparse + d0 + xhat d1 + t_y d2   (norm.cu's rule expanded; sparse lives on the arg-max pixels)
//        at_W = sum_p at_y X = rstd gamma GW + (rstd t_gamma - rstd^2 sdot gamma) Gd + d0 s + d1 G + d2 (t_W S2 + t_b s)
//        with GW[o,k] = sum_w mask at_q[w,o] X[p*(w,o),k]  -- the only per-iteration pass, at POOLED resolution.
//
// So neither y-sized tangents nor y-sized adjoints exist: per iteration the block reads x and streams pooled-size
// arrays (a quarter of y).  All arithmetic fp32 (x / y / q are read in the dtype the forward recorded).
// Verified against the composition of the three member rules (oracle/plan_interp.py) and autograd's double backward.
#include <cuda_bf16.h>
#include <stdlib.h>

#include "../../include/betty_b200.h"
#include "bb_common.cuh"
#include "plan.h"

namespace {

constexpr int KP = 28;          // padded CKK (<= 27 used)
constexpr int NT = 256;         // threads per CTA
constexpr int MAXO = 64;
constexpr int NSUM = 4;         // per-channel scalar sums carried next to GW

// pooled arrays (xhat*, dxhat*, mask a_q) are fp32 for fp32 graphs and bf16 for reduced-precision graphs (they are
// streamed once per iteration: 13 -> 7 bytes per pooled element)
template <typename PT> __device__ __forceinline__ float ldp(const PT* p, int64_t i);
template <> __device__ __forceinline__ float ldp<float>(const float* p, int64_t i) { return p[i]; }
template <> __device__ __forceinline__ float ldp<__nv_bfloat16>(const __nv_bfloat16* p, int64_t i) { return __bfloat162float(p[i]); }
template <typename PT> __device__ __forceinline__ void stp(PT* p, int64_t i, float v);
template <> __device__ __forceinline__ void stp<float>(float* p, int64_t i, float v) { p[i] = v; }
template <> __device__ __forceinline__ void stp<__nv_bfloat16>(__nv_bfloat16* p, int64_t i, float v) { p[i] = __float2bfloat16(v); }

struct CbGeom {
  int N, C, H, W, O, HO, WO, ph, pw, HP, WP, relu;
  int ps;         // bytes per pooled-array element (4 or 2)
  int R;          // window rows per tile
  int tiles_per_img, ntiles;
  int xrows, xpitch;   // input tile rows / pitch in shared memory
  int wpitch;          // pooled tile pitch in shared memory
};

// workspace layout (floats unless noted); offsets in bytes, filled by cb_layout()
struct CbWs {
  double* dsum;     // [2*O] sum y, sum y^2  + [KP + O + O*KP + KP*KP] double accumulators of the Gram pass
  float* mean;      // [O]
  float* rstd;      // [O]
  float* s;         // [KP]
  float* sx;        // [O]
  float* G;         // [O*KP]
  float* S2;        // [KP*KP]
  float* Gd;        // [O*KP]   gather-wgrad of the masked base adjoint
  float* Sa;        // [O]      sum mask a_q
  float* Saxh;      // [O]      sum mask a_q xhat_sel
  float* mean_t;    // [O]      written by the tangent-forward pass
  float* sdot;      // [O]
  float* part;      // [grid][O][KP + NSUM] per-CTA partial sums of the reduce kernels
  unsigned char* sel;   // [N*HP*WP*O] NHWC: code (dy*2+dx) | mask << 2
  void* xh;         // [N*HP*WP*O] NHWC xhat at the arg-max pixel          (fp32 or bf16, CbGeom::ps)
  void* dxh;        // [N*HP*WP*O] NHWC dxhat at the arg-max pixel (per iteration)
  void* aqm;        // [N*HP*WP*O] NHWC mask * a_q
  size_t bytes;
};

constexpr int GRID_MAX = 4 * BB_SM_COUNT;

__host__ __device__ inline size_t up(size_t x) { return (x + 255) & ~(size_t)255; }

CbWs cb_layout(void* base, const CbGeom& g) {
  CbWs w{};
  size_t at = 0;
  uint8_t* b = reinterpret_cast<uint8_t*>(base);
  auto take = [&](size_t bytes) { size_t o = at; at = up(at + bytes); return b ? b + o : nullptr; };
  const size_t O = g.O, pooled = (size_t)g.N * g.HP * g.WP * g.O;
  w.dsum = reinterpret_cast<double*>(take(sizeof(double) * (2 * O + KP + O + O * KP + KP * KP)));
  w.mean = reinterpret_cast<float*>(take(4 * O));
  w.rstd = reinterpret_cast<float*>(take(4 * O));
  w.s = reinterpret_cast<float*>(take(4 * KP));
  w.sx = reinterpret_cast<float*>(take(4 * O));
  w.G = reinterpret_cast<float*>(take(4 * O * KP));
  w.S2 = reinterpret_cast<float*>(take(4 * KP * KP));
  w.Gd = reinterpret_cast<float*>(take(4 * O * KP));
  w.Sa = reinterpret_cast<float*>(take(4 * O));
  w.Saxh = reinterpret_cast<float*>(take(4 * O));
  w.mean_t = reinterpret_cast<float*>(take(4 * O));
  w.sdot = reinterpret_cast<float*>(take(4 * O));
  w.part = reinterpret_cast<float*>(take(4 * (size_t)GRID_MAX * O * (KP + NSUM)));
  w.sel = reinterpret_cast<unsigned char*>(take(pooled));
  const size_t ps = g.ps == 2 ? 2 : 4;
  w.xh = take(ps * pooled);
  w.dxh = take(ps * pooled);
  w.aqm = take(ps * pooled);
  w.bytes = at;
  return w;
}

struct CbArgs {
  CbGeom g;
  CbWs w;
  const void* x; int dtx;
  const void* y; int dty;
  const void* q; int dtq;
  const int64_t* idx;
  const float* gamma;
  float eps;
  // parameter tangents (direction arena slices) and adjoint-tangent slices (H.d arena); null when absent
  const float *t_W, *t_b, *t_gamma, *t_beta;
  float *at_W, *at_b, *at_gamma, *at_beta;
  // pooled output buffers of the plan (fp32, NCHW)
  float* t_q;
  __nv_bfloat16* tq_nhwc;     // when the consumer is a fused block: t_q as bf16 NHWC (its TMA operand), t_q unused
  const float* a_q;
  const float* at_q;
  // when the consumer is a fused inner block it hands a_q / at_q over as bf16 padded NHWC [N][HP+2][WP+2][64]
  const __nv_bfloat16* aq_nhwc;
  const __nv_bfloat16* atq_nhwc;
  int nparts;     // CTAs that wrote partials
};

CbGeom make_geom(const bb_node& nd) {
  CbGeom g{};
  g.N = (int)nd.dims[0]; g.C = (int)nd.dims[1]; g.H = (int)nd.dims[2]; g.W = (int)nd.dims[3]; g.O = (int)nd.dims[4];
  g.HO = (int)nd.dims[7]; g.WO = (int)nd.dims[8]; g.ph = (int)nd.dims[11]; g.pw = (int)nd.dims[12];
  g.HP = (int)nd.dims[13]; g.WP = (int)nd.dims[14]; g.relu = (int)nd.dims[15];
  g.ps = (nd.kind & 1) ? 2 : 4;
  int R = 96 / (g.WP > 0 ? g.WP : 1);     // ~96 windows x O channels per tile (two pooled rows of the 84x84 layer)
  if (R < 1) R = 1;
  if (R > g.HP) R = g.HP;
  g.R = R;
  g.tiles_per_img = (g.HP + R - 1) / R;
  g.ntiles = g.N * g.tiles_per_img;
  g.xrows = 2 * R + 2;
  g.xpitch = (g.WO + 2) | 1;              // odd pitch: the four candidate pixels of a window fall into distinct banks
  g.wpitch = (R * g.WP) | 1;
  return g;
}

// ---------------------------------------------------------------------------------------------------------------
// base statistics of y (once per call)
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) cb_stats_kernel(const CbArgs A) {
  // grid (O, chunks): each block sums a slice of the channel's N*HO*WO values
  __shared__ double red[32];
  const int o = blockIdx.x, HW = A.g.HO * A.g.WO;
  const int64_t per = (int64_t)A.g.N * HW;
  double s0 = 0, s1 = 0;
  for (int64_t i = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; i < per; i += (int64_t)gridDim.y * blockDim.x) {
    const int64_t n = i / HW, r = i - n * HW;
    const float v = bb::ldf(A.y, (n * A.g.O + o) * HW + r, A.dty);
    s0 += v;
    s1 += (double)v * v;
  }
  s0 = bb::block_sum<double>(s0, red);
  s1 = bb::block_sum<double>(s1, red);
  if (threadIdx.x == 0) {
    atomicAdd(&A.w.dsum[o], s0);
    atomicAdd(&A.w.dsum[A.g.O + o], s1);
  }
}

__global__ void cb_stats_finish_kernel(const CbArgs A) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= A.g.O) return;
  const double cnt = (double)A.g.N * A.g.HO * A.g.WO;
  const double mean = A.w.dsum[o] / cnt;
  const double var = A.w.dsum[A.g.O + o] / cnt - mean * mean;
  A.w.mean[o] = (float)mean;
  A.w.rstd[o] = (float)rsqrt((var > 0 ? var : 0) + (double)A.eps);
}

// ---------------------------------------------------------------------------------------------------------------
// shared tile helpers
// ---------------------------------------------------------------------------------------------------------------
// input rows needed by window rows [hp0, hp0+R): conv-output rows 2*hp0 .. 2*(hp0+R)-1, taps -ph .. 2-ph
template <int C>
__device__ __forceinline__ void load_x_tile(const CbArgs& A, int n, int hp0, float* xs) {
  const CbGeom& g = A.g;
  const int rows = g.xrows, pitch = g.xpitch, cols = g.WO + 2;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  // warp per (channel, row), lanes along the row: no divisions, coalesced reads
  for (int cr = wid; cr < C * rows; cr += NT / 32) {
    const int c = cr / rows, r = cr - c * rows;
    const int iy = 2 * hp0 - g.ph + r;
    float* d = xs + cr * pitch;
    if (iy < 0 || iy >= g.H) {
      for (int col = lane; col < cols; col += 32) d[col] = 0.f;
      continue;
    }
    const int64_t rowbase = (((int64_t)n * C + c) * g.H + iy) * g.W - g.pw;
    for (int col = lane; col < cols; col += 32) {
      const int ix = col - g.pw;
      d[col] = (ix >= 0 && ix < g.W) ? bb::ldf(A.x, rowbase + col, A.dtx) : 0.f;
    }
  }
}

// Bulk copy of a contiguous global range into shared memory with cp.async (no register staging: every thread has
// several independent 16-byte requests in flight, so the pooled arrays of a whole tile stream in at DRAM latency
// once instead of once per window -- the first version of these kernels ran at 10 % of DRAM throughput on exactly
// that dependency, profiles/r02_convblock_v0_ncu.md).  Falls back to plain loads for unaligned ranges.
__device__ __forceinline__ void tile_copy_async(void* dst_smem, const void* src, int nbytes) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst_smem);
  const uintptr_t a = reinterpret_cast<uintptr_t>(src);
  if (((a | d | (uintptr_t)nbytes) & 15) == 0) {
    for (int i = threadIdx.x * 16; i < nbytes; i += NT * 16)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d + i), "l"(a + i) : "memory");
  } else if (((a | d | (uintptr_t)nbytes) & 3) == 0) {
    for (int i = threadIdx.x * 4; i < nbytes; i += NT * 4)
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d + i), "l"(a + i) : "memory");
  } else {
    const unsigned char* sp = reinterpret_cast<const unsigned char*>(src);
    unsigned char* dp = reinterpret_cast<unsigned char*>(dst_smem);
    for (int i = threadIdx.x; i < nbytes; i += NT) dp[i] = sp[i];
  }
}
__device__ __forceinline__ void tile_copy_wait() {
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// dot of the 3x3xC patch whose top-left tap sits at xs[base] with per-lane weights tw[C*9]
template <int C>
__device__ __forceinline__ float patch_dot(const float* xs, int base, int plane, int pitch, const float* tw) {
  float acc = 0.f;
#pragma unroll
  for (int c = 0; c < C; ++c)
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) acc = fmaf(xs[base + c * plane + i * pitch + j], tw[(c * 3 + i) * 3 + j], acc);
  return acc;
}

// ---------------------------------------------------------------------------------------------------------------
// base-backward preparation: arg-max codes, ReLU mask, xhat at the arg-max pixel, masked base adjoint (NHWC)
// ---------------------------------------------------------------------------------------------------------------
template <typename PT>
__global__ void __launch_bounds__(256) cb_prep_kernel(const CbArgs A) {
  const CbGeom& g = A.g;
  const int64_t total = (int64_t)g.N * g.HP * g.WP * g.O;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    // i walks the NCHW pooled tensor (coalesced reads of idx / q / a_q); the NHWC writes are scattered, once per call
    const int wp = (int)(i % g.WP);
    int64_t r = i / g.WP;
    const int hp = (int)(r % g.HP);
    r /= g.HP;
    const int o = (int)(r % g.O);
    const int n = (int)(r / g.O);
    const int64_t id = A.idx[i];
    const int iy = (int)(id / g.WO), ix = (int)(id - (int64_t)iy * g.WO);
    const int dy = iy - 2 * hp, dx = ix - 2 * wp;
    const bool m = g.relu ? (bb::ldf(A.q, i, A.dtq) > 0.f) : true;
    const float yv = bb::ldf(A.y, ((int64_t)n * g.O + o) * g.HO * g.WO + id, A.dty);
    const int64_t pi = (((int64_t)n * g.HP + hp) * g.WP + wp) * g.O + o;
    A.w.sel[pi] = (unsigned char)((dy & 1) * 2 + (dx & 1) + (m ? 4 : 0));
    stp<PT>(reinterpret_cast<PT*>(A.w.xh), pi, (yv - A.w.mean[o]) * A.w.rstd[o]);
    const float aq = A.aq_nhwc ? __bfloat162float(A.aq_nhwc[((((int64_t)n * (g.HP + 2) + hp + 1) * (g.WP + 2)) + wp + 1) * 64 + o])
                               : A.a_q[i];
    stp<PT>(reinterpret_cast<PT*>(A.w.aqm), pi, m ? aq : 0.f);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Gram pass (once per call): s, sx, G, S2 over every pixel of y
// ---------------------------------------------------------------------------------------------------------------
template <int C>
__global__ void __launch_bounds__(NT) cb_gram_kernel(const CbArgs A) {
  // tile = (image n, band of 2R conv-output rows) -- the same input tile as the K-loop kernels.
  //   warps 0 .. NW-3: lane <-> channel o, warp <-> (channel group, pixel group); per-lane G[o,:] / sx[o] accumulators,
  //                    the pixel's patch is a shared-memory broadcast
  //   warps NW-2, NW-1: lane k < CKK owns row k of S2 and s[k] (even / odd pixels)
  extern __shared__ float sm[];
  const CbGeom& g = A.g;
  constexpr int CKK = C * 9;
  constexpr int NW = NT / 32;
  const int band = 2 * g.R;
  const int bands = (g.HO + band - 1) / band;
  float* xs = sm;
  const int plane = g.xrows * g.xpitch;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int cgs = (g.O + 31) / 32, pgs = (NW - 2) / cgs;     // channel groups x pixel groups among the G warps
  const bool s2warp = wid >= NW - 2;
  const int cg = wid % cgs, pg = wid / cgs;
  const int o = cg * 32 + lane;
  const bool gwarp = !s2warp && pg < pgs;
  const bool och = gwarp && o < g.O;
  float acc[CKK];
#pragma unroll
  for (int k = 0; k < CKK; ++k) acc[k] = 0.f;
  float a1 = 0.f;
  const float mean = och ? A.w.mean[o] : 0.f, rstd = och ? A.w.rstd[o] : 0.f;
  const int HW = g.HO * g.WO;
  for (int tile = blockIdx.x; tile < g.N * bands; tile += gridDim.x) {
    const int n = tile / bands, b = tile - n * bands;
    __syncthreads();
    load_x_tile<C>(A, n, b * g.R, xs);
    __syncthreads();
    const int row0 = b * band;
    const int npix = min(band, g.HO - row0) * g.WO;
    if (s2warp) {
      if (lane < CKK) {
        for (int p = wid - (NW - 2); p < npix; p += 2) {
          const int oy = p / g.WO, ox = p - oy * g.WO;
          const int base = oy * g.xpitch + ox;
          const int c0 = lane / 9, r0 = lane - c0 * 9, i0 = r0 / 3, j0 = r0 - i0 * 3;
          const float mine = xs[base + c0 * plane + i0 * g.xpitch + j0];
          a1 += mine;
#pragma unroll
          for (int c = 0; c < C; ++c)
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
              for (int j = 0; j < 3; ++j)
                acc[(c * 3 + i) * 3 + j] = fmaf(mine, xs[base + c * plane + i * g.xpitch + j], acc[(c * 3 + i) * 3 + j]);
        }
      }
    } else if (och) {
      const int64_t ybase = ((int64_t)n * g.O + o) * HW + (int64_t)row0 * g.WO;
      for (int p = pg; p < npix; p += pgs) {
        const int oy = p / g.WO, ox = p - oy * g.WO;
        const int base = oy * g.xpitch + ox;
        const float xh = (bb::ldf(A.y, ybase + p, A.dty) - mean) * rstd;
        a1 += xh;
#pragma unroll
        for (int c = 0; c < C; ++c)
#pragma unroll
          for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j)
              acc[(c * 3 + i) * 3 + j] = fmaf(xh, xs[base + c * plane + i * g.xpitch + j], acc[(c * 3 + i) * 3 + j]);
      }
    }
  }
  // one double atomic per accumulator and thread (once per call)
  double* D = A.w.dsum + 2 * g.O;                         // [KP] s | [O] sx | [O*KP] G | [KP*KP] S2
  if (s2warp) {
    if (lane < CKK) {
      atomicAdd(&D[lane], (double)a1);
#pragma unroll
      for (int k = 0; k < CKK; ++k) atomicAdd(&D[KP + g.O + g.O * KP + lane * KP + k], (double)acc[k]);
    }
  } else if (och) {
    atomicAdd(&D[KP + o], (double)a1);
#pragma unroll
    for (int k = 0; k < CKK; ++k) atomicAdd(&D[KP + g.O + o * KP + k], (double)acc[k]);
  }
}

__global__ void cb_gram_finish_kernel(const CbArgs A) {
  const CbGeom& g = A.g;
  const double* D = A.w.dsum + 2 * g.O;
  const int n = KP + g.O + g.O * KP + KP * KP;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float v = (float)D[i];
    if (i < KP) A.w.s[i] = v;
    else if (i < KP + g.O) A.w.sx[i - KP] = v;
    else if (i < KP + g.O + g.O * KP) A.w.G[i - KP - g.O] = v;
    else A.w.S2[i - KP - g.O - g.O * KP] = v;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// tangent forward: t_q and dxhat at the arg-max pixels
//   lane <-> channel (its 27 direction weights live in registers), warp <-> (channel group, window group); the four
//   candidate pixels of a window are four distinct shared-memory words -> conflict-free broadcast reads.  Window
//   indices advance without divisions; the pooled NHWC arrays (code, xhat) of the next window are fetched while the
//   current one is being computed.
// ---------------------------------------------------------------------------------------------------------------
template <int C, typename PT>
__global__ void __launch_bounds__(NT, 4) cb_tf_kernel(const CbArgs A) {
  extern __shared__ float sm[];
  const CbGeom g = A.g;
  constexpr int CKK = C * 9;
  const int plane = g.xrows * g.xpitch;
  float* xs = sm;
  float* outs = sm + C * plane;                    // [O][wpitch] pooled tile, transposed for the NCHW store
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int cgs = (g.O + 31) / 32, wgs = (NT / 32) / cgs;   // channel groups x window groups
  const int cg = wid % cgs, wg = wid / cgs;
  const int o = cg * 32 + lane;
  const bool och = o < g.O && wg < wgs;
  const double invP = 1.0 / ((double)g.N * g.HO * g.WO);
  const unsigned char* __restrict__ sel = A.w.sel;
  const PT* __restrict__ xhp = reinterpret_cast<const PT*>(A.w.xh);
  PT* __restrict__ dxhp = reinterpret_cast<PT*>(A.w.dxh);
  // per-lane channel constants
  float tw[CKK];
  float mean_t = 0.f, sdot = 0.f, rstd = 0.f, gam = 1.f, tgam = 0.f, tbeta = 0.f, tb = 0.f;
  if (och) {
    double d0 = 0, d1 = 0;
#pragma unroll
    for (int k = 0; k < CKK; ++k) {
      tw[k] = A.t_W[o * CKK + k];
      d0 += (double)tw[k] * A.w.s[k];
      d1 += (double)tw[k] * A.w.G[o * KP + k];
    }
    tb = A.t_b ? A.t_b[o] : 0.f;
    const double mt = d0 * invP + tb;
    const double sx = A.w.sx[o];
    mean_t = (float)mt;
    sdot = (float)((d1 + tb * sx - mt * sx) * invP);
    rstd = A.w.rstd[o];
    gam = A.gamma ? A.gamma[o] : 1.f;
    tgam = A.t_gamma ? A.t_gamma[o] : 0.f;
    tbeta = A.t_beta ? A.t_beta[o] : 0.f;
    if (blockIdx.x == 0 && wg == 0) {
      A.w.mean_t[o] = mean_t;
      A.w.sdot[o] = sdot;
    }
  } else {
#pragma unroll
    for (int k = 0; k < CKK; ++k) tw[k] = 0.f;
  }
  // t_q = mask * (c_y * t_y + c_x * xhat + c_0)
  const float c_y = gam * rstd, c_x = tgam - gam * rstd * sdot, c_0 = tbeta + gam * rstd * (tb - mean_t);
  const float e_0 = (tb - mean_t) * rstd, e_x = -sdot * rstd;          // dxhat = rstd*t_y + e_x*xhat + e_0
  const int pitch = g.xpitch, WPc = g.WP, Oc = g.O;
  int roff[C * 3];                                                      // (channel, tap row) offsets inside the x tile
#pragma unroll
  for (int r = 0; r < C * 3; ++r) roff[r] = (r / 3) * plane + (r % 3) * pitch;
  const int tcap = g.R * g.WP * g.O;                 // pooled elements of a full tile
  PT* xh_s = reinterpret_cast<PT*>(outs + (A.tq_nhwc ? 0 : g.O * g.wpitch));   // [nw][O] xhat* of the tile (no NCHW
                                                                                // staging tile in NHWC mode)
  unsigned char* sel_s = reinterpret_cast<unsigned char*>(xh_s + tcap);
  for (int tile = blockIdx.x; tile < g.ntiles; tile += gridDim.x) {
    const int n = tile / g.tiles_per_img, tr = tile - n * g.tiles_per_img;
    const int hp0 = tr * g.R, rows = min(g.R, g.HP - hp0), nw = rows * g.WP;
    const int64_t t0 = ((int64_t)n * g.HP + hp0) * g.WP * g.O;             // NHWC index of (window 0, channel 0)
    __syncthreads();
    tile_copy_async(xh_s, xhp + t0, nw * g.O * (int)sizeof(PT));
    tile_copy_async(sel_s, sel + t0, nw * g.O);
    load_x_tile<C>(A, n, hp0, xs);
    tile_copy_wait();
    __syncthreads();
    if (och) {
      // strength-reduced indices: everything below advances by additions (the first version spent ~50 integer
      // multiply-adds per window on address arithmetic, profiles/r02_convblock_v1_ncu.md)
      int wr = 0, wc = wg;
      while (wc >= WPc) { wc -= WPc; ++wr; }
      int rowbase = 2 * wr * pitch + 2 * wc;                    // xs offset of the window's top-left candidate pixel
      int sidx = wg * Oc + o;                                   // index into the staged pooled arrays
      PT* dxo = dxhp + t0 + sidx;                               // global dxhat* slot
      __nv_bfloat16* tqo = A.tq_nhwc
          ? A.tq_nhwc + ((((int64_t)n * (g.HP + 2) + hp0 + wr + 1) * (WPc + 2)) + wc + 1) * 64 + o : nullptr;
      for (int wl = wg; wl < nw; wl += wgs) {
        const unsigned code_c = sel_s[sidx];
        const float xh_c = ldp<PT>(xh_s, sidx);
        const float* px = xs + rowbase + ((code_c & 2) ? pitch : 0) + (code_c & 1);
        float ty = 0.f;
#pragma unroll
        for (int r = 0; r < C * 3; ++r) {
          const float* q = px + roff[r];
          ty = fmaf(q[0], tw[3 * r], ty);
          ty = fmaf(q[1], tw[3 * r + 1], ty);
          ty = fmaf(q[2], tw[3 * r + 2], ty);
        }
        stp<PT>(dxo, 0, fmaf(rstd, ty, fmaf(e_x, xh_c, e_0)));
        const float tq = (code_c & 4) ? fmaf(c_y, ty, fmaf(c_x, xh_c, c_0)) : 0.f;
        if (tqo)
          *tqo = __float2bfloat16(tq);
        else
          outs[o * g.wpitch + wl] = tq;
        sidx += wgs * Oc;
        dxo += wgs * Oc;
        wc += wgs;
        rowbase += 2 * wgs;
        if (tqo) tqo += wgs * 64;
        while (wc >= WPc) {
          wc -= WPc; ++wr;
          rowbase += 2 * pitch - 2 * WPc;
          if (tqo) tqo += 2 * 64;                                // skip the right border of this row and the left of the next
        }
      }
    }
    if (!A.tq_nhwc) {
      __syncthreads();
      // NCHW store: for each channel the tile's windows are `nw` consecutive floats; warp per channel, lanes along them
      float* dst = A.t_q + (((int64_t)n * g.O) * g.HP + hp0) * g.WP;
      for (int oo = wid; oo < g.O; oo += NT / 32) {
        float* d = dst + (int64_t)oo * g.HP * g.WP;
        const float* sp = outs + oo * g.wpitch;
        for (int wl = lane; wl < nw; wl += 32) d[wl] = sp[wl];
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// pooled-resolution reduction (tangent backward; BASE = the same sums for the base adjoint, once per call)
//   part[cta][o][k < KP]  = sum_w v(w,o) X[p*(w,o), k]         v = mask*at_q   (BASE: mask*a_q)
//   part[cta][o][KP + 0]  = sum_w v
//   part[cta][o][KP + 1]  = sum_w v * xhat_sel
//   part[cta][o][KP + 2]  = sum_w mask*a_q * dxhat_sel         (TB only)
// ---------------------------------------------------------------------------------------------------------------
template <int C, bool BASE, typename PT>
__global__ void __launch_bounds__(NT) cb_reduce_kernel(const CbArgs A) {
  extern __shared__ float sm[];
  const CbGeom g = A.g;
  constexpr int CKK = C * 9;
  const int plane = g.xrows * g.xpitch;
  float* xs = sm;
  float* ins = sm + C * plane;                     // [O][wpitch] at_q tile (NCHW load, transposed reads)
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int cgs = (g.O + 31) / 32, wgs = (NT / 32) / cgs;
  const int cg = wid % cgs, wg = wid / cgs;
  const int o = cg * 32 + lane;
  const bool och = o < g.O && wg < wgs;
  const unsigned char* __restrict__ sel = A.w.sel;
  const PT* __restrict__ xhp = reinterpret_cast<const PT*>(A.w.xh);
  const PT* __restrict__ dxhp = reinterpret_cast<const PT*>(A.w.dxh);
  const PT* __restrict__ aqm = reinterpret_cast<const PT*>(A.w.aqm);
  float gw[CKK];
#pragma unroll
  for (int k = 0; k < CKK; ++k) gw[k] = 0.f;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f;
  const int pitch = g.xpitch, WPc = g.WP, Oc = g.O;
  int roff[C * 3];
#pragma unroll
  for (int r = 0; r < C * 3; ++r) roff[r] = (r / 3) * plane + (r % 3) * pitch;
  const int tcap = g.R * g.WP * g.O;
  // pooled arrays of the tile; in NHWC mode the transposed NCHW staging tile `ins` is not needed and its place is taken
  // by the adjoint-tangent rows themselves ([rows][WP][64] bf16, copied row by row out of the padded array)
  const bool nhwc_mode = A.atq_nhwc != nullptr || A.tq_nhwc != nullptr;     // (the launcher sizes shared memory by this)
  const bool nhwc_in = !BASE && A.atq_nhwc != nullptr;
  PT* xh_s = reinterpret_cast<PT*>(ins + (nhwc_mode ? 0 : g.O * g.wpitch));
  PT* aq_s = xh_s + tcap;
  PT* dx_s = aq_s + tcap;
  unsigned char* sel_s = reinterpret_cast<unsigned char*>(dx_s + tcap);
  __nv_bfloat16* at_s = reinterpret_cast<__nv_bfloat16*>(sel_s + ((tcap + 15) & ~15));
  for (int tile = blockIdx.x; tile < g.ntiles; tile += gridDim.x) {
    const int n = tile / g.tiles_per_img, tr = tile - n * g.tiles_per_img;
    const int hp0 = tr * g.R, rows = min(g.R, g.HP - hp0), nw = rows * g.WP;
    const int64_t t0 = ((int64_t)n * g.HP + hp0) * g.WP * g.O;
    __syncthreads();
    tile_copy_async(xh_s, xhp + t0, nw * g.O * (int)sizeof(PT));
    tile_copy_async(aq_s, aqm + t0, nw * g.O * (int)sizeof(PT));
    if (!BASE) tile_copy_async(dx_s, dxhp + t0, nw * g.O * (int)sizeof(PT));
    tile_copy_async(sel_s, sel + t0, nw * g.O);
    if (nhwc_in)
      for (int rr = 0; rr < rows; ++rr)
        tile_copy_async(at_s + rr * WPc * 64, A.atq_nhwc + ((((int64_t)n * (g.HP + 2) + hp0 + rr + 1) * (WPc + 2)) + 1) * 64,
                        WPc * 64 * 2);
    load_x_tile<C>(A, n, hp0, xs);
    if (!BASE && !A.atq_nhwc) {
      const float* src = A.at_q + (((int64_t)n * g.O) * g.HP + hp0) * g.WP;
      for (int oo = wid; oo < g.O; oo += NT / 32) {
        const float* sp = src + (int64_t)oo * g.HP * g.WP;
        float* d = ins + oo * g.wpitch;
        for (int wl = lane; wl < nw; wl += 32) d[wl] = sp[wl];
      }
    }
    tile_copy_wait();
    __syncthreads();
    if (och) {
      int wr = 0, wc = wg;
      while (wc >= WPc) { wc -= WPc; ++wr; }
      int rowbase = 2 * wr * pitch + 2 * wc;
      int si = wg * Oc + o;
      for (int wl = wg; wl < nw; wl += wgs) {
        const unsigned code_c = sel_s[si];
        const float xh_c = ldp<PT>(xh_s, si), aq_c = ldp<PT>(aq_s, si);
        float v;
        if (BASE) {
          v = aq_c;
        } else {
          const float a = nhwc_in ? __bfloat162float(at_s[wl * 64 + o]) : ins[o * g.wpitch + wl];
          v = (code_c & 4) ? a : 0.f;
          s2 = fmaf(aq_c, ldp<PT>(dx_s, si), s2);
        }
        s0 += v;
        s1 = fmaf(v, xh_c, s1);
        const float* px = xs + rowbase + ((code_c & 2) ? pitch : 0) + (code_c & 1);
#pragma unroll
        for (int r = 0; r < C * 3; ++r) {
          const float* q = px + roff[r];
          gw[3 * r] = fmaf(v, q[0], gw[3 * r]);
          gw[3 * r + 1] = fmaf(v, q[1], gw[3 * r + 1]);
          gw[3 * r + 2] = fmaf(v, q[2], gw[3 * r + 2]);
        }
        si += wgs * Oc;
        wc += wgs;
        rowbase += 2 * wgs;
        while (wc >= WPc) {
          wc -= WPc; ++wr;
          rowbase += 2 * pitch - 2 * WPc;
        }
      }
    }
  }
  // cross-window-group reduction through shared memory, then one partial row per (cta, channel)
  __syncthreads();
  float* red = sm;      // [wgs][O][KP + NSUM] -- reuses the tile memory
  const int stride = KP + NSUM;
  if (och) {
    float* r = red + ((size_t)wg * g.O + o) * stride;
#pragma unroll
    for (int k = 0; k < CKK; ++k) r[k] = gw[k];
    for (int k = CKK; k < KP; ++k) r[k] = 0.f;
    r[KP + 0] = s0; r[KP + 1] = s1; r[KP + 2] = s2; r[KP + 3] = 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < g.O * stride; i += NT) {
    float acc = 0.f;
    for (int w = 0; w < wgs; ++w) acc += red[(size_t)w * g.O * stride + i];
    A.w.part[(size_t)blockIdx.x * g.O * stride + i] = acc;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// finish: combine the per-CTA partials (fixed order: deterministic) and write the parameter slices of H.d
// ---------------------------------------------------------------------------------------------------------------
template <bool BASE>
__global__ void __launch_bounds__(64) cb_finish_kernel(const CbArgs A, int CKK) {
  // one block per channel; thread k < KP + NSUM owns one column of the partial rows
  const CbGeom& g = A.g;
  const int o = blockIdx.x, k = threadIdx.x;
  const int stride = KP + NSUM;
  __shared__ float col[KP + NSUM];
  __shared__ float tyx[KP];
  if (k < stride) {
    double acc = 0.0;
    for (int c = 0; c < A.nparts; ++c) acc += A.w.part[((size_t)c * g.O + o) * stride + k];
    col[k] = (float)acc;
  }
  __syncthreads();
  if (BASE) {
    if (k < KP) A.w.Gd[o * KP + k] = col[k];
    if (k == 0) { A.w.Sa[o] = col[KP + 0]; A.w.Saxh[o] = col[KP + 1]; }
    return;
  }
  const double P = (double)g.N * g.HO * g.WO;
  const float rstd = A.w.rstd[o], sdot = A.w.sdot[o], mean_t = A.w.mean_t[o];
  const float gam = A.gamma ? A.gamma[o] : 1.f, tgam = A.t_gamma ? A.t_gamma[o] : 0.f;
  const float tb = A.t_b ? A.t_b[o] : 0.f;
  const float Sa = A.w.Sa[o], Saxh = A.w.Saxh[o];
  const float S_at = col[KP + 0], S_atxh = col[KP + 1], S_adxh = col[KP + 2];
  const float m1 = (float)(gam * Sa / P), m2 = (float)(gam * Saxh / P);
  const float mt1 = (float)((gam * S_at + tgam * Sa) / P);
  const float mt2 = (float)(((double)gam * S_atxh + (double)tgam * Saxh + (double)gam * S_adxh) / P);
  const float d2 = -rstd * rstd * m2;
  const float d1 = -rstd * mt2 + 2.f * rstd * rstd * sdot * m2;
  const float d0 = -rstd * mt1 + rstd * rstd * sdot * m1 + rstd * rstd * m2 * mean_t;
  const float cw = rstd * gam, cd = rstd * tgam - rstd * rstd * sdot * gam;
  if (k < CKK) {
    // (t_W S2)[o,k] + t_b s[k]
    double acc = 0.0;
    for (int l = 0; l < CKK; ++l) acc += (double)A.t_W[o * CKK + l] * A.w.S2[l * KP + k];
    tyx[k] = (float)(acc + (double)tb * A.w.s[k]);
    const float v = cw * col[k] + cd * A.w.Gd[o * KP + k] + d0 * A.w.s[k] + d1 * A.w.G[o * KP + k] + d2 * tyx[k];
    A.at_W[o * CKK + k] += v;
  }
  if (k == 0) {
    if (A.at_gamma) A.at_gamma[o] += S_atxh + S_adxh;
    if (A.at_beta) A.at_beta[o] += S_at;
    if (A.at_b) A.at_b[o] += (float)((double)cw * S_at + (double)cd * Sa + (double)d0 * P + (double)d1 * A.w.sx[o] +
                                     (double)d2 * P * mean_t);
  }
}

template <int C, typename PT>
int run(const CbArgs& A0, int pass, cudaStream_t s) {
  CbArgs A = A0;
  const CbGeom& g = A.g;
  // x tile + transposed pooled tile (+ staged pooled arrays of one tile: xhat*, mask a_q, dxhat* 4 B each, codes 1 B)
  const size_t tcap = (size_t)g.R * g.WP * g.O;
  const size_t tile = 4 * ((size_t)C * g.xrows * g.xpitch + (size_t)g.O * g.wpitch);
  const int wgs = (NT / 32) / ((g.O + 31) / 32);
  const size_t smem_part = 4 * (size_t)wgs * g.O * (KP + NSUM);
  // NHWC mode (fused neighbour): no transposed NCHW staging tile; the reduce kernel stages the bf16 adjoint-tangent rows
  const bool nhwc = A.tq_nhwc != nullptr || A.atq_nhwc != nullptr;
  const size_t xt = 4 * (size_t)C * g.xrows * g.xpitch, nchw_tile = nhwc ? 0 : 4 * (size_t)g.O * g.wpitch;
  const size_t smem_tf = xt + nchw_tile + (sizeof(PT) + 1) * tcap + 64;
  size_t smem_red = xt + nchw_tile + (3 * sizeof(PT) + 1) * tcap + (nhwc ? 2 * (size_t)g.R * g.WP * 64 : 0) + 128;
  if (smem_red < smem_part) smem_red = smem_part;
  static BbOncePerDevice once;
  if (once.need()) {
    BB_CUDA_TRY(cudaFuncSetAttribute(cb_tf_kernel<C, PT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    BB_CUDA_TRY(cudaFuncSetAttribute(cb_reduce_kernel<C, false, PT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    BB_CUDA_TRY(cudaFuncSetAttribute(cb_reduce_kernel<C, true, PT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    BB_CUDA_TRY(cudaFuncSetAttribute(cb_gram_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  }
  if (smem_red > 200 * 1024) return BB_ERR_UNSUPPORTED;
  // persistent grids: exactly the CTAs that are resident at once (a partial second wave would run at a fraction of
  // the occupancy for as long as a full one)
  auto resident = [&](const void* fn, size_t smem) {
    int per_sm = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, NT, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    int gmax = per_sm * BB_SM_COUNT;
    if (gmax > GRID_MAX) gmax = GRID_MAX;
    return g.ntiles < gmax ? (g.ntiles < 1 ? 1 : g.ntiles) : gmax;
  };
  if (pass == BB_PASS_BASE_BWD) {
    const size_t nd = sizeof(double) * (2 * g.O + KP + g.O + g.O * KP + KP * KP);
    BB_CUDA_TRY(cudaMemsetAsync(A.w.dsum, 0, nd, s));
    const int64_t per = (int64_t)g.N * g.HO * g.WO;
    int chunks = (int)((per + 256 * 64 - 1) / (256 * 64));
    if (chunks > 64) chunks = 64;
    if (chunks < 1) chunks = 1;
    cb_stats_kernel<<<dim3(g.O, chunks), 256, 0, s>>>(A);
    cb_stats_finish_kernel<<<1, 64, 0, s>>>(A);
    cb_prep_kernel<PT><<<4 * BB_SM_COUNT, 256, 0, s>>>(A);
    const int bands = (g.HO + 2 * g.R - 1) / (2 * g.R);
    int ggrid = g.N * bands < GRID_MAX ? g.N * bands : GRID_MAX;
    cb_gram_kernel<C><<<ggrid, NT, tile, s>>>(A);
    cb_gram_finish_kernel<<<8, 256, 0, s>>>(A);
    const int grid = resident((const void*)cb_reduce_kernel<C, true, PT>, smem_red);
    A.nparts = grid;
    cb_reduce_kernel<C, true, PT><<<grid, NT, smem_red, s>>>(A);
    cb_finish_kernel<true><<<g.O, 64, 0, s>>>(A, C * 9);
    bb_launch_tally += 9;
    BB_LAUNCH_CHECK();
    return BB_OK;
  }
  if (pass == BB_PASS_TAN_FWD) {
    const int grid = resident((const void*)cb_tf_kernel<C, PT>, smem_tf);
    cb_tf_kernel<C, PT><<<grid, NT, smem_tf, s>>>(A);
    bb_launch_tally += 1;
    BB_LAUNCH_CHECK();
    return BB_OK;
  }
  const int grid = resident((const void*)cb_reduce_kernel<C, false, PT>, smem_red);
  A.nparts = grid;
  cb_reduce_kernel<C, false, PT><<<grid, NT, smem_red, s>>>(A);
  cb_finish_kernel<false><<<g.O, 64, 0, s>>>(A, C * 9);
  bb_launch_tally += 2;
  BB_LAUNCH_CHECK();
  return BB_OK;
}

}  // namespace

// node layout (plan.py _n_convblock):
//   dims = N,C,H,W,O,KH,KW,HO,WO,sh,sw,ph,pw,HP,WP,relu      f[0] = eps
//   slot 0 = W (t, at), slot 1 = conv bias (t, at; may be 0), slot 2 = gamma (base = fp32 values; t, at; may be 0)
//   slot 3 = pooled output q (base, t, a, at)
//   base[0] = x (dt[0]), base[1] = y = conv output (dt[1])
//   aux[0] = workspace (bb_convblock_ws_bytes), aux[1] = t_beta, aux[2] = at_beta, aux[3] = int64 arg-max indices
int bb_launch_convblock(const bb_node& nd, int pass, cudaStream_t s) {
  CbArgs A{};
  A.g = make_geom(nd);
  if (nd.dims[5] != 3 || nd.dims[6] != 3 || A.g.O > MAXO || (A.g.C != 1 && A.g.C != 3)) return BB_ERR_UNSUPPORTED;
  A.w = cb_layout(nd.aux[0], A.g);
  A.x = nd.base[0]; A.dtx = nd.dt[0];
  A.y = nd.base[1]; A.dty = nd.dt[1];
  A.q = nd.base[3]; A.dtq = nd.dt[3];
  A.idx = reinterpret_cast<const int64_t*>(nd.aux[3]);
  A.gamma = reinterpret_cast<const float*>(nd.base[2]);
  A.eps = (float)nd.f[0];
  A.t_W = reinterpret_cast<const float*>(nd.t[0]); A.at_W = reinterpret_cast<float*>(nd.at[0]);
  A.t_b = reinterpret_cast<const float*>(nd.t[1]); A.at_b = reinterpret_cast<float*>(nd.at[1]);
  A.t_gamma = reinterpret_cast<const float*>(nd.t[2]); A.at_gamma = reinterpret_cast<float*>(nd.at[2]);
  A.t_beta = reinterpret_cast<const float*>(nd.aux[1]); A.at_beta = reinterpret_cast<float*>(nd.aux[2]);
  // kind bit 2: the pooled tangent is the next fused block's bf16 padded-NHWC operand ([N][HP+2][WP+2][64], O == 64)
  const bool tq_nhwc = (nd.kind & 4) && A.g.O == 64;
  A.t_q = tq_nhwc ? nullptr : reinterpret_cast<float*>(nd.t[3]);
  A.tq_nhwc = tq_nhwc ? reinterpret_cast<__nv_bfloat16*>(nd.t[3]) : nullptr;
  // kind bit 3: a_q / at_q arrive as bf16 padded NHWC too (written by the next fused block's input-gradient epilogue)
  const bool adj_nhwc = (nd.kind & 8) && A.g.O == 64;
  A.a_q = adj_nhwc ? nullptr : reinterpret_cast<const float*>(nd.a[3]);
  A.at_q = adj_nhwc ? nullptr : reinterpret_cast<const float*>(nd.at[3]);
  A.aq_nhwc = adj_nhwc ? reinterpret_cast<const __nv_bfloat16*>(nd.a[3]) : nullptr;
  A.atq_nhwc = adj_nhwc ? reinterpret_cast<const __nv_bfloat16*>(nd.at[3]) : nullptr;
  if (A.g.ps == 2) return A.g.C == 1 ? run<1, __nv_bfloat16>(A, pass, s) : run<3, __nv_bfloat16>(A, pass, s);
  return A.g.C == 1 ? run<1, float>(A, pass, s) : run<3, float>(A, pass, s);
}

extern "C" int64_t bb_convblock_ws_bytes(int N, int C, int H, int W, int O, int HO, int WO, int HP, int WP) {
  CbGeom g{};
  g.ps = 4;    // upper bound (fp32 pooled arrays)
  g.N = N; g.C = C; g.H = H; g.W = W; g.O = O; g.HO = HO; g.WO = WO; g.HP = HP; g.WP = WP;
  return (int64_t)cb_layout(nullptr, g).bytes;
}
