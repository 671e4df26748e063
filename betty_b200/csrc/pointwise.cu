// K8 (pointwise part): element-wise second-order rules -- unary activations (ReLU, GELU, tanh, sigmoid,
// pow, scale), residual adds, products, strided copies -- and the full reduction used by loss terms.
// Rule semantics: betty_b200/ir.py; executable spec: oracle/plan_interp.py (tf_/bb_/tb_ unary, copy,
// add2, mulc, mul2, sumall).  These replace the ATen element-wise double-backward kernels autograd
// dispatches for reference neumann.py:62 / cg.py:39-41.
#include <stdlib.h>

#include "bb_common.cuh"
#include "../../include/betty_b200.h"
#include "plan.h"

thread_local int bb_launch_tally = 0;

namespace {

struct EwArgs {
  int op, pass, kind, linear, ndim;
  int beta0, beta1;
  int dt0, dt1;
  int64_t n;
  float s0, sa, sb;
  const void* x0;
  const void* x1;
  const float* c;
  const float* t0;
  const float* t1;
  float* ty;
  const float* gy;  // adjoint (BB) or adjoint-tangent (TB) of the output
  const float* ay;  // adjoint of the output (TB curvature terms)
  float* g0;
  float* g1;
  int64_t sizes[BB_MAX_DIMS];
  int64_t st[4][BB_MAX_DIMS];
};

__device__ __forceinline__ void d12(int kind, float x, float s, float& d1, float& d2) {
  switch (kind) {
    case BB_U_RELU:
      d1 = x > 0.f ? 1.f : 0.f;
      d2 = 0.f;
      break;
    case BB_U_TANH: {
      const float y = tanhf(x);
      d1 = 1.f - y * y;
      d2 = -2.f * y * d1;
    } break;
    case BB_U_SIGMOID: {
      const float sg = 1.f / (1.f + expf(-x));
      d1 = sg * (1.f - sg);
      d2 = d1 * (1.f - 2.f * sg);
    } break;
    case BB_U_GELU: {
      const float pdf = 0.3989422804014327f * expf(-0.5f * x * x);
      const float cdf = 0.5f * (1.f + erff(x * 0.7071067811865476f));
      d1 = cdf + x * pdf;
      d2 = pdf * (2.f - x * x);
    } break;
    case BB_U_POW:
      if (s == 2.f) {
        d1 = 2.f * x;
        d2 = 2.f;
      } else {
        d1 = s * powf(x, s - 1.f);
        d2 = (s == 1.f) ? 0.f : s * (s - 1.f) * powf(x, s - 2.f);
      }
      break;
    default:  // BB_U_SCALE
      d1 = s;
      d2 = 0.f;
  }
}

// a null destination means "this input's adjoint is not needed" (base adjoints of parameters)
__device__ __forceinline__ void put(float* p, int64_t o, float v, int beta) {
  if (p != nullptr) p[o] = beta ? p[o] + v : v;
}

constexpr int kEwThreads = 256;

__global__ void __launch_bounds__(kEwThreads) ew_kernel(const __grid_constant__ EwArgs A) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < A.n; i += stride) {
    int64_t o0 = i, o1 = i, o2 = i, o3 = i;
    if (!A.linear) {
      int64_t rem = i;
      o0 = o1 = o2 = o3 = 0;
#pragma unroll 1
      for (int d = A.ndim - 1; d >= 0; --d) {
        const int64_t q = rem / A.sizes[d];
        const int64_t r = rem - q * A.sizes[d];
        rem = q;
        o0 += r * A.st[0][d];
        o1 += r * A.st[1][d];
        o2 += r * A.st[2][d];
        o3 += r * A.st[3][d];
      }
    }
    switch (A.op) {
      case BB_OP_UNARY: {
        float d1, d2;
        d12(A.kind, bb::ldf(A.x0, o0, A.dt0), A.s0, d1, d2);
        if (A.pass == BB_PASS_TAN_FWD) {
          A.ty[o3] = d1 * A.t0[o0];
        } else if (A.pass == BB_PASS_BASE_BWD) {
          put(A.g0, o0, d1 * A.gy[o3], A.beta0);
        } else {
          put(A.g0, o0, d1 * A.gy[o3] + d2 * A.t0[o0] * A.ay[o3], A.beta0);
        }
      } break;
      case BB_OP_COPY:
        if (A.pass == BB_PASS_TAN_FWD) A.ty[o3] = A.t0[o0];
        else put(A.g0, o0, A.gy[o3], A.beta0);
        break;
      case BB_OP_ADD2:
        if (A.pass == BB_PASS_TAN_FWD) {
          A.ty[o3] = A.sa * A.t0[o0] + A.sb * A.t1[o1];
        } else {
          const float g = A.gy[o3];
          put(A.g0, o0, A.sa * g, A.beta0);
          put(A.g1, o1, A.sb * g, A.beta1);
        }
        break;
      case BB_OP_MULC: {
        const float c = A.c[o2];
        if (A.pass == BB_PASS_TAN_FWD) A.ty[o3] = c * A.t0[o0];
        else put(A.g0, o0, c * A.gy[o3], A.beta0);
      } break;
      case BB_OP_MUL2: {
        const float a = bb::ldf(A.x0, o0, A.dt0), b = bb::ldf(A.x1, o1, A.dt1);
        if (A.pass == BB_PASS_TAN_FWD) {
          A.ty[o3] = A.t0[o0] * b + a * A.t1[o1];
        } else if (A.pass == BB_PASS_BASE_BWD) {
          const float g = A.gy[o3];
          put(A.g0, o0, g * b, A.beta0);
          put(A.g1, o1, g * a, A.beta1);
        } else {
          const float gt = A.gy[o3], g = A.ay[o3];
          const float ta = A.t0[o0], tb = A.t1[o1];
          put(A.g0, o0, gt * b + g * tb, A.beta0);
          put(A.g1, o1, gt * a + g * ta, A.beta1);
        }
      } break;
    }
  }
}

// ---- 128-bit variant for dense operands with identical strides (nd.linear), n % 4 == 0, 16-byte aligned buffers:
// four elements per thread, float4 loads / stores (8-byte loads for 16-bit base tensors).  Operands whose
// coefficient is identically zero are not read at all (ReLU and scale have no second-derivative term: the tangent
// backward pass does not need t_x and a_y) -- the scalar kernel paid 8 of its 18 bytes per element for those.
__device__ __forceinline__ void ld4f(const float* p, int64_t i, float* o) {
  const float4 q = *reinterpret_cast<const float4*>(p + i);
  o[0] = q.x; o[1] = q.y; o[2] = q.z; o[3] = q.w;
}
__device__ __forceinline__ void ld4b(const void* p, int64_t i, int dt, float* o) {
  if (dt == BB_F32) {
    ld4f(reinterpret_cast<const float*>(p), i, o);
    return;
  }
  const uint2 r = *reinterpret_cast<const uint2*>(reinterpret_cast<const unsigned short*>(p) + i);
  if (dt == BB_BF16) {
    o[0] = __uint_as_float(r.x << 16); o[1] = __uint_as_float(r.x & 0xffff0000u);
    o[2] = __uint_as_float(r.y << 16); o[3] = __uint_as_float(r.y & 0xffff0000u);
  } else {
    const float2 f0 = __half22float2(*reinterpret_cast<const __half2*>(&r.x));
    const float2 f1 = __half22float2(*reinterpret_cast<const __half2*>(&r.y));
    o[0] = f0.x; o[1] = f0.y; o[2] = f1.x; o[3] = f1.y;
  }
}
__device__ __forceinline__ void put4(float* p, int64_t i, const float* v, int beta) {
  if (p == nullptr) return;
  float4* q = reinterpret_cast<float4*>(p + i);
  float4 w = make_float4(v[0], v[1], v[2], v[3]);
  if (beta) {
    const float4 o = *q;
    w.x += o.x; w.y += o.y; w.z += o.z; w.w += o.w;
  }
  *q = w;
}

template <int OP>
__global__ void __launch_bounds__(kEwThreads) ew_vec4_kernel(const __grid_constant__ EwArgs A) {
  const int64_t n4 = A.n >> 2, stride = (int64_t)gridDim.x * blockDim.x;
  const bool curved = A.kind != BB_U_RELU && A.kind != BB_U_SCALE;   // unary ops with a non-zero second derivative
  for (int64_t i4 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i4 < n4; i4 += stride) {
    const int64_t i = i4 << 2;
    float v[4], w[4];
    if (OP == BB_OP_UNARY) {
      float x[4], d1[4], d2[4];
      ld4b(A.x0, i, A.dt0, x);
#pragma unroll
      for (int e = 0; e < 4; ++e) d12(A.kind, x[e], A.s0, d1[e], d2[e]);
      if (A.pass == BB_PASS_TAN_FWD) {
        float t[4];
        ld4f(A.t0, i, t);
#pragma unroll
        for (int e = 0; e < 4; ++e) v[e] = d1[e] * t[e];
        put4(A.ty, i, v, 0);
      } else {
        float g[4];
        ld4f(A.gy, i, g);
#pragma unroll
        for (int e = 0; e < 4; ++e) v[e] = d1[e] * g[e];
        if (A.pass == BB_PASS_TAN_BWD && curved) {
          float t[4], a[4];
          ld4f(A.t0, i, t);
          ld4f(A.ay, i, a);
#pragma unroll
          for (int e = 0; e < 4; ++e) v[e] += d2[e] * t[e] * a[e];
        }
        put4(A.g0, i, v, A.beta0);
      }
    } else if (OP == BB_OP_COPY) {
      if (A.pass == BB_PASS_TAN_FWD) {
        ld4f(A.t0, i, v);
        put4(A.ty, i, v, 0);
      } else {
        ld4f(A.gy, i, v);
        put4(A.g0, i, v, A.beta0);
      }
    } else if (OP == BB_OP_ADD2) {
      if (A.pass == BB_PASS_TAN_FWD) {
        float a[4], b[4];
        ld4f(A.t0, i, a);
        ld4f(A.t1, i, b);
#pragma unroll
        for (int e = 0; e < 4; ++e) v[e] = A.sa * a[e] + A.sb * b[e];
        put4(A.ty, i, v, 0);
      } else {
        float g[4];
        ld4f(A.gy, i, g);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          v[e] = A.sa * g[e];
          w[e] = A.sb * g[e];
        }
        put4(A.g0, i, v, A.beta0);
        put4(A.g1, i, w, A.beta1);
      }
    } else {   // BB_OP_MUL2
      float a[4], b[4];
      ld4b(A.x0, i, A.dt0, a);
      ld4b(A.x1, i, A.dt1, b);
      if (A.pass == BB_PASS_TAN_FWD) {
        float ta[4], tb[4];
        ld4f(A.t0, i, ta);
        ld4f(A.t1, i, tb);
#pragma unroll
        for (int e = 0; e < 4; ++e) v[e] = ta[e] * b[e] + a[e] * tb[e];
        put4(A.ty, i, v, 0);
      } else if (A.pass == BB_PASS_BASE_BWD) {
        float g[4];
        ld4f(A.gy, i, g);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          v[e] = g[e] * b[e];
          w[e] = g[e] * a[e];
        }
        put4(A.g0, i, v, A.beta0);
        put4(A.g1, i, w, A.beta1);
      } else {
        float gt[4], g[4], ta[4], tb[4];
        ld4f(A.gy, i, gt);
        ld4f(A.ay, i, g);
        ld4f(A.t0, i, ta);
        ld4f(A.t1, i, tb);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          v[e] = gt[e] * b[e] + g[e] * tb[e];
          w[e] = gt[e] * a[e] + g[e] * ta[e];
        }
        put4(A.g0, i, v, A.beta0);
        put4(A.g1, i, w, A.beta1);
      }
    }
  }
}

// ---- sumall ----------------------------------------------------------------------------------
constexpr int kRedThreads = 512;
constexpr int kRedMaxGrid = BB_SM_COUNT * 2;

// scratch layout (doubles): [0, kRedMaxGrid) partials, then one uint ticket
__global__ void __launch_bounds__(kRedThreads) sum_fwd_kernel(const float* __restrict__ tx, float* ty, float scale,
                                                              int64_t n, double* scratch) {
  __shared__ double red[32];
  float acc = 0.f;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) acc += tx[i];
  const double b = bb::block_sum<double>((double)acc, red);
  double total;
  unsigned int* ticket = reinterpret_cast<unsigned int*>(scratch + kRedMaxGrid);
  if (bb::grid_sum_finish(b, scratch, ticket, &total, red)) ty[0] = (float)(scale * total);
}

__global__ void __launch_bounds__(kEwThreads) sum_bwd_kernel(float* gx, const float* __restrict__ gy, float scale,
                                                             int64_t n, int beta) {
  const float g = scale * gy[0];
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) gx[i] = beta ? gx[i] + g : g;
}

inline int grid_1d(int64_t n, int threads, int max_blocks) {
  int64_t g = (n + threads - 1) / threads;
  if (g < 1) g = 1;
  if (g > max_blocks) g = max_blocks;
  return (int)g;
}

}  // namespace

int bb_launch_ew(const bb_node& nd, int pass, cudaStream_t s) {
  EwArgs A;
  A.op = nd.op;
  A.pass = pass;
  A.kind = nd.kind;
  A.linear = nd.linear;
  A.ndim = nd.ndim;
  A.beta0 = nd.beta[0];
  A.beta1 = nd.beta[1];
  A.dt0 = nd.dt[0];
  A.dt1 = nd.dt[1];
  A.n = nd.n;
  A.s0 = (float)nd.f[0];
  A.sa = (float)nd.f[1];
  A.sb = (float)nd.f[2];
  A.x0 = nd.base[0];
  A.x1 = nd.base[1];
  A.c = reinterpret_cast<const float*>(nd.aux[0]);
  A.t0 = reinterpret_cast<const float*>(nd.t[0]);
  A.t1 = reinterpret_cast<const float*>(nd.t[1]);
  A.ty = reinterpret_cast<float*>(nd.t[3]);
  A.ay = reinterpret_cast<const float*>(nd.a[3]);
  if (pass == BB_PASS_BASE_BWD) {
    // nd.pad0 = mask of inputs whose base adjoint is needed (activations, not parameters)
    A.gy = reinterpret_cast<const float*>(nd.a[3]);
    A.g0 = (nd.pad0 & 1) ? reinterpret_cast<float*>(nd.a[0]) : nullptr;
    A.g1 = (nd.pad0 & 2) ? reinterpret_cast<float*>(nd.a[1]) : nullptr;
    if (A.g0 == nullptr && A.g1 == nullptr) return BB_OK;
  } else {
    A.gy = reinterpret_cast<const float*>(nd.at[3]);
    A.g0 = reinterpret_cast<float*>(nd.at[0]);
    A.g1 = reinterpret_cast<float*>(nd.at[1]);
  }
  for (int d = 0; d < BB_MAX_DIMS; ++d) {
    A.sizes[d] = nd.sizes[d];
    for (int k = 0; k < 4; ++k) A.st[k][d] = nd.stride[k][d];
  }
  if (A.n <= 0) return BB_OK;
  // 128-bit path: every operand this op/pass touches must be dense-linear and 16-byte aligned (8 for 16-bit bases)
  auto al = [](const void* p, uintptr_t m) { return (reinterpret_cast<uintptr_t>(p) & m) == 0; };
  const bool two = A.op == BB_OP_ADD2 || A.op == BB_OP_MUL2;
  const bool reads_base0 = A.op == BB_OP_UNARY || A.op == BB_OP_MUL2, reads_base1 = A.op == BB_OP_MUL2;
  bool vec = A.linear && (A.n & 3) == 0 && A.op != BB_OP_MULC && !getenv("BB200_EW_SCALAR");
  vec = vec && al(A.t0, 15) && al(A.ty, 15) && al(A.gy, 15) && al(A.ay, 15) && al(A.g0, 15) && al(A.g1, 15);
  vec = vec && (!two || al(A.t1, 15));
  vec = vec && (!reads_base0 || al(A.x0, A.dt0 == BB_F32 ? 15 : 7)) && (!reads_base1 || al(A.x1, A.dt1 == BB_F32 ? 15 : 7));
  if (vec) {
    const int grid = grid_1d(A.n >> 2, kEwThreads, BB_SM_COUNT * 8);
    switch (A.op) {
      case BB_OP_UNARY: ew_vec4_kernel<BB_OP_UNARY><<<grid, kEwThreads, 0, s>>>(A); break;
      case BB_OP_COPY: ew_vec4_kernel<BB_OP_COPY><<<grid, kEwThreads, 0, s>>>(A); break;
      case BB_OP_ADD2: ew_vec4_kernel<BB_OP_ADD2><<<grid, kEwThreads, 0, s>>>(A); break;
      default: ew_vec4_kernel<BB_OP_MUL2><<<grid, kEwThreads, 0, s>>>(A);
    }
    bb_launch_tally += 1;
    BB_LAUNCH_CHECK();
    return BB_OK;
  }
  ew_kernel<<<grid_1d(A.n, kEwThreads, BB_SM_COUNT * 8), kEwThreads, 0, s>>>(A);
  bb_launch_tally += 1;
  BB_LAUNCH_CHECK();
  return BB_OK;
}

int bb_launch_sumall(const bb_node& nd, int pass, cudaStream_t s) {
  const float scale = (float)nd.f[0];
  if (nd.n <= 0) return BB_OK;
  if (pass == BB_PASS_TAN_FWD) {
    sum_fwd_kernel<<<grid_1d(nd.n, kRedThreads, kRedMaxGrid), kRedThreads, 0, s>>>(
        reinterpret_cast<const float*>(nd.t[0]), reinterpret_cast<float*>(nd.t[3]), scale, nd.n,
        reinterpret_cast<double*>(nd.aux[0]));
  } else {
    const bool bb = pass == BB_PASS_BASE_BWD;
    if (bb && !(nd.pad0 & 1)) return BB_OK;
    sum_bwd_kernel<<<grid_1d(nd.n, kEwThreads, BB_SM_COUNT * 8), kEwThreads, 0, s>>>(
        reinterpret_cast<float*>(bb ? nd.a[0] : nd.at[0]), reinterpret_cast<const float*>(bb ? nd.a[3] : nd.at[3]),
        scale, nd.n, nd.beta[0]);
  }
  bb_launch_tally += 1;
  BB_LAUNCH_CHECK();
  return BB_OK;
}
