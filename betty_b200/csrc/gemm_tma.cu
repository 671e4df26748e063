// K5/K6 tensor-core path, TMA-fed: dual-product GEMM and stride-1 convolution (forward / input-gradient form) on
// tcgen05.mma with operands brought into shared memory by the TMA unit.
//
//   D[m][n] (beta/atomic)= sum over up to two operand pairs p of  sum_k A_p[m][k] * B_p[n][k]   (+ bias[n])
//
// The software-staged kernel in gemm_tc.cu spends >95 % of its time in the producer warps (address arithmetic,
// 4-byte gathers, bf16 conversion, swizzled stores; ncu: tensor pipe ~1 %).  Here every operand is first made
// TMA-addressable -- bf16, unit stride along one matrix dimension, 16-byte aligned rows; base activations and
// autocast weight copies already are, tangents / adjoints (fp32 arena slices) go through a streaming pack kernel --
// and one elected thread issues cp.async.bulk.tensor loads that land in the SWIZZLE_128B layout the UMMA
// descriptors read:
//   K-major operand  (k contiguous in memory): one box of 64 k x ROWS rows      -> rows of 128 B
//   MN-major operand (m/n contiguous):         ROWS/64 boxes of 64 mn x 64 k    -> 8x64 atoms, SBO 1024 B, LBO 8 KB
//   convolution A    (NHWC bf16 activations):  one 4-D box 64 ch x Wb x Hb x 1 per (tap, channel block); the
//                    window displacement is a coordinate offset, zero padding is TMA out-of-bounds fill
// so transposed views cost nothing and an implicit-GEMM convolution is nine shifted box loads per 64 channels.
//
// Roles (192 threads): warp 0 = TMA producer (one lane), warp 1 = TMEM alloc + MMA issue (one lane), warps 2-5 =
// epilogue (tcgen05.ld 32 lanes x 32 columns, warp w owns TMEM lanes 32*(w%4)..).  3-4 stage mbarrier ring; two
// CTAs per SM so one CTA's epilogue overlaps the other's main loop.
#include <cuda_bf16.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <string>
#include <unordered_map>

#include "../../include/betty_b200.h"
#include "bb_common.cuh"
#include "gemm_tma.h"
#include "plan.h"
#include "tc_ptx.cuh"
#include "tma.h"

thread_local BbScratch bb_scratch = {nullptr, 0, 0};
thread_local uint64_t bb_scratch_gen = 0;

namespace {

using namespace bbtc;

constexpr int BM = 128, BK = 64;
constexpr int NTHREADS = 192;
constexpr int A_TILE = BM * BK * 2;   // 16 KB

constexpr int MAX_STAGES = 8;

template <int BN_>
struct Cfg {
  static constexpr int kStagesShared = BN_ == 64 ? 4 : 3;   // two CTAs per SM
  static constexpr int kStagesDeep = BN_ == 64 ? 8 : 6;     // one CTA per SM, ~190 KB in flight
  static constexpr int kBTile = BN_ * BK * 2;
  static constexpr int kStage = A_TILE + kBTile;
  static constexpr size_t smem(int stages) { return (size_t)stages * kStage + 1024 + 256; }
};

// Epilogue store of one 32-column chunk of a thread's accumulator row.  Written for a low instruction count: the
// epilogue warps are four single warps per CTA, so their time is (instructions x dependent-issue latency), not
// bandwidth -- the first version spent ~30 integer instructions per stored value on 64-bit index arithmetic and
// per-element predicates (ncu source page: 7.7 us per 128x64 tile, see profiles/r01_persist_ncu.md).  Here every mode
// decision is made once per chunk and addresses advance by pointer increments.
__device__ __forceinline__ void epi_store(const TmaGemmArgs& G, uint32_t (&v)[32], float* dst, int64_t col_stride, int col0,
                                          bool add_bias, bool vec, bool vec_red) {
  const int ncols = (G.N - col0) < 32 ? (int)(G.N - col0) : 32;
  const bool full = ncols == 32;
  if (add_bias) {
    const float* bp = G.bias + (int64_t)col0 * G.bias_stride;
    if (full && G.bias_stride == 1 && (reinterpret_cast<uintptr_t>(bp) & 15) == 0) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 b = *reinterpret_cast<const float4*>(bp + j);
        v[j] = __float_as_uint(__uint_as_float(v[j]) + b.x);
        v[j + 1] = __float_as_uint(__uint_as_float(v[j + 1]) + b.y);
        v[j + 2] = __float_as_uint(__uint_as_float(v[j + 2]) + b.z);
        v[j + 3] = __float_as_uint(__uint_as_float(v[j + 3]) + b.w);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < ncols) v[j] = __float_as_uint(__uint_as_float(v[j]) + bp[(int64_t)j * G.bias_stride]);
    }
  }
  if (full && vec) {
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      float4 o = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                             __uint_as_float(v[j + 3]));
      float4* q = reinterpret_cast<float4*>(dst + j);
      if (G.beta) {
        const float4 old = *q;
        o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
      }
      *q = o;
    }
    return;
  }
  if (full && vec_red) {
    // split-K partial sums: 128-bit reductions (red.global.add.v4.f32), a quarter of the atomic traffic
#pragma unroll
    for (int j = 0; j < 32; j += 4)
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + j), "f"(__uint_as_float(v[j])),
                   "f"(__uint_as_float(v[j + 1])), "f"(__uint_as_float(v[j + 2])), "f"(__uint_as_float(v[j + 3]))
                   : "memory");
    return;
  }
  float* q = dst;
  if (G.ksplit > 1) {
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      if (j < ncols) atomicAdd(q, __uint_as_float(v[j]));
      q += col_stride;
    }
  } else if (G.beta) {
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      if (j < ncols) *q += __uint_as_float(v[j]);
      q += col_stride;
    }
  } else if (full) {
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      *q = __uint_as_float(v[j]);
      q += col_stride;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      if (j < ncols) *q = __uint_as_float(v[j]);
      q += col_stride;
    }
  }
}

template <int BN_>
__global__ void __launch_bounds__(NTHREADS, 2) gemm_tma_kernel(const __grid_constant__ TmaGemmArgs G) {
  using C = Cfg<BN_>;
  const int STAGES = G.stages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * C::kStage);
  const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + MAX_STAGES), accum = smem_u32(bars + 2 * MAX_STAGES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * MAX_STAGES + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n0 = blockIdx.x * BN_;
  const int bz = blockIdx.z / G.ksplit;            // batch index (ksplit == 1 when batched)
  const int split = blockIdx.z - bz * G.ksplit;
  const int64_t kblocks_total = (G.K + BK - 1) / BK;
  const int64_t kb_per = (kblocks_total + G.ksplit - 1) / G.ksplit;
  const int64_t kb_beg = (int64_t)split * kb_per;
  int64_t kb_end = kb_beg + kb_per;
  if (kb_end > kblocks_total) kb_end = kblocks_total;
  const int nkb = (int)(kb_end > kb_beg ? kb_end - kb_beg : 0);
  const int total_kb = nkb * G.npairs;

  const bool conv = G.a_kind[0] == TMA_CONV;
  int m0 = 0, img = 0, h0 = 0;
  if (conv) {
    img = blockIdx.y / G.tiles_per_img;
    h0 = (blockIdx.y - img * G.tiles_per_img) * G.Hb;
  } else {
    m0 = blockIdx.y * BM;
  }

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, 1);
    }
    mbar_init(accum, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), (uint32_t)BN_);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ---------------- TMA producer ----------------
    if (elect_one()) {
      for (int p = 0; p < G.npairs; ++p) {
        tma_prefetch_desc(&G.a[p]);
        tma_prefetch_desc(&G.b[p]);
      }
      for (int it = 0; it < total_kb; ++it) {
        const int s = it % STAGES;
        if (it >= STAGES) mbar_wait(empty0 + 8 * s, ((it / STAGES) - 1) & 1);
        const int pair = it / nkb;
        const int kb = (int)kb_beg + (it - pair * nkb);
        const int k0 = kb * BK;
        const uint32_t bar = full0 + 8 * s;
        const uint32_t sa = smem_u32(smem + s * C::kStage), sb = sa + A_TILE;
        mbar_expect_tx(bar, G.a_bytes + G.b_bytes);
        const int ak = G.a_kind[pair];
        if (ak == TMA_KMAJ) {
          tma_load_3d(sa, &G.a[pair], bar, k0, m0, bz);
        } else if (ak == TMA_MNMAJ) {
          tma_load_3d(sa, &G.a[pair], bar, m0, k0, bz);
          tma_load_3d(sa + 8192, &G.a[pair], bar, m0 + 64, k0, bz);
        } else {
          const int tap = kb / G.cblocks, cb = kb - tap * G.cblocks;
          const int i = tap / G.KW, j = tap - i * G.KW;
          const int dy = G.flip ? G.ph - i : i - G.ph, dx = G.flip ? G.pw - j : j - G.pw;
          tma_load_4d(sa, &G.a[pair], bar, cb * 64, dx, h0 + dy, img);
        }
        const int bzb = ak == TMA_CONV ? 0 : bz;
        if (G.b_kind[pair] == TMA_KMAJ) {
          tma_load_3d(sb, &G.b[pair], bar, k0, n0, bzb);
        } else {
#pragma unroll
          for (int q = 0; q < BN_ / 64; ++q) tma_load_3d(sb + q * 8192, &G.b[pair], bar, n0 + q * 64, k0, bzb);
        }
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer ----------------
    if (elect_one()) {
      for (int it = 0; it < total_kb; ++it) {
        const int s = it % STAGES;
        const int pair = it / nkb;
        const bool a_mn = G.a_kind[pair] == TMA_MNMAJ, b_mn = G.b_kind[pair] == TMA_MNMAJ;
        const uint32_t idesc = idesc_bf16(BM, BN_, a_mn, b_mn);
        mbar_wait(full0 + 8 * s, (it / STAGES) & 1);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + s * C::kStage), b_addr = a_addr + A_TILE;
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          const uint64_t da = a_mn ? desc_mn(a_addr + k * 2048, 8192) : desc_k(a_addr + k * 32);
          const uint64_t db = b_mn ? desc_mn(b_addr + k * 2048, 8192) : desc_k(b_addr + k * 32);
          umma_bf16(tmem_base, da, db, idesc, (it > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(empty0 + 8 * s);
      }
      if (total_kb > 0) umma_commit(accum);
    }
    __syncwarp();
  } else {
    // ---------------- epilogue (warps 2..5) ----------------
    if (total_kb > 0) {
      mbar_wait(accum, 0, 100);
      tc_fence_after();
    }
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    bool row_ok;
    int64_t row_base, col_stride;
    if (G.omode == 1) {
      const int64_t q = (int64_t)h0 * G.Wb + r;   // Wb == output width
      row_ok = r < G.Wb * G.Hb && q < G.OHW;
      row_base = (int64_t)img * G.OCH * G.OHW + q;
      col_stride = G.OHW;
    } else if (G.omode == 2) {
      const int64_t row = (int64_t)m0 + r;
      row_ok = row < G.M;
      const int64_t im = row / G.OHW;
      row_base = im * G.OCH * G.OHW + (row - im * G.OHW);
      col_stride = G.OHW;
    } else {
      const int64_t row = (int64_t)m0 + r;
      row_ok = row < G.M;
      row_base = (int64_t)bz * G.obs + row * G.ors;
      col_stride = G.ocs;
    }
    const bool row_vec = G.omode == 0 && G.ocs == 1 && (G.ors & 3) == 0 && (G.obs & 3) == 0 &&
                         ((reinterpret_cast<uintptr_t>(G.out) & 15) == 0);
    const bool vec = row_vec && G.ksplit == 1, vec_red = row_vec && G.ksplit > 1;
#pragma unroll 1
    for (int c = 0; c < BN_ / 32; ++c) {
      uint32_t v[32];
      if (total_kb > 0) {
        tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(c * 32), v);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0u;
      }
      if (!row_ok) continue;
      const int col0 = n0 + c * 32;
      if (col0 >= G.N) continue;
      epi_store(G, v, G.out + row_base + (int64_t)col0 * col_stride, col_stride, col0, G.bias != nullptr && split == 0, vec,
                vec_red);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)BN_);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Persistent variant for launches with many more tiles than SMs (convolutions: one tile per 128 pixels).  A CTA
// walks tiles blockIdx.x, +gridDim.x, ...; the TMA producer and the MMA warp run ahead across tile boundaries on the
// same stage ring, and the accumulator is double-buffered in TMEM (2 x BN columns) so the epilogue of tile i
// (tcgen05.ld + global stores) overlaps the loads and MMAs of tile i+1.  The one-tile-per-CTA kernel above pays
// barrier init + TMEM allocation + a cold TMA round trip per tile (~4-8 us for a 1-9 k-block tile).
//   barriers: full[S] / empty[S] (stage ring), acc_full[2] (MMA -> epilogue), acc_empty[2] (4 epilogue warps -> MMA)
// No split-K, no batch (conv and plane-output GEMMs only).
template <int BN_>
__global__ void __launch_bounds__(NTHREADS, 2) gemm_tma_persist_kernel(const __grid_constant__ TmaGemmArgs G, int ntiles_n,
                                                                       int total_tiles) {
  using C = Cfg<BN_>;
  const int STAGES = G.stages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * C::kStage);
  const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + MAX_STAGES);
  const uint32_t accf0 = smem_u32(bars + 2 * MAX_STAGES), acce0 = smem_u32(bars + 2 * MAX_STAGES + 2);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * MAX_STAGES + 4);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nkb = (int)((G.K + BK - 1) / BK);
  const int per_tile = nkb * G.npairs;
  const bool conv = G.a_kind[0] == TMA_CONV;

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(accf0 + 8 * b, 1);
      mbar_init(acce0 + 8 * b, 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), (uint32_t)(2 * BN_));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ---------------- TMA producer ----------------
    if (elect_one()) {
      for (int p = 0; p < G.npairs; ++p) {
        tma_prefetch_desc(&G.a[p]);
        tma_prefetch_desc(&G.b[p]);
      }
      int git = 0;   // stage-ring position, continues across tiles
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int mt = tile / ntiles_n, n0 = (tile - mt * ntiles_n) * BN_;
        int m0 = 0, img = 0, h0 = 0;
        if (conv) {
          img = mt / G.tiles_per_img;
          h0 = (mt - img * G.tiles_per_img) * G.Hb;
        } else {
          m0 = mt * BM;
        }
        for (int it = 0; it < per_tile; ++it, ++git) {
          const int s = git % STAGES;
          if (git >= STAGES) mbar_wait(empty0 + 8 * s, ((git / STAGES) - 1) & 1);
          const int pair = it / nkb, kb = it - pair * nkb, k0 = kb * BK;
          const uint32_t bar = full0 + 8 * s;
          const uint32_t sa = smem_u32(smem + s * C::kStage), sb = sa + A_TILE;
          mbar_expect_tx(bar, G.a_bytes + G.b_bytes);
          const int ak = G.a_kind[pair];
          if (ak == TMA_KMAJ) {
            tma_load_3d(sa, &G.a[pair], bar, k0, m0, 0);
          } else if (ak == TMA_MNMAJ) {
            tma_load_3d(sa, &G.a[pair], bar, m0, k0, 0);
            tma_load_3d(sa + 8192, &G.a[pair], bar, m0 + 64, k0, 0);
          } else {
            const int tap = kb / G.cblocks, cb = kb - tap * G.cblocks;
            const int i = tap / G.KW, j = tap - i * G.KW;
            const int dy = G.flip ? G.ph - i : i - G.ph, dx = G.flip ? G.pw - j : j - G.pw;
            tma_load_4d(sa, &G.a[pair], bar, cb * 64, dx, h0 + dy, img);
          }
          if (G.b_kind[pair] == TMA_KMAJ) {
            tma_load_3d(sb, &G.b[pair], bar, k0, n0, 0);
          } else {
#pragma unroll
            for (int q = 0; q < BN_ / 64; ++q) tma_load_3d(sb + q * 8192, &G.b[pair], bar, n0 + q * 64, k0, 0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer ----------------
    if (elect_one()) {
      int git = 0, lt = 0;   // lt = this CTA's tile counter
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++lt) {
        const int buf = lt & 1;
        if (lt >= 2) {   // the epilogue must have drained this accumulator (tile lt-2)
          mbar_wait(acce0 + 8 * buf, ((lt >> 1) - 1) & 1);
          tc_fence_after();
        }
        const uint32_t tacc = tmem_base + (uint32_t)(buf * BN_);
        for (int it = 0; it < per_tile; ++it, ++git) {
          const int s = git % STAGES;
          const int pair = it / nkb;
          const bool a_mn = G.a_kind[pair] == TMA_MNMAJ, b_mn = G.b_kind[pair] == TMA_MNMAJ;
          const uint32_t idesc = idesc_bf16(BM, BN_, a_mn, b_mn);
          mbar_wait(full0 + 8 * s, (git / STAGES) & 1);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + s * C::kStage), b_addr = a_addr + A_TILE;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t da = a_mn ? desc_mn(a_addr + k * 2048, 8192) : desc_k(a_addr + k * 32);
            const uint64_t db = b_mn ? desc_mn(b_addr + k * 2048, 8192) : desc_k(b_addr + k * 32);
            umma_bf16(tacc, da, db, idesc, (it > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(empty0 + 8 * s);
        }
        umma_commit(accf0 + 8 * buf);
      }
    }
    __syncwarp();
  } else {
    // ---------------- epilogue (warps 2..5) ----------------
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const bool vec = G.omode == 0 && G.ocs == 1 && (G.ors & 3) == 0 && ((reinterpret_cast<uintptr_t>(G.out) & 15) == 0);
    int lt = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++lt) {
      const int buf = lt & 1;
      const int mt = tile / ntiles_n, n0 = (tile - mt * ntiles_n) * BN_;
      bool row_ok;
      int64_t row_base, col_stride;
      if (G.omode == 1) {
        const int img = mt / G.tiles_per_img, h0 = (mt - img * G.tiles_per_img) * G.Hb;
        const int64_t q = (int64_t)h0 * G.Wb + r;
        row_ok = r < G.Wb * G.Hb && q < G.OHW;
        row_base = (int64_t)img * G.OCH * G.OHW + q;
        col_stride = G.OHW;
      } else if (G.omode == 2) {
        const int64_t row = (int64_t)mt * BM + r;
        row_ok = row < G.M;
        const int64_t im = row / G.OHW;
        row_base = im * G.OCH * G.OHW + (row - im * G.OHW);
        col_stride = G.OHW;
      } else {
        const int64_t row = (int64_t)mt * BM + r;
        row_ok = row < G.M;
        row_base = row * G.ors;
        col_stride = G.ocs;
      }
      mbar_wait(accf0 + 8 * buf, (lt >> 1) & 1, 60);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < BN_ / 32; ++c) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * BN_ + c * 32), v);
        if (c == BN_ / 32 - 1) {   // last read of this accumulator: hand it back to the MMA warp before storing
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(acce0 + 8 * buf);
        }
        const int col0 = n0 + c * 32;
        if (!row_ok || col0 >= G.N) continue;
        epi_store(G, v, G.out + row_base + (int64_t)col0 * col_stride, col_stride, col0, G.bias != nullptr, vec, false);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)(2 * BN_));
  }
}

template <int BN_>
int launch_tma(TmaGemmArgs& G, int64_t mtiles, cudaStream_t s, int batch) {
  using C = Cfg<BN_>;
  static BbOncePerDevice configured;
  if (configured.need()) {
    BB_CUDA_TRY(cudaFuncSetAttribute(gemm_tma_kernel<BN_>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)C::smem(C::kStagesDeep)));
  }
  dim3 grid((unsigned)((G.N + BN_ - 1) / BN_), (unsigned)mtiles, (unsigned)(G.ksplit * batch));
  // A grid that does not fill the machine is latency-bound on its k-loop (each TMA round trip is ~1.5 us): give the
  // lone CTA of each SM the whole shared memory as prefetch depth.  Full grids keep two CTAs per SM instead, so one
  // CTA's epilogue overlaps the other's main loop.
  static const int forced = getenv("BB200_TMA_STAGES") ? atoi(getenv("BB200_TMA_STAGES")) : 0;
  const int64_t ctas = (int64_t)grid.x * grid.y * grid.z;
  G.stages = ctas <= BB_SM_COUNT ? C::kStagesDeep : C::kStagesShared;
  if (forced >= 2 && forced <= C::kStagesDeep) G.stages = forced;
  static const bool no_persist = getenv("BB200_TMA_NO_PERSIST") != nullptr;
  if (!no_persist && G.ksplit == 1 && batch == 1 && ctas > 4 * BB_SM_COUNT) {
    static BbOncePerDevice configured_p;
    if (configured_p.need()) {
      BB_CUDA_TRY(cudaFuncSetAttribute(gemm_tma_persist_kernel<BN_>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)C::smem(C::kStagesDeep)));
    }
    G.stages = C::kStagesShared;
    gemm_tma_persist_kernel<BN_><<<2 * BB_SM_COUNT, NTHREADS, C::smem(G.stages), s>>>(G, (int)grid.x, (int)(grid.x * grid.y));
    bb_launch_tally += 1;
    BB_LAUNCH_CHECK();
    return BB_OK;
  }
  gemm_tma_kernel<BN_><<<grid, NTHREADS, C::smem(G.stages), s>>>(G);
  bb_launch_tally += 1;
  BB_LAUNCH_CHECK();
  return BB_OK;
}

// ------------------------------------------------------------------------------------------------
// tensor maps
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = []() -> EncodeTiledFn {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess) return nullptr;
    if (q != cudaDriverEntryPointSuccess) return nullptr;
    return reinterpret_cast<EncodeTiledFn>(f);
  }();
  return fn;
}

std::mutex g_map_mu;
std::unordered_map<std::string, CUtensorMap> g_maps;

int encode_cached(CUtensorMap* out, int rank, const void* p, const cuuint64_t* dims, const cuuint64_t* strides,
                  const cuuint32_t* box) {
  std::string key(reinterpret_cast<const char*>(&p), sizeof(p));
  key.append(reinterpret_cast<const char*>(dims), sizeof(cuuint64_t) * rank);
  key.append(reinterpret_cast<const char*>(strides), sizeof(cuuint64_t) * (rank - 1));
  key.append(reinterpret_cast<const char*>(box), sizeof(cuuint32_t) * rank);
  std::lock_guard<std::mutex> lock(g_map_mu);
  auto it = g_maps.find(key);
  if (it != g_maps.end()) {
    *out = it->second;
    return BB_OK;
  }
  EncodeTiledFn fn = encode_fn();
  if (fn == nullptr) return BB_ERR_UNSUPPORTED;
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  alignas(64) CUtensorMap m;
  const CUresult rc = fn(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(p), dims, strides, box,
                         estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) return BB_ERR_ARG;
  if (g_maps.size() > 65536) g_maps.clear();
  g_maps.emplace(std::move(key), m);
  *out = m;
  return BB_OK;
}

// ------------------------------------------------------------------------------------------------
// packs
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

// dst[o][i] = bf16(src[o*os + i*is]); one thread = 8 consecutive destination elements (16 bytes).  Up to four
// operands per launch (blockIdx.y = job): a dual product needs at most four packs, and launches -- not bytes -- are
// what a 1 us pack costs.
struct PackJob {
  const void* src;
  uint4* dst;
  int64_t os, is, outer, inner, dp8;
  int64_t batch, sbs;   // batches (>= 1) and source batch stride; destination batches are dense (outer*dp8 chunks)
  int dt, vec;          // vec: fp32, is == 1, rows 16-byte aligned, inner % 8 == 0, dp == inner
};
struct PackJobs {
  PackJob j[4];
  int n;
};

__global__ void __launch_bounds__(256) pack2d_kernel(const __grid_constant__ PackJobs J) {
  const PackJob& q = J.j[blockIdx.y];
  const int64_t per = q.outer * q.dp8, total = per * q.batch;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = t / per, tb = t - b * per;
    const int64_t o = tb / q.dp8, i0 = (tb - o * q.dp8) * 8;
    const int64_t sb = b * q.sbs;
    float v[8];
    if (q.vec) {
      const float* f = reinterpret_cast<const float*>(q.src) + sb + o * q.os + i0;
      const float4 a = bb::ld4_stream(f), b = bb::ld4_stream(f + 4);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = (i0 + e < q.inner) ? bb::ldf(q.src, sb + o * q.os + (i0 + e) * q.is, q.dt) : 0.f;
    }
    uint4 out;
    out.x = pack_bf16(v[0], v[1]); out.y = pack_bf16(v[2], v[3]); out.z = pack_bf16(v[4], v[5]); out.w = pack_bf16(v[6], v[7]);
    q.dst[t] = out;
  }
}

// NCHW -> NHWC bf16: block = 64 channels x 64 pixels of one image through shared memory
__global__ void __launch_bounds__(256) pack_nhwc_kernel(const void* __restrict__ src, int dt, int C, int HW,
                                                        __nv_bfloat16* __restrict__ dst, int Cp) {
  __shared__ float tile[64][65];
  const int img = blockIdx.z, c0 = blockIdx.y * 64, p0 = blockIdx.x * 64;
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;   // 64 x 4
  const int64_t sbase = (int64_t)img * C * HW;
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const int c = c0 + ty + k * 4, p = p0 + tx;
    tile[ty + k * 4][tx] = (c < C && p < HW) ? bb::ldf(src, sbase + (int64_t)c * HW + p, dt) : 0.f;
  }
  __syncthreads();
  // write: 32 threads cover 64 channels (2 each) of one pixel; 8 pixels per pass
  const int cx = (threadIdx.x & 31) * 2, py = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int pl = py + k * 8, p = p0 + pl;
    if (p < HW) {
      const uint32_t w = pack_bf16(tile[cx][pl], tile[cx + 1][pl]);
      *reinterpret_cast<uint32_t*>(dst + ((int64_t)img * HW + p) * Cp + c0 + cx) = w;
    }
  }
}

__global__ void __launch_bounds__(256) pack_convw_kernel(const void* __restrict__ src, int dt, int O, int C, int taps,
                                                         int transpose, __nv_bfloat16* __restrict__ dst, int Qp) {
  const int R = transpose ? C : O, Q = transpose ? O : C;
  const int64_t total = (int64_t)R * taps * Qp;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int q = (int)(t % Qp);
    const int64_t rt = t / Qp;
    const int tap = (int)(rt % taps), r = (int)(rt / taps);
    float v = 0.f;
    if (q < Q) {
      const int o = transpose ? q : r, c = transpose ? r : q;
      v = bb::ldf(src, ((int64_t)o * C + c) * taps + tap, dt);
    }
    dst[t] = __float2bfloat16(v);
  }
}

// im2col: one thread = one pixel x 8 consecutive k
__global__ void __launch_bounds__(256) pack_im2col_kernel(const void* __restrict__ src, int dt, const Im2colGeom g,
                                                          uint4* __restrict__ dst, int kp8) {
  const int64_t P = (int64_t)g.N * g.HO * g.WO, total = P * kp8;
  const int KK = g.KH * g.KW, CKK = g.C * KK;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t pix = t / kp8;
    const int k0 = (int)(t - pix * kp8) * 8;
    const int hw = g.HO * g.WO;
    const int img = (int)(pix / hw), q = (int)(pix - (int64_t)img * hw), y = q / g.WO, x = q - y * g.WO;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int k = k0 + e;
      float val = 0.f;
      if (k < CKK) {
        const int c = k / KK, r = k - c * KK, i = r / g.KW, j = r - i * g.KW;
        const int h = y * g.sh - g.ph + i * g.dh, w = x * g.sw - g.pw + j * g.dw;
        if ((unsigned)h < (unsigned)g.H && (unsigned)w < (unsigned)g.W)
          val = bb::ldf(src, (((int64_t)img * g.C + c) * g.H + h) * g.W + w, dt);
      }
      v[e] = val;
    }
    uint4 out;
    out.x = pack_bf16(v[0], v[1]); out.y = pack_bf16(v[2], v[3]); out.z = pack_bf16(v[4], v[5]); out.w = pack_bf16(v[6], v[7]);
    dst[t] = out;
  }
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
inline int64_t round8(int64_t x) { return (x + 7) / 8 * 8; }

// Packs already written to the scratch since the last bb_scratch_reset() (= by earlier launches of the same node):
// the output adjoints of a Linear feed both its input-gradient and its weight-gradient products, and a K-major view
// and its transposed MN-major view are the same bytes.
struct PackKey {
  const void* src;
  int64_t os, is, outer, inner, batch, sbs;
  int dt;
  void* dst;
};
thread_local PackKey t_cache[16];
thread_local int t_ncache = 0;
thread_local uint64_t t_cache_gen = 0;   // bb_scratch_gen the entries belong to

void* cached_pack(const PackKey& want) {
  if (t_cache_gen != bb_scratch_gen) {   // scratch was reset since
    t_ncache = 0;
    t_cache_gen = bb_scratch_gen;
  }
  for (int i = 0; i < t_ncache; ++i) {
    const PackKey& k = t_cache[i];
    if (k.src == want.src && k.dt == want.dt && k.os == want.os && k.is == want.is && k.outer == want.outer &&
        k.inner == want.inner && k.batch == want.batch && k.sbs == want.sbs)
      return k.dst;
  }
  return nullptr;
}

int queue_pack(PackJobs& J, const void* src, int dt, int64_t os, int64_t is, int64_t outer, int64_t inner, int64_t batch,
               int64_t sbs, void** dst, int64_t* dp) {
  *dp = round8(inner);
  PackKey key{src, os, is, outer, inner, batch, batch > 1 ? sbs : 0, dt, nullptr};
  if (void* hit = cached_pack(key)) {
    *dst = hit;
    return BB_OK;
  }
  void* d = bb_scratch_alloc((size_t)batch * outer * *dp * 2);
  if (!d || J.n >= 4) return BB_DECLINED;
  PackJob& q = J.j[J.n++];
  q.src = src; q.dst = reinterpret_cast<uint4*>(d); q.os = os; q.is = is; q.outer = outer; q.inner = inner;
  q.dp8 = *dp / 8; q.dt = dt; q.batch = batch; q.sbs = key.sbs;
  q.vec = (dt == BB_F32 && is == 1 && inner % 8 == 0 && os % 4 == 0 && key.sbs % 4 == 0 && aligned16(src)) ? 1 : 0;
  key.dst = d;
  if (t_ncache < 16) t_cache[t_ncache++] = key;
  *dst = d;
  return BB_OK;
}

int flush_packs(const PackJobs& J, cudaStream_t s) {
  if (J.n == 0) return BB_OK;
  int64_t most = 0;
  for (int i = 0; i < J.n; ++i) {
    const int64_t t = J.j[i].batch * J.j[i].outer * J.j[i].dp8;
    if (t > most) most = t;
  }
  int64_t blocks = (most + 255) / 256;
  if (blocks > 8 * BB_SM_COUNT) blocks = 8 * BB_SM_COUNT;
  if (blocks < 1) blocks = 1;
  pack2d_kernel<<<dim3((unsigned)blocks, (unsigned)J.n), 256, 0, s>>>(J);
  bb_launch_tally += 1;
  BB_LAUNCH_CHECK();
  return BB_OK;
}

// An operand after preparation: a bf16 matrix (per batch) whose rows are the view's rows (TMA_KMAJ, pitch >= K) or the
// view's k (TMA_MNMAJ, pitch >= R), `bstride` elements between batches.
struct Prepared {
  const void* mat;
  int64_t pitch, bstride;
  int kind;
};

inline bool batch_ok(const TmaView& v, int64_t batch) { return batch == 1 || (v.bs % 8 == 0 && v.bs > 0); }
inline bool direct_k(const TmaView& v, int64_t K, int64_t batch) {
  return v.cs == 1 && v.dt == BB_BF16 && v.rs % 8 == 0 && v.rs >= K && aligned16(v.p) && batch_ok(v, batch);
}
inline bool direct_mn(const TmaView& v, int64_t R, int64_t batch) {
  return v.rs == 1 && v.cs != 1 && v.dt == BB_BF16 && v.cs % 8 == 0 && v.cs >= R && aligned16(v.p) && batch_ok(v, batch);
}
size_t pack_bytes(const TmaView& v, int64_t R, int64_t K, int64_t batch) {
  if (v.cs == 1 || v.rs != 1) return direct_k(v, K, batch) ? 0 : (size_t)batch * R * round8(K) * 2;
  return direct_mn(v, R, batch) ? 0 : (size_t)batch * K * round8(R) * 2;
}

// Make one operand view TMA-addressable; packs are queued in J.
int prepare_operand(PackJobs& J, const TmaView& v, int64_t R, int64_t K, int64_t batch, Prepared* out) {
  void* d = nullptr;
  if (v.cs == 1 || v.rs != 1) {
    // K-major (or neither stride unit: gathered into K-major)
    out->kind = TMA_KMAJ;
    if (direct_k(v, K, batch)) {
      out->mat = v.p; out->pitch = v.rs; out->bstride = batch > 1 ? v.bs : R * v.rs;
      return BB_OK;
    }
    const int rc = queue_pack(J, v.p, v.dt, v.rs, v.cs, R, K, batch, v.bs, &d, &out->pitch);
    out->mat = d; out->bstride = R * out->pitch;
    return rc;
  }
  out->kind = TMA_MNMAJ;
  if (direct_mn(v, R, batch)) {
    out->mat = v.p; out->pitch = v.cs; out->bstride = batch > 1 ? v.bs : K * v.cs;
    return BB_OK;
  }
  const int rc = queue_pack(J, v.p, v.dt, v.cs, 1, K, R, batch, v.bs, &d, &out->pitch);
  out->mat = d; out->bstride = K * out->pitch;
  return rc;
}

}  // namespace

int bb_tma_map_3d(CUtensorMap* out, const void* p, int64_t batch, int64_t rows, int64_t cols, int64_t pitch,
                  int64_t bstride, int box_rows) {
  const cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)batch};
  const cuuint64_t strides[2] = {(cuuint64_t)pitch * 2, (cuuint64_t)bstride * 2};
  const cuuint32_t box[3] = {64u, (cuuint32_t)box_rows, 1u};
  return encode_cached(out, 3, p, dims, strides, box);
}

int bb_tma_map_2d(CUtensorMap* out, const void* p, int64_t rows, int64_t cols, int64_t pitch, int box_rows) {
  return bb_tma_map_3d(out, p, 1, rows, cols, pitch, rows * pitch, box_rows);
}

int bb_tma_map_nhwc(CUtensorMap* out, const void* p, int N, int H, int W, int Cp, int bw, int bh) {
  const cuuint64_t dims[4] = {(cuuint64_t)Cp, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  const cuuint64_t strides[3] = {(cuuint64_t)Cp * 2, (cuuint64_t)W * Cp * 2, (cuuint64_t)H * W * Cp * 2};
  const cuuint32_t box[4] = {64u, (cuuint32_t)bw, (cuuint32_t)bh, 1u};
  return encode_cached(out, 4, p, dims, strides, box);
}

int bb_tma_map_nhwc_padded(CUtensorMap* out, const void* p, int N, int H, int W, int bw, int bh) {
  const cuuint64_t Wp = (cuuint64_t)W + 2, Hp = (cuuint64_t)H + 2;
  const cuuint64_t dims[4] = {64u, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  const cuuint64_t strides[3] = {64u * 2, Wp * 64 * 2, Hp * Wp * 64 * 2};
  const cuuint32_t box[4] = {64u, (cuuint32_t)bw, (cuuint32_t)bh, 1u};
  const char* interior = reinterpret_cast<const char*>(p) + (Wp + 1) * 64 * 2;
  return encode_cached(out, 4, interior, dims, strides, box);
}

int bb_pack2d(const void* src, int dt, int64_t os, int64_t is, int64_t outer, int64_t inner, void* dst, int64_t dp,
              cudaStream_t s) {
  if (outer <= 0 || inner <= 0) return BB_OK;
  PackJobs J{};
  PackJob& q = J.j[0];
  J.n = 1;
  q.src = src; q.dst = reinterpret_cast<uint4*>(dst); q.os = os; q.is = is; q.outer = outer; q.inner = inner;
  q.dp8 = dp / 8; q.dt = dt; q.batch = 1; q.sbs = 0;
  q.vec = (dt == BB_F32 && is == 1 && inner % 8 == 0 && dp == inner && os % 4 == 0 && aligned16(src)) ? 1 : 0;
  return flush_packs(J, s);
}

int bb_pack_nhwc(const void* src, int dt, int N, int C, int HW, void* dst, int Cp, cudaStream_t s) {
  dim3 grid((unsigned)((HW + 63) / 64), (unsigned)(Cp / 64), (unsigned)N);
  pack_nhwc_kernel<<<grid, 256, 0, s>>>(src, dt, C, HW, reinterpret_cast<__nv_bfloat16*>(dst), Cp);
  bb_launch_tally += 1;
  BB_LAUNCH_CHECK();
  return BB_OK;
}

int bb_pack_convw(const void* src, int dt, int O, int C, int taps, int transpose, void* dst, int Qp, cudaStream_t s) {
  const int64_t total = (int64_t)(transpose ? C : O) * taps * Qp;
  int64_t blocks = (total + 255) / 256;
  if (blocks > 4 * BB_SM_COUNT) blocks = 4 * BB_SM_COUNT;
  pack_convw_kernel<<<(unsigned)blocks, 256, 0, s>>>(src, dt, O, C, taps, transpose, reinterpret_cast<__nv_bfloat16*>(dst), Qp);
  bb_launch_tally += 1;
  BB_LAUNCH_CHECK();
  return BB_OK;
}

int bb_pack_im2col(const void* src, int dt, const Im2colGeom& g, void* dst, int kp, cudaStream_t s) {
  const int64_t total = (int64_t)g.N * g.HO * g.WO * (kp / 8);
  int64_t blocks = (total + 255) / 256;
  if (blocks > 16 * BB_SM_COUNT) blocks = 16 * BB_SM_COUNT;
  if (blocks < 1) blocks = 1;
  pack_im2col_kernel<<<(unsigned)blocks, 256, 0, s>>>(src, dt, g, reinterpret_cast<uint4*>(dst), kp / 8);
  bb_launch_tally += 1;
  BB_LAUNCH_CHECK();
  return BB_OK;
}

int bb_gemm_tma_launch(TmaGemmArgs& G, int bn, int64_t mtiles, cudaStream_t s, int batch) {
  return bn == 64 ? launch_tma<64>(G, mtiles, s, batch) : launch_tma<128>(G, mtiles, s, batch);
}

int bb_gemm_tma_prepack(const TmaPackReq* reqs, int n, int64_t batch, cudaStream_t s) {
  if (encode_fn() == nullptr || bb_scratch.base == nullptr) return BB_OK;
  PackJobs J{};
  for (int i = 0; i < n; ++i) {
    const size_t need = pack_bytes(reqs[i].v, reqs[i].rows, reqs[i].k, batch);
    if (need == 0) continue;
    if (bb_scratch.used + need + 512 > bb_scratch.bytes) break;   // the products will decline or pack on their own
    Prepared pr;
    if (prepare_operand(J, reqs[i].v, reqs[i].rows, reqs[i].k, batch, &pr) != BB_OK) break;
    if (J.n == 4) {
      const int rc = flush_packs(J, s);
      if (rc) return rc;
      J.n = 0;
    }
  }
  return flush_packs(J, s);
}

int bb_gemm_tma_run(int64_t M, int64_t N, int64_t K, int npairs, const TmaView* A, const TmaView* B, float* out,
                    int64_t ors, int64_t ocs, int beta, const float* bias, int64_t bias_stride, bool out_dense,
                    cudaStream_t s, int plane_ohw, int min_n, int64_t batch, int64_t obs) {
  if (npairs < 1 || npairs > 2 || batch < 1) return BB_DECLINED;
  if (batch == 1) {
    if (M < 64 || N < min_n || K < 8) return BB_DECLINED;
    if (K < 64 && min_n >= 64) return BB_DECLINED;
  } else {
    // batched (attention score / context products): the tiles are mostly padding, but a 128 x 64 tcgen05 tile costs
    // less than the SIMT kernel's inner loop as soon as the products are not tiny
    if (M < 32 || N < 32 || K < 32 || batch > 65535 || plane_ohw > 0 || bias != nullptr) return BB_DECLINED;
  }
  if (M > INT32_MAX / 2 || N > INT32_MAX / 2 || K > INT32_MAX / 2) return BB_DECLINED;
  if (encode_fn() == nullptr || bb_scratch.base == nullptr) return BB_DECLINED;
  // scratch admission: bytes of the operands that need a pack (plan.py sizes the scratch with an upper bound of this)
  {
    size_t need = 1024;
    for (int p = 0; p < npairs; ++p) need += pack_bytes(A[p], M, K, batch) + pack_bytes(B[p], N, K, batch) + 512;
    if (bb_scratch.used + need > bb_scratch.bytes) return BB_DECLINED;
  }
  alignas(64) TmaGemmArgs G;
  memset(&G, 0, sizeof(G));
  // Tile width and split-K by a cost model.  The kernel is bound by what one SM can pull through the TMA unit
  // (measured ~43 B/clk/SM = the chip's ~6300 B/clk L2 throughput / 148; a 128 x bn x 64 k-block costs 16 + bn/8 KB),
  // so the time of a configuration is (waves of CTAs) x (k-blocks per CTA) x (bytes per k-block) / per-SM rate; split-K
  // adds a memset launch and the reduction traffic of its partial tiles.
  const int64_t mtiles = (M + BM - 1) / BM;
  const int64_t kblocks = (K + BK - 1) / BK;
  const bool can_split = batch == 1 && plane_ohw == 0 && (beta || out_dense) && kblocks >= 8;
  int bn = 64, ksplit = 1;
  {
    double best = 1e30;
    const int bn_opts[2] = {128, 64};
    for (int bi = 0; bi < 2; ++bi) {
      const int cand = bn_opts[bi];
      if (cand == 128 && N <= 64) continue;
      const int64_t tiles = mtiles * ((N + cand - 1) / cand) * batch;
      for (int sp = 1; sp <= 256; sp = sp < 4 ? sp + 1 : sp * 2) {
        if (sp > 1 && (!can_split || kblocks / sp < 4)) break;
        const int64_t ctas = tiles * sp;
        const int64_t waves = (ctas + BB_SM_COUNT - 1) / BB_SM_COUNT;
        const int64_t iters = (kblocks + sp - 1) / sp * npairs;
        double us = (double)waves * (double)iters * (16.0 + cand / 8.0) * 1024.0 / 84e3 + 2.0;   // 84 GB/s per SM
        if (sp > 1) us += (beta ? 0.0 : 2.5) + (double)M * (double)N * 4.0 * sp / 2.0e6;          // ~2 TB/s of reductions
        if (us < best) {
          best = us;
          bn = cand;
          ksplit = sp;
        }
      }
    }
  }
  PackJobs J{};
  Prepared pa[2], pb[2];
  for (int p = 0; p < npairs; ++p) {
    int rc = prepare_operand(J, A[p], M, K, batch, &pa[p]);
    if (rc) return rc;
    rc = prepare_operand(J, B[p], N, K, batch, &pb[p]);
    if (rc) return rc;
  }
  {
    const int rc = flush_packs(J, s);
    if (rc) return rc;
  }
  for (int p = 0; p < npairs; ++p) {
    G.a_kind[p] = pa[p].kind;
    G.b_kind[p] = pb[p].kind;
    int rc = pa[p].kind == TMA_KMAJ ? bb_tma_map_3d(&G.a[p], pa[p].mat, batch, M, K, pa[p].pitch, pa[p].bstride, BM)
                                    : bb_tma_map_3d(&G.a[p], pa[p].mat, batch, K, M, pa[p].pitch, pa[p].bstride, 64);
    if (rc) return rc;
    rc = pb[p].kind == TMA_KMAJ ? bb_tma_map_3d(&G.b[p], pb[p].mat, batch, N, K, pb[p].pitch, pb[p].bstride, bn)
                                : bb_tma_map_3d(&G.b[p], pb[p].mat, batch, K, N, pb[p].pitch, pb[p].bstride, 64);
    if (rc) return rc;
  }
  G.M = M; G.N = N; G.K = K; G.npairs = npairs;
  G.a_bytes = A_TILE; G.b_bytes = (uint32_t)bn * BK * 2;
  G.out = out; G.omode = 0; G.ors = ors; G.ocs = ocs; G.obs = obs; G.beta = beta; G.bias = bias; G.bias_stride = bias_stride;
  if (plane_ohw > 0) {
    G.omode = 2; G.OCH = (int)N; G.OHW = plane_ohw;
  }
  if (ksplit > 1 && !beta) {
    BB_CUDA_TRY(cudaMemsetAsync(out, 0, sizeof(float) * M * N, s));
    bb_launch_tally += 1;
  }
  G.ksplit = ksplit;
  return bb_gemm_tma_launch(G, bn, mtiles, s, (int)batch);
}

extern "C" int bb_gemm_bf16_tma(int64_t M, int64_t N, int64_t K, const void* A, int dtA, int64_t ars, int64_t acs,
                                const void* B, int dtB, int64_t brs, int64_t bcs, float* C, int64_t crs, int64_t ccs,
                                int beta, void* scratch, int64_t scratch_bytes, void* stream) {
  const BbScratch saved = bb_scratch;
  bb_scratch = BbScratch{reinterpret_cast<uint8_t*>(scratch), (size_t)scratch_bytes, 0};
  bb_scratch_reset();
  const TmaView a{A, dtA, ars, acs}, b{B, dtB, bcs, brs};   // rows of the B view = n
  const bool dense = (ccs == 1 && crs == N) || (crs == 1 && ccs == M);
  const int rc = bb_gemm_tma_run(M, N, K, 1, &a, &b, C, crs, ccs, beta, nullptr, 0, dense, (cudaStream_t)stream);
  bb_scratch = saved;
  bb_scratch_reset();
  return rc;
}
