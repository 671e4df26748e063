// Argument block of the tcgen05 dual-product GEMM (gemm_tc.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

struct TcGemmArgs {
  int64_t M, N, K;
  int npairs, ksplit;
  const void* a[2];          // A_p[m][k] = a[p][m*ars + k*acs]
  int dta[2];
  int64_t ars[2], acs[2];
  int a_kfast[2];
  const void* b[2];          // B_p[k][n] = b[p][k*brs + n*bcs]
  int dtb[2];
  int64_t brs[2], bcs[2];
  int b_kfast[2];
  float* out;                // D[m][n] = out[m*ors + n*ocs]
  int64_t ors, ocs;
  int beta;
  const float* bias;
  int64_t bias_stride;
  int allow_split, out_dense;
};

bool bb_gemm_tc_eligible(int64_t M, int64_t N, int64_t K, int64_t batch);
int bb_gemm_tc_run(const TcGemmArgs& G, cudaStream_t s);
