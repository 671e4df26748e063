// TMA-fed tensor-core path of the Conv2d second-order rules (conv_tma.cu).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

#include "plan.h"

// node shape / pass this path handles (stride 1, <= 9 taps, 32..64 channels each way, bf16-autocast graph, scratch set)
bool bb_conv_tma_ok(const bb_node& nd, int pass);
// base-backward pass hook: packs the node's K-loop constants into the plan's persistent arena (no-op otherwise)
int bb_conv_tma_prepare(const bb_node& nd, cudaStream_t s);
// scratch bytes the node's packs need (plan.py mirrors this bound)
size_t bb_conv_tma_scratch(const bb_node& nd);
// TF: writes t_y.  TB: writes at_x and at_W (the bias adjoint stays with the caller).  BB_DECLINED: nothing usable was
// produced, caller takes its other path.
int bb_conv_tma_run(const bb_node& nd, int pass, cudaStream_t s);
