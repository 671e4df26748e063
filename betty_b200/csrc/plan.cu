// Executor of the second-order tape: walks the node list (forward for tangents, reverse for adjoints),
// and drives whole Neumann / CG K-loops natively so Python is out of the loop: one iteration
// (tangent forward, tangent backward, vector update kernels) is captured into a CUDA graph and replayed.
// Replaces reference neumann.py:61-64 and cg.py:38-55 (Python loops around autograd.grad).
#include <vector>

#include "../../include/betty_b200.h"
#include "bb_common.cuh"
#include "plan.h"
#include "tma.h"

struct bb_plan {
  std::vector<bb_node> nodes;
  std::vector<void*> zero_ptr[3];
  std::vector<int64_t> zero_bytes[3];
  int launches[3] = {0, 0, 0};
  void* scratch = nullptr;      // bf16 operand packs of the TMA-fed tensor-core path (tma.h)
  int64_t scratch_bytes = 0;
  uint8_t* persist = nullptr;   // plan-lifetime packs of K-loop constants, bump-allocated per node
  int64_t persist_bytes = 0, persist_used = 0;
  std::vector<void*> node_persist;
};

namespace {
thread_local bb_plan* t_plan = nullptr;
thread_local int t_node = -1;
}  // namespace

void* bb_persist_get(size_t bytes, bool* fresh, int slot_id) {
  *fresh = false;
  bb_plan* p = t_plan;
  if (!p || !p->persist || t_node < 0 || slot_id < 0 || slot_id >= BB_PERSIST_SLOTS) return nullptr;
  const size_t at_slot = (size_t)t_node * BB_PERSIST_SLOTS + slot_id;
  if (at_slot >= p->node_persist.size()) return nullptr;
  void*& slot = p->node_persist[at_slot];
  if (slot) return slot;
  const int64_t at = (p->persist_used + 255) & ~(int64_t)255;
  if (at + (int64_t)bytes > p->persist_bytes) return nullptr;
  slot = p->persist + at;
  p->persist_used = at + (int64_t)bytes;
  *fresh = true;
  return slot;
}

namespace {

int dispatch(const bb_node& nd, int pass, cudaStream_t s) {
  switch (nd.op) {
    case BB_OP_UNARY:
    case BB_OP_COPY:
    case BB_OP_ADD2:
    case BB_OP_MULC:
    case BB_OP_MUL2:
      return bb_launch_ew(nd, pass, s);
    case BB_OP_SUMALL:
      return bb_launch_sumall(nd, pass, s);
    case BB_OP_GEMM:
      return bb_launch_gemm(nd, pass, s);
    case BB_OP_CONV2D:
      return bb_launch_conv2d(nd, pass, s);
    case BB_OP_MAXPOOL2D:
      return bb_launch_maxpool2d(nd, pass, s);
    case BB_OP_BATCHNORM:
      return bb_launch_batchnorm(nd, pass, s);
    case BB_OP_LAYERNORM:
      return bb_launch_layernorm(nd, pass, s);
    case BB_OP_SOFTMAX:
    case BB_OP_LOGSOFTMAX:
      return bb_launch_softmax(nd, pass, s);
    case BB_OP_NLL:
      return bb_launch_nll(nd, pass, s);
    case BB_OP_BCE_LOGITS:
      return bb_launch_bce(nd, pass, s);
    case BB_OP_EMBEDDING:
      return bb_launch_embedding(nd, pass, s);
    case BB_OP_AVGPOOL2D:
      return bb_launch_avgpool2d(nd, pass, s);
    case BB_OP_DIAGSHIFT: {
      if (pass != BB_PASS_TAN_BWD) return BB_OK;
      bb_launch_tally += 1;
      return bb_mt_axpby(reinterpret_cast<const bb_mt_chunk*>(nd.aux[0]), (int)nd.dims[0], (float)nd.f[0], nullptr, 1.0f,
                         (void*)s);
    }
    default:
      return BB_ERR_UNSUPPORTED;
  }
}

int run_pass(bb_plan* p, int pass, cudaStream_t s) {
  if (pass < 0 || pass > 2) return BB_ERR_ARG;
  bb_scratch = BbScratch{reinterpret_cast<uint8_t*>(p->scratch), (size_t)p->scratch_bytes, 0};
  bb_scratch_reset();
  const int tally0 = bb_launch_tally;
  for (size_t i = 0; i < p->zero_ptr[pass].size(); ++i) {
    BB_CUDA_TRY(cudaMemsetAsync(p->zero_ptr[pass][i], 0, (size_t)p->zero_bytes[pass][i], s));
    bb_launch_tally += 1;
  }
  const int n = (int)p->nodes.size();
  t_plan = p;
  if (pass == BB_PASS_TAN_FWD) {
    for (int i = 0; i < n; ++i) {
      t_node = i;
      const int rc = dispatch(p->nodes[i], pass, s);
      if (rc) return rc;
    }
  } else {
    for (int i = n - 1; i >= 0; --i) {
      t_node = i;
      const int rc = dispatch(p->nodes[i], pass, s);
      if (rc) return rc;
    }
  }
  t_node = -1;
  p->launches[pass] = bb_launch_tally - tally0;
  return BB_OK;
}

template <class Body>
int loop_with_graph(int iterations, int use_graph, cudaStream_t s, Body body) {
  if (iterations <= 0) return BB_OK;
  if (!use_graph || iterations == 1) {
    for (int k = 0; k < iterations; ++k) {
      const int rc = body();
      if (rc) return rc;
    }
    return BB_OK;
  }
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  BB_CUDA_TRY(cudaStreamBeginCapture(s, cudaStreamCaptureModeRelaxed));
  const int rc = body();
  cudaError_t e = cudaStreamEndCapture(s, &graph);
  if (rc) {
    if (graph) cudaGraphDestroy(graph);
    return rc;
  }
  if (e != cudaSuccess) return (int)e;
  e = cudaGraphInstantiate(&exec, graph, 0);
  if (e != cudaSuccess) {
    cudaGraphDestroy(graph);
    return (int)e;
  }
  int out = BB_OK;
  for (int k = 0; k < iterations; ++k) {
    e = cudaGraphLaunch(exec, s);
    if (e != cudaSuccess) {
      out = (int)e;
      break;
    }
  }
  cudaGraphExecDestroy(exec);
  cudaGraphDestroy(graph);
  return out;
}

}  // namespace

extern "C" {

int bb_node_bytes(void) { return (int)sizeof(bb_node); }

int bb_plan_create(const struct bb_node* nodes, int n_nodes, bb_plan** out) {
  if (!out || n_nodes < 0 || (n_nodes > 0 && !nodes)) return BB_ERR_ARG;
  bb_plan* p = new bb_plan();
  p->nodes.assign(nodes, nodes + n_nodes);
  p->node_persist.assign((size_t)n_nodes * BB_PERSIST_SLOTS, nullptr);
  *out = p;
  return BB_OK;
}

int bb_plan_destroy(bb_plan* plan) {
  delete plan;
  return BB_OK;
}

int bb_plan_set_zero_regions(bb_plan* plan, int pass, void* const* ptrs, const int64_t* bytes, int n) {
  if (!plan || pass < 0 || pass > 2) return BB_ERR_ARG;
  plan->zero_ptr[pass].assign(ptrs, ptrs + n);
  plan->zero_bytes[pass].assign(bytes, bytes + n);
  return BB_OK;
}

int bb_plan_set_scratch(bb_plan* plan, void* ptr, int64_t bytes) {
  if (!plan || bytes < 0) return BB_ERR_ARG;
  plan->scratch = ptr;
  plan->scratch_bytes = bytes;
  return BB_OK;
}

int bb_plan_set_persistent(bb_plan* plan, void* ptr, int64_t bytes) {
  if (!plan || bytes < 0) return BB_ERR_ARG;
  plan->persist = reinterpret_cast<uint8_t*>(ptr);
  plan->persist_bytes = bytes;
  plan->persist_used = 0;
  plan->node_persist.assign(plan->nodes.size() * BB_PERSIST_SLOTS, nullptr);
  return BB_OK;
}

int bb_plan_run(bb_plan* plan, int pass, void* stream) {
  if (!plan) return BB_ERR_ARG;
  return run_pass(plan, pass, (cudaStream_t)stream);
}

int bb_plan_launch_count(const bb_plan* plan, int pass) {
  if (!plan || pass < 0 || pass > 2) return BB_ERR_ARG;
  return plan->launches[pass];
}

int bb_plan_hvp(bb_plan* plan, void* stream) {
  if (!plan) return BB_ERR_ARG;
  int rc = run_pass(plan, BB_PASS_TAN_FWD, (cudaStream_t)stream);
  if (rc) return rc;
  return run_pass(plan, BB_PASS_TAN_BWD, (cudaStream_t)stream);
}

int bb_plan_profile(bb_plan* plan, int pass, float* ms_per_node, void* stream) {
  if (!plan || pass < 0 || pass > 2 || !ms_per_node) return BB_ERR_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  bb_scratch = BbScratch{reinterpret_cast<uint8_t*>(plan->scratch), (size_t)plan->scratch_bytes, 0};
  bb_scratch_reset();
  const int n = (int)plan->nodes.size();
  std::vector<cudaEvent_t> ev(n + 1);
  for (auto& e : ev) BB_CUDA_TRY(cudaEventCreate(&e));
  for (size_t i = 0; i < plan->zero_ptr[pass].size(); ++i)
    BB_CUDA_TRY(cudaMemsetAsync(plan->zero_ptr[pass][i], 0, (size_t)plan->zero_bytes[pass][i], s));
  int rc = BB_OK;
  for (int j = 0; j < n && !rc; ++j) {
    const int i = (pass == BB_PASS_TAN_FWD) ? j : n - 1 - j;
    cudaEventRecord(ev[j], s);
    t_plan = plan;
    t_node = i;
    rc = dispatch(plan->nodes[i], pass, s);
    ms_per_node[i] = 0.f;
  }
  cudaEventRecord(ev[n], s);
  cudaStreamSynchronize(s);
  if (!rc) {
    for (int j = 0; j < n; ++j) {
      const int i = (pass == BB_PASS_TAN_FWD) ? j : n - 1 - j;
      cudaEventElapsedTime(&ms_per_node[i], ev[j], ev[j + 1]);
    }
  }
  for (auto& e : ev) cudaEventDestroy(e);
  return rc;
}

int bb_plan_neumann_loop(bb_plan* plan, int iterations, float alpha, float* v, float* p, const float* hv, int64_t n,
                         int use_graph, void* stream) {
  if (!plan) return BB_ERR_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  return loop_with_graph(iterations, use_graph, s, [&]() -> int {
    int rc = bb_plan_hvp(plan, s);                            // hv <- H v        (neumann.py:62)
    if (rc) return rc;
    return bb_neumann_update(v, p, hv, alpha, 0.f, n, s);     // v, p updates     (neumann.py:63-64)
  });
}

int bb_plan_cg_loop(bb_plan* plan, int iterations, float cg_alpha, float* x, float* r, float* p, const float* hp,
                    int64_t n, void* ws, int use_graph, void* stream) {
  if (!plan) return BB_ERR_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  int rc = bb_cg_init(r, n, ws, s);                           // rr = r.r         (cg.py:45, first iteration)
  if (rc) return rc;
  return loop_with_graph(iterations, use_graph, s, [&]() -> int {
    int rc2 = bb_plan_hvp(plan, s);                           // hp <- H p        (cg.py:39-41)
    if (rc2) return rc2;
    rc2 = bb_cg_dots(r, hp, p, cg_alpha, 0, n, ws, s);        // alpha            (cg.py:42-47)
    if (rc2) return rc2;
    rc2 = bb_cg_update_xr(x, r, p, hp, n, ws, s);             // x, r, beta       (cg.py:49-52)
    if (rc2) return rc2;
    return bb_cg_update_p(p, r, n, ws, s);                    // p                (cg.py:53)
  });
}

const char* bb_version(void) { return "betty_b200 0.1.0 (sm_100a)"; }

}  // extern "C"
