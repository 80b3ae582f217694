// Early exit of a captured decode loop (modeling_utils.py:865-867 `if cur_unfinished.max() == 0: break`, :1071-1073
// `if all(done): break`) WITHOUT a host round trip: while a stream is being captured into a CUDA graph, every decode step is
// recorded as the body of a conditional IF node whose condition a one-block kernel evaluates on the device from the search
// state ("is any sequence still unfinished?"). Once every caption of the batch has ended, the remaining steps of the replayed
// graph cost one tiny kernel each instead of ~33.
//
//   graph_if_any_begin(flags, n, invert, capture_stream, body_stream)
//       capture_stream is capturing (torch.cuda.graph): appends [set-condition kernel] -> [IF node] to its graph, makes the IF
//       node the stream's capture dependency, and starts capturing body_stream INTO the IF node's body graph. The caller then
//       launches the step's kernels on body_stream.
//   graph_if_end(body_stream)
//       ends the body capture; capture_stream continues after the IF node.
#include "common.cuh"

namespace vc {

namespace {
// condition = any(flags[i] != 0) (invert = 0: `unfinished` per sequence) or any(flags[i] == 0) (invert = 1: `done` per image)
__global__ void __launch_bounds__(256) set_cond_any_kernel(cudaGraphConditionalHandle h, const int* __restrict__ flags, int n, int invert) {
  int any = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) any |= invert ? (flags[i] == 0) : (flags[i] != 0);
  any = __syncthreads_or(any);
  if (threadIdx.x == 0) cudaGraphSetConditional(h, any ? 1u : 0u);
}

int fail(const char* what, cudaError_t e) {
  set_last_error("graph_if: %s: %s", what, cudaGetErrorString(e));
  cudaGetLastError();
  return VC_ERR_LAUNCH;
}
}  // namespace

int graph_if_any_begin(const int* flags, int n, int invert, cudaStream_t cap, cudaStream_t body) {
  if (flags == nullptr || n <= 0 || cap == body) { set_last_error("graph_if_any_begin: bad args"); return VC_ERR_BAD_ARG; }
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  cudaGraph_t g = nullptr;
  const cudaGraphNode_t* deps = nullptr;
  size_t nd = 0;
  cudaError_t e = cudaStreamGetCaptureInfo(cap, &st, nullptr, &g, &deps, &nd);
  if (e != cudaSuccess) return fail("cudaStreamGetCaptureInfo", e);
  if (st != cudaStreamCaptureStatusActive || g == nullptr) { set_last_error("graph_if_any_begin: the stream is not capturing"); return VC_ERR_BAD_ARG; }
  cudaGraphConditionalHandle h;
  e = cudaGraphConditionalHandleCreate(&h, g, 1, cudaGraphCondAssignDefault);
  if (e != cudaSuccess) return fail("cudaGraphConditionalHandleCreate", e);
  set_cond_any_kernel<<<1, 256, 0, cap>>>(h, flags, n, invert);              // captured: a kernel node of g
  e = cudaGetLastError();
  if (e != cudaSuccess) return fail("set_cond_any_kernel", e);
  e = cudaStreamGetCaptureInfo(cap, &st, nullptr, &g, &deps, &nd);         // the dependencies now end in that kernel node
  if (e != cudaSuccess) return fail("cudaStreamGetCaptureInfo", e);
  cudaGraphNodeParams p = {};
  p.type = cudaGraphNodeTypeConditional;
  p.conditional.handle = h;
  p.conditional.type = cudaGraphCondTypeIf;
  p.conditional.size = 1;
  cudaGraphNode_t node;
  e = cudaGraphAddNode(&node, g, deps, nd, &p);
  if (e != cudaSuccess) return fail("cudaGraphAddNode(conditional)", e);
  cudaGraph_t body_graph = p.conditional.phGraph_out[0];
  e = cudaStreamUpdateCaptureDependencies(cap, &node, 1, cudaStreamSetCaptureDependencies);
  if (e != cudaSuccess) return fail("cudaStreamUpdateCaptureDependencies", e);
  e = cudaStreamBeginCaptureToGraph(body, body_graph, nullptr, nullptr, 0, cudaStreamCaptureModeGlobal);
  if (e != cudaSuccess) return fail("cudaStreamBeginCaptureToGraph", e);
  return VC_OK;
}

int graph_if_end(cudaStream_t body) {
  cudaGraph_t g = nullptr;
  cudaError_t e = cudaStreamEndCapture(body, &g);
  if (e != cudaSuccess) return fail("cudaStreamEndCapture(body)", e);
  return VC_OK;
}

}  // namespace vc
