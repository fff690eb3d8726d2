// Feasibility + cost probe: a CUDA-graph WHILE node whose body is a chain of 7 small kernels, the loop condition set on the
// device (cudaGraphSetConditional) - the shape of the host-free wave loop.  Prints microseconds per iteration.
#include <cuda_runtime.h>
#include <stdio.h>
__global__ void k_body(int* ctr, int which) { if (threadIdx.x == 0 && blockIdx.x == 0) atomicAdd(&ctr[which], 1); }
__global__ void k_cond(int* ctr, int n_iter, cudaGraphConditionalHandle h) {
  if (threadIdx.x == 0 && blockIdx.x == 0) { int it = atomicAdd(&ctr[7], 1) + 1; cudaGraphSetConditional(h, it < n_iter ? 1 : 0); }
}
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s -> %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
int main(int argc, char** argv) {
  int n_iter = argc > 1 ? atoi(argv[1]) : 1000, grid = argc > 2 ? atoi(argv[2]) : 148 * 4;
  int* ctr; CK(cudaMalloc(&ctr, 64)); CK(cudaMemset(ctr, 0, 64));
  cudaStream_t s; CK(cudaStreamCreate(&s));
  cudaGraph_t g; CK(cudaGraphCreate(&g, 0));
  cudaGraphConditionalHandle h; CK(cudaGraphConditionalHandleCreate(&h, g, 1, cudaGraphCondAssignDefault));
  cudaGraphNodeParams p = {}; p.type = cudaGraphNodeTypeConditional; p.conditional.handle = h; p.conditional.type = cudaGraphCondTypeWhile; p.conditional.size = 1;
  cudaGraphNode_t node; CK(cudaGraphAddNode(&node, g, nullptr, 0, &p));
  cudaGraph_t body = p.conditional.phGraph_out[0];
  // body built by stream capture into the body graph
  CK(cudaStreamBeginCaptureToGraph(s, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
  for (int k = 0; k < 7; ++k) k_body<<<grid, 128, 0, s>>>(ctr, k);
  k_cond<<<1, 32, 0, s>>>(ctr, n_iter, h);
  CK(cudaStreamEndCapture(s, nullptr));
  cudaGraphExec_t ex; CK(cudaGraphInstantiate(&ex, g, 0));
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int rep = 0; rep < 3; ++rep) {
    CK(cudaMemsetAsync(ctr, 0, 64, s));
    cudaEventRecord(a, s); CK(cudaGraphLaunch(ex, s)); cudaEventRecord(b, s); CK(cudaStreamSynchronize(s));
    float ms; cudaEventElapsedTime(&ms, a, b);
    int hc[8]; cudaMemcpy(hc, ctr, 32, cudaMemcpyDeviceToHost);
    printf("graph while: %d iterations x 8 kernels (grid %d): %.3f ms = %.2f us / iteration (body kernel ran %d times)\n", hc[7], grid, ms, 1e3 * ms / hc[7], hc[0]);
  }
  // the same chain as plain stream launches, one host synchronise per iteration (today's growth phase) and none (lookahead)
  for (int mode = 0; mode < 2; ++mode) {
    cudaEventRecord(a, s);
    for (int it = 0; it < n_iter; ++it) { for (int k = 0; k < 7; ++k) k_body<<<grid, 128, 0, s>>>(ctr, k); k_body<<<1, 32, 0, s>>>(ctr, 7); if (mode == 0) cudaStreamSynchronize(s); }
    cudaEventRecord(b, s); cudaStreamSynchronize(s);
    float ms; cudaEventElapsedTime(&ms, a, b);
    printf("stream launches, %s: %.2f us / iteration\n", mode == 0 ? "sync per iteration" : "no sync", 1e3 * ms / n_iter);
  }
  return 0;
}
