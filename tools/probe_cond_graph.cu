#include <cuda_runtime.h>
#include <cstdio>
__global__ void body(int* c, cudaGraphConditionalHandle h) {
    int v = atomicAdd(c, 1);
    if (threadIdx.x == 0) cudaGraphSetConditional(h, v < 5 ? 1 : 0);
}
int main() {
    cudaStream_t st; cudaStreamCreate(&st);
    int* c; cudaMalloc(&c, 4); cudaMemset(c, 0, 4);
    cudaGraph_t g; cudaGraphCreate(&g, 0);
    cudaGraphConditionalHandle h;
    cudaGraphConditionalHandleCreate(&h, g, 1, cudaGraphCondAssignDefault);
    cudaGraphNodeParams p = {};
    p.type = cudaGraphNodeTypeConditional;
    p.conditional.handle = h; p.conditional.type = cudaGraphCondTypeWhile; p.conditional.size = 1;
    cudaGraphNode_t node;
    printf("add %d\n", cudaGraphAddNode(&node, g, nullptr, 0, &p));
    cudaGraph_t bodyg = p.conditional.phGraph_out[0];
    printf("begin %d\n", cudaStreamBeginCaptureToGraph(st, bodyg, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
    body<<<1, 1, 0, st>>>(c, h);
    cudaGraph_t out; printf("end %d\n", cudaStreamEndCapture(st, &out));
    cudaGraphExec_t ex; printf("inst %d\n", cudaGraphInstantiate(&ex, g, 0));
    cudaGraphLaunch(ex, st); cudaStreamSynchronize(st);
    int hc; cudaMemcpy(&hc, c, 4, cudaMemcpyDeviceToHost); printf("count %d\n", hc);
}
