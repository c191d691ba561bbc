// Does kernel execution priority let a chain of short kernels run INSIDE a long, GPU-filling kernel on this GPU?
//   nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/priority_probe tools/priority_probe.cu && /tmp/priority_probe
// Stream A: one "bulk" kernel (15000 CTAs x 256 threads, 8 resident per SM = every thread slot, ~30 us per CTA).  Stream B: 15 dependent
// short kernels (326 CTAs, ~10 us each).  Reported: when the chain finishes and when the bulk kernel finishes, both
// measured from the start of the bulk kernel, for
//   plain      both streams at default priority
//   streams    B on a high-priority stream, A on a low-priority one
//   attribute  both streams high priority, the bulk kernel launched with cudaLaunchAttributePriority = lowest
//   graph      the 'attribute' schedule captured (fork / join) from a high-priority stream and replayed, instantiated
//              by default and with cudaGraphInstantiateFlagUseNodePriority
#include <cuda_runtime.h>
#include <cstdio>

__global__ void __launch_bounds__(256, 4) spin_kernel(long long cycles, int* sink) {
  const long long t0 = clock64();
  while (clock64() - t0 < cycles) {
  }
  if (sink && threadIdx.x == 0 && blockIdx.x == 0x7fffffff) *sink = 1;
}

static void launch(cudaStream_t s, int grid, long long cycles, bool low_attr, int low) {
  if (!low_attr) {
    spin_kernel<<<grid, 256, 0, s>>>(cycles, nullptr);
    return;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(256);
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributePriority;
  at[0].val.priority = low;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, spin_kernel, cycles, (int*)nullptr);
}

int main() {
  int low = 0, high = 0;
  cudaDeviceGetStreamPriorityRange(&low, &high);
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const long long bulk_cycles = (long long)khz * 30 / 1000, small_cycles = (long long)khz * 10 / 1000;
  printf("priority range: lowest %d, highest %d; clock %d kHz\n", low, high, khz);
  cudaStream_t def_a, def_b, lo_a, hi_a, hi_b, hi_cap;
  cudaStreamCreateWithPriority(&def_a, cudaStreamNonBlocking, low);
  cudaStreamCreateWithPriority(&def_b, cudaStreamNonBlocking, low);
  cudaStreamCreateWithPriority(&lo_a, cudaStreamNonBlocking, low);
  cudaStreamCreateWithPriority(&hi_a, cudaStreamNonBlocking, high);
  cudaStreamCreateWithPriority(&hi_b, cudaStreamNonBlocking, high);
  cudaStreamCreateWithPriority(&hi_cap, cudaStreamNonBlocking, high);
  cudaEvent_t e0, ea, eb, fork, join;
  cudaEventCreate(&e0); cudaEventCreate(&ea); cudaEventCreate(&eb);
  cudaEventCreateWithFlags(&fork, cudaEventDisableTiming); cudaEventCreateWithFlags(&join, cudaEventDisableTiming);

  auto run = [&](const char* name, cudaStream_t a, cudaStream_t b, bool low_attr) {
    for (int rep = 0; rep < 3; ++rep) {
      cudaDeviceSynchronize();
      cudaEventRecord(e0, a);
      cudaEventRecord(fork, a);
      cudaStreamWaitEvent(b, fork, 0);
      launch(a, 15000, bulk_cycles, low_attr, low);
      for (int k = 0; k < 15; ++k) launch(b, 326, small_cycles, false, low);
      cudaEventRecord(ea, a);
      cudaEventRecord(eb, b);
      cudaDeviceSynchronize();
      float ta = 0, tb = 0;
      cudaEventElapsedTime(&ta, e0, ea);
      cudaEventElapsedTime(&tb, e0, eb);
      if (rep == 2) printf("%-10s chain done at %.3f ms, bulk done at %.3f ms\n", name, tb, ta);
    }
  };
  // baselines: each alone
  for (int rep = 0; rep < 2; ++rep) {
    cudaDeviceSynchronize();
    cudaEventRecord(e0, def_a);
    launch(def_a, 15000, bulk_cycles, false, low);
    cudaEventRecord(ea, def_a);
    cudaEventRecord(eb, def_a);
    for (int k = 0; k < 15; ++k) launch(def_a, 326, small_cycles, false, low);
    cudaEventRecord(fork, def_a);
    cudaEvent_t ec;
    cudaEventCreate(&ec);
    cudaEventRecord(ec, def_a);
    cudaDeviceSynchronize();
    float ta = 0, tc = 0;
    cudaEventElapsedTime(&ta, e0, ea);
    cudaEventElapsedTime(&tc, eb, ec);
    if (rep == 1) printf("alone: bulk %.3f ms, chain of 15 %.3f ms (serial sum %.3f)\n", ta, tc, ta + tc);
    cudaEventDestroy(ec);
  }
  run("plain", def_a, def_b, false);
  run("streams", lo_a, hi_b, false);
  run("attribute", hi_a, hi_b, true);

  // the 'attribute' schedule as a graph captured from a high-priority stream
  cudaGraph_t graph;
  cudaGraphExec_t exec;
  cudaStreamBeginCapture(hi_cap, cudaStreamCaptureModeThreadLocal);
  cudaEventRecord(fork, hi_cap);
  cudaStreamWaitEvent(hi_b, fork, 0);
  launch(hi_cap, 15000, bulk_cycles, true, low);
  for (int k = 0; k < 15; ++k) launch(hi_b, 326, small_cycles, false, low);
  cudaEventRecord(join, hi_b);
  cudaStreamWaitEvent(hi_cap, join, 0);
  cudaStreamEndCapture(hi_cap, &graph);
  for (int flag = 0; flag < 2; ++flag) {
    // cudaGraphInstantiateFlagUseNodePriority: run the nodes at THEIR priorities, not at the launch stream's
    cudaGraphInstantiateWithFlags(&exec, graph, flag ? cudaGraphInstantiateFlagUseNodePriority : 0);
    for (int rep = 0; rep < 3; ++rep) {
      cudaDeviceSynchronize();
      cudaEventRecord(e0, hi_a);
      cudaGraphLaunch(exec, hi_a);
      cudaEventRecord(ea, hi_a);
      cudaDeviceSynchronize();
      float ta = 0;
      cudaEventElapsedTime(&ta, e0, ea);
      if (rep == 2)
        printf("graph %-22s bulk + chain done at %.3f ms (serial = the 'alone' sum)\n",
               flag ? "(UseNodePriority)" : "(default instantiate)", ta);
    }
  }
  {  // the priorities the captured nodes carry
    size_t n = 0;
    cudaGraphGetNodes(graph, nullptr, &n);
    cudaGraphNode_t nodes[64];
    n = n > 64 ? 64 : n;
    cudaGraphGetNodes(graph, nodes, &n);
    printf("node priorities:");
    for (size_t i = 0; i < n; ++i) {
      cudaGraphNodeType ty;
      cudaGraphNodeGetType(nodes[i], &ty);
      if (ty != cudaGraphNodeTypeKernel) continue;
      cudaLaunchAttributeValue v;
      cudaGraphKernelNodeGetAttribute(nodes[i], cudaLaunchAttributePriority, &v);
      printf(" %d", v.priority);
    }
    printf("\n");
  }
  // and the same graph without any priority
  cudaStreamBeginCapture(def_a, cudaStreamCaptureModeThreadLocal);
  cudaEventRecord(fork, def_a);
  cudaStreamWaitEvent(def_b, fork, 0);
  launch(def_a, 15000, bulk_cycles, false, low);
  for (int k = 0; k < 15; ++k) launch(def_b, 326, small_cycles, false, low);
  cudaEventRecord(join, def_b);
  cudaStreamWaitEvent(def_a, join, 0);
  cudaStreamEndCapture(def_a, &graph);
  cudaGraphInstantiate(&exec, graph, 0);
  for (int rep = 0; rep < 3; ++rep) {
    cudaDeviceSynchronize();
    cudaEventRecord(e0, def_a);
    cudaGraphLaunch(exec, def_a);
    cudaEventRecord(ea, def_a);
    cudaDeviceSynchronize();
    float ta = 0;
    cudaEventElapsedTime(&ta, e0, ea);
    if (rep == 2) printf("graph-plain bulk + chain done at %.3f ms\n", ta);
  }
  printf("last error: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
