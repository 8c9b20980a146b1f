// cluster_probe.cu -- where does the hardware put the CTAs of a thread-block cluster?  (developer probe, not product code)
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tools/probe/cluster_probe tools/probe/cluster_probe.cu
// For cluster sizes 2, 4, 8 and CTAs sized so that TWO fit on an SM (512 threads, ~104 KiB of shared memory): the number of
// co-resident clusters, the SMs in use, and how many distinct SMs one cluster spans.
#include <cooperative_groups.h>
#include <cstdio>
#include <set>
#include <vector>
#include <map>
namespace cg = cooperative_groups;

__global__ void probe(int* smid_of, long long spin) {
  extern __shared__ char sm[];
  unsigned s;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(s));
  if (threadIdx.x == 0) smid_of[blockIdx.x] = static_cast<int>(s);
  cg::this_cluster().sync();
  const long long t0 = clock64();
  while (clock64() - t0 < spin) {}      // keep every CTA resident until all are placed
  if (threadIdx.x == 0) sm[0] = 1;
}

int main() {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  printf("SMs %d\n", sms);
  cudaFuncSetAttribute(probe, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  const int cfgs[][3] = {{2, 1024, 224 * 1024}, {2, 512, 104 * 1024}, {4, 512, 104 * 1024}, {8, 512, 104 * 1024},
                         {8, 256, 52 * 1024}, {4, 1024, 224 * 1024}};
  for (auto& c : cfgs) {
    const int cs = c[0], nt = c[1], smem = c[2];
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.gridDim = dim3(cs * 2 * sms); cfg.blockDim = dim3(nt); cfg.dynamicSmemBytes = smem; cfg.attrs = at; cfg.numAttrs = 1;
    int nc = 0;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&nc, probe, &cfg);
    printf("cluster %d x %4d threads, %3d KiB: max active clusters %d (%s)\n", cs, nt, smem / 1024, nc, cudaGetErrorString(e));
    if (e != cudaSuccess || nc <= 0) { cudaGetLastError(); continue; }
    const int ctas = cs * nc;
    int* d; cudaMalloc(&d, ctas * sizeof(int));
    cfg.gridDim = dim3(ctas);
    e = cudaLaunchKernelEx(&cfg, probe, d, 2000000LL);
    cudaError_t e2 = cudaDeviceSynchronize();
    if (e != cudaSuccess || e2 != cudaSuccess) { printf("  launch failed: %s / %s\n", cudaGetErrorString(e), cudaGetErrorString(e2)); cudaGetLastError(); continue; }
    std::vector<int> h(ctas);
    cudaMemcpy(h.data(), d, ctas * sizeof(int), cudaMemcpyDeviceToHost);
    std::set<int> used;
    std::map<int, int> per_sm, span_hist;
    for (int i = 0; i < ctas; ++i) { used.insert(h[i]); per_sm[h[i]]++; }
    for (int k = 0; k < nc; ++k) {
      std::set<int> s;
      for (int r = 0; r < cs; ++r) s.insert(h[k * cs + r]);
      span_hist[static_cast<int>(s.size())]++;
    }
    std::map<int, int> occ_hist;
    for (auto& kv : per_sm) occ_hist[kv.second]++;
    printf("  SMs used %zu;", used.size());
    for (auto& kv : occ_hist) printf(" %d SMs with %d CTAs;", kv.second, kv.first);
    printf("  distinct SMs per cluster:");
    for (auto& kv : span_hist) printf(" %d clusters span %d SMs;", kv.second, kv.first);
    printf("\n  first clusters:");
    for (int k = 0; k < 3 && k < nc; ++k) { printf(" ["); for (int r = 0; r < cs; ++r) printf("%d ", h[k * cs + r]); printf("]"); }
    printf("\n");
    cudaFree(d);
  }
  return 0;
}
