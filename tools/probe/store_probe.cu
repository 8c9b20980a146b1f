// store_probe.cu -- how fast can ONE SM hand 128 KiB to the memory system?  (developer probe, not product code)
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tools/probe/store_probe tools/probe/store_probe.cu
// (a) STG.128 from the registers of 1024 threads (what step_pair.cu's backward does): clocks until the warps are through
//     their stores; (b) the same bytes from shared memory with cp.async.bulk (TMA): clocks to issue, clocks until the source
//     has been read, clocks until complete.  With 1 CTA (nothing else on the chip) and with 148.
#include <cstdio>
#include <cstdint>
#include <vector>
#include <algorithm>

__global__ void __launch_bounds__(1024, 1) stg_kernel(float4* out, long long* clocks, int reps) {
  const size_t per_cta = 8 * 1024;       // float4s per repetition: 128 KiB
  float4 v = make_float4(threadIdx.x, 1.f, 2.f, 3.f);
  long long tot = 0;
  for (int r = 0; r < reps; ++r) {
    float4* dst = out + (static_cast<size_t>(r) * gridDim.x + blockIdx.x) * per_cta + threadIdx.x;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll
    for (int it = 0; it < 8; ++it) dst[it * 1024] = v;
    __syncthreads();
    tot += clock64() - t0;
    v.x += 1.f;
  }
  if (threadIdx.x == 0) clocks[blockIdx.x] = tot / reps;
}

__global__ void __launch_bounds__(1024, 1) tma_kernel(float4* out, long long* clocks, int reps) {
  extern __shared__ __align__(128) unsigned char smem[];
  const size_t per_cta = 8 * 1024;
  float4* s = reinterpret_cast<float4*>(smem);
  long long t_sts = 0, t_issue = 0, t_read = 0, t_done = 0;
  for (int r = 0; r < reps; ++r) {
    float4* dst = out + (static_cast<size_t>(r) * gridDim.x + blockIdx.x) * per_cta;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll
    for (int it = 0; it < 8; ++it) s[it * 1024 + threadIdx.x] = make_float4(threadIdx.x, r, it, 3.f);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) {
      const uint32_t sa = static_cast<uint32_t>(__cvta_generic_to_shared(smem));
      for (int c = 0; c < 4; ++c)
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + c * 2048), "r"(sa + c * 32768), "r"(32768) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    const long long t2 = clock64();
    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    __syncthreads();
    const long long t3 = clock64();
    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    __syncthreads();
    const long long t4 = clock64();
    t_sts += t1 - t0; t_issue += t2 - t1; t_read += t3 - t1; t_done += t4 - t1;
  }
  if (threadIdx.x == 0) {
    clocks[4 * blockIdx.x] = t_sts / reps; clocks[4 * blockIdx.x + 1] = t_issue / reps;
    clocks[4 * blockIdx.x + 2] = t_read / reps; clocks[4 * blockIdx.x + 3] = t_done / reps;
  }
}

static long long median(std::vector<long long> v) { std::sort(v.begin(), v.end()); return v[v.size() / 2]; }

int main() {
  const int reps = 40;
  float4* out; long long* d;
  cudaMalloc(&out, static_cast<size_t>(reps) * 148 * 131072);
  cudaMalloc(&d, 148 * 4 * sizeof(long long));
  cudaFuncSetAttribute(tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072);
  for (int ctas : {1, 2, 16, 74, 148}) {
    std::vector<long long> h(148 * 4);
    stg_kernel<<<ctas, 1024>>>(out, d, reps);
    stg_kernel<<<ctas, 1024>>>(out, d, reps);
    cudaDeviceSynchronize();
    cudaMemcpy(h.data(), d, ctas * sizeof(long long), cudaMemcpyDeviceToHost);
    const long long stg = median(std::vector<long long>(h.begin(), h.begin() + ctas));
    tma_kernel<<<ctas, 1024, 131072>>>(out, d, reps);
    tma_kernel<<<ctas, 1024, 131072>>>(out, d, reps);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h.data(), d, ctas * 4 * sizeof(long long), cudaMemcpyDeviceToHost);
    std::vector<long long> a, b, c, f;
    for (int i = 0; i < ctas; ++i) { a.push_back(h[4 * i]); b.push_back(h[4 * i + 1]); c.push_back(h[4 * i + 2]); f.push_back(h[4 * i + 3]); }
    printf("%3d CTAs x 128 KiB: STG.128 from registers %lld clk (%.1f B/clk/SM) | STS %lld clk, bulk store issue %lld, source read after %lld "
           "(%.1f B/clk/SM), complete after %lld clk  [%s]\n", ctas, stg, 131072.0 / stg, median(a), median(b), median(c),
           131072.0 / median(c), median(f), cudaGetErrorString(e));
  }
  return 0;
}
