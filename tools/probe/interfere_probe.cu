// interfere_probe.cu -- does a CTA that streams stores slow down its neighbour on the same SM?  (developer probe)
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tools/probe/interfere_probe tools/probe/interfere_probe.cu
// 2 CTAs of 512 threads per SM.  CTA A (blockIdx < 148): mode 0 idle, 1 = STG.128 storm from registers, 2 = STS + TMA bulk
// stores.  CTA B (blockIdx >= 148) measures the clocks of fixed pieces of work: (a) 64 KiB of LDS.128, (b) 32 MUFU per
// thread, (c) a bulk load of 8 KiB from global + wait, (d) 64 KiB of STG.128.
#include <cstdio>
#include <cstdint>
#include <vector>
#include <algorithm>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(bar), "r"(parity) : "memory");
}

__global__ void __launch_bounds__(512, 2) probe(float4* out, const float4* in, long long* res, int* smids, int mode, int reps, volatile int* stop) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) unsigned long long bar[2];
  float4* s = reinterpret_cast<float4*>(smem);
  const int tid = threadIdx.x;
  unsigned sm;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
  if (tid == 0) smids[blockIdx.x] = sm;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[0])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[1])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (blockIdx.x < 148) {      // ---- A: the disturber, runs until B says stop (bounded)
    float4* dst = out + static_cast<size_t>(blockIdx.x) * (1 << 18);      // 4 MiB window per CTA, cycled
    float4 v = make_float4(tid, 1.f, 2.f, 3.f);
    for (int r = 0; r < 4000 && !*stop; ++r) {
      float4* d = dst + (r & 63) * 4096;
      if (mode == 1) {
#pragma unroll
        for (int it = 0; it < 8; ++it) d[it * 512 + tid] = v;
      } else if (mode == 2) {
#pragma unroll
        for (int it = 0; it < 8; ++it) s[it * 512 + tid] = v;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(d), "r"(smem_u32(smem)), "r"(65536) : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        __syncthreads();
      } else {
        const long long t0 = clock64();
        while (clock64() - t0 < 2000) {}
      }
      v.x += 1.f;
    }
    return;
  }
  // ---- B: the measured neighbour
  const int b = blockIdx.x - 148;
  long long t_lds = 0, t_mufu = 0, t_tma = 0, t_stg = 0;
  float acc = 0.f;
  for (int it = 0; it < 8; ++it) s[it * 512 + tid] = make_float4(tid, it, 1.f, 2.f);
  __syncthreads();
  float4* dst = out + static_cast<size_t>(148 + b) * (1 << 18);
  for (int r = 0; r < reps; ++r) {
    __syncthreads();
    long long t0 = clock64();
#pragma unroll
    for (int it = 0; it < 8; ++it) { const float4 q = s[it * 512 + tid]; acc += q.x + q.y + q.z + q.w; }
    __syncthreads();
    long long t1 = clock64();
    t_lds += t1 - t0;
    float e = acc * 1e-9f;
#pragma unroll
    for (int i = 0; i < 32; ++i) e = exp2f(e) * 0.5f;
    acc += e;
    __syncthreads();
    long long t2 = clock64();
    t_mufu += t2 - t1;
    if (tid == 0) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[0])), "r"(8192) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem) + 65536),
                   "l"(in + (static_cast<size_t>(b) * reps + r) * 512), "r"(8192), "r"(smem_u32(&bar[0])) : "memory");
    }
    mbar_wait(smem_u32(&bar[0]), r & 1);
    __syncthreads();
    long long t3 = clock64();
    t_tma += t3 - t2;
    long long t4 = t3;
#pragma unroll
    for (int it = 0; it < 8; ++it) dst[(r & 63) * 4096 + it * 512 + tid] = make_float4(acc, 1.f, 2.f, 3.f);
    __syncthreads();
    long long t5 = clock64();
    t_stg += t5 - t4;
  }
  if (tid == 0) {
    long long* o = res + 5 * b;
    o[0] = t_lds / reps; o[1] = t_mufu / reps; o[2] = t_tma / reps; o[3] = 0; o[4] = t_stg / reps;
    if (acc == 12345.678f) o[0] = 0;
    if (b == 0) *stop = 1;
  }
}

int main() {
  const int reps = 200;
  float4 *out, *in; long long* d; int* sm; int* stop;
  cudaMalloc(&out, static_cast<size_t>(296) * (1 << 18) * 16);
  cudaMalloc(&in, static_cast<size_t>(148) * reps * 8192);
  cudaMalloc(&d, 148 * 5 * sizeof(long long));
  cudaMalloc(&sm, 296 * sizeof(int));
  cudaMalloc(&stop, sizeof(int));
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const char* names[3] = {"idle", "STG.128 storm", "STS + TMA bulk store storm"};
  for (int mode = 0; mode < 3; ++mode) {
    cudaMemset(stop, 0, sizeof(int));
    probe<<<296, 512, 100 * 1024>>>(out, in, d, sm, mode, reps, stop);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<long long> h(148 * 5);
    std::vector<int> s(296);
    cudaMemcpy(h.data(), d, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
    cudaMemcpy(s.data(), sm, s.size() * sizeof(int), cudaMemcpyDeviceToHost);
    int paired = 0;
    std::vector<int> cntA(148, 0);
    for (int i = 0; i < 148; ++i) cntA[s[i]]++;
    for (int i = 148; i < 296; ++i) paired += cntA[s[i]] == 1;
    long long m[5];
    for (int k = 0; k < 5; ++k) { std::vector<long long> v; for (int b = 0; b < 148; ++b) v.push_back(h[5 * b + k]); std::sort(v.begin(), v.end()); m[k] = v[74]; }
    printf("neighbour %-28s (%3d of 148 B-CTAs share an SM with exactly one A): 64 KiB LDS.128 %5lld clk | 32 MUFU/thread %5lld | 8 KiB bulk load %5lld | "
           "64 KiB STG.128 %5lld  [%s]\n", names[mode], paired, m[0], m[1], m[2], m[4], cudaGetErrorString(e));
  }
  return 0;
}
