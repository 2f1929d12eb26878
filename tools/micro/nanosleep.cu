// How long does __nanosleep(t) really sleep on this GPU?  (tools/micro: measurement helpers, not product code)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(unsigned t, int iters, long long *out) {
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) __nanosleep(t);
  long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = (t1 - t0) / iters;
}
__global__ void kbar(unsigned hint_ns, int iters, long long *out) {
  __shared__ unsigned long long bar;
  if (threadIdx.x == 0) asm volatile("mbarrier.init.shared.b64 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(&bar)));
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
    unsigned ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared.b64 p, [%1], 0, %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"((unsigned)__cvta_generic_to_shared(&bar)), "r"(hint_ns));
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = (t1 - t0) / iters;
}
int main() {
  long long *d; cudaMalloc(&d, 8 * 1024);
  long long h[4];
  for (unsigned t : {64u, 250u, 1000u, 4000u, 16000u, 100000u}) {
    k<<<1, 32>>>(t, 200, d); cudaMemcpy(h, d, 8, cudaMemcpyDeviceToHost);
    long long a = h[0];
    k<<<148 * 4, 480>>>(t, 200, d); cudaMemcpy(h, d, 8, cudaMemcpyDeviceToHost);
    printf("nanosleep(%u): %lld cycles alone, %lld cycles with 148x4 CTAs of 480\n", t, a, h[0]);
  }
  for (unsigned t : {250u, 1000u, 4000u, 16000u}) {
    kbar<<<1, 32>>>(t, 200, d); cudaMemcpy(h, d, 8, cudaMemcpyDeviceToHost);
    printf("mbarrier.try_wait hint %u: %lld cycles\n", t, h[0]);
  }
  return 0;
}
