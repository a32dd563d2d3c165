// What does mbarrier.pending_count report for the state returned by mbarrier.arrive?  (one block, 8 arrivals)
#include <cstdio>
#include <cstdint>
__global__ void k(unsigned* out) {
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&bar)), "r"(8));
  }
  __syncthreads();
  for (int round = 0; round < 2; ++round) {
    for (int w = 0; w < 8; ++w) {
      if ((threadIdx.x >> 5) == w && (threadIdx.x & 31) == 0) {
        uint64_t state; unsigned pending;
        asm volatile("mbarrier.arrive.shared::cta.b64 %0, [%1];" : "=l"(state) : "r"((uint32_t)__cvta_generic_to_shared(&bar)) : "memory");
        asm volatile("mbarrier.pending_count.b64 %0, %1;" : "=r"(pending) : "l"(state));
        out[round * 8 + w] = pending;
      }
      __syncthreads();
    }
  }
}
int main() {
  unsigned* d; cudaMalloc(&d, 64); cudaMemset(d, 0xff, 64);
  k<<<1, 256>>>(d);
  unsigned h[16]; cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
  for (int i = 0; i < 16; ++i) printf("%u ", h[i]);
  printf("\n%s\n", cudaGetErrorString(cudaGetLastError()));
}
