// Probe: register <-> (row, column) mapping of tcgen05.ld.16x256b.x2 against the known 32x32b layout (lane = row, register = column).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void probe(int* out) {
  __shared__ uint32_t slot;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)), "r"(32));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot;
  const uint32_t taddr = base + ((uint32_t)(warp * 32) << 16);
  uint32_t v[16];
  for (int c = 0; c < 16; c++) v[c] = (warp * 32 + lane) * 100 + c;     // value = row*100 + col
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
               "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]),
               "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  uint32_t r[16];
  for (int h = 0; h < 2; h++) {
    const uint32_t ta = taddr + ((uint32_t)(h * 16) << 16);
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[8 * h + 0]), "=r"(r[8 * h + 1]), "=r"(r[8 * h + 2]), "=r"(r[8 * h + 3]), "=r"(r[8 * h + 4]), "=r"(r[8 * h + 5]), "=r"(r[8 * h + 6]), "=r"(r[8 * h + 7])
                 : "r"(ta));
  }
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int i = 0; i < 16; i++) out[(warp * 32 + lane) * 16 + i] = (int)r[i];
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(32));
}
int main() {
  int* d; cudaMalloc(&d, 128 * 16 * 4);
  probe<<<1, 128>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status %s\n", cudaGetErrorString(e));
  static int h[128 * 16]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  for (int w = 0; w < 2; w++)
    for (int l = 0; l < 32; l++) {
      printf("w%d lane %2d:", w, l);
      for (int i = 0; i < 16; i++) printf(" (%d,%d)", h[(w * 32 + l) * 16 + i] / 100 - w * 32, h[(w * 32 + l) * 16 + i] % 100);
      printf("\n");
    }
  return 0;
}
