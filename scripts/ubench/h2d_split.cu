// Pinned host -> device bandwidth of one 170 MB buffer (the observation array of BASELINE configs[2]) sent as
// 1, 2, 4 or 8 concurrent chunks on separate streams:  nvcc -O3 -o h2d_split h2d_split.cu && ./h2d_split
#include <cstdio>
#include <cstring>
#include <cuda_runtime.h>
int main() {
  const size_t bytes = 170401416;
  unsigned char *h, *d;
  cudaHostAlloc((void**)&h, bytes, cudaHostAllocDefault);
  cudaMalloc((void**)&d, bytes);
  memset(h, 1, bytes);
  cudaStream_t st[8];
  for (auto& s : st) cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int n : {1, 2, 4, 8, 1}) {
    float best = 1e9f;
    for (int rep = 0; rep < 8; ++rep) {
      cudaDeviceSynchronize();
      cudaEventRecord(e0, st[0]);
      const size_t per = ((bytes + n - 1) / n + 255) & ~(size_t)255;
      for (int i = 1; i < n; ++i) cudaStreamWaitEvent(st[i], e0, 0);
      for (int i = 0; i < n; ++i) {
        const size_t off = i * per, len = off < bytes ? (bytes - off < per ? bytes - off : per) : 0;
        if (len) cudaMemcpyAsync(d + off, h + off, len, cudaMemcpyHostToDevice, st[i]);
      }
      cudaEvent_t done[8];
      for (int i = 1; i < n; ++i) { cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming); cudaEventRecord(done[i], st[i]); cudaStreamWaitEvent(st[0], done[i], 0); }
      cudaEventRecord(e1, st[0]);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      if (rep >= 2 && ms < best) best = ms;
      for (int i = 1; i < n; ++i) cudaEventDestroy(done[i]);
    }
    printf("%d stream(s): %.3f ms  %.1f GB/s\n", n, best, bytes / best * 1e-6);
  }
  return 0;
}
