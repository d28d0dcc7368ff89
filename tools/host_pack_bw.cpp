// microbench: host narrow/widen bandwidth vs threads (dev tool).  g++ -O2 -std=c++17 -I/usr/local/cuda/include tools/host_pack_bw.cpp grok_b200/csrc/host_pack.cpp -lpthread
#include "../grok_b200/csrc/b2k_internal.h"
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
int main(int argc, char** argv)
{
  const size_t W = 8192, H = 8192 * 3;
  int32_t* a = (int32_t*)aligned_alloc(4096, W * H * 4);
  uint16_t* n = (uint16_t*)aligned_alloc(4096, W * H * 2);
  for(size_t i = 0; i < W * H; ++i) a[i] = (int32_t)(i * 2654435761u >> 20);
  memset(n, 0, W * H * 2);
  for(int i = 1; i < argc; ++i)
  {
    const int nt = atoi(argv[i]);
    b2k_host_set_threads(nt);
    double best[2] = {1e9, 1e9};
    for(int it = 0; it < 5; ++it)
      for(int widen = 0; widen < 2; ++widen)
      {
        const auto t0 = std::chrono::steady_clock::now();
        for(int k = 0; k < 8; ++k) // chunked like the engine: 8 fork-joins of 3 rects
        {
          b2k_host_rect r[3];
          for(int c = 0; c < 3; ++c)
          {
            const size_t off = ((size_t)c * 8192 + (size_t)k * 1024) * W;
            if(widen) r[c] = {n + off, a + off, W, W, W, 1024};
            else r[c] = {a + off, n + off, W, W, W, 1024};
          }
          b2k_host_convert(r, 3, widen, false);
        }
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        if(ms < best[widen]) best[widen] = ms;
      }
    printf("threads %2d narrow %.2f ms (%.0f GB/s r+w)  widen %.2f ms (%.0f GB/s r+w)\n", nt, best[0], W * H * 6 / best[0] / 1e6, best[1],
           W * H * 6 / best[1] / 1e6);
  }
  return 0;
}
