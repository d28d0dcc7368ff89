// CPU check of grok_b200/csrc/host_pack.cpp (container conversion + fork-join pool); built and run by tests/test_host.py
#include "../grok_b200/csrc/b2k_internal.h"
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <vector>

static uint32_t rng = 12345;
static uint32_t next() { rng = rng * 1664525u + 1013904223u; return rng >> 8; }

int main()
{
  int bad = 0;
  const int thread_counts[] = {1, 3, 8};
  for(int tc : thread_counts)
  {
    b2k_host_set_threads(tc);
    for(int sgnd = 0; sgnd < 2; ++sgnd)
      for(int trial = 0; trial < 6; ++trial)
      {
        const size_t w = 1 + next() % 700, h = 1 + next() % 90, ss = w + next() % 9, ds = w + next() % 5, off = next() % 7;
        std::vector<int32_t> a(ss * h + 16), back(ss * h + 16, -7);
        std::vector<uint16_t> n(ds * h + 16 + off, 0xABCD);
        for(auto& v : a)
          v = sgnd ? (int32_t)(next() % 65536) - 32768 : (int32_t)(next() % 65536);
        b2k_host_rect r1{a.data(), n.data() + off, ss, ds, w, h};
        b2k_host_rect two[2] = {r1, r1}; // same rect twice: exercises multi-rect task indexing (idempotent)
        b2k_host_convert(two, 2, false, sgnd != 0);
        for(size_t y = 0; y < h; ++y)
          for(size_t x = 0; x < w; ++x)
            if(n[off + y * ds + x] != (uint16_t)a[y * ss + x]) ++bad;
        for(size_t y = 0; y < h; ++y)
          for(size_t x = w; x < ds; ++x)
            if(n[off + y * ds + x] != 0xABCD) ++bad; // nothing written past the row
        b2k_host_rect r2{n.data() + off, back.data(), ds, ss, w, h};
        b2k_host_convert(&r2, 1, true, sgnd != 0);
        for(size_t y = 0; y < h; ++y)
        {
          for(size_t x = 0; x < w; ++x)
            if(back[y * ss + x] != a[y * ss + x]) ++bad;
          for(size_t x = w; x < ss && y * ss + x < back.size(); ++x)
            if(back[y * ss + x] != -7) ++bad;
        }
      }
  }
  /* fork-join pool under churn: resized pools, spinning sessions on and off, empty and tiny loops -- every index
     must be visited exactly once */
  for(int round = 0; round < 24; ++round)
  {
    b2k_host_set_threads(1 + (int)(next() % 8));
    const bool hot = next() & 1;
    if(hot) b2k_host_session(true);
    for(int it = 0; it < 1000; ++it)
    {
      const size_t n = next() % 40;
      std::vector<std::atomic<int>> hits(n);
      for(auto& h : hits) h.store(0);
      b2k_host_parallel(n, [&](size_t i) { hits[i].fetch_add(1); });
      for(size_t i = 0; i < n; ++i)
        if(hits[i].load() != 1) ++bad;
    }
    if(hot) b2k_host_session(false);
  }
  b2k_host_set_threads(0);
  if(b2k_host_threads() != 0) ++bad;
  printf("host_pack_check bad=%d\n", bad);
  return bad ? 1 : 0;
}
